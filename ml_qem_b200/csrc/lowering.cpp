// K0 -- circuit lowering stage (host C++).
//
// Turns the flat gate stream of one circuit (bwq_batch) plus the per-backend noise table into a
// "sweep program" for the density-matrix kernels:
//   gate + attached error -> real Pauli-transfer matrices (error applied AFTER its gate, the
//   density_matrix-method semantics of Aer reached from blackwater/data/utils.py:427-429),
//   consecutive 1-qubit maps on a qubit are multiplied into one 4x4 and absorbed into the
//   neighbouring 2-qubit register pass, qubits no instruction touches are truncated (Aer does
//   the same), and passes are packed greedily into shared-memory tile sweeps so one HBM
//   read+write of the state serves as many gates as the tile's qubit set allows.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>

#include "program.h"

namespace bwq {

using cd = std::complex<double>;
static const cd I_(0.0, 1.0);

// ----------------------------------------------------------------------------- noise table
int NoiseTable::set(const bwq_noise_table* t, char* err, size_t errlen) {
  data.clear();
  entries.clear();
  one_q.clear();
  fixed_ptm.clear();
  fixed_ok.clear();
  if (!t || t->n_entries == 0) return BWQ_OK;
  if (t->n_entries < 0 || !t->opcode || !t->q0 || !t->q1 || !t->kind || !t->data_off || !t->data) {
    snprintf(err, errlen, "noise table: null array");
    return BWQ_ERR_ARG;
  }
  // entries are re-packed 4-double aligned so the kernels may use 16-byte loads
  for (int i = 0; i < t->n_entries; ++i) {
    int need = t->kind[i] == BWQ_NOISE_DENSE1 ? 16 : t->kind[i] == BWQ_NOISE_DENSE2 ? 256
             : t->kind[i] == BWQ_NOISE_RELAX2 ? 25 : -1;
    if (need < 0 || t->data_off[i] < 0 || t->data_off[i] + need > t->n_data) {
      snprintf(err, errlen, "noise table: entry %d has bad kind/offset", i);
      return BWQ_ERR_ARG;
    }
    bool two = gate_is_2q(t->opcode[i]);
    if (two != (t->kind[i] != BWQ_NOISE_DENSE1)) {
      snprintf(err, errlen, "noise table: entry %d kind does not match gate arity", i);
      return BWQ_ERR_ARG;
    }
    uint32_t key = (uint32_t(t->opcode[i]) << 16) | (uint32_t(t->q0[i]) << 8) | t->q1[i];
    while (data.size() % 4) data.push_back(0.0);
    entries.push_back({key, NoiseEntry{t->kind[i], (int64_t)data.size()}});
    data.insert(data.end(), t->data + t->data_off[i], t->data + t->data_off[i] + need);
  }
  while (data.size() % 4) data.push_back(0.0);  // the kernels read a 25-double entry as 13 pairs
  data.push_back(0.0);
  std::sort(entries.begin(), entries.end(),
            [](const auto& a, const auto& b) { return a.first < b.first; });
  one_q.clear();  // find() below searches while the table is being filled
  std::vector<int32_t> tab(32 * 64, -1);
  for (int op = 0; op < 32; ++op)
    for (int q = 0; q < 64; ++q) {
      const NoiseEntry* e = find((uint16_t)op, q, 255);
      if (!e) continue;
      for (size_t i = 0; i < entries.size(); ++i)
        if (&entries[i].second == e) { tab[op * 64 + q] = (int32_t)i; break; }
    }
  one_q.swap(tab);
  build_fixed();
  return BWQ_OK;
}

const NoiseEntry* NoiseTable::find(uint16_t opcode, int q0, int q1) const {
  if (entries.empty()) return nullptr;
  if (q1 == 255 && opcode < 32 && q0 >= 0 && q0 < 64 && !one_q.empty()) {
    const int32_t i = one_q[opcode * 64 + q0];
    return i < 0 ? nullptr : &entries[(size_t)i].second;
  }
  auto look = [&](uint32_t key) -> const NoiseEntry* {
    auto it = std::lower_bound(entries.begin(), entries.end(), key,
                               [](const auto& a, uint32_t k) { return a.first < k; });
    return (it != entries.end() && it->first == key) ? &it->second : nullptr;
  };
  // a local error for the exact ordered qubit tuple overrides the all-qubit default (Aer)
  if (const NoiseEntry* e = look((uint32_t(opcode) << 16) | (uint32_t(q0) << 8) | uint32_t(q1 & 255)))
    return e;
  return look((uint32_t(opcode) << 16) | (255u << 8) | 255u);
}

// ----------------------------------------------------------------------------- gate library
bool gate_is_2q(uint16_t op) { return (op >= BWQ_G_CX && op <= BWQ_G_ECR) || op == BWQ_G_UNITARY2; }

int gate_num_params(uint16_t op) {
  switch (op) {
    case BWQ_G_RX: case BWQ_G_RY: case BWQ_G_RZ: case BWQ_G_P:
    case BWQ_G_CRX: case BWQ_G_CRY: case BWQ_G_CRZ: case BWQ_G_CP:
    case BWQ_G_RZZ: case BWQ_G_RXX: case BWQ_G_RYY: case BWQ_G_RZX: return 1;
    case BWQ_G_U2: return 2;
    case BWQ_G_U3: case BWQ_G_CU3: return 3;
    case BWQ_G_UNITARY1: return 8;
    case BWQ_G_UNITARY2: return 32;
    default: return 0;
  }
}

static void u3_mat(double th, double ph, double la, cd* m) {
  double c = std::cos(th / 2), s = std::sin(th / 2);
  m[0] = c;
  m[1] = -std::exp(I_ * la) * s;
  m[2] = std::exp(I_ * ph) * s;
  m[3] = std::exp(I_ * (ph + la)) * c;
}

static bool unitary1(uint16_t op, const double* p, cd* m) {
  const double r = std::sqrt(0.5);
  switch (op) {
    case BWQ_G_ID: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = 1; return true;
    case BWQ_G_X: m[0] = 0; m[1] = 1; m[2] = 1; m[3] = 0; return true;
    case BWQ_G_Y: m[0] = 0; m[1] = -I_; m[2] = I_; m[3] = 0; return true;
    case BWQ_G_Z: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = -1; return true;
    case BWQ_G_H: m[0] = r; m[1] = r; m[2] = r; m[3] = -r; return true;
    case BWQ_G_S: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = I_; return true;
    case BWQ_G_SDG: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = -I_; return true;
    case BWQ_G_T: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = std::exp(I_ * (M_PI / 4)); return true;
    case BWQ_G_TDG: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = std::exp(-I_ * (M_PI / 4)); return true;
    case BWQ_G_SX: m[0] = cd(.5, .5); m[1] = cd(.5, -.5); m[2] = cd(.5, -.5); m[3] = cd(.5, .5); return true;
    case BWQ_G_SXDG: m[0] = cd(.5, -.5); m[1] = cd(.5, .5); m[2] = cd(.5, .5); m[3] = cd(.5, -.5); return true;
    case BWQ_G_RX: { double c = std::cos(p[0] / 2), s = std::sin(p[0] / 2);
      m[0] = c; m[1] = -I_ * s; m[2] = -I_ * s; m[3] = c; return true; }
    case BWQ_G_RY: { double c = std::cos(p[0] / 2), s = std::sin(p[0] / 2);
      m[0] = c; m[1] = -s; m[2] = s; m[3] = c; return true; }
    // exp(+-i t) written as (cos t, +-sin t): what std::exp(complex) returns, without its overhead
    case BWQ_G_RZ: { const double c = std::cos(p[0] / 2), sn = std::sin(p[0] / 2);
      m[0] = cd(c, -sn); m[1] = 0; m[2] = 0; m[3] = cd(c, sn); return true; }
    case BWQ_G_P: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = cd(std::cos(p[0]), std::sin(p[0])); return true;
    case BWQ_G_U2: u3_mat(M_PI / 2, p[0], p[1], m); return true;
    case BWQ_G_U3: u3_mat(p[0], p[1], p[2], m); return true;
    case BWQ_G_UNITARY1: for (int i = 0; i < 4; ++i) m[i] = cd(p[2 * i], p[2 * i + 1]); return true;
    default: return false;
  }
}

// local index = i_q0 + 2*i_q1; controlled gates: control = q0, target = q1
static void controlled(const cd* u, cd* m) {
  for (int i = 0; i < 16; ++i) m[i] = 0;
  m[0 * 4 + 0] = 1;
  m[2 * 4 + 2] = 1;
  for (int tr = 0; tr < 2; ++tr)
    for (int tc = 0; tc < 2; ++tc) m[(1 + 2 * tr) * 4 + (1 + 2 * tc)] = u[tr * 2 + tc];
}

// exp(-i th/2 * A(x)B) with A on q0, B on q1 (A, B Paulis: 1=X 2=Y 3=Z)
static void pauli_rot2(int a, int b, double th, cd* m) {
  static const cd P[4][4] = {{1, 0, 0, 1}, {0, 1, 1, 0}, {0, cd(0, -1), cd(0, 1), 0}, {1, 0, 0, -1}};
  double c = std::cos(th / 2), s = std::sin(th / 2);
  for (int r0 = 0; r0 < 2; ++r0) for (int r1 = 0; r1 < 2; ++r1)
    for (int c0 = 0; c0 < 2; ++c0) for (int c1 = 0; c1 < 2; ++c1) {
      cd pp = P[a][r0 * 2 + c0] * P[b][r1 * 2 + c1];
      cd id = (r0 == c0 && r1 == c1) ? 1.0 : 0.0;
      m[(r0 + 2 * r1) * 4 + (c0 + 2 * c1)] = c * id - I_ * s * pp;
    }
}

static bool unitary2(uint16_t op, const double* p, cd* m) {
  cd u[4];
  switch (op) {
    case BWQ_G_CX: unitary1(BWQ_G_X, p, u); controlled(u, m); return true;
    case BWQ_G_CY: unitary1(BWQ_G_Y, p, u); controlled(u, m); return true;
    case BWQ_G_CZ: unitary1(BWQ_G_Z, p, u); controlled(u, m); return true;
    case BWQ_G_CH: unitary1(BWQ_G_H, p, u); controlled(u, m); return true;
    case BWQ_G_CRX: unitary1(BWQ_G_RX, p, u); controlled(u, m); return true;
    case BWQ_G_CRY: unitary1(BWQ_G_RY, p, u); controlled(u, m); return true;
    case BWQ_G_CRZ: unitary1(BWQ_G_RZ, p, u); controlled(u, m); return true;
    case BWQ_G_CP: unitary1(BWQ_G_P, p, u); controlled(u, m); return true;
    case BWQ_G_CU3: unitary1(BWQ_G_U3, p, u); controlled(u, m); return true;
    case BWQ_G_SWAP:
      for (int i = 0; i < 16; ++i) m[i] = 0;
      m[0] = 1; m[1 * 4 + 2] = 1; m[2 * 4 + 1] = 1; m[15] = 1; return true;
    case BWQ_G_ISWAP:
      for (int i = 0; i < 16; ++i) m[i] = 0;
      m[0] = 1; m[1 * 4 + 2] = I_; m[2 * 4 + 1] = I_; m[15] = 1; return true;
    case BWQ_G_RZZ: pauli_rot2(3, 3, p[0], m); return true;
    case BWQ_G_RXX: pauli_rot2(1, 1, p[0], m); return true;
    case BWQ_G_RYY: pauli_rot2(2, 2, p[0], m); return true;
    case BWQ_G_RZX: pauli_rot2(3, 1, p[0], m); return true;  // Z on q0, X on q1 (qiskit RZXGate)
    case BWQ_G_ECR: {
      const double r = std::sqrt(0.5);
      const cd e[16] = {0, 1, 0, I_, 1, 0, -I_, 0, 0, I_, 0, 1, -I_, 0, 1, 0};
      for (int i = 0; i < 16; ++i) m[i] = r * e[i];
      return true; }
    case BWQ_G_UNITARY2: for (int i = 0; i < 16; ++i) m[i] = cd(p[2 * i], p[2 * i + 1]); return true;
    default: return false;
  }
}

bool gate_unitary(uint16_t op, const double* p, double* out) {
  cd m[16];
  if (gate_is_2q(op)) {
    if (!unitary2(op, p, m)) return false;
    for (int i = 0; i < 16; ++i) { out[2 * i] = m[i].real(); out[2 * i + 1] = m[i].imag(); }
    return true;
  }
  if (!unitary1(op, p, m)) return false;
  for (int i = 0; i < 4; ++i) { out[2 * i] = m[i].real(); out[2 * i + 1] = m[i].imag(); }
  return true;
}

static const cd kPauli[4][4] = {{1, 0, 0, 1}, {0, 1, 1, 0}, {0, cd(0, -1), cd(0, 1), 0}, {1, 0, 0, -1}};

// R[i][j] = 1/2 Tr(P_i U P_j U^dag)
void ptm_from_unitary1(const double* uu, double* r) {
  cd u[4];
  for (int i = 0; i < 4; ++i) u[i] = cd(uu[2 * i], uu[2 * i + 1]);
  for (int j = 0; j < 4; ++j) {
    cd t[4], e[4];  // t = U P_j, e = t U^dag
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b)
      t[a * 2 + b] = u[a * 2 + 0] * kPauli[j][0 * 2 + b] + u[a * 2 + 1] * kPauli[j][1 * 2 + b];
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b)
      e[a * 2 + b] = t[a * 2 + 0] * std::conj(u[b * 2 + 0]) + t[a * 2 + 1] * std::conj(u[b * 2 + 1]);
    for (int i = 0; i < 4; ++i) {
      cd tr = 0;
      for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) tr += kPauli[i][a * 2 + b] * e[b * 2 + a];
      r[i * 4 + j] = 0.5 * tr.real();
    }
  }
}

// 16x16: index = digit_q0 + 4*digit_q1, basis index = b_q0 + 2*b_q1
void ptm_from_unitary2(const double* uu, double* r) {
  cd u[16], P[16][16];
  for (int i = 0; i < 16; ++i) u[i] = cd(uu[2 * i], uu[2 * i + 1]);
  for (int d0 = 0; d0 < 4; ++d0) for (int d1 = 0; d1 < 4; ++d1)
    for (int r0 = 0; r0 < 2; ++r0) for (int r1 = 0; r1 < 2; ++r1)
      for (int c0 = 0; c0 < 2; ++c0) for (int c1 = 0; c1 < 2; ++c1)
        P[d0 + 4 * d1][(r0 + 2 * r1) * 4 + (c0 + 2 * c1)] = kPauli[d0][r0 * 2 + c0] * kPauli[d1][r1 * 2 + c1];
  for (int j = 0; j < 16; ++j) {
    cd t[16], e[16];
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) {
      cd s = 0;
      for (int k = 0; k < 4; ++k) s += u[a * 4 + k] * P[j][k * 4 + b];
      t[a * 4 + b] = s;
    }
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) {
      cd s = 0;
      for (int k = 0; k < 4; ++k) s += t[a * 4 + k] * std::conj(u[b * 4 + k]);
      e[a * 4 + b] = s;
    }
    for (int i = 0; i < 16; ++i) {
      cd tr = 0;
      for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) tr += P[i][a * 4 + b] * e[b * 4 + a];
      r[i * 16 + j] = 0.25 * tr.real();
    }
  }
}

// ----------------------------------------------------------------------------- 4x4 helpers
static void mat4_identity(double* m) { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
// out = a*b (out may alias a or b).  Row i of the product is a linear combination of the rows of
// b, evaluated 4 wide in the order ((a0 b0 + a1 b1) + a2 b2) + a3 b3 -- the summation order of
// the scalar triple loop, so the programs stay bit-identical.  K0 spends most of its time here
// (two products per 1-qubit gate).
typedef double v4d __attribute__((vector_size(32), aligned(8)));
static inline void mat4_mul(const double* a, const double* b, double* out) {
  v4d b0, b1, b2, b3, r[4];
  std::memcpy(&b0, b, 32); std::memcpy(&b1, b + 4, 32); std::memcpy(&b2, b + 8, 32); std::memcpy(&b3, b + 12, 32);
  for (int i = 0; i < 4; ++i) r[i] = ((a[4 * i] * b0 + a[4 * i + 1] * b1) + a[4 * i + 2] * b2) + a[4 * i + 3] * b3;
  std::memcpy(out, r, sizeof r);
}
// out = R * b for R = identity outside rows/columns 1, 2 (rz, p, z, s, t: a rotation of X, Y)
static inline void mat4_rot_mul(double c, double sn, const double* b, double* out) {
  v4d b1, b2;
  std::memcpy(&b1, b + 4, 32); std::memcpy(&b2, b + 8, 32);
  const v4d r1 = c * b1 + (-sn) * b2, r2 = sn * b1 + c * b2;
  if (out != b) { std::memcpy(out, b, 32); std::memcpy(out + 12, b + 12, 32); }
  std::memcpy(out + 4, &r1, 32); std::memcpy(out + 8, &r2, 32);
}
static bool gate_ptm1(uint16_t op, const double* p, double* r) {
  switch (op) {
    case BWQ_G_ID: mat4_identity(r); return true;
    case BWQ_G_X: mat4_identity(r); r[10] = -1; r[15] = -1; return true;
    case BWQ_G_Z: mat4_identity(r); r[5] = -1; r[10] = -1; return true;
    case BWQ_G_RZ: case BWQ_G_P: {
      double c = std::cos(p[0]), s = std::sin(p[0]);
      mat4_identity(r); r[5] = c; r[6] = -s; r[9] = s; r[10] = c; return true; }
    case BWQ_G_SX:  // rx(pi/2): Y -> Z, Z -> -Y
      for (int i = 0; i < 16; ++i) r[i] = 0;
      r[0] = 1; r[5] = 1; r[2 * 4 + 3] = -1; r[3 * 4 + 2] = 1; return true;
    case BWQ_G_RESET:  // rho -> Tr(rho)|0><0|
      for (int i = 0; i < 16; ++i) r[i] = 0;
      r[0] = 1; r[3 * 4 + 0] = 1; return true;
    default: {
      double u[8];
      if (!gate_unitary(op, p, u)) return false;
      ptm_from_unitary1(u, r);
      return true; }
  }
}

void NoiseTable::build_fixed() {
  fixed_ptm.assign((size_t)32 * 64 * 16, 0.0);
  fixed_ok.assign((size_t)32 * 64, 0);
  for (int op = 0; op < 32; ++op) {
    if (gate_num_params((uint16_t)op) != 0 || op == BWQ_G_RESET) continue;
    for (int q = 0; q < 64; ++q) {
      const NoiseEntry* ne = find((uint16_t)op, q, 255);
      double r[16];
      if (!ne || !gate_ptm1((uint16_t)op, nullptr, r)) continue;
      mat4_mul(&data[ne->off], r, &fixed_ptm[(size_t)(op * 64 + q) * 16]);  // same product as the gate loop
      fixed_ok[op * 64 + q] = 1;
    }
  }
}

// ----------------------------------------------------------------------------- DM lowering
namespace {
struct HostPass {
  int qa, qb;
  std::vector<MacroOp> ops;
};
// Result of the first lowering stage (gates + noise -> register passes), reusable for every ZNE
// fold of the circuit: a fold only repeats the cx ops inside the passes.
struct DmStageA {
  bool valid = false;
  bool cx_only = true;          // every 2-qubit gate is a cx (self-inverse: folding = repetition)
  CircuitProgram head;          // active qubits, term indices, matrix scratch, status
  std::vector<HostPass> passes;
};
}  // namespace

static int64_t push_mat(std::vector<double>& mats, const double* m, int n) {
  int64_t off = (int64_t)mats.size();
  mats.insert(mats.end(), m, m + n);
  return off;
}

static void lower_terms(const bwq_batch& b, int c, const std::vector<int>& digit_of, int nq,
                        CircuitProgram* out) {
  int64_t o0 = b.obs_offsets[c], o1 = b.obs_offsets[c + 1];
  int64_t t0 = b.term_offsets[o0], t1 = b.term_offsets[o1];
  out->term_index.resize(t1 - t0);
  out->term_coeff.resize(t1 - t0);
  uint64_t valid = nq >= 64 ? ~0ull : ((1ull << nq) - 1);
  for (int64_t t = t0; t < t1; ++t) {
    uint64_t x = b.term_x[t], z = b.term_z[t];
    int64_t idx = 0;
    if ((x | z) & ~valid) { out->status = BWQ_CIRC_BAD_QUBIT; idx = -1; }
    for (int q = 0; q < nq && idx >= 0; ++q) {
      int xb = (x >> q) & 1, zb = (z >> q) & 1;
      if (!xb && !zb) continue;
      if (digit_of[q] < 0) {           // idle qubit stays |0>: <Z>=1, <X>=<Y>=0
        if (xb) idx = -1;
        continue;
      }
      int digit = xb ? (zb ? 2 : 1) : 3;
      idx += (int64_t)digit << (2 * digit_of[q]);
    }
    out->term_index[t - t0] = idx;
    out->term_coeff[t - t0] = b.term_coeff[t];
  }
}

// op list of a pass -> PassSig (straight-line kernel body) or SIG_GENERIC.  Only passes that do not
// target slot 0 have fast bodies (their two register groups share every 16-byte access).
static PassSig classify_pass(const std::vector<MacroOp>& ops, bool fast_layout) {
  if (!fast_layout || ops.empty() || ops.size() > 2) return SIG_GENERIC;
  auto is = [](const MacroOp& o, uint8_t a, uint8_t b, uint8_t q) { return o.pre_a == a && o.pre_b == b && o.twoq == q; };
  const MacroOp& o0 = ops[0];
  if (ops.size() == 1) {
    if (is(o0, P_AFF, P_AFF, Q_CXN_AB)) return SIG_AAC;
    if (is(o0, P_ROT, P_AFF, Q_CXN_AB)) return SIG_RAC;
    if (is(o0, P_NONE, P_AFF, Q_CXN_AB)) return SIG_0AC;
    return SIG_GENERIC;
  }
  const MacroOp& o1 = ops[1];
  if (is(o0, P_AFF, P_AFF, Q_CXN_AB)) {
    if (is(o1, P_AFF, P_AFF, Q_NONE)) return SIG_AAC_AA;
    if (is(o1, P_NONE, P_ROT, Q_CXN_AB)) return SIG_AAC_RC;
    if (is(o1, P_AFF, P_NONE, Q_NONE)) return SIG_AAC_A0;
    if (is(o1, P_NONE, P_AFF, Q_NONE)) return SIG_AAC_0A;
  }
  if (is(o0, P_NONE, P_NONE, Q_CXN_AB) && is(o1, P_NONE, P_ROT, Q_CXN_AB)) return SIG_C_RC;
  return SIG_GENERIC;
}

static void lower_dm_impl(const NoiseTable& noise, const bwq_batch& b, int c, const LowerOptions& opt, CircuitProgram* out,
                          DmStageA* cache, int fold);
static void lower_dm_gates(const NoiseTable& noise, const bwq_batch& b, int c, CircuitProgram* out, std::vector<HostPass>* passes_out);
static void pack_dm_passes(const NoiseTable& noise, const LowerOptions& opt, const int nd, std::vector<HostPass>& passes, CircuitProgram* out);

void lower_dm_circuit(const NoiseTable& noise, const bwq_batch& b, int c, const LowerOptions& opt,
                      CircuitProgram* out) {
  lower_dm_impl(noise, b, c, opt, out, nullptr, 1);
}

// Programs of circuit c at the noise factors folds[0..n) (local folding of the 2-qubit gates, as
// LocalFoldingAmplifier(gates_to_fold=2) of docs/tutorials/zne_parallel.py:172-183): the gates are
// lowered ONCE, every fold re-packs the passes with its cx ops repeated.  Returns false when the
// circuit has 2-qubit gates other than cx (their folds need inverse gates: lower the expanded
// gate stream of bwq_variants instead).
bool lower_dm_circuit_folds(const NoiseTable& noise, const bwq_batch& b, int c, const LowerOptions& opt, const int32_t* folds,
                            int n_folds, CircuitProgram* outs) {
  for (int64_t g = b.op_offsets[c]; g < b.op_offsets[c + 1]; ++g)
    if (gate_is_2q(b.ops[g].opcode) && b.ops[g].opcode != BWQ_G_CX) return false;
  DmStageA A;
  for (int f = 0; f < n_folds; ++f) lower_dm_impl(noise, b, c, opt, &outs[f], &A, folds[f]);
  return true;
}

static void lower_dm_impl(const NoiseTable& noise, const bwq_batch& b, int c, const LowerOptions& opt, CircuitProgram* out,
                          DmStageA* cache, int fold) {
  *out = CircuitProgram();
  std::vector<HostPass> passes;
  if (cache && cache->valid) {
    *out = cache->head;
    passes = cache->passes;
  } else {
    lower_dm_gates(noise, b, c, out, &passes);
    if (cache) { cache->valid = true; cache->head = *out; cache->passes = passes; }
  }
  if (out->status) return;
  if (fold > 1) {  // G -> G (G^dagger G)^((fold-1)/2) = fold copies of the self-inverse cx, each with its error
    for (HostPass& p : passes) {
      std::vector<MacroOp> ops;
      ops.reserve(p.ops.size() * fold);
      for (size_t i = 0; i < p.ops.size(); ++i) {
        const MacroOp& o = p.ops[i];
        ops.push_back(o);
        if (o.twoq == Q_CXN_AB || o.twoq == Q_CXN_BA) {  // cx fused with its relaxation error
          MacroOp rep{};
          rep.twoq = o.twoq; rep.off_2 = o.off_2;
          for (int r = 1; r < fold; ++r) ops.push_back(rep);
        } else if (o.twoq == Q_CX_AB || o.twoq == Q_CX_BA) {
          // bare cx: its error, if any, is the next op (the circuit has no other 2-qubit gates) -- repeat the pair
          MacroOp rep{};
          rep.twoq = o.twoq;
          const bool has_err = i + 1 < p.ops.size() && p.ops[i + 1].pre_a == P_NONE && p.ops[i + 1].pre_b == P_NONE &&
                               p.ops[i + 1].twoq >= Q_RELAX;
          if (has_err) ops.push_back(p.ops[i + 1]);
          for (int r = 1; r < fold; ++r) { ops.push_back(rep); if (has_err) ops.push_back(p.ops[i + 1]); }
          if (has_err) ++i;
        }
      }
      p.ops.swap(ops);
    }
    int64_t n_cx = 0;
    for (int64_t g = b.op_offsets[c]; g < b.op_offsets[c + 1]; ++g) n_cx += b.ops[g].opcode == BWQ_G_CX;
    out->n_gates += n_cx * (fold - 1);
  }
  const int nd = out->n_digits;
  pack_dm_passes(noise, opt, nd, passes, out);
}

// first stage: gates + attached errors -> register passes (and the circuit's digits / Pauli terms)
static void lower_dm_gates(const NoiseTable& noise, const bwq_batch& b, int c, CircuitProgram* out, std::vector<HostPass>* passes_out) {
  std::vector<HostPass>& passes = *passes_out;
  const int nq = b.n_qubits[c];
  const int64_t g0 = b.op_offsets[c], g1 = b.op_offsets[c + 1];
  out->n_gates = g1 - g0;
  if (nq < 0 || nq > 64) { out->status = BWQ_CIRC_BAD_QUBIT; return; }

  // active qubits -> digits (ascending physical order)
  std::vector<int> digit_of(nq, -1);
  {
    std::vector<char> used(nq, 0);
    for (int64_t g = g0; g < g1; ++g) {
      const bwq_op& op = b.ops[g];
      bool two = gate_is_2q(op.opcode);
      if (op.q0 >= nq || (two && (op.q1 >= nq || op.q1 == op.q0))) { out->status = BWQ_CIRC_BAD_QUBIT; return; }
      used[op.q0] = 1;
      if (two) used[op.q1] = 1;
    }
    for (int q = 0; q < nq; ++q) if (used[q]) { digit_of[q] = (int)out->active.size(); out->active.push_back(q); }
  }
  while (out->active.size() < 2) out->active.push_back(-1);  // pad: idle digits stay I/Z = 1
  const int nd = out->n_digits = (int)out->active.size();
  if (nd > kMaxDmQubits) { out->status = BWQ_CIRC_TOO_WIDE; return; }
  lower_terms(b, c, digit_of, nq, out);
  if (out->status) return;

  // ---- gates -> passes
  out->mats.reserve((size_t)(g1 - g0) * 6 + 64);
  passes.reserve(64);
  std::vector<double> pend(16 * nd);
  std::vector<char> has(nd, 0);
  std::vector<int> last(nd, -1);
  auto local = [&](const double* m, int n) { return kLocalMat | (uint32_t)push_mat(out->mats, m, n); };
  // the macro-op a 1-qubit map on slot a/b (or a 2-qubit op) goes into
  auto slot_for_pre = [&](HostPass& p, bool on_a) -> MacroOp& {
    if (p.ops.empty() || p.ops.back().twoq != Q_NONE || (on_a ? p.ops.back().pre_a : p.ops.back().pre_b) != P_NONE)
      p.ops.push_back(MacroOp{});
    return p.ops.back();
  };
  auto slot_for_two = [&](HostPass& p) -> MacroOp& {
    if (p.ops.empty() || p.ops.back().twoq != Q_NONE) p.ops.push_back(MacroOp{});
    return p.ops.back();
  };
  auto flush = [&](int d, HostPass& p) {
    if (!has[d]) return;
    has[d] = 0;
    const double* m = &pend[16 * d];
    const bool on_a = d == p.qa;
    uint8_t kind;
    uint32_t off;
    const bool tp = m[0] == 1.0 && m[1] == 0.0 && m[2] == 0.0 && m[3] == 0.0;
    const bool rot = tp && m[4] == 0.0 && m[8] == 0.0 && m[12] == 0.0 && m[7] == 0.0 && m[11] == 0.0 &&
                     m[13] == 0.0 && m[14] == 0.0 && m[15] == 1.0 && m[5] == m[10] && m[6] == -m[9];
    if (!tp) {  // not trace preserving: general 4x4
      kind = P_DENSE; off = local(m, 16);
    } else if (rot) {
      double c = m[5], sn = m[9];
      if (c == 1.0 && sn == 0.0) return;  // identity
      // R(theta) = sign * Shear_x(-t) Shear_y(s) Shear_x(-t), theta' = theta (-/+ pi when cos < 0)
      double sign = 1.0;
      if (c < 0.0) { c = -c; sn = -sn; sign = -1.0; }
      const double t = sn / (1.0 + c);  // tan(theta'/2), |t| <= 1
      const double sh[4] = {t, sn, sign, 0.0};
      kind = P_ROT; off = local(sh, 4);
    } else {
      kind = P_AFF; off = local(m + 4, 12);
    }
    MacroOp& mo = slot_for_pre(p, on_a);
    if (on_a) { mo.pre_a = kind; mo.off_a = off; } else { mo.pre_b = kind; mo.off_b = off; }
  };
  for (int64_t g = g0; g < g1; ++g) {
    const bwq_op& op = b.ops[g];
    const int npar = gate_num_params(op.opcode);
    if (npar && (int64_t)op.param_idx + npar > b.n_params) { out->status = BWQ_CIRC_BAD_OP; return; }
    const double* par = npar ? b.params + op.param_idx : nullptr;
    if (!gate_is_2q(op.opcode)) {
      double r[16];
      int d = digit_of[op.q0];
      const NoiseEntry* ne = noise.find(op.opcode, op.q0, 255);
      if (!ne && has[d] && (op.opcode == BWQ_G_RZ || op.opcode == BWQ_G_P)) {  // noise-free virtual rotation
        mat4_rot_mul(std::cos(par[0]), std::sin(par[0]), &pend[16 * d], &pend[16 * d]);
        continue;
      }
      const double* gm = r;
      if (ne && op.opcode < 32 && op.q0 < 64 && !noise.fixed_ok.empty() && noise.fixed_ok[op.opcode * 64 + op.q0]) {
        gm = &noise.fixed_ptm[(size_t)(op.opcode * 64 + op.q0) * 16];  // error x gate, precomputed
      } else {
        if (!gate_ptm1(op.opcode, par, r)) { out->status = BWQ_CIRC_BAD_OP; return; }
        if (ne) mat4_mul(&noise.data[ne->off], r, r);
      }
      if (has[d]) mat4_mul(gm, &pend[16 * d], &pend[16 * d]);
      else { std::memcpy(&pend[16 * d], gm, sizeof r); has[d] = 1; }
      continue;
    }
    int d0 = digit_of[op.q0], d1 = digit_of[op.q1];
    int pi;
    if (last[d0] >= 0 && last[d0] == last[d1]) pi = last[d0];
    else {
      pi = (int)passes.size();
      passes.push_back(HostPass{d0, d1, {}});
      last[d0] = last[d1] = pi;
    }
    HostPass& p = passes[pi];
    flush(d0, p);
    flush(d1, p);
    const bool same = (d0 == p.qa);
    const NoiseEntry* ne = noise.find(op.opcode, op.q0, op.q1);
    if (op.opcode == BWQ_G_CX) {
      MacroOp& mo = slot_for_two(p);
      if (ne && ne->kind == BWQ_NOISE_RELAX2) {  // CX + its relaxation/depolarizing error: one fused op
        mo.twoq = same ? Q_CXN_AB : Q_CXN_BA; mo.off_2 = (uint32_t)ne->off;
        ne = nullptr;
      } else {
        mo.twoq = same ? Q_CX_AB : Q_CX_BA;
      }
    } else {
      double u[32], r[256];
      if (!gate_unitary(op.opcode, par, u)) { out->status = BWQ_CIRC_BAD_OP; return; }
      ptm_from_unitary2(u, r);
      MacroOp& mo = slot_for_two(p);
      mo.twoq = same ? Q_DENSE : Q_DENSE_SW; mo.off_2 = local(r, 256);
    }
    if (ne) {
      MacroOp& mo = slot_for_two(p);
      mo.twoq = ne->kind == BWQ_NOISE_RELAX2 ? (same ? Q_RELAX : Q_RELAX_SW) : (same ? Q_DENSE : Q_DENSE_SW);
      mo.off_2 = (uint32_t)ne->off;
    }
  }
  for (int d = 0; d < nd; ++d) {
    if (!has[d]) continue;
    if (last[d] < 0) {  // qubit with 1-qubit gates only: give it a pass with any partner
      int partner = d == 0 ? 1 : 0;
      passes.push_back(HostPass{d, partner, {}});
      last[d] = (int)passes.size() - 1;
    }
    flush(d, passes[last[d]]);
  }

}

// second stage: passes -> sweeps, each one self-contained program block
static void pack_dm_passes(const NoiseTable& noise, const LowerOptions& opt, const int nd, std::vector<HostPass>& passes, CircuitProgram* out) {
  // ---- parameter sizes (8-byte words) and sources
  auto pre_words = [](uint8_t k) { return k == P_ROT ? 4 : k == P_AFF ? 12 : k == P_DENSE ? 16 : 0; };
  auto two_words = [](uint8_t k) {
    return (k == Q_CXN_AB || k == Q_CXN_BA || k == Q_RELAX || k == Q_RELAX_SW) ? 26 : (k == Q_DENSE || k == Q_DENSE_SW) ? 256 : 0;
  };
  auto src_ptr = [&](uint32_t off) -> const double* {
    return (off & kLocalMat) ? &out->mats[off & ~kLocalMat] : &noise.data[off];
  };
  // TMA tile layout: circuits wider than the 6-digit tile, default tiling (see program.h)
  const bool tma = opt.tma && nd > 6 && nd <= 15 && std::min(std::max(opt.tile_qubits, 3), std::min(nd, kMaxTileQubits)) == 6 &&
                   std::min(std::max(opt.low_qubits, 1), 4) == 2;
  out->tma = tma;
  // bytes a pass adds to a sweep block; parameters already present (same source) are shared
  struct Seen { uint32_t key; uint16_t off; };
  auto pass_bytes = [&](const HostPass& p, const std::vector<Seen>& seen) {
    int bytes = (int)sizeof(PassHdr) + (int)sizeof(BlockOp) * (int)p.ops.size() + (tma ? 64 : 0);
    std::vector<uint32_t> mine;
    auto add = [&](uint32_t key, int words) {
      if (!words) return;
      if (!(key & kLocalMat)) {  // circuit-local matrices are unique per use: only noise-table entries are shared
        for (const Seen& sn : seen) if (sn.key == key) return;
        for (uint32_t k : mine) if (k == key) return;
        mine.push_back(key);
      }
      bytes += ((words + 1) & ~1) * 8;
    };
    for (const MacroOp& o : p.ops) { add(o.off_a, pre_words(o.pre_a)); add(o.off_b, pre_words(o.pre_b)); add(o.off_2, two_words(o.twoq)); }
    return bytes;
  };
  const int block_bytes = kBlockBytes;
  // split passes whose block alone would not fit the shared-memory program buffer
  {
    const int cap = block_bytes - (int)sizeof(BlockHdr);
    std::vector<HostPass> split;
    for (HostPass& p : passes) {
      if (pass_bytes(p, {}) <= cap) { split.push_back(std::move(p)); continue; }
      HostPass cur{p.qa, p.qb, {}};
      for (const MacroOp& o : p.ops) {
        cur.ops.push_back(o);
        if (pass_bytes(cur, {}) > cap) {
          cur.ops.pop_back();
          if (cur.ops.empty()) { out->status = BWQ_CIRC_BAD_OP; return; }  // a single op always fits
          split.push_back(cur);
          cur.ops.assign(1, o);
        }
      }
      if (!cur.ops.empty()) split.push_back(cur);
    }
    passes.swap(split);
  }

  // ---- passes -> sweeps (greedy in program order; a pass that does not fit blocks its qubits)
  // tile = min(tile_qubits, n) digits, at least 3 when the state is larger than the tile so that
  // digit 0 (always resident: 16-byte global accesses) leaves two free slots for any pass
  const int kq = std::min(std::max(opt.tile_qubits, 3), std::min(nd, kMaxTileQubits));
  const int mlow = (nd <= kq) ? 0 : std::min(std::max(opt.low_qubits, 1), kq - 2);
  const int np = (int)passes.size();
  out->n_passes = np;
  std::vector<char> done(np, 0);
  int first = 0, remaining = np;
  uint32_t touched_mask = 0;  // digits acted on so far (the others still hold |0><0|: I/Z only)
  std::vector<char> in_tile(nd), blocked(nd);
  std::vector<int> sel;
  std::vector<Seen> seen;
  while (remaining > 0) {
    std::fill(in_tile.begin(), in_tile.end(), 0);
    std::fill(blocked.begin(), blocked.end(), 0);
    int nt = 0, nblocked = 0;
    for (int d = 0; d < mlow; ++d) { in_tile[d] = 1; ++nt; }
    sel.clear();
    seen.clear();
    int bytes = (int)sizeof(BlockHdr) + (tma ? 64 : 0);  // TMA layout: room for the direct-store corner table
    while (first < np && done[first]) ++first;
    for (int i = first; i < np && nblocked < nd; ++i) {
      if (done[i]) continue;
      const HostPass& p = passes[i];
      auto block = [&](int d) { if (!blocked[d]) { blocked[d] = 1; ++nblocked; } };
      if (blocked[p.qa] || blocked[p.qb]) { block(p.qa); block(p.qb); continue; }
      const int need = (!in_tile[p.qa]) + (!in_tile[p.qb]);
      const int add = pass_bytes(p, seen);
      if (nt + need > kq || bytes + add > block_bytes) { block(p.qa); block(p.qb); continue; }
      if (!in_tile[p.qa]) { in_tile[p.qa] = 1; ++nt; }
      if (!in_tile[p.qb]) { in_tile[p.qb] = 1; ++nt; }
      done[i] = 1;
      --remaining;
      sel.push_back(i);
      bytes += add;
      // remember the parameters this pass brings (offsets are assigned at emission)
      for (const MacroOp& o : p.ops) {
        auto note = [&](uint32_t key, int words) {
          if (!words || (key & kLocalMat)) return;
          for (const Seen& sn : seen) if (sn.key == key) return;
          seen.push_back(Seen{key, 0});
        };
        note(o.off_a, pre_words(o.pre_a)); note(o.off_b, pre_words(o.pre_b)); note(o.off_2, two_words(o.twoq));
      }
    }
    if (sel.empty()) { out->status = BWQ_CIRC_BAD_OP; return; }  // cannot happen
    for (int d = 0; d < nd && nt < kq; ++d) if (!in_tile[d]) { in_tile[d] = 1; ++nt; }
    int slot_of[kMaxDmQubits];
    SweepDesc sw{};
    int s = 0;
    for (int d = 0; d < nd; ++d) if (in_tile[d]) { sw.pos[s] = (uint8_t)d; slot_of[d] = s++; }
    if (tma) {
      // box order of the four upper digits: a digit that shares a pass with digit 1 (or 0) should not
      // sit in slot 2 (3), where its bits feed the 128-byte swizzle of the chunk bits its partner
      // occupies (2-way bank conflicts, program.h) -- such digits go to the last box dims
      int key[4], ord[4] = {0, 1, 2, 3};
      for (int k = 0; k < 4; ++k) {
        const int d = sw.pos[2 + k];
        key[k] = 0;
        for (int i : sel) {
          const int other = passes[i].qa == d ? passes[i].qb : passes[i].qb == d ? passes[i].qa : -1;
          if (other == 1) key[k] = 2;
          else if (other == 0 && key[k] < 1) key[k] = 1;
        }
      }
      std::stable_sort(ord, ord + 4, [&](int a, int c) { return key[a] < key[c]; });
      uint8_t code = 0;
      for (int k = 0; k < 4; ++k) { slot_of[sw.pos[2 + ord[k]]] = 2 + k; code |= (uint8_t)(ord[k] << (2 * k)); }
      sw.pos[6] = code;
      sw.pos[7] = kTmaSweep;
    }

    // ---- direct passes: passes on disjoint qubits commute, so an eligible pass (neither target in
    // the two lowest slots) with no predecessor / successor on its qubits inside this sweep is
    // moved to the front / back and exchanges its register groups with global memory directly
    auto eligible = [&](int i) { return slot_of[passes[i].qa] >= 2 && slot_of[passes[i].qb] >= 2; };
    auto shares = [&](int i, int j) {
      return passes[i].qa == passes[j].qa || passes[i].qa == passes[j].qb || passes[i].qb == passes[j].qa || passes[i].qb == passes[j].qb;
    };
    bool first_direct = false, last_direct = false;
    if (!tma && (opt.direct & kPassLoadDirect)) {
      for (size_t k = 0; k < sel.size() && !first_direct; ++k) {
        if (!eligible(sel[k])) continue;
        bool free_ = true;
        for (size_t e = 0; e < k && free_; ++e) free_ = !shares(sel[e], sel[k]);
        if (!free_) continue;
        std::rotate(sel.begin(), sel.begin() + k, sel.begin() + k + 1);
        first_direct = true;
      }
    }
    if ((opt.direct & kPassStoreDirect) && (!tma || opt.tma_direct_store)) {  // TMA layout: direct 16-byte stores on request
      const size_t stop = (first_direct && sel.size() > 1) ? 1 : 0;  // the front pass stays in front
      for (size_t k = sel.size(); k-- > stop && !last_direct;) {
        if (!eligible(sel[k])) continue;
        bool free_ = true;
        for (size_t l = k + 1; l < sel.size() && free_; ++l) free_ = !shares(sel[l], sel[k]);
        if (!free_) continue;
        std::rotate(sel.begin() + k, sel.begin() + k + 1, sel.end());
        last_direct = true;
      }
    }

    // ---- emit the block
    size_t n_ops_total = 0;
    for (int i : sel) n_ops_total += passes[i].ops.size();
    const size_t ext_words = tma ? 8 * sel.size() + 8 : 0;  // corner tables: 16 x u32 per pass + the direct-store table
    const size_t hdr_words = (sizeof(BlockHdr) + sizeof(PassHdr) * sel.size() + sizeof(BlockOp) * n_ops_total) / 8 + ext_words;
    const size_t blk_begin = out->prog.size();
    out->prog.resize(blk_begin + (size_t)bytes / 8, 0);
    uint64_t* blk = out->prog.data() + blk_begin;
    BlockHdr* bh = reinterpret_cast<BlockHdr*>(blk);
    bh->n_passes = (int32_t)sel.size();
    bh->ext_q16 = tma ? (int32_t)((hdr_words - ext_words) / 2) : 0;
    uint32_t* ext = reinterpret_cast<uint32_t*>(blk + (hdr_words - ext_words));
    if (tma) for (int d = 0; d < nd; ++d) if (in_tile[d]) bh->slot_pos[slot_of[d]] = (uint8_t)d;
    PassHdr* ph = reinterpret_cast<PassHdr*>(blk + sizeof(BlockHdr) / 8);
    BlockOp* bo = reinterpret_cast<BlockOp*>(blk + (sizeof(BlockHdr) + sizeof(PassHdr) * sel.size()) / 8);
    size_t op_cursor = 0, par_cursor = hdr_words;
    seen.clear();
    auto place = [&](uint32_t key, int words) -> uint16_t {
      if (!words) return 0;
      const bool shared = !(key & kLocalMat);
      if (shared) for (const Seen& sn : seen) if (sn.key == key) return sn.off;
      const uint16_t off = (uint16_t)par_cursor;
      std::memcpy(blk + par_cursor, src_ptr(key), (size_t)words * 8);
      par_cursor += (size_t)((words + 1) & ~1);
      if (shared) seen.push_back(Seen{key, off});
      return off;
    };
    for (size_t k = 0; k < sel.size(); ++k) {
      const HostPass& p = passes[sel[k]];
      ph[k].ops_q16 = (uint16_t)((sizeof(BlockHdr) + sizeof(PassHdr) * sel.size() + sizeof(BlockOp) * op_cursor) / 16);
      ph[k].n_ops = (uint16_t)p.ops.size();
      ph[k].sa = (uint8_t)slot_of[p.qa];
      ph[k].sb = (uint8_t)slot_of[p.qb];
      if (k == 0 && first_direct) {
        ph[k].flags |= kPassLoadDirect;
        sw.pos[7] = (uint8_t)(kFirstDirect | ph[k].sa | (ph[k].sb << 3));
      }
      if (k + 1 == sel.size() && last_direct) ph[k].flags |= kPassStoreDirect;
      if (tma) {
        const int sa = ph[k].sa, sb = ph[k].sb;
        ph[k].row = (uint8_t)tma_pair_row(sa, sb);
        ph[k].gofs = 8u * tswz(1u << tma_pass_layout(std::min(sa, sb), std::max(sa, sb)).beta);
        for (int db = 0; db < 4; ++db)
          for (int da = 0; da < 4; ++da) ext[16 * k + da + 4 * db] = 8u * tswz((uint32_t(da) << (2 * sa)) | (uint32_t(db) << (2 * sb)));
        ph[k].sig = (uint8_t)classify_pass(p.ops, sa != 0 && sb != 0);
        if (ph[k].flags & kPassStoreDirect)
          for (int db = 0; db < 4; ++db)
            for (int da = 0; da < 4; ++da)
              ext[16 * sel.size() + da + 4 * db] = (uint32_t(da) << (2 * p.qa)) | (uint32_t(db) << (2 * p.qb));
      }
      for (const MacroOp& o : p.ops) {
        BlockOp& d = bo[op_cursor++];
        d.pre_a = o.pre_a; d.pre_b = o.pre_b; d.twoq = o.twoq;
        d.off_a = place(o.off_a, pre_words(o.pre_a));
        d.off_b = place(o.off_b, pre_words(o.pre_b));
        d.off_2 = place(o.off_2, two_words(o.twoq));
        if (o.pre_a == P_DENSE || o.pre_b == P_DENSE || o.twoq == Q_DENSE || o.twoq == Q_DENSE_SW) out->needs_dense = true;
      }
    }
    if (par_cursor * 8 != (size_t)bytes) { out->status = BWQ_CIRC_BAD_OP; return; }  // accounting bug guard
    sw.blk_q16 = (uint32_t)(blk_begin / 2);
    for (int i : sel) { touched_mask |= 1u << passes[i].qa; touched_mask |= 1u << passes[i].qb; }
    sw.blk_len_q16 = (uint32_t)(bytes / 16) | ((~touched_mask & ((1u << nd) - 1u)) << 16);
    out->sweeps.push_back(sw);
  }
}

// ----------------------------------------------------------------------------- SV lowering
// 2x2 complex product in plain real arithmetic: the (ac - bd, ad + bc) of std::complex without
// its NaN recovery path (identical for finite values)
static inline cd cmulf(const cd& x, const cd& y) {
  return cd(x.real() * y.real() - x.imag() * y.imag(), x.real() * y.imag() + x.imag() * y.real());
}
static void mat2c_mul(const cd* a, const cd* b, cd* out) {
  cd t[4];
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) t[i * 2 + j] = cmulf(a[i * 2], b[j]) + cmulf(a[i * 2 + 1], b[2 + j]);
  for (int i = 0; i < 4; ++i) out[i] = t[i];
}

void lower_sv_circuit(const bwq_batch& b, int c, SvProgram* out) {
  *out = SvProgram();
  const int nq = b.n_qubits[c];
  const int64_t g0 = b.op_offsets[c], g1 = b.op_offsets[c + 1];
  out->n_gates = g1 - g0;
  if (nq < 0 || nq > 64) { out->status = BWQ_CIRC_BAD_QUBIT; return; }
  std::vector<int> bit_of(nq, -1);
  {
    std::vector<char> used(nq, 0);
    for (int64_t g = g0; g < g1; ++g) {
      const bwq_op& op = b.ops[g];
      bool two = gate_is_2q(op.opcode);
      if (op.q0 >= nq || (two && (op.q1 >= nq || op.q1 == op.q0))) { out->status = BWQ_CIRC_BAD_QUBIT; return; }
      if (op.opcode == BWQ_G_RESET) { out->status = BWQ_CIRC_BAD_OP; return; }
      used[op.q0] = 1;
      if (two) used[op.q1] = 1;
    }
    for (int q = 0; q < nq; ++q) if (used[q]) { bit_of[q] = (int)out->active.size(); out->active.push_back(q); }
  }
  if (out->active.empty()) out->active.push_back(-1);
  const int nb = out->n_bits = (int)out->active.size();
  if (nb > kMaxSvQubits) { out->status = BWQ_CIRC_TOO_WIDE; return; }

  // terms
  {
    int64_t o0 = b.obs_offsets[c], o1 = b.obs_offsets[c + 1];
    int64_t t0 = b.term_offsets[o0], t1 = b.term_offsets[o1];
    uint64_t valid = nq >= 64 ? ~0ull : ((1ull << nq) - 1);
    for (int64_t t = t0; t < t1; ++t) {
      uint64_t x = b.term_x[t], z = b.term_z[t];
      if ((x | z) & ~valid) { out->status = BWQ_CIRC_BAD_QUBIT; return; }
      uint32_t cx = 0, cz = 0;
      int ny = 0;
      double coeff = b.term_coeff[t];
      for (int q = 0; q < nq; ++q) {
        int xb = (x >> q) & 1, zb = (z >> q) & 1;
        if (!xb && !zb) continue;
        if (bit_of[q] < 0) { if (xb) coeff = 0.0; continue; }
        if (xb) cx |= 1u << bit_of[q];
        if (zb) cz |= 1u << bit_of[q];
        if (xb && zb) ++ny;
      }
      out->term_x.push_back(cx); out->term_z.push_back(cz); out->term_ny.push_back(ny);
      out->term_coeff.push_back(coeff);
    }
  }

  std::vector<cd> pend(4 * nb);
  std::vector<char> has(nb, 0);
  auto flush = [&](int q) {
    if (!has[q]) return;
    double m[8];
    for (int i = 0; i < 4; ++i) { m[2 * i] = pend[4 * q + i].real(); m[2 * i + 1] = pend[4 * q + i].imag(); }
    SvOp o{}; o.kind = SV_U1; o.q0 = (uint8_t)q; o.q1 = 0; o.off = push_mat(out->mats, m, 8);
    out->ops.push_back(o);
    has[q] = 0;
  };
  for (int64_t g = g0; g < g1; ++g) {
    const bwq_op& op = b.ops[g];
    const int npar = gate_num_params(op.opcode);
    if (npar && (int64_t)op.param_idx + npar > b.n_params) { out->status = BWQ_CIRC_BAD_OP; return; }
    const double* par = npar ? b.params + op.param_idx : nullptr;
    if (!gate_is_2q(op.opcode)) {
      cd u[4];
      if (!unitary1(op.opcode, par, u)) { out->status = BWQ_CIRC_BAD_OP; return; }
      int q = bit_of[op.q0];
      if (has[q]) mat2c_mul(u, &pend[4 * q], &pend[4 * q]);
      else { for (int i = 0; i < 4; ++i) pend[4 * q + i] = u[i]; has[q] = 1; }
      continue;
    }
    int q0 = bit_of[op.q0], q1 = bit_of[op.q1];
    flush(q0); flush(q1);
    SvOp o{}; o.q0 = (uint8_t)q0; o.q1 = (uint8_t)q1;
    if (op.opcode == BWQ_G_CX) { o.kind = SV_CX; o.off = 0; }
    else {
      double u[32];
      if (!gate_unitary(op.opcode, par, u)) { out->status = BWQ_CIRC_BAD_OP; return; }
      o.kind = SV_U2; o.off = push_mat(out->mats, u, 32);
    }
    out->ops.push_back(o);
  }
  for (int q = 0; q < nb; ++q) flush(q);
}

}  // namespace bwq
