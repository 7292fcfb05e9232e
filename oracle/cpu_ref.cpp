// Aer-style CPU restatement of the exact expectation-value path -- TEST / BASELINE INFRASTRUCTURE.
//
// The reference's CPU implementation of this path is qiskit-aer (C++/OpenMP, not in tree and not
// installable here; call sites blackwater/data/utils.py:422-430).  This file restates Aer's
// density_matrix and statevector methods with Aer's algorithmic choices so that bench.py can time
// "the reference's own CPU path" on the GPU box's host cores:
//   * state = complex128 vec(rho), column-stacked: qubit q <-> bits q and q+n      [3P Aer]
//   * every gate applied as its own pass: cx / x as index permutations, rz as a diagonal,
//     other unitaries as conj(U)(x)U superoperators; every attached error as ONE dense
//     superoperator (4x4 / 16x16 complex) right after its gate                     [3P Aer]
//   * untouched qubits truncated; optional fusion of consecutive ops on the same <=2 qubits into
//     one 16x16 superoperator for n >= 7 (Aer: fusion_max_qubit=2 for density matrices)
//   * OpenMP over circuits (Aer max_parallel_experiments) for small states, over amplitudes above
//   * Tr(rho P) / <psi|P|psi> by x/z-mask traversal (Aer expval_pauli)
// It is checked against the numpy oracle (tests/test_cpu_ref.py) and is never linked into the product.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/bwq.h"

using cd = std::complex<double>;
static const cd I_(0.0, 1.0);

extern "C" {
typedef struct {
  int32_t n_entries;
  const uint16_t* opcode;
  const uint8_t* q0;
  const uint8_t* q1;       // 255 = 1-qubit entry; q0 = 255 = all-qubit default
  const int64_t* data_off; // offset in complex numbers; 16 (1q) or 256 (2q) entries, row-major
  const double* data;      // interleaved re/im
} cpuref_noise;
}

namespace {

bool is2q(uint16_t op) { return (op >= BWQ_G_CX && op <= BWQ_G_ECR) || op == BWQ_G_UNITARY2; }

// Gate matrices are DATA here: oracle/cpu_ref.py evaluates oracle/gates.py (the numpy restatement of
// Qiskit's standard gates) for every op of the batch and hands the table over, so this file shares
// no gate library with the product (ml_qem_b200/csrc/lowering.cpp) -- a wrong matrix there cannot be
// common-mode here.  Entry g: 4 (1-qubit) or 16 (2-qubit) complex numbers, row-major, local index
// i_q0 + 2 i_q1; reset has no matrix (its superoperator is built below).
struct GateTable {
  const double* mats = nullptr;   // interleaved re/im
  const int64_t* off = nullptr;   // per op of the batch: offset in complex numbers, -1 = unsupported
};
const GateTable* g_tab = nullptr;  // set for the duration of one cpuref_*_run call (calls are not re-entrant)

bool unitary1(int64_t g, cd* m) {
  if (!g_tab || !g_tab->off || g_tab->off[g] < 0) return false;
  const double* p = g_tab->mats + 2 * g_tab->off[g];
  for (int i = 0; i < 4; ++i) m[i] = cd(p[2 * i], p[2 * i + 1]);
  return true;
}
bool unitary2(int64_t g, cd* m) {
  if (!g_tab || !g_tab->off || g_tab->off[g] < 0) return false;
  const double* p = g_tab->mats + 2 * g_tab->off[g];
  for (int i = 0; i < 16; ++i) m[i] = cd(p[2 * i], p[2 * i + 1]);
  return true;
}

// ---- generic k-"bit" matrix application on a 2^N vector (bits ascending order in `bits`)
inline uint64_t insert_zero_bits(uint64_t g, const int* sorted_bits, int k) {
  for (int i = 0; i < k; ++i) {
    uint64_t low = g & ((1ull << sorted_bits[i]) - 1);
    g = ((g >> sorted_bits[i]) << (sorted_bits[i] + 1)) | low;
  }
  return g;
}

// v <- M v on local bits bits[0..k-1] (local index bit j <-> global bit bits[j]); M row-major 2^k x 2^k
void apply_matrix(cd* v, int nbits, const int* bits, int k, const cd* m, bool par) {
  const int dim = 1 << k;
  int sorted[8];
  for (int i = 0; i < k; ++i) sorted[i] = bits[i];
  std::sort(sorted, sorted + k);
  uint64_t offs[256];
  for (int l = 0; l < dim; ++l) {
    uint64_t o = 0;
    for (int j = 0; j < k; ++j) if ((l >> j) & 1) o |= 1ull << bits[j];
    offs[l] = o;
  }
  const int64_t groups = int64_t(1) << (nbits - k);
#pragma omp parallel for if (par) schedule(static)
  for (int64_t g = 0; g < groups; ++g) {
    const uint64_t b = insert_zero_bits((uint64_t)g, sorted, k);
    cd x[16], y[16];
    for (int l = 0; l < dim; ++l) x[l] = v[b | offs[l]];
    for (int r = 0; r < dim; ++r) {
      cd s = 0;
      for (int c = 0; c < dim; ++c) s += m[r * dim + c] * x[c];
      y[r] = s;
    }
    for (int l = 0; l < dim; ++l) v[b | offs[l]] = y[l];
  }
}

void apply_diag1(cd* v, int nbits, int bit, cd d0, cd d1, bool par) {
  const int64_t N = int64_t(1) << nbits;
#pragma omp parallel for if (par) schedule(static)
  for (int64_t i = 0; i < N; ++i) v[i] *= ((i >> bit) & 1) ? d1 : d0;
}

void apply_x_bit(cd* v, int nbits, int bit, bool par) {
  const int64_t groups = int64_t(1) << (nbits - 1);
  const int sorted[1] = {bit};
#pragma omp parallel for if (par) schedule(static)
  for (int64_t g = 0; g < groups; ++g) {
    uint64_t b = insert_zero_bits((uint64_t)g, sorted, 1);
    std::swap(v[b], v[b | (1ull << bit)]);
  }
}

void apply_cx_bits(cd* v, int nbits, int cbit, int tbit, bool par) {
  const int64_t groups = int64_t(1) << (nbits - 2);
  int sorted[2] = {std::min(cbit, tbit), std::max(cbit, tbit)};
#pragma omp parallel for if (par) schedule(static)
  for (int64_t g = 0; g < groups; ++g) {
    uint64_t b = insert_zero_bits((uint64_t)g, sorted, 2) | (1ull << cbit);
    std::swap(v[b], v[b | (1ull << tbit)]);
  }
}

void superop_from_unitary(const cd* u, int dim, cd* s) {  // S = conj(U) (x) U : index r + c*dim
  for (int c = 0; c < dim; ++c) for (int r = 0; r < dim; ++r)
    for (int c2 = 0; c2 < dim; ++c2) for (int r2 = 0; r2 < dim; ++r2)
      s[(r + c * dim) * dim * dim + (r2 + c2 * dim)] = std::conj(u[c * dim + c2]) * u[r * dim + r2];
}

struct NoiseLookup {
  const cpuref_noise* t;
  const cd* find(uint16_t op, int q0, int q1) const {
    if (!t) return nullptr;
    const cd* def = nullptr;
    for (int i = 0; i < t->n_entries; ++i) {
      if (t->opcode[i] != op) continue;
      if (t->q0[i] == q0 && t->q1[i] == (q1 & 255)) return reinterpret_cast<const cd*>(t->data) + t->data_off[i];
      if (t->q0[i] == 255) def = reinterpret_cast<const cd*>(t->data) + t->data_off[i];
    }
    return def;
  }
};

struct Compact {
  std::vector<int> pos;  // physical -> compact (-1 idle)
  int n = 0;
};

Compact compact_qubits(const bwq_batch& b, int c) {
  Compact out;
  out.pos.assign(b.n_qubits[c], -1);
  std::vector<char> used(b.n_qubits[c], 0);
  for (int64_t g = b.op_offsets[c]; g < b.op_offsets[c + 1]; ++g) {
    used[b.ops[g].q0] = 1;
    if (is2q(b.ops[g].opcode)) used[b.ops[g].q1] = 1;
  }
  for (int q = 0; q < b.n_qubits[c]; ++q) if (used[q]) out.pos[q] = out.n++;
  if (out.n == 0) out.n = 1;
  return out;
}

// pending fused 2-qubit superoperator (Aer-style fusion, n >= threshold)
struct Fused {
  bool active = false;
  int a = -1, b = -1;
  cd s[256];
};

void mat16_mul(const cd* x, const cd* y, cd* out) {
  cd t[256];
  for (int i = 0; i < 16; ++i) for (int j = 0; j < 16; ++j) {
    cd s = 0;
    for (int k = 0; k < 16; ++k) s += x[i * 16 + k] * y[k * 16 + j];
    t[i * 16 + j] = s;
  }
  std::memcpy(out, t, sizeof t);
}

// embed a 4x4 superop on local qubit `which` (0/1) of a pair into 16x16; index r0 + 2 r1 + 4 c0 + 8 c1
void embed_1q(const cd* s4, int which, cd* s16) {
  for (int i = 0; i < 256; ++i) s16[i] = 0;
  for (int r = 0; r < 2; ++r) for (int c = 0; c < 2; ++c) for (int r2 = 0; r2 < 2; ++r2) for (int c2 = 0; c2 < 2; ++c2)
    for (int ro = 0; ro < 2; ++ro) for (int co = 0; co < 2; ++co) {
      int row = which == 0 ? (r + 2 * ro + 4 * c + 8 * co) : (ro + 2 * r + 4 * co + 8 * c);
      int col = which == 0 ? (r2 + 2 * ro + 4 * c2 + 8 * co) : (ro + 2 * r2 + 4 * co + 8 * c2);
      s16[row * 16 + col] = s4[(r + 2 * c) * 4 + (r2 + 2 * c2)];
    }
}

void swap_pair_order(cd* s16) {  // (q0,q1) -> (q1,q0)
  cd t[256];
  auto sw = [](int i) { int r0 = i & 1, r1 = (i >> 1) & 1, c0 = (i >> 2) & 1, c1 = (i >> 3) & 1; return r1 + 2 * r0 + 4 * c1 + 8 * c0; };
  for (int i = 0; i < 16; ++i) for (int j = 0; j < 16; ++j) t[sw(i) * 16 + sw(j)] = s16[i * 16 + j];
  std::memcpy(s16, t, sizeof t);
}

int dm_circuit(const bwq_batch& b, int c, const cpuref_noise* noise, double* out, bool par, int fusion_threshold) {
  Compact cq = compact_qubits(b, c);
  const int n = cq.n;
  if (n > 15) return BWQ_CIRC_TOO_WIDE;
  const int nbits = 2 * n;
  std::vector<cd> v(size_t(1) << nbits, cd(0));
  v[0] = 1;
  NoiseLookup nl{noise};
  const bool fuse = n >= fusion_threshold;
  Fused f;
  auto flush = [&]() {
    if (!f.active) return;
    int bits[4] = {f.a, f.b, f.a + n, f.b + n};
    apply_matrix(v.data(), nbits, bits, 4, f.s, par);
    f.active = false;
  };
  auto fuse_1q = [&](int q, const cd* s4) -> bool {  // absorb into the pending pair if it involves q
    if (!f.active || (q != f.a && q != f.b)) return false;
    cd e[256];
    embed_1q(s4, q == f.a ? 0 : 1, e);
    mat16_mul(e, f.s, f.s);
    return true;
  };
  auto fuse_2q = [&](int q0, int q1, const cd* s16in) {
    cd s[256];
    std::memcpy(s, s16in, sizeof s);
    if (f.active && ((f.a == q0 && f.b == q1) || (f.a == q1 && f.b == q0))) {
      if (f.a != q0) swap_pair_order(s);
      mat16_mul(s, f.s, f.s);
      return;
    }
    flush();
    f.active = true; f.a = q0; f.b = q1;
    std::memcpy(f.s, s, sizeof s);
  };
  for (int64_t g = b.op_offsets[c]; g < b.op_offsets[c + 1]; ++g) {
    const bwq_op& op = b.ops[g];
    if (!is2q(op.opcode)) {
      const int q = cq.pos[op.q0];
      cd s4[16];
      bool have_s4 = false;
      if (op.opcode == BWQ_G_RESET) {
        for (int i = 0; i < 16; ++i) s4[i] = 0;
        s4[0] = 1; s4[3] = 1; have_s4 = true;
      }
      cd u[4];
      if (!have_s4 && !unitary1(g, u)) return BWQ_CIRC_BAD_OP;
      const cd* ns = nl.find(op.opcode, op.q0, 255);
      if (fuse && f.active && (q == f.a || q == f.b)) {
        if (!have_s4) superop_from_unitary(u, 2, s4);
        fuse_1q(q, s4);
        if (ns) fuse_1q(q, ns);
        continue;
      }
      if (have_s4) { int bits[2] = {q, q + n}; apply_matrix(v.data(), nbits, bits, 2, s4, par); }
      else if (op.opcode == BWQ_G_X) { apply_x_bit(v.data(), nbits, q, par); apply_x_bit(v.data(), nbits, q + n, par); }
      else if (op.opcode == BWQ_G_RZ || op.opcode == BWQ_G_P || op.opcode == BWQ_G_Z || op.opcode == BWQ_G_S ||
               op.opcode == BWQ_G_SDG || op.opcode == BWQ_G_T || op.opcode == BWQ_G_TDG) {
        apply_diag1(v.data(), nbits, q, u[0], u[3], par);
        apply_diag1(v.data(), nbits, q + n, std::conj(u[0]), std::conj(u[3]), par);
      } else if (op.opcode != BWQ_G_ID) {
        superop_from_unitary(u, 2, s4);
        int bits[2] = {q, q + n};
        apply_matrix(v.data(), nbits, bits, 2, s4, par);
      }
      if (ns) { int bits[2] = {q, q + n}; apply_matrix(v.data(), nbits, bits, 2, ns, par); }
      continue;
    }
    const int q0 = cq.pos[op.q0], q1 = cq.pos[op.q1];
    const cd* ns = nl.find(op.opcode, op.q0, op.q1);
    if (fuse) {
      cd u[16], s[256];
      if (!unitary2(g, u)) return BWQ_CIRC_BAD_OP;
      superop_from_unitary(u, 4, s);
      fuse_2q(q0, q1, s);
      if (ns) fuse_2q(q0, q1, ns);
      continue;
    }
    if (op.opcode == BWQ_G_CX) {
      apply_cx_bits(v.data(), nbits, q0, q1, par);
      apply_cx_bits(v.data(), nbits, q0 + n, q1 + n, par);
    } else {
      cd u[16], s[256];
      if (!unitary2(g, u)) return BWQ_CIRC_BAD_OP;
      superop_from_unitary(u, 4, s);
      int bits[4] = {q0, q1, q0 + n, q1 + n};
      apply_matrix(v.data(), nbits, bits, 4, s, par);
    }
    if (ns) { int bits[4] = {q0, q1, q0 + n, q1 + n}; apply_matrix(v.data(), nbits, bits, 4, ns, par); }
  }
  flush();
  // expectation values: Tr(rho P) = i^ny sum_c (-1)^popc(c&z) rho[c, c^x]
  const int64_t dim = int64_t(1) << n;
  for (int64_t o = b.obs_offsets[c]; o < b.obs_offsets[c + 1]; ++o) {
    double total = 0;
    for (int64_t t = b.term_offsets[o]; t < b.term_offsets[o + 1]; ++t) {
      uint64_t x = 0, z = 0;
      int ny = 0;
      bool dead = false;
      for (int q = 0; q < b.n_qubits[c]; ++q) {
        int xb = (b.term_x[t] >> q) & 1, zb = (b.term_z[t] >> q) & 1;
        if (!xb && !zb) continue;
        if (cq.pos[q] < 0) { if (xb) dead = true; continue; }
        if (xb) x |= 1ull << cq.pos[q];
        if (zb) z |= 1ull << cq.pos[q];
        if (xb && zb) ++ny;
      }
      if (dead) continue;
      double re = 0, im = 0;
#pragma omp parallel for if (par) reduction(+ : re, im) schedule(static)
      for (int64_t ci = 0; ci < dim; ++ci) {
        cd a = v[ci + ((ci ^ (int64_t)x) << n)];
        if (__builtin_popcountll((uint64_t)ci & z) & 1) a = -a;
        re += a.real(); im += a.imag();
      }
      cd val = cd(re, im);
      for (int k = 0; k < (ny & 3); ++k) val *= I_;
      total += b.term_coeff[t] * val.real();
    }
    out[o] = total;
  }
  return 0;
}

int sv_circuit(const bwq_batch& b, int c, double* out, bool par) {
  Compact cq = compact_qubits(b, c);
  const int n = cq.n;
  if (n > 30) return BWQ_CIRC_TOO_WIDE;
  std::vector<cd> v(size_t(1) << n, cd(0));
  v[0] = 1;
  for (int64_t g = b.op_offsets[c]; g < b.op_offsets[c + 1]; ++g) {
    const bwq_op& op = b.ops[g];
    if (!is2q(op.opcode)) {
      cd u[4];
      if (op.opcode == BWQ_G_RESET || !unitary1(g, u)) return BWQ_CIRC_BAD_OP;
      int bits[1] = {cq.pos[op.q0]};
      if (op.opcode == BWQ_G_X) apply_x_bit(v.data(), n, bits[0], par);
      else if (u[1] == cd(0) && u[2] == cd(0)) apply_diag1(v.data(), n, bits[0], u[0], u[3], par);
      else apply_matrix(v.data(), n, bits, 1, u, par);
    } else if (op.opcode == BWQ_G_CX) {
      apply_cx_bits(v.data(), n, cq.pos[op.q0], cq.pos[op.q1], par);
    } else {
      cd u[16];
      if (!unitary2(g, u)) return BWQ_CIRC_BAD_OP;
      int bits[2] = {cq.pos[op.q0], cq.pos[op.q1]};
      apply_matrix(v.data(), n, bits, 2, u, par);
    }
  }
  const int64_t dim = int64_t(1) << n;
  for (int64_t o = b.obs_offsets[c]; o < b.obs_offsets[c + 1]; ++o) {
    double total = 0;
    for (int64_t t = b.term_offsets[o]; t < b.term_offsets[o + 1]; ++t) {
      uint64_t x = 0, z = 0;
      int ny = 0;
      bool dead = false;
      for (int q = 0; q < b.n_qubits[c]; ++q) {
        int xb = (b.term_x[t] >> q) & 1, zb = (b.term_z[t] >> q) & 1;
        if (!xb && !zb) continue;
        if (cq.pos[q] < 0) { if (xb) dead = true; continue; }
        if (xb) x |= 1ull << cq.pos[q];
        if (zb) z |= 1ull << cq.pos[q];
        if (xb && zb) ++ny;
      }
      if (dead) continue;
      double re = 0, im = 0;
#pragma omp parallel for if (par) reduction(+ : re, im) schedule(static)
      for (int64_t ci = 0; ci < dim; ++ci) {
        cd a = std::conj(v[ci ^ (int64_t)x]) * v[ci];
        if (__builtin_popcountll((uint64_t)ci & z) & 1) a = -a;
        re += a.real(); im += a.imag();
      }
      cd val = cd(re, im);
      for (int k = 0; k < (ny & 3); ++k) val *= I_;
      total += b.term_coeff[t] * val.real();
    }
    out[o] = total;
  }
  return 0;
}

}  // namespace

extern "C" {

// threads <= 0 -> omp_get_max_threads().  amplitude_parallel_qubits: states with at least this many
// active qubits are parallelised over amplitudes, smaller ones over circuits (Aer: experiments).
int cpuref_dm_run(const bwq_batch* b, const cpuref_noise* noise, const double* gate_mats, const int64_t* gate_off, double* out,
                  int32_t* status, int threads, int amplitude_parallel_qubits, int fusion_threshold) {
  const GateTable tab{gate_mats, gate_off};
  g_tab = &tab;
  if (threads <= 0) threads = omp_get_max_threads();
  omp_set_num_threads(threads);
  if (amplitude_parallel_qubits <= 0) amplitude_parallel_qubits = 7;
  if (fusion_threshold <= 0) fusion_threshold = 7;
  std::vector<int> small, big;
  for (int c = 0; c < b->n_circuits; ++c) (compact_qubits(*b, c).n >= amplitude_parallel_qubits ? big : small).push_back(c);
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t i = 0; i < small.size(); ++i) status[small[i]] = dm_circuit(*b, small[i], noise, out, false, fusion_threshold);
  for (int c : big) status[c] = dm_circuit(*b, c, noise, out, true, fusion_threshold);
  g_tab = nullptr;
  return 0;
}

int cpuref_sv_run(const bwq_batch* b, const double* gate_mats, const int64_t* gate_off, double* out, int32_t* status, int threads,
                  int amplitude_parallel_qubits) {
  const GateTable tab{gate_mats, gate_off};
  g_tab = &tab;
  if (threads <= 0) threads = omp_get_max_threads();
  omp_set_num_threads(threads);
  if (amplitude_parallel_qubits <= 0) amplitude_parallel_qubits = 14;
  std::vector<int> small, big;
  for (int c = 0; c < b->n_circuits; ++c) (compact_qubits(*b, c).n >= amplitude_parallel_qubits ? big : small).push_back(c);
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t i = 0; i < small.size(); ++i) status[small[i]] = sv_circuit(*b, small[i], out, false);
  for (int c : big) status[c] = sv_circuit(*b, c, out, true);
  g_tab = nullptr;
  return 0;
}

int cpuref_max_threads(void) { return omp_get_max_threads(); }
}
