// dm_onchip_kernel -- K0..K4 in one launch for circuits whose state fits on chip (<= 5 active qubits:
// 4^5 Pauli-basis elements = 8 KiB of shared memory).
//
// The reference's smallest configurations (cfg1 of BASELINE.json: 4-qubit Trotter circuits on
// FakeLima, ZNE folds 1/3/5 -- docs/tutorials/zne_parallel.py:168-189) are thousands of tiny
// circuits: the sweeps take 0.3 ms while the host-side lowering of the variants takes 10 ms.  Here
// one WARP interprets the flat gate stream of one (circuit, variant) directly -- no lowered program,
// no variant expansion on the host, the upload is the 8-byte ops of the BASE circuits only:
//   scan    : active qubits -> digits, validity (same status codes as lower_dm_circuit)
//   1q gate : the gate's transfer matrix (x its error) is multiplied into the qubit's pending 4x4
//             by lanes 0..15 (one element each; products through shuffles) -- lowering.cpp's fusion
//   cx      : the 4^(n-2) register groups of the pair are loaded from the warp's shared-memory
//             state, the pending maps of both qubits and `fold` copies of cx (+ its error) applied
//             (the register ops of kernels.cuh), and stored back
//   twirls  : the Pauli pair of (seed, circuit, twirl, cx index) is drawn with the generator of
//             variants.cpp and enters as ordinary 1-qubit gates (with their errors)
//   values  : sum_k c_k rho[index(P_k)], one warp reduction per observable
// Circuits it does not cover (2-qubit gates other than cx, more than 5 active qubits) report
// kOnchipNotHandled and the caller runs the batch through the tile-sweep path instead.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace bwq {

#ifndef BWQ_ONCHIP_MAXD
#define BWQ_ONCHIP_MAXD 5
#endif
#ifndef BWQ_ONCHIP_MINB
#define BWQ_ONCHIP_MINB 16
#endif
constexpr int kOnchipMaxDigits = BWQ_ONCHIP_MAXD;
#ifndef BWQ_ONCHIP_WARPS
#define BWQ_ONCHIP_WARPS 1
#endif
constexpr int kOnchipWarps = BWQ_ONCHIP_WARPS;  // warps (circuits) per CTA; 1: a finished circuit frees its slot at once
constexpr int kOnchipNotHandled = 100;     // internal status: not an on-chip circuit

struct OnchipNoise {            // all null / 0 for the ideal evolution
  const int32_t* g1;            // [33][64] entry of (1-qubit opcode < 32 | row 32 = unitary1, qubit), -1 = none
  const int32_t* cx;            // [64][64] entry of cx (control, target), -1 = none
  const int2* ent;              // [n] {kind, offset into data (doubles)}
  const uint8_t* fixed_ok;      // [32*64] error x gate tabulated
  const double* fixed_ptm;      // [32*64][16]
  const double* data;
};

struct OnchipLaunch {
  int32_t n_circuits, n_folds, n_twirls, twirl;   // n_folds, n_twirls >= 1 here; twirl = draw Paulis
  int32_t circuit_base;         // batch index of this launch's circuit 0 (a launch covers a range of the batch; twirl draws use the batch index)
  uint64_t seed;
  const int32_t* folds;         // [n_folds] or null (factor 1)
  const int32_t* n_qubits;
  const int64_t* op_offsets;
  const bwq_op* ops;
  const double* params;
  int64_t n_params;
  const int64_t* obs_offsets;
  const int64_t* term_offsets;
  const uint64_t* term_x;
  const uint64_t* term_z;
  const double* term_coeff;
  OnchipNoise noise;
  double* out;                  // [(obs of circuit c) x variants]: circuit-major, variant, observable
  int32_t* status;              // [n_circuits * n_variants]
  // with_ideal: one more warp per circuit evolves the base circuit WITHOUT the noise table (the
  // ideal value of every variant: folds and twirls are identities) -> out_ideal [obs], status_ideal [n_circuits]
  int32_t with_ideal;
  double* out_ideal;
  int32_t* status_ideal;
  // statevector semantics for the whole launch (bwq_sv_run): `reset` is not a unitary -- the
  // statevector path reports BWQ_CIRC_BAD_OP for it (sv_lowering.cpp), and so do the ideal warps
  int32_t sv_mode;
};

__device__ __forceinline__ int dev_num_params(uint32_t op) {
  switch (op) {
    case BWQ_G_RX: case BWQ_G_RY: case BWQ_G_RZ: case BWQ_G_P: return 1;
    case BWQ_G_U2: return 2;
    case BWQ_G_U3: return 3;
    case BWQ_G_UNITARY1: return 8;
    default: return 0;
  }
}
__device__ __forceinline__ bool dev_is_1q(uint32_t op) { return op <= BWQ_G_RESET || op == BWQ_G_UNITARY1; }


__device__ __forceinline__ double2 cmulc(double2 a, double2 b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a conj(b)
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

// 2x2 unitary of a 1-qubit gate (row-major), the table of lowering.cpp:unitary1
__device__ __forceinline__ void dev_unitary1(uint32_t op, double p0, const double* __restrict__ pp, double2 (&u)[4]) {
  const double r = 0.70710678118654757;  // sqrt(0.5) as the host computes it
  u[0] = make_double2(1, 0); u[1] = make_double2(0, 0); u[2] = make_double2(0, 0); u[3] = make_double2(1, 0);
  switch (op) {
    case BWQ_G_X: u[0].x = 0; u[1].x = 1; u[2].x = 1; u[3].x = 0; break;
    case BWQ_G_Z: u[3].x = -1; break;
    case BWQ_G_SX: u[0] = make_double2(.5, .5); u[1] = make_double2(.5, -.5); u[2] = make_double2(.5, -.5); u[3] = make_double2(.5, .5); break;
    case BWQ_G_RZ: { double s, c; sincos(0.5 * p0, &s, &c); u[0] = make_double2(c, -s); u[3] = make_double2(c, s); break; }
    case BWQ_G_P: { double s, c; sincos(p0, &s, &c); u[3] = make_double2(c, s); break; }
    case BWQ_G_Y: u[0].x = 0; u[1] = make_double2(0, -1); u[2] = make_double2(0, 1); u[3].x = 0; break;
    case BWQ_G_H: u[0].x = r; u[1].x = r; u[2].x = r; u[3].x = -r; break;
    case BWQ_G_S: u[3] = make_double2(0, 1); break;
    case BWQ_G_SDG: u[3] = make_double2(0, -1); break;
    case BWQ_G_T: { double s, c; sincos(0.78539816339744828, &s, &c); u[3] = make_double2(c, s); break; }
    case BWQ_G_TDG: { double s, c; sincos(0.78539816339744828, &s, &c); u[3] = make_double2(c, -s); break; }
    case BWQ_G_SXDG: u[0] = make_double2(.5, -.5); u[1] = make_double2(.5, .5); u[2] = make_double2(.5, .5); u[3] = make_double2(.5, -.5); break;
    case BWQ_G_RX: { double s, c; sincos(0.5 * p0, &s, &c); u[0].x = c; u[1] = make_double2(0, -s); u[2] = make_double2(0, -s); u[3].x = c; break; }
    case BWQ_G_RY: { double s, c; sincos(0.5 * p0, &s, &c); u[0].x = c; u[1].x = -s; u[2].x = s; u[3].x = c; break; }
    case BWQ_G_U2: case BWQ_G_U3: {
      const double th = op == BWQ_G_U2 ? 1.5707963267948966 : p0;
      const double ph = op == BWQ_G_U2 ? p0 : pp[1], la = op == BWQ_G_U2 ? pp[1] : pp[2];
      double s, c, sl, cl, sp, cp, spl, cpl;
      sincos(0.5 * th, &s, &c); sincos(la, &sl, &cl); sincos(ph, &sp, &cp); sincos(ph + la, &spl, &cpl);
      u[0] = make_double2(c, 0); u[1] = make_double2(-cl * s, -sl * s); u[2] = make_double2(cp * s, sp * s); u[3] = make_double2(cpl * c, spl * c);
      break; }
    case BWQ_G_UNITARY1:
#pragma unroll
      for (int i = 0; i < 4; ++i) u[i] = make_double2(pp[2 * i], pp[2 * i + 1]);
      break;
    default: break;  // id (x, z, sx, rz, p, reset have closed-form transfer matrices below)
  }
}

// the whole 4x4 (row-major) -- one lane computes the matrix of its own op of a 32-op fetch
__device__ __forceinline__ void dev_gate_ptm_full(uint32_t op, double p0, const double* __restrict__ pp, double (&r)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = 0.0;
  switch (op) {
    case BWQ_G_ID: r[0] = 1; r[5] = 1; r[10] = 1; r[15] = 1; return;
    case BWQ_G_X: r[0] = 1; r[5] = 1; r[10] = -1; r[15] = -1; return;
    case BWQ_G_Z: r[0] = 1; r[5] = -1; r[10] = -1; r[15] = 1; return;
    case BWQ_G_RZ: case BWQ_G_P: {
      double s, c;
      sincos(p0, &s, &c);
      r[0] = 1; r[5] = c; r[6] = -s; r[9] = s; r[10] = c; r[15] = 1; return; }
    case BWQ_G_SX: r[0] = 1; r[5] = 1; r[11] = -1; r[14] = 1; return;
    case BWQ_G_RESET: r[0] = 1; r[12] = 1; return;
    default: break;
  }
  double2 u[4];
  dev_unitary1(op, p0, pp, u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double2 t[4], e[4];
    if (j == 0) { t[0] = u[0]; t[1] = u[1]; t[2] = u[2]; t[3] = u[3]; }
    else if (j == 1) { t[0] = u[1]; t[1] = u[0]; t[2] = u[3]; t[3] = u[2]; }
    else if (j == 2) {
      t[0] = make_double2(-u[1].y, u[1].x); t[1] = make_double2(u[0].y, -u[0].x);
      t[2] = make_double2(-u[3].y, u[3].x); t[3] = make_double2(u[2].y, -u[2].x);
    } else { t[0] = u[0]; t[1] = make_double2(-u[1].x, -u[1].y); t[2] = u[2]; t[3] = make_double2(-u[3].x, -u[3].y); }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) e[a * 2 + b] = cadd(cmulc(t[a * 2 + 0], u[b * 2 + 0]), cmulc(t[a * 2 + 1], u[b * 2 + 1]));
    r[0 + j] = 0.5 * (e[0].x + e[3].x);
    r[4 + j] = 0.5 * (e[1].x + e[2].x);
    r[8 + j] = 0.5 * (e[2].y - e[1].y);
    r[12 + j] = 0.5 * (e[0].x - e[3].x);
  }
}

__device__ __forceinline__ uint64_t dev_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

constexpr int kOnchipGRow = 18;  // doubles per staged gate matrix (16 + pad: conflict-free 16-byte rows)
constexpr int kOnchipWarpDoubles = (1 << (2 * kOnchipMaxDigits)) + kOnchipMaxDigits * 16 + 32 * kOnchipGRow;
constexpr size_t kOnchipSmem = sizeof(double) * kOnchipWarpDoubles * kOnchipWarps;

// Bank swizzle of the warp's state: element idx (digits d0..d4) lives at idx ^ S(idx), where S folds
// the upper digits into the low four index bits (= the 8-byte bank) through a spread of GF(2)^4:
//   bits 0..1 ^= d2 ^ d3 ^ d4,   bits 2..3 ^= d2 ^ w(d3) ^ w(w(d4)),   w = multiplication by the GF(4) generator.
// The five 2-bit subspaces are pairwise complementary, so the 16 register groups of ANY digit pair
// (lanes enumerate the two other digits of a 4-qubit state) touch 16 distinct banks for each of
// their 16 elements; S is linear over XOR, so address(x | a << 2da | b << 2db) = P(x) ^ P(a << 2da) ^ P(b << 2db).
__host__ __device__ constexpr uint32_t oc_w(uint32_t x) { return ((x >> 1) & 1u) | (((x ^ (x >> 1)) & 1u) << 1); }
__host__ __device__ constexpr uint32_t oc_nib(int d, uint32_t k) {  // S of digit d = k (d = 2, 3, 4)
  return d == 2 ? (k | (k << 2)) : d == 3 ? (k | (oc_w(k) << 2)) : (k | ((k ^ oc_w(k)) << 2));
}
__host__ __device__ constexpr uint64_t oc_table() {  // nibble 4 * (d - 2) * 4 + k ... : 16 bits per digit
  uint64_t t = 0;
  for (int d = 2; d <= 4; ++d)
    for (uint32_t k = 0; k < 4; ++k) t |= (uint64_t)oc_nib(d, k) << (16 * (d - 2) + 4 * k);
  return t;
}
constexpr uint64_t kOcTable = oc_table();
__device__ __forceinline__ uint32_t oc_phys(uint32_t idx) {
  const uint32_t hi = idx >> 4;  // digits 2..4, 2 bits each -> nibble index 4 * (d - 2) + k: shift 16 (d - 2) + 4 k
  const uint32_t s2 = (uint32_t)(kOcTable >> ((hi & 3u) << 2)), s3 = (uint32_t)(kOcTable >> (16u + (((hi >> 2) & 3u) << 2)));
  const uint32_t s4 = (uint32_t)(kOcTable >> (32u + (((hi >> 4) & 3u) << 2)));
  return idx ^ ((s2 ^ s3 ^ s4) & 15u);
}
// swizzled image of digit d = k alone (a basis vector of the XOR-linear map)
__device__ __forceinline__ uint32_t oc_phys_digit(int d, uint32_t k) {
  const uint32_t sw = d < 2 ? 0u : (uint32_t)(kOcTable >> (16 * (d - 2) + 4 * k)) & 15u;
  return (k << (2 * d)) ^ sw;
}

struct OnchipWarp {
  double* st;      // 4^n state of this warp (swizzled: oc_phys)
  double* pend;    // [kOnchipMaxDigits][16] pending 1-qubit maps
  double* gbuf;    // [32][kOnchipGRow] transfer matrices of the fetched ops
  uint32_t has;    // digits with a pending map (warp-uniform)
  int nd, lane;
};

// error x gate of a 1-qubit op on physical qubit q, whole matrix (lane-private)
__device__ __forceinline__ void onchip_gate_matrix(const OnchipNoise& N, uint32_t op, int q, double p0, const double* __restrict__ pp,
                                                   double (&g)[16]) {
  int e = -1;
  if (N.g1 != nullptr) e = __ldg(N.g1 + (op < 32u ? op : 32u) * 64 + q);
  if (e >= 0 && op < 32u && __ldg(N.fixed_ok + op * 64 + q)) {
    const double2* f = reinterpret_cast<const double2*>(N.fixed_ptm + (size_t)(op * 64 + q) * 16);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const double2 x = __ldg(f + i); g[2 * i] = x.x; g[2 * i + 1] = x.y; }
    return;
  }
  if (e < 0) { dev_gate_ptm_full(op, p0, pp, g); return; }
  double r[16];
  dev_gate_ptm_full(op, p0, pp, r);
  const double2* nm = reinterpret_cast<const double2*>(N.data + __ldg(&N.ent[e]).y);  // error after the gate: N * R
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2 n01 = __ldg(nm + 2 * i), n23 = __ldg(nm + 2 * i + 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) g[4 * i + j] = ((n01.x * r[j] + n01.y * r[4 + j]) + n23.x * r[8 + j]) + n23.y * r[12 + j];
  }
}

// pend[d] <- G * pend[d], G row-major in shared memory; lanes 0..15 own one element each
__device__ __forceinline__ void onchip_push1(OnchipWarp& W, int d, const double* __restrict__ G) {
  const int l = W.lane & 15, i = l >> 2, j = l & 3;
  double* P = W.pend + 16 * d;
  double r;
  if ((W.has >> d) & 1u) {
    const double2 g01 = *reinterpret_cast<const double2*>(G + 4 * i), g23 = *reinterpret_cast<const double2*>(G + 4 * i + 2);
    r = ((g01.x * P[j] + g01.y * P[4 + j]) + g23.x * P[8 + j]) + g23.y * P[12 + j];  // the summation order of lowering.cpp:mat4_mul
    __syncwarp();
  } else {
    r = G[l];
  }
  if (W.lane < 16) P[l] = r;
  W.has |= 1u << d;
  __syncwarp();
}

// a gate outside the fetched stream (twirl Paulis): lane 0 computes the matrix into row 0 of a scratch
__device__ __forceinline__ void onchip_gate1(OnchipWarp& W, const OnchipNoise& N, double* scratch, uint32_t op, int q, int d, double p0) {
  if (W.lane == 0) {
    double g[16];
    onchip_gate_matrix(N, op, q, p0, nullptr, g);
#pragma unroll
    for (int i = 0; i < 16; ++i) scratch[i] = g[i];
  }
  __syncwarp();
  onchip_push1(W, d, scratch);
}

// Pauli of a twirl in the backend basis (variants.cpp): Y = rz(pi) x, Z = rz(pi), X = x
__device__ __forceinline__ void onchip_pauli(OnchipWarp& W, const OnchipNoise& N, double* scratch, int p, int q, int d) {
  if (p == 2 || p == 3) onchip_gate1(W, N, scratch, BWQ_G_RZ, q, d, 3.14159265358979323846);
  if (p == 1 || p == 2) onchip_gate1(W, N, scratch, BWQ_G_X, q, d, 0.0);
}

// one cx (+ its error of `kind` at nm) on the register group
__device__ __forceinline__ void onchip_cx(double (&v)[1][16], int kind, const double* __restrict__ nm) {
  if (kind == BWQ_NOISE_RELAX2) op_relax2<false, true, 1>(v, nm);
  else {
    op_cx<false, 1>(v);
    if (kind == BWQ_NOISE_DENSE2) op_dense2<false, 1>(v, nm);
  }
}

// register groups of the digit pair (da, db): pending maps, then `reps` copies of cx (+ its error);
// reps = 0: only the pending map of da (end of circuit)
__device__ __forceinline__ void onchip_pair(OnchipWarp& W, int da, int db, int reps, int kind, const double* __restrict__ nm) {
  const int lo = min(da, db), hi = max(da, db);
  const int groups = 1 << (2 * (W.nd - 2));
  const bool pa = (W.has >> da) & 1u, pb = reps > 0 && ((W.has >> db) & 1u);
  uint32_t oa[4], ob[4];  // byte offsets
#pragma unroll
  for (int k = 0; k < 4; ++k) { oa[k] = oc_phys_digit(da, (uint32_t)k) << 3; ob[k] = oc_phys_digit(db, (uint32_t)k) << 3; }
  char* const base = reinterpret_cast<char*>(W.st);
  for (int g = W.lane; g < groups; g += 32) {
    uint32_t x = (uint32_t)g;
    x = ((x >> (2 * lo)) << (2 * lo + 2)) | (x & ((1u << (2 * lo)) - 1u));
    x = ((x >> (2 * hi)) << (2 * hi + 2)) | (x & ((1u << (2 * hi)) - 1u));
    const uint32_t px = oc_phys(x) << 3;
    double v[1][16];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int a = 0; a < 4; ++a) v[0][a + 4 * b] = *reinterpret_cast<const double*>(base + (px ^ oa[a] ^ ob[b]));
    if (pa) op_dense1<false, 1>(v, W.pend + 16 * da);
    if (pb) op_dense1<true, 1>(v, W.pend + 16 * db);
    // folds are odd: one application, then pairs (the cx permutation is an involution, so a pair
    // returns the register assignment to where it started: no moves at the loop edge)
    if (reps & 1) onchip_cx(v, kind, nm);
    for (int r = 0; r < (reps >> 1); ++r) { onchip_cx(v, kind, nm); onchip_cx(v, kind, nm); }
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int a = 0; a < 4; ++a) *reinterpret_cast<double*>(base + (px ^ oa[a] ^ ob[b])) = v[0][a + 4 * b];
  }
  W.has &= ~(1u << da);
  if (reps > 0) W.has &= ~(1u << db);
  __syncwarp();
}

// Ideal side (and bwq_sv_run on small circuits): the noise-free state is pure, so the warp evolves
// the 2^n <= 32 complex amplitudes instead of 4^n Pauli coefficients -- lane x owns amplitude x, a
// 1-qubit gate is one butterfly through a shuffle, cx a conditional shuffle, <P> one partner
// shuffle per Pauli term.  The 2x2 unitaries of a fetch of 32 ops are prepared 32 wide (one lane
// per op) into the gate-matrix rows of shared memory, as on the noisy side.
__device__ __forceinline__ void onchip_sv_warp(const OnchipLaunch& L, double* __restrict__ gbuf, int lane, uint64_t used,
                                               int64_t g0, int64_t g1, int64_t o0, int64_t o1, double* __restrict__ out) {
  auto digit_of = [&](int q) { return __popcll(used & ((1ull << q) - 1ull)); };
  double2 psi = make_double2(lane == 0 ? 1.0 : 0.0, 0.0);
  for (int64_t gb = g0; gb < g1; gb += 32) {
    uint32_t my_dig = 0u;  // opcode | digit(q0) << 16 | digit(q1) << 24
    if (gb + lane < g1) {
      const unsigned long long raw = __ldg(reinterpret_cast<const unsigned long long*>(L.ops) + gb + lane);
      const uint32_t opc = (uint32_t)(raw & 0xffffu);
      const int q0 = (int)((raw >> 16) & 0xffu), q1 = (int)((raw >> 24) & 0xffu);
      my_dig = opc | ((uint32_t)digit_of(q0) << 16) | ((uint32_t)digit_of(q1 & 63) << 24);
      if (opc != BWQ_G_CX) {
        const double* pp = L.params + (uint32_t)(raw >> 32);
        double2 u[4];
        dev_unitary1(opc, dev_num_params(opc) > 0 ? __ldg(pp) : 0.0, pp, u);
        double2* dst = reinterpret_cast<double2*>(gbuf + lane * kOnchipGRow);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = u[i];
      }
    }
    __syncwarp();
    const int cnt = (int)min((int64_t)32, g1 - gb);
    for (int k = 0; k < cnt; ++k) {
      const uint32_t dig = __shfl_sync(0xffffffffu, my_dig, k);
      const int da = (int)((dig >> 16) & 0xffu), db = (int)(dig >> 24);
      if ((dig & 0xffffu) == BWQ_G_CX) {  // control = q0: amplitudes with the control bit set swap along the target bit
        const double vx = __shfl_xor_sync(0xffffffffu, psi.x, 1 << db), vy = __shfl_xor_sync(0xffffffffu, psi.y, 1 << db);
        if ((lane >> da) & 1) psi = make_double2(vx, vy);
        continue;
      }
      const double2* u = reinterpret_cast<const double2*>(gbuf + k * kOnchipGRow);
      const double2 p = make_double2(__shfl_xor_sync(0xffffffffu, psi.x, 1 << da), __shfl_xor_sync(0xffffffffu, psi.y, 1 << da));
      const bool hi = (lane >> da) & 1;
      const double2 ua = hi ? u[3] : u[0], ub = hi ? u[2] : u[1];  // new = u[bit][bit] * mine + u[bit][1 - bit] * partner
      psi = make_double2(ua.x * psi.x - ua.y * psi.y + ub.x * p.x - ub.y * p.y, ua.x * psi.y + ua.y * psi.x + ub.x * p.y + ub.y * p.x);
    }
    __syncwarp();
  }
  // <P> = sum_j conj(psi_j) phase(j ^ xm) psi_(j ^ xm), phase(k) = i^(#Y) (-1)^popc(k & zm)
  for (int64_t o = o0; o < o1; ++o) {
    const int64_t t0 = __ldg(L.term_offsets + o), t1 = __ldg(L.term_offsets + o + 1);
    double acc = 0.0;
    for (int64_t t = t0; t < t1; ++t) {
      const uint64_t x = __ldg(L.term_x + t), z = __ldg(L.term_z + t);
      uint32_t xm = 0, zm = 0;
      uint64_t m = (x | z) & used;
      while (m) {
        const int q = __ffsll((long long)m) - 1;
        m &= m - 1;
        const int d = digit_of(q);
        xm |= (uint32_t)((x >> q) & 1ull) << d;
        zm |= (uint32_t)((z >> q) & 1ull) << d;
      }
      const double2 p = make_double2(__shfl_xor_sync(0xffffffffu, psi.x, xm), __shfl_xor_sync(0xffffffffu, psi.y, xm));
      if ((x & ~used) != 0ull) continue;  // X / Y on an idle qubit (warp-uniform)
      const int ny = __popc(xm & zm) & 3;
      const double sgn = (__popc((uint32_t)(lane ^ xm) & zm) & 1) ? -1.0 : 1.0;
      // conj(psi) * p = (re, im); times i^ny: real part is re, -im, -re, im for ny = 0..3
      const double re = psi.x * p.x + psi.y * p.y, im = psi.x * p.y - psi.y * p.x;
      const double v = ny == 0 ? re : ny == 1 ? -im : ny == 2 ? -re : im;
      acc += __ldg(L.term_coeff + t) * sgn * v;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) out[o - o0] = acc;
  }
}

__global__ void __launch_bounds__(32 * kOnchipWarps, BWQ_ONCHIP_MINB / kOnchipWarps) dm_onchip_kernel(const OnchipLaunch L) {
  extern __shared__ __align__(16) double s_dyn[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_var = L.n_folds * L.n_twirls;
  const int64_t w = (int64_t)blockIdx.x * kOnchipWarps + warp;
  if (w >= (int64_t)L.n_circuits * (n_var + L.with_ideal)) return;
  // variant-major launch order, last variant first: the folds are ascending, so the longest
  // circuits (highest noise factor) start first and the short ones (the ideal side last) fill the tail
  const int c = (int)(w % L.n_circuits);
  const int v_rev = (int)(w / L.n_circuits);
  const bool ideal = v_rev == n_var;
  const int vi = ideal ? 0 : n_var - 1 - v_rev;
  const int fac = (L.folds && !ideal) ? __ldg(L.folds + vi / L.n_twirls) : 1;
  const int tw = vi % L.n_twirls;
  const bool twirl = L.twirl && !ideal;
  const int nq = __ldg(L.n_qubits + c);
  const int64_t g0 = __ldg(L.op_offsets + c), g1 = __ldg(L.op_offsets + c + 1);
  const int64_t o0 = __ldg(L.obs_offsets + c), o1 = __ldg(L.obs_offsets + c + 1);
  double* out = ideal ? L.out_ideal + o0 : L.out + o0 * n_var + (int64_t)vi * (o1 - o0);

  // ---- scan: active qubits and validity (status precedence of lower_dm_circuit)
  uint64_t used = 0;
  uint32_t bad = 0;  // 1: bad qubit, 2: bad op, 4: not an on-chip circuit
  if (nq < 0 || nq > 64) bad |= 1u;
  for (int64_t g = g0 + lane; g < g1 && !(bad & 1u); g += 32) {
    const bwq_op op = L.ops[g];
    const bool one = dev_is_1q(op.opcode);
    if (!one && op.opcode != BWQ_G_CX) { bad |= (op.opcode > BWQ_G_ECR && op.opcode != BWQ_G_UNITARY2) ? 2u : 4u; }
    const bool two = !one && ((op.opcode >= BWQ_G_CX && op.opcode <= BWQ_G_ECR) || op.opcode == BWQ_G_UNITARY2);
    if ((int)op.q0 >= nq || (two && ((int)op.q1 >= nq || op.q1 == op.q0))) { bad |= 1u; break; }
    used |= 1ull << op.q0;
    if (two) used |= 1ull << op.q1;
    const int np = one ? dev_num_params(op.opcode) : 0;
    if (np && (int64_t)op.param_idx + np > L.n_params) bad |= 2u;
    if (op.opcode == BWQ_G_RESET && (ideal || L.sv_mode)) bad |= 2u;
  }
  {
    const uint64_t valid = nq >= 64 ? ~0ull : ((1ull << max(nq, 0)) - 1ull);
    const int64_t t0 = __ldg(L.term_offsets + o0), t1 = __ldg(L.term_offsets + o1);
    for (int64_t t = t0 + lane; t < t1; t += 32)
      if ((__ldg(L.term_x + t) | __ldg(L.term_z + t)) & ~valid) bad |= 1u;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    used |= __shfl_xor_sync(0xffffffffu, used, s);
    bad |= __shfl_xor_sync(0xffffffffu, bad, s);
  }
  const int n_active = __popcll(used);
  int status = 0;
  if (bad & 1u) status = BWQ_CIRC_BAD_QUBIT;
  else if (n_active > kOnchipMaxDigits || (bad & 4u)) status = kOnchipNotHandled;
  else if (bad & 2u) status = BWQ_CIRC_BAD_OP;
  if (lane == 0) { if (ideal) L.status_ideal[c] = status; else L.status[(int64_t)c * n_var + vi] = status; }
  if (status) {
    for (int64_t o = lane; o < o1 - o0; o += 32) out[o] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  if (ideal || L.sv_mode) {  // pure state: 2^n amplitudes in registers
    onchip_sv_warp(L, s_dyn + warp * kOnchipWarpDoubles + (1 << (2 * kOnchipMaxDigits)) + kOnchipMaxDigits * 16, lane, used, g0, g1, o0, o1, out);
    return;
  }
  OnchipWarp W;
  W.st = s_dyn + warp * kOnchipWarpDoubles;
  W.pend = W.st + (1 << (2 * kOnchipMaxDigits));
  W.gbuf = W.pend + kOnchipMaxDigits * 16;
  W.has = 0u; W.lane = lane;
  W.nd = max(n_active, 2);  // padded with idle digits (I/Z = 1)
  auto digit_of = [&](int q) { return __popcll(used & ((1ull << q) - 1ull)); };

  // ---- |0..0><0..0|: 1 on the {I, Z} strings
  for (int j = lane; j < (1 << (2 * W.nd)); j += 32) W.st[oc_phys((uint32_t)j)] = ((j ^ (j >> 1)) & 0x55555555) == 0 ? 1.0 : 0.0;
  __syncwarp();

  // ---- the gate stream, 32 ops per fetch: every lane prepares ITS op (transfer matrix x error of a
  // 1-qubit gate into shared memory, the error entry of a cx into a register) -- table lookups,
  // sincos and the 4x4 products run 32 wide; the sequential part below is shared-memory work only
  uint64_t k_cx = 0;
  for (int64_t gb = g0; gb < g1; gb += 32) {
    unsigned long long my_raw = 0ull, my_aux = 0ull;
    uint32_t my_dig = 0u;  // opcode | digit(q0) << 16 | digit(q1) << 24
    if (gb + lane < g1) {
      my_raw = __ldg(reinterpret_cast<const unsigned long long*>(L.ops) + gb + lane);
      const uint32_t opc = (uint32_t)(my_raw & 0xffffu);
      const int q0 = (int)((my_raw >> 16) & 0xffu), q1 = (int)((my_raw >> 24) & 0xffu);
      my_dig = opc | ((uint32_t)digit_of(q0) << 16) | ((uint32_t)digit_of(q1 & 63) << 24);
      if (opc == BWQ_G_CX) {
        const int e = L.noise.cx ? __ldg(L.noise.cx + q0 * 64 + q1) : -1;
        if (e >= 0) { const int2 en = __ldg(&L.noise.ent[e]); my_aux = ((unsigned long long)(uint32_t)en.x << 32) | (uint32_t)en.y; }
      } else {
        const double* pp = L.params + (uint32_t)(my_raw >> 32);
        double g[16];
        onchip_gate_matrix(L.noise, opc, q0, dev_num_params(opc) > 0 ? __ldg(pp) : 0.0, pp, g);
        double2* dst = reinterpret_cast<double2*>(W.gbuf + lane * kOnchipGRow);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_double2(g[2 * i], g[2 * i + 1]);
      }
    }
    __syncwarp();
    const int cnt = (int)min((int64_t)32, g1 - gb);
    for (int k = 0; k < cnt; ++k) {
      const uint32_t dig = __shfl_sync(0xffffffffu, my_dig, k);
      const int da = (int)((dig >> 16) & 0xffu), db = (int)(dig >> 24);
      if ((dig & 0xffffu) != BWQ_G_CX) {
        onchip_push1(W, da, W.gbuf + k * kOnchipGRow);
        continue;
      }
      const unsigned long long aux = __shfl_sync(0xffffffffu, my_aux, k);
      int q0 = 0, q1 = 0;
      if (twirl) {
        const unsigned long long raw = __shfl_sync(0xffffffffu, my_raw, k);
        q0 = (int)((raw >> 16) & 0xffu); q1 = (int)((raw >> 24) & 0xffu);
      }
      // twirl Paulis are staged in the gate-matrix row of the cx itself (a cx leaves its row unused)
      double* scratch = W.gbuf + k * kOnchipGRow;
      int qc = 0, qt = 0;
      if (twirl) {
        const uint32_t d = (uint32_t)(dev_splitmix64(dev_splitmix64(dev_splitmix64(L.seed ^ (uint64_t)(c + L.circuit_base)) ^ (uint64_t)tw) ^ k_cx) & 15u);
        ++k_cx;
        const int pc = (int)(d & 3u), pt = (int)(d >> 2);
        // CX conjugation (sign dropped): X_c -> X_c X_t, Z_t -> Z_c Z_t
        const int xc = (pc == 1 || pc == 2), zc = (pc >= 2), xt = (pt == 1 || pt == 2), zt = (pt >= 2);
        const int zc2 = zc ^ zt, xt2 = xt ^ xc;
        qc = xc ? (zc2 ? 2 : 1) : (zc2 ? 3 : 0);
        qt = xt2 ? (zt ? 2 : 1) : (zt ? 3 : 0);
        onchip_pauli(W, L.noise, scratch, pc, q0, da);
        onchip_pauli(W, L.noise, scratch, pt, q1, db);
      }
      onchip_pair(W, da, db, fac, (int)(aux >> 32), L.noise.data + (uint32_t)aux);
      if (twirl) {
        onchip_pauli(W, L.noise, scratch, qc, q0, da);
        onchip_pauli(W, L.noise, scratch, qt, q1, db);
      }
    }
    __syncwarp();
  }
  // ---- pending maps left at the end of the circuit
  for (int d = 0; d < W.nd; ++d)
    if ((W.has >> d) & 1u) onchip_pair(W, d, d == 0 ? 1 : 0, 0, 0, nullptr);

  // ---- values: sum_k c_k rho[index(P_k)] (idle qubits stay |0>: <Z> = 1, <X> = <Y> = 0)
  for (int64_t o = o0; o < o1; ++o) {
    const int64_t t0 = __ldg(L.term_offsets + o), t1 = __ldg(L.term_offsets + o + 1);
    double acc = 0.0;
    for (int64_t t = t0 + lane; t < t1; t += 32) {
      const uint64_t x = __ldg(L.term_x + t), z = __ldg(L.term_z + t);
      bool zero = (x & ~used) != 0ull;
      uint32_t idx = 0;
      uint64_t m = (x | z) & used;
      while (m) {
        const int q = __ffsll((long long)m) - 1;
        m &= m - 1;
        const int xb = (int)((x >> q) & 1ull), zb = (int)((z >> q) & 1ull);
        idx += (uint32_t)(xb ? (zb ? 2 : 1) : 3) << (2 * digit_of(q));
      }
      if (!zero) acc += __ldg(L.term_coeff + t) * W.st[oc_phys(idx)];
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) out[o - o0] = acc;
  }
}

}  // namespace bwq
