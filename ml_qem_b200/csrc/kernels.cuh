// CUDA kernels of the exact expectation-value engine (sm_100a).
//
//   dm_sweep_kernel<KQ>   K1/K2/K3: one HBM read+write of every Pauli-basis density-matrix
//                         element; a CTA stages one 4^KQ-element tile in shared memory and runs
//                         every register pass (2-qubit group) the lowering packed into the sweep.
//   dm_expval_kernel      K4: sum_k c_k Tr(rho P_k) -- in the Pauli basis a gather + warp reduction.
//   sv_circuit_kernel     K5 (small n): one CTA evolves one statevector (shared memory up to 12
//                         qubits, global scratch above) and reduces <psi|P|psi> per Pauli term.
#pragma once
#include <cuda_runtime.h>

#include "program.h"

namespace bwq {

struct DmLaunch {
  double* states;              // chunk base; circuit slot s owns [s * stride, (s+1) * stride)
  int64_t stride;              // 4^n doubles
  int32_t n_digits;
  int32_t prefetch_dist;       // CTA b prefetches the tile of CTA b + prefetch_dist into L2 (0 = off)
  const SweepDesc* sweeps;     // THIS launch: one descriptor per circuit slot (sweep-major table)
  const uint4* prog;           // sweep blocks (16-byte units)
  const uint32_t* b0_table;    // [21][1024] group bases (bytes) of the register passes (fill_b0_table)
};

// ---------------------------------------------------------------------------------------------
// register-pass ops; v[da + 4*db]
// ---------------------------------------------------------------------------------------------
template <bool SW> __device__ __forceinline__ constexpr int idx2(int d0, int d1) {
  return SW ? (d1 + 4 * d0) : (d0 + 4 * d1);  // (q0,q1) digits -> register index
}
template <bool ON_B> __device__ __forceinline__ constexpr int idx1(int d, int other) {
  return ON_B ? (other + 4 * d) : (d + 4 * other);
}

// Every op runs on NG register groups at once (v[g][da + 4*db]): the parameters -- uniform
// shared-memory loads, one wavefront each -- and the op decode are paid once per NG groups.

// general 4x4 (kept for channels that are not trace preserving)
template <bool ON_B, int NG> __device__ __forceinline__ void op_dense1(double (&v)[NG][16], const double* __restrict__ m) {
  const double2* m2 = reinterpret_cast<const double2*>(m);
#pragma unroll
  for (int i4 = 0; i4 < 1; ++i4) {
    double x[NG][4][4];
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int d = 0; d < 4; ++d) x[g][o][d] = v[g][idx1<ON_B>(d, o)];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double2 a = m2[2 * i], b = m2[2 * i + 1];
#pragma unroll
      for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int o = 0; o < 4; ++o)
          v[g][idx1<ON_B>(i, o)] = fma(b.y, x[g][o][3], fma(b.x, x[g][o][2], fma(a.y, x[g][o][1], a.x * x[g][o][0])));
    }
  }
}

// trace-preserving 1-qubit channel: row 0 of the transfer matrix is (1,0,0,0); m = rows 1..3.
template <bool ON_B, int NG> __device__ __forceinline__ void op_aff1(double (&v)[NG][16], const double* __restrict__ m) {
  const double2* m2 = reinterpret_cast<const double2*>(m);
  double2 a[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) a[i] = m2[i];
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const double x0 = v[g][idx1<ON_B>(0, o)], x1 = v[g][idx1<ON_B>(1, o)], x2 = v[g][idx1<ON_B>(2, o)], x3 = v[g][idx1<ON_B>(3, o)];
      const double p1 = fma(a[1].x, x2, fma(a[0].y, x1, a[0].x * x0));
      const double p2 = fma(a[3].x, x2, fma(a[2].y, x1, a[2].x * x0));
      const double p3 = fma(a[5].x, x2, fma(a[4].y, x1, a[4].x * x0));
      v[g][idx1<ON_B>(1, o)] = fma(a[1].y, x3, p1);
      v[g][idx1<ON_B>(2, o)] = fma(a[3].y, x3, p2);
      v[g][idx1<ON_B>(3, o)] = fma(a[5].y, x3, p3);
    }
}

// rz / phase as three in-place shears: x -= t y; y += s x; x -= t y; then the optional sign
template <bool ON_B, int NG> __device__ __forceinline__ void op_rot(double (&v)[NG][16], const double* __restrict__ m) {
  const double2 ts = *reinterpret_cast<const double2*>(m);
  const double sign = m[2];
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      double x = v[g][idx1<ON_B>(1, o)], y = v[g][idx1<ON_B>(2, o)];
      x = fma(-ts.x, y, x);
      y = fma(ts.y, x, y);
      x = fma(-ts.x, y, x);
      v[g][idx1<ON_B>(1, o)] = sign < 0.0 ? -x : x;
      v[g][idx1<ON_B>(2, o)] = sign < 0.0 ? -y : y;
    }
}

// CX Pauli-transfer matrix = signed permutation (an involution: six transpositions);
// index = d_control + 4 * d_target:  y[i] = sgn[i] * v[src[i]]
__device__ constexpr int kCxSrc[16] = {0, 5, 6, 3, 4, 1, 2, 7, 11, 14, 13, 8, 15, 10, 9, 12};
__device__ constexpr int kCxSgn[16] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, 1, 1, -1, 1, 1};

template <bool CTRL_B, int NG> __device__ __forceinline__ void op_cx(double (&v)[NG][16]) {
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    double w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = v[g][i];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int j = kCxSrc[i];
      const double x = w[idx2<CTRL_B>(j & 3, j >> 2)];
      v[g][idx2<CTRL_B>(i & 3, i >> 2)] = kCxSgn[i] > 0 ? x : -x;
    }
  }
}

// out[i] = d[i] in[i]; out[Z,b] += ca[b] in[I,b]; out[a,Z] += cb[a] in[a,I]; out[Z,Z] += cab in[I,I]
// with (d0,d1) = digits of the error's (q0,q1).  WITH_CX: in = CX(v) first, control = q0 (the
// noise of a cx is keyed by (control, target)); the permutation and signs fold into operand
// selection, so the fused op costs the 25 multiply-adds of the noise alone.
// m: d[16] (index d0 + 4 d1), ca[4], cb[4], cab, pad -- 26 doubles, read as 16-byte uniform loads
template <bool SW, bool WITH_CX, int NG> __device__ __forceinline__ void op_relax2(double (&v)[NG][16], const double* __restrict__ m) {
  const double2* m2 = reinterpret_cast<const double2*>(m);
  double y[NG][16];  // input in (q0,q1) index order, after the optional CX
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int j = WITH_CX ? kCxSrc[i] : i;
      const double x = v[g][idx2<SW>(j & 3, j >> 2)];
      y[g][i] = (WITH_CX && kCxSgn[i] < 0) ? -x : x;
    }
  const double2 cb01 = m2[10], cb23 = m2[11], cabp = m2[12];
  const double cb[4] = {cb01.x, cb01.y, cb23.x, cb23.y};
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const double2 d01 = m2[2 * b], d23 = m2[2 * b + 1];
    const double2 ca2 = m2[8 + (b >> 1)];
    const double ca = (b & 1) ? ca2.y : ca2.x;
    const double d[4] = {d01.x, d01.y, d23.x, d23.y};
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        double r = d[a] * y[g][a + 4 * b];
        if (a == 3) r = fma(ca, y[g][0 + 4 * b], r);
        if (b == 3) r = fma(cb[a], y[g][a + 0], r);
        if (a == 3 && b == 3) r = fma(cabp.x, y[g][0], r);
        v[g][idx2<SW>(a, b)] = r;
      }
  }
}

template <bool SW, int NG> __device__ __forceinline__ void op_dense2(double (&v)[NG][16], const double* __restrict__ m) {
  double w[NG][16];
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int i = 0; i < 16; ++i) w[g][i] = v[g][i];
  const double2* m2 = reinterpret_cast<const double2*>(m);
#pragma unroll
  for (int i1 = 0; i1 < 4; ++i1)
#pragma unroll
    for (int i0 = 0; i0 < 4; ++i0) {
      const int row = i0 + 4 * i1;
      double s[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) s[g] = 0.0;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const double2 t = m2[row * 8 + jj];
        const int j = 2 * jj;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          s[g] = fma(t.x, w[g][idx2<SW>(j & 3, j >> 2)], s[g]);
          s[g] = fma(t.y, w[g][idx2<SW>((j + 1) & 3, (j + 1) >> 2)], s[g]);
        }
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) v[g][idx2<SW>(i0, i1)] = s[g];
    }
}

// ops and parameters of a pass, read from the sweep block staged in shared memory
template <bool FULL, int NG>
__device__ __forceinline__ void run_ops(double (&v)[NG][16], const double* __restrict__ blk, int ops_q16, int n_ops) {
  const uint4* ops = reinterpret_cast<const uint4*>(blk) + ops_q16;
  for (int o = 0; o < n_ops; ++o) {
    const uint4 raw = ops[o];
    const uint32_t pre_a = raw.x & 0xffu, pre_b = (raw.x >> 8) & 0xffu, twoq = (raw.x >> 16) & 0xffu;
    const double* ma = blk + (raw.y & 0xffffu);
    const double* mb = blk + (raw.y >> 16);
    const double* m2q = blk + (raw.z & 0xffffu);
    if (pre_a == P_AFF) op_aff1<false, NG>(v, ma);
    else if (pre_a == P_ROT) op_rot<false, NG>(v, ma);
    else if (FULL && pre_a == P_DENSE) op_dense1<false, NG>(v, ma);
    if (pre_b == P_AFF) op_aff1<true, NG>(v, mb);
    else if (pre_b == P_ROT) op_rot<true, NG>(v, mb);
    else if (FULL && pre_b == P_DENSE) op_dense1<true, NG>(v, mb);
    switch (twoq) {
      case Q_CXN_AB: op_relax2<false, true, NG>(v, m2q); break;
      case Q_CXN_BA: op_relax2<true, true, NG>(v, m2q); break;
      case Q_CX_AB: op_cx<false, NG>(v); break;
      case Q_CX_BA: op_cx<true, NG>(v); break;
      case Q_RELAX: op_relax2<false, false, NG>(v, m2q); break;
      case Q_RELAX_SW: op_relax2<true, false, NG>(v, m2q); break;
      case Q_DENSE: if constexpr (FULL) op_dense2<false, NG>(v, m2q); break;
      case Q_DENSE_SW: if constexpr (FULL) op_dense2<true, NG>(v, m2q); break;
      default: break;
    }
  }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// shared-memory swizzle.  Tile element j (digits D0..D6, 2 bits each) lives at swz(j): the low
// two digits are XOR-ed with GF(4)-linear combinations of the higher digits,
//   digit0' = D0 + D2 + D3 + D4 + D5 + D6,   digit1' = D1 + D2 + w D3 + w^2 D4 + w^2 D5 + w D6,
// i.e. D0..D4 sit on the five distinct points of the projective line over GF(4).  A half-warp of
// a register pass varies the two lowest non-target digits (always among D0..D3), so its 16
// 8-byte accesses hit 16 distinct bank pairs for EVERY choice of target slots; the linear
// load/store phases stay conflict-free as well.  swz is XOR-linear: swz(a^b) = swz(a)^swz(b).
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr uint32_t gf4_mulw(uint32_t b) { return ((((b >> 1) ^ b) & 1u) << 1) | (b >> 1); }
__host__ __device__ constexpr uint32_t swz(uint32_t j) {
  const uint32_t d2 = (j >> 4) & 3u, d3 = (j >> 6) & 3u, d4 = (j >> 8) & 3u, d5 = (j >> 10) & 3u, d6 = (j >> 12) & 3u;
  const uint32_t d45 = d4 ^ d5, d36 = d3 ^ d6;
  const uint32_t x0 = d2 ^ d36 ^ d45;
  const uint32_t x1 = d2 ^ gf4_mulw(d36) ^ gf4_mulw(d45) ^ d45;
  return j ^ (x0 | (x1 << 2));
}
// swz(d << 2s) for a single digit d in slot s (uniform lookup for the register-pass offsets)
__constant__ const uint16_t kSwzDigit[7][4] = {
    {0, 1, 2, 3},          {0, 4, 8, 12},         {0, 21, 42, 63},        {0, 73, 142, 199},
    {0, 269, 518, 779},    {0, 1037, 2054, 3083}, {0, 4105, 8206, 12295}};
// the same offsets in bytes (8-byte elements)
__constant__ const uint32_t kSwzDigit8[7][4] = {
    {0, 8, 16, 24},          {0, 32, 64, 96},         {0, 168, 336, 504},        {0, 584, 1136, 1592},
    {0, 2152, 4144, 6232},   {0, 8296, 16432, 24664}, {0, 32840, 65648, 98360}};
static_assert(swz(1u << 4) == 21 && swz(2u << 6) == 142 && swz(3u << 8) == 779 && swz(1u << 10) == 1037 &&
              swz(3u << 12) == 12295 && swz(2u << 4) == 42 && swz(3u << 6) == 199, "swizzle table");

// ---------------------------------------------------------------------------------------------
// K1/K2/K3: tile sweep
// ---------------------------------------------------------------------------------------------
template <int KQ> struct SweepCfg {
  static constexpr int kElems = 1 << (2 * KQ);
  static constexpr int kGroups = kElems / 16;
#ifndef BWQ_KQ6_NG
#define BWQ_KQ6_NG 2
#endif
#ifndef BWQ_KQ6_THREADS
#define BWQ_KQ6_THREADS (256 / BWQ_KQ6_NG)
#endif
  static constexpr int kNG = KQ == 6 ? BWQ_KQ6_NG : (kGroups >= 64 ? 2 : 1);   // register groups per thread
  static constexpr int kThreads = KQ == 6 ? BWQ_KQ6_THREADS : (kGroups / kNG >= 256 ? 256 : (kGroups / kNG >= 32 ? kGroups / kNG : 32));
#ifndef BWQ_KQ6_BLOCKS
#define BWQ_KQ6_BLOCKS 4   // 128 registers (56 B of spills), 16 warps per SM: +6..7 % over 3 x 168 registers
#endif
  static constexpr int kMinBlocks = KQ >= 7 ? 1 : (KQ == 6 ? BWQ_KQ6_BLOCKS : 4);
};

// default distance of the L2 tile prefetch: 2/3 of the 148 x 3 resident CTAs (KQ = 6); measured
// on tfim13: off 4262 GB/s, 148: 4436, 296: 4460, 444: 4417, 888: 4317
constexpr int kDmPrefetchDist = 296;

// tile-local index j (2 bits per slot) -> offset in the state (2 bits per digit position)
// (element offsets fit 32 bits: kMaxDmQubits = 16 digits)
template <int KQ> __device__ __forceinline__ uint32_t deposit(uint32_t j, const int (&pos)[KQ]) {
  uint32_t off = 0;
#pragma unroll
  for (int s = 0; s < KQ; ++s) off |= ((j >> (2 * s)) & 3u) << (2 * pos[s]);
  return off;
}
static_assert(kMaxDmQubits <= 16, "32-bit element offsets");

// tile id -> offset of the tile's first element: the bits of t fill the digit positions that are
// NOT resident in the tile, i.e. the gaps between consecutive resident positions (pos[] ascending)
template <int KQ> __device__ __forceinline__ uint32_t tile_base(uint32_t t, const int (&pos)[KQ]) {
  uint32_t base = 0, rest = t;
  int next = 0;  // next free digit position
#pragma unroll
  for (int s = 0; s < KQ; ++s) {
    const int gap = pos[s] - next;  // digits in [next, pos[s]) come from t
    base |= (rest & ((1u << (2 * gap)) - 1u)) << (2 * next);
    rest = gap >= 16 ? 0u : rest >> (2 * gap);
    next = pos[s] + 1;
  }
  return next >= 16 ? base : (base | (rest << (2 * next)));
}

// one bit per digit position (d -> bit 2d) whose digit is X or Y
__device__ __forceinline__ uint32_t xy_mask(uint32_t elem) { return (elem ^ (elem >> 1)) & 0x55555555u; }
// 16-bit set of digit positions -> bits 2d
__device__ __forceinline__ uint32_t spread_digits(uint32_t m) {
  m = (m | (m << 8)) & 0x00ff00ffu;
  m = (m | (m << 4)) & 0x0f0f0f0fu;
  m = (m | (m << 2)) & 0x33333333u;
  return (m | (m << 1)) & 0x55555555u;
}

// tile-local index of register group grp of a pass on slots (lo, hi): zero digits inserted at lo, hi
__device__ __forceinline__ uint32_t group_tile_index(uint32_t grp, int lo, int hi) {
  const uint32_t low = grp & ((1u << (2 * lo)) - 1u);
  const uint32_t mid = (grp >> (2 * lo)) & ((1u << (2 * (hi - lo - 1))) - 1u);
  const uint32_t high = grp >> (2 * (hi - 1));
  return low | (mid << (2 * lo + 2)) | (high << (2 * hi + 2));
}

// Direct pass <-> HBM transfers.  The first pass of a sweep may take its register groups straight
// from global memory and the last one may store them straight back, which removes the staging
// hop through shared memory (2 of the 2P+2 tile traversals each).  The lowering stage requests it
// only when neither target slot is one of the two lowest slots: consecutive lanes then walk the
// 16 contiguous elements of slots 0 and 1, i.e. every 8-byte warp access covers two full 128-byte
// lines.  synth: first sweep of a circuit, the tile of |0..0><0..0| is generated instead of read.
template <int KQ, int NG>
__device__ __forceinline__ void group_load_global(double (&v)[NG][16], const double* __restrict__ gtile, const int (&pos)[KQ],
                                                  const uint64_t pk, const int sa, const int sb, const int grp, const bool synth) {
  constexpr int G = SweepCfg<KQ>::kGroups, T = SweepCfg<KQ>::kThreads;
  const int lo = min(sa, sb), hi = max(sa, sb);
  const int sha = 2 * int((pk >> (8 * sa)) & 0xffu), shb = 2 * int((pk >> (8 * sb)) & 0xffu);
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const uint32_t j0 = group_tile_index(uint32_t(min(grp + g * T, G - 1)), lo, hi);
    if (synth) {
      const bool ok = ((j0 ^ (j0 >> 1)) & 0x55555555u) == 0u;  // every other digit of the group is I or Z
#pragma unroll
      for (int i = 0; i < 16; ++i) v[g][i] = (ok && (i == 0 || i == 3 || i == 12 || i == 15)) ? 1.0 : 0.0;
    } else {
      // one 64-bit pointer per group and row; the corner offsets occupy digit positions the group
      // base leaves zero, so they add (uniform byte strides) instead of being OR-ed per element
      const char* const base = reinterpret_cast<const char*>(gtile + deposit<KQ>(j0, pos));
      const uint64_t stra = uint64_t(8) << sha, strb = uint64_t(8) << shb;
#pragma unroll
      for (int db = 0; db < 4; ++db) {
        const char* const row = base + db * strb;
#pragma unroll
        for (int da = 0; da < 4; ++da) v[g][da + 4 * db] = __ldcg(reinterpret_cast<const double*>(row + da * stra));
      }
    }
  }
}
template <int KQ, int NG>
__device__ __forceinline__ void group_store_global(const double (&v)[NG][16], double* __restrict__ gtile, const int (&pos)[KQ],
                                                   const uint64_t pk, const int sa, const int sb, const int grp) {
  constexpr int G = SweepCfg<KQ>::kGroups, T = SweepCfg<KQ>::kThreads;
  const int lo = min(sa, sb), hi = max(sa, sb);
  const int sha = 2 * int((pk >> (8 * sa)) & 0xffu), shb = 2 * int((pk >> (8 * sb)) & 0xffu);
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    if (NG > 1 && grp + g * T >= G) break;
    char* const base = reinterpret_cast<char*>(gtile + deposit<KQ>(group_tile_index(uint32_t(grp + g * T), lo, hi), pos));
    const uint64_t stra = uint64_t(8) << sha, strb = uint64_t(8) << shb;
#pragma unroll
    for (int db = 0; db < 4; ++db) {
      char* const row = base + db * strb;
#pragma unroll
      for (int da = 0; da < 4; ++da) *reinterpret_cast<double*>(row + da * stra) = v[g][da + 4 * db];
    }
  }
}

// L2 prefetch of the tile CTA (blockIdx.x + prefetch_dist) will sweep: CTAs are dispatched in index
// order, so that CTA starts about one CTA lifetime from now and finds its 32 KiB in L2 instead of
// HBM (the load phase is latency bound: 12 warps per SM).  One warp issues the 4^(KQ-2) line
// prefetches; tiles that are provably zero or whose two lowest slots are not contiguous are skipped.
template <int KQ>
__device__ __forceinline__ void prefetch_next_tile(const DmLaunch& L, const int tid) {
  if (L.prefetch_dist == 0 || tid >= 32) return;
  const uint32_t nb = blockIdx.x + uint32_t(L.prefetch_dist);
  if (nb >= gridDim.x) return;
  const int tiles_log2 = 2 * (L.n_digits - KQ);
  const uint32_t slot = nb >> tiles_log2;
  const int4 d = __ldg(reinterpret_cast<const int4*>(L.sweeps + slot));
  const uint64_t pk = (uint64_t(uint32_t(d.w)) << 32) | uint32_t(d.z);
  int pos[KQ];
#pragma unroll
  for (int s = 0; s < KQ; ++s) pos[s] = int((pk >> (8 * s)) & 0xff);
  if (pos[0] != 0 || pos[1] != 1) return;
  const uint32_t base = tile_base<KQ>(nb & ((1u << tiles_log2) - 1u), pos);
  if (xy_mask(base) & spread_digits(uint32_t(d.y) >> 16)) return;
  const double* g = L.states + int64_t(slot) * L.stride + base;
  constexpr int kRuns = 1 << (2 * KQ - 4);  // 128-byte runs (slots 0 and 1) per tile
#pragma unroll 4
  for (int r = tid; r < kRuns; r += 32)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(g + deposit<KQ>(uint32_t(r) << 4, pos)));
}

// All register passes of one sweep block.  The tile is staged in shared memory unless the first
// pass loads directly (first_desc, from SweepDesc::pos[7]: known before the program block lands,
// so the loads are issued ahead of the block wait).  Returns true when the last pass already
// stored the tile to global memory.
template <int KQ, bool FULL>
__device__ __forceinline__ bool run_passes(double* __restrict__ tile, const double* __restrict__ pbuf,
                                           const uint32_t* __restrict__ b0_table, const int tid,
                                           double* __restrict__ gtile, const int (&pos)[KQ], const uint64_t pk,
                                           const uint32_t first_desc, const bool synth, const DmLaunch& L) {
  constexpr int G = SweepCfg<KQ>::kGroups, T = SweepCfg<KQ>::kThreads, NG = SweepCfg<KQ>::kNG;
  static_assert(NG == 1 || G % (T * NG) == 0, "two-group configurations cover the tile exactly");
  char* const tile_b = reinterpret_cast<char*>(tile);
  double v[NG][16];
  const bool first_direct = (first_desc & 0x80u) != 0u;
  if (first_direct && tid < G)
    group_load_global<KQ, NG>(v, gtile, pos, pk, int(first_desc & 7u), int((first_desc >> 3) & 7u), tid, synth);
  if (first_direct) prefetch_next_tile<KQ>(L, tid);
  cp_async_wait_all();
  __syncthreads();
  const int n_passes = reinterpret_cast<const int*>(pbuf)[0];
  bool stored = false;
  for (int p = 0; p < n_passes; ++p) {
    if (p) __syncthreads();
    const uint2 praw = *reinterpret_cast<const uint2*>(pbuf + 2 * (1 + p));
    const int ops_q16 = praw.x & 0xffffu, n_ops = praw.x >> 16;
    const int sa = praw.y & 0xffu, sb = (praw.y >> 8) & 0xffu;
    const bool ld_g = p == 0 && first_direct;
    const bool st_g = ((praw.y >> 16) & kPassStoreDirect) != 0u;
    const int lo = min(sa, sb), hi = max(sa, sb);
    // swizzled byte offsets of the 16 (da, db) corners (uniform) and of the thread's group base
    // (table built on the host: index = swz(grp with zero digits inserted at lo and hi))
    uint32_t oab[16];
#pragma unroll
    for (int db = 0; db < 4; ++db)
#pragma unroll
      for (int da = 0; da < 4; ++da) oab[da + 4 * db] = kSwzDigit8[sa][da] ^ kSwzDigit8[sb][db];
    const uint32_t* __restrict__ b0row = b0_table + ((hi * (hi - 1) / 2 + lo) << 10);
    for (int grp = tid; grp < G; grp += T * NG) {
      uint32_t b0[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) b0[g] = __ldg(b0row + min(grp + g * T, G - 1));
      if (ld_g) {
        if (grp != tid) group_load_global<KQ, NG>(v, gtile, pos, pk, sa, sb, grp, synth);
      } else {
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
          for (int i = 0; i < 16; ++i) v[g][i] = *reinterpret_cast<const double*>(tile_b + (b0[g] ^ oab[i]));
      }
      run_ops<FULL, NG>(v, pbuf, ops_q16, n_ops);
      if (st_g) {
        group_store_global<KQ, NG>(v, gtile, pos, pk, sa, sb, grp);
      } else {
#pragma unroll
        for (int g = 0; g < NG; ++g)
          if (NG == 1 || grp + g * T < G) {
#pragma unroll
            for (int i = 0; i < 16; ++i) *reinterpret_cast<double*>(tile_b + (b0[g] ^ oab[i])) = v[g][i];
          }
      }
    }
    stored = st_g;
  }
  return stored;
}

// FULL = also carries the dense 4x4 / 16x16 ops (coherent errors, non-basis 2-qubit gates); the
// lean instantiation keeps the register budget at 3 CTAs per SM.
template <int KQ, bool FULL>
__global__ void __launch_bounds__(SweepCfg<KQ>::kThreads, FULL ? 1 : SweepCfg<KQ>::kMinBlocks)
dm_sweep_kernel(const DmLaunch L, const int sweep_idx) {
  constexpr int E = SweepCfg<KQ>::kElems, T = SweepCfg<KQ>::kThreads;
  constexpr int U = E / 2;                       // double2 units per tile
  constexpr int NIT = (U + T - 1) / T;           // load/store iterations per thread
  // KQ <= 6: static shared memory, so the tile base is a compile-time constant and the pass
  // gather is one XOR + LDS [reg + imm] per element; KQ = 7 (128 KiB) needs the dynamic window
  extern __shared__ __align__(16) double dyn_tile[];
  __shared__ __align__(16) double st_tile[KQ <= 6 ? E + kBlockBytes / 8 : 2];
  double* const tile = KQ <= 6 ? st_tile : dyn_tile;
  const int tid = threadIdx.x;
  const int tiles_log2 = 2 * (L.n_digits - KQ);
  const uint32_t tile_mask = (1u << tiles_log2) - 1u;
  const uint32_t slot = blockIdx.x >> tiles_log2;
  const int4 swraw = __ldg(reinterpret_cast<const int4*>(L.sweeps + slot));
  double* pbuf = tile + E;
  int pos[KQ];
  const uint64_t pk = (uint64_t(uint32_t(swraw.w)) << 32) | uint32_t(swraw.z);  // pos[0..7]
#pragma unroll
  for (int s = 0; s < KQ; ++s) pos[s] = int((pk >> (8 * s)) & 0xff);
  const uint32_t first_desc = uint32_t(swraw.w) >> 24;  // pos[7]: kFirstDirect | sa | sb << 3
  const bool first_direct = (first_desc & 0x80u) != 0u;
  const uint32_t base = tile_base<KQ>(blockIdx.x & tile_mask, pos);
  double* __restrict__ g = L.states + int64_t(slot) * L.stride + base;

  // Unit u = tid + k*T covers tile elements j = 2u, 2u+1.  deposit() and swz() are bitwise
  // linear, so the per-thread part (tid) is computed once and the per-iteration part (k*T) is
  // uniform / compile-time.  swz(j + 1) = swz(j) ^ 1: the pair goes to shared memory as two
  // 8-byte accesses (same bank traffic as one 16-byte access, no select on the swizzle parity).
  const uint32_t j_thr = 2u * uint32_t(tid);
  const uint32_t off_thr = deposit<KQ>(j_thr, pos);
  const uint32_t p_thr = swz(j_thr);

  // Sparsity of the early state: |0..0><0..0| is 1 on the {I,Z}^n strings and 0 elsewhere, and a
  // digit no pass has acted on yet still is I or Z.  A tile with X or Y on such an (outside) digit
  // is all zero and stays zero under the linear passes (the affine parts act through the tile's
  // own I-component).  First sweep: plain zero store -- only 2^-(n-KQ) of its tiles run passes,
  // the rest is write-bandwidth bound; later sweeps: nothing to read, compute or write.
  const uint32_t untouched = spread_digits(uint32_t(swraw.y) >> 16);  // digit positions d -> bit 2d
  const bool tile_ok = (xy_mask(base) & untouched) == 0u;
  if (sweep_idx > 0 && !tile_ok) return;
  if (sweep_idx == 0 && !tile_ok) {
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      if (NIT * T != U && tid + k * T >= U) break;
      const uint32_t off = off_thr | deposit<KQ>(2u * uint32_t(k * T), pos);
      __stcg(reinterpret_cast<double2*>(g + off), make_double2(0.0, 0.0));
    }
    return;
  }
  // stage the sweep's program block (descriptors + parameters) while the tile streams in
  {
    const uint4* src = L.prog + uint32_t(swraw.x);
    const int len = swraw.y & 0xffff;
    for (int i = tid; i < len; i += T) cp_async16(reinterpret_cast<uint4*>(pbuf) + i, src + i);
  }

  // ---- load (or synthesise |0..0><0..0| on the first sweep); skipped when the first pass reads
  //      its register groups straight from global memory
  if (first_direct) {
  } else if (sweep_idx == 0) {
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      if (NIT * T != U && tid + k * T >= U) break;
      const uint32_t j = j_thr + 2u * uint32_t(k * T);  // D0 of j is 0 or 2
      const bool hi_ok = tile_ok && ((((j >> 2) ^ (j >> 3)) & 0x15555555u) == 0u);
      const bool d0_is2 = (j & 2u) != 0u;
      const uint32_t p = p_thr ^ swz(2u * uint32_t(k * T));
      tile[p] = (hi_ok && !d0_is2) ? 1.0 : 0.0;
      tile[p ^ 1u] = (hi_ok && d0_is2) ? 1.0 : 0.0;
    }
  } else {
    double2 val[NIT];
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      if (NIT * T != U && tid + k * T >= U) break;
      const uint32_t off = off_thr | deposit<KQ>(2u * uint32_t(k * T), pos);
      val[k] = __ldcg(reinterpret_cast<const double2*>(g + off));  // stream past L1
    }
    prefetch_next_tile<KQ>(L, tid);
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      if (NIT * T != U && tid + k * T >= U) break;
      const uint32_t p = p_thr ^ swz(2u * uint32_t(k * T));
      tile[p] = val[k].x;
      tile[p ^ 1u] = val[k].y;
    }
  }

  // ---- register passes (waits for the program block; the last pass may store the tile itself)
  if (run_passes<KQ, FULL>(tile, pbuf, L.b0_table, tid, g, pos, pk, first_desc, sweep_idx == 0, L)) return;
  __syncthreads();

  // ---- store
#pragma unroll
  for (int k = 0; k < NIT; ++k) {
    if (NIT * T != U && tid + k * T >= U) break;
    const uint32_t off = off_thr | deposit<KQ>(2u * uint32_t(k * T), pos);
    const uint32_t p = p_thr ^ swz(2u * uint32_t(k * T));
    *reinterpret_cast<double2*>(g + off) = make_double2(tile[p], tile[p ^ 1u]);
  }
}

// group-base table of the register passes: row (lo, hi) -> swz(grp with zero digits at slots lo
// and hi) in bytes, grp < 1024 (KQ = 7); filled once per context
constexpr int kB0Pairs = 21, kB0Groups = 1024;
inline void fill_b0_table(uint32_t* t) {
  for (int hi = 1; hi < 7; ++hi)
    for (int lo = 0; lo < hi; ++lo)
      for (uint32_t grp = 0; grp < (uint32_t)kB0Groups; ++grp) {
        const uint32_t low = grp & ((1u << (2 * lo)) - 1u);
        const uint32_t mid = (grp >> (2 * lo)) & ((1u << (2 * (hi - lo - 1))) - 1u);
        const uint32_t high = grp >> (2 * (hi - 1));
        t[((hi * (hi - 1) / 2 + lo) << 10) + grp] =
            8u * swz(low | (mid << (2 * lo + 2)) | (high << (2 * hi + 2)));
      }
}

// ---------------------------------------------------------------------------------------------
// K4: expectation values.  One warp per observable.
// ---------------------------------------------------------------------------------------------
struct ExpvalLaunch {
  const double* states;
  int64_t stride;
  int32_t n_obs;
  const int64_t* obs_desc;    // per observable (chunk-local): {term_begin, term_end, slot, out_index}
  const int64_t* term_index;
  const double* term_coeff;
  double* out;
};

__global__ void dm_expval_kernel(const ExpvalLaunch L) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= L.n_obs) return;
  const int64_t t0 = L.obs_desc[4 * warp], t1 = L.obs_desc[4 * warp + 1];
  const double* st = L.states + L.obs_desc[4 * warp + 2] * L.stride;
  double acc = 0.0;
  for (int64_t t = t0 + lane; t < t1; t += 32) {
    const int64_t idx = L.term_index[t];
    if (idx >= 0) acc = fma(L.term_coeff[t], st[idx], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) L.out[L.obs_desc[4 * warp + 3]] = acc;
}

// ---------------------------------------------------------------------------------------------
// K5 (small n): statevector, one CTA per circuit
// ---------------------------------------------------------------------------------------------
struct SvLaunch {
  int32_t first_circuit;
  int32_t n_circuits;
  const int32_t* circ_desc;   // per sorted circuit: {n_bits, op_begin, op_end, obs_begin, obs_end, 0,0,0}
  const SvOp* ops;
  const double* mats;
  const int64_t* obs_desc;    // per observable: {term_begin, term_end, out_index, 0}
  const uint32_t* term_x;
  const uint32_t* term_z;
  const int32_t* term_ny;
  const double* term_coeff;
  double* out;
  double2* scratch;           // global-memory states for n_bits > smem_bits
  int64_t scratch_stride;
  int32_t smem_bits;
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {
  return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

constexpr int kSvThreads = 256;

__global__ void __launch_bounds__(kSvThreads) sv_circuit_kernel(const SvLaunch L) {
  extern __shared__ __align__(16) double2 sv_smem[];
  __shared__ double red[kSvThreads / 32];
  const int tid = threadIdx.x;
  const int circ = L.first_circuit + blockIdx.x;
  const int32_t* cdsc = L.circ_desc + 8 * circ;
  const int n = cdsc[0];
  const uint32_t N = 1u << n;
  double2* psi = (n <= L.smem_bits) ? sv_smem : (L.scratch + int64_t(blockIdx.x) * L.scratch_stride);
  for (uint32_t i = tid; i < N; i += kSvThreads) psi[i] = make_double2(i == 0 ? 1.0 : 0.0, 0.0);
  __syncthreads();
  for (int o = cdsc[1]; o < cdsc[2]; ++o) {
    const SvOp op = L.ops[o];
    const double2* m = reinterpret_cast<const double2*>(L.mats + op.off);
    if (op.kind == SV_U1) {
      const double2 u00 = __ldg(m), u01 = __ldg(m + 1), u10 = __ldg(m + 2), u11 = __ldg(m + 3);
      const uint32_t q = op.q0, lowm = (1u << q) - 1u;
      for (uint32_t p = tid; p < N / 2; p += kSvThreads) {
        const uint32_t i0 = (p & lowm) | ((p & ~lowm) << 1), i1 = i0 | (1u << q);
        const double2 a = psi[i0], b = psi[i1];
        psi[i0] = cfma(u01, b, cmul(u00, a));
        psi[i1] = cfma(u11, b, cmul(u10, a));
      }
    } else {
      const uint32_t qa = min(op.q0, op.q1), qb = max(op.q0, op.q1);
      const uint32_t m0 = 1u << op.q0, m1 = 1u << op.q1;
      const uint32_t lowa = (1u << qa) - 1u, lowb = (1u << qb) - 1u;
      for (uint32_t p = tid; p < N / 4; p += kSvThreads) {
        uint32_t i = (p & lowa) | ((p & ~lowa) << 1);
        i = (i & lowb) | ((i & ~lowb) << 1);
        if (op.kind == SV_CX) {  // control q0, target q1
          const double2 a = psi[i | m0], b = psi[i | m0 | m1];
          psi[i | m0] = b;
          psi[i | m0 | m1] = a;
        } else {  // 4x4 on local index i_q0 + 2 i_q1
          double2 x[4] = {psi[i], psi[i | m0], psi[i | m1], psi[i | m0 | m1]};
          double2 y[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            double2 s = cmul(__ldg(m + r * 4), x[0]);
            s = cfma(__ldg(m + r * 4 + 1), x[1], s);
            s = cfma(__ldg(m + r * 4 + 2), x[2], s);
            s = cfma(__ldg(m + r * 4 + 3), x[3], s);
            y[r] = s;
          }
          psi[i] = y[0]; psi[i | m0] = y[1]; psi[i | m1] = y[2]; psi[i | m0 | m1] = y[3];
        }
      }
    }
    __syncthreads();
  }
  // <psi|P|psi> = Re sum_c conj(psi[c^x]) i^ny (-1)^popc(c&z) psi[c]
  for (int ob = cdsc[3]; ob < cdsc[4]; ++ob) {
    const int64_t t0 = L.obs_desc[4 * ob], t1 = L.obs_desc[4 * ob + 1];
    double total = 0.0;
    for (int64_t t = t0; t < t1; ++t) {
      const double coeff = L.term_coeff[t];
      if (coeff == 0.0) continue;
      const uint32_t x = L.term_x[t], z = L.term_z[t];
      const int ny = L.term_ny[t] & 3;
      double acc = 0.0;
      for (uint32_t c = tid; c < N; c += kSvThreads) {
        const double2 a = psi[c ^ x], b = psi[c];
        // conj(a) * b
        const double re = a.x * b.x + a.y * b.y, im = a.x * b.y - a.y * b.x;
        // times i^ny: 0: re, 1: -im, 2: -re, 3: im   (real part)
        double r = (ny == 0) ? re : (ny == 1) ? -im : (ny == 2) ? -re : im;
        acc += (__popc(c & z) & 1) ? -r : r;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      __syncthreads();
      if ((tid & 31) == 0) red[tid >> 5] = acc;
      __syncthreads();
      if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < kSvThreads / 32; ++w) s += red[w];
        total = fma(coeff, s, total);
      }
    }
    if (tid == 0) L.out[L.obs_desc[4 * ob + 2]] = total;
  }
}

}  // namespace bwq
