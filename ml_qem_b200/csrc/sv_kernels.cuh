// Statevector sweep kernels (sm_100a): wide registers (13..31 qubits), batched over circuits or
// amplitude-sharded across GPUs.  Program layout and execution model: program.h (SvxProgram).
//
//   sv_sweep_kernel    K5: one HBM read+write of every amplitude; a CTA stages 2^K amplitudes in
//                      shared memory (16-byte complex128, XOR-swizzled) and runs the register
//                      passes the planner packed into the sweep.  Algorithmic bytes per state
//                      sweep: 2 x 16 B x 2^n_local.
//   sv_zexp_kernel     signed sums of |amplitude|^2 for up to 32 Z-type Pauli strings per pass
//                      over the state (X/Y terms were rotated into Z by the planner).
//   sv_zexp_finalize   deterministic reduction of the per-CTA partials into the observables.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "program.h"

namespace bwq {

struct SvxLaunch {
  double2* states;             // slot s owns [s * stride, (s+1) * stride)
  int64_t stride;              // 2^n_local amplitudes
  int32_t n_local, tile_bits, low_bits;
  int32_t first_circuit;
  const int32_t* sweep_range;  // per circuit {begin, end} of this stage; nullptr: sweep_idx is absolute
  const SweepDesc* sweeps;
  const uint32_t* sweep_untouched;  // per sweep: physical bits still |0> (see sv_lowering.cpp)
  const uint4* prog;
  uint32_t hi_bits;            // rank << n_local: the global part of the physical index
  int32_t init;                // sweep_idx == 0 synthesises |0...0> instead of reading
};

constexpr int kSvxThreads = 256;

// shared-memory swizzle for 16-byte elements: LDS.128 is served per quarter-warp, so the eight
// element indices of a quarter-warp must differ in their low 3 bits.  A register pass removes two
// slots from the thread->index map, so bits 3 and 4 fold into the low bits (7 = 111b, 3 = 011b:
// any three of {001, 010, 100, 111, 011} but {001,010,011} and {100,111,011} are independent;
// those two cases -- passes on slots (2,3) and (0,1) -- pay a 2-way conflict).
__device__ __forceinline__ uint32_t svz(uint32_t j) {
  return j ^ (((j >> 3) & 1u) * 7u) ^ (((j >> 4) & 1u) * 3u);
}

__device__ __forceinline__ double2 cmul_d(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cfma_d(double2 a, double2 b, double2 c) {
  return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

// tile-local index (low bits + 8 free slots) -> offset inside the shard.  pk = the 8 slot
// positions, one byte each.
__device__ __forceinline__ uint32_t svx_deposit_hi(uint32_t up, uint64_t pk) {
  uint32_t off = 0;
#pragma unroll
  for (int s = 0; s < 8; ++s) off |= ((up >> s) & 1u) << (uint32_t(pk >> (8 * s)) & 0xffu);
  return off;
}

// 2x2 (or X) on the target slot; pairs along it: indices (0,1),(2,3) for slot a, (0,2),(1,3) for b
template <int NG, bool ON_B, bool COND>
__device__ __forceinline__ void sv_op_1q(double2 (&v)[NG][4], const double2* __restrict__ m, const bool isx,
                                         const uint32_t (&gidx0)[NG], const uint32_t cbit, const uint32_t want,
                                         const uint32_t po) {
  constexpr int P0 = 0, P1 = ON_B ? 2 : 1, Q0 = ON_B ? 1 : 2, Q1 = 3;
  if (isx) {
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      if (!COND || ((gidx0[k] >> cbit) & 1u) == want) { const double2 a = v[k][P0]; v[k][P0] = v[k][P1]; v[k][P1] = a; }
      if (!COND || (((gidx0[k] | po) >> cbit) & 1u) == want) { const double2 a = v[k][Q0]; v[k][Q0] = v[k][Q1]; v[k][Q1] = a; }
    }
    return;
  }
  const double2 u00 = m[0], u01 = m[1], u10 = m[2], u11 = m[3];
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    if (!COND || ((gidx0[k] >> cbit) & 1u) == want) {
      const double2 a = v[k][P0], b = v[k][P1];
      v[k][P0] = cfma_d(u01, b, cmul_d(u00, a));
      v[k][P1] = cfma_d(u11, b, cmul_d(u10, a));
    }
    if (!COND || (((gidx0[k] | po) >> cbit) & 1u) == want) {
      const double2 a = v[k][Q0], b = v[k][Q1];
      v[k][Q0] = cfma_d(u01, b, cmul_d(u00, a));
      v[k][Q1] = cfma_d(u11, b, cmul_d(u10, a));
    }
  }
}

// structured 2x2 (global phase dropped by the planner): XT = false: real matrix {m00, m01, m10, m11};
// XT = true: [[d0, i o01], [i o10, d1]] as {d0, d1, o01, o10} -- 4 multiply-adds per amplitude
// instead of the 8 of a complex 2x2 (rx / ry / sx / h layers of the Trotter circuits)
template <int NG, bool ON_B, bool XT>
__device__ __forceinline__ void sv_op_1s(double2 (&v)[NG][4], const double* __restrict__ m) {
  constexpr int P0 = 0, P1 = ON_B ? 2 : 1, Q0 = ON_B ? 1 : 2, Q1 = 3;
  const double2 m01 = *reinterpret_cast<const double2*>(m), m23 = *reinterpret_cast<const double2*>(m + 2);
#pragma unroll
  for (int k = 0; k < NG; ++k) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int A = h ? Q0 : P0, B = h ? Q1 : P1;
      const double2 a = v[k][A], b = v[k][B];
      if (XT) {
        v[k][A] = make_double2(fma(-m23.x, b.y, m01.x * a.x), fma(m23.x, b.x, m01.x * a.y));
        v[k][B] = make_double2(fma(-m23.y, a.y, m01.y * b.x), fma(m23.y, a.x, m01.y * b.y));
      } else {
        v[k][A] = make_double2(fma(m01.y, b.x, m01.x * a.x), fma(m01.y, b.y, m01.x * a.y));
        v[k][B] = make_double2(fma(m23.y, b.x, m23.x * a.x), fma(m23.y, b.y, m23.x * a.y));
      }
    }
  }
}

// 4x4, matrix index i_first + 2 i_second; ON_B: (first, second) = (slot b, slot a)
template <int NG, bool ON_B>
__device__ __forceinline__ void sv_op_u2(double2 (&v)[NG][4], const double2* __restrict__ m) {
  constexpr int I1 = ON_B ? 2 : 1, I2 = ON_B ? 1 : 2;
  double2 y[NG][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const double2 m0 = m[4 * r], m1 = m[4 * r + 1], m2 = m[4 * r + 2], m3 = m[4 * r + 3];
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      double2 s = cmul_d(m0, v[k][0]);
      s = cfma_d(m1, v[k][I1], s);
      s = cfma_d(m2, v[k][I2], s);
      y[k][r] = cfma_d(m3, v[k][3], s);
    }
  }
#pragma unroll
  for (int k = 0; k < NG; ++k) { v[k][0] = y[k][0]; v[k][I1] = y[k][1]; v[k][I2] = y[k][2]; v[k][3] = y[k][3]; }
}

// One register pass on NG groups per thread (groups grp0 + k * kSvxThreads).  Every thread keeps
// the 4 amplitudes v[k][ka + 2 kb] of its groups in registers and walks the op list once, so the
// op decode and the (uniform) parameter loads are shared by the NG groups.
// ld_g / st_g: direct pass -- the groups come from / go to global memory (gt = the tile's base in
// the shard; the offset of a group is the deposit of its tile index, the 4 corners add the slot
// bits 1 << pa, 1 << pb) instead of the shared-memory tile; synth: first sweep, |0...0> is generated.
template <int NG>
__device__ __forceinline__ void sv_run_pass(double2* __restrict__ tile, const double* __restrict__ pbuf,
                                            const uint32_t* __restrict__ dep, const uint2* __restrict__ ops, const int n_ops,
                                            const uint32_t grp0, const uint32_t n_grp, const int lo, const int hi,
                                            const uint32_t ma, const uint32_t mb, const uint32_t pa, const uint32_t pb,
                                            const uint32_t gbase, const int LB, const bool needs_index,
                                            double2* __restrict__ gt, const bool ld_g, const bool st_g, const bool synth) {
  uint32_t idx[NG][4], gidx0[NG], goff[NG];
  double2 v[NG][4];
  const uint32_t lowmask = (1u << LB) - 1u;
  const uint32_t ca = 1u << pa, cb = 1u << pb;
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    const uint32_t grp = min(grp0 + k * kSvxThreads, n_grp - 1u);
    uint32_t b0 = (grp & ((1u << lo) - 1u)) | ((grp >> lo) << (lo + 1));
    b0 = (b0 & ((1u << hi) - 1u)) | ((b0 >> hi) << (hi + 1));
    idx[k][0] = svz(b0); idx[k][1] = svz(b0 | ma); idx[k][2] = svz(b0 | mb); idx[k][3] = svz(b0 | ma | mb);
    goff[k] = (needs_index || ld_g || st_g) ? ((b0 & lowmask) | dep[b0 >> LB]) : 0u;
    gidx0[k] = gbase | goff[k];
    if (ld_g) {
      if (synth) {
        v[k][0] = make_double2(gidx0[k] == 0u ? 1.0 : 0.0, 0.0);
        v[k][1] = v[k][2] = v[k][3] = make_double2(0.0, 0.0);
      } else {
        const double2* __restrict__ src = gt + goff[k];
        v[k][0] = __ldcg(src); v[k][1] = __ldcg(src + ca); v[k][2] = __ldcg(src + cb); v[k][3] = __ldcg(src + (ca | cb));
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) v[k][c] = tile[idx[k][c]];
    }
  }
  for (int o = 0; o < n_ops; ++o) {
    const uint2 raw = ops[o];
    const uint32_t kind = raw.x & 0xffu, flags = (raw.x >> 8) & 0xffu;
    const uint32_t qa = (raw.x >> 16) & 0xffu, qb = raw.x >> 24;
    const double2* m = reinterpret_cast<const double2*>(pbuf + (raw.y & 0xffffu));
    if (kind == SVO_DZZ) {
      // fused layer of ZZ-type bonds: phase = table[number of bonds with odd parity]
      const uint2* hdr = reinterpret_cast<const uint2*>(m);
      const uint32_t n_d = hdr[0].x;
      const double2* tab = reinterpret_cast<const double2*>(hdr + ((n_d + 2u) & ~1u));
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        for (uint32_t j = 0; j < n_d; ++j) {
          const uint2 dm = hdr[1 + j];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t x = gidx0[k] | ((c & 1) ? ca : 0u) | ((c & 2) ? cb : 0u);
            w[c] += __popc((x ^ (x >> dm.x)) & dm.y);
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) v[k][c] = cmul_d(tab[w[c]], v[k][c]);
      }
      continue;
    }
    if (kind == SVO_D2 || kind == SVO_D1) {
      // phase index b_qa (+ 2 b_qb); gidx0 has zeros at pa and pb, so the slot bits OR in
      const bool two = kind == SVO_D2;
      const uint32_t da = uint32_t(pa == qa) | (two ? (uint32_t(pa == qb) << 1) : 0u);
      const uint32_t db = uint32_t(pb == qa) | (two ? (uint32_t(pb == qb) << 1) : 0u);
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        const uint32_t s0 = ((gidx0[k] >> qa) & 1u) | (two ? (((gidx0[k] >> qb) & 1u) << 1) : 0u);
        v[k][0] = cmul_d(m[s0], v[k][0]);
        v[k][1] = cmul_d(m[s0 | da], v[k][1]);
        v[k][2] = cmul_d(m[s0 | db], v[k][2]);
        v[k][3] = cmul_d(m[s0 | da | db], v[k][3]);
      }
      continue;
    }
    const bool on_b = (flags & SVF_ON_B) != 0u;
    const uint32_t cbit = (raw.y >> 16) & 0xffu;
    const uint32_t want = (flags & SVF_COND_VAL) ? 1u : 0u;
    const uint32_t po = 1u << (on_b ? pa : pb);  // bit of the non-target slot
    // uniform branches select straight-line variants (no per-element selects)
    if (kind == SVO_U1 || kind == SVO_X) {
      const bool isx = kind == SVO_X;
      if (flags & SVF_COND) {
        if (on_b) sv_op_1q<NG, true, true>(v, m, isx, gidx0, cbit, want, po);
        else sv_op_1q<NG, false, true>(v, m, isx, gidx0, cbit, want, po);
      } else {
        if (on_b) sv_op_1q<NG, true, false>(v, m, isx, gidx0, cbit, want, po);
        else sv_op_1q<NG, false, false>(v, m, isx, gidx0, cbit, want, po);
      }
    } else if (kind == SVO_X1) {
      if (on_b) sv_op_1s<NG, true, true>(v, reinterpret_cast<const double*>(m));
      else sv_op_1s<NG, false, true>(v, reinterpret_cast<const double*>(m));
    } else if (kind == SVO_R1) {
      if (on_b) sv_op_1s<NG, true, false>(v, reinterpret_cast<const double*>(m));
      else sv_op_1s<NG, false, false>(v, reinterpret_cast<const double*>(m));
    } else if (kind == SVO_U2) {
      if (on_b) sv_op_u2<NG, true>(v, m);
      else sv_op_u2<NG, false>(v, m);
    } else if (kind == SVO_SWAP) {
#pragma unroll
      for (int k = 0; k < NG; ++k) { const double2 a = v[k][1]; v[k][1] = v[k][2]; v[k][2] = a; }
    }
  }
#pragma unroll
  for (int k = 0; k < NG; ++k)
    if (NG == 1 ? (grp0 < n_grp) : true) {
      if (st_g) {
        double2* __restrict__ dst = gt + goff[k];
        dst[0] = v[k][0]; dst[ca] = v[k][1]; dst[cb] = v[k][2]; dst[ca | cb] = v[k][3];
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) tile[idx[k][c]] = v[k][c];
      }
    }
}

__global__ void __launch_bounds__(kSvxThreads, 3) sv_sweep_kernel(const SvxLaunch L, const int sweep_idx) {
  extern __shared__ __align__(16) double2 sv_tile[];
  const int tid = threadIdx.x;
  const int K = L.tile_bits, LB = L.low_bits;
  const uint32_t E = 1u << K;
  const int tiles_log2 = L.n_local - K;
  const int64_t slot = int64_t(blockIdx.x) >> tiles_log2;
  const uint32_t t = uint32_t(blockIdx.x) & ((1u << tiles_log2) - 1u);
  int sw_i = sweep_idx;
  if (L.sweep_range != nullptr) {
    const int circ = L.first_circuit + int(slot);
    sw_i = __ldg(L.sweep_range + 2 * circ) + sweep_idx;
    if (sw_i >= __ldg(L.sweep_range + 2 * circ + 1)) return;
  }
  const int4 swraw = __ldg(reinterpret_cast<const int4*>(L.sweeps + sw_i));
  const uint32_t untouched = __ldg(L.sweep_untouched + sw_i);
  double* pbuf = reinterpret_cast<double*>(sv_tile + E);
  const uint64_t pk = (uint64_t(uint32_t(swraw.w)) << 32) | uint32_t(swraw.z);  // slot positions
  // tile id -> the physical bits that are not resident
  uint32_t base = 0;
  {
    uint32_t rest = t;
    int next = LB;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      if (s < K - LB) {
        const int ps = int((pk >> (8 * s)) & 0xff);
        const int gap = ps - next;
        base |= (rest & ((1u << gap) - 1u)) << next;
        rest >>= gap;
        next = ps + 1;
      }
    }
    base |= rest << next;
  }
  const uint32_t gbase = L.hi_bits | base;
  double2* __restrict__ g = L.states + slot * L.stride + base;
  // early-state sparsity: an outside bit set on a qubit that is still |0> means the tile is all
  // zero and stays zero -- nothing to do, except that the very first sweep has to store the zeros
  const bool first = L.init && sweep_idx == 0;
  const bool dead = (gbase & untouched) != 0u;
  if (dead && !first) return;
  // deposit table of the 8 free slots (one entry per thread), after the program block
  uint32_t* dep = reinterpret_cast<uint32_t*>(pbuf + kBlockBytes / 8);
  dep[tid] = svx_deposit_hi(uint32_t(tid), pk);
  __syncthreads();
  const uint32_t lowmask = (1u << LB) - 1u;
#define SVX_DEPOSIT(j) (((j) & lowmask) | dep[(j) >> LB])
  if (dead) {
    for (uint32_t u = tid; u < E; u += kSvxThreads) __stcg(g + SVX_DEPOSIT(u), make_double2(0.0, 0.0));
    return;
  }
  {
    const uint4* src = L.prog + uint32_t(swraw.x);
    const int len = swraw.y & 0xffff;
    for (int i = tid; i < len; i += kSvxThreads) cp_async16(reinterpret_cast<uint4*>(pbuf) + i, src + i);
  }
  // first pass direct (descriptor in the high half of blk_len): no staging of the tile; its lines
  // are prefetched into L2 while the program block is in flight
  const bool first_direct = ((uint32_t(swraw.y) >> 16) & kSvFirstDirect) != 0u;

  const uint32_t p_thr = svz(uint32_t(tid));
  if (first_direct) {
    if (!first)
      for (uint32_t u = 8u * tid; u < E; u += 8u * kSvxThreads)   // one 128-byte line = 8 amplitudes
        asm volatile("prefetch.global.L2 [%0];" ::"l"(g + SVX_DEPOSIT(u)));
  } else if (first) {
    for (uint32_t u0 = 0; u0 < E; u0 += kSvxThreads) {
      const uint32_t u = u0 + tid;
      if (u < E) sv_tile[p_thr ^ svz(u0)] = make_double2((gbase == 0u && u == 0u) ? 1.0 : 0.0, 0.0);
    }
  } else {
    for (uint32_t u0 = 0; u0 < E; u0 += 4 * kSvxThreads) {
      double2 val[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t u = u0 + k * kSvxThreads + tid;
        if (u < E) val[k] = __ldcg(g + SVX_DEPOSIT(u));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t uk = u0 + k * kSvxThreads;
        if (uk + tid < E) sv_tile[p_thr ^ svz(uk)] = val[k];
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int n_passes = reinterpret_cast<const int*>(pbuf)[0];
  const uint2* phdr = reinterpret_cast<const uint2*>(pbuf + 2);
  bool stored = false;  // the last pass wrote the tile back itself
  for (int p = 0; p < n_passes; ++p) {
    if (p) __syncthreads();
    const uint2 praw = phdr[p];
    const int ops_q8 = praw.x & 0xffffu, n_ops = praw.x >> 16;
    const int sa = praw.y & 0xffu, sb = (praw.y >> 8) & 0xffu;
    const bool needs_index = ((praw.y >> 16) & 0xffu) != 0u;
    const bool ld_g = p == 0 && first_direct;
    const bool st_g = ((praw.y >> 24) & kPassStoreDirect) != 0u;
    stored = st_g;
    const int lo = min(sa, sb), hi = max(sa, sb);
    const uint32_t ma = 1u << sa, mb = 1u << sb;
    // physical positions of the two slots
    const uint32_t pa = sa < LB ? uint32_t(sa) : (uint32_t(pk >> (8 * (sa - LB))) & 0xffu);
    const uint32_t pb = sb < LB ? uint32_t(sb) : (uint32_t(pk >> (8 * (sb - LB))) & 0xffu);
    const uint2* ops = reinterpret_cast<const uint2*>(pbuf) + ops_q8;
    const uint32_t n_grp = E >> 2;
    if (n_grp % (2u * kSvxThreads) == 0u)
      for (uint32_t g0 = 0; g0 < n_grp; g0 += 2u * kSvxThreads)
        sv_run_pass<2>(sv_tile, pbuf, dep, ops, n_ops, tid + g0, n_grp, lo, hi, ma, mb, pa, pb, gbase, LB, needs_index,
                       g, ld_g, st_g, first);
    else
      for (uint32_t g0 = 0; g0 < n_grp; g0 += kSvxThreads)
        sv_run_pass<1>(sv_tile, pbuf, dep, ops, n_ops, tid + g0, n_grp, lo, hi, ma, mb, pa, pb, gbase, LB, needs_index,
                       g, ld_g, st_g, first);
  }
  if (stored) return;
  __syncthreads();

  for (uint32_t u0 = 0; u0 < E; u0 += kSvxThreads) {
    const uint32_t u = u0 + tid;
    if (u < E) g[SVX_DEPOSIT(u)] = sv_tile[p_thr ^ svz(u0)];
  }
#undef SVX_DEPOSIT
}

// ---------------------------------------------------------------------------------------------
// EXCHANGE over NVLink peer memory: the top g local index bits swap with the rank bits, i.e.
// block b of this rank's new shard is block `rank` of rank b's old shard.  Every rank moves its
// 2^g blocks straight between the shards (P2P loads or stores through NVSwitch; symmetric-memory
// mappings supplied by the caller) -- no staging buffers, no send/recv channels.  CTA c serves
// partner rank (rank + 1 + c) mod world, so at any moment the traffic is spread over all peers
// (every link carries 1/world of each rank) instead of all ranks hitting rank 0 first.
// ---------------------------------------------------------------------------------------------
constexpr int kSvxMaxWorld = 16;
struct SvxPeers { double2* ptr[kSvxMaxWorld]; };

// PUSH = false: ptr[w] = rank w's OLD shard, `local` = this rank's new shard (P2P loads);
// PUSH = true:  ptr[w] = rank w's NEW shard, `local` = this rank's old shard (P2P stores: block b
//               of the old shard becomes block `rank` of rank b's new shard).
template <bool PUSH>
__global__ void __launch_bounds__(256) svx_exchange_kernel(double2* __restrict__ local, const SvxPeers peers, const int world,
                                                           const int rank, const int64_t blk) {
  const int b = (rank + 1 + int(blockIdx.x % world)) % world;          // partner rank of this CTA
  const int64_t cta = blockIdx.x / world, n_cta = gridDim.x / world;    // CTAs sharing the block
  const double2* __restrict__ src = PUSH ? local + int64_t(b) * blk : peers.ptr[b] + int64_t(rank) * blk;
  double2* __restrict__ out = PUSH ? peers.ptr[b] + int64_t(rank) * blk : local + int64_t(b) * blk;
  constexpr int U = 8;                                                  // 16-byte accesses in flight per thread
  for (int64_t i0 = (cta * 256 + threadIdx.x); i0 < blk; i0 += n_cta * 256 * U) {
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + int64_t(u) * n_cta * 256;
      if (i < blk) v[u] = __ldcg(src + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + int64_t(u) * n_cta * 256;
      if (i < blk) out[i] = v[u];
    }
  }
}

// |0...0> for circuits whose first stage has no sweep (slot list)
__global__ void sv_init_kernel(double2* states, int64_t stride, const int32_t* slots, int n_slots, uint32_t hi_bits) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  if (s >= n_slots || i >= stride) return;
  states[int64_t(slots ? slots[s] : s) * stride + i] = make_double2((i == 0 && hi_bits == 0u) ? 1.0 : 0.0, 0.0);
}

// ---------------------------------------------------------------------------------------------
// Z-type expectation values
// ---------------------------------------------------------------------------------------------
constexpr int kZexpTerms = 32;   // terms per pass over the state
constexpr int kZexpThreads = 256;

struct ZexpLaunch {
  const double2* states;
  int64_t stride;
  uint32_t hi_bits;
  int32_t splits;
  const int32_t* group_desc;   // per group {state slot, first z-term, n_terms (<= 32), 0}
  const uint32_t* zt_mask;
  double* partial;             // [group][split][32]
};

// A thread owns 8 amplitudes per unit of 2048 (index = unit base | warp << 8 | j << 5 | lane, so
// every load instruction of a warp covers 512 contiguous bytes).  Their probabilities go through an
// 8-point Walsh-Hadamard transform over j (24 adds): w[c] = sum_j (-1)^popc(j & c) p_j is the
// thread's signed sum for every term whose mask has c on index bits 5..7.  A term then costs one
// parity of the remaining bits, one shared-memory read of w[c] (uniform c: conflict free) and one
// add per EIGHT amplitudes instead of a parity, select and add per amplitude.
__global__ void __launch_bounds__(kZexpThreads) sv_zexp_kernel(const ZexpLaunch L) {
  __shared__ double red[kZexpThreads / 32][kZexpTerms];
  __shared__ double wsm[8][kZexpThreads];
  const int grp = blockIdx.x / L.splits, sp = blockIdx.x % L.splits;
  const int4 d = __ldg(reinterpret_cast<const int4*>(L.group_desc) + grp);
  const int nt = d.z;
  uint32_t mask[kZexpTerms];
#pragma unroll
  for (int t = 0; t < kZexpTerms; ++t) mask[t] = t < nt ? __ldg(L.zt_mask + d.y + t) : 0u;
  double acc[kZexpTerms];
#pragma unroll
  for (int t = 0; t < kZexpTerms; ++t) acc[t] = 0.0;
  const double2* st = L.states + int64_t(d.x) * L.stride;
  if ((L.stride & 2047) == 0) {
    const int tid = threadIdx.x;
    const int64_t units = L.stride >> 11;
    const int64_t u0 = units * sp / L.splits, u1 = units * (sp + 1) / L.splits;
    const uint32_t in_unit = (uint32_t(tid >> 5) << 8) | uint32_t(tid & 31);  // bits 5..7 = j = 0
    for (int64_t u = u0; u < u1; ++u) {
      const double2* src = st + (u << 11) + in_unit;
      double w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double2 a = __ldcs(src + 32 * j);
        w[j] = fma(a.x, a.x, a.y * a.y);
      }
#pragma unroll
      for (int h = 1; h < 8; h <<= 1)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (!(j & h)) { const double x = w[j], y = w[j | h]; w[j] = x + y; w[j | h] = x - y; }
#pragma unroll
      for (int c = 0; c < 8; ++c) wsm[c][tid] = w[c];
      const uint32_t gi = L.hi_bits | uint32_t(u << 11) | in_unit;
#pragma unroll
      for (int t = 0; t < kZexpTerms; ++t)
        if (t < nt) {
          const double x = wsm[(mask[t] >> 5) & 7u][tid];
          acc[t] += (__popc(gi & mask[t] & ~0xe0u) & 1) ? -x : x;
        }
    }
  } else {  // shards smaller than one unit (tests, simulated ranks)
    const int64_t chunk = (L.stride + L.splits - 1) / L.splits;
    const int64_t i0 = chunk * sp, i1 = min(L.stride, i0 + chunk);
    for (int64_t i = i0 + threadIdx.x; i < i1; i += kZexpThreads) {
      const double2 a = __ldcs(st + i);
      const double p = fma(a.x, a.x, a.y * a.y);
      const uint32_t gi = L.hi_bits | uint32_t(i);
#pragma unroll
      for (int t = 0; t < kZexpTerms; ++t)
        if (t < nt) acc[t] += (__popc(gi & mask[t]) & 1) ? -p : p;
    }
  }
#pragma unroll
  for (int t = 0; t < kZexpTerms; ++t) {
    double a = acc[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][t] = a;
  }
  __syncthreads();
  if (threadIdx.x < kZexpTerms) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kZexpThreads / 32; ++w) s += red[w][threadIdx.x];
    L.partial[(int64_t(blockIdx.x)) * kZexpTerms + threadIdx.x] = s;
  }
}

struct ZexpFinalize {
  const int32_t* circ_desc;    // per circuit {first group, n_groups, obs base, 0}
  const int32_t* group_desc;
  const double* zt_coeff;
  const int32_t* zt_obs;
  const double* partial;
  int32_t splits;
  double* out;                 // out[obs base + zt_obs] += coeff * value
};

// one CTA per circuit: the split partials of every term are summed in a fixed order, then thread
// 0 accumulates the observables term by term (deterministic)
__global__ void __launch_bounds__(128) sv_zexp_finalize(const ZexpFinalize F) {
  __shared__ double val[kZexpTerms];
  const int4 cd = __ldg(reinterpret_cast<const int4*>(F.circ_desc) + blockIdx.x);
  for (int gi = cd.x; gi < cd.x + cd.y; ++gi) {
    const int4 gd = __ldg(reinterpret_cast<const int4*>(F.group_desc) + gi);
    __syncthreads();
    if (threadIdx.x < gd.z) {
      double s = 0.0;
      for (int sp = 0; sp < F.splits; ++sp) s += F.partial[(int64_t(gi) * F.splits + sp) * kZexpTerms + threadIdx.x];
      val[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int t = 0; t < gd.z; ++t) {
        const double c = F.zt_coeff[gd.y + t];
        if (c != 0.0) F.out[cd.z + F.zt_obs[gd.y + t]] += c * val[t];
      }
  }
}

}  // namespace bwq
