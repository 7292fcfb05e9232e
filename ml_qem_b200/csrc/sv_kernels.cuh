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
  double* pbuf = reinterpret_cast<double*>(sv_tile + E);
  {
    const uint4* src = L.prog + uint32_t(swraw.x);
    const int len = swraw.y;
    for (int i = tid; i < len; i += kSvxThreads) cp_async16(reinterpret_cast<uint4*>(pbuf) + i, src + i);
  }
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
  // deposit table of the 8 free slots (one entry per thread), after the program block
  uint32_t* dep = reinterpret_cast<uint32_t*>(pbuf + kBlockBytes / 8);
  dep[tid] = svx_deposit_hi(uint32_t(tid), pk);
  __syncthreads();
  const uint32_t lowmask = (1u << LB) - 1u;
#define SVX_DEPOSIT(j) (((j) & lowmask) | dep[(j) >> LB])

  const uint32_t p_thr = svz(uint32_t(tid));
  if (L.init && sweep_idx == 0) {
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
  for (int p = 0; p < n_passes; ++p) {
    if (p) __syncthreads();
    const uint2 praw = phdr[p];
    const int ops_q8 = praw.x & 0xffffu, n_ops = praw.x >> 16;
    const int sa = praw.y & 0xffu, sb = (praw.y >> 8) & 0xffu;
    const bool needs_index = ((praw.y >> 16) & 0xffu) != 0u;
    const int lo = min(sa, sb), hi = max(sa, sb);
    const uint32_t ma = 1u << sa, mb = 1u << sb;
    // physical positions of the two slots
    const uint32_t pa = sa < LB ? uint32_t(sa) : (uint32_t(pk >> (8 * (sa - LB))) & 0xffu);
    const uint32_t pb = sb < LB ? uint32_t(sb) : (uint32_t(pk >> (8 * (sb - LB))) & 0xffu);
    const uint2* ops = reinterpret_cast<const uint2*>(pbuf) + ops_q8;
    for (uint32_t grp = tid; grp < (E >> 2); grp += kSvxThreads) {
      uint32_t b0 = (grp & ((1u << lo) - 1u)) | ((grp >> lo) << (lo + 1));
      b0 = (b0 & ((1u << hi) - 1u)) | ((b0 >> hi) << (hi + 1));
      const uint32_t i0 = svz(b0), i1 = svz(b0 | ma), i2 = svz(b0 | mb), i3 = svz(b0 | ma | mb);
      double2 v0 = sv_tile[i0], v1 = sv_tile[i1], v2 = sv_tile[i2], v3 = sv_tile[i3];  // v[ka + 2 kb]
      const uint32_t gidx0 = needs_index ? (gbase | SVX_DEPOSIT(b0)) : 0u;
      for (int o = 0; o < n_ops; ++o) {
        const uint2 raw = ops[o];
        const uint32_t kind = raw.x & 0xffu, flags = (raw.x >> 8) & 0xffu;
        const uint32_t qa = (raw.x >> 16) & 0xffu, qb = raw.x >> 24;
        const double2* m = reinterpret_cast<const double2*>(pbuf + (raw.y & 0xffffu));
        if (kind == SVO_D2) {
          // phase index b_qa + 2 b_qb; gidx0 has zeros at pa and pb, so the slot bits OR in
          const uint32_t s0 = ((gidx0 >> qa) & 1u) | (((gidx0 >> qb) & 1u) << 1);
          const uint32_t da = uint32_t(pa == qa) | (uint32_t(pa == qb) << 1);
          const uint32_t db = uint32_t(pb == qa) | (uint32_t(pb == qb) << 1);
          v0 = cmul_d(m[s0], v0);
          v1 = cmul_d(m[s0 | da], v1);
          v2 = cmul_d(m[s0 | db], v2);
          v3 = cmul_d(m[s0 | da | db], v3);
          continue;
        }
        if (kind == SVO_D1) {
          const uint32_t s0 = (gidx0 >> qa) & 1u;
          const uint32_t da = uint32_t(pa == qa), db = uint32_t(pb == qa);
          v0 = cmul_d(m[s0], v0);
          v1 = cmul_d(m[s0 | da], v1);
          v2 = cmul_d(m[s0 | db], v2);
          v3 = cmul_d(m[s0 | da | db], v3);
          continue;
        }
        const bool on_b = (flags & SVF_ON_B) != 0u;
        // conditional ops: does the pair with the OTHER slot's bit = k qualify?
        bool c0 = true, c1 = true;
        if (flags & SVF_COND) {
          const uint32_t cbit = (raw.y >> 16) & 0xffu;
          const uint32_t want = (flags & SVF_COND_VAL) ? 1u : 0u;
          const uint32_t po = on_b ? pa : pb;  // position of the non-target slot
          c0 = ((gidx0 >> cbit) & 1u) == want;
          c1 = (((gidx0 | (1u << po)) >> cbit) & 1u) == want;
        }
        if (kind == SVO_U1) {
          const double2 u00 = m[0], u01 = m[1], u10 = m[2], u11 = m[3];
          // pairs along the target slot: (x0,x1) and (y0,y1)
          double2 x0 = v0, x1 = on_b ? v2 : v1, y0 = on_b ? v1 : v2, y1 = v3;
          if (c0) { const double2 a = x0, b = x1; x0 = cfma_d(u01, b, cmul_d(u00, a)); x1 = cfma_d(u11, b, cmul_d(u10, a)); }
          if (c1) { const double2 a = y0, b = y1; y0 = cfma_d(u01, b, cmul_d(u00, a)); y1 = cfma_d(u11, b, cmul_d(u10, a)); }
          v0 = x0; v3 = y1;
          if (on_b) { v2 = x1; v1 = y0; } else { v1 = x1; v2 = y0; }
        } else if (kind == SVO_X) {
          double2 x0 = v0, x1 = on_b ? v2 : v1, y0 = on_b ? v1 : v2, y1 = v3;
          if (c0) { const double2 a = x0; x0 = x1; x1 = a; }
          if (c1) { const double2 a = y0; y0 = y1; y1 = a; }
          v0 = x0; v3 = y1;
          if (on_b) { v2 = x1; v1 = y0; } else { v1 = x1; v2 = y0; }
        } else if (kind == SVO_U2) {
          // matrix index i_first + 2 i_second; on_b: (first, second) = (slot b, slot a)
          const double2 x0 = v0, x1 = on_b ? v2 : v1, x2 = on_b ? v1 : v2, x3 = v3;
          double2 y[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            double2 s = cmul_d(m[4 * r], x0);
            s = cfma_d(m[4 * r + 1], x1, s);
            s = cfma_d(m[4 * r + 2], x2, s);
            s = cfma_d(m[4 * r + 3], x3, s);
            y[r] = s;
          }
          v0 = y[0]; v3 = y[3];
          if (on_b) { v2 = y[1]; v1 = y[2]; } else { v1 = y[1]; v2 = y[2]; }
        } else if (kind == SVO_SWAP) {
          const double2 a = v1; v1 = v2; v2 = a;
        }
      }
      sv_tile[i0] = v0; sv_tile[i1] = v1; sv_tile[i2] = v2; sv_tile[i3] = v3;
    }
  }
  __syncthreads();

  for (uint32_t u0 = 0; u0 < E; u0 += kSvxThreads) {
    const uint32_t u = u0 + tid;
    if (u < E) g[SVX_DEPOSIT(u)] = sv_tile[p_thr ^ svz(u0)];
  }
#undef SVX_DEPOSIT
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

__global__ void __launch_bounds__(kZexpThreads) sv_zexp_kernel(const ZexpLaunch L) {
  __shared__ double red[kZexpThreads / 32][kZexpTerms];
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
