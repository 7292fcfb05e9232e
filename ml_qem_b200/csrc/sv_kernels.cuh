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
  // EXCHANGE fused into this sweep's store (push_g > 0, amplitude-sharded runs): the swept tile is
  // written straight into the NEW shards of the peers over NVLink -- the top push_g local index
  // bits of an amplitude select the destination rank, this rank's id takes their place (what
  // svx_exchange_kernel<PUSH> does as a separate read + write of the whole shard)
  int32_t push_g, push_rank;
  double2* push_ptr[16];       // rank w's NEW shard as mapped into this process (kSvxMaxWorld entries)
};

constexpr int kSvxThreads = 256;

// Shared-memory layout: amplitude with tile-local index j at 16 * svz12(j) (program.h): the low three
// index bits are XOR-ed with three higher bit triples, so that the eight lanes of a quarter warp
// (LDS.128 / STS.128 are served per quarter warp) hit eight distinct 16-byte chunks for every
// choice of the four pass slots; the thread -> element map comes with the pass header (tb[]).
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

// ---- register-resident ops of the fast passes: v[ia + 2 ib + 4 ic + 8 id], pass slot S compile time
// structured / general 2x2 on pass slot S: pairs (c, c | 1 << S)
template <int S, int KIND>
__device__ __forceinline__ void sv_slot_op(double2 (&v)[16], const double* __restrict__ m) {
  if constexpr (KIND == SVO_U1) {
    const double2* mc = reinterpret_cast<const double2*>(m);
    const double2 u00 = mc[0], u01 = mc[1], u10 = mc[2], u11 = mc[3];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int c0 = ((h >> S) << (S + 1)) | (h & ((1 << S) - 1)), c1 = c0 | (1 << S);
      const double2 a = v[c0], b = v[c1];
      v[c0] = cfma_d(u01, b, cmul_d(u00, a));
      v[c1] = cfma_d(u11, b, cmul_d(u10, a));
    }
  } else {
    const double2 m01 = *reinterpret_cast<const double2*>(m), m23 = *reinterpret_cast<const double2*>(m + 2);
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int c0 = ((h >> S) << (S + 1)) | (h & ((1 << S) - 1)), c1 = c0 | (1 << S);
      const double2 a = v[c0], b = v[c1];
      if constexpr (KIND == SVO_X1) {  // [[d0, i o01], [i o10, d1]] as {d0, d1, o01, o10}
        v[c0] = make_double2(fma(-m23.x, b.y, m01.x * a.x), fma(m23.x, b.x, m01.x * a.y));
        v[c1] = make_double2(fma(-m23.y, a.y, m01.y * b.x), fma(m23.y, a.x, m01.y * b.y));
      } else {                         // real {m00, m01, m10, m11}
        v[c0] = make_double2(fma(m01.y, b.x, m01.x * a.x), fma(m01.y, b.y, m01.x * a.y));
        v[c1] = make_double2(fma(m23.y, b.x, m23.x * a.x), fma(m23.y, b.y, m23.x * a.y));
      }
    }
  }
}

// phase of one amplitude under a diagonal op; x = its full physical index
__device__ __forceinline__ double2 sv_diag_phase(const uint32_t kind, const uint2 raw, const double* __restrict__ pbuf, const uint32_t x) {
  const double* m = pbuf + (raw.y & 0xffffu);
  if (kind == SVO_DZZ) {  // fused layer of ZZ-type bonds: table[number of bonds with odd parity]
    const uint2* hdr = reinterpret_cast<const uint2*>(m);
    const uint32_t n_d = hdr[0].x;
    uint32_t w = 0;
    for (uint32_t j = 0; j < n_d; ++j) {
      const uint2 dm = hdr[1 + j];
      w += __popc((x ^ (x >> dm.x)) & dm.y);
    }
    return reinterpret_cast<const double2*>(hdr + ((n_d + 2u) & ~1u))[w];
  }
  const uint32_t qa = (raw.x >> 16) & 0xffu, qb = raw.x >> 24;
  const uint32_t idx = ((x >> qa) & 1u) | (kind == SVO_D2 ? (((x >> qb) & 1u) << 1) : 0u);
  return reinterpret_cast<const double2*>(m)[idx];
}

// Pass context (uniform per CTA except base / gidx0 / active)
struct SvPassCtx {
  char* tile_b;             // shared-memory tile
  const double* pbuf;       // program block
  const SvPassHdr* ph;      // this pass's header (shared memory)
  const uint2* ops;
  uint32_t base;            // thread's swizzled byte offset (corner 0)
  uint32_t gidx0;           // full physical index of corner 0
  uint32_t goff;            // local offset of corner 0 in the shard (direct passes)
  double2* gt;              // tile base in the shard
  bool ld_g, st_g, synth, active;
};

// diagonal ops [o0, o1) of a pass, applied to the thread's 16 amplitudes IN the shared-memory tile
// (runtime loops: one small copy of this code serves every pass shape; a thread only touches its
// own amplitudes, so no barrier separates this from the gather / scatter of the same pass)
__device__ __forceinline__ void sv_diag_smem(const SvPassCtx& C, const int o0, const int o1) {
  if (!C.active || o0 >= o1) return;
  const SvPassHdr* ph = C.ph;
  const uint32_t pbit[4] = {1u << ph->pp[0], 1u << ph->pp[1], 1u << ph->pp[2], 1u << ph->pp[3]};
#pragma unroll 4
  for (uint32_t c = 0; c < 16; ++c) {
    const uint32_t x = C.gidx0 | ((c & 1u) ? pbit[0] : 0u) | ((c & 2u) ? pbit[1] : 0u) | ((c & 4u) ? pbit[2] : 0u) | ((c & 8u) ? pbit[3] : 0u);
    double2* a = reinterpret_cast<double2*>(C.tile_b + (C.base ^ ph->cor[c]));
    double2 v = *a;
    for (int o = o0; o < o1; ++o) {
      const uint2 raw = C.ops[o];
      v = cmul_d(sv_diag_phase(raw.x & 0xffu, raw, C.pbuf, x), v);
    }
    *a = v;
  }
}

// Straight-line body of a fast pass: gather the 16 amplitudes (16-byte accesses, or global loads /
// |0..0> synthesis on a direct first pass), N structured 1-qubit ops of one kind on pass slots
// 0..N-1, scatter (or direct global stores on a direct last pass).
template <int KIND, int N>
__device__ __forceinline__ void sv_pass_fast(const SvPassCtx& C, const int first_op) {
  if (!C.active) return;
  const SvPassHdr* ph = C.ph;
  double2 v[16];
  if (C.ld_g) {
    if (C.synth) {
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = make_double2((c == 0 && C.gidx0 == 0u) ? 1.0 : 0.0, 0.0);
    } else {
      const double2* __restrict__ src = C.gt + C.goff;
#pragma unroll
      for (int c = 0; c < 16; ++c)
        v[c] = __ldcg(src + (((c & 1) ? 1u << ph->pp[0] : 0u) | ((c & 2) ? 1u << ph->pp[1] : 0u) | ((c & 4) ? 1u << ph->pp[2] : 0u) |
                             ((c & 8) ? 1u << ph->pp[3] : 0u)));
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 c4 = reinterpret_cast<const uint4*>(ph->cor)[q];
      v[4 * q] = *reinterpret_cast<const double2*>(C.tile_b + (C.base ^ c4.x));
      v[4 * q + 1] = *reinterpret_cast<const double2*>(C.tile_b + (C.base ^ c4.y));
      v[4 * q + 2] = *reinterpret_cast<const double2*>(C.tile_b + (C.base ^ c4.z));
      v[4 * q + 3] = *reinterpret_cast<const double2*>(C.tile_b + (C.base ^ c4.w));
    }
  }
  if constexpr (N >= 1) sv_slot_op<0, KIND>(v, C.pbuf + (C.ops[first_op].y & 0xffffu));
  if constexpr (N >= 2) sv_slot_op<1, KIND>(v, C.pbuf + (C.ops[first_op + 1].y & 0xffffu));
  if constexpr (N >= 3) sv_slot_op<2, KIND>(v, C.pbuf + (C.ops[first_op + 2].y & 0xffffu));
  if constexpr (N >= 4) sv_slot_op<3, KIND>(v, C.pbuf + (C.ops[first_op + 3].y & 0xffffu));
  if (C.st_g) {
    double2* __restrict__ dst = C.gt + C.goff;
#pragma unroll
    for (int c = 0; c < 16; ++c)
      dst[((c & 1) ? 1u << ph->pp[0] : 0u) | ((c & 2) ? 1u << ph->pp[1] : 0u) | ((c & 4) ? 1u << ph->pp[2] : 0u) |
          ((c & 8) ? 1u << ph->pp[3] : 0u)] = v[c];
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 c4 = reinterpret_cast<const uint4*>(ph->cor)[q];
      *reinterpret_cast<double2*>(C.tile_b + (C.base ^ c4.x)) = v[4 * q];
      *reinterpret_cast<double2*>(C.tile_b + (C.base ^ c4.y)) = v[4 * q + 1];
      *reinterpret_cast<double2*>(C.tile_b + (C.base ^ c4.z)) = v[4 * q + 2];
      *reinterpret_cast<double2*>(C.tile_b + (C.base ^ c4.w)) = v[4 * q + 3];
    }
  }
}

// Any op list: the ops run one after the other on the thread's 16 amplitudes IN the shared-memory
// tile (runtime slots: no register-resident state, hence no register shuffling at the op dispatch).
// Conditional / general 1-qubit ops, 4x4 unitaries and SWAPs take this path.
__device__ __forceinline__ void sv_pass_generic(const SvPassCtx& C) {
  if (!C.active) return;
  const SvPassHdr* ph = C.ph;
  const uint32_t pbit[4] = {1u << ph->pp[0], 1u << ph->pp[1], 1u << ph->pp[2], 1u << ph->pp[3]};
  auto amp = [&](uint32_t c) { return reinterpret_cast<double2*>(C.tile_b + (C.base ^ ph->cor[c])); };
  auto gidx = [&](uint32_t c) {
    return C.gidx0 | ((c & 1u) ? pbit[0] : 0u) | ((c & 2u) ? pbit[1] : 0u) | ((c & 4u) ? pbit[2] : 0u) | ((c & 8u) ? pbit[3] : 0u);
  };
  for (int o = 0; o < ph->n_ops; ++o) {
    const uint2 raw = C.ops[o];
    const uint32_t kind = raw.x & 0xffu, flags = (raw.x >> 8) & 0xffu;
    const uint32_t sa = (raw.x >> 16) & 0xffu, sb = raw.x >> 24;
    const double2* m = reinterpret_cast<const double2*>(C.pbuf + (raw.y & 0xffffu));
    if (kind == SVO_D1 || kind == SVO_D2 || kind == SVO_DZZ) {
      for (uint32_t c = 0; c < 16; ++c) {
        double2* a = amp(c);
        *a = cmul_d(sv_diag_phase(kind, raw, C.pbuf, gidx(c)), *a);
      }
    } else if (kind == SVO_U2 || kind == SVO_SWAP) {
      const uint32_t ma = 1u << sa, mb = 1u << sb;
      for (uint32_t c = 0; c < 16; ++c) {
        if (c & (ma | mb)) continue;
        double2 *a0 = amp(c), *a1 = amp(c | ma), *a2 = amp(c | mb), *a3 = amp(c | ma | mb);
        const double2 x0 = *a0, x1 = *a1, x2 = *a2, x3 = *a3;
        if (kind == SVO_SWAP) { *a1 = x2; *a2 = x1; continue; }
        double2 y[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) y[r] = cfma_d(m[4 * r + 3], x3, cfma_d(m[4 * r + 2], x2, cfma_d(m[4 * r + 1], x1, cmul_d(m[4 * r], x0))));
        *a0 = y[0]; *a1 = y[1]; *a2 = y[2]; *a3 = y[3];
      }
    } else {  // 1-qubit op on pass slot sa, optionally conditional on a physical bit
      const uint32_t ma = 1u << sa;
      const uint32_t cbit = (raw.y >> 16) & 0xffu, want = (flags & SVF_COND_VAL) ? 1u : 0u;
      for (uint32_t c = 0; c < 16; ++c) {
        if (c & ma) continue;
        if ((flags & SVF_COND) && ((gidx(c) >> cbit) & 1u) != want) continue;
        double2 *pa = amp(c), *pb = amp(c | ma);
        const double2 a = *pa, b = *pb;
        if (kind == SVO_X) { *pa = b; *pb = a; }
        else if (kind == SVO_U1) { *pa = cfma_d(m[1], b, cmul_d(m[0], a)); *pb = cfma_d(m[3], b, cmul_d(m[2], a)); }
        else {
          const double* r = reinterpret_cast<const double*>(m);
          if (kind == SVO_X1) {
            *pa = make_double2(fma(-r[2], b.y, r[0] * a.x), fma(r[2], b.x, r[0] * a.y));
            *pb = make_double2(fma(-r[3], a.y, r[1] * b.x), fma(r[3], a.x, r[1] * b.y));
          } else {
            *pa = make_double2(fma(r[1], b.x, r[0] * a.x), fma(r[1], b.y, r[0] * a.y));
            *pb = make_double2(fma(r[3], b.x, r[2] * a.x), fma(r[3], b.y, r[2] * a.y));
          }
        }
      }
    }
  }
}

// NT threads x 16 amplitudes: NT = 256 for 2^12-amplitude tiles (two CTAs per SM at 128 registers:
// the 16 register-resident amplitudes are 64 of them), NT = 128 for tiles of 2^11 and fewer (four
// smaller CTAs per SM: the same 16 warps, but a pass barrier only stalls four of them)
template <int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 4) sv_sweep_kernel(const SvxLaunch L, const int sweep_idx) {
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
  const bool push = L.push_g > 0;
  if (dead && !first && !push) return;  // (a pushed sweep still has to deliver the zeros to the new shards)
  const int push_sh = L.n_local - L.push_g;
  const uint32_t push_low = (1u << push_sh) - 1u, push_me = uint32_t(L.push_rank) << push_sh;
  const int64_t slot_off = slot * L.stride;
  // deposit table of the 8 free slots (one entry per thread), after the program block
  uint32_t* dep = reinterpret_cast<uint32_t*>(pbuf + kBlockBytes / 8);
  for (int i = tid; i < 256; i += NT) dep[i] = svx_deposit_hi(uint32_t(i), pk);
  __syncthreads();
  const uint32_t lowmask = (1u << LB) - 1u;
#define SVX_DEPOSIT(j) (((j) & lowmask) | dep[(j) >> LB])
// destination of the amplitude with local index i when the sweep pushes its output to the peers
#define SVX_PUSH_DST(i) (L.push_ptr[(i) >> push_sh] + slot_off + (push_me | ((i) & push_low)))
  if (dead) {
    if (push) { for (uint32_t u = tid; u < E; u += NT) { const uint32_t i = base | SVX_DEPOSIT(u); __stcg(SVX_PUSH_DST(i), make_double2(0.0, 0.0)); } }
    else for (uint32_t u = tid; u < E; u += NT) __stcg(g + SVX_DEPOSIT(u), make_double2(0.0, 0.0));
    return;
  }
  {
    const uint4* src = L.prog + uint32_t(swraw.x);
    const int len = swraw.y & 0xffff;
    for (int i = tid; i < len; i += NT) cp_async16(reinterpret_cast<uint4*>(pbuf) + i, src + i);
  }
  // first pass direct (flag in the high half of blk_len): no staging of the tile; its lines are
  // prefetched into L2 while the program block is in flight
  const bool first_direct = ((uint32_t(swraw.y) >> 16) & kSvFirstDirect) != 0u;

  const uint32_t p_thr = svz12(uint32_t(tid));
  if (first_direct) {
    if (!first)
      for (uint32_t u = 8u * tid; u < E; u += 8u * NT)   // one 128-byte line = 8 amplitudes
        asm volatile("prefetch.global.L2 [%0];" ::"l"(g + SVX_DEPOSIT(u)));
  } else if (first) {
    for (uint32_t u0 = 0; u0 < E; u0 += NT) {
      const uint32_t u = u0 + tid;
      if (u < E) sv_tile[p_thr ^ svz12(u0)] = make_double2((gbase == 0u && u == 0u) ? 1.0 : 0.0, 0.0);
    }
  } else {
    for (uint32_t u0 = 0; u0 < E; u0 += 4 * NT) {
      double2 val[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t u = u0 + k * NT + tid;
        if (u < E) val[k] = __ldcg(g + SVX_DEPOSIT(u));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t uk = u0 + k * NT;
        if (uk + tid < E) sv_tile[p_thr ^ svz12(uk)] = val[k];
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int n_passes = reinterpret_cast<const int*>(pbuf)[0];
  const SvPassHdr* phdr = reinterpret_cast<const SvPassHdr*>(pbuf + 2);
  bool stored = false;  // the last pass wrote the tile back itself
  SvPassCtx C;
  C.tile_b = reinterpret_cast<char*>(sv_tile);
  C.pbuf = pbuf;
  C.gt = g;
  C.synth = first;
  C.active = uint32_t(tid) < (E >> 4);
  for (int p = 0; p < n_passes; ++p) {
    if (p) __syncthreads();
    const SvPassHdr* ph = phdr + p;
    C.ph = ph;
    C.ops = reinterpret_cast<const uint2*>(pbuf) + ph->ops_q8;
    C.ld_g = p == 0 && first_direct;
    C.st_g = (ph->flags & kPassStoreDirect) != 0u && !push;  // pushed sweeps store from the staged tile
    stored = C.st_g;
    // thread -> tile-local index of its corner 0 (zeros at the four pass slots)
    uint32_t j = 0;
    {
      const uint2 tbw = *reinterpret_cast<const uint2*>(ph->tb);
      const uint64_t tbk = (uint64_t(tbw.y) << 32) | tbw.x;
#pragma unroll
      for (int k = 0; k < 8; ++k) j |= ((uint32_t(tid) >> k) & 1u) << (uint32_t(tbk >> (8 * k)) & 31u);
      j &= E - 1u;  // tiles smaller than 2^12: unused thread bits point at bit 31
    }
    C.base = 16u * svz12(j);
    C.goff = (ph->needs_index || C.ld_g || C.st_g) ? SVX_DEPOSIT(j) : 0u;
    C.gidx0 = gbase | C.goff;
    // fast passes: [diagonal ops on the tile] [register-resident slot ops] [diagonal ops on the tile]
    // (the planner gives a direct first pass no leading and a direct last pass no trailing diagonal op)
    const int n_pre = ph->n_pre, sig = ph->sig;
    const int n_slot = sig == SVS_GENERIC ? 0 : (sig == SVS_DIAG ? 0 : ((sig - 1) & 3) + 1);
    if (sig != SVS_GENERIC) sv_diag_smem(C, 0, n_pre);
    switch (sig) {
      case SVS_X1 + 0: sv_pass_fast<SVO_X1, 1>(C, n_pre); break;
      case SVS_X1 + 1: sv_pass_fast<SVO_X1, 2>(C, n_pre); break;
      case SVS_X1 + 2: sv_pass_fast<SVO_X1, 3>(C, n_pre); break;
      case SVS_X1 + 3: sv_pass_fast<SVO_X1, 4>(C, n_pre); break;
      case SVS_R1 + 0: sv_pass_fast<SVO_R1, 1>(C, n_pre); break;
      case SVS_R1 + 1: sv_pass_fast<SVO_R1, 2>(C, n_pre); break;
      case SVS_R1 + 2: sv_pass_fast<SVO_R1, 3>(C, n_pre); break;
      case SVS_R1 + 3: sv_pass_fast<SVO_R1, 4>(C, n_pre); break;
      case SVS_U1 + 0: sv_pass_fast<SVO_U1, 1>(C, n_pre); break;
      case SVS_U1 + 1: sv_pass_fast<SVO_U1, 2>(C, n_pre); break;
      case SVS_U1 + 2: sv_pass_fast<SVO_U1, 3>(C, n_pre); break;
      case SVS_U1 + 3: sv_pass_fast<SVO_U1, 4>(C, n_pre); break;
      case SVS_DIAG: break;
      default: sv_pass_generic(C); break;
    }
    if (sig != SVS_GENERIC) sv_diag_smem(C, n_pre + n_slot, ph->n_ops);
  }
  if (stored) return;
  __syncthreads();

  if (push) {
    for (uint32_t u0 = 0; u0 < E; u0 += NT) {
      const uint32_t u = u0 + tid;
      if (u < E) { const uint32_t i = base | SVX_DEPOSIT(u); *SVX_PUSH_DST(i) = sv_tile[p_thr ^ svz12(u0)]; }
    }
    return;
  }
  for (uint32_t u0 = 0; u0 < E; u0 += NT) {
    const uint32_t u = u0 + tid;
    if (u < E) g[SVX_DEPOSIT(u)] = sv_tile[p_thr ^ svz12(u0)];
  }
#undef SVX_PUSH_DST
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
