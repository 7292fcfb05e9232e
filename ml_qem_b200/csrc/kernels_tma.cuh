// dm_sweep_tma_kernel (sm_100a): the tile sweep of the Pauli-basis density matrix for circuits
// wider than one 6-digit tile, with the tile moved by the Tensor Memory Accelerator.
//
//   HBM -> shared memory : ONE cp.async.bulk.tensor.5d per CTA (box 16 x 4 x 4 x 4 x 4 doubles =
//                          32 KiB: the 128-byte run of digits 0,1 times the four other resident
//                          digits, each a tensor dimension of extent 4 and stride 8 * 4^pos bytes),
//                          the program block with one cp.async.bulk, both completing on one mbarrier
//   register passes      : 128 threads x 32 elements; a thread's two register groups differ in the
//                          low bit of digit 0, so every shared-memory access is 16 bytes wide
//                          (LDS.128 / STS.128 serve corner i of both groups), conflict free under
//                          CU_TENSOR_MAP_SWIZZLE_128B for the thread -> element map of program.h;
//                          the common op lists (PassSig) run as straight-line bodies: no op dispatch
//                          inside the pass, hence no register reconciliation at dispatch joins
//   shared memory -> HBM : one cp.async.bulk.tensor.5d store per CTA
//
// No thread ever computes a global address of a tile element: the former deposit() / row-pointer
// arithmetic, the LDG/STS staging and the direct global gathers of dm_sweep_kernel are gone.
// Replaces Aer's DensityMatrix::apply_superop_matrix sweeps (reached from
// blackwater/data/utils.py:422-430) like dm_sweep_kernel does; same arithmetic, bit-identical values.
#pragma once
#include <cuda.h>

#include "kernels.cuh"

namespace bwq {

struct DmTmaLaunch {
  double* states;              // chunk base (zero-tile stores of the first sweep)
  int64_t stride;              // 4^n doubles
  int32_t n_digits;
  int32_t prefetch_dist;       // CTA b prefetches the tile of CTA b + prefetch_dist into L2 (0 = off)
  const SweepDesc* sweeps;     // THIS launch: one descriptor per circuit slot
  const uint4* prog;
  const uint32_t* a_table;     // [kTmaPairs][128] thread-base byte offsets (fill_tma_a_table)
  const CUtensorMap* maps;     // [map id][window]: window w covers circuit slots [w << window_log2, ...)
  int32_t n_windows, window_log2;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_init_n(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n .reg .pred p;\n BWQ_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra BWQ_DONE;\n bra BWQ_WAIT;\n BWQ_DONE:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* map, uint64_t* bar, int c0) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %4, %4, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(0)
      : "memory");
}
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* map, const void* src, int c0) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %3, %3, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(0)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_tile(const CUtensorMap* map, int c0) {
  asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %2, %2, %2}];" ::"l"(map), "r"(c0), "r"(0) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// one macro-op with compile-time kinds
template <int PA, int PB, int TQ>
__device__ __forceinline__ void macro_op(double (&v)[2][16], const double* __restrict__ blk, const uint4 raw) {
  if constexpr (PA == P_AFF) op_aff1<false, 2>(v, blk + (raw.y & 0xffffu));
  if constexpr (PA == P_ROT) op_rot<false, 2>(v, blk + (raw.y & 0xffffu));
  if constexpr (PB == P_AFF) op_aff1<true, 2>(v, blk + (raw.y >> 16));
  if constexpr (PB == P_ROT) op_rot<true, 2>(v, blk + (raw.y >> 16));
  if constexpr (TQ == Q_CXN_AB) op_relax2<false, true, 2>(v, blk + (raw.z & 0xffffu));
}

// Straight-line pass body: 16-byte gathers (corner i of both register groups), the macro-ops of
// the signature, 16-byte scatters.  base = thread offset ^ nothing else; cor = the pass's 16 corner offsets.
// Destination of a pass's register groups: the shared-memory tile, or -- last pass of a sweep whose
// targets are not slots 0/1 -- global memory directly: gdst = the thread's first element in the state,
// gcor = the 16 corner offsets (elements); a quarter warp then writes one full 128-byte line per store
// and the tile skips its last two shared-memory traversals (scatter + TMA read).
struct PassDst {
  double* gdst;        // nullptr: scatter to the tile
  const uint4* gcor;
};
__device__ __forceinline__ void scatter_groups(const double (&v)[2][16], char* __restrict__ tile_b, const uint32_t (&off)[16], const PassDst& dst) {
  if (dst.gdst != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 c = dst.gcor[q];
      __stcg(reinterpret_cast<double2*>(dst.gdst + c.x), make_double2(v[0][4 * q], v[1][4 * q]));
      __stcg(reinterpret_cast<double2*>(dst.gdst + c.y), make_double2(v[0][4 * q + 1], v[1][4 * q + 1]));
      __stcg(reinterpret_cast<double2*>(dst.gdst + c.z), make_double2(v[0][4 * q + 2], v[1][4 * q + 2]));
      __stcg(reinterpret_cast<double2*>(dst.gdst + c.w), make_double2(v[0][4 * q + 3], v[1][4 * q + 3]));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) *reinterpret_cast<double2*>(tile_b + off[i]) = make_double2(v[0][i], v[1][i]);
  }
}

template <int PA1, int PB1, int TQ1, bool TWO, int PA2, int PB2, int TQ2>
__device__ __forceinline__ void pass_fast(char* __restrict__ tile_b, const double* __restrict__ pbuf, const uint32_t base,
                                          const uint4* __restrict__ cor, const int ops_q16, const PassDst& dst) {
  double v[2][16];
  uint32_t off[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 c = cor[q];
    off[4 * q] = base ^ c.x; off[4 * q + 1] = base ^ c.y; off[4 * q + 2] = base ^ c.z; off[4 * q + 3] = base ^ c.w;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const double2 t = *reinterpret_cast<const double2*>(tile_b + off[i]);
    v[0][i] = t.x;
    v[1][i] = t.y;
  }
  const uint4* ops = reinterpret_cast<const uint4*>(pbuf) + ops_q16;
  macro_op<PA1, PB1, TQ1>(v, pbuf, ops[0]);
  if constexpr (TWO) macro_op<PA2, PB2, TQ2>(v, pbuf, ops[1]);
  scatter_groups(v, tile_b, off, dst);
}

// Any op list.  gofs == 8: the second group is the other half of every 16-byte access; otherwise
// (slot 0 is a target) the groups are separate 8-byte gathers at base and base ^ gofs.
template <bool FULL>
__device__ __forceinline__ void pass_generic(char* __restrict__ tile_b, const double* __restrict__ pbuf, const uint32_t base,
                                             const uint4* __restrict__ cor, const int ops_q16, const int n_ops, const uint32_t gofs,
                                             const PassDst& dst) {
  double v[2][16];
  uint32_t off[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 c = cor[q];
    off[4 * q] = base ^ c.x; off[4 * q + 1] = base ^ c.y; off[4 * q + 2] = base ^ c.z; off[4 * q + 3] = base ^ c.w;
  }
  const bool wide = gofs == 8u;
  if (wide) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const double2 t = *reinterpret_cast<const double2*>(tile_b + off[i]);
      v[0][i] = t.x;
      v[1][i] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      v[0][i] = *reinterpret_cast<const double*>(tile_b + off[i]);
      v[1][i] = *reinterpret_cast<const double*>(tile_b + (off[i] ^ gofs));
    }
  }
  run_ops<FULL, 2>(v, pbuf, ops_q16, n_ops);
  if (wide) {
    scatter_groups(v, tile_b, off, dst);  // direct stores are only requested for wide passes
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      *reinterpret_cast<double*>(tile_b + off[i]) = v[0][i];
      *reinterpret_cast<double*>(tile_b + (off[i] ^ gofs)) = v[1][i];
    }
  }
}

#ifndef BWQ_TMA_BLOCKS
#define BWQ_TMA_BLOCKS 4
#endif

// barrier among the 128 compute threads of a CTA: the whole CTA (classic kernel) or named barrier 1
// (persistent kernel: the producer warp does not take part)
template <bool NAMED> __device__ __forceinline__ void compute_sync() {
  if constexpr (NAMED) asm volatile("bar.sync 1, 128;" ::: "memory");
  else __syncthreads();
}

// all register passes of the sweep block in pbuf on the tile in tile_b
// returns true when the last pass stored the tile to global memory itself (gtile = the tile's first element)
template <bool FULL, bool NAMED>
__device__ __forceinline__ bool run_tma_passes(char* __restrict__ tile_b, const double* __restrict__ pbuf,
                                               const uint32_t* __restrict__ a_table, const int tid, double* __restrict__ gtile,
                                               uint64_t* tile_bar = nullptr) {
  constexpr int T = kTmaThreads;
  const int n_passes = reinterpret_cast<const int*>(pbuf)[0];
  bool stored = false;
  const uint4* const ext = reinterpret_cast<const uint4*>(pbuf) + reinterpret_cast<const int*>(pbuf)[1];
  uint4 hraw = reinterpret_cast<const uint4*>(pbuf)[1];
  uint32_t a_next = __ldg(a_table + (hraw.z & 0xffu) * T + tid);
  if (tile_bar != nullptr) mbar_wait(tile_bar, 0);  // the tile: its flight covered the header / thread-base lookups
  for (int p = 0; p < n_passes; ++p) {
    if (p) compute_sync<NAMED>();
    const uint4 h = hraw;
    const uint32_t a_thr = a_next;
    if (p + 1 < n_passes) {  // next pass: header and thread base in flight during this pass
      hraw = reinterpret_cast<const uint4*>(pbuf)[2 + p];
      a_next = __ldg(a_table + (hraw.z & 0xffu) * T + tid);
    }
    const int ops_q16 = h.x & 0xffffu, n_ops = h.x >> 16;
    const uint32_t sig = h.y >> 24;
    const uint4* cor = ext + 4 * p;
    PassDst dst{nullptr, ext + 4 * n_passes};
    if ((h.y >> 16) & kPassStoreDirect) {
      // the thread's first element in the state: its tile-local index (tswz is an involution) with
      // every digit moved to the position of its slot
      const uint32_t j = tswz(a_thr >> 3);
      const uint2 sp = reinterpret_cast<const uint2*>(pbuf)[1];  // BlockHdr::slot_pos
      const uint64_t spk = (uint64_t(sp.y) << 32) | sp.x;
      uint32_t goff = 0;
#pragma unroll
      for (int s = 0; s < 6; ++s) goff |= ((j >> (2 * s)) & 3u) << (2u * (uint32_t(spk >> (8 * s)) & 0xffu));
      dst.gdst = gtile + goff;
      stored = true;
    }
    switch (sig) {
      case SIG_AAC: pass_fast<P_AFF, P_AFF, Q_CXN_AB, false, 0, 0, 0>(tile_b, pbuf, a_thr, cor, ops_q16, dst); break;
      case SIG_AAC_AA: pass_fast<P_AFF, P_AFF, Q_CXN_AB, true, P_AFF, P_AFF, Q_NONE>(tile_b, pbuf, a_thr, cor, ops_q16, dst); break;
      case SIG_AAC_RC: pass_fast<P_AFF, P_AFF, Q_CXN_AB, true, P_NONE, P_ROT, Q_CXN_AB>(tile_b, pbuf, a_thr, cor, ops_q16, dst); break;
      case SIG_C_RC: pass_fast<P_NONE, P_NONE, Q_CXN_AB, true, P_NONE, P_ROT, Q_CXN_AB>(tile_b, pbuf, a_thr, cor, ops_q16, dst); break;
      case SIG_RAC: pass_fast<P_ROT, P_AFF, Q_CXN_AB, false, 0, 0, 0>(tile_b, pbuf, a_thr, cor, ops_q16, dst); break;
      case SIG_0AC: pass_fast<P_NONE, P_AFF, Q_CXN_AB, false, 0, 0, 0>(tile_b, pbuf, a_thr, cor, ops_q16, dst); break;
      case SIG_AAC_A0: pass_fast<P_AFF, P_AFF, Q_CXN_AB, true, P_AFF, P_NONE, Q_NONE>(tile_b, pbuf, a_thr, cor, ops_q16, dst); break;
      case SIG_AAC_0A: pass_fast<P_AFF, P_AFF, Q_CXN_AB, true, P_NONE, P_AFF, Q_NONE>(tile_b, pbuf, a_thr, cor, ops_q16, dst); break;
      default: pass_generic<FULL>(tile_b, pbuf, a_thr, cor, ops_q16, n_ops, h.w, dst); break;
    }
  }
  return stored;
}

template <bool FULL>
__global__ void __launch_bounds__(kTmaThreads, FULL ? 2 : BWQ_TMA_BLOCKS) dm_sweep_tma_kernel(const DmTmaLaunch L, const int sweep_idx) {
  constexpr int KQ = 6, E = 1 << (2 * KQ), T = kTmaThreads;
  __shared__ __align__(1024) double tile[E];
  __shared__ __align__(16) double pbuf[kBlockBytes / 8];
  __shared__ __align__(8) uint64_t bar[2];  // [0]: program block landed, [1]: tile landed
  const int tid = threadIdx.x;
  const int tiles_log2 = 2 * (L.n_digits - KQ);
  const uint32_t tile_mask = (1u << tiles_log2) - 1u;
  const uint32_t slot = blockIdx.x >> tiles_log2;
  const int4 swraw = __ldg(reinterpret_cast<const int4*>(L.sweeps + slot));
  int pos[KQ];
  const uint64_t pk = (uint64_t(uint32_t(swraw.w)) << 32) | uint32_t(swraw.z);  // pos[0..5] ascending | map id
#pragma unroll
  for (int s = 0; s < KQ; ++s) pos[s] = int((pk >> (8 * s)) & 0xff);
  const uint32_t map_id = uint32_t(swraw.w) >> 16;
  const uint32_t base = tile_base<KQ>(blockIdx.x & tile_mask, pos);

  // zero tiles of the early state (see dm_sweep_kernel): skipped, or stored once by the first sweep
  const uint32_t untouched = spread_digits(uint32_t(swraw.y) >> 16);
  const bool tile_ok = (xy_mask(base) & untouched) == 0u;
  if (sweep_idx > 0 && !tile_ok) return;
  if (sweep_idx == 0 && !tile_ok) {
    double* __restrict__ g = L.states + int64_t(slot) * L.stride + base;
    const uint32_t off_thr = deposit<KQ>(2u * uint32_t(tid), pos);
#pragma unroll
    for (int k = 0; k < E / 2 / T; ++k)
      __stcg(reinterpret_cast<double2*>(g + (off_thr | deposit<KQ>(2u * uint32_t(k * T), pos))), make_double2(0.0, 0.0));
    return;
  }

  const CUtensorMap* const map = L.maps + map_id * uint32_t(L.n_windows) + (slot >> L.window_log2);
  const int c0 = int(((slot & ((1u << L.window_log2) - 1u)) << (2 * L.n_digits)) + base);
  if (tid == 0) {
    // both copies are issued before the CTA-wide barrier that publishes the mbarriers: the loads are
    // in flight while the other warps arrive; the (small) program block gets its own mbarrier so the
    // first pass header and thread-base lookup overlap the tile's flight
    mbar_init_n(&bar[0], 1);
    mbar_init_n(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t blk_bytes = uint32_t(swraw.y & 0xffff) * 16u;
    mbar_expect_tx(&bar[0], blk_bytes);
    bulk_load(pbuf, L.prog + uint32_t(swraw.x), blk_bytes, &bar[0]);
    if (sweep_idx > 0) {
      mbar_expect_tx(&bar[1], uint32_t(E * 8));
      tma_load_tile(tile, map, &bar[1], c0);
    }
    // L2 prefetch of the tile a later CTA will sweep (CTAs are dispatched in index order)
    if (L.prefetch_dist > 0 && sweep_idx > 0) {
      const uint32_t nb = blockIdx.x + uint32_t(L.prefetch_dist);
      if (nb < gridDim.x) {
        const uint32_t nslot = nb >> tiles_log2;
        const int4 d = __ldg(reinterpret_cast<const int4*>(L.sweeps + nslot));
        const uint64_t npk = (uint64_t(uint32_t(d.w)) << 32) | uint32_t(d.z);
        int npos[KQ];
#pragma unroll
        for (int s = 0; s < KQ; ++s) npos[s] = int((npk >> (8 * s)) & 0xff);
        const uint32_t nbase = tile_base<KQ>(nb & tile_mask, npos);
        if ((xy_mask(nbase) & spread_digits(uint32_t(d.y) >> 16)) == 0u)
          tma_prefetch_tile(L.maps + (uint32_t(d.w) >> 16) * uint32_t(L.n_windows) + (nslot >> L.window_log2),
                            int(((nslot & ((1u << L.window_log2) - 1u)) << (2 * L.n_digits)) + nbase));
      }
    }
  }
  char* const tile_b = reinterpret_cast<char*>(tile);
  if (sweep_idx == 0) {
    // |0..0><0..0| restricted to the tile: 1 on the {I,Z}^6 strings (the outside digits are I/Z: tile_ok)
#pragma unroll
    for (int k = 0; k < E / 2 / T; ++k) {
      const uint32_t j = 2u * uint32_t(tid + k * T);  // digit 0 of j is 0 or 2
      const bool hi_ok = (((j >> 2) ^ (j >> 3)) & 0x155u) == 0u;  // digits 1..5 in {I, Z}
      const bool d0_is2 = (j & 2u) != 0u;
      *reinterpret_cast<double2*>(tile_b + 8u * tswz(j)) = make_double2((hi_ok && !d0_is2) ? 1.0 : 0.0, (hi_ok && d0_is2) ? 1.0 : 0.0);
    }
  }
  __syncthreads();  // mbarrier inits visible; first sweep: the synthesised tile is complete
  mbar_wait(&bar[0], 0);

  if (run_tma_passes<FULL, false>(tile_b, pbuf, L.a_table, tid, L.states + int64_t(slot) * L.stride + base, sweep_idx > 0 ? &bar[1] : nullptr))
    return;
  // generic-proxy writes -> async proxy, then one thread stores the tile and keeps the CTA (and
  // its shared memory) alive until the bulk store has read it
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) tma_store_tile(map, tile, c0);
}

// ---------------------------------------------------------------------------------------------
// Persistent variant (sweeps after the first): CTAs stay resident (two per SM) and pull tiles from
// a per-launch counter.  Warp 4 is the producer: one thread owns all TMA traffic of the CTA -- it
// loads tile k+1 (and its program block) into the second buffer while the four compute warps run
// the passes of tile k, and stores a finished tile before it reuses the buffer.  full[b]: bytes of
// buffer b have landed (producer arrive.expect_tx + TMA complete_tx); done[b]: the compute warps
// are through with buffer b (their writes fenced to the async proxy).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_issue(const CUtensorMap* map, const void* src, int c0) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %3, %3, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(0)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

constexpr int kTmaPersistThreads = kTmaThreads + 32;
constexpr size_t kTmaPersistSmem = 2 * (4096 * 8 + kBlockBytes) + 1024 /* alignment slack */ + 64;

template <bool FULL>
__global__ void __launch_bounds__(kTmaPersistThreads, 2) dm_sweep_tma_persistent_kernel(const DmTmaLaunch L, unsigned int* __restrict__ counter,
                                                                                       const uint32_t n_tiles_total) {
  constexpr int KQ = 6, E = 1 << (2 * KQ);
  extern __shared__ unsigned char tma_dyn_smem[];
  // carve: [tile0 | tile1] 1024-aligned, [pbuf0 | pbuf1], barriers, flags
  // (offset arithmetic on the array keeps the shared address space: LDS/STS, not generic LD/ST)
  unsigned char* base_p = tma_dyn_smem + ((1024u - (smem_u32(tma_dyn_smem) & 1023u)) & 1023u);
  double* const tiles = reinterpret_cast<double*>(base_p);
  double* const pbufs = tiles + 2 * E;
  uint64_t* const full = reinterpret_cast<uint64_t*>(pbufs + 2 * (kBlockBytes / 8));
  uint64_t* const done = full + 2;
  volatile int* const valid = reinterpret_cast<volatile int*>(done + 2);
  volatile int* const stored_flag = valid + 2;                                   // the last pass stored the tile itself
  volatile int64_t* const gt_off = reinterpret_cast<volatile int64_t*>(valid + 4);  // tile's first element in the chunk
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init_n(&full[0], 1); mbar_init_n(&full[1], 1);
    mbar_init_n(&done[0], 1); mbar_init_n(&done[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= kTmaThreads) {
    // ---------------- producer (one thread)
    if (tid != kTmaThreads) return;
    const int tiles_log2 = 2 * (L.n_digits - KQ);
    const uint32_t tile_mask = (1u << tiles_log2) - 1u;
    const CUtensorMap* st_map[2] = {nullptr, nullptr};
    int st_c0[2] = {0, 0};
    uint32_t n_loaded = 0, n_stored = 0, n_real = 0;
    bool finished = false;
    for (;;) {
      if (!finished && n_loaded - n_stored < 2u) {
        const int b = int(n_loaded & 1u);
        // next tile that is not provably zero
        uint32_t t;
        const CUtensorMap* map = nullptr;
        int c0 = 0;
        uint32_t blk_q16 = 0, blk_bytes = 0;
        for (;;) {
          t = atomicAdd(counter, 1u);
          if (t >= n_tiles_total) break;
          const uint32_t slot = t >> tiles_log2;
          const int4 d = __ldg(reinterpret_cast<const int4*>(L.sweeps + slot));
          const uint64_t pk = (uint64_t(uint32_t(d.w)) << 32) | uint32_t(d.z);
          int pos[KQ];
#pragma unroll
          for (int s = 0; s < KQ; ++s) pos[s] = int((pk >> (8 * s)) & 0xff);
          const uint32_t tb = tile_base<KQ>(t & tile_mask, pos);
          if (xy_mask(tb) & spread_digits(uint32_t(d.y) >> 16)) continue;  // zero tile: nothing to do
          map = L.maps + (uint32_t(d.w) >> 16) * uint32_t(L.n_windows) + (slot >> L.window_log2);
          c0 = int(((slot & ((1u << L.window_log2) - 1u)) << (2 * L.n_digits)) + tb);
          blk_q16 = uint32_t(d.x);
          blk_bytes = uint32_t(d.y & 0xffff) * 16u;
          break;
        }
        if (t >= n_tiles_total) {  // end marker for the compute warps
          finished = true;
          valid[b] = 0;
          mbar_arrive(&full[b]);
          ++n_loaded;
          continue;
        }
        valid[b] = 1;
        gt_off[b] = int64_t(t >> tiles_log2) * L.stride + int64_t(c0 & int((1u << (2 * L.n_digits)) - 1u));
        st_map[b] = map; st_c0[b] = c0;
        mbar_expect_tx(&full[b], blk_bytes + uint32_t(E * 8));
        bulk_load(pbufs + b * (kBlockBytes / 8), L.prog + blk_q16, blk_bytes, &full[b]);
        tma_load_tile(tiles + b * E, map, &full[b], c0);
        ++n_loaded; ++n_real;
        continue;
      }
      if (n_stored < n_real) {
        const int b = int(n_stored & 1u);
        mbar_wait(&done[b], (n_stored >> 1) & 1u);
        if (!stored_flag[b]) {
          tma_store_issue(st_map[b], tiles + b * E, st_c0[b]);
          tma_store_wait_read();  // the buffer may be overwritten by the next load
        }
        ++n_stored;
        continue;
      }
      break;
    }
    tma_store_wait_all();
    return;
  }

  // ---------------- compute warps
  for (uint32_t k = 0;; ++k) {
    const int b = int(k & 1u);
    mbar_wait(&full[b], (k >> 1) & 1u);
    if (!valid[b]) break;
    const bool stored = run_tma_passes<FULL, true>(reinterpret_cast<char*>(tiles + b * E), pbufs + b * (kBlockBytes / 8), L.a_table, tid,
                                                   L.states + gt_off[b]);
    if (tid == 0) stored_flag[b] = stored ? 1 : 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    compute_sync<true>();
    if (tid == 0) mbar_arrive(&done[b]);
  }
}

}  // namespace bwq
