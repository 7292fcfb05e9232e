// TMA feasibility probe for the density-matrix tile sweep (sm_100a), run under gpurun:
//   1. does cuTensorMapEncodeTiled accept a rank-5 map whose dim 0 spans the whole state buffer
//      (stride 1) and whose dims 1..4 are 4-element digits with arbitrary strides 8 * 4^pos?
//   2. what exactly is the shared-memory layout of CU_TENSOR_MAP_SWIZZLE_128B for that box?
//   3. how fast is a pure TMA load + store of 32 KiB tiles (box 16 x 4 x 4 x 4 x 4 doubles) for
//      contiguous and strided digit positions, next to an LDG/STG copy of the same tiles?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_LOOP:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE;\n bra WAIT_LOOP;\n DONE:\n}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kTile = 4096;

// tile id -> element offset of the tile: bits of t fill the digit positions not in pos[] (pos[0..5] ascending)
__host__ __device__ inline uint32_t tile_base(uint32_t t, const int* pos) {
  uint32_t base = 0, rest = t;
  int next = 0;
  for (int s = 0; s < 6; ++s) {
    int gap = pos[s] - next;
    base |= (rest & ((1u << (2 * gap)) - 1u)) << (2 * next);
    rest = gap >= 16 ? 0u : rest >> (2 * gap);
    next = pos[s] + 1;
  }
  return next >= 16 ? base : (base | (rest << (2 * next)));
}

struct Pos { int p[6]; };

// mode 0: load, dump the smem image to `dump` (tile 0 only), add 1.0 to every element, store
// mode 1: load + store (bandwidth)
__global__ void __launch_bounds__(128) tma_tile_kernel(const CUtensorMap* map, Pos pos, double* dump, int mode) {
  __shared__ __align__(1024) double tile[kTile];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  const uint32_t base = tile_base(blockIdx.x, pos.p);
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, kTile * 8);
    tma_load_5d(tile, map, &bar, (int)base, 0, 0, 0, 0);
  }
  mbar_wait(&bar, 0);
  if (mode == 0) {
    if (blockIdx.x == 0 && dump)
      for (int i = tid; i < kTile; i += 128) dump[i] = tile[i];
    for (int i = tid; i < kTile; i += 128) tile[i] += 1.0;
  }
  fence_async_smem();
  __syncthreads();
  if (tid == 0) {
    tma_store_5d(map, tile, (int)base, 0, 0, 0, 0);
    tma_store_commit_wait();
  }
}

// LDG/STG copy of the same tiles through shared memory (16-byte accesses, 128-byte runs)
__global__ void __launch_bounds__(128) ldg_tile_kernel(double* g, Pos pos) {
  __shared__ __align__(16) double tile[kTile];
  const int tid = threadIdx.x;
  const uint32_t base = tile_base(blockIdx.x, pos.p);
  double2 v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t j = 2u * (tid + 128 * k);
    uint32_t off = 0;
#pragma unroll
    for (int s = 0; s < 6; ++s) off |= ((j >> (2 * s)) & 3u) << (2 * pos.p[s]);
    v[k] = __ldcg(reinterpret_cast<const double2*>(g + base + off));
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) reinterpret_cast<double2*>(tile)[tid + 128 * k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t j = 2u * (tid + 128 * k);
    uint32_t off = 0;
#pragma unroll
    for (int s = 0; s < 6; ++s) off |= ((j >> (2 * s)) & 3u) << (2 * pos.p[s]);
    *reinterpret_cast<double2*>(g + base + off) = reinterpret_cast<double2*>(tile)[tid + 128 * k];
  }
}

static PFN_cuTensorMapEncodeTiled get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !fn) { printf("no cuTensorMapEncodeTiled entry point\n"); exit(1); }
  return (PFN_cuTensorMapEncodeTiled)fn;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 12;     // digits
  const int64_t N = int64_t(1) << (2 * n);
  const int n_states = argc > 2 ? atoi(argv[2]) : 4;
  const int64_t total = N * n_states;
  printf("n=%d states=%d total=%lld doubles (%.2f GiB)\n", n, n_states, (long long)total, total * 8.0 / (1 << 30));
  double* d = nullptr;
  CK(cudaMalloc(&d, total * 8));
  std::vector<double> h(total);
  for (int64_t i = 0; i < total; ++i) h[i] = (double)i;
  CK(cudaMemcpy(d, h.data(), total * 8, cudaMemcpyHostToDevice));
  double* dump = nullptr;
  CK(cudaMalloc(&dump, kTile * 8));
  CUtensorMap* dmap = nullptr;
  CK(cudaMalloc(&dmap, sizeof(CUtensorMap) * 4));
  auto encode = get_encode();

  std::vector<Pos> cases;
  cases.push_back(Pos{{0, 1, 2, 3, 4, 5}});
  cases.push_back(Pos{{0, 1, 3, 5, 7, 9}});
  cases.push_back(Pos{{0, 1, n - 4, n - 3, n - 2, n - 1}});
  cases.push_back(Pos{{0, 1, 2, 3, n - 2, n - 1}});
  const CUtensorMapSwizzle swz_modes[2] = {CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_NONE};
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (size_t ci = 0; ci < cases.size(); ++ci) {
    Pos P = cases[ci];
    // box-dim order = slots 2..5 in ascending position order here; the engine may permute them
    for (int sm = 0; sm < 2; ++sm) {
      alignas(64) CUtensorMap m;
      cuuint64_t gdim[5] = {(cuuint64_t)total, 4, 4, 4, 4};
      cuuint64_t gstr[4] = {8ull << (2 * P.p[2]), 8ull << (2 * P.p[3]), 8ull << (2 * P.p[4]), 8ull << (2 * P.p[5])};
      cuuint32_t box[5] = {16, 4, 4, 4, 4};
      cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz_modes[sm],
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("case %zu pos={%d,%d,%d,%d,%d,%d} swizzle=%s: encode -> %d\n", ci, P.p[0], P.p[1], P.p[2], P.p[3], P.p[4], P.p[5],
             sm == 0 ? "128B" : "NONE", (int)r);
      if (r != CUDA_SUCCESS) continue;
      CK(cudaMemcpy(dmap, &m, sizeof m, cudaMemcpyHostToDevice));
      const unsigned tiles = (unsigned)(total / kTile);
      // ---- correctness: +1 on every element through load/modify/store, smem image of tile 0
      CK(cudaMemcpy(d, h.data(), total * 8, cudaMemcpyHostToDevice));
      tma_tile_kernel<<<tiles, 128>>>(dmap, P, dump, 0);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      std::vector<double> back(total), img(kTile);
      CK(cudaMemcpy(back.data(), d, total * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(img.data(), dump, kTile * 8, cudaMemcpyDeviceToHost));
      int64_t bad = 0;
      for (int64_t i = 0; i < total; ++i) bad += back[i] != h[i] + 1.0;
      // expected layout: lin = e + 16*(D2 + 4 D3 + 16 D4 + 64 D5); byte = 8 lin; 128B swizzle: bits[4:6] ^= bits[7:9]
      int64_t bad_img = 0;
      for (int j = 0; j < kTile; ++j) {
        uint32_t off = 0;
        for (int s = 0; s < 6; ++s) off |= ((j >> (2 * s)) & 3u) << (2 * P.p[s]);
        uint32_t byte = 8u * (uint32_t)j;
        if (sm == 0) byte ^= ((byte >> 7) & 7u) << 4;
        bad_img += img[byte / 8] != (double)off;
      }
      printf("   roundtrip mismatches %lld / %lld, smem layout mismatches vs formula %lld / %d\n", (long long)bad, (long long)total,
             (long long)bad_img, kTile);
      if (bad_img && ci == 0) {
        printf("   first 40 smem words hold global offsets:");
        for (int i = 0; i < 40; ++i) printf(" %d", (int)img[i]);
        printf("\n");
      }
      // ---- bandwidth
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        for (int it = 0; it < 5; ++it) tma_tile_kernel<<<tiles, 128>>>(dmap, P, nullptr, 1);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep) printf("   TMA load+store: %.3f ms per sweep, %.0f GB/s\n", ms / 5, 2.0 * total * 8 / (ms / 5 * 1e6));
      }
    }
    for (int rep = 0; rep < 2; ++rep) {
      const unsigned tiles = (unsigned)(total / kTile);
      CK(cudaEventRecord(e0));
      for (int it = 0; it < 5; ++it) ldg_tile_kernel<<<tiles, 128>>>(d, P);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep) printf("   LDG/STG copy  : %.3f ms per sweep, %.0f GB/s\n", ms / 5, 2.0 * total * 8 / (ms / 5 * 1e6));
    }
  }
  // reference: plain device-to-device memcpy
  {
    double* d2 = nullptr;
    CK(cudaMalloc(&d2, total * 8));
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      for (int it = 0; it < 5; ++it) CK(cudaMemcpyAsync(d2, d, total * 8, cudaMemcpyDeviceToDevice));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep) printf("cudaMemcpy D2D: %.3f ms, %.0f GB/s (read+write)\n", ms / 5, 2.0 * total * 8 / (ms / 5 * 1e6));
    }
  }
  return 0;
}
