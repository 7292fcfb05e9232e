// C ABI (include/bwq.h): context, batch scheduling, launches.  There is no CPU fallback: every
// *_run entry point needs a bwq_ctx, and bwq_create fails without a CUDA device.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <cuda.h>
#include <cudaTypedefs.h>
#include <nvtx3/nvToolsExt.h>

#include "kernels.cuh"
#include "kernels_tma.cuh"
#include "sv_kernels.cuh"
#include "onchip.cuh"

using namespace bwq;

struct bwq_program {
  CircuitProgram p;
};

namespace {

thread_local std::string g_create_error;

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// BWQ_TRACE=1: wall-clock phase marks of the *_run entry points on stderr (host-side tuning)
struct Trace {
  bool on;
  double t0, last;
  explicit Trace(const char* what) : on(std::getenv("BWQ_TRACE") != nullptr), t0(now_ms()), last(t0) { if (on) std::fprintf(stderr, "[bwq trace] %s\n", what); }
  void mark(const char* what) {
    if (!on) return;
    const double t = now_ms();
    std::fprintf(stderr, "[bwq trace]   %-34s +%7.3f ms  (at %7.3f)\n", what, t - last, t - t0);
    last = t;
  }
};

// NVTX range over a host-side stage (visible in ncu / nsys timelines; header-only, no cost without a tool attached)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// grow-only device / pinned-host buffers
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { cudaGetLastError(); want = bytes; e = cudaMalloc(&p, want); }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// one blob = several arrays packed with 256-B alignment, uploaded with a single cudaMemcpyAsync
struct Blob {
  std::vector<size_t> offs;
  size_t total = 0;
  size_t add(size_t bytes) {
    size_t o = total;
    offs.push_back(o);
    total += (bytes + 255) & ~size_t(255);
    return o;
  }
};

}  // namespace

struct DmChunk {
  int first = 0, count = 0, nd = 0, kq = 0;
  bool full = false;
  bool tma = false;               // sweeps in the TMA tile layout: dm_sweep_tma_kernel
  int window_log2 = 0, n_windows = 1;  // circuit slots per tensor map (dim 0 must stay below 2^31 elements)
  size_t map_first = 0;           // first tensor map of the chunk in the plan's map array
  std::vector<uint32_t> map_keys; // per map id: the four upper digit positions in box order, one byte each
  std::vector<int> live;  // circuits that still have a sweep s, per sweep index
  std::vector<int64_t> desc_off;  // per sweep index: first descriptor of the launch (sweep-major table)
  std::vector<int64_t> bytes;  // algorithmic bytes of launch s: tiles that are not provably zero
};
struct DmPlan {
  bool valid = false;
  int tile_qubits = 6;
  int64_t n_obs = 0;
  size_t o_sweeps = 0, o_prog = 0, o_tidx = 0, o_tcoef = 0, o_obs = 0, blob_bytes = 0;
  std::vector<int64_t> ob_off;
  std::vector<DmChunk> chunks;
  std::vector<std::pair<int64_t, double>> host_fix;
  int64_t max_chunk_bytes = 0, n_gates = 0, n_passes = 0;
  double lower_ms = 0, h2d_ms = 0;
  size_t n_maps = 0;              // tensor maps of all TMA chunks
  const void* maps_for = nullptr; // state buffer the uploaded maps were encoded for
};

struct SvGroup {
  int first = 0, count = 0, nb = 0, per_launch = 0;
  size_t smem = 0;
  int64_t stride = 0;
};
// wide circuits (> kSvSmallBits active qubits): tile-sweep path, batched per chunk of equal width
struct SvWideStage {
  int max_sweeps = 0;       // sweep launches of this stage
  size_t range_off = 0;     // int32 index of the chunk's {begin,end} array for this stage
  int group_first = 0, n_groups = 0;  // index into the group_desc array
  int cdesc_first = 0, n_cdesc = 0;   // circuits that evaluate terms in this stage
};
struct SvWideChunk {
  int first = 0, count = 0, nb = 0, tile_bits = 0, low_bits = 0;
  std::vector<SvWideStage> stages;
  size_t init_off = 0;      // int32 index of the slots that need an explicit |0..0>
  int n_init = 0;
};
struct SvWidePlan {
  size_t o_range = 0, o_sweeps = 0, o_unt = 0, o_prog = 0, o_ztm = 0, o_ztc = 0, o_zto = 0, o_gdesc = 0, o_cdesc = 0, o_init = 0;
  size_t blob_bytes = 0;
  int64_t bytes_per_exec = 0;   // algorithmic bytes of one execute (live tiles only)
  std::vector<SvWideChunk> chunks;
  int64_t max_state_bytes = 0, n_passes = 0;
  bool any = false;
};
struct SvPlan {
  bool valid = false;
  int64_t n_obs = 0, n_gates = 0;
  size_t o_cd = 0, o_ops = 0, o_mats = 0, o_obs = 0, o_tx = 0, o_tz = 0, o_tny = 0, o_tc = 0, blob_bytes = 0;
  std::vector<SvGroup> groups;
  std::vector<int64_t> nan_obs;
  SvWidePlan wide;
  double lower_ms = 0, h2d_ms = 0;
};

constexpr int kMaxPersistLaunches = 16384;

struct bwq_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::string error;
  bwq_options opt{};
  NoiseTable noise;
  DevBuf d_b0, d_noise, d_sv_prog, d_states, d_out, d_scratch, d_wide_prog, d_partial;
  PinBuf h_sv_prog, h_out, h_wide_prog;
  bwq_stats stats{};
  // density-matrix program slots: slot 0 is the prepared batch of bwq_dm_prepare / bwq_dm_execute;
  // bwq_dm_run alternates between both so that segment k+1 is lowered on the host threads while
  // the GPU executes segment k
  struct DmSlot { DmPlan plan; PinBuf h_prog, h_maps; DevBuf d_prog, d_maps; cudaEvent_t h2d_done = nullptr; } dm[2];
  DevBuf d_tma_a;                     // thread-base table of dm_sweep_tma_kernel
  DevBuf d_counters;                  // tile counters of the persistent launches (one per launch of an execute)
  PFN_cuTensorMapEncodeTiled encode_tiled = nullptr;  // driver entry point; null => no TMA path
  SvPlan sv_plan;
  std::vector<cudaEvent_t> chunk_ev;  // begin/end of each chunk's sweep launches
  bwq_ctx* companion = nullptr;       // statevector side of bwq_meas_data_run (created on first use)
  // dm_onchip_kernel: raw batch blob, result buffers, device copy of the noise lookup tables
  DevBuf d_oc, d_oc_out, d_oc_noise;
  PinBuf h_oc, h_oc_out;
  OnchipNoise oc_noise{};
  bool oc_noise_valid = false;
  int64_t budget_cache = 0; size_t budget_cap = 0; double budget_time_ms = 0; uint64_t budget_key = 0;  // dm_state_budget
  cudaStream_t oc_copy_stream = nullptr;          // uploads of range r+1 overlap the kernel of range r
  cudaEvent_t oc_copied[8] = {}, oc_k0[8] = {}, oc_k1[8] = {};
  size_t smem_optin = 0;
  int sm_count = 0;
};

static int fail(bwq_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->error = buf; else g_create_error = buf;
  return code;
}
#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(ctx, BWQ_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),      \
                  __FILE__, __LINE__);                                                            \
  } while (0)

extern "C" int bwq_version(void) { return BWQ_VERSION; }

extern "C" const char* bwq_last_error(const bwq_ctx* ctx) {
  return ctx ? ctx->error.c_str() : g_create_error.c_str();
}

extern "C" int bwq_create(int device, bwq_ctx** out) {
  if (!out) return fail(nullptr, BWQ_ERR_ARG, "bwq_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(nullptr, BWQ_ERR_NO_DEVICE,
                "bwq_create: no CUDA device (%s); this engine has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= count) return fail(nullptr, BWQ_ERR_ARG, "bwq_create: device %d out of range [0,%d)", device, count);
  bwq_ctx* ctx = new bwq_ctx();
  ctx->device = device;
  auto bail = [&](cudaError_t ce, const char* what) {
    fail(nullptr, BWQ_ERR_CUDA, "bwq_create: %s: %s", what, cudaGetErrorString(ce));
    delete ctx;
    return BWQ_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
  for (auto& ev : ctx->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(e, "cudaEventCreate");
  for (auto& sl : ctx->dm)
    if ((e = cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
  ctx->chunk_ev.resize(128);
  for (auto& ev : ctx->chunk_ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(e, "cudaEventCreate");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  ctx->sm_count = prop.multiProcessorCount;
  if ((e = cudaFuncSetAttribute(dm_sweep_kernel<7, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (8 << 14) + kBlockBytes)) != cudaSuccess)
    return bail(e, "cudaFuncSetAttribute(dm_sweep_kernel<7>) -- was the library built for this GPU (sm_100a)?");
  if ((e = cudaFuncSetAttribute(dm_sweep_kernel<7, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (8 << 14) + kBlockBytes)) != cudaSuccess)
    return bail(e, "cudaFuncSetAttribute(dm_sweep_kernel<7, full>)");
  if ((e = cudaFuncSetAttribute(dm_onchip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOnchipSmem)) != cudaSuccess)
    return bail(e, "cudaFuncSetAttribute(dm_onchip_kernel)");
  if ((e = cudaFuncSetAttribute(sv_circuit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 << 12)) != cudaSuccess)
    return bail(e, "cudaFuncSetAttribute(sv_circuit_kernel)");
  if ((e = cudaFuncSetAttribute(sv_sweep_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (16 << kSvTileBitsMax) + kBlockBytes + 1024)) != cudaSuccess)
    return bail(e, "cudaFuncSetAttribute(sv_sweep_kernel)");
  if ((e = cudaFuncSetAttribute(sv_sweep_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (16 << 11) + kBlockBytes + 1024)) != cudaSuccess)
    return bail(e, "cudaFuncSetAttribute(sv_sweep_kernel<128>)");
  {
    std::vector<uint32_t> tab((size_t)kB0Pairs * kB0Groups);
    fill_b0_table(tab.data());
    if ((e = ctx->d_b0.reserve(tab.size() * sizeof(uint32_t))) != cudaSuccess) return bail(e, "cudaMalloc(b0 table)");
    if ((e = cudaMemcpy(ctx->d_b0.p, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess)
      return bail(e, "cudaMemcpy(b0 table)");
  }
  {
    std::vector<uint32_t> tab((size_t)kTmaPairs * kTmaThreads);
    fill_tma_a_table(tab.data());
    if ((e = ctx->d_tma_a.reserve(tab.size() * sizeof(uint32_t))) != cudaSuccess) return bail(e, "cudaMalloc(tma table)");
    if ((e = cudaMemcpy(ctx->d_tma_a.p, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess)
      return bail(e, "cudaMemcpy(tma table)");
    if ((e = ctx->d_counters.reserve(sizeof(unsigned int) * kMaxPersistLaunches)) != cudaSuccess) return bail(e, "cudaMalloc(counters)");
    if ((e = cudaFuncSetAttribute(dm_sweep_tma_persistent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaPersistSmem)) != cudaSuccess)
      return bail(e, "cudaFuncSetAttribute(dm_sweep_tma_persistent_kernel)");
    if ((e = cudaFuncSetAttribute(dm_sweep_tma_persistent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaPersistSmem)) != cudaSuccess)
      return bail(e, "cudaFuncSetAttribute(dm_sweep_tma_persistent_kernel<full>)");
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      ctx->encode_tiled = (PFN_cuTensorMapEncodeTiled)fn;
    else
      cudaGetLastError();
  }
  *out = ctx;
  return BWQ_OK;
}

extern "C" int bwq_destroy(bwq_ctx* ctx) {
  if (!ctx) return BWQ_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->d_b0.release(); ctx->d_noise.release(); ctx->d_states.release(); ctx->d_out.release();
  for (auto& sl : ctx->dm) { sl.d_prog.release(); sl.h_prog.release(); sl.d_maps.release(); sl.h_maps.release(); if (sl.h2d_done) cudaEventDestroy(sl.h2d_done); }
  ctx->d_tma_a.release(); ctx->d_counters.release();
  ctx->d_oc.release(); ctx->d_oc_out.release(); ctx->d_oc_noise.release(); ctx->h_oc.release(); ctx->h_oc_out.release();
  if (ctx->oc_copy_stream) {
    cudaStreamDestroy(ctx->oc_copy_stream);
    for (int i = 0; i < 8; ++i) { cudaEventDestroy(ctx->oc_copied[i]); cudaEventDestroy(ctx->oc_k0[i]); cudaEventDestroy(ctx->oc_k1[i]); }
  }
  if (ctx->companion) bwq_destroy(ctx->companion);
  ctx->d_scratch.release(); ctx->h_out.release();
  ctx->d_sv_prog.release(); ctx->h_sv_prog.release();
  ctx->d_wide_prog.release(); ctx->h_wide_prog.release(); ctx->d_partial.release();
  for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->chunk_ev) if (ev) cudaEventDestroy(ev);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return BWQ_OK;
}

extern "C" int bwq_set_options(bwq_ctx* ctx, const bwq_options* opt) {
  if (!ctx) return BWQ_ERR_ARG;
  if (!opt) { ctx->opt = bwq_options{}; return BWQ_OK; }
  if (opt->tile_qubits && (opt->tile_qubits < 2 || opt->tile_qubits > kMaxTileQubits))
    return fail(ctx, BWQ_ERR_ARG, "tile_qubits must be in [2,%d]", kMaxTileQubits);
  if (opt->low_qubits < 0 || opt->low_qubits > kMaxTileQubits) return fail(ctx, BWQ_ERR_ARG, "low_qubits out of range");
  ctx->opt = *opt;
  return BWQ_OK;
}

extern "C" int bwq_set_noise_table(bwq_ctx* ctx, const bwq_noise_table* table) {
  if (!ctx) return BWQ_ERR_ARG;
  char err[256] = {0};
  int rc = ctx->noise.set(table, err, sizeof err);
  ctx->oc_noise_valid = false;
  if (rc) return fail(ctx, rc, "%s", err);
  CK(cudaSetDevice(ctx->device));
  if (!ctx->noise.data.empty()) {
    CK(ctx->d_noise.reserve(ctx->noise.data.size() * sizeof(double)));
    CK(cudaMemcpyAsync(ctx->d_noise.p, ctx->noise.data.data(), ctx->noise.data.size() * sizeof(double),
                       cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return BWQ_OK;
}

extern "C" int bwq_get_stats(const bwq_ctx* ctx, bwq_stats* out) {
  if (!ctx || !out) return BWQ_ERR_ARG;
  *out = ctx->stats;
  return BWQ_OK;
}

extern "C" int bwq_sync(bwq_ctx* ctx) {
  if (!ctx) return BWQ_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return BWQ_OK;
}

static int check_batch(bwq_ctx* ctx, const bwq_batch* b, const void* out, const void* status) {
  if (!b || !out || !status) return fail(ctx, BWQ_ERR_ARG, "null batch/out/status");
  if (b->n_circuits < 0) return fail(ctx, BWQ_ERR_ARG, "n_circuits < 0");
  if (b->n_circuits == 0) return BWQ_OK;
  if (!b->n_qubits || !b->op_offsets || !b->obs_offsets || !b->term_offsets)
    return fail(ctx, BWQ_ERR_ARG, "batch: null offset arrays");
  for (int c = 0; c < b->n_circuits; ++c) {
    if (b->op_offsets[c + 1] < b->op_offsets[c] || b->obs_offsets[c + 1] < b->obs_offsets[c])
      return fail(ctx, BWQ_ERR_ARG, "batch: offsets of circuit %d not monotone", c);
  }
  int64_t n_obs = b->obs_offsets[b->n_circuits];
  for (int64_t o = 0; o < n_obs; ++o)
    if (b->term_offsets[o + 1] < b->term_offsets[o]) return fail(ctx, BWQ_ERR_ARG, "batch: term_offsets not monotone");
  if (b->op_offsets[b->n_circuits] > 0 && !b->ops) return fail(ctx, BWQ_ERR_ARG, "batch: ops is NULL");
  if (n_obs > 0 && b->term_offsets[n_obs] > 0 && (!b->term_x || !b->term_z || !b->term_coeff))
    return fail(ctx, BWQ_ERR_ARG, "batch: term arrays NULL");
  return BWQ_OK;
}

// Threads are spawned per call.  A persistent pool parked on a condition variable was measured and
// dropped: it saves ~3 ms of lowering per cfg2 call on a quiet host (6.5 vs 9.5 ms, hidden behind the
// sweeps anyway) but on two ranks sharing 16 cores one run in three fell into an 81 ms-per-call mode
// (47 ms with spawned threads; same box, alternating runs of both builds, commit 676d3d0 has the pool).
template <class F> static void parallel_for(bwq_ctx* /*ctx*/, int n, int threads, F f) {
  threads = std::max(1, std::min(threads, n));
  if (threads == 1) { for (int i = 0; i < n; ++i) f(i); return; }
  std::atomic<int> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&] { for (;;) { int i = next.fetch_add(16); if (i >= n) break; for (int j = i; j < std::min(n, i + 16); ++j) f(j); } });
  for (auto& th : pool) th.join();
}

static int host_threads(const bwq_ctx* ctx) {
  if (ctx->opt.host_threads > 0) return ctx->opt.host_threads;
  unsigned hc = std::thread::hardware_concurrency();
  return hc ? (int)std::min(hc, 32u) : 4;
}

template <int KQ, bool FULL> static cudaError_t launch_sweep(const DmLaunch& L, int sweep, int64_t n_cta, cudaStream_t s) {
  const size_t dyn = KQ <= 6 ? 0 : (sizeof(double) << (2 * KQ)) + kBlockBytes;  // KQ <= 6: static shared memory
  dm_sweep_kernel<KQ, FULL><<<(unsigned)n_cta, SweepCfg<KQ>::kThreads, dyn, s>>>(L, sweep);
  return cudaGetLastError();
}

template <bool FULL> static cudaError_t launch_sweep_kq(int kq, const DmLaunch& L, int sweep, int64_t n_cta, cudaStream_t s) {
  switch (kq) {
    case 2: return launch_sweep<2, FULL>(L, sweep, n_cta, s);
    case 3: return launch_sweep<3, FULL>(L, sweep, n_cta, s);
    case 4: return launch_sweep<4, FULL>(L, sweep, n_cta, s);
    case 5: return launch_sweep<5, FULL>(L, sweep, n_cta, s);
    case 6: return launch_sweep<6, FULL>(L, sweep, n_cta, s);
    case 7: return launch_sweep<7, FULL>(L, sweep, n_cta, s);
    default: return cudaErrorInvalidValue;
  }
}

// ------------------------------------------------------------------------------------------------
// density-matrix run = prepare (K0 lowering on host threads + one H2D of the program) + execute
// (sweeps, expectation values, D2H of the values).  The prepared plan stays in the ctx so that
// bwq_dm_execute can be repeated with the program resident in HBM.
// ------------------------------------------------------------------------------------------------
// Host part: lowers circuits [c0, c1) of the batch into slot `sl` (program blob in pinned memory,
// chunk plan).  Output indices are relative to the first observable of c0.  Touches only the slot,
// so it may run on a helper thread while the GPU executes the other slot.
// The batch is the expansion of `base` into n_folds fold variants per circuit (variant-major inside a
// circuit, no twirls): circuit i of the batch = base circuit i / n_folds at factor folds[i % n_folds].
struct FoldInfo {
  const bwq_batch* base;
  const int32_t* folds;
  int n_folds;
};

static int dm_lower_impl(bwq_ctx* ctx, bwq_ctx::DmSlot& sl, const bwq_batch* b, int c0, int c1, int32_t* out_status,
                         int64_t budget, const FoldInfo* fi = nullptr) {
  NvtxRange nvtx_range("bwq:dm_lower (K0)");
  DmPlan& P = sl.plan;
  P = DmPlan();
  const int N = c1 - c0;
  const int64_t obs0 = N > 0 ? b->obs_offsets[c0] : 0;
  P.n_obs = N > 0 ? b->obs_offsets[c1] - obs0 : 0;
  P.valid = true;
  if (N == 0) return BWQ_OK;
  CK(cudaSetDevice(ctx->device));

  // ---- K0: lowering (host threads)
  double t0 = now_ms();
  LowerOptions lo;
  lo.tile_qubits = ctx->opt.tile_qubits ? ctx->opt.tile_qubits : 6;
  lo.low_qubits = ctx->opt.low_qubits > 0 ? ctx->opt.low_qubits : 2;
  lo.direct = (ctx->opt.flags & BWQ_OPT_NO_DIRECT_LOAD ? 0 : kPassLoadDirect) | (ctx->opt.flags & BWQ_OPT_NO_DIRECT_STORE ? 0 : kPassStoreDirect);
  lo.tma = ctx->encode_tiled != nullptr && !(ctx->opt.flags & BWQ_OPT_NO_TMA);
  lo.tma_direct_store = (ctx->opt.flags & BWQ_OPT_TMA_DIRECT_STORE) != 0;
  P.tile_qubits = lo.tile_qubits;
  std::vector<CircuitProgram> progs(N);  // progs[c] <-> batch circuit c0 + c
  if (fi && fi->n_folds > 0 && c0 % fi->n_folds == 0 && N % fi->n_folds == 0) {
    // fold variants: the gates of a base circuit are lowered once, every fold re-packs the passes
    const int nf = fi->n_folds;
    parallel_for(ctx, N / nf, host_threads(ctx), [&](int k) {
      const int base = c0 / nf + k;
      if (!lower_dm_circuit_folds(ctx->noise, *fi->base, base, lo, fi->folds, nf, &progs[(size_t)k * nf]))
        for (int f = 0; f < nf; ++f) lower_dm_circuit(ctx->noise, *b, c0 + k * nf + f, lo, &progs[(size_t)k * nf + f]);
    });
  } else {
    parallel_for(ctx, N, host_threads(ctx), [&](int c) { lower_dm_circuit(ctx->noise, *b, c0 + c, lo, &progs[c]); });
  }

  std::vector<int> order;
  order.reserve(N);
  for (int c = 0; c < N; ++c) {
    out_status[c0 + c] = progs[c].status;
    if (progs[c].status == 0 && !progs[c].sweeps.empty()) order.push_back(c);
  }
  // wide first, then by sweep count (so the circuits of a chunk finish together)
  std::sort(order.begin(), order.end(), [&](int a, int c) {
    if (progs[a].n_digits != progs[c].n_digits) return progs[a].n_digits > progs[c].n_digits;
    if (progs[a].sweeps.size() != progs[c].sweeps.size()) return progs[a].sweeps.size() > progs[c].sweeps.size();
    return a < c;
  });
  const int M = (int)order.size();

  // ---- merge the programs into one blob (sorted order)
  std::vector<int64_t> sw_off(M + 1, 0), pg_off(M + 1, 0), tm_off(M + 1, 0);
  P.ob_off.assign(M + 1, 0);
  for (int i = 0; i < M; ++i) {
    const CircuitProgram& p = progs[order[i]];
    sw_off[i + 1] = sw_off[i] + (int64_t)p.sweeps.size();
    pg_off[i + 1] = pg_off[i] + (int64_t)p.prog.size();
    tm_off[i + 1] = tm_off[i] + (int64_t)p.term_index.size();
    P.ob_off[i + 1] = P.ob_off[i] + (b->obs_offsets[c0 + order[i] + 1] - b->obs_offsets[c0 + order[i]]);
  }
  if (sw_off[M] > INT32_MAX || pg_off[M] / 2 >= (int64_t(1) << 32))
    return fail(ctx, BWQ_ERR_ARG, "batch too large for 32-bit program indices; split the batch");
  Blob blob;
  P.o_sweeps = blob.add(sizeof(SweepDesc) * (size_t)sw_off[M]);
  P.o_prog = blob.add(sizeof(uint64_t) * (size_t)pg_off[M]);
  P.o_tidx = blob.add(sizeof(int64_t) * (size_t)tm_off[M]);
  P.o_tcoef = blob.add(sizeof(double) * (size_t)tm_off[M]);
  P.o_obs = blob.add(sizeof(int64_t) * 4 * (size_t)P.ob_off[M]);
  P.blob_bytes = blob.total;
  if (M > 0) CK(sl.h_prog.reserve(blob.total));
  char* hb = (char*)sl.h_prog.p;
  parallel_for(ctx, M, host_threads(ctx), [&](int i) {
    const int c = order[i];
    const CircuitProgram& p = progs[c];
    SweepDesc* sw = (SweepDesc*)(hb + P.o_sweeps) + sw_off[i];
    for (size_t k = 0; k < p.sweeps.size(); ++k) {
      sw[k] = p.sweeps[k];
      sw[k].blk_q16 += (uint32_t)(pg_off[i] / 2);
    }
    if (!p.prog.empty()) std::memcpy((uint64_t*)(hb + P.o_prog) + pg_off[i], p.prog.data(), p.prog.size() * sizeof(uint64_t));
    if (!p.term_index.empty()) {
      std::memcpy((int64_t*)(hb + P.o_tidx) + tm_off[i], p.term_index.data(), p.term_index.size() * sizeof(int64_t));
      std::memcpy((double*)(hb + P.o_tcoef) + tm_off[i], p.term_coeff.data(), p.term_coeff.size() * sizeof(double));
    }
    // observables: {term_begin, term_end, chunk-local slot (filled below), out_index}
    int64_t* od = (int64_t*)(hb + P.o_obs) + 4 * P.ob_off[i];
    const int64_t ob0 = b->obs_offsets[c0 + c], ob1 = b->obs_offsets[c0 + c + 1];
    const int64_t tb = b->term_offsets[ob0];
    for (int64_t o = ob0; o < ob1; ++o) {
      int64_t* d = od + 4 * (o - ob0);
      d[0] = tm_off[i] + (b->term_offsets[o] - tb);
      d[1] = tm_off[i] + (b->term_offsets[o + 1] - tb);
      d[2] = i;
      d[3] = o - obs0;
    }
  });
  for (int i = 0; i < M; ++i) { P.n_gates += progs[order[i]].n_gates; P.n_passes += progs[order[i]].n_passes; }

  // ---- chunk plan: circuits of equal width, as many resident states as the budget allows
  for (int i = 0; i < M;) {
    const int nd = progs[order[i]].n_digits;
    const int64_t sbytes = (int64_t)sizeof(double) << (2 * nd);
    int64_t fit = budget / sbytes;
    if (fit < 1) {
      int j = i;
      while (j < M && progs[order[j]].n_digits == nd) out_status[c0 + order[j++]] = BWQ_CIRC_TOO_WIDE;
      i = j;
      continue;
    }
    if (ctx->opt.chunk_circuits > 0) fit = std::min<int64_t>(fit, ctx->opt.chunk_circuits);
    const int kq = std::min(std::min(nd, std::max(lo.tile_qubits, 3)), kMaxTileQubits);
    const int64_t tiles = int64_t(1) << (2 * (nd - kq));
    fit = std::min<int64_t>(fit, (int64_t(1) << 30) / tiles);  // grid.x limit
    int j = i;
    while (j < M && progs[order[j]].n_digits == nd && j - i < fit) ++j;
    DmChunk ch;
    ch.first = i; ch.count = j - i; ch.nd = nd; ch.kq = kq;
    // circuits are sorted by sweep count (descending) inside a width group, so sweep s only needs
    // the leading circuits that still have an s-th sweep
    const size_t max_sweeps = progs[order[i]].sweeps.size();
    int live = ch.count;
    for (size_t sidx = 0; sidx < max_sweeps; ++sidx) {
      while (live > 0 && progs[order[i + live - 1]].sweeps.size() <= sidx) --live;
      ch.live.push_back(live);
    }
    for (int k = i; k < j; ++k) ch.full = ch.full || progs[order[k]].needs_dense;
    ch.tma = progs[order[i]].tma;  // a function of the options and the width: uniform inside a chunk
    // bytes per launch: a tile with X/Y on an outside digit no pass has touched yet is all zero
    // (kernels.cuh): the first sweep stores it (8 B/element, no read), later sweeps skip it
    ch.bytes.assign(max_sweeps, 0);
    for (int k = i; k < j; ++k) {
      const CircuitProgram& p = progs[order[k]];
      for (size_t sidx = 0; sidx < p.sweeps.size(); ++sidx) {
        uint32_t outside = (nd >= 32 ? ~0u : ((1u << nd) - 1u));
        for (int s2 = 0; s2 < kq; ++s2) outside &= ~(1u << p.sweeps[sidx].pos[s2]);
        const int u = __builtin_popcount(outside & (p.sweeps[sidx].blk_len_q16 >> 16));
        const int64_t full = (int64_t)sizeof(double) << (2 * nd);
        ch.bytes[sidx] += sidx == 0 ? full + (full >> u) * 0 : 2 * (full >> u);
      }
    }
    // sweep-major descriptor table of the chunk: launch s reads one descriptor per circuit slot
    // (circuits with an s-th sweep are the leading `live[s]` slots), a single lookup per CTA
    {
      SweepDesc* all = (SweepDesc*)(hb + P.o_sweeps);
      const int64_t base = sw_off[i];
      std::vector<SweepDesc> tmp(all + base, all + sw_off[j]);
      int64_t off = base;
      for (size_t sidx = 0; sidx < max_sweeps; ++sidx) {
        ch.desc_off.push_back(off);
        for (int slot = 0; slot < ch.live[sidx]; ++slot) all[off++] = tmp[(sw_off[i + slot] - base) + (int64_t)sidx];
      }
    }
    if (ch.tma) {
      // tensor maps: one per ordered tuple of upper digit positions (box order) and window of
      // circuit slots; the sweep descriptors get the 16-bit map id in place of pos[6..7]
      ch.window_log2 = std::max(0, 31 - 2 * nd);
      if (const char* wenv = std::getenv("BWQ_TMA_WINDOW_LOG2"))  // tests: several windows at small widths
        ch.window_log2 = std::max(0, std::min(ch.window_log2, std::atoi(wenv)));
      ch.n_windows = (ch.count + (1 << ch.window_log2) - 1) >> ch.window_log2;
      ch.map_first = P.n_maps;
      SweepDesc* all = (SweepDesc*)(hb + P.o_sweeps);
      for (int64_t d = sw_off[i]; d < sw_off[j]; ++d) {
        SweepDesc& sd = all[d];
        uint32_t key = 0;
        for (int k = 0; k < 4; ++k) key |= uint32_t(sd.pos[2 + ((sd.pos[6] >> (2 * k)) & 3)]) << (8 * k);
        size_t id = 0;
        while (id < ch.map_keys.size() && ch.map_keys[id] != key) ++id;
        if (id == ch.map_keys.size()) ch.map_keys.push_back(key);
        sd.pos[6] = (uint8_t)(id & 0xff);
        sd.pos[7] = (uint8_t)(id >> 8);
      }
      if (ch.map_keys.size() >= 65536) return fail(ctx, BWQ_ERR_ARG, "too many distinct tile layouts in one chunk");
      P.n_maps += ch.map_keys.size() * (size_t)ch.n_windows;
    }
    P.chunks.push_back(std::move(ch));
    i = j;
  }
  for (auto& ch : P.chunks) {
    P.max_chunk_bytes = std::max(P.max_chunk_bytes, ((int64_t)sizeof(double) << (2 * ch.nd)) * ch.count);
    for (int i = ch.first; i < ch.first + ch.count; ++i) {
      int64_t* od = (int64_t*)(hb + P.o_obs) + 4 * P.ob_off[i];
      for (int64_t k = 0; k < P.ob_off[i + 1] - P.ob_off[i]; ++k) od[4 * k + 2] = i - ch.first;
    }
  }
  // circuits the GPU does not touch: failures (NaN) and circuits without any gate (state stays
  // |0..0>: <P> = 1 for I/Z strings, 0 otherwise)
  for (int c = 0; c < N; ++c) {
    const bool gpu = out_status[c0 + c] == 0 && !progs[c].sweeps.empty();
    if (gpu) continue;
    for (int64_t o = b->obs_offsets[c0 + c]; o < b->obs_offsets[c0 + c + 1]; ++o) {
      double v = 0.0;
      if (out_status[c0 + c] == 0)
        for (int64_t t = b->term_offsets[o]; t < b->term_offsets[o + 1]; ++t)
          if (b->term_x[t] == 0) v += b->term_coeff[t];
      P.host_fix.push_back({o - obs0, out_status[c0 + c] == 0 ? v : std::nan("")});
    }
  }
  // thousands of small circuits leave ~10 heap blocks each: release them on the host threads too
  if (N >= 1024) parallel_for(ctx, N, host_threads(ctx), [&](int c) { progs[c] = CircuitProgram(); });
  P.lower_ms = now_ms() - t0;
  return BWQ_OK;
}

static int dm_encode_maps(bwq_ctx* ctx, bwq_ctx::DmSlot& sl);

// Device part of the preparation: buffers + one H2D copy of the slot's program blob.
// sync = false (pipelined run): the value buffers were sized for the whole batch by the caller and
// the copy is only enqueued; sl.h2d_done tells the lowering thread when the pinned blob is free.
static int dm_upload_impl(bwq_ctx* ctx, bwq_ctx::DmSlot& sl, bool sync) {
  NvtxRange nvtx_range("bwq:dm_upload");
  DmPlan& P = sl.plan;
  const size_t blob_total = P.blob_bytes;
  const char* hb = (const char*)sl.h_prog.p;
  const bool have = !P.chunks.empty();
  if (have) CK(sl.d_prog.reserve(blob_total));
  if (P.max_chunk_bytes > 0) CK(ctx->d_states.reserve((size_t)P.max_chunk_bytes));
  if (sync && P.n_obs > 0) {
    CK(ctx->d_out.reserve(sizeof(double) * (size_t)P.n_obs));
    CK(ctx->h_out.reserve(sizeof(double) * (size_t)P.n_obs));
  }
  cudaStream_t st = ctx->stream;
  if (sync) CK(cudaEventRecord(ctx->ev[0], st));
  if (have) {
    // the pinned map staging may still feed the copy of this slot's previous segment
    if (!sync && P.n_maps) CK(cudaEventSynchronize(sl.h2d_done));
    int rc = dm_encode_maps(ctx, sl);
    if (rc) return rc;
  }
  if (have) CK(cudaMemcpyAsync(sl.d_prog.p, hb, blob_total, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(sync ? ctx->ev[1] : sl.h2d_done, st));
  if (sync) {
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    P.h2d_ms = ms;
  }
  return BWQ_OK;
}

// Tensor maps of the slot's TMA chunks, encoded for the current state buffer: rank 5, dim 0 = the
// whole window of states (stride 1: the tile's 128-byte runs start at any multiple of 16 elements),
// dims 1..4 = one resident digit each (extent 4, stride 8 * 4^pos bytes), box 16 x 4 x 4 x 4 x 4,
// 128-byte swizzle.  Uploaded on the engine stream (pinned staging).
static int dm_encode_maps(bwq_ctx* ctx, bwq_ctx::DmSlot& sl) {
  DmPlan& P = sl.plan;
  if (P.n_maps == 0) return BWQ_OK;
  if (!ctx->encode_tiled) return fail(ctx, BWQ_ERR_UNSUPPORTED, "TMA program without cuTensorMapEncodeTiled");
  CK(sl.h_maps.reserve(P.n_maps * sizeof(CUtensorMap)));
  CK(sl.d_maps.reserve(P.n_maps * sizeof(CUtensorMap)));
  CUtensorMap* hm = (CUtensorMap*)sl.h_maps.p;
  for (const DmChunk& ch : P.chunks) {
    if (!ch.tma) continue;
    const uint64_t state_elems = uint64_t(1) << (2 * ch.nd);
    for (size_t id = 0; id < ch.map_keys.size(); ++id)
      for (int w = 0; w < ch.n_windows; ++w) {
        const uint32_t key = ch.map_keys[id];
        cuuint64_t gdim[5] = {state_elems << ch.window_log2, 4, 4, 4, 4};
        cuuint64_t gstr[4];
        for (int k = 0; k < 4; ++k) gstr[k] = cuuint64_t(8) << (2 * ((key >> (8 * k)) & 0xffu));
        cuuint32_t box[5] = {16, 4, 4, 4, 4}, estr[5] = {1, 1, 1, 1, 1};
        void* base = (double*)ctx->d_states.p + (uint64_t(w) << ch.window_log2) * state_elems;
        CUresult r = ctx->encode_tiled(&hm[ch.map_first + id * ch.n_windows + w], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, base, gdim, gstr, box,
                                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(ctx, BWQ_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
      }
  }
  CK(cudaMemcpyAsync(sl.d_maps.p, hm, P.n_maps * sizeof(CUtensorMap), cudaMemcpyHostToDevice, ctx->stream));
  P.maps_for = ctx->d_states.p;
  return BWQ_OK;
}

static int64_t dm_state_budget(bwq_ctx* ctx, int64_t* out, const bwq_batch* b = nullptr) {
  if (ctx->opt.max_state_bytes > 0) { *out = ctx->opt.max_state_bytes; return BWQ_OK; }
  // cudaMemGetInfo costs 0.3-2 ms and sits in front of the first launch of every call: its answer is
  // kept for two seconds for back-to-back calls with a batch of the SAME shape (a generation loop: same
  // plan, so the state buffer already has its size and nothing is allocated against the stale figure)
  const double t = now_ms();
  const uint64_t key = b && b->n_circuits > 0 ? ((uint64_t)b->n_circuits * 0x9E3779B97F4A7C15ull) ^ (uint64_t)b->op_offsets[b->n_circuits] : 0;
  if (key && ctx->budget_cache > 0 && ctx->budget_key == key && ctx->budget_cap == ctx->d_states.cap && t - ctx->budget_time_ms < 2000.0) {
    *out = ctx->budget_cache;
    return BWQ_OK;
  }
  ctx->budget_key = key;
  size_t free_b = 0, total_b = 0;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemGetInfo(&free_b, &total_b));
  *out = (int64_t)((free_b + ctx->d_states.cap) * 0.8);
  ctx->budget_cache = *out; ctx->budget_cap = ctx->d_states.cap; ctx->budget_time_ms = t;
  return BWQ_OK;
}

static int dm_prepare_impl(bwq_ctx* ctx, const bwq_batch* b, int32_t* out_status) {
  if (!ctx) return BWQ_ERR_ARG;
  int rc = check_batch(ctx, b, out_status, out_status);
  if (rc) return rc;
  ctx->stats = bwq_stats{};
  int64_t budget = 0;
  if ((rc = (int)dm_state_budget(ctx, &budget))) return rc;
  bwq_ctx::DmSlot& sl = ctx->dm[0];
  if ((rc = dm_lower_impl(ctx, sl, b, 0, b->n_circuits, out_status, budget))) return rc;
  if (b->n_circuits > 0 && (rc = dm_upload_impl(ctx, sl, true))) return rc;
  const DmPlan& P = sl.plan;
  ctx->stats.lower_ms = P.lower_ms;
  ctx->stats.h2d_ms = P.h2d_ms;
  ctx->stats.h2d_bytes = (int64_t)P.blob_bytes;
  ctx->stats.n_gates = P.n_gates;
  ctx->stats.n_passes = P.n_passes;
  return BWQ_OK;
}

// deferred = true (pipelined run): segment values go to offset obs_off of the batch-sized buffers,
// nothing is synchronised, timed or copied to the caller here (dm_run_impl does that once).
static int dm_execute_impl(bwq_ctx* ctx, bwq_ctx::DmSlot& sl, double* out_vals, bool out_on_device, bool deferred = false,
                           int64_t obs_off = 0) {
  NvtxRange nvtx_range("bwq:dm_execute (sweeps + expval)");
  if (!ctx || !out_vals) return BWQ_ERR_ARG;
  DmPlan& P = sl.plan;
  if (!P.valid) return fail(ctx, BWQ_ERR_ARG, "bwq_dm_execute: no prepared batch (call bwq_dm_prepare first)");
  CK(cudaSetDevice(ctx->device));
  bwq_stats& S = ctx->stats;
  S = bwq_stats{};
  S.lower_ms = P.lower_ms; S.h2d_ms = P.h2d_ms; S.h2d_bytes = (int64_t)P.blob_bytes;
  S.n_gates = P.n_gates; S.n_passes = P.n_passes;
  cudaStream_t st = ctx->stream;
  double* d_out = out_on_device ? out_vals : (double*)ctx->d_out.p + obs_off;
  const char* db = (const char*)sl.d_prog.p;
  // the state buffer moved since the maps were encoded (another prepare grew it): encode again
  if (P.n_maps && P.maps_for != ctx->d_states.p) {
    CK(cudaStreamSynchronize(st));
    int rc = dm_encode_maps(ctx, sl);
    if (rc) return rc;
  }
  if (!deferred) CK(cudaEventRecord(ctx->ev[1], st));
  if (P.n_obs > 0) CK(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)P.n_obs, st));
  int n_persist = 0;
  const bool persist_ok = (ctx->opt.flags & BWQ_OPT_PERSIST) != 0;
  if (P.n_maps && persist_ok) CK(cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(unsigned int) * kMaxPersistLaunches, st));
  const size_t n_ev = deferred ? 0 : std::min(P.chunks.size(), ctx->chunk_ev.size() / 2);
  for (size_t ci = 0; ci < P.chunks.size(); ++ci) {
    const DmChunk& ch = P.chunks[ci];
    DmLaunch L;
    L.states = (double*)ctx->d_states.p;
    L.stride = int64_t(1) << (2 * ch.nd);
    L.n_digits = ch.nd;
    const int pf_opt = (ctx->opt.flags >> 8) & 0xffff;  // kernel experiments: prefetch distance, 0xffff = off
    const int pf_dist = (ch.nd <= ch.kq || pf_opt == 0xffff) ? 0 : (pf_opt ? pf_opt : kDmPrefetchDist);
    L.prog = (const uint4*)(db + P.o_prog);
    L.b0_table = (const uint32_t*)ctx->d_b0.p;
    const int64_t tiles = int64_t(1) << (2 * (ch.nd - ch.kq));
    if (ci < n_ev) CK(cudaEventRecord(ctx->chunk_ev[2 * ci], st));
    for (size_t sidx = 0; sidx < ch.live.size(); ++sidx) {
      const int live = ch.live[sidx];
      L.sweeps = (const SweepDesc*)(db + P.o_sweeps) + ch.desc_off[sidx];
      L.prefetch_dist = sidx == 0 ? 0 : pf_dist;  // the first sweep synthesises |0..0><0..0|, nothing to read
      if (ch.tma) {
        DmTmaLaunch TL;
        TL.states = L.states; TL.stride = L.stride; TL.n_digits = L.n_digits; TL.prefetch_dist = L.prefetch_dist;
        TL.sweeps = L.sweeps; TL.prog = L.prog;
        TL.a_table = (const uint32_t*)ctx->d_tma_a.p;
        TL.maps = (const CUtensorMap*)sl.d_maps.p + ch.map_first;
        TL.n_windows = ch.n_windows; TL.window_log2 = ch.window_log2;
        if (sidx > 0 && persist_ok && n_persist < kMaxPersistLaunches) {
          // resident CTAs (two per SM) pull tiles from this launch's counter; tile k+1 streams in while k is swept
          const unsigned total = (unsigned)(tiles * live);
          const unsigned grid = std::min<unsigned>(total, 2u * (unsigned)ctx->sm_count);
          unsigned int* cnt = (unsigned int*)ctx->d_counters.p + n_persist++;
          if (ch.full) dm_sweep_tma_persistent_kernel<true><<<grid, kTmaPersistThreads, kTmaPersistSmem, st>>>(TL, cnt, total);
          else dm_sweep_tma_persistent_kernel<false><<<grid, kTmaPersistThreads, kTmaPersistSmem, st>>>(TL, cnt, total);
        } else if (ch.full) dm_sweep_tma_kernel<true><<<(unsigned)(tiles * live), kTmaThreads, 0, st>>>(TL, (int)sidx);
        else dm_sweep_tma_kernel<false><<<(unsigned)(tiles * live), kTmaThreads, 0, st>>>(TL, (int)sidx);
        CK(cudaGetLastError());
        S.n_tma_sweep_launches++;
      } else {
        CK(ch.full ? launch_sweep_kq<true>(ch.kq, L, (int)sidx, tiles * live, st)
                   : launch_sweep_kq<false>(ch.kq, L, (int)sidx, tiles * live, st));
      }
      S.n_sweep_launches++;
      S.n_state_sweeps += live;
      S.state_bytes_swept += ch.bytes[sidx];
    }
    if (ci < n_ev) CK(cudaEventRecord(ctx->chunk_ev[2 * ci + 1], st));
    const int64_t nob = P.ob_off[ch.first + ch.count] - P.ob_off[ch.first];
    if (nob > 0) {
      ExpvalLaunch E;
      E.states = L.states;
      E.stride = L.stride;
      E.n_obs = (int32_t)nob;
      E.obs_desc = (const int64_t*)(db + P.o_obs) + 4 * P.ob_off[ch.first];
      E.term_index = (const int64_t*)(db + P.o_tidx);
      E.term_coeff = (const double*)(db + P.o_tcoef);
      E.out = d_out;
      const int wpb = 8;
      dm_expval_kernel<<<(unsigned)((nob + wpb - 1) / wpb), wpb * 32, 0, st>>>(E);
      CK(cudaGetLastError());
      S.n_other_launches++;
    }
  }
  if (!deferred) CK(cudaEventRecord(ctx->ev[2], st));
  if (!out_on_device && P.n_obs > 0) {
    CK(cudaMemcpyAsync((double*)ctx->h_out.p + obs_off, d_out, sizeof(double) * (size_t)P.n_obs, cudaMemcpyDeviceToHost, st));
    S.d2h_bytes = (int64_t)sizeof(double) * P.n_obs;
  }
  if (deferred) return BWQ_OK;
  CK(cudaEventRecord(ctx->ev[3], st));
  CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2])); S.kernel_ms = ms;
  CK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3])); S.d2h_ms = ms;
  if (n_ev == P.chunks.size()) {
    double tot = 0;
    for (size_t ci = 0; ci < n_ev; ++ci) {
      CK(cudaEventElapsedTime(&ms, ctx->chunk_ev[2 * ci], ctx->chunk_ev[2 * ci + 1]));
      tot += ms;
    }
    S.sweep_kernel_ms = tot;
  } else {
    S.sweep_kernel_ms = S.kernel_ms;
  }
  if (!out_on_device) {
    if (P.n_obs > 0) std::memcpy(out_vals, ctx->h_out.p, sizeof(double) * (size_t)P.n_obs);
    for (auto& f : P.host_fix) out_vals[f.first] = f.second;
  } else {
    for (auto& f : P.host_fix) CK(cudaMemcpy(out_vals + f.first, &f.second, sizeof(double), cudaMemcpyHostToDevice));
  }
  return BWQ_OK;
}

extern "C" int bwq_dm_prepare(bwq_ctx* ctx, const bwq_batch* b, int32_t* out_status) {
  if (!out_status) return fail(ctx, BWQ_ERR_ARG, "null status");
  return dm_prepare_impl(ctx, b, out_status);
}
extern "C" int bwq_dm_execute(bwq_ctx* ctx, double* out_vals) { return ctx ? dm_execute_impl(ctx, ctx->dm[0], out_vals, false) : BWQ_ERR_ARG; }
extern "C" int bwq_dm_execute_device_out(bwq_ctx* ctx, double* d_out_vals) { return ctx ? dm_execute_impl(ctx, ctx->dm[0], d_out_vals, true) : BWQ_ERR_ARG; }

// Whole-batch run from host buffers.  Large batches are cut into segments of consecutive circuits
// and software-pipelined over the two program slots: while the GPU sweeps segment k, the host
// threads lower segment k+1 (K0 is otherwise 20-25 % of the call for 10-qubit circuits).  The
// values do not depend on the segmentation (circuits are independent).
// ------------------------------------------------------------------------------------------------
// on-chip path (onchip.cuh): circuits of at most 5 active qubits whose 2-qubit gates are all cx --
// one warp interprets the gate stream of one (circuit, variant); the upload is the raw base batch
// ------------------------------------------------------------------------------------------------
static int onchip_upload_noise(bwq_ctx* ctx) {
  if (ctx->oc_noise_valid) return BWQ_OK;
  ctx->oc_noise = OnchipNoise{};
  const NoiseTable& nt = ctx->noise;
  if (!nt.empty()) {
    const size_t ne = nt.entries.size();
    Blob blob;
    const size_t o_g1 = blob.add(sizeof(int32_t) * 33 * 64), o_cx = blob.add(sizeof(int32_t) * 64 * 64);
    const size_t o_ent = blob.add(sizeof(int2) * ne), o_ok = blob.add(32 * 64), o_fx = blob.add(sizeof(double) * 32 * 64 * 16);
    std::vector<char> h(blob.total, 0);
    int32_t* g1 = (int32_t*)(h.data() + o_g1);
    int32_t* cx = (int32_t*)(h.data() + o_cx);
    int2* ent = (int2*)(h.data() + o_ent);
    auto index_of = [&](const NoiseEntry* e) -> int32_t {
      if (!e) return -1;
      return (int32_t)(((const char*)e - (const char*)&nt.entries[0].second) / (ptrdiff_t)sizeof(nt.entries[0]));
    };
    for (int op = 0; op <= 32; ++op)
      for (int q = 0; q < 64; ++q) {
        const uint16_t opc = op < 32 ? (uint16_t)op : (uint16_t)BWQ_G_UNITARY1;
        const NoiseEntry* e = gate_is_2q(opc) ? nullptr : nt.find(opc, q, 255);
        g1[op * 64 + q] = (e && e->kind == BWQ_NOISE_DENSE1) ? index_of(e) : -1;
      }
    for (int a = 0; a < 64; ++a)
      for (int t = 0; t < 64; ++t) cx[a * 64 + t] = a == t ? -1 : index_of(nt.find(BWQ_G_CX, a, t));
    for (size_t i = 0; i < ne; ++i) ent[i] = make_int2((int)nt.entries[i].second.kind, (int)nt.entries[i].second.off);
    if (!nt.fixed_ok.empty()) {
      std::memcpy(h.data() + o_ok, nt.fixed_ok.data(), std::min<size_t>(nt.fixed_ok.size(), 32 * 64));
      std::memcpy(h.data() + o_fx, nt.fixed_ptm.data(), std::min<size_t>(nt.fixed_ptm.size(), (size_t)32 * 64 * 16) * sizeof(double));
    }
    CK(ctx->d_oc_noise.reserve(blob.total));
    CK(cudaMemcpyAsync(ctx->d_oc_noise.p, h.data(), blob.total, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const char* d = (const char*)ctx->d_oc_noise.p;
    ctx->oc_noise.g1 = (const int32_t*)(d + o_g1);
    ctx->oc_noise.cx = (const int32_t*)(d + o_cx);
    ctx->oc_noise.ent = (const int2*)(d + o_ent);
    ctx->oc_noise.fixed_ok = (const uint8_t*)(d + o_ok);
    ctx->oc_noise.fixed_ptm = (const double*)(d + o_fx);
    ctx->oc_noise.data = (const double*)ctx->d_noise.p;
  }
  ctx->oc_noise_valid = true;
  return BWQ_OK;
}

// Runs the batch (times its variants) on dm_onchip_kernel when every circuit qualifies; *handled =
// false leaves out_vals / out_status for the tile-sweep path.  out_status: one entry per variant
// circuit (circuit-major), out_vals as bwq_dm_run_variants.
static int onchip_run(bwq_ctx* ctx, const bwq_batch* b, const bwq_variants* v, bool noisy, double* out_vals, bool out_on_device,
                      int32_t* out_status, bool* handled, double* out_ideal = nullptr, int32_t* status_ideal = nullptr) {
  NvtxRange nvtx_range("bwq:onchip");
  *handled = false;
  const bool with_ideal = out_ideal != nullptr;  // the same launch also evolves every base circuit without noise
  if (ctx->opt.flags & BWQ_OPT_NO_ONCHIP) return BWQ_OK;
  const int N = b->n_circuits;
  if (N <= 0) return BWQ_OK;
  const int n_folds = v && v->n_folds > 0 ? v->n_folds : 1, n_tw = v && v->n_twirls > 0 ? v->n_twirls : 1;
  if (v && v->n_folds > 0) {
    if (!v->folds) return BWQ_OK;
    for (int f = 0; f < n_folds; ++f)
      if (v->folds[f] < 1 || v->folds[f] % 2 == 0) return BWQ_OK;  // the expansion path reports the error
  }
  const int64_t NV = (int64_t)N * n_folds * n_tw;
  const int64_t g_lo = b->op_offsets[0], g_hi = b->op_offsets[N];
  const int64_t ob_lo = b->obs_offsets[0], ob_hi = b->obs_offsets[N];
  if (NV > INT32_MAX || b->n_params >= (int64_t(1) << 32)) return BWQ_OK;
  // routing probe (the kernel decides per circuit): a few circuits must have at most 5 active qubits
  for (int probe = 0; probe < 3; ++probe) {
    const int c = probe == 0 ? 0 : probe == 1 ? N / 2 : N - 1;
    uint64_t used = 0;
    for (int64_t g = b->op_offsets[c]; g < b->op_offsets[c + 1]; ++g) {
      const bwq_op& op = b->ops[g];
      if (gate_is_2q(op.opcode)) {
        if (op.opcode != BWQ_G_CX) return BWQ_OK;
        used |= 1ull << (op.q1 & 63);
      }
      used |= 1ull << (op.q0 & 63);
    }
    if (__builtin_popcountll(used) > kOnchipMaxDigits) return BWQ_OK;
  }
  const double t0 = now_ms();
  CK(cudaSetDevice(ctx->device));
  int rc = noisy ? onchip_upload_noise(ctx) : BWQ_OK;
  if (rc) return rc;
  const int64_t n_obs = ob_hi - ob_lo, n_out = n_obs * n_folds * n_tw;
  // ---- ranges of circuits: range r+1 is staged into pinned memory and uploaded (copy stream) while the
  // kernel of range r runs, so a very large batch pays for staging + H2D + kernel of ONE range plus
  // the longer of (all copies, all kernels) instead of their sum.  Shared by all ranges: parameters,
  // folds.  Only for batches of >= 64 MiB of ops (~100 k cfg1-sized circuits): a launch lasts at least
  // as long as its deepest circuit's serial chain (~0.3 ms), so cutting cfg1's 6000 circuits into 8
  // ranges was measured at 3.0 ms of kernels instead of 1.0 ms -- ranges must stay far larger than
  // the ~2400 warps the GPU holds.
  constexpr int kMaxRanges = 8;
  int R = (int)std::min<int64_t>(kMaxRanges, std::max<int64_t>(1, ((g_hi - g_lo) * (int64_t)sizeof(bwq_op)) >> 25));  // >= 32 MiB of ops each
  if (const char* re = std::getenv("BWQ_ONCHIP_RANGES")) R = std::max(1, std::min(kMaxRanges, std::atoi(re)));  // tests
  R = std::min(R, N);
  int rc0[kMaxRanges + 1];
  rc0[0] = 0;
  for (int r = 1; r < R; ++r) {
    const int64_t want = g_lo + (g_hi - g_lo) * r / R;
    int c = (int)(std::lower_bound(b->op_offsets, b->op_offsets + N, want) - b->op_offsets);
    rc0[r] = std::max(rc0[r - 1] + 1, std::min(c, N - (R - r)));
  }
  rc0[R] = N;
  struct Rng { size_t o_nq, o_go, o_ops, o_oo, o_to, o_tx, o_tz, o_tc, begin, end; int64_t g0, ob0, t0, nt, nob; };
  Rng rg[kMaxRanges];
  Blob blob;
  const size_t o_par = blob.add(sizeof(double) * (size_t)b->n_params), o_fd = blob.add(sizeof(int32_t) * (size_t)n_folds);
  for (int r = 0; r < R; ++r) {
    const int c0 = rc0[r], c1 = rc0[r + 1], nc = c1 - c0;
    Rng& g = rg[r];
    g.g0 = b->op_offsets[c0]; g.ob0 = b->obs_offsets[c0]; g.nob = b->obs_offsets[c1] - g.ob0;
    g.t0 = g.nob > 0 ? b->term_offsets[g.ob0] : 0; g.nt = g.nob > 0 ? b->term_offsets[g.ob0 + g.nob] - g.t0 : 0;
    g.begin = blob.total;
    g.o_nq = blob.add(sizeof(int32_t) * (size_t)nc); g.o_go = blob.add(sizeof(int64_t) * (size_t)(nc + 1));
    g.o_ops = blob.add(sizeof(bwq_op) * (size_t)(b->op_offsets[c1] - g.g0));
    g.o_oo = blob.add(sizeof(int64_t) * (size_t)(nc + 1)); g.o_to = blob.add(sizeof(int64_t) * (size_t)(g.nob + 1));
    g.o_tx = blob.add(sizeof(uint64_t) * (size_t)g.nt); g.o_tz = blob.add(sizeof(uint64_t) * (size_t)g.nt); g.o_tc = blob.add(sizeof(double) * (size_t)g.nt);
    g.end = blob.total;
  }
  if (blob.total > (size_t(1) << 30)) return BWQ_OK;  // larger batches: the segmented tile-sweep pipeline
  CK(ctx->h_oc.reserve(blob.total));
  CK(ctx->d_oc.reserve(blob.total));
  char* h = (char*)ctx->h_oc.p;
  struct Piece { size_t off; const void* src; size_t bytes; };
  auto stage = [&](const Piece* pieces, int n_pieces) {  // 256 KiB slices over a few host threads when large
    struct Slice { char* dst; const char* src; size_t bytes; };
    std::vector<Slice> slices;
    for (int i = 0; i < n_pieces; ++i) {
      const Piece& pc = pieces[i];
      if (!pc.src || !pc.bytes) continue;
      for (size_t o = 0; o < pc.bytes; o += size_t(1) << 18)
        slices.push_back({h + pc.off + o, (const char*)pc.src + o, std::min(pc.bytes - o, size_t(1) << 18)});
    }
    parallel_for(ctx, (int)slices.size(), std::min<int>({(int)slices.size() / 4, 8, host_threads(ctx)}), [&](int i) { std::memcpy(slices[i].dst, slices[i].src, slices[i].bytes); });
  };
  auto stage_range = [&](int r) {
    const int c0 = rc0[r], nc = rc0[r + 1] - c0;
    const Rng& g = rg[r];
    const Piece pieces[] = {
        {g.o_nq, b->n_qubits + c0, sizeof(int32_t) * (size_t)nc}, {g.o_go, b->op_offsets + c0, sizeof(int64_t) * (size_t)(nc + 1)},
        {g.o_ops, b->ops ? b->ops + g.g0 : nullptr, sizeof(bwq_op) * (size_t)(b->op_offsets[c0 + nc] - g.g0)},
        {g.o_oo, b->obs_offsets + c0, sizeof(int64_t) * (size_t)(nc + 1)}, {g.o_to, b->term_offsets + g.ob0, sizeof(int64_t) * (size_t)(g.nob + 1)},
        {g.o_tx, b->term_x ? b->term_x + g.t0 : nullptr, sizeof(uint64_t) * (size_t)g.nt},
        {g.o_tz, b->term_z ? b->term_z + g.t0 : nullptr, sizeof(uint64_t) * (size_t)g.nt},
        {g.o_tc, b->term_coeff ? b->term_coeff + g.t0 : nullptr, sizeof(double) * (size_t)g.nt}};
    stage(pieces, 8);
  };
  // results: [status int32 x NV | status_ideal x N | values | ideal values]; values straight into the
  // caller's buffer when it is on the device
  const size_t st_bytes = (sizeof(int32_t) * (size_t)(NV + (with_ideal ? N : 0)) + 255) & ~size_t(255);
  const size_t val_bytes = out_on_device ? 0 : ((sizeof(double) * (size_t)n_out + 255) & ~size_t(255));
  const size_t res_bytes = st_bytes + val_bytes + (with_ideal ? sizeof(double) * (size_t)n_obs : 0);
  CK(ctx->d_oc_out.reserve(res_bytes));
  CK(ctx->h_oc_out.reserve(res_bytes));
  if (!ctx->oc_copy_stream) {
    CK(cudaStreamCreateWithFlags(&ctx->oc_copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < kMaxRanges; ++i) {
      CK(cudaEventCreateWithFlags(&ctx->oc_copied[i], cudaEventDisableTiming));
      CK(cudaEventCreate(&ctx->oc_k0[i]));
      CK(cudaEventCreate(&ctx->oc_k1[i]));
    }
  }
  cudaStream_t st = ctx->stream, cs = ctx->oc_copy_stream;
  const char* d = (const char*)ctx->d_oc.p;
  const int n_var = n_folds * n_tw;
  double* d_vals = out_on_device ? out_vals : (double*)((char*)ctx->d_oc_out.p + st_bytes);
  double staged_ms = 0.0;
  {
    const Piece shared[] = {{o_par, b->params, sizeof(double) * (size_t)b->n_params},
                            {o_fd, v && v->n_folds > 0 ? v->folds : nullptr, sizeof(int32_t) * (size_t)n_folds}};
    stage(shared, 2);
    stage_range(0);
    staged_ms += now_ms() - t0;
  }
  CK(cudaEventRecord(ctx->ev[0], cs));
  CK(cudaMemcpyAsync(ctx->d_oc.p, h, rg[0].end, cudaMemcpyHostToDevice, cs));  // shared part + range 0 (contiguous)
  CK(cudaEventRecord(ctx->oc_copied[0], cs));
  for (int r = 0; r < R; ++r) {
    const int c0 = rc0[r], nc = rc0[r + 1] - c0;
    const Rng& g = rg[r];
    OnchipLaunch L{};
    L.n_circuits = nc; L.circuit_base = c0; L.n_folds = n_folds; L.n_twirls = n_tw; L.twirl = v && v->n_twirls > 0; L.seed = v ? v->seed : 0;
    L.folds = v && v->n_folds > 0 ? (const int32_t*)(d + o_fd) : nullptr;
    L.n_qubits = (const int32_t*)(d + g.o_nq);
    L.op_offsets = (const int64_t*)(d + g.o_go);
    L.ops = (const bwq_op*)(d + g.o_ops) - g.g0;  // indexed with the batch's absolute offsets
    L.params = (const double*)(d + o_par);
    L.n_params = b->n_params;
    L.obs_offsets = (const int64_t*)(d + g.o_oo);
    L.term_offsets = (const int64_t*)(d + g.o_to) - g.ob0;
    L.term_x = (const uint64_t*)(d + g.o_tx) - g.t0;
    L.term_z = (const uint64_t*)(d + g.o_tz) - g.t0;
    L.term_coeff = (const double*)(d + g.o_tc) - g.t0;
    if (noisy) L.noise = ctx->oc_noise;
    L.sv_mode = noisy ? 0 : 1;  // the ideal call keeps the statevector path's semantics (reset = unsupported op)
    L.out = d_vals - ob_lo * (int64_t)n_var;
    L.status = (int32_t*)ctx->d_oc_out.p + (int64_t)c0 * n_var;
    L.with_ideal = with_ideal ? 1 : 0;
    L.out_ideal = with_ideal ? (double*)((char*)ctx->d_oc_out.p + st_bytes + val_bytes) - ob_lo : nullptr;
    L.status_ideal = (int32_t*)ctx->d_oc_out.p + NV + c0;
    const int64_t warps_r = (int64_t)nc * (n_var + (with_ideal ? 1 : 0));
    CK(cudaStreamWaitEvent(st, ctx->oc_copied[r], 0));
    CK(cudaEventRecord(ctx->oc_k0[r], st));
    dm_onchip_kernel<<<(unsigned)((warps_r + kOnchipWarps - 1) / kOnchipWarps), 32 * kOnchipWarps, kOnchipSmem, st>>>(L);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->oc_k1[r], st));
    if (r + 1 < R) {  // next range: host staging and its upload overlap this range's kernel
      const double ts = now_ms();
      stage_range(r + 1);
      staged_ms += now_ms() - ts;
      CK(cudaMemcpyAsync((char*)ctx->d_oc.p + rg[r + 1].begin, h + rg[r + 1].begin, rg[r + 1].end - rg[r + 1].begin, cudaMemcpyHostToDevice, cs));
      CK(cudaEventRecord(ctx->oc_copied[r + 1], cs));
    }
  }
  CK(cudaEventRecord(ctx->ev[1], cs));
  const int64_t n_warps = NV + (with_ideal ? N : 0);
  CK(cudaEventRecord(ctx->ev[2], st));
  CK(cudaMemcpyAsync(ctx->h_oc_out.p, ctx->d_oc_out.p, res_bytes, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(ctx->ev[3], st));
  CK(cudaStreamSynchronize(cs));
  CK(cudaStreamSynchronize(st));
  const int32_t* hs = (const int32_t*)ctx->h_oc_out.p;
  for (int64_t i = 0; i < n_warps; ++i)
    if (hs[i] == kOnchipNotHandled) return BWQ_OK;  // mixed batch: everything goes through the tile sweeps
  std::memcpy(out_status, hs, sizeof(int32_t) * (size_t)NV);
  if (!out_on_device && n_out > 0) std::memcpy(out_vals, (const char*)ctx->h_oc_out.p + st_bytes, sizeof(double) * (size_t)n_out);
  if (with_ideal) {
    std::memcpy(status_ideal, hs + NV, sizeof(int32_t) * (size_t)N);
    if (n_obs > 0) std::memcpy(out_ideal, (const char*)ctx->h_oc_out.p + st_bytes + val_bytes, sizeof(double) * (size_t)n_obs);
  }
  bwq_stats S{};
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); S.h2d_ms = ms;  // first to last upload (copy stream; overlaps the kernels)
  for (int r = 0; r < R; ++r) { CK(cudaEventElapsedTime(&ms, ctx->oc_k0[r], ctx->oc_k1[r])); S.kernel_ms += ms; }  // the launches alone
  CK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3])); S.d2h_ms = ms;
  S.lower_ms = staged_ms;
  S.h2d_bytes = (int64_t)blob.total; S.d2h_bytes = (int64_t)res_bytes;
  S.n_other_launches = R;
  S.n_onchip_circuits = n_warps;
  for (int c = 0; c < N; ++c) S.n_gates += b->op_offsets[c + 1] - b->op_offsets[c];
  ctx->stats = S;
  *handled = true;
  return BWQ_OK;
}

static int dm_run_impl(bwq_ctx* ctx, const bwq_batch* b, double* out_vals, int32_t* out_status, bool out_on_device,
                       const FoldInfo* fi = nullptr, bool allow_onchip = true) {
  if (!ctx) return BWQ_ERR_ARG;
  Trace tr("dm_run");
  int rc = check_batch(ctx, b, out_status, out_status);
  if (rc) return rc;
  tr.mark("check_batch");
  if (allow_onchip) {  // circuits that fit on chip: one launch on the raw gate stream, no lowering
    bool handled = false;
    if ((rc = onchip_run(ctx, b, nullptr, true, out_vals, out_on_device, out_status, &handled))) return rc;
    if (handled) return BWQ_OK;
  }
  ctx->stats = bwq_stats{};
  const int N = b->n_circuits;
  int64_t budget = 0;
  if ((rc = (int)dm_state_budget(ctx, &budget, b))) return rc;
  // segments: at least kMinSeg circuits each, at most kMaxSegs of them -- but only when the sweeps
  // are worth hiding behind: a rough estimate of the HBM traffic (8 B x 4^active qubits per sweep,
  // about one sweep per four 2-qubit gates) must reach a few milliseconds of GPU time; batches of
  // tiny circuits (cfg1: 4 qubits, on chip) are lowered in one piece (the per-segment thread and
  // launch overhead would exceed the kernel time)
  constexpr int kMinSeg = 128, kMaxSegs = 8;
  int n_seg = (ctx->opt.flags & BWQ_OPT_NO_PIPELINE) ? 1 : std::max(1, std::min(kMaxSegs, N / kMinSeg));
  if (n_seg > 1 && !(ctx->opt.flags & BWQ_OPT_FORCE_PIPELINE)) {
    // (a sample of at most 64 evenly spaced circuits, scaled: the scan is serial and sits in front of
    // the first segment's lowering)
    double est_bytes = 0.0;
    const int stride = std::max(1, N / 64);
    for (int c = 0; c < N && est_bytes < 2e10; c += stride) {
      uint64_t used = 0;
      int64_t n2 = 0;
      for (int64_t g = b->op_offsets[c]; g < b->op_offsets[c + 1]; ++g) {
        const bwq_op& op = b->ops[g];
        used |= 1ull << (op.q0 & 63);
        if (gate_is_2q(op.opcode)) { used |= 1ull << (op.q1 & 63); ++n2; }
      }
      const int na = std::min(__builtin_popcountll(used), kMaxDmQubits);
      if (na > 6) est_bytes += (double)stride * 16.0 * std::ldexp(1.0, 2 * na) * (1.0 + 0.25 * (double)n2);
    }
    if (est_bytes < 2e10) n_seg = 1;  // < ~5 ms of sweeps
  }
  tr.mark("budget + traffic estimate");
  const int unit = fi && fi->n_folds > 0 && N % fi->n_folds == 0 ? fi->n_folds : 1;  // the variants of a circuit stay in one segment
  // segment 0 is lowered before anything runs on the GPU (exposed host time), so it gets a quarter of
  // a share and segment 1 three quarters: boundaries at (k - 3/4) shares for k >= 1
  auto seg_begin = [&](int k) {
    if (k <= 0) return 0;
    if (k >= n_seg) return N;
    const int64_t units = N / unit;
    const int64_t x = n_seg > 2 ? units * (4 * k - 3) / (4 * n_seg - 6) : units * k / n_seg;  // k = 1: 1/(4n-6) ... k = n-1: (4n-7)/(4n-6)
    return (int)std::min<int64_t>(units, std::max<int64_t>(1, x)) * unit;
  };
  if (n_seg == 1) {
    if ((rc = dm_lower_impl(ctx, ctx->dm[0], b, 0, N, out_status, budget, fi))) return rc;
    if (N > 0 && (rc = dm_upload_impl(ctx, ctx->dm[0], true))) return rc;
    return dm_execute_impl(ctx, ctx->dm[0], out_vals, out_on_device);
  }
  // pipelined: nothing is synchronised between the segments -- upload, sweeps and the D2H of the
  // values of segment k are enqueued behind segment k-1 while the helper thread lowers segment k+1
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t n_obs = b->obs_offsets[N];
  if (!out_on_device && n_obs > 0) {
    CK(ctx->d_out.reserve(sizeof(double) * (size_t)n_obs));
    CK(ctx->h_out.reserve(sizeof(double) * (size_t)n_obs));
  }
  bwq_stats total{};
  std::vector<std::pair<int64_t, double>> fixes;
  if ((rc = dm_lower_impl(ctx, ctx->dm[0], b, seg_begin(0), seg_begin(1), out_status, budget, fi))) return rc;
  tr.mark("lower segment 0 (exposed)");
  const double pre_ms = now_ms() - tr.t0;
  CK(cudaEventRecord(ctx->ev[1], st));
  for (int k = 0; k < n_seg; ++k) {
    bwq_ctx::DmSlot& cur = ctx->dm[k & 1];
    std::thread helper;
    int next_rc = BWQ_OK;
    if (k + 1 < n_seg)
      helper = std::thread([&, k] {
        bwq_ctx::DmSlot& nxt = ctx->dm[(k + 1) & 1];
        cudaSetDevice(ctx->device);
        cudaEventSynchronize(nxt.h2d_done);  // the blob of segment k-1 has left the pinned buffer
        next_rc = dm_lower_impl(ctx, nxt, b, seg_begin(k + 1), seg_begin(k + 2), out_status, budget, fi);
      });
    const int64_t obs0 = b->obs_offsets[seg_begin(k)];
    rc = dm_upload_impl(ctx, cur, false);
    if (!rc) rc = dm_execute_impl(ctx, cur, out_on_device ? out_vals + obs0 : out_vals, out_on_device, true, obs0);
    for (auto& f : cur.plan.host_fix) fixes.push_back({obs0 + f.first, f.second});
    const bwq_stats S = ctx->stats;
    const double seg_lower_ms = cur.plan.lower_ms;
    tr.mark("  upload + launches enqueued");
    if (helper.joinable()) helper.join();
    tr.mark("  next segment lowered (join)");
    if (rc || next_rc) { cudaStreamSynchronize(st); return rc ? rc : next_rc; }
    total.n_sweep_launches += S.n_sweep_launches; total.n_state_sweeps += S.n_state_sweeps;
    total.n_tma_sweep_launches += S.n_tma_sweep_launches;
    total.n_passes += S.n_passes; total.n_gates += S.n_gates; total.state_bytes_swept += S.state_bytes_swept;
    total.n_other_launches += S.n_other_launches; total.lower_ms += seg_lower_ms;
    total.h2d_bytes += S.h2d_bytes; total.d2h_bytes += S.d2h_bytes;
  }
  CK(cudaEventRecord(ctx->ev[2], st));
  CK(cudaStreamSynchronize(st));
  tr.mark("stream synchronised");
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]));
  total.kernel_ms = total.sweep_kernel_ms = ms;  // uploads and value copies of the segments included
  if (!out_on_device) {
    if (n_obs > 0) std::memcpy(out_vals, ctx->h_out.p, sizeof(double) * (size_t)n_obs);
    for (auto& f : fixes) out_vals[f.first] = f.second;
  } else {
    for (auto& f : fixes) CK(cudaMemcpy(out_vals + f.first, &f.second, sizeof(double), cudaMemcpyHostToDevice));
  }
  total.host_pre_ms = pre_ms;
  total.call_wall_ms = now_ms() - tr.t0;
  ctx->stats = total;
  ctx->dm[0].plan.valid = false;  // slot 0 no longer holds a whole prepared batch
  return BWQ_OK;
}

extern "C" int bwq_dm_run(bwq_ctx* ctx, const bwq_batch* b, double* out_vals, int32_t* out_status) {
  if (!out_vals || !out_status) return fail(ctx, BWQ_ERR_ARG, "null out/status");
  return dm_run_impl(ctx, b, out_vals, out_status, false);
}
extern "C" int bwq_dm_run_device_out(bwq_ctx* ctx, const bwq_batch* b, double* d_out_vals, int32_t* out_status) {
  if (!d_out_vals || !out_status) return fail(ctx, BWQ_ERR_ARG, "null out/status");
  return dm_run_impl(ctx, b, d_out_vals, out_status, true);
}

// ------------------------------------------------------------------------------------------------
// wide statevectors (13..31 active qubits on one GPU): tile sweeps, batched per chunk of equal width
// ------------------------------------------------------------------------------------------------
static cudaError_t launch_sv_sweep(const SvxLaunch& L, int sweep, int64_t n_cta, cudaStream_t s) {
  const size_t smem = (sizeof(double2) << L.tile_bits) + kBlockBytes + 1024;  // tile | program | deposit table
  if (L.tile_bits <= 11) sv_sweep_kernel<128><<<(unsigned)n_cta, 128, smem, s>>>(L, sweep);
  else sv_sweep_kernel<256><<<(unsigned)n_cta, 256, smem, s>>>(L, sweep);
  return cudaGetLastError();
}

static int zexp_splits(const bwq_ctx* ctx, int n_groups, int n_local) {
  const int64_t amps = int64_t(1) << n_local;
  int64_t want = (4 * (int64_t)ctx->sm_count + n_groups - 1) / std::max(1, n_groups);
  want = std::min<int64_t>(want, std::max<int64_t>(1, amps / 8192));
  return (int)std::max<int64_t>(1, want);
}

static int sv_wide_prepare(bwq_ctx* ctx, const bwq_batch* b, const std::vector<int>& wide, int32_t* out_status) {
  SvPlan& SP = ctx->sv_plan;
  SvWidePlan& W = SP.wide;
  const int NW = (int)wide.size();
  std::vector<SvxProgram> progs(NW);
  SvxOptions so;
  so.tile_bits = ctx->opt.sv_tile_bits > 0 ? ctx->opt.sv_tile_bits : kSvTileBitsDefault;
  parallel_for(ctx, NW, host_threads(ctx), [&](int i) { lower_svx_circuit(*b, wide[i], so, &progs[i]); });
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  const int64_t budget = ctx->opt.max_state_bytes > 0 ? ctx->opt.max_state_bytes
                                                       : (int64_t)((free_b + ctx->d_states.cap) * 0.8);
  std::vector<int> order;
  for (int i = 0; i < NW; ++i) {
    const int c = wide[i];
    if (progs[i].status == 0 && ((int64_t)sizeof(double2) << progs[i].n_bits) > budget) progs[i].status = BWQ_CIRC_TOO_WIDE;
    out_status[c] = progs[i].status;
    if (progs[i].status == 0) order.push_back(i);
    else for (int64_t o = b->obs_offsets[c]; o < b->obs_offsets[c + 1]; ++o) SP.nan_obs.push_back(o);
  }
  std::sort(order.begin(), order.end(), [&](int a, int c) {
    if (progs[a].n_bits != progs[c].n_bits) return progs[a].n_bits > progs[c].n_bits;
    return a < c;
  });
  const int M = (int)order.size();
  if (M == 0) return BWQ_OK;
  // per circuit: stages = (optional sweep segment, expval segment)
  struct Stage { int sw_first = 0, sw_count = 0, zt_first = 0, zt_count = 0; };
  std::vector<std::vector<Stage>> stages(M);
  std::vector<int64_t> sw_off(M + 1, 0), pg_off(M + 1, 0), zt_off(M + 1, 0);
  for (int i = 0; i < M; ++i) {
    const SvxProgram& p = progs[order[i]];
    Stage cur;
    for (const SvxSegment& sg : p.segs) {
      if (sg.kind == SVSEG_SWEEPS) {  // consecutive sweep segments are contiguous
        if (cur.sw_count == 0) cur.sw_first = sg.first;
        cur.sw_count = sg.first + sg.count - cur.sw_first;
      }
      else if (sg.kind == SVSEG_EXPVAL) { cur.zt_first = sg.first; cur.zt_count = sg.count; stages[i].push_back(cur); cur = Stage(); }
    }
    sw_off[i + 1] = sw_off[i] + (int64_t)p.sweeps.size();
    pg_off[i + 1] = pg_off[i] + (int64_t)p.prog.size();
    zt_off[i + 1] = zt_off[i] + (int64_t)p.zt_mask.size();
    SP.n_gates += p.n_gates;
    W.n_passes += p.n_passes;
    {  // algorithmic bytes: live tiles only; the first sweep writes every tile and reads none
      const int K = p.tile_bits, LBk = std::max(0, K - kSvFreeSlots);
      const int64_t full = (int64_t)sizeof(double2) << p.n_bits;
      for (size_t k = 0; k < p.sweeps.size(); ++k) {
        uint32_t outside = p.n_bits >= 32 ? ~0u : ((1u << p.n_bits) - 1u);
        outside &= ~((1u << LBk) - 1u);
        for (int s2 = 0; s2 < K - LBk; ++s2) outside &= ~(1u << p.sweeps[k].pos[s2]);
        const int u = __builtin_popcount(outside & p.sweep_untouched[k]);
        W.bytes_per_exec += k == 0 ? full : 2 * (full >> u);
      }
      for (const SvxSegment& sg : p.segs)
        if (sg.kind == SVSEG_EXPVAL) W.bytes_per_exec += full * ((sg.count + kZexpTerms - 1) / kZexpTerms);
    }
  }
  if (sw_off[M] > INT32_MAX || pg_off[M] / 2 >= (int64_t(1) << 32) || zt_off[M] > INT32_MAX)
    return fail(ctx, BWQ_ERR_ARG, "statevector batch too large for 32-bit program indices; split the batch");
  // chunks + per-stage tables
  std::vector<int32_t> ranges, gdesc, cdesc, inits;
  for (int i = 0; i < M;) {
    const int nb = progs[order[i]].n_bits;
    const int64_t sbytes = (int64_t)sizeof(double2) << nb;
    int64_t fit = std::max<int64_t>(1, budget / sbytes);
    if (ctx->opt.chunk_circuits > 0) fit = std::min<int64_t>(fit, ctx->opt.chunk_circuits);
    const int tb = progs[order[i]].tile_bits;
    fit = std::min<int64_t>(fit, (int64_t(1) << 30) >> (nb - tb));
    int j = i;
    while (j < M && progs[order[j]].n_bits == nb && j - i < fit) ++j;
    SvWideChunk ch;
    ch.first = i; ch.count = j - i; ch.nb = nb; ch.tile_bits = tb; ch.low_bits = std::max(0, tb - kSvFreeSlots);
    size_t n_st = 0;
    for (int k = i; k < j; ++k) n_st = std::max(n_st, stages[k].size());
    ch.init_off = inits.size();
    for (int k = i; k < j; ++k)
      if (stages[k].empty() || stages[k][0].sw_count == 0) inits.push_back(k - i);
    ch.n_init = (int)(inits.size() - ch.init_off);
    for (size_t f = 0; f < n_st; ++f) {
      SvWideStage st;
      st.range_off = ranges.size();
      st.group_first = (int)(gdesc.size() / 4);
      st.cdesc_first = (int)(cdesc.size() / 4);
      for (int k = i; k < j; ++k) {
        Stage sg = f < stages[k].size() ? stages[k][f] : Stage();
        ranges.push_back((int32_t)(sw_off[k] + sg.sw_first));
        ranges.push_back((int32_t)(sw_off[k] + sg.sw_first + sg.sw_count));
        st.max_sweeps = std::max(st.max_sweeps, sg.sw_count);
        if (sg.zt_count > 0) {
          const int g_first = (int)(gdesc.size() / 4) - st.group_first;
          int ng = 0;
          for (int t0 = 0; t0 < sg.zt_count; t0 += kZexpTerms, ++ng) {
            gdesc.push_back(k - i);
            gdesc.push_back((int32_t)(zt_off[k] + sg.zt_first + t0));
            gdesc.push_back(std::min(kZexpTerms, sg.zt_count - t0));
            gdesc.push_back(0);
          }
          cdesc.push_back(g_first); cdesc.push_back(ng);
          cdesc.push_back((int32_t)b->obs_offsets[wide[order[k]]]); cdesc.push_back(0);
        }
      }
      st.n_groups = (int)(gdesc.size() / 4) - st.group_first;
      st.n_cdesc = (int)(cdesc.size() / 4) - st.cdesc_first;
      ch.stages.push_back(st);
    }
    W.max_state_bytes = std::max(W.max_state_bytes, sbytes * ch.count);
    W.chunks.push_back(std::move(ch));
    i = j;
  }
  if (b->obs_offsets[b->n_circuits] > INT32_MAX) return fail(ctx, BWQ_ERR_ARG, "too many observables");
  Blob blob;
  W.o_range = blob.add(sizeof(int32_t) * ranges.size());
  W.o_sweeps = blob.add(sizeof(SweepDesc) * (size_t)sw_off[M]);
  W.o_unt = blob.add(sizeof(uint32_t) * (size_t)sw_off[M]);
  W.o_prog = blob.add(sizeof(uint64_t) * (size_t)pg_off[M]);
  W.o_ztm = blob.add(sizeof(uint32_t) * (size_t)zt_off[M]);
  W.o_ztc = blob.add(sizeof(double) * (size_t)zt_off[M]);
  W.o_zto = blob.add(sizeof(int32_t) * (size_t)zt_off[M]);
  W.o_gdesc = blob.add(sizeof(int32_t) * gdesc.size());
  W.o_cdesc = blob.add(sizeof(int32_t) * cdesc.size());
  W.o_init = blob.add(sizeof(int32_t) * inits.size());
  W.blob_bytes = blob.total;
  CK(ctx->h_wide_prog.reserve(blob.total));
  CK(ctx->d_wide_prog.reserve(blob.total));
  char* hb = (char*)ctx->h_wide_prog.p;
  if (!ranges.empty()) std::memcpy(hb + W.o_range, ranges.data(), sizeof(int32_t) * ranges.size());
  if (!gdesc.empty()) std::memcpy(hb + W.o_gdesc, gdesc.data(), sizeof(int32_t) * gdesc.size());
  if (!cdesc.empty()) std::memcpy(hb + W.o_cdesc, cdesc.data(), sizeof(int32_t) * cdesc.size());
  if (!inits.empty()) std::memcpy(hb + W.o_init, inits.data(), sizeof(int32_t) * inits.size());
  parallel_for(ctx, M, host_threads(ctx), [&](int i) {
    const SvxProgram& p = progs[order[i]];
    SweepDesc* sw = (SweepDesc*)(hb + W.o_sweeps) + sw_off[i];
    for (size_t k = 0; k < p.sweeps.size(); ++k) { sw[k] = p.sweeps[k]; sw[k].blk_q16 += (uint32_t)(pg_off[i] / 2); }
    if (!p.sweeps.empty()) std::memcpy((uint32_t*)(hb + W.o_unt) + sw_off[i], p.sweep_untouched.data(), p.sweeps.size() * sizeof(uint32_t));
    if (!p.prog.empty()) std::memcpy((uint64_t*)(hb + W.o_prog) + pg_off[i], p.prog.data(), p.prog.size() * sizeof(uint64_t));
    const size_t nt = p.zt_mask.size();
    if (nt) {
      std::memcpy((uint32_t*)(hb + W.o_ztm) + zt_off[i], p.zt_mask.data(), nt * sizeof(uint32_t));
      std::memcpy((double*)(hb + W.o_ztc) + zt_off[i], p.zt_coeff.data(), nt * sizeof(double));
      std::memcpy((int32_t*)(hb + W.o_zto) + zt_off[i], p.zt_obs.data(), nt * sizeof(int32_t));
    }
  });
  CK(ctx->d_states.reserve((size_t)W.max_state_bytes));
  // per-CTA partials of the largest expectation launch
  int64_t max_partial = 0;
  for (const SvWideChunk& ch : W.chunks)
    for (const SvWideStage& st : ch.stages)
      max_partial = std::max<int64_t>(max_partial, (int64_t)st.n_groups * zexp_splits(ctx, st.n_groups, ch.nb) * kZexpTerms);
  if (max_partial > 0) CK(ctx->d_partial.reserve(sizeof(double) * (size_t)max_partial));
  W.any = true;
  return BWQ_OK;
}

static int sv_wide_execute(bwq_ctx* ctx, double* d_out) {
  SvWidePlan& W = ctx->sv_plan.wide;
  bwq_stats& S = ctx->stats;
  S.sv_state_bytes_swept = W.bytes_per_exec;
  cudaStream_t st = ctx->stream;
  const char* db = (const char*)ctx->d_wide_prog.p;
  for (const SvWideChunk& ch : W.chunks) {
    SvxLaunch L{};
    L.states = (double2*)ctx->d_states.p;
    L.stride = int64_t(1) << ch.nb;
    L.n_local = ch.nb; L.tile_bits = ch.tile_bits; L.low_bits = ch.low_bits;
    L.first_circuit = 0;
    L.sweeps = (const SweepDesc*)(db + W.o_sweeps);
    L.sweep_untouched = (const uint32_t*)(db + W.o_unt);
    L.prog = (const uint4*)(db + W.o_prog);
    L.hi_bits = 0;
    if (ch.n_init > 0) {
      dim3 grid((unsigned)((L.stride + 255) / 256), (unsigned)ch.n_init);
      sv_init_kernel<<<grid, 256, 0, st>>>(L.states, L.stride, (const int32_t*)(db + W.o_init) + ch.init_off, ch.n_init, 0u);
      CK(cudaGetLastError());
      S.n_other_launches++;
    }
    const int64_t tiles = int64_t(1) << (ch.nb - ch.tile_bits);
    for (size_t f = 0; f < ch.stages.size(); ++f) {
      const SvWideStage& sg = ch.stages[f];
      L.sweep_range = (const int32_t*)(db + W.o_range) + sg.range_off;
      L.init = f == 0 ? 1 : 0;
      for (int sidx = 0; sidx < sg.max_sweeps; ++sidx) {
        CK(launch_sv_sweep(L, sidx, tiles * ch.count, st));
        S.n_other_launches++;
      }
      if (sg.n_groups > 0) {
        ZexpLaunch Z;
        Z.states = L.states; Z.stride = L.stride; Z.hi_bits = 0;
        Z.splits = zexp_splits(ctx, sg.n_groups, ch.nb);
        Z.group_desc = (const int32_t*)(db + W.o_gdesc) + 4 * (size_t)sg.group_first;
        Z.zt_mask = (const uint32_t*)(db + W.o_ztm);
        Z.partial = (double*)ctx->d_partial.p;
        sv_zexp_kernel<<<(unsigned)(sg.n_groups * Z.splits), kZexpThreads, 0, st>>>(Z);
        CK(cudaGetLastError());
        ZexpFinalize F;
        F.circ_desc = (const int32_t*)(db + W.o_cdesc) + 4 * (size_t)sg.cdesc_first;
        F.group_desc = Z.group_desc;
        F.zt_coeff = (const double*)(db + W.o_ztc);
        F.zt_obs = (const int32_t*)(db + W.o_zto);
        F.partial = Z.partial; F.splits = Z.splits; F.out = d_out;
        sv_zexp_finalize<<<(unsigned)sg.n_cdesc, 128, 0, st>>>(F);
        CK(cudaGetLastError());
        S.n_other_launches += 2;
      }
    }
  }
  return BWQ_OK;
}

// ------------------------------------------------------------------------------------------------
// statevector run (ideal labels)
// ------------------------------------------------------------------------------------------------
static int sv_prepare_impl(bwq_ctx* ctx, const bwq_batch* b, int32_t* out_status) {
  NvtxRange nvtx_range("bwq:sv_prepare (K0)");
  if (!ctx) return BWQ_ERR_ARG;
  int rc = check_batch(ctx, b, out_status, out_status);
  if (rc) return rc;
  SvPlan& P = ctx->sv_plan;
  P = SvPlan();
  ctx->stats = bwq_stats{};
  const int N = b->n_circuits;
  P.n_obs = N ? b->obs_offsets[N] : 0;
  P.valid = true;
  if (N == 0) return BWQ_OK;
  CK(cudaSetDevice(ctx->device));
  double t0 = now_ms();
  std::vector<SvProgram> progs(N);
  parallel_for(ctx, N, host_threads(ctx), [&](int c) { lower_sv_circuit(*b, c, &progs[c]); });
  // <= kSvSmallBits active qubits: one CTA per circuit, state in shared memory; wider: tile sweeps
  std::vector<int> order, wide;
  for (int c = 0; c < N; ++c) {
    if (progs[c].status == 0 && progs[c].n_bits > kSvSmallBits) { wide.push_back(c); continue; }
    out_status[c] = progs[c].status;
    if (progs[c].status == 0) order.push_back(c);
    else for (int64_t o = b->obs_offsets[c]; o < b->obs_offsets[c + 1]; ++o) P.nan_obs.push_back(o);
  }
  if (!wide.empty()) {
    int rc2 = sv_wide_prepare(ctx, b, wide, out_status);
    if (rc2) return rc2;
  }
  std::sort(order.begin(), order.end(), [&](int a, int c) {
    if (progs[a].n_bits != progs[c].n_bits) return progs[a].n_bits > progs[c].n_bits;
    return a < c;
  });
  const int M = (int)order.size();
  std::vector<int64_t> op_off(M + 1, 0), mt_off(M + 1, 0), tm_off(M + 1, 0), ob_off(M + 1, 0);
  for (int i = 0; i < M; ++i) {
    const SvProgram& p = progs[order[i]];
    op_off[i + 1] = op_off[i] + (int64_t)p.ops.size();
    mt_off[i + 1] = mt_off[i] + (int64_t)p.mats.size();
    tm_off[i + 1] = tm_off[i] + (int64_t)p.term_coeff.size();
    ob_off[i + 1] = ob_off[i] + (b->obs_offsets[order[i] + 1] - b->obs_offsets[order[i]]);
    P.n_gates += p.n_gates;
  }
  if (op_off[M] > INT32_MAX || ob_off[M] > INT32_MAX) return fail(ctx, BWQ_ERR_ARG, "batch too large; split it");
  Blob blob;
  P.o_cd = blob.add(sizeof(int32_t) * 8 * (size_t)M);
  P.o_ops = blob.add(sizeof(SvOp) * (size_t)op_off[M]);
  P.o_mats = blob.add(sizeof(double) * (size_t)mt_off[M]);
  P.o_obs = blob.add(sizeof(int64_t) * 4 * (size_t)ob_off[M]);
  P.o_tx = blob.add(sizeof(uint32_t) * (size_t)tm_off[M]);
  P.o_tz = blob.add(sizeof(uint32_t) * (size_t)tm_off[M]);
  P.o_tny = blob.add(sizeof(int32_t) * (size_t)tm_off[M]);
  P.o_tc = blob.add(sizeof(double) * (size_t)tm_off[M]);
  P.blob_bytes = blob.total;
  if (M > 0) {
    CK(ctx->h_sv_prog.reserve(blob.total));
    CK(ctx->d_sv_prog.reserve(blob.total));
  }
  char* hb = (char*)ctx->h_sv_prog.p;
  parallel_for(ctx, M, host_threads(ctx), [&](int i) {
    const int c = order[i];
    const SvProgram& p = progs[c];
    int32_t* cd = (int32_t*)(hb + P.o_cd) + 8 * i;
    cd[0] = p.n_bits; cd[1] = (int32_t)op_off[i]; cd[2] = (int32_t)op_off[i + 1];
    cd[3] = (int32_t)ob_off[i]; cd[4] = (int32_t)ob_off[i + 1]; cd[5] = cd[6] = cd[7] = 0;
    SvOp* ops = (SvOp*)(hb + P.o_ops) + op_off[i];
    for (size_t k = 0; k < p.ops.size(); ++k) { ops[k] = p.ops[k]; ops[k].off += mt_off[i]; }
    if (!p.mats.empty()) std::memcpy((double*)(hb + P.o_mats) + mt_off[i], p.mats.data(), p.mats.size() * sizeof(double));
    const size_t nt = p.term_coeff.size();
    if (nt) {
      std::memcpy((uint32_t*)(hb + P.o_tx) + tm_off[i], p.term_x.data(), nt * sizeof(uint32_t));
      std::memcpy((uint32_t*)(hb + P.o_tz) + tm_off[i], p.term_z.data(), nt * sizeof(uint32_t));
      std::memcpy((int32_t*)(hb + P.o_tny) + tm_off[i], p.term_ny.data(), nt * sizeof(int32_t));
      std::memcpy((double*)(hb + P.o_tc) + tm_off[i], p.term_coeff.data(), nt * sizeof(double));
    }
    int64_t* od = (int64_t*)(hb + P.o_obs) + 4 * ob_off[i];
    const int64_t ob0 = b->obs_offsets[c], ob1 = b->obs_offsets[c + 1];
    const int64_t tb = b->term_offsets[ob0];
    for (int64_t o = ob0; o < ob1; ++o) {
      int64_t* d = od + 4 * (o - ob0);
      d[0] = tm_off[i] + (b->term_offsets[o] - tb);
      d[1] = tm_off[i] + (b->term_offsets[o + 1] - tb);
      d[2] = o;
      d[3] = 0;
    }
  });
  // launch groups: circuits of equal width
  constexpr int kSmemBits = kSvSmallBits;
  int64_t scratch = 0;
  for (int i = 0; i < M;) {
    const int nb = progs[order[i]].n_bits;
    int j = i;
    while (j < M && progs[order[j]].n_bits == nb) ++j;
    SvGroup g;
    g.first = i; g.count = j - i; g.nb = nb; g.per_launch = j - i;
    if (nb <= kSmemBits) g.smem = sizeof(double2) << nb;
    else {
      g.stride = int64_t(1) << nb;
      g.per_launch = std::min(g.per_launch, std::max(1, std::min(4 * ctx->sm_count, (int)((int64_t(8) << 30) / (g.stride * 16)))));
      scratch = std::max(scratch, (int64_t)sizeof(double2) * g.stride * g.per_launch);
    }
    P.groups.push_back(g);
    i = j;
  }
  P.lower_ms = now_ms() - t0;
  if (scratch > 0) CK(ctx->d_scratch.reserve((size_t)scratch));
  if (P.n_obs > 0) {
    CK(ctx->d_out.reserve(sizeof(double) * (size_t)P.n_obs));
    CK(ctx->h_out.reserve(sizeof(double) * (size_t)P.n_obs));
  }
  cudaStream_t st = ctx->stream;
  CK(cudaEventRecord(ctx->ev[0], st));
  if (M > 0) CK(cudaMemcpyAsync(ctx->d_sv_prog.p, hb, blob.total, cudaMemcpyHostToDevice, st));
  if (P.wide.any) CK(cudaMemcpyAsync(ctx->d_wide_prog.p, ctx->h_wide_prog.p, P.wide.blob_bytes, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(ctx->ev[1], st));
  CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  P.h2d_ms = ms;
  P.blob_bytes += P.wide.blob_bytes;
  ctx->stats.lower_ms = P.lower_ms; ctx->stats.h2d_ms = ms; ctx->stats.h2d_bytes = (int64_t)P.blob_bytes;
  ctx->stats.n_gates = P.n_gates;
  return BWQ_OK;
}

static int sv_execute_impl(bwq_ctx* ctx, double* out_vals) {
  NvtxRange nvtx_range("bwq:sv_execute");
  if (!ctx || !out_vals) return BWQ_ERR_ARG;
  SvPlan& P = ctx->sv_plan;
  if (!P.valid) return fail(ctx, BWQ_ERR_ARG, "bwq_sv_execute: no prepared batch (call bwq_sv_prepare first)");
  CK(cudaSetDevice(ctx->device));
  bwq_stats& S = ctx->stats;
  S = bwq_stats{};
  S.lower_ms = P.lower_ms; S.h2d_ms = P.h2d_ms; S.h2d_bytes = (int64_t)P.blob_bytes; S.n_gates = P.n_gates;
  double* d_out = (double*)ctx->d_out.p;
  cudaStream_t st = ctx->stream;
  const char* db = (const char*)ctx->d_sv_prog.p;
  CK(cudaEventRecord(ctx->ev[1], st));
  if (P.n_obs > 0) CK(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)P.n_obs, st));
  for (const SvGroup& g : P.groups) {
    for (int f = g.first; f < g.first + g.count; f += g.per_launch) {
      SvLaunch L;
      L.first_circuit = f;
      L.n_circuits = std::min(g.per_launch, g.first + g.count - f);
      L.circ_desc = (const int32_t*)(db + P.o_cd);
      L.ops = (const SvOp*)(db + P.o_ops);
      L.mats = (const double*)(db + P.o_mats);
      L.obs_desc = (const int64_t*)(db + P.o_obs);
      L.term_x = (const uint32_t*)(db + P.o_tx);
      L.term_z = (const uint32_t*)(db + P.o_tz);
      L.term_ny = (const int32_t*)(db + P.o_tny);
      L.term_coeff = (const double*)(db + P.o_tc);
      L.out = d_out;
      L.scratch = (double2*)ctx->d_scratch.p;
      L.scratch_stride = g.stride;
      L.smem_bits = 12;
      sv_circuit_kernel<<<L.n_circuits, kSvThreads, g.smem, st>>>(L);
      CK(cudaGetLastError());
      S.n_other_launches++;
    }
  }
  if (P.wide.any) {
    int rc2 = sv_wide_execute(ctx, d_out);
    if (rc2) return rc2;
  }
  CK(cudaEventRecord(ctx->ev[2], st));
  if (P.n_obs > 0) {
    CK(cudaMemcpyAsync(ctx->h_out.p, d_out, sizeof(double) * (size_t)P.n_obs, cudaMemcpyDeviceToHost, st));
    S.d2h_bytes = (int64_t)sizeof(double) * P.n_obs;
  }
  CK(cudaEventRecord(ctx->ev[3], st));
  CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2])); S.kernel_ms = ms;
  CK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3])); S.d2h_ms = ms;
  if (P.n_obs > 0) std::memcpy(out_vals, ctx->h_out.p, sizeof(double) * (size_t)P.n_obs);
  for (int64_t o : P.nan_obs) out_vals[o] = std::nan("");
  return BWQ_OK;
}

extern "C" int bwq_sv_prepare(bwq_ctx* ctx, const bwq_batch* b, int32_t* out_status) {
  if (!out_status) return fail(ctx, BWQ_ERR_ARG, "null status");
  return sv_prepare_impl(ctx, b, out_status);
}
extern "C" int bwq_sv_execute(bwq_ctx* ctx, double* out_vals) { return sv_execute_impl(ctx, out_vals); }
extern "C" int bwq_sv_run(bwq_ctx* ctx, const bwq_batch* b, double* out_vals, int32_t* out_status) {
  if (!ctx) return BWQ_ERR_ARG;
  if (!out_vals || !out_status) return fail(ctx, BWQ_ERR_ARG, "null out/status");
  int rc = check_batch(ctx, b, out_vals, out_status);
  if (rc) return rc;
  {  // circuits that fit on chip: the noise-free Pauli-basis evolution gives the same <P> (onchip.cuh)
    bool handled = false;
    if ((rc = onchip_run(ctx, b, nullptr, false, out_vals, false, out_status, &handled))) return rc;
    if (handled) return BWQ_OK;
  }
  rc = sv_prepare_impl(ctx, b, out_status);
  return rc ? rc : sv_execute_impl(ctx, out_vals);
}

// statevector side of the (ideal, noisy) calls: a companion context with its own stream and buffers
static int ensure_companion(bwq_ctx* ctx) {
  if (!ctx->companion) {
    int rc = bwq_create(ctx->device, &ctx->companion);
    if (rc) return fail(ctx, rc, "companion context: %s", bwq_last_error(nullptr));
    // the statevector side is a chain of short dependent launches: on a high-priority stream its
    // CTAs are placed ahead of the queued density-matrix tiles instead of waiting for a whole sweep
    int least = 0, greatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    cudaStream_t hp = nullptr;
    CK(cudaStreamCreateWithPriority(&hp, cudaStreamNonBlocking, greatest));
    cudaStreamDestroy(ctx->companion->stream);
    ctx->companion->stream = hp;
  }
  return BWQ_OK;
}

extern "C" int bwq_meas_data_run(bwq_ctx* ctx, const bwq_batch* b, double* out_ideal, double* out_noisy,
                                 int32_t* status_ideal, int32_t* status_noisy) {
  if (!ctx) return BWQ_ERR_ARG;
  if (!out_ideal || !out_noisy || !status_ideal || !status_noisy) return fail(ctx, BWQ_ERR_ARG, "null out/status");
  {
    int rc = check_batch(ctx, b, out_ideal, status_ideal);
    if (rc) return rc;
    bool handled = false;
    // small circuits: both sides in one launch of dm_onchip_kernel (ideal = without the noise table)
    if ((rc = onchip_run(ctx, b, nullptr, true, out_noisy, false, status_noisy, &handled, out_ideal, status_ideal))) return rc;
    if (handled) return BWQ_OK;
  }
  int rc_c = ensure_companion(ctx);
  if (rc_c) return rc_c;
  // both lowering stages run at once: split the host threads (statevector lowering is the cheaper
  // one) instead of oversubscribing the cores, which shows up as multi-millisecond join tails
  const int all_threads = host_threads(ctx), saved_threads = ctx->opt.host_threads;
  ctx->companion->opt = ctx->opt;
  ctx->companion->opt.host_threads = std::max(1, all_threads / 4);
  ctx->opt.host_threads = std::max(1, all_threads - ctx->companion->opt.host_threads);
  int rc_sv = BWQ_OK;
  Trace tr("meas_data_run");
  std::thread ideal([&] { rc_sv = bwq_sv_run(ctx->companion, b, out_ideal, status_ideal); });
  const int rc_dm = dm_run_impl(ctx, b, out_noisy, status_noisy, false, nullptr, false);
  tr.mark("density-matrix side returned");
  ideal.join();
  tr.mark("statevector side joined");
  ctx->opt.host_threads = saved_threads;
  if (rc_dm) return rc_dm;
  if (rc_sv) return fail(ctx, rc_sv, "statevector side: %s", bwq_last_error(ctx->companion));
  return BWQ_OK;
}

// variants: expand on the host (K0), noisy side on every variant, ideal side on the base circuits
extern "C" int bwq_meas_data_run_variants(bwq_ctx* ctx, const bwq_batch* b, const bwq_variants* v, double* out_ideal, double* out_noisy,
                                          int32_t* status_ideal, int32_t* status_noisy) {
  if (!ctx) return BWQ_ERR_ARG;
  if (!b || !v || !out_ideal || !out_noisy || !status_ideal || !status_noisy) return fail(ctx, BWQ_ERR_ARG, "null argument");
  int rc = check_batch(ctx, b, out_ideal, status_ideal);
  if (rc) return rc;
  {
    bool handled = false;
    const int64_t n_var = (int64_t)std::max(1, v->n_folds) * std::max(1, v->n_twirls);
    std::vector<int32_t> st_all((size_t)std::max<int64_t>(1, b->n_circuits * n_var));
    if ((rc = onchip_run(ctx, b, v, true, out_noisy, false, st_all.data(), &handled, out_ideal, status_ideal))) return rc;
    if (handled) {
      for (int c = 0; c < b->n_circuits; ++c) {
        status_noisy[c] = 0;
        for (int64_t k = 0; k < n_var && !status_noisy[c]; ++k) status_noisy[c] = st_all[(size_t)(c * n_var + k)];
      }
      return BWQ_OK;
    }
  }
  if ((rc = ensure_companion(ctx))) return rc;
  const double t0 = now_ms();
  ExpandedBatch X;
  if ((rc = expand_variants(*b, *v, &X, host_threads(ctx)))) return fail(ctx, rc, "bad variants descriptor (folds must be odd positive factors)");
  const double expand_ms = now_ms() - t0;
  std::vector<int32_t> st_var((size_t)X.view.n_circuits);
  const int all_threads = host_threads(ctx), saved_threads = ctx->opt.host_threads;
  ctx->companion->opt = ctx->opt;
  // host threads in proportion to the lowering work of each side: the ideal side lowers the base
  // circuits only; the noisy side every variant (about a third of that per extra fold of a base circuit)
  {
    const double sv_work = 16.0 * b->n_circuits;
    const double dm_work = (v->n_twirls == 0 ? 19.0 + 6.0 * (X.n_variants - 1) : 19.0 * X.n_variants) * b->n_circuits;
    ctx->companion->opt.host_threads = std::max(1, (int)std::lround(all_threads * sv_work / (sv_work + dm_work)));
  }
  ctx->opt.host_threads = std::max(1, all_threads - ctx->companion->opt.host_threads);
  int rc_sv = BWQ_OK;
  std::thread ideal([&] { rc_sv = bwq_sv_run(ctx->companion, b, out_ideal, status_ideal); });
  const FoldInfo fi{b, v->folds, v->n_folds};
  const int rc_dm = dm_run_impl(ctx, &X.view, out_noisy, st_var.data(), false, v->n_twirls == 0 && v->n_folds > 0 ? &fi : nullptr, false);
  ideal.join();
  ctx->opt.host_threads = saved_threads;
  if (rc_dm) return rc_dm;
  if (rc_sv) return fail(ctx, rc_sv, "statevector side: %s", bwq_last_error(ctx->companion));
  for (int c = 0; c < b->n_circuits; ++c) {
    status_noisy[c] = X.status[c];
    for (int k = 0; k < X.n_variants && !status_noisy[c]; ++k) status_noisy[c] = st_var[(size_t)c * X.n_variants + k];
    if (X.status[c])  // a gate without an inverse rule: the folded variants are not what was asked for
      for (int64_t o = X.view.obs_offsets[(size_t)c * X.n_variants]; o < X.view.obs_offsets[(size_t)(c + 1) * X.n_variants]; ++o) out_noisy[o] = std::nan("");
  }
  ctx->stats.lower_ms += expand_ms;
  return BWQ_OK;
}

extern "C" int bwq_dm_run_variants(bwq_ctx* ctx, const bwq_batch* b, const bwq_variants* v, double* out_vals, int32_t* out_status) {
  if (!ctx) return BWQ_ERR_ARG;
  if (!b || !v || !out_vals || !out_status) return fail(ctx, BWQ_ERR_ARG, "null argument");
  int rc = check_batch(ctx, b, out_vals, out_status);
  if (rc) return rc;
  {
    bool handled = false;
    if ((rc = onchip_run(ctx, b, v, true, out_vals, false, out_status, &handled))) return rc;
    if (handled) return BWQ_OK;
  }
  const double t0 = now_ms();
  ExpandedBatch X;
  if ((rc = expand_variants(*b, *v, &X, host_threads(ctx)))) return fail(ctx, rc, "bad variants descriptor (folds must be odd positive factors)");
  const double expand_ms = now_ms() - t0;
  const FoldInfo fi{b, v->folds, v->n_folds};
  if ((rc = dm_run_impl(ctx, &X.view, out_vals, out_status, false, v->n_twirls == 0 && v->n_folds > 0 ? &fi : nullptr, false))) return rc;
  for (int c = 0; c < b->n_circuits; ++c)
    if (X.status[c])
      for (int k = 0; k < X.n_variants; ++k) {
        const size_t vi = (size_t)c * X.n_variants + k;
        out_status[vi] = X.status[c];
        for (int64_t o = X.view.obs_offsets[vi]; o < X.view.obs_offsets[vi + 1]; ++o) out_vals[o] = std::nan("");
      }
  ctx->stats.lower_ms += expand_ms;
  return BWQ_OK;
}

extern "C" int bwq_expand_variants(const bwq_batch* b, const bwq_variants* v, int64_t sizes[4], int64_t* op_offsets, bwq_op* ops,
                                   double* params) {
  if (!b || !v || !sizes) return BWQ_ERR_ARG;
  ExpandedBatch X;
  int rc = expand_variants(*b, *v, &X, 4);
  if (rc) return rc;
  sizes[0] = X.view.n_circuits; sizes[1] = (int64_t)X.ops.size(); sizes[2] = (int64_t)X.params.size(); sizes[3] = X.n_variants;
  if (op_offsets) std::memcpy(op_offsets, X.op_offsets.data(), sizeof(int64_t) * X.op_offsets.size());
  if (ops && !X.ops.empty()) std::memcpy(ops, X.ops.data(), sizeof(bwq_op) * X.ops.size());
  if (params && !X.params.empty()) std::memcpy(params, X.params.data(), sizeof(double) * X.params.size());
  return BWQ_OK;
}

// ------------------------------------------------------------------------------------------------
// host-only lowering introspection
// ------------------------------------------------------------------------------------------------
extern "C" int bwq_lower_dm_ex(const bwq_noise_table* table, const bwq_batch* batch, int32_t circuit, int32_t tile_qubits,
                               int32_t low_qubits, int32_t flags, bwq_program** out);
extern "C" int bwq_lower_dm(const bwq_noise_table* table, const bwq_batch* batch, int32_t circuit,
                            int32_t tile_qubits, int32_t low_qubits, bwq_program** out) {
  return bwq_lower_dm_ex(table, batch, circuit, tile_qubits, low_qubits, 0, out);
}
extern "C" int bwq_lower_dm_ex(const bwq_noise_table* table, const bwq_batch* batch, int32_t circuit, int32_t tile_qubits,
                               int32_t low_qubits, int32_t flags, bwq_program** out) {
  if (!batch || !out || circuit < 0 || circuit >= batch->n_circuits) return BWQ_ERR_ARG;
  NoiseTable nt;
  char err[256];
  int rc = nt.set(table, err, sizeof err);
  if (rc) { g_create_error = err; return rc; }
  LowerOptions lo;
  lo.tile_qubits = tile_qubits ? tile_qubits : 6;
  lo.low_qubits = low_qubits < 0 ? 0 : (low_qubits ? low_qubits : 2);
  lo.tma = (flags & 1) != 0;
  lo.tma_direct_store = (flags & 2) != 0;
  bwq_program* p = new bwq_program();
  const int32_t fold = (flags >> 8) & 0xff;  // ZNE noise factor: the fold-aware path of bwq_*_variants
  if (fold > 1) {
    if (!lower_dm_circuit_folds(nt, *batch, circuit, lo, &fold, 1, &p->p)) p->p.status = BWQ_CIRC_BAD_OP;
  } else {
    lower_dm_circuit(nt, *batch, circuit, lo, &p->p);
  }
  *out = p;
  return BWQ_OK;
}

extern "C" void bwq_program_free(bwq_program* p) { delete p; }

extern "C" int bwq_program_sizes(const bwq_program* p, int64_t s[8]) {
  if (!p || !s) return BWQ_ERR_ARG;
  s[0] = p->p.n_digits; s[1] = (int64_t)p->p.sweeps.size(); s[2] = p->p.n_passes;
  s[3] = (int64_t)p->p.prog.size(); s[4] = p->p.needs_dense ? 1 : 0; s[5] = p->p.status;
  s[6] = (int64_t)p->p.term_index.size(); s[7] = p->p.n_gates;
  return BWQ_OK;
}

extern "C" int bwq_program_read(const bwq_program* p, int32_t* active, int32_t* sweeps, uint64_t* prog,
                                int64_t* term_index, double* term_coeff) {
  if (!p) return BWQ_ERR_ARG;
  const CircuitProgram& q = p->p;
  if (active) for (size_t i = 0; i < q.active.size(); ++i) active[i] = q.active[i];
  if (sweeps)
    for (size_t i = 0; i < q.sweeps.size(); ++i) {
      int32_t* s = sweeps + 10 * i;
      s[0] = (int32_t)q.sweeps[i].blk_q16;
      for (int k = 0; k < 8; ++k) s[1 + k] = q.sweeps[i].pos[k];
      s[9] = (int32_t)(q.sweeps[i].blk_len_q16 & 0xffffu);
    }
  if (prog && !q.prog.empty()) std::memcpy(prog, q.prog.data(), q.prog.size() * sizeof(uint64_t));
  if (term_index && !q.term_index.empty()) std::memcpy(term_index, q.term_index.data(), q.term_index.size() * sizeof(int64_t));
  if (term_coeff && !q.term_coeff.empty()) std::memcpy(term_coeff, q.term_coeff.data(), q.term_coeff.size() * sizeof(double));
  return BWQ_OK;
}

// ------------------------------------------------------------------------------------------------
// wide / amplitude-sharded statevector of ONE circuit on caller-owned device memory
// ------------------------------------------------------------------------------------------------
struct bwq_svx_program {
  SvxProgram p;
  int64_t n_obs = 0;
  int device = -1;
  DevBuf d_blob, d_partial;
  size_t o_sweeps = 0, o_unt = 0, o_prog = 0, o_ztm = 0, o_ztc = 0, o_zto = 0, o_gdesc = 0, o_cdesc = 0;
  std::vector<int> seg_group_first, seg_n_groups;  // per segment (EXPVAL only)
  bool uploaded = false;
};

extern "C" int bwq_svx_lower(const bwq_batch* batch, int32_t circuit, int32_t tile_bits, int32_t n_global_bits,
                             bwq_svx_program** out) {
  if (!batch || !out || circuit < 0 || circuit >= batch->n_circuits || n_global_bits < 0 || n_global_bits > 8)
    return fail(nullptr, BWQ_ERR_ARG, "bwq_svx_lower: bad arguments");
  bwq_svx_program* h = new bwq_svx_program();
  SvxOptions so;
  so.tile_bits = tile_bits > 0 ? tile_bits : kSvTileBitsDefault;
  so.n_global = n_global_bits;
  lower_svx_circuit(*batch, circuit, so, &h->p);
  h->n_obs = batch->obs_offsets[circuit + 1] - batch->obs_offsets[circuit];
  *out = h;
  return BWQ_OK;
}

extern "C" void bwq_svx_free(bwq_svx_program* h) {
  if (!h) return;
  if (h->uploaded) { cudaSetDevice(h->device); h->d_blob.release(); h->d_partial.release(); }
  delete h;
}

extern "C" int bwq_svx_sizes(const bwq_svx_program* h, int64_t s[12]) {
  if (!h || !s) return BWQ_ERR_ARG;
  const SvxProgram& p = h->p;
  s[0] = p.status; s[1] = p.n_bits; s[2] = p.n_local; s[3] = p.n_global; s[4] = p.tile_bits;
  s[5] = (int64_t)p.sweeps.size(); s[6] = (int64_t)p.prog.size(); s[7] = (int64_t)p.segs.size();
  s[8] = (int64_t)p.zt_mask.size(); s[9] = p.n_passes; s[10] = p.n_exchanges; s[11] = h->n_obs;
  return BWQ_OK;
}

extern "C" int bwq_svx_read(const bwq_svx_program* h, int32_t* active, int32_t* sweeps, uint64_t* prog, int32_t* segs,
                            uint32_t* zt_mask, double* zt_coeff, int32_t* zt_obs) {
  if (!h) return BWQ_ERR_ARG;
  const SvxProgram& q = h->p;
  if (active) for (size_t i = 0; i < q.active.size(); ++i) active[i] = q.active[i];
  if (sweeps)
    for (size_t i = 0; i < q.sweeps.size(); ++i) {
      int32_t* s = sweeps + 10 * i;
      s[0] = (int32_t)q.sweeps[i].blk_q16;
      for (int k = 0; k < 8; ++k) s[1 + k] = q.sweeps[i].pos[k];
      s[9] = (int32_t)q.sweeps[i].blk_len_q16;
    }
  if (prog && !q.prog.empty()) std::memcpy(prog, q.prog.data(), q.prog.size() * sizeof(uint64_t));
  if (segs)
    for (size_t i = 0; i < q.segs.size(); ++i) {
      segs[4 * i] = q.segs[i].kind; segs[4 * i + 1] = q.segs[i].first; segs[4 * i + 2] = q.segs[i].count; segs[4 * i + 3] = 0;
    }
  if (zt_mask && !q.zt_mask.empty()) std::memcpy(zt_mask, q.zt_mask.data(), q.zt_mask.size() * sizeof(uint32_t));
  if (zt_coeff && !q.zt_coeff.empty()) std::memcpy(zt_coeff, q.zt_coeff.data(), q.zt_coeff.size() * sizeof(double));
  if (zt_obs && !q.zt_obs.empty()) std::memcpy(zt_obs, q.zt_obs.data(), q.zt_obs.size() * sizeof(int32_t));
  return BWQ_OK;
}

extern "C" int64_t bwq_svx_bytes(const bwq_svx_program* h, int32_t rank) {
  if (!h) return -1;
  const SvxProgram& p = h->p;
  const int K = p.tile_bits, LBk = std::max(0, K - kSvFreeSlots);
  const int64_t full = (int64_t)sizeof(double2) << p.n_local;
  const uint32_t hi = uint32_t(rank) << p.n_local;
  int64_t bytes = 0;
  for (size_t k = 0; k < p.sweeps.size(); ++k) {
    if (k > 0 && (hi & p.sweep_untouched[k])) continue;  // the whole shard is still zero
    uint32_t outside = (1u << p.n_local) - 1u;
    outside &= ~((1u << LBk) - 1u);
    for (int s2 = 0; s2 < K - LBk; ++s2) outside &= ~(1u << p.sweeps[k].pos[s2]);
    const int u = __builtin_popcount(outside & p.sweep_untouched[k]);
    bytes += k == 0 ? full : 2 * (full >> u);
  }
  for (const SvxSegment& sg : p.segs)
    if (sg.kind == SVSEG_EXPVAL) bytes += full * ((sg.count + kZexpTerms - 1) / kZexpTerms);
  return bytes;
}

extern "C" int bwq_svx_upload(bwq_ctx* ctx, bwq_svx_program* h) {
  if (!ctx || !h) return BWQ_ERR_ARG;
  const SvxProgram& p = h->p;
  if (p.status != 0) return fail(ctx, BWQ_ERR_ARG, "bwq_svx_upload: program status %d", p.status);
  CK(cudaSetDevice(ctx->device));
  std::vector<int32_t> gdesc, cdesc;
  h->seg_group_first.assign(p.segs.size(), 0);
  h->seg_n_groups.assign(p.segs.size(), 0);
  int max_groups = 0;
  for (size_t i = 0; i < p.segs.size(); ++i) {
    if (p.segs[i].kind != SVSEG_EXPVAL) continue;
    h->seg_group_first[i] = (int)(gdesc.size() / 4);
    int ng = 0;
    for (int t0 = 0; t0 < p.segs[i].count; t0 += kZexpTerms, ++ng) {
      gdesc.push_back(0); gdesc.push_back(p.segs[i].first + t0);
      gdesc.push_back(std::min(kZexpTerms, p.segs[i].count - t0)); gdesc.push_back(0);
    }
    h->seg_n_groups[i] = ng;
    max_groups = std::max(max_groups, ng);
    cdesc.push_back(0); cdesc.push_back(ng); cdesc.push_back(0); cdesc.push_back(0);  // one per EXPVAL segment
  }
  Blob blob;
  h->o_sweeps = blob.add(sizeof(SweepDesc) * p.sweeps.size());
  h->o_unt = blob.add(sizeof(uint32_t) * p.sweeps.size());
  h->o_prog = blob.add(sizeof(uint64_t) * p.prog.size());
  h->o_ztm = blob.add(sizeof(uint32_t) * p.zt_mask.size());
  h->o_ztc = blob.add(sizeof(double) * p.zt_coeff.size());
  h->o_zto = blob.add(sizeof(int32_t) * p.zt_obs.size());
  h->o_gdesc = blob.add(sizeof(int32_t) * gdesc.size());
  h->o_cdesc = blob.add(sizeof(int32_t) * 4);
  std::vector<char> hb(blob.total + 256, 0);
  auto put = [&](size_t off, const void* src, size_t n) { if (n) std::memcpy(hb.data() + off, src, n); };
  put(h->o_sweeps, p.sweeps.data(), sizeof(SweepDesc) * p.sweeps.size());
  put(h->o_unt, p.sweep_untouched.data(), sizeof(uint32_t) * p.sweeps.size());
  put(h->o_prog, p.prog.data(), sizeof(uint64_t) * p.prog.size());
  put(h->o_ztm, p.zt_mask.data(), sizeof(uint32_t) * p.zt_mask.size());
  put(h->o_ztc, p.zt_coeff.data(), sizeof(double) * p.zt_coeff.size());
  put(h->o_zto, p.zt_obs.data(), sizeof(int32_t) * p.zt_obs.size());
  put(h->o_gdesc, gdesc.data(), sizeof(int32_t) * gdesc.size());
  h->device = ctx->device;
  CK(h->d_blob.reserve(hb.size()));
  CK(cudaMemcpy(h->d_blob.p, hb.data(), hb.size(), cudaMemcpyHostToDevice));
  if (max_groups > 0)
    CK(h->d_partial.reserve(sizeof(double) * (size_t)max_groups * zexp_splits(ctx, 1, p.n_local) * kZexpTerms));
  h->uploaded = true;
  return BWQ_OK;
}

static int svx_exchange_impl(bwq_ctx* ctx, double* d_local, const uint64_t* peers, int32_t world, int32_t rank,
                             int64_t n_local_amps, void* stream, bool push) {
  if (!ctx || !d_local || !peers) return BWQ_ERR_ARG;
  if (world < 2 || world > kSvxMaxWorld || (world & (world - 1)) || rank < 0 || rank >= world)
    return fail(ctx, BWQ_ERR_ARG, "bwq_svx_exchange: world must be a power of two in [2,%d], rank inside it", kSvxMaxWorld);
  if (n_local_amps < world || n_local_amps % world) return fail(ctx, BWQ_ERR_ARG, "shard size must be a multiple of world");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  SvxPeers P{};
  for (int w = 0; w < world; ++w) {
    if (!peers[w]) return fail(ctx, BWQ_ERR_ARG, "peer pointer %d is NULL", w);
    P.ptr[w] = reinterpret_cast<double2*>(peers[w]);
  }
  const int64_t blk = n_local_amps / world;
  // CTAs per partner rank: enough 16-byte accesses in flight for the NVLink round trip, about two waves
  const int64_t per = std::max<int64_t>(1, std::min<int64_t>((blk + 256 * 8 - 1) / (256 * 8), 2 * (int64_t)ctx->sm_count * 8 / world));
  if (push) svx_exchange_kernel<true><<<(unsigned)(per * world), 256, 0, st>>>((double2*)d_local, P, world, rank, blk);
  else svx_exchange_kernel<false><<<(unsigned)(per * world), 256, 0, st>>>((double2*)d_local, P, world, rank, blk);
  CK(cudaGetLastError());
  return BWQ_OK;
}
extern "C" int bwq_svx_exchange_pull(bwq_ctx* ctx, double* d_dst, const uint64_t* peer_src, int32_t world, int32_t rank,
                                     int64_t n_local_amps, void* stream) {
  return svx_exchange_impl(ctx, d_dst, peer_src, world, rank, n_local_amps, stream, false);
}
extern "C" int bwq_svx_exchange_push(bwq_ctx* ctx, const double* d_src, const uint64_t* peer_dst, int32_t world, int32_t rank,
                                     int64_t n_local_amps, void* stream) {
  return svx_exchange_impl(ctx, const_cast<double*>(d_src), peer_dst, world, rank, n_local_amps, stream, true);
}

static int svx_run_segment_impl(bwq_ctx* ctx, const bwq_svx_program* h, int32_t segment, double* d_state, int32_t rank, double* d_obs,
                                void* stream, const uint64_t* peer_dst, int32_t world);

extern "C" int bwq_svx_run_segment(bwq_ctx* ctx, const bwq_svx_program* h, int32_t segment, double* d_state,
                                   int32_t rank, double* d_obs, void* stream) {
  return svx_run_segment_impl(ctx, h, segment, d_state, rank, d_obs, stream, nullptr, 0);
}

extern "C" int bwq_svx_run_segment_push(bwq_ctx* ctx, const bwq_svx_program* h, int32_t segment, double* d_state, int32_t rank,
                                        const uint64_t* peer_dst, int32_t world, void* stream) {
  if (!ctx || !h || !peer_dst) return BWQ_ERR_ARG;
  const SvxProgram& p = h->p;
  if (segment < 0 || segment + 1 >= (int)p.segs.size() || p.segs[segment].kind != SVSEG_SWEEPS || p.segs[segment + 1].kind != SVSEG_EXCHANGE)
    return fail(ctx, BWQ_ERR_ARG, "bwq_svx_run_segment_push: segment must be a SWEEPS segment followed by an EXCHANGE");
  if (world != (1 << p.n_global) || world > kSvxMaxWorld) return fail(ctx, BWQ_ERR_ARG, "bwq_svx_run_segment_push: world must be 2^n_global (<= %d)", kSvxMaxWorld);
  return svx_run_segment_impl(ctx, h, segment, d_state, rank, nullptr, stream, peer_dst, world);
}

static int svx_run_segment_impl(bwq_ctx* ctx, const bwq_svx_program* h, int32_t segment, double* d_state, int32_t rank, double* d_obs,
                                void* stream, const uint64_t* peer_dst, int32_t world) {
  if (!ctx || !h || !d_state) return BWQ_ERR_ARG;
  if (!h->uploaded) return fail(ctx, BWQ_ERR_ARG, "bwq_svx_run_segment: call bwq_svx_upload first");
  const SvxProgram& p = h->p;
  if (segment < 0 || segment >= (int)p.segs.size()) return fail(ctx, BWQ_ERR_ARG, "segment out of range");
  if (rank < 0 || rank >= (1 << p.n_global)) return fail(ctx, BWQ_ERR_ARG, "rank out of range");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  const SvxSegment& sg = p.segs[segment];
  const char* db = (const char*)h->d_blob.p;
  const uint32_t hi = uint32_t(rank) << p.n_local;
  const int64_t stride = int64_t(1) << p.n_local;
  // a program that starts with an expectation value (no gates) still needs |0...0>
  if (segment == 0 && sg.kind != SVSEG_SWEEPS) {
    sv_init_kernel<<<dim3((unsigned)((stride + 255) / 256), 1), 256, 0, st>>>((double2*)d_state, stride, nullptr, 1, hi);
    CK(cudaGetLastError());
  }
  if (sg.kind == SVSEG_EXCHANGE)
    return fail(ctx, BWQ_ERR_UNSUPPORTED, "EXCHANGE segments are executed by the caller (all-to-all of the top local bits)");
  if (sg.kind == SVSEG_SWEEPS) {
    SvxLaunch L{};
    L.states = (double2*)d_state; L.stride = stride;
    L.n_local = p.n_local; L.tile_bits = p.tile_bits; L.low_bits = std::max(0, p.tile_bits - kSvFreeSlots);
    L.first_circuit = 0; L.sweep_range = nullptr;
    L.sweeps = (const SweepDesc*)(db + h->o_sweeps);
    L.sweep_untouched = (const uint32_t*)(db + h->o_unt);
    L.prog = (const uint4*)(db + h->o_prog);
    L.hi_bits = hi; L.init = 1;
    const int64_t tiles = int64_t(1) << (p.n_local - p.tile_bits);
    for (int s = sg.first; s < sg.first + sg.count; ++s) {
      if (peer_dst && s + 1 == sg.first + sg.count) {  // last sweep: its store is the EXCHANGE (P2P stores)
        L.push_g = p.n_global; L.push_rank = rank;
        for (int w = 0; w < world; ++w) L.push_ptr[w] = (double2*)peer_dst[w];
      }
      CK(launch_sv_sweep(L, s, tiles, st));
    }
    return BWQ_OK;
  }
  if (!d_obs) return fail(ctx, BWQ_ERR_ARG, "EXPVAL segment needs d_obs");
  const int ng = h->seg_n_groups[segment];
  if (ng == 0) return BWQ_OK;
  ZexpLaunch Z;
  Z.states = (const double2*)d_state; Z.stride = stride; Z.hi_bits = hi;
  Z.splits = zexp_splits(ctx, ng, p.n_local);
  Z.group_desc = (const int32_t*)(db + h->o_gdesc) + 4 * (size_t)h->seg_group_first[segment];
  Z.zt_mask = (const uint32_t*)(db + h->o_ztm);
  Z.partial = (double*)h->d_partial.p;
  sv_zexp_kernel<<<(unsigned)(ng * Z.splits), kZexpThreads, 0, st>>>(Z);
  CK(cudaGetLastError());
  // circ_desc {0, ng, 0, 0}: passed through a small device constant written at upload time
  int32_t cd[4] = {0, ng, 0, 0};
  CK(cudaMemcpyAsync((void*)(db + h->o_cdesc), cd, sizeof cd, cudaMemcpyHostToDevice, st));
  ZexpFinalize F;
  F.circ_desc = (const int32_t*)(db + h->o_cdesc);
  F.group_desc = Z.group_desc;
  F.zt_coeff = (const double*)(db + h->o_ztc);
  F.zt_obs = (const int32_t*)(db + h->o_zto);
  F.partial = Z.partial; F.splits = Z.splits; F.out = d_obs;
  sv_zexp_finalize<<<1, 128, 0, st>>>(F);
  CK(cudaGetLastError());
  return BWQ_OK;
}
