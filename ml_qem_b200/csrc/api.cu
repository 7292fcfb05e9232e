// C ABI (include/bwq.h): context, batch scheduling, launches.  There is no CPU fallback: every
// *_run entry point needs a bwq_ctx, and bwq_create fails without a CUDA device.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"

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

struct bwq_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::string error;
  bwq_options opt{};
  NoiseTable noise;
  DevBuf d_noise, d_prog, d_states, d_out, d_scratch;
  PinBuf h_prog, h_out;
  bwq_stats stats{};
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
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  ctx->sm_count = prop.multiProcessorCount;
  if ((e = cudaFuncSetAttribute(dm_sweep_kernel<7, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 << 14)) != cudaSuccess)
    return bail(e, "cudaFuncSetAttribute(dm_sweep_kernel<7>) -- was the library built for this GPU (sm_100a)?");
  if ((e = cudaFuncSetAttribute(dm_sweep_kernel<7, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 << 14)) != cudaSuccess)
    return bail(e, "cudaFuncSetAttribute(dm_sweep_kernel<7, full>)");
  if ((e = cudaFuncSetAttribute(sv_circuit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 << 12)) != cudaSuccess)
    return bail(e, "cudaFuncSetAttribute(sv_circuit_kernel)");
  *out = ctx;
  return BWQ_OK;
}

extern "C" int bwq_destroy(bwq_ctx* ctx) {
  if (!ctx) return BWQ_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->d_noise.release(); ctx->d_prog.release(); ctx->d_states.release(); ctx->d_out.release();
  ctx->d_scratch.release(); ctx->h_prog.release(); ctx->h_out.release();
  for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
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

template <class F> static void parallel_for(int n, int threads, F f) {
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
  dm_sweep_kernel<KQ, FULL><<<(unsigned)n_cta, SweepCfg<KQ>::kThreads, sizeof(double) << (2 * KQ), s>>>(L, sweep);
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
// density-matrix run
// ------------------------------------------------------------------------------------------------
static int dm_run_impl(bwq_ctx* ctx, const bwq_batch* b, double* out_vals, bool out_on_device, int32_t* out_status) {
  if (!ctx) return BWQ_ERR_ARG;
  int rc = check_batch(ctx, b, out_vals, out_status);
  if (rc) return rc;
  ctx->stats = bwq_stats{};
  const int N = b->n_circuits;
  if (N == 0) return BWQ_OK;
  CK(cudaSetDevice(ctx->device));
  const int64_t n_obs = b->obs_offsets[N];

  // ---- K0: lowering (host threads)
  double t0 = now_ms();
  LowerOptions lo;
  lo.tile_qubits = ctx->opt.tile_qubits ? ctx->opt.tile_qubits : 6;
  lo.low_qubits = ctx->opt.low_qubits ? ctx->opt.low_qubits : 2;
  if (ctx->opt.low_qubits < 0) lo.low_qubits = 0;
  std::vector<CircuitProgram> progs(N);
  parallel_for(N, host_threads(ctx), [&](int c) { lower_dm_circuit(ctx->noise, *b, c, lo, &progs[c]); });

  // host-evaluated circuits: failures, and circuits without any gate (state stays |0..0>)
  std::vector<double> host_vals;  // only used for those
  std::vector<int> order;
  order.reserve(N);
  for (int c = 0; c < N; ++c) {
    out_status[c] = progs[c].status;
    if (progs[c].status == 0 && !progs[c].sweeps.empty()) order.push_back(c);
  }
  // sort: wide first, then by sweep count (so a chunk's circuits finish together)
  std::sort(order.begin(), order.end(), [&](int a, int c) {
    if (progs[a].n_digits != progs[c].n_digits) return progs[a].n_digits > progs[c].n_digits;
    if (progs[a].sweeps.size() != progs[c].sweeps.size()) return progs[a].sweeps.size() > progs[c].sweeps.size();
    return a < c;
  });
  const int M = (int)order.size();

  // ---- merge programs into one blob (sorted order)
  std::vector<int64_t> sw_off(M + 1, 0), ps_off(M + 1, 0), op_off(M + 1, 0), mt_off(M + 1, 0), tm_off(M + 1, 0), ob_off(M + 1, 0);
  for (int i = 0; i < M; ++i) {
    const CircuitProgram& p = progs[order[i]];
    sw_off[i + 1] = sw_off[i] + (int64_t)p.sweeps.size();
    ps_off[i + 1] = ps_off[i] + (int64_t)p.passes.size();
    op_off[i + 1] = op_off[i] + (int64_t)p.ops.size();
    mt_off[i + 1] = mt_off[i] + (int64_t)p.mats.size();
    tm_off[i + 1] = tm_off[i] + (int64_t)p.term_index.size();
    ob_off[i + 1] = ob_off[i] + (b->obs_offsets[order[i] + 1] - b->obs_offsets[order[i]]);
  }
  if (ps_off[M] > INT32_MAX || op_off[M] > INT32_MAX || sw_off[M] > INT32_MAX)
    return fail(ctx, BWQ_ERR_ARG, "batch too large for 32-bit program indices; split the batch");
  Blob blob;
  const size_t o_range = blob.add(sizeof(int32_t) * 2 * (size_t)M);
  const size_t o_sweeps = blob.add(sizeof(SweepDesc) * (size_t)sw_off[M]);
  const size_t o_passes = blob.add(sizeof(PassDesc) * (size_t)ps_off[M]);
  const size_t o_ops = blob.add(sizeof(DevOp) * (size_t)op_off[M]);
  const size_t o_mats = blob.add(sizeof(double) * (size_t)mt_off[M]);
  const size_t o_tidx = blob.add(sizeof(int64_t) * (size_t)tm_off[M]);
  const size_t o_tcoef = blob.add(sizeof(double) * (size_t)tm_off[M]);
  const size_t o_obs = blob.add(sizeof(int64_t) * 4 * (size_t)ob_off[M]);
  if (M > 0) {
    CK(ctx->h_prog.reserve(blob.total));
    CK(ctx->d_prog.reserve(blob.total));
  }
  char* hb = (char*)ctx->h_prog.p;
  parallel_for(M, host_threads(ctx), [&](int i) {
    const int c = order[i];
    const CircuitProgram& p = progs[c];
    int32_t* range = (int32_t*)(hb + o_range) + 2 * i;
    range[0] = (int32_t)sw_off[i];
    range[1] = (int32_t)sw_off[i + 1];
    SweepDesc* sw = (SweepDesc*)(hb + o_sweeps) + sw_off[i];
    for (size_t k = 0; k < p.sweeps.size(); ++k) {
      sw[k] = p.sweeps[k];
      sw[k].pass_begin += (int32_t)ps_off[i];
      sw[k].pass_end += (int32_t)ps_off[i];
    }
    PassDesc* ps = (PassDesc*)(hb + o_passes) + ps_off[i];
    for (size_t k = 0; k < p.passes.size(); ++k) {
      ps[k] = p.passes[k];
      ps[k].op_begin += (int32_t)op_off[i];
      ps[k].op_end += (int32_t)op_off[i];
    }
    DevOp* ops = (DevOp*)(hb + o_ops) + op_off[i];
    for (size_t k = 0; k < p.ops.size(); ++k) {
      ops[k] = p.ops[k];
      if (ops[k].src == 0) ops[k].off += mt_off[i];
    }
    if (!p.mats.empty()) std::memcpy((double*)(hb + o_mats) + mt_off[i], p.mats.data(), p.mats.size() * sizeof(double));
    if (!p.term_index.empty()) {
      std::memcpy((int64_t*)(hb + o_tidx) + tm_off[i], p.term_index.data(), p.term_index.size() * sizeof(int64_t));
      std::memcpy((double*)(hb + o_tcoef) + tm_off[i], p.term_coeff.data(), p.term_coeff.size() * sizeof(double));
    }
    // observables: {term_begin, term_end, slot (filled per chunk = sorted index), out_index}
    int64_t* od = (int64_t*)(hb + o_obs) + 4 * ob_off[i];
    const int64_t ob0 = b->obs_offsets[c], ob1 = b->obs_offsets[c + 1];
    const int64_t tb = b->term_offsets[ob0];
    for (int64_t o = ob0; o < ob1; ++o) {
      int64_t* d = od + 4 * (o - ob0);
      d[0] = tm_off[i] + (b->term_offsets[o] - tb);
      d[1] = tm_off[i] + (b->term_offsets[o + 1] - tb);
      d[2] = i;  // rewritten to the chunk-local slot below
      d[3] = o;
    }
  });
  for (int i = 0; i < M; ++i) { ctx->stats.n_gates += progs[order[i]].n_gates; ctx->stats.n_passes += (int64_t)progs[order[i]].passes.size(); }
  ctx->stats.lower_ms = now_ms() - t0;

  // ---- chunk plan: circuits of equal width, as many resident states as the budget allows
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  int64_t budget = ctx->opt.max_state_bytes > 0 ? ctx->opt.max_state_bytes
                                                 : (int64_t)((free_b + ctx->d_states.cap) * 0.8);
  struct Chunk { int first, count, nd; };
  std::vector<Chunk> chunks;
  for (int i = 0; i < M;) {
    const int nd = progs[order[i]].n_digits;
    const int64_t sbytes = (int64_t)sizeof(double) << (2 * nd);
    int64_t fit = budget / sbytes;
    if (fit < 1) {
      int j = i;
      while (j < M && progs[order[j]].n_digits == nd) out_status[order[j++]] = BWQ_CIRC_TOO_WIDE;
      i = j;
      continue;
    }
    if (ctx->opt.chunk_circuits > 0) fit = std::min<int64_t>(fit, ctx->opt.chunk_circuits);
    const int kq = std::min(std::min(nd, lo.tile_qubits), kMaxTileQubits);
    const int64_t tiles = int64_t(1) << (2 * (nd - kq));
    fit = std::min<int64_t>(fit, (int64_t(1) << 30) / tiles);  // grid.x limit
    int j = i;
    while (j < M && progs[order[j]].n_digits == nd && j - i < fit) ++j;
    chunks.push_back({i, j - i, nd});
    i = j;
  }
  int64_t max_chunk_bytes = 0;
  for (auto& ch : chunks) max_chunk_bytes = std::max(max_chunk_bytes, ((int64_t)sizeof(double) << (2 * ch.nd)) * ch.count);
  // chunk-local slots for the observables
  for (auto& ch : chunks)
    for (int i = ch.first; i < ch.first + ch.count; ++i) {
      int64_t* od = (int64_t*)(hb + o_obs) + 4 * ob_off[i];
      for (int64_t k = 0; k < ob_off[i + 1] - ob_off[i]; ++k) od[4 * k + 2] = i - ch.first;
    }

  // ---- device: upload, sweeps, expectation values
  double* d_out = nullptr;
  if (out_on_device) d_out = out_vals;
  else if (n_obs > 0) {
    CK(ctx->d_out.reserve(sizeof(double) * (size_t)n_obs));
    CK(ctx->h_out.reserve(sizeof(double) * (size_t)n_obs));
    d_out = (double*)ctx->d_out.p;
  }
  if (max_chunk_bytes > 0) CK(ctx->d_states.reserve((size_t)max_chunk_bytes));
  cudaStream_t st = ctx->stream;
  CK(cudaEventRecord(ctx->ev[0], st));
  if (M > 0) CK(cudaMemcpyAsync(ctx->d_prog.p, hb, blob.total, cudaMemcpyHostToDevice, st));
  if (n_obs > 0) CK(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)n_obs, st));
  CK(cudaEventRecord(ctx->ev[1], st));
  const char* db = (const char*)ctx->d_prog.p;
  float sweep_ms_total = 0.f;
  for (auto& ch : chunks) {
    const int kq = std::min(std::min(ch.nd, lo.tile_qubits), kMaxTileQubits);
    DmLaunch L;
    L.states = (double*)ctx->d_states.p;
    L.stride = int64_t(1) << (2 * ch.nd);
    L.n_digits = ch.nd;
    L.first_circuit = ch.first;
    L.sweep_range = (const int32_t*)(db + o_range);
    L.sweeps = (const SweepDesc*)(db + o_sweeps);
    L.passes = (const PassDesc*)(db + o_passes);
    L.ops = (const DevOp*)(db + o_ops);
    L.mats = (const double*)(db + o_mats);
    L.noise = (const double*)ctx->d_noise.p;
    const int64_t tiles = int64_t(1) << (2 * (ch.nd - kq));
    // circuits are sorted by sweep count (descending) inside a width group, so sweep s only needs
    // the leading circuits that still have an s-th sweep
    size_t max_sweeps = progs[order[ch.first]].sweeps.size();
    bool full = false;
    for (int i = ch.first; i < ch.first + ch.count; ++i) full = full || progs[order[i]].needs_dense;
    int live = ch.count;
    for (size_t s = 0; s < max_sweeps; ++s) {
      while (live > 0 && progs[order[ch.first + live - 1]].sweeps.size() <= s) --live;
      CK(full ? launch_sweep_kq<true>(kq, L, (int)s, tiles * live, st) : launch_sweep_kq<false>(kq, L, (int)s, tiles * live, st));
      ctx->stats.n_sweep_launches++;
      ctx->stats.n_state_sweeps += live;
      ctx->stats.state_bytes_swept += 2 * (int64_t)sizeof(double) * L.stride * live;
    }
    const int64_t nob = ob_off[ch.first + ch.count] - ob_off[ch.first];
    if (nob > 0) {
      ExpvalLaunch E;
      E.states = L.states;
      E.stride = L.stride;
      E.n_obs = (int32_t)nob;
      E.obs_desc = (const int64_t*)(db + o_obs) + 4 * ob_off[ch.first];
      E.term_index = (const int64_t*)(db + o_tidx);
      E.term_coeff = (const double*)(db + o_tcoef);
      E.out = d_out;
      const int wpb = 8;
      dm_expval_kernel<<<(unsigned)((nob + wpb - 1) / wpb), wpb * 32, 0, st>>>(E);
      CK(cudaGetLastError());
      ctx->stats.n_other_launches++;
    }
  }
  CK(cudaEventRecord(ctx->ev[2], st));
  if (!out_on_device && n_obs > 0)
    CK(cudaMemcpyAsync(ctx->h_out.p, d_out, sizeof(double) * (size_t)n_obs, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(ctx->ev[3], st));
  CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); ctx->stats.h2d_ms = ms;
  CK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2])); ctx->stats.kernel_ms = ms;
  CK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3])); ctx->stats.d2h_ms = ms;
  ctx->stats.sweep_kernel_ms = sweep_ms_total > 0 ? sweep_ms_total : ctx->stats.kernel_ms;

  // ---- host-evaluated circuits: no gates at all => |0..0>: <P> = 1 for I/Z strings else 0
  std::vector<std::pair<int64_t, double>> host_fix;
  for (int c = 0; c < N; ++c) {
    const bool gpu = out_status[c] == 0 && !progs[c].sweeps.empty();
    if (gpu) continue;
    const int64_t ob0 = b->obs_offsets[c], ob1 = b->obs_offsets[c + 1];
    for (int64_t o = ob0; o < ob1; ++o) {
      double v = 0.0;
      if (out_status[c] == 0)
        for (int64_t t = b->term_offsets[o]; t < b->term_offsets[o + 1]; ++t)
          if (b->term_x[t] == 0) v += b->term_coeff[t];
      host_fix.push_back({o, out_status[c] == 0 ? v : std::nan("")});
    }
  }
  if (!out_on_device) {
    if (n_obs > 0) std::memcpy(out_vals, ctx->h_out.p, sizeof(double) * (size_t)n_obs);
    for (auto& f : host_fix) out_vals[f.first] = f.second;
  } else {
    for (auto& f : host_fix) CK(cudaMemcpy(out_vals + f.first, &f.second, sizeof(double), cudaMemcpyHostToDevice));
  }
  return BWQ_OK;
}

extern "C" int bwq_dm_run(bwq_ctx* ctx, const bwq_batch* b, double* out_vals, int32_t* out_status) {
  return dm_run_impl(ctx, b, out_vals, false, out_status);
}
extern "C" int bwq_dm_run_device_out(bwq_ctx* ctx, const bwq_batch* b, double* d_out_vals, int32_t* out_status) {
  return dm_run_impl(ctx, b, d_out_vals, true, out_status);
}

// ------------------------------------------------------------------------------------------------
// statevector run (ideal labels)
// ------------------------------------------------------------------------------------------------
extern "C" int bwq_sv_run(bwq_ctx* ctx, const bwq_batch* b, double* out_vals, int32_t* out_status) {
  if (!ctx) return BWQ_ERR_ARG;
  int rc = check_batch(ctx, b, out_vals, out_status);
  if (rc) return rc;
  ctx->stats = bwq_stats{};
  const int N = b->n_circuits;
  if (N == 0) return BWQ_OK;
  CK(cudaSetDevice(ctx->device));
  const int64_t n_obs = b->obs_offsets[N];
  double t0 = now_ms();
  std::vector<SvProgram> progs(N);
  parallel_for(N, host_threads(ctx), [&](int c) { lower_sv_circuit(*b, c, &progs[c]); });
  constexpr int kSimpleMaxBits = 24;  // one-CTA-per-circuit kernel; wider needs the sharded path
  std::vector<int> order;
  for (int c = 0; c < N; ++c) {
    if (progs[c].status == 0 && progs[c].n_bits > kSimpleMaxBits) progs[c].status = BWQ_CIRC_TOO_WIDE;
    out_status[c] = progs[c].status;
    if (progs[c].status == 0) order.push_back(c);
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
    ctx->stats.n_gates += p.n_gates;
  }
  if (op_off[M] > INT32_MAX || ob_off[M] > INT32_MAX) return fail(ctx, BWQ_ERR_ARG, "batch too large; split it");
  Blob blob;
  const size_t o_cd = blob.add(sizeof(int32_t) * 8 * (size_t)M);
  const size_t o_ops = blob.add(sizeof(SvOp) * (size_t)op_off[M]);
  const size_t o_mats = blob.add(sizeof(double) * (size_t)mt_off[M]);
  const size_t o_obs = blob.add(sizeof(int64_t) * 4 * (size_t)ob_off[M]);
  const size_t o_tx = blob.add(sizeof(uint32_t) * (size_t)tm_off[M]);
  const size_t o_tz = blob.add(sizeof(uint32_t) * (size_t)tm_off[M]);
  const size_t o_tny = blob.add(sizeof(int32_t) * (size_t)tm_off[M]);
  const size_t o_tc = blob.add(sizeof(double) * (size_t)tm_off[M]);
  if (M > 0) {
    CK(ctx->h_prog.reserve(blob.total));
    CK(ctx->d_prog.reserve(blob.total));
  }
  char* hb = (char*)ctx->h_prog.p;
  parallel_for(M, host_threads(ctx), [&](int i) {
    const int c = order[i];
    const SvProgram& p = progs[c];
    int32_t* cd = (int32_t*)(hb + o_cd) + 8 * i;
    cd[0] = p.n_bits; cd[1] = (int32_t)op_off[i]; cd[2] = (int32_t)op_off[i + 1];
    cd[3] = (int32_t)ob_off[i]; cd[4] = (int32_t)ob_off[i + 1]; cd[5] = cd[6] = cd[7] = 0;
    SvOp* ops = (SvOp*)(hb + o_ops) + op_off[i];
    for (size_t k = 0; k < p.ops.size(); ++k) { ops[k] = p.ops[k]; ops[k].off += mt_off[i]; }
    if (!p.mats.empty()) std::memcpy((double*)(hb + o_mats) + mt_off[i], p.mats.data(), p.mats.size() * sizeof(double));
    const size_t nt = p.term_coeff.size();
    if (nt) {
      std::memcpy((uint32_t*)(hb + o_tx) + tm_off[i], p.term_x.data(), nt * sizeof(uint32_t));
      std::memcpy((uint32_t*)(hb + o_tz) + tm_off[i], p.term_z.data(), nt * sizeof(uint32_t));
      std::memcpy((int32_t*)(hb + o_tny) + tm_off[i], p.term_ny.data(), nt * sizeof(int32_t));
      std::memcpy((double*)(hb + o_tc) + tm_off[i], p.term_coeff.data(), nt * sizeof(double));
    }
    int64_t* od = (int64_t*)(hb + o_obs) + 4 * ob_off[i];
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
  ctx->stats.lower_ms = now_ms() - t0;

  if (n_obs > 0) {
    CK(ctx->d_out.reserve(sizeof(double) * (size_t)n_obs));
    CK(ctx->h_out.reserve(sizeof(double) * (size_t)n_obs));
  }
  double* d_out = (double*)ctx->d_out.p;
  cudaStream_t st = ctx->stream;
  CK(cudaEventRecord(ctx->ev[0], st));
  if (M > 0) CK(cudaMemcpyAsync(ctx->d_prog.p, hb, blob.total, cudaMemcpyHostToDevice, st));
  if (n_obs > 0) CK(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)n_obs, st));
  CK(cudaEventRecord(ctx->ev[1], st));
  const char* db = (const char*)ctx->d_prog.p;
  constexpr int kSmemBits = 12;
  for (int i = 0; i < M;) {
    const int nb = progs[order[i]].n_bits;
    int j = i;
    while (j < M && progs[order[j]].n_bits == nb) ++j;
    int per_launch = j - i;
    size_t smem = 0;
    int64_t stride = 0;
    if (nb <= kSmemBits) smem = sizeof(double2) << nb;
    else {
      stride = int64_t(1) << nb;
      per_launch = std::min(per_launch, std::max(1, std::min(4 * ctx->sm_count, (int)((int64_t(8) << 30) / (stride * 16)))));
      CK(ctx->d_scratch.reserve(sizeof(double2) * (size_t)stride * per_launch));
    }
    for (int f = i; f < j; f += per_launch) {
      SvLaunch L;
      L.first_circuit = f;
      L.n_circuits = std::min(per_launch, j - f);
      L.circ_desc = (const int32_t*)(db + o_cd);
      L.ops = (const SvOp*)(db + o_ops);
      L.mats = (const double*)(db + o_mats);
      L.obs_desc = (const int64_t*)(db + o_obs);
      L.term_x = (const uint32_t*)(db + o_tx);
      L.term_z = (const uint32_t*)(db + o_tz);
      L.term_ny = (const int32_t*)(db + o_tny);
      L.term_coeff = (const double*)(db + o_tc);
      L.out = d_out;
      L.scratch = (double2*)ctx->d_scratch.p;
      L.scratch_stride = stride;
      L.smem_bits = kSmemBits;
      sv_circuit_kernel<<<L.n_circuits, kSvThreads, smem, st>>>(L);
      CK(cudaGetLastError());
      ctx->stats.n_other_launches++;
    }
    i = j;
  }
  CK(cudaEventRecord(ctx->ev[2], st));
  if (n_obs > 0) CK(cudaMemcpyAsync(ctx->h_out.p, d_out, sizeof(double) * (size_t)n_obs, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(ctx->ev[3], st));
  CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); ctx->stats.h2d_ms = ms;
  CK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2])); ctx->stats.kernel_ms = ms;
  CK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3])); ctx->stats.d2h_ms = ms;
  if (n_obs > 0) std::memcpy(out_vals, ctx->h_out.p, sizeof(double) * (size_t)n_obs);
  for (int c = 0; c < N; ++c)
    if (out_status[c] != 0)
      for (int64_t o = b->obs_offsets[c]; o < b->obs_offsets[c + 1]; ++o) out_vals[o] = std::nan("");
  return BWQ_OK;
}

// ------------------------------------------------------------------------------------------------
// host-only lowering introspection
// ------------------------------------------------------------------------------------------------
extern "C" int bwq_lower_dm(const bwq_noise_table* table, const bwq_batch* batch, int32_t circuit,
                            int32_t tile_qubits, int32_t low_qubits, bwq_program** out) {
  if (!batch || !out || circuit < 0 || circuit >= batch->n_circuits) return BWQ_ERR_ARG;
  NoiseTable nt;
  char err[256];
  int rc = nt.set(table, err, sizeof err);
  if (rc) { g_create_error = err; return rc; }
  LowerOptions lo;
  lo.tile_qubits = tile_qubits ? tile_qubits : 6;
  lo.low_qubits = low_qubits < 0 ? 0 : (low_qubits ? low_qubits : 2);
  bwq_program* p = new bwq_program();
  lower_dm_circuit(nt, *batch, circuit, lo, &p->p);
  // table-sourced ops reference the (re-packed) table: export them as batch matrices instead
  for (auto& op : p->p.ops) {
    if (op.src != 1) continue;
    int n = (op.kind == K_RELAX2 || op.kind == K_RELAX2_SW) ? 25 : 256;
    int64_t off = (int64_t)p->p.mats.size();
    p->p.mats.insert(p->p.mats.end(), nt.data.begin() + op.off, nt.data.begin() + op.off + n);
    while (p->p.mats.size() % 4) p->p.mats.push_back(0.0);
    op.src = 0;
    op.off = off;
  }
  *out = p;
  return BWQ_OK;
}

extern "C" void bwq_program_free(bwq_program* p) { delete p; }

extern "C" int bwq_program_sizes(const bwq_program* p, int64_t s[8]) {
  if (!p || !s) return BWQ_ERR_ARG;
  s[0] = p->p.n_digits; s[1] = (int64_t)p->p.sweeps.size(); s[2] = (int64_t)p->p.passes.size();
  s[3] = (int64_t)p->p.ops.size(); s[4] = (int64_t)p->p.mats.size(); s[5] = p->p.status;
  s[6] = (int64_t)p->p.term_index.size(); s[7] = p->p.n_gates;
  return BWQ_OK;
}

extern "C" int bwq_program_read(const bwq_program* p, int32_t* active, int32_t* sweeps, int32_t* passes,
                                int64_t* ops, double* mats, int64_t* term_index, double* term_coeff) {
  if (!p) return BWQ_ERR_ARG;
  const CircuitProgram& q = p->p;
  const int kq = q.sweeps.empty() ? 0 : 0;
  (void)kq;
  if (active) for (size_t i = 0; i < q.active.size(); ++i) active[i] = q.active[i];
  if (sweeps)
    for (size_t i = 0; i < q.sweeps.size(); ++i) {
      int32_t* s = sweeps + 10 * i;
      s[0] = q.sweeps[i].pass_begin;
      for (int k = 0; k < 8; ++k) s[1 + k] = q.sweeps[i].pos[k];
      s[9] = q.sweeps[i].pass_end;
    }
  if (passes)
    for (size_t i = 0; i < q.passes.size(); ++i) {
      passes[3 * i] = q.passes[i].sa; passes[3 * i + 1] = q.passes[i].sb; passes[3 * i + 2] = q.passes[i].op_end;
    }
  if (ops)
    for (size_t i = 0; i < q.ops.size(); ++i) { ops[2 * i] = q.ops[i].kind | (int64_t(q.ops[i].src) << 8); ops[2 * i + 1] = q.ops[i].off; }
  if (mats && !q.mats.empty()) std::memcpy(mats, q.mats.data(), q.mats.size() * sizeof(double));
  if (term_index && !q.term_index.empty()) std::memcpy(term_index, q.term_index.data(), q.term_index.size() * sizeof(int64_t));
  if (term_coeff && !q.term_coeff.empty()) std::memcpy(term_coeff, q.term_coeff.data(), q.term_coeff.size() * sizeof(double));
  return BWQ_OK;
}
