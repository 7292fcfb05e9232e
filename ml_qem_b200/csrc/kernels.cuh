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
  int32_t first_circuit;       // index (sorted order) of slot 0
  const int32_t* sweep_range;  // [2 * n_circuits] absolute {begin, end} per sorted circuit
  const SweepDesc* sweeps;
  const PassDesc* passes;
  const DevOp* ops;
  const double* mats;
  const double* noise;
};

// ---------------------------------------------------------------------------------------------
// register-pass ops; v[da + 4*db]
// ---------------------------------------------------------------------------------------------
template <bool SW> __device__ __forceinline__ constexpr int idx2(int d0, int d1) {
  return SW ? (d1 + 4 * d0) : (d0 + 4 * d1);  // (q0,q1) digits -> register index
}

template <bool ON_B> __device__ __forceinline__ void op_dense1(double (&v)[16], const double* __restrict__ m) {
  double a[16];
  const double2* m2 = reinterpret_cast<const double2*>(m);
#pragma unroll
  for (int i = 0; i < 8; ++i) { double2 t = __ldg(m2 + i); a[2 * i] = t.x; a[2 * i + 1] = t.y; }
#pragma unroll
  for (int o = 0; o < 4; ++o) {  // the other digit
    double x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = v[ON_B ? (o + 4 * j) : (j + 4 * o)];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double s = a[i * 4] * x[0];
      s = fma(a[i * 4 + 1], x[1], s);
      s = fma(a[i * 4 + 2], x[2], s);
      s = fma(a[i * 4 + 3], x[3], s);
      v[ON_B ? (o + 4 * i) : (i + 4 * o)] = s;
    }
  }
}

// CX Pauli-transfer matrix = signed permutation; index = d_control + 4 * d_target
template <bool CTRL_B> __device__ __forceinline__ void op_cx(double (&v)[16]) {
  constexpr int src[16] = {0, 5, 6, 3, 4, 1, 2, 7, 11, 14, 13, 8, 15, 10, 9, 12};
  constexpr int sgn[16] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, 1, 1, -1, 1, 1};
  double w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = v[i];
#pragma unroll
  for (int dc = 0; dc < 4; ++dc)
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      const int i = dc + 4 * dt, j = src[i];
      const int jc = j & 3, jt = j >> 2;
      const double x = w[CTRL_B ? (jt + 4 * jc) : (jc + 4 * jt)];
      v[CTRL_B ? (dt + 4 * dc) : (dc + 4 * dt)] = sgn[i] > 0 ? x : -x;
    }
}

// out[i] = d[i] in[i]; out[Z,b] += ca[b] in[I,b]; out[a,Z] += cb[a] in[a,I]; out[Z,Z] += cab in[I,I]
template <bool SW> __device__ __forceinline__ void op_relax2(double (&v)[16], const double* __restrict__ m) {
  double w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = v[i];
#pragma unroll
  for (int d1 = 0; d1 < 4; ++d1)
#pragma unroll
    for (int d0 = 0; d0 < 4; ++d0) v[idx2<SW>(d0, d1)] = __ldg(m + d0 + 4 * d1) * w[idx2<SW>(d0, d1)];
#pragma unroll
  for (int b = 0; b < 4; ++b) v[idx2<SW>(3, b)] = fma(__ldg(m + 16 + b), w[idx2<SW>(0, b)], v[idx2<SW>(3, b)]);
#pragma unroll
  for (int a = 0; a < 4; ++a) v[idx2<SW>(a, 3)] = fma(__ldg(m + 20 + a), w[idx2<SW>(a, 0)], v[idx2<SW>(a, 3)]);
  v[idx2<SW>(3, 3)] = fma(__ldg(m + 24), w[idx2<SW>(0, 0)], v[idx2<SW>(3, 3)]);
}

template <bool SW> __device__ __forceinline__ void op_dense2(double (&v)[16], const double* __restrict__ m) {
  double w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = v[i];
  const double2* m2 = reinterpret_cast<const double2*>(m);
#pragma unroll
  for (int i1 = 0; i1 < 4; ++i1)
#pragma unroll
    for (int i0 = 0; i0 < 4; ++i0) {
      const int row = i0 + 4 * i1;
      double s = 0.0;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const double2 t = __ldg(m2 + row * 8 + jj);
        const int j = 2 * jj;
        s = fma(t.x, w[idx2<SW>(j & 3, j >> 2)], s);
        s = fma(t.y, w[idx2<SW>((j + 1) & 3, (j + 1) >> 2)], s);
      }
      v[idx2<SW>(i0, i1)] = s;
    }
}

__device__ __forceinline__ void run_ops(double (&v)[16], const DmLaunch& L, int op_begin, int op_end) {
  for (int o = op_begin; o < op_end; ++o) {
    const int4 raw = __ldg(reinterpret_cast<const int4*>(L.ops + o));
    const int kind = raw.x;
    const int64_t off = (int64_t(uint32_t(raw.w)) << 32) | uint32_t(raw.z);
    const double* m = (raw.y ? L.noise : L.mats) + off;
    switch (kind) {
      case K_DENSE1_A: op_dense1<false>(v, m); break;
      case K_DENSE1_B: op_dense1<true>(v, m); break;
      case K_CX_AB: op_cx<false>(v); break;
      case K_CX_BA: op_cx<true>(v); break;
      case K_RELAX2: op_relax2<false>(v, m); break;
      case K_RELAX2_SW: op_relax2<true>(v, m); break;
      case K_DENSE2: op_dense2<false>(v, m); break;
      case K_DENSE2_SW: op_dense2<true>(v, m); break;
      default: break;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K1/K2/K3: tile sweep
// ---------------------------------------------------------------------------------------------
template <int KQ> struct SweepCfg {
  static constexpr int kElems = 1 << (2 * KQ);
  static constexpr int kGroups = kElems / 16;
  static constexpr int kThreads = kGroups >= 256 ? 256 : (kGroups >= 32 ? kGroups : 32);
};

template <int KQ>
__global__ void __launch_bounds__(SweepCfg<KQ>::kThreads)
dm_sweep_kernel(const DmLaunch L, const int sweep_idx) {
  constexpr int E = SweepCfg<KQ>::kElems, G = SweepCfg<KQ>::kGroups, T = SweepCfg<KQ>::kThreads;
  extern __shared__ __align__(16) double tile[];
  const int tid = threadIdx.x;
  const int tiles_log2 = 2 * (L.n_digits - KQ);
  const int64_t slot = int64_t(blockIdx.x) >> tiles_log2;
  const uint32_t t = uint32_t(blockIdx.x) & ((1u << tiles_log2) - 1u);
  const int circ = L.first_circuit + int(slot);
  const int sw_i = __ldg(L.sweep_range + 2 * circ) + sweep_idx;
  if (sw_i >= __ldg(L.sweep_range + 2 * circ + 1)) return;  // this circuit has fewer sweeps
  const int4 swraw = __ldg(reinterpret_cast<const int4*>(L.sweeps + sw_i));
  const int pass_begin = swraw.x, pass_end = swraw.y;
  int pos[KQ];
  {
    const uint64_t pk = (uint64_t(uint32_t(swraw.w)) << 32) | uint32_t(swraw.z);
#pragma unroll
    for (int s = 0; s < KQ; ++s) pos[s] = int((pk >> (8 * s)) & 0xff);
  }
  // scatter the tile id over the digit positions that are NOT resident in the tile
  int64_t base = 0;
  {
    uint32_t rest = t;
    int s = 0;
    for (int d = 0; d < L.n_digits; ++d) {
      if (s < KQ && pos[s] == d) { ++s; continue; }
      base |= int64_t(rest & 3u) << (2 * d);
      rest >>= 2;
    }
  }
  double* __restrict__ g = L.states + slot * L.stride + base;

  // ---- load (or synthesise |0..0><0..0| on the first sweep)
  if (sweep_idx == 0) {
    const bool tile_ok = ((t ^ (t >> 1)) & 0x55555555u) == 0u;  // all outside digits in {I,Z}
#pragma unroll
    for (int u = tid; u < E / 2; u += T) {
      const uint32_t j = 2u * u;
      const bool ok0 = tile_ok && (((j ^ (j >> 1)) & 0x55555555u) == 0u);
      const bool ok1 = tile_ok && ((((j + 1) ^ ((j + 1) >> 1)) & 0x55555555u) == 0u);
      reinterpret_cast<double2*>(tile)[u] = make_double2(ok0 ? 1.0 : 0.0, ok1 ? 1.0 : 0.0);
    }
  } else {
#pragma unroll
    for (int u = tid; u < E / 2; u += T) {
      const uint32_t j = 2u * u;
      int64_t off = 0;
#pragma unroll
      for (int s = 0; s < KQ; ++s) off |= int64_t((j >> (2 * s)) & 3u) << (2 * pos[s]);
      reinterpret_cast<double2*>(tile)[u] = *reinterpret_cast<const double2*>(g + off);
    }
  }

  // ---- register passes
  for (int p = pass_begin; p < pass_end; ++p) {
    __syncthreads();
    const int4 praw = __ldg(reinterpret_cast<const int4*>(L.passes + p));
    const int sa = praw.z & 0xff, sb = (praw.z >> 8) & 0xff;
    const int lo = min(sa, sb), hi = max(sa, sb);
    const int stride_a = 1 << (2 * sa), stride_b = 1 << (2 * sb);
    for (int grp = tid; grp < G; grp += T) {
      const uint32_t low = grp & ((1u << (2 * lo)) - 1u);
      const uint32_t mid = (grp >> (2 * lo)) & ((1u << (2 * (hi - lo - 1))) - 1u);
      const uint32_t high = grp >> (2 * (hi - 1));
      const uint32_t b0 = low | (mid << (2 * lo + 2)) | (high << (2 * hi + 2));
      double v[16];
#pragma unroll
      for (int db = 0; db < 4; ++db)
#pragma unroll
        for (int da = 0; da < 4; ++da) v[da + 4 * db] = tile[b0 + da * stride_a + db * stride_b];
      run_ops(v, L, praw.x, praw.y);
#pragma unroll
      for (int db = 0; db < 4; ++db)
#pragma unroll
        for (int da = 0; da < 4; ++da) tile[b0 + da * stride_a + db * stride_b] = v[da + 4 * db];
    }
  }
  __syncthreads();

  // ---- store
#pragma unroll
  for (int u = tid; u < E / 2; u += T) {
    const uint32_t j = 2u * u;
    int64_t off = 0;
#pragma unroll
    for (int s = 0; s < KQ; ++s) off |= int64_t((j >> (2 * s)) & 3u) << (2 * pos[s]);
    *reinterpret_cast<double2*>(g + off) = reinterpret_cast<const double2*>(tile)[u];
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
