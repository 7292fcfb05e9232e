// K0, variant generation: ZNE gate folds and Pauli twirls of a base batch, expanded on the flat gate
// stream inside the library -- no per-variant circuit objects on the caller's side.
//
// Replaces the reference's Python-level variant loops: LocalFoldingAmplifier(gates_to_fold=2) with
// noise factors (1, 3, 5) (docs/tutorials/zne_parallel.py:168-189, 256-270;
// docs/tutorials/derek_files/phase_diagram.ipynb:960) and add_pauli_twirls
// (docs/tutorials/derek_files/phase_diagram.ipynb:776), which build a new QuantumCircuit per variant.
//
// Variant order of base circuit c: for fold f (outer), for twirl t (inner) -> variant f * n_twirls + t.
//   twirl : before every cx a Pauli pair (P_c, P_t) drawn from a counter-based generator
//           (bwq_twirl_draw: a pure function of seed, circuit, twirl and cx index, restated in
//           tests/test_variants.py), after it CX (P_c P_t) CX; Paulis in the backend basis
//           X = x, Y = rz(pi) x, Z = rz(pi), identity emits nothing (ml_qem_b200/families.py);
//           twirl 0 of n_twirls == 0 means "no twirling"
//   fold  : every 2-qubit gate G -> G (G^dagger G)^((f-1)/2) between the twirl Paulis; self-inverse
//           gates repeat f times, rotations negate their angle, cu3 / unitary2 invert explicitly
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>

#include "program.h"

namespace bwq {

static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// Pauli pair (control, target) in {0=I, 1=X, 2=Y, 3=Z}^2 of cx number k of twirl t of circuit c
uint32_t twirl_draw(uint64_t seed, uint64_t c, uint64_t t, uint64_t k) {
  const uint64_t h = splitmix64(splitmix64(splitmix64(seed ^ c) ^ t) ^ k);
  return (uint32_t)(h & 15u);  // pc = bits 0..1, pt = bits 2..3
}

// CX conjugation of a Pauli pair, sign dropped: X_c -> X_c X_t, Z_t -> Z_c Z_t
static inline void cx_conjugate(int pc, int pt, int* qc, int* qt) {
  static const int sx[4] = {0, 1, 1, 0}, sz[4] = {0, 0, 1, 1};  // I X Y Z -> (x, z)
  static const int code[2][2] = {{0, 3}, {1, 2}};               // [x][z] -> Pauli
  *qc = code[sx[pc]][sz[pc] ^ sz[pt]];
  *qt = code[sx[pt] ^ sx[pc]][sz[pt]];
}

// Two passes over the base circuits, both parallel over circuits (threads): (1) sizes of every
// variant (ops, inverse parameters), (2) fill at the offsets the prefix sums give.
template <class F> static void par_for(int n, int threads, F f) {
  threads = std::max(1, std::min(threads, n));
  if (threads == 1) { for (int i = 0; i < n; ++i) f(i); return; }
  std::atomic<int> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&] { for (;;) { int i = next.fetch_add(8); if (i >= n) break; for (int j = i; j < std::min(n, i + 8); ++j) f(j); } });
  for (auto& th : pool) th.join();
}

int expand_variants(const bwq_batch& b, const bwq_variants& v, ExpandedBatch* out, int threads, const ParallelRunner* runner) {
  // the context's persistent worker pool when the caller has one, else threads spawned per call
  auto par_for = [&](int n, int thr, const std::function<void(int)>& f) {
    if (runner && *runner) (*runner)(n, f);
    else bwq::par_for(n, thr, f);
  };
  const int n_folds = v.n_folds > 0 ? v.n_folds : 1;
  const int n_tw = v.n_twirls > 0 ? v.n_twirls : 1;
  const bool twirl = v.n_twirls > 0;
  int max_fac = 1;
  for (int f = 0; f < n_folds; ++f) {
    const int fac = v.n_folds > 0 ? v.folds[f] : 1;
    if (fac < 1 || fac % 2 == 0) return BWQ_ERR_ARG;  // local folding: odd factors
    max_fac = std::max(max_fac, fac);
  }
  const int n_var = n_folds * n_tw;
  const int64_t N = (int64_t)b.n_circuits * n_var;
  if (N > INT32_MAX) return BWQ_ERR_ARG;
  out->n_variants = n_var;
  out->n_qubits.resize(N);
  out->op_offsets.assign(N + 1, 0);
  out->obs_offsets.assign(N + 1, 0);
  out->status.assign(b.n_circuits, 0);
  static const int pauli_ops[4] = {0, 1, 2, 1};  // gates a Pauli costs: I -, X x, Y rz x, Z rz
  auto inverse_params = [](uint16_t opc) {       // parameters the inverse gate appends (-1: no rule)
    switch (opc) {
      case BWQ_G_CX: case BWQ_G_CY: case BWQ_G_CZ: case BWQ_G_CH: case BWQ_G_SWAP: case BWQ_G_ECR: return 0;
      case BWQ_G_CRX: case BWQ_G_CRY: case BWQ_G_CRZ: case BWQ_G_CP:
      case BWQ_G_RZZ: case BWQ_G_RXX: case BWQ_G_RYY: case BWQ_G_RZX: return 1;
      case BWQ_G_CU3: return 3;
      case BWQ_G_UNITARY2: return 32;
      default: return -1;
    }
  };
  // ---- pass 1: ops per variant, inverse parameters per base circuit (shared by its variants)
  std::vector<int64_t> inv_cnt(b.n_circuits + 1, 0);
  par_for(b.n_circuits, threads, [&](int c) {
    const int64_t g0 = b.op_offsets[c], g1 = b.op_offsets[c + 1];
    int64_t n1 = 0, n2 = 0, ninv = 0;
    for (int64_t g = g0; g < g1; ++g) {
      const bwq_op& op = b.ops[g];
      if (!gate_is_2q(op.opcode)) { ++n1; continue; }
      ++n2;
      if (max_fac > 1) {
        const int k = inverse_params(op.opcode);
        if (k < 0 || (int64_t)op.param_idx + k > b.n_params) out->status[c] = BWQ_CIRC_BAD_OP;
        else ninv += k;
      }
    }
    inv_cnt[c + 1] = ninv;
    for (int f = 0; f < n_folds; ++f) {
      const int fac = v.n_folds > 0 ? v.folds[f] : 1;
      for (int t = 0; t < n_tw; ++t) {
        int64_t extra = 0;
        if (twirl) {
          uint64_t k_cx = 0;
          for (int64_t g = g0; g < g1; ++g) {
            if (b.ops[g].opcode != BWQ_G_CX) continue;
            const uint32_t d = twirl_draw(v.seed, (uint64_t)c, (uint64_t)t, k_cx++);
            int qc, qt;
            cx_conjugate((int)(d & 3u), (int)(d >> 2), &qc, &qt);
            extra += pauli_ops[d & 3u] + pauli_ops[d >> 2] + pauli_ops[qc] + pauli_ops[qt];
          }
        }
        out->op_offsets[(int64_t)c * n_var + (int64_t)f * n_tw + t + 1] = n1 + n2 * fac + extra;
      }
    }
  });
  for (int64_t i = 0; i < N; ++i) out->op_offsets[i + 1] += out->op_offsets[i];
  for (int c = 0; c < b.n_circuits; ++c) inv_cnt[c + 1] += inv_cnt[c];
  const int64_t n_par_total = b.n_params + 1 + inv_cnt[b.n_circuits];
  if (n_par_total >= (int64_t(1) << 32)) return BWQ_ERR_ARG;
  out->params.resize((size_t)n_par_total);
  if (b.n_params) std::memcpy(out->params.data(), b.params, sizeof(double) * (size_t)b.n_params);
  const uint32_t pi_idx = (uint32_t)b.n_params;
  out->params[pi_idx] = M_PI;
  out->ops.resize((size_t)out->op_offsets[N]);
  // observables replicated per variant
  {
    std::vector<int64_t> tcount(b.n_circuits + 1, 0);
    for (int c = 0; c < b.n_circuits; ++c) {
      const int64_t o0 = b.obs_offsets[c], o1 = b.obs_offsets[c + 1];
      tcount[c + 1] = tcount[c] + (b.term_offsets[o1] - b.term_offsets[o0]) * n_var;
      for (int k = 0; k < n_var; ++k) out->obs_offsets[(int64_t)c * n_var + k + 1] = o1 - o0;
    }
    for (int64_t i = 0; i < N; ++i) out->obs_offsets[i + 1] += out->obs_offsets[i];
    out->term_offsets.assign((size_t)out->obs_offsets[N] + 1, 0);
    out->term_x.resize((size_t)tcount[b.n_circuits]);
    out->term_z.resize((size_t)tcount[b.n_circuits]);
    out->term_coeff.resize((size_t)tcount[b.n_circuits]);
    par_for(b.n_circuits, threads, [&](int c) {
      const int64_t o0 = b.obs_offsets[c], o1 = b.obs_offsets[c + 1];
      const int64_t t0 = b.term_offsets[o0], nt = b.term_offsets[o1] - t0;
      for (int k = 0; k < n_var; ++k) {
        const int64_t dst = tcount[c] + (int64_t)k * nt;
        if (nt) {
          std::memcpy(&out->term_x[dst], b.term_x + t0, sizeof(uint64_t) * nt);
          std::memcpy(&out->term_z[dst], b.term_z + t0, sizeof(uint64_t) * nt);
          std::memcpy(&out->term_coeff[dst], b.term_coeff + t0, sizeof(double) * nt);
        }
        const int64_t ob = out->obs_offsets[(int64_t)c * n_var + k];
        for (int64_t o = o0; o < o1; ++o) out->term_offsets[ob + (o - o0) + 1] = dst + (b.term_offsets[o + 1] - t0);
      }
    });
  }
  // ---- pass 2: fill
  par_for(b.n_circuits, threads, [&](int c) {
    const int64_t g0 = b.op_offsets[c], g1 = b.op_offsets[c + 1];
    // inverse parameters of this circuit's parametrised 2-qubit gates (only when some factor > 1)
    std::vector<uint32_t> inv_idx;
    if (max_fac > 1 && !out->status[c]) {
      int64_t cur = b.n_params + 1 + inv_cnt[c];
      for (int64_t g = g0; g < g1; ++g) {
        const bwq_op& op = b.ops[g];
        if (!gate_is_2q(op.opcode)) continue;
        const int k = inverse_params(op.opcode);
        inv_idx.push_back(k > 0 ? (uint32_t)cur : op.param_idx);
        const double* p = b.params + op.param_idx;
        double* q = out->params.data() + cur;
        if (k == 1) q[0] = -p[0];
        else if (k == 3) { q[0] = -p[0]; q[1] = -p[2]; q[2] = -p[1]; }  // u3(t, p, l)^-1 = u3(-t, -l, -p)
        else if (k == 32)
          for (int r = 0; r < 4; ++r)
            for (int cc = 0; cc < 4; ++cc) { q[2 * (r * 4 + cc)] = p[2 * (cc * 4 + r)]; q[2 * (r * 4 + cc) + 1] = -p[2 * (cc * 4 + r) + 1]; }
        cur += std::max(k, 0);
      }
    }
    for (int f = 0; f < n_folds; ++f) {
      const int fac = out->status[c] ? 1 : (v.n_folds > 0 ? v.folds[f] : 1);
      for (int t = 0; t < n_tw; ++t) {
        const int64_t vi = (int64_t)c * n_var + (int64_t)f * n_tw + t;
        out->n_qubits[vi] = b.n_qubits[c];
        bwq_op* w = out->ops.data() + out->op_offsets[vi];
        bwq_op* const w_end = out->ops.data() + out->op_offsets[vi + 1];
        auto pauli = [&](int p, uint8_t q) {
          if (p == 2 || p == 3) *w++ = bwq_op{BWQ_G_RZ, q, 0, pi_idx};
          if (p == 1 || p == 2) *w++ = bwq_op{BWQ_G_X, q, 0, 0};
        };
        uint64_t k_cx = 0;
        size_t k2 = 0;
        for (int64_t g = g0; g < g1; ++g) {
          const bwq_op op = b.ops[g];
          if (!gate_is_2q(op.opcode)) { *w++ = op; continue; }
          int qc = 0, qt = 0;
          const bool tw = twirl && op.opcode == BWQ_G_CX;
          if (tw) {
            const uint32_t d = twirl_draw(v.seed, (uint64_t)c, (uint64_t)t, k_cx++);
            cx_conjugate((int)(d & 3u), (int)(d >> 2), &qc, &qt);
            pauli((int)(d & 3u), op.q0); pauli((int)(d >> 2), op.q1);
          }
          bwq_op inv = op;  // G (G^dagger G)^((fac-1)/2)
          if (!inv_idx.empty()) inv.param_idx = inv_idx[k2];
          ++k2;
          for (int r = 0; r < fac; ++r) *w++ = (r & 1) ? inv : op;
          if (tw) { pauli(qc, op.q0); pauli(qt, op.q1); }
        }
        // a circuit whose folds are impossible (status set) keeps factor 1; pad what pass 1 reserved
        while (w < w_end) *w++ = bwq_op{BWQ_G_ID, 0, 0, 0};
      }
    }
  });
  if (out->params.size() >= (size_t(1) << 32)) return BWQ_ERR_ARG;
  out->view.n_circuits = (int32_t)N;
  out->view.n_qubits = out->n_qubits.data();
  out->view.op_offsets = out->op_offsets.data();
  out->view.ops = out->ops.data();
  out->view.params = out->params.data();
  out->view.n_params = (int64_t)out->params.size();
  out->view.obs_offsets = out->obs_offsets.data();
  out->view.term_offsets = out->term_offsets.data();
  out->view.term_x = out->term_x.data();
  out->view.term_z = out->term_z.data();
  out->view.term_coeff = out->term_coeff.data();
  return BWQ_OK;
}

}  // namespace bwq
