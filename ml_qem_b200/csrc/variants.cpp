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
#include <cmath>
#include <cstring>

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

int expand_variants(const bwq_batch& b, const bwq_variants& v, ExpandedBatch* out) {
  const int n_folds = v.n_folds > 0 ? v.n_folds : 1;
  const int n_tw = v.n_twirls > 0 ? v.n_twirls : 1;
  const bool twirl = v.n_twirls > 0;
  for (int f = 0; f < n_folds; ++f) {
    const int fac = v.n_folds > 0 ? v.folds[f] : 1;
    if (fac < 1 || fac % 2 == 0) return BWQ_ERR_ARG;  // local folding: odd factors
  }
  const int n_var = n_folds * n_tw;
  const int64_t N = (int64_t)b.n_circuits * n_var;
  if (N > INT32_MAX) return BWQ_ERR_ARG;
  out->n_variants = n_var;
  out->n_qubits.resize(N);
  out->op_offsets.assign(N + 1, 0);
  out->obs_offsets.assign(N + 1, 0);
  out->status.assign(b.n_circuits, 0);
  out->params.assign(b.params, b.params + b.n_params);
  const uint32_t pi_idx = (uint32_t)out->params.size();
  out->params.push_back(M_PI);
  out->ops.clear();
  out->term_offsets.assign(1, 0);
  out->term_x.clear(); out->term_z.clear(); out->term_coeff.clear();
  auto pauli = [&](int p, uint8_t q) {
    if (p == 2 || p == 3) out->ops.push_back(bwq_op{BWQ_G_RZ, q, 0, pi_idx});
    if (p == 1 || p == 2) out->ops.push_back(bwq_op{BWQ_G_X, q, 0, 0});
  };
  for (int c = 0; c < b.n_circuits; ++c) {
    const int64_t g0 = b.op_offsets[c], g1 = b.op_offsets[c + 1];
    const int64_t o0 = b.obs_offsets[c], o1 = b.obs_offsets[c + 1];
    for (int f = 0; f < n_folds; ++f) {
      const int fac = v.n_folds > 0 ? v.folds[f] : 1;
      for (int t = 0; t < n_tw; ++t) {
        const int64_t vi = (int64_t)c * n_var + (int64_t)f * n_tw + t;
        out->n_qubits[vi] = b.n_qubits[c];
        uint64_t k_cx = 0;
        for (int64_t g = g0; g < g1; ++g) {
          const bwq_op op = b.ops[g];
          if (!gate_is_2q(op.opcode)) { out->ops.push_back(op); continue; }
          int pc = 0, pt = 0, qc = 0, qt = 0;
          if (twirl && op.opcode == BWQ_G_CX) {
            const uint32_t d = twirl_draw(v.seed, (uint64_t)c, (uint64_t)t, k_cx++);
            pc = (int)(d & 3u); pt = (int)(d >> 2);
            cx_conjugate(pc, pt, &qc, &qt);
            pauli(pc, op.q0); pauli(pt, op.q1);
          }
          // G (G^dagger G)^((fac-1)/2)
          bwq_op inv = op;
          if (fac > 1) {
            switch (op.opcode) {
              case BWQ_G_CX: case BWQ_G_CY: case BWQ_G_CZ: case BWQ_G_CH: case BWQ_G_SWAP: case BWQ_G_ECR: break;
              case BWQ_G_CRX: case BWQ_G_CRY: case BWQ_G_CRZ: case BWQ_G_CP:
              case BWQ_G_RZZ: case BWQ_G_RXX: case BWQ_G_RYY: case BWQ_G_RZX: {
                if ((int64_t)op.param_idx + 1 > b.n_params) { out->status[c] = BWQ_CIRC_BAD_OP; break; }
                inv.param_idx = (uint32_t)out->params.size();
                out->params.push_back(-b.params[op.param_idx]);
                break; }
              case BWQ_G_CU3: {  // u3(t, p, l)^-1 = u3(-t, -l, -p)
                if ((int64_t)op.param_idx + 3 > b.n_params) { out->status[c] = BWQ_CIRC_BAD_OP; break; }
                inv.param_idx = (uint32_t)out->params.size();
                const double* p = b.params + op.param_idx;
                out->params.push_back(-p[0]); out->params.push_back(-p[2]); out->params.push_back(-p[1]);
                break; }
              case BWQ_G_UNITARY2: {
                if ((int64_t)op.param_idx + 32 > b.n_params) { out->status[c] = BWQ_CIRC_BAD_OP; break; }
                inv.param_idx = (uint32_t)out->params.size();
                const double* p = b.params + op.param_idx;
                for (int r = 0; r < 4; ++r)
                  for (int cc = 0; cc < 4; ++cc) { out->params.push_back(p[2 * (cc * 4 + r)]); out->params.push_back(-p[2 * (cc * 4 + r) + 1]); }
                break; }
              default: out->status[c] = BWQ_CIRC_BAD_OP; break;  // no inverse rule (iswap)
            }
          }
          for (int r = 0; r < fac; ++r) out->ops.push_back((r & 1) ? inv : op);
          if (twirl && op.opcode == BWQ_G_CX) { pauli(qc, op.q0); pauli(qt, op.q1); }
        }
        out->op_offsets[vi + 1] = (int64_t)out->ops.size();
        // observables of the base circuit, replicated
        for (int64_t o = o0; o < o1; ++o) {
          for (int64_t tt = b.term_offsets[o]; tt < b.term_offsets[o + 1]; ++tt) {
            out->term_x.push_back(b.term_x[tt]); out->term_z.push_back(b.term_z[tt]); out->term_coeff.push_back(b.term_coeff[tt]);
          }
          out->term_offsets.push_back((int64_t)out->term_x.size());
        }
        out->obs_offsets[vi + 1] = out->obs_offsets[vi] + (o1 - o0);
      }
    }
  }
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
