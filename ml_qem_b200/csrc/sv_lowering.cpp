// Statevector sweep planner (host C++): gate stream of one circuit -> SvxProgram.
//
// Replaces, for ideal labels, what the reference obtains from qiskit.primitives.Estimator /
// Statevector.evolve (docs/tutorials/h13_ising_data_gen_tomo.ipynb:811,
// docs/tutorials/vqe_data_gen_parallel.py:31) -- here for wide registers (13..30+ qubits) and,
// with n_global > 0, for amplitudes sharded across 2^n_global GPUs.  See program.h for the
// program layout and the execution model.
//
// Stages: (1) gates -> ops: 1-qubit gates are multiplied into a pending 2x2 per qubit, runs of
// 2-qubit gates on one pair (with the 1-qubit gates sandwiched between them) into one 4x4 that is
// then classified: diagonal (cx rz cx = exp(-i t ZZ): no residency needed), controlled (only the
// target must be resident), SWAP, or dense; (2) ops -> register passes on slot pairs; (3) passes
// -> tile sweeps, greedy in program order; a pass whose slot qubit sits on a global bit blocks,
// and when nothing is placeable the planner evicts the local qubits used furthest in the future
// to the top local positions (SWAP passes) and emits an EXCHANGE; (4) observables: Z-type terms
// are evaluated directly, X/Y terms are grouped into qubit-wise commuting families whose basis
// rotation (H, H Sdg) is appended as ordinary gates, so every reduction is a signed sum of
// |amplitude|^2 and never reads a partner amplitude (or a partner GPU).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "program.h"

namespace bwq {
namespace {

using cd = std::complex<double>;

struct HOp {
  uint8_t kind = 0;
  int8_t target = -1;          // U1 / X: logical target
  int8_t qa = -1, qb = -1;     // U2 / SWAP: logical pair (index i_qa + 2 i_qb); D1 / D2: qubits
  int8_t cond_q = -1;
  uint8_t cond_val = 0;
  int32_t off = -1;            // into Planner::mats (doubles)
};
struct ZzLayer {               // bonds of one fused diagonal layer (logical qubits) and its phase pair
  std::vector<std::pair<int, int>> pairs;
  cd pe, po;
};
struct HPass {
  int q[4] = {-1, -1, -1, -1};  // logical qubits of pass slots 0..3 (-1: any other resident slot)
  std::vector<HOp> ops;
  uint64_t touch = 0;          // logical qubits this pass reads or writes (ordering)
  // fast shape [diagonal ops][one structured 1-qubit op of one kind per slot 0..n-1][diagonal ops]:
  bool fast = true;            // still has that shape
  int n_pre = 0, n_slot = 0;   // diagonal ops before the slot ops, slot ops
  uint8_t slot_kind = 0;       // their common kind
  bool post = false;           // a diagonal op follows the slot ops
  int slot_of(int lq) const { for (int i = 0; i < 4; ++i) if (q[i] == lq) return i; return -1; }
  int n_q() const { int c = 0; for (int i = 0; i < 4; ++i) c += q[i] >= 0; return c; }
};
inline bool is_slot_kind(uint8_t k) { return k == SVO_U1 || k == SVO_R1 || k == SVO_X1; }

inline bool is_zero(cd x) { return x.real() == 0.0 && x.imag() == 0.0; }
inline bool is_one(cd x) { return x.real() == 1.0 && x.imag() == 0.0; }

struct Planner {
  SvxProgram* out;
  int n = 0, nl = 0, g = 0, K = 0, L = 0;
  std::vector<int> phys, logical_at;   // logical <-> physical bit position
  uint64_t touched = 0;                // logical qubits a non-diagonal op has acted on so far
  bool direct_passes = true;           // SvxOptions::direct
  std::vector<double> mats;
  // stage state
  std::vector<HPass> passes;
  std::vector<int> last;
  std::vector<cd> pend;                // 2x2 per qubit
  std::vector<char> has;
  struct Block { int a = -1, b = -1; cd m[16]; bool open = false; };
  std::vector<Block> blocks;           // open 2-qubit runs
  std::vector<int> open_of;            // qubit -> index into blocks or -1
  // ZZ-type diagonal bonds waiting to be fused into one layer op (they commute with every other
  // diagonal op; the layer is emitted before the next non-diagonal op on one of its qubits)
  std::vector<ZzLayer> zz_layers;      // emitted layers (HOp::off of SVO_DZZ)
  ZzLayer pending;
  uint64_t pending_qubits = 0;
  bool fuse_layers = true;
  int zz_chunk = 62;                   // bonds per fused layer op (sharded runs use small chunks: see lower_svx_circuit)

  int32_t push(const double* m, int nd) {
    while (mats.size() % 2) mats.push_back(0.0);
    int32_t off = (int32_t)mats.size();
    mats.insert(mats.end(), m, m + nd);
    return off;
  }
  int32_t push_c(const cd* m, int nc) {
    std::vector<double> t(2 * nc);
    for (int i = 0; i < nc; ++i) { t[2 * i] = m[i].real(); t[2 * i + 1] = m[i].imag(); }
    return push(t.data(), 2 * nc);
  }

  void begin_stage() {
    passes.clear();
    last.assign(n, -1);
    pend.assign(4 * n, cd(0));
    has.assign(n, 0);
    blocks.clear();
    open_of.assign(n, -1);
  }

  // ---- ops -> passes ---------------------------------------------------------------------
  // need: qubits that must own a slot of the pass; touch: every qubit the op reads or writes
  void flush_zz() {
    if (pending.pairs.empty()) return;
    ZzLayer L;
    L.pairs.swap(pending.pairs);
    L.pe = pending.pe; L.po = pending.po;
    const uint64_t touch = pending_qubits;
    pending_qubits = 0;
    HOp op;
    if (L.pairs.size() == 1) {  // a single bond stays a plain 2-qubit diagonal
      op.kind = SVO_D2; op.qa = (int8_t)L.pairs[0].first; op.qb = (int8_t)L.pairs[0].second;
      const cd ph[4] = {L.pe, L.po, L.po, L.pe};
      op.off = push_c(ph, 4);
    } else {
      op.kind = SVO_DZZ;
      op.off = (int32_t)zz_layers.size();
      zz_layers.push_back(std::move(L));
    }
    emit(op, 0, touch);
  }
  void add_zz(int a, int b, cd pe, cd po) {
    bool dup = false;
    for (auto& pr : pending.pairs) dup = dup || (pr.first == a && pr.second == b) || (pr.first == b && pr.second == a);
    if (!pending.pairs.empty() && (dup || pending.pe != pe || pending.po != po || (int)pending.pairs.size() >= zz_chunk)) flush_zz();
    pending.pairs.push_back({a, b});
    pending.pe = pe; pending.po = po;
    pending_qubits |= (1ull << a) | (1ull << b);
  }

  void emit(const HOp& op, uint64_t need, uint64_t touch) {
    touch |= need;
    // a non-diagonal op (it owns slots) or a conditional one on a qubit of the pending layer:
    // the layer goes first (diagonal ops commute with it and may overtake it)
    if ((need || op.cond_q >= 0) && (touch & pending_qubits)) {
      // open 2-qubit runs that are complete bonds of the same layer join it first (their qubits
      // have seen no later op: a later non-diagonal op would have closed the run)
      for (size_t bi = 0; bi < blocks.size(); ++bi) {
        const Block& B = blocks[bi];
        if (!B.open || ((touch >> B.a) & 1) || ((touch >> B.b) & 1)) continue;
        bool diag = true;
        for (int r = 0; r < 4 && diag; ++r)
          for (int c = 0; c < 4; ++c) if (r != c && !is_zero(B.m[r * 4 + c])) { diag = false; break; }
        if (diag && B.m[0] == B.m[15] && B.m[5] == B.m[10] && B.m[0] == pending.pe && B.m[5] == pending.po) close_block((int)bi);
      }
      flush_zz();
    }
    int P = -1;
    for (int q = 0; q < n; ++q) if ((touch >> q) & 1) P = std::max(P, last[q]);
    const bool slot_op = is_slot_kind(op.kind) && op.cond_q < 0;
    const bool diag_op = need == 0;
    // does the op fit pass c (index >= P: everything it depends on sits in passes <= P)?
    auto fits = [&](const HPass& c) {
      int free_slots = 4 - c.n_q();
      for (int q = 0; q < n; ++q)
        if (((need >> q) & 1) && c.slot_of(q) < 0 && --free_slots < 0) return false;
      // keep a fast pass fast: its slot ops are one per slot, of one kind, no diagonal op between them
      if (c.fast && slot_op) {
        const int sl = c.slot_of(op.target);
        if (c.post || c.n_slot >= 4 || (c.n_slot > 0 && c.slot_kind != op.kind)) return false;
        if (sl >= 0 ? sl != c.n_slot : c.q[c.n_slot] >= 0) return false;  // its slot must be pass slot n_slot
      }
      // a pass stays well inside the shared-memory program buffer
      return pass_bytes(c) + 8 * op_words_max(op) + 64 <= kBlockBytes / 2;
    };
    // diagonal ops go to the earliest legal pass (they must not delay later ops on their qubits),
    // slot ops to the most recent pass with room (four slot ops share one gather / scatter)
    int chosen = -1;
    const int n_pass = (int)passes.size();
    if (diag_op || !slot_op) {
      for (int c = std::max(P, 0); c < n_pass && c < std::max(P, 0) + 4; ++c)
        if ((P >= 0 || c == n_pass - 1) && fits(passes[c])) { chosen = c; break; }
    } else {
      for (int c = n_pass - 1; c >= std::max(P, 0) && c >= n_pass - 6; --c)
        if (fits(passes[c])) { chosen = c; break; }
    }
    if (chosen < 0) {
      chosen = n_pass;
      passes.push_back(HPass());
    }
    P = chosen;
    HPass& p = passes[P];
    if (p.fast && slot_op) {
      p.q[p.n_slot] = op.target;  // (empty or already this qubit)
      p.slot_kind = op.kind;
      ++p.n_slot;
    } else if (p.fast && diag_op) {
      if (p.n_slot == 0) ++p.n_pre; else p.post = true;
    } else {
      p.fast = false;
    }
    for (int q = 0; q < n; ++q)
      if (((need >> q) & 1) && p.slot_of(q) < 0)
        for (int i = 0; i < 4; ++i) if (p.q[i] < 0) { p.q[i] = q; break; }
    p.ops.push_back(op);
    p.touch |= touch;
    for (int q = 0; q < n; ++q) if ((touch >> q) & 1) last[q] = P;
  }

  static bool diag2(const cd* m) { return is_zero(m[1]) && is_zero(m[2]); }
  static bool ident2(const cd* m) { return diag2(m) && is_one(m[0]) && is_one(m[3]); }
  static bool is_x(const cd* m) { return is_zero(m[0]) && is_zero(m[3]) && is_one(m[1]) && is_one(m[2]); }

  // 1-qubit map on q (optionally conditional on cond_q == cond_val); hint: extra slot request
  void emit_u1(int q, const cd* m, int cond_q, int cond_val, uint64_t hint) {
    if (ident2(m)) return;
    HOp op;
    op.cond_q = (int8_t)cond_q; op.cond_val = (uint8_t)cond_val;
    uint64_t touch = (1ull << q) | (cond_q >= 0 ? (1ull << cond_q) : 0);
    if (diag2(m)) {
      if (cond_q < 0) {
        op.kind = SVO_D1; op.qa = (int8_t)q;
        const cd ph[2] = {m[0], m[3]};
        op.off = push_c(ph, 2);
      } else {  // conditional phase = 2-qubit diagonal, index b_q + 2 b_cond
        op.kind = SVO_D2; op.qa = (int8_t)q; op.qb = (int8_t)cond_q; op.cond_q = -1;
        cd ph[4] = {1, 1, 1, 1};
        ph[0 + 2 * cond_val] = m[0]; ph[1 + 2 * cond_val] = m[3];
        op.off = push_c(ph, 4);
      }
      emit(op, 0, touch);
      return;
    }
    op.target = (int8_t)q;
    double st[4];
    int skind = 0;
    if (is_x(m)) op.kind = SVO_X;
    else if (cond_q < 0 && structured && (skind = classify_u1(m, st)) != 0) { op.kind = (uint8_t)skind; op.off = push(st, 4); }
    else { op.kind = SVO_U1; op.off = push_c(m, 4); }
    emit(op, (1ull << q) | hint, touch);
  }

  // Unconditional 1-qubit unitary up to a global phase: all entries real (SVO_R1) or real diagonal
  // with imaginary off-diagonal (SVO_X1) -- half the multiply-adds of the general complex 2x2.
  bool structured = true;
  static int classify_u1(const cd* m, double* out) {
    int big = 0;
    for (int i = 1; i < 4; ++i) if (std::abs(m[i]) > std::abs(m[big])) big = i;
    const double a0 = std::arg(m[big]);
    const double tol = 1e-14;
    for (int shift = 0; shift < 2; ++shift) {
      const cd rot = std::polar(1.0, -(a0 - shift * (M_PI / 2)));
      cd r[4];
      for (int i = 0; i < 4; ++i) r[i] = m[i] * rot;
      bool real = true, xt = true;
      for (int i = 0; i < 4; ++i) real = real && std::abs(r[i].imag()) <= tol;
      xt = std::abs(r[0].imag()) <= tol && std::abs(r[3].imag()) <= tol && std::abs(r[1].real()) <= tol && std::abs(r[2].real()) <= tol;
      if (real) { for (int i = 0; i < 4; ++i) out[i] = r[i].real(); return SVO_R1; }
      if (xt) { out[0] = r[0].real(); out[1] = r[3].real(); out[2] = r[1].imag(); out[3] = r[2].imag(); return SVO_X1; }
    }
    return 0;
  }

  void flush1(int q, uint64_t hint) {
    if (!has[q]) return;
    has[q] = 0;
    cd m[4] = {pend[4 * q], pend[4 * q + 1], pend[4 * q + 2], pend[4 * q + 3]};
    emit_u1(q, m, -1, 0, hint);
  }

  void close_block(int bi) {
    Block& B = blocks[bi];
    if (!B.open) return;
    B.open = false;
    open_of[B.a] = open_of[B.b] = -1;
    const int a = B.a, b = B.b;
    const cd* m = B.m;
    bool diag = true, ctl_a = true, ctl_b = true;
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        if (is_zero(m[r * 4 + c])) continue;
        if (r != c) diag = false;
        if ((r & 1) != (c & 1)) ctl_a = false;
        if ((r >> 1) != (c >> 1)) ctl_b = false;
      }
    if (diag) {
      if (is_one(m[0]) && is_one(m[5]) && is_one(m[10]) && is_one(m[15])) return;
      if (fuse_layers && m[0] == m[15] && m[5] == m[10]) { add_zz(a, b, m[0], m[5]); return; }  // exp(-i t ZZ) type
      HOp op; op.kind = SVO_D2; op.qa = (int8_t)a; op.qb = (int8_t)b;
      const cd ph[4] = {m[0], m[5], m[10], m[15]};
      op.off = push_c(ph, 4);
      emit(op, 0, (1ull << a) | (1ull << b));
      return;
    }
    if (ctl_a) {  // block diagonal in a: target b
      for (int v = 0; v < 2; ++v) {
        const cd s[4] = {m[(v + 0) * 4 + (v + 0)], m[(v + 0) * 4 + (v + 2)], m[(v + 2) * 4 + (v + 0)], m[(v + 2) * 4 + (v + 2)]};
        emit_u1(b, s, a, v, 0);
      }
      return;
    }
    if (ctl_b) {
      for (int v = 0; v < 2; ++v) {
        const cd s[4] = {m[(2 * v) * 4 + 2 * v], m[(2 * v) * 4 + 2 * v + 1], m[(2 * v + 1) * 4 + 2 * v], m[(2 * v + 1) * 4 + 2 * v + 1]};
        emit_u1(a, s, b, v, 0);
      }
      return;
    }
    HOp op; op.qa = (int8_t)a; op.qb = (int8_t)b;
    bool swp = true;
    for (int r = 0; r < 4 && swp; ++r)
      for (int c = 0; c < 4; ++c) {
        const int want = (r == 0 && c == 0) || (r == 3 && c == 3) || (r == 1 && c == 2) || (r == 2 && c == 1);
        if (want ? !is_one(m[r * 4 + c]) : !is_zero(m[r * 4 + c])) { swp = false; break; }
      }
    if (swp) op.kind = SVO_SWAP;
    else { op.kind = SVO_U2; op.off = push_c(m, 16); }
    emit(op, (1ull << a) | (1ull << b), 0);
  }

  void gate1(int q, const cd* u) {
    if (has[q]) {
      cd t[4];
      const cd* p = &pend[4 * q];
      for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) t[i * 2 + j] = u[i * 2] * p[j] + u[i * 2 + 1] * p[2 + j];
      for (int i = 0; i < 4; ++i) pend[4 * q + i] = t[i];
    } else {
      for (int i = 0; i < 4; ++i) pend[4 * q + i] = u[i];
      has[q] = 1;
    }
  }

  // 4x4 u on (q0,q1), local index i_q0 + 2 i_q1
  void gate2(int q0, int q1, const cd* u) {
    int bi = open_of[q0];
    if (bi >= 0 && bi == open_of[q1]) {
      Block& B = blocks[bi];
      // orient u to the block's (a,b)
      cd v[16];
      if (B.a == q0) std::memcpy(v, u, sizeof v);
      else
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
          const int rr = ((r & 1) << 1) | (r >> 1), cc = ((c & 1) << 1) | (c >> 1);
          v[rr * 4 + cc] = u[r * 4 + c];
        }
      // absorb the pending 1-qubit gates of a and b: M <- v * (Pb (x) Pa) * M
      cd k[16];
      cd ia[4] = {1, 0, 0, 1}, ib[4] = {1, 0, 0, 1};
      const cd* pa = has[B.a] ? &pend[4 * B.a] : ia;
      const cd* pb = has[B.b] ? &pend[4 * B.b] : ib;
      for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c)
        k[r * 4 + c] = pa[(r & 1) * 2 + (c & 1)] * pb[(r >> 1) * 2 + (c >> 1)];
      has[B.a] = has[B.b] = 0;
      cd t[16], w[16];
      for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
        cd s = 0;
        for (int x = 0; x < 4; ++x) s += k[r * 4 + x] * B.m[x * 4 + c];
        t[r * 4 + c] = s;
      }
      for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
        cd s = 0;
        for (int x = 0; x < 4; ++x) s += v[r * 4 + x] * t[x * 4 + c];
        w[r * 4 + c] = s;
      }
      std::memcpy(B.m, w, sizeof w);
      return;
    }
    if (open_of[q0] >= 0) close_block(open_of[q0]);
    if (open_of[q1] >= 0) close_block(open_of[q1]);
    // pending 1-qubit gates go first; non-diagonal ones ask for the pair's slots so that the
    // whole run shares one register pass
    uint64_t hint = 0;
    for (int q : {q0, q1})
      if (has[q] && !diag2(&pend[4 * q])) hint |= 1ull << q;
    flush1(q0, hint);
    flush1(q1, hint);
    Block B;
    B.a = q0; B.b = q1; B.open = true;
    std::memcpy(B.m, u, sizeof B.m);
    blocks.push_back(B);
    open_of[q0] = open_of[q1] = (int)blocks.size() - 1;
  }

  void end_gates() {
    for (size_t i = 0; i < blocks.size(); ++i) close_block((int)i);
    for (int q = 0; q < n; ++q) flush1(q, 0);
    flush_zz();
  }

  // ---- passes -> sweeps / exchanges --------------------------------------------------------
  static int op_words(uint8_t k) {
    return k == SVO_U1 ? 8 : k == SVO_U2 ? 32 : (k == SVO_D1 || k == SVO_R1 || k == SVO_X1) ? 4 : k == SVO_D2 ? 8 : 0;
  }
  // upper bound for the block budget: a fused layer's size depends on the physical mapping at emission
  int op_words_max(const HOp& o) const {
    if (o.kind != SVO_DZZ) return op_words(o.kind);
    const int K = (int)zz_layers[o.off].pairs.size();
    return 2 + K + 2 * (K + 1);
  }
  int pass_bytes(const HPass& p) const {
    int b = (int)sizeof(SvPassHdr) + (int)sizeof(SvBlockOp) * (int)p.ops.size();
    for (const HOp& o : p.ops) b += 8 * op_words_max(o) + 8;
    return b;
  }
  // parameters of a fused layer under the current logical -> physical map (8-byte words)
  std::vector<uint64_t> zz_params(const ZzLayer& L) const {
    std::vector<std::pair<uint32_t, uint32_t>> dm;  // (distance, mask)
    for (auto& pr : L.pairs) {
      const uint32_t pa = (uint32_t)phys[pr.first], pb = (uint32_t)phys[pr.second];
      const uint32_t lo = std::min(pa, pb), d = std::max(pa, pb) - lo;
      size_t j = 0;
      while (j < dm.size() && dm[j].first != d) ++j;
      if (j == dm.size()) dm.push_back({d, 0u});
      dm[j].second |= 1u << lo;
    }
    const uint32_t K = (uint32_t)L.pairs.size();
    std::vector<uint64_t> w;
    w.push_back((uint64_t)dm.size() | ((uint64_t)K << 32));
    for (auto& e : dm) w.push_back((uint64_t)e.first | ((uint64_t)e.second << 32));
    if (w.size() % 2) w.push_back(0);
    const double ae = std::arg(L.pe), ao = std::arg(L.po);
    for (uint32_t k = 0; k <= K; ++k) {
      const cd t = std::polar(1.0, (double)(K - k) * ae + (double)k * ao);
      double re = t.real(), im = t.imag();
      uint64_t a, b;
      std::memcpy(&a, &re, 8); std::memcpy(&b, &im, 8);
      w.push_back(a); w.push_back(b);
    }
    return w;
  }

  static bool is_diag_kind(uint8_t k) { return k == SVO_D1 || k == SVO_D2 || k == SVO_DZZ; }
  std::vector<HOp> merge_diagonal_runs(const HPass& p, int* n_pre) {
    std::vector<HOp> outv;
    *n_pre = 0;
    bool seen_non_diag = false;
    for (size_t i = 0; i < p.ops.size();) {
      if (!is_diag_kind(p.ops[i].kind) || p.ops[i].cond_q >= 0) { outv.push_back(p.ops[i++]); seen_non_diag = true; continue; }
      size_t j = i;
      while (j < p.ops.size() && is_diag_kind(p.ops[j].kind) && p.ops[j].cond_q < 0) ++j;
      const size_t first = outv.size();
      std::vector<ZzLayer> groups;
      std::vector<HOp> singles;  // the original op of a group's first bond
      for (size_t k = i; k < j; ++k) {
        const HOp& o = p.ops[k];
        bool zz = false;
        if (o.kind == SVO_D2) {
          const cd* ph = reinterpret_cast<const cd*>(&mats[o.off]);
          zz = ph[0] == ph[3] && ph[1] == ph[2];
          if (zz) {
            size_t gi = 0;
            for (; gi < groups.size(); ++gi) {
              bool dup = false;
              for (auto& pr : groups[gi].pairs) dup = dup || (pr.first == o.qa && pr.second == o.qb) || (pr.first == o.qb && pr.second == o.qa);
              if (!dup && groups[gi].pe == ph[0] && groups[gi].po == ph[1] && groups[gi].pairs.size() < 62) break;
            }
            if (gi == groups.size()) { groups.push_back(ZzLayer()); groups[gi].pe = ph[0]; groups[gi].po = ph[1]; singles.push_back(o); }
            groups[gi].pairs.push_back({o.qa, o.qb});
          }
        }
        if (!zz) outv.push_back(o);
      }
      for (size_t gi = 0; gi < groups.size(); ++gi) {
        if (groups[gi].pairs.size() == 1) { outv.push_back(singles[gi]); continue; }
        HOp op;
        op.kind = SVO_DZZ;
        op.off = (int32_t)zz_layers.size();
        zz_layers.push_back(groups[gi]);
        outv.push_back(op);
      }
      if (!seen_non_diag) *n_pre = (int)outv.size();
      (void)first;
      i = j;
    }
    return outv;
  }

  void emit_sweep(const std::vector<int>& sel_in, std::vector<char>& in_tile) {
    std::vector<int> sel(sel_in);
    int nt = 0;
    for (int p = 0; p < nl; ++p) nt += in_tile[p];
    for (int p = 0; p < nl && nt < K; ++p) if (!in_tile[p]) { in_tile[p] = 1; ++nt; }
    int slot_of[64];
    SweepDesc sw{};
    int s = 0;
    for (int p = 0; p < nl; ++p)
      if (in_tile[p]) {
        slot_of[p] = s;
        if (s >= L) sw.pos[s - L] = (uint8_t)p;
        ++s;
      }
    for (int i = K - L; i < 8; ++i) sw.pos[i] = 31;  // unused slots
    // direct passes: a pass that commutes with everything before / after it in this sweep (no
    // shared touched qubit) and whose slots are free slots moves to the front / back and exchanges
    // its register groups with global memory itself (sv_kernels.cuh)
    // tile slots of the pass slots: real qubits first, then fillers (free slots preferred, so a
    // pass on free slots stays eligible for direct global loads / stores)
    auto slots_of = [&](const HPass& p, uint8_t (&ts)[4]) {
      bool used[32] = {};
      for (int i = 0; i < 4; ++i)
        if (p.q[i] >= 0) { ts[i] = (uint8_t)slot_of[phys[p.q[i]]]; used[ts[i]] = true; }
      for (int i = 0; i < 4; ++i) {
        if (p.q[i] >= 0) continue;
        int pick = -1;
        for (int t = K - 1; t >= L && pick < 0; --t) if (!used[t]) pick = t;
        for (int t = 0; t < L && pick < 0; ++t) if (!used[t]) pick = t;
        ts[i] = (uint8_t)pick; used[pick] = true;
      }
    };
    // only the register-resident part of a fast pass exchanges amplitudes with global memory:
    // a direct first pass must not start with diagonal ops (they run on the tile), a direct last
    // pass must not end with them
    auto eligible = [&](int i, bool as_first) {
      uint8_t ts[4];
      slots_of(passes[i], ts);
      const HPass& p = passes[i];
      if (!p.fast || p.n_slot == 0 || (as_first ? p.n_pre > 0 : p.post)) return false;
      return L >= 3 && ts[0] >= L && ts[1] >= L && ts[2] >= L && ts[3] >= L;
    };
    bool first_direct = false, last_direct = false;
    if (direct_passes) {
      for (size_t k = 0; k < sel.size() && !first_direct; ++k) {
        if (!eligible(sel[k], true)) continue;
        bool free_ = true;
        for (size_t e = 0; e < k && free_; ++e) free_ = !(passes[sel[e]].touch & passes[sel[k]].touch);
        if (!free_) continue;
        std::rotate(sel.begin(), sel.begin() + k, sel.begin() + k + 1);
        first_direct = true;
      }
      const size_t stop = (first_direct && sel.size() > 1) ? 1 : 0;
      for (size_t k = sel.size(); k-- > stop && !last_direct;) {
        if (!eligible(sel[k], false)) continue;
        bool free_ = true;
        for (size_t l = k + 1; l < sel.size() && free_; ++l) free_ = !(passes[sel[l]].touch & passes[sel[k]].touch);
        if (!free_) continue;
        std::rotate(sel.begin() + k, sel.begin() + k + 1, sel.end());
        last_direct = true;
      }
    }
    // diagonal runs of a pass: ZZ-type bonds with one phase pair that landed in the same run become
    // one fused-layer op (with n_global > 0 the planner keeps the bonds separate so that the light
    // cone of a global qubit stays narrow; they are fused here, where they meet)
    std::vector<std::vector<HOp>> mops(sel.size());
    std::vector<int> m_pre(sel.size(), 0);
    for (size_t k = 0; k < sel.size(); ++k) mops[k] = merge_diagonal_runs(passes[sel[k]], &m_pre[k]);
    size_t n_ops = 0;
    for (size_t k = 0; k < sel.size(); ++k) n_ops += mops[k].size();
    auto al16 = [](size_t x) { return (x + 15) & ~size_t(15); };
    const size_t o_pass = sizeof(BlockHdr);
    const size_t o_ops = al16(o_pass + sizeof(SvPassHdr) * sel.size());
    size_t o_par = al16(o_ops + sizeof(SvBlockOp) * n_ops);
    size_t bytes = o_par;
    for (size_t k = 0; k < sel.size(); ++k)
      for (const HOp& o : mops[k])
        bytes += o.kind == SVO_DZZ ? al16(8 * zz_params(zz_layers[o.off]).size()) : al16(8 * (size_t)op_words(o.kind));
    const size_t blk_begin = out->prog.size();
    out->prog.resize(blk_begin + bytes / 8, 0);
    uint64_t* blk = out->prog.data() + blk_begin;
    reinterpret_cast<BlockHdr*>(blk)->n_passes = (int32_t)sel.size();
    SvPassHdr* ph = reinterpret_cast<SvPassHdr*>(reinterpret_cast<char*>(blk) + o_pass);
    SvBlockOp* bo = reinterpret_cast<SvBlockOp*>(reinterpret_cast<char*>(blk) + o_ops);
    size_t oc = 0, pc = o_par;
    for (size_t k = 0; k < sel.size(); ++k) {
      const HPass& p = passes[sel[k]];
      uint8_t ts[4];
      slots_of(p, ts);
      if (k == 0 && first_direct) ph[k].flags |= kPassLoadDirect;
      if (k + 1 == sel.size() && last_direct) ph[k].flags |= kPassStoreDirect;
      ph[k].ops_q8 = (uint16_t)((o_ops + sizeof(SvBlockOp) * oc) / 8);
      ph[k].n_ops = (uint16_t)mops[k].size();
      // slot -> physical position: low slots are their own position, free slots come from the sweep
      for (int i = 0; i < 4; ++i) {
        ph[k].s[i] = ts[i];
        ph[k].pp[i] = ts[i] < L ? ts[i] : sw.pos[ts[i] - L];
      }
      sv_thread_bits(ph[k].s, K, ph[k].tb);
      for (uint32_t c = 0; c < 16; ++c) {
        uint32_t j = 0;
        for (int i = 0; i < 4; ++i) j |= ((c >> i) & 1u) << ts[i];
        ph[k].cor[c] = 16u * svz12(j);
      }
      ph[k].sig = SVS_GENERIC;
      if (p.fast) {
        ph[k].n_pre = (uint8_t)m_pre[k];
        ph[k].sig = p.n_slot == 0 ? SVS_DIAG
                  : (uint8_t)((p.slot_kind == SVO_X1 ? SVS_X1 : p.slot_kind == SVO_R1 ? SVS_R1 : SVS_U1) + (p.n_slot - 1));
      }
      for (const HOp& o : mops[k]) {
        SvBlockOp& d = bo[oc++];
        d.kind = o.kind;
        d.flags = 0;
        d.qa = d.qb = 0;
        if (o.kind == SVO_DZZ) {
          const std::vector<uint64_t> w = zz_params(zz_layers[o.off]);
          ph[k].needs_index = 1;
          d.off = (uint16_t)(pc / 8);
          std::memcpy(reinterpret_cast<char*>(blk) + pc, w.data(), 8 * w.size());
          pc += al16(8 * w.size());
          continue;
        }
        if (o.kind == SVO_U1 || o.kind == SVO_X || o.kind == SVO_R1 || o.kind == SVO_X1) {
          d.qa = (uint8_t)p.slot_of(o.target);
        } else if (o.kind == SVO_U2 || o.kind == SVO_SWAP) {
          d.qa = (uint8_t)p.slot_of(o.qa);
          d.qb = (uint8_t)p.slot_of(o.qb);
        } else {
          d.qa = (uint8_t)phys[o.qa];
          d.qb = (uint8_t)(o.qb >= 0 ? phys[o.qb] : 0);
          ph[k].needs_index = 1;
        }
        if (o.cond_q >= 0) {
          d.flags |= SVF_COND | (o.cond_val ? SVF_COND_VAL : 0);
          d.cond_bit = (uint8_t)phys[o.cond_q];
          ph[k].needs_index = 1;
        }
        const int w = op_words(o.kind);
        if (w) {
          d.off = (uint16_t)(pc / 8);
          std::memcpy(reinterpret_cast<char*>(blk) + pc, &mats[o.off], 8 * (size_t)w);
          pc += al16(8 * (size_t)w);
        }
      }
    }
    sw.blk_q16 = (uint32_t)(blk_begin / 2);
    sw.blk_len_q16 = (uint32_t)(bytes / 16);
    if (first_direct) sw.blk_len_q16 |= kSvFirstDirect << 16;
    out->sweeps.push_back(sw);
    // a qubit only diagonal ops / control tests have seen is still |0>: amplitudes with a 1 on its
    // bit are zero, so tiles (or whole shards) with such an outside bit set can be skipped
    for (int i : sel)
      for (int k = 0; k < 4; ++k) if (passes[i].q[k] >= 0) touched |= 1ull << passes[i].q[k];
    uint32_t um = 0;
    for (int q = 0; q < n; ++q) if (!((touched >> q) & 1)) um |= 1u << phys[q];
    out->sweep_untouched.push_back(um);
    out->n_passes += (int64_t)sel.size();
  }

  // greedy: fills sweeps until no remaining pass is placeable; returns passes left
  int pack_local(std::vector<char>& done, int remaining) {
    const int np = (int)passes.size();
    const int seg_first = (int)out->sweeps.size();
    std::vector<char> in_tile(std::max(nl, 1));
    std::vector<int> sel;
    while (remaining > 0) {
      std::fill(in_tile.begin(), in_tile.end(), 0);
      int nt = 0;
      for (int p = 0; p < L; ++p) { in_tile[p] = 1; ++nt; }
      uint64_t blocked = 0;
      const uint64_t all = n >= 64 ? ~0ull : ((1ull << n) - 1);
      sel.clear();
      int bytes = (int)sizeof(BlockHdr) + 32;  // header + alignment slack
      for (int i = 0; i < np && blocked != all; ++i) {
        if (done[i]) continue;
        const HPass& p = passes[i];
        if (p.touch & blocked) { blocked |= p.touch; continue; }
        bool ok = true;
        int need = 0;
        int newpos[4];
        for (int q : p.q) {
          if (q < 0) continue;
          const int pp = phys[q];
          if (pp >= nl) { ok = false; break; }
          if (!in_tile[pp]) newpos[need++] = pp;
        }
        int add = pass_bytes(p);
        if (!ok || nt + need > K || bytes + add > kBlockBytes) { blocked |= p.touch; continue; }
        for (int k = 0; k < need; ++k) { in_tile[newpos[k]] = 1; ++nt; }
        bytes += add;
        done[i] = 1;
        --remaining;
        sel.push_back(i);
      }
      if (sel.empty()) break;
      emit_sweep(sel, in_tile);
    }
    const int cnt = (int)out->sweeps.size() - seg_first;
    if (cnt > 0) out->segs.push_back(SvxSegment{SVSEG_SWEEPS, seg_first, cnt, 0});
    return remaining;
  }

  // evicts the g local qubits whose next slot use lies furthest ahead (or that appear in `keep`
  // least urgently) to the top local positions, then exchanges them with the global positions
  void exchange(const std::vector<int64_t>& next_use) {
    std::vector<int> cand;
    for (int q = 0; q < n; ++q) if (phys[q] < nl) cand.push_back(q);
    std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) {
      if (next_use[a] != next_use[b]) return next_use[a] > next_use[b];
      return phys[a] > phys[b];
    });
    cand.resize(g);
    std::vector<char> is_victim(n, 0);
    for (int q : cand) is_victim[q] = 1;
    // victims outside the top region swap with non-victims inside it
    std::vector<HPass> saved;
    saved.swap(passes);
    std::vector<std::pair<int, int>> swaps;
    int top = nl - g;
    for (int q : cand) {
      if (phys[q] >= nl - g) continue;
      while (is_victim[logical_at[top]]) ++top;
      swaps.push_back({q, logical_at[top]});
      ++top;
    }
    for (size_t i = 0; i < swaps.size();) {
      // as many swaps per sweep as the tile holds
      std::vector<char> in_tile(nl, 0);
      int nt = 0;
      for (int p = 0; p < L; ++p) { in_tile[p] = 1; ++nt; }
      std::vector<int> sel;
      passes.clear();
      size_t j = i;
      for (; j < swaps.size(); ++j) {
        const int pa = phys[swaps[j].first], pb = phys[swaps[j].second];
        const int need = (!in_tile[pa]) + (!in_tile[pb]);
        if (nt + need > K) break;
        if (!in_tile[pa]) { in_tile[pa] = 1; ++nt; }
        if (!in_tile[pb]) { in_tile[pb] = 1; ++nt; }
        HPass p;
        p.q[0] = swaps[j].first; p.q[1] = swaps[j].second;
        p.fast = false;
        p.touch = (1ull << p.q[0]) | (1ull << p.q[1]);
        HOp op; op.kind = SVO_SWAP; op.qa = (int8_t)p.q[0]; op.qb = (int8_t)p.q[1];
        p.ops.push_back(op);
        sel.push_back((int)passes.size());
        passes.push_back(p);
      }
      const int first = (int)out->sweeps.size();
      emit_sweep(sel, in_tile);
      out->segs.push_back(SvxSegment{SVSEG_SWEEPS, first, 1, 0});
      for (size_t k = i; k < j; ++k) {
        const int a = swaps[k].first, b = swaps[k].second;
        std::swap(phys[a], phys[b]);
        logical_at[phys[a]] = a;
        logical_at[phys[b]] = b;
      }
      i = j;
    }
    passes.swap(saved);
    out->segs.push_back(SvxSegment{SVSEG_EXCHANGE, 0, 0, 0});
    out->n_exchanges++;
    for (int i = 0; i < g; ++i) {
      const int pl = nl - g + i, pg = nl + i;
      const int a = logical_at[pl], b = logical_at[pg];
      phys[a] = pg; phys[b] = pl;
      logical_at[pg] = a; logical_at[pl] = b;
    }
  }

  bool pack_stage() {
    const int np = (int)passes.size();
    std::vector<char> done(np, 0);
    int remaining = np;
    int guard = 0;
    while (remaining > 0) {
      const int before = remaining;
      remaining = pack_local(done, remaining);
      if (remaining == 0) break;
      if (g == 0 || (remaining == before && ++guard > 4 * n + 8)) return false;
      if (remaining != before) guard = 0;
      std::vector<int64_t> next_use(n, INT64_MAX);
      for (int i = np - 1; i >= 0; --i) {
        if (done[i]) continue;
        for (int k = 0; k < 4; ++k) if (passes[i].q[k] >= 0) next_use[passes[i].q[k]] = i;
      }
      exchange(next_use);
    }
    return true;
  }
};

}  // namespace

void lower_svx_circuit(const bwq_batch& b, int c, const SvxOptions& opt, SvxProgram* out) {
  *out = SvxProgram();
  const int nq = b.n_qubits[c];
  const int64_t g0 = b.op_offsets[c], g1 = b.op_offsets[c + 1];
  out->n_gates = g1 - g0;
  if (nq < 0 || nq > 64) { out->status = BWQ_CIRC_BAD_QUBIT; return; }
  std::vector<int> bit_of(nq, -1);
  {
    std::vector<char> used(nq, 0);
    for (int64_t g = g0; g < g1; ++g) {
      const bwq_op& op = b.ops[g];
      const bool two = gate_is_2q(op.opcode);
      if (op.q0 >= nq || (two && (op.q1 >= nq || op.q1 == op.q0))) { out->status = BWQ_CIRC_BAD_QUBIT; return; }
      if (op.opcode == BWQ_G_RESET) { out->status = BWQ_CIRC_BAD_OP; return; }
      used[op.q0] = 1;
      if (two) used[op.q1] = 1;
    }
    for (int q = 0; q < nq; ++q) if (used[q]) { bit_of[q] = (int)out->active.size(); out->active.push_back(q); }
  }
  const int gl = std::max(0, opt.n_global);
  // a register pass owns four tile slots, and an exchange (the top g local bits leave) must be able
  // to keep the up to four qubits of a pass local: n_local >= g + 4
  while ((int)out->active.size() < 2 * gl + 4) out->active.push_back(-1);
  const int n = out->n_bits = (int)out->active.size();
  if (n > kMaxSvQubits + 1) { out->status = BWQ_CIRC_TOO_WIDE; return; }
  Planner P;
  P.out = out;
  P.direct_passes = opt.direct != 0 && std::getenv("BWQ_SVX_NO_DIRECT") == nullptr;  // env: kernel experiments
  // whole-layer fusion makes every qubit of the layer wait for the layer: right on one GPU (one
  // table lookup per amplitude and layer), wrong when amplitudes are sharded -- a global qubit would
  // need an exchange per Trotter layer instead of one per light cone; there the bonds stay separate
  // and are fused pass by pass at emission (merge_diagonal_runs)
  P.fuse_layers = std::getenv("BWQ_SVX_NO_FUSE") == nullptr && gl == 0;
  if (const char* ce = std::getenv("BWQ_SVX_ZZ_CHUNK")) {  // experiment: chunked layer ops when sharded
    const int c = std::atoi(ce);
    if (c >= 2 && std::getenv("BWQ_SVX_NO_FUSE") == nullptr) { P.fuse_layers = true; P.zz_chunk = gl > 0 ? std::min(c, 62) : 62; }
  }
  P.structured = std::getenv("BWQ_SVX_NO_STRUCT") == nullptr;
  P.n = n; P.g = gl; P.nl = n - gl;
  P.K = std::min(std::min(std::max(opt.tile_bits, 4), kSvTileBitsMax), P.nl);
  P.L = std::max(0, P.K - kSvFreeSlots);
  out->n_local = P.nl; out->n_global = gl; out->tile_bits = P.K;
  P.phys.resize(n); P.logical_at.resize(n);
  for (int q = 0; q < n; ++q) P.phys[q] = P.logical_at[q] = q;

  // ---- stage 0: the circuit
  P.begin_stage();
  for (int64_t g = g0; g < g1; ++g) {
    const bwq_op& op = b.ops[g];
    const int npar = gate_num_params(op.opcode);
    if (npar && (int64_t)op.param_idx + npar > b.n_params) { out->status = BWQ_CIRC_BAD_OP; return; }
    const double* par = npar ? b.params + op.param_idx : nullptr;
    double u[32];
    if (!gate_unitary(op.opcode, par, u)) { out->status = BWQ_CIRC_BAD_OP; return; }
    cd m[16];
    if (gate_is_2q(op.opcode)) {
      for (int i = 0; i < 16; ++i) m[i] = cd(u[2 * i], u[2 * i + 1]);
      P.gate2(bit_of[op.q0], bit_of[op.q1], m);
    } else {
      for (int i = 0; i < 4; ++i) m[i] = cd(u[2 * i], u[2 * i + 1]);
      P.gate1(bit_of[op.q0], m);
    }
  }
  P.end_gates();
  if (!P.pack_stage()) { out->status = BWQ_CIRC_BAD_OP; return; }

  // ---- observables: Pauli terms -> qubit-wise commuting families
  const int64_t o0 = b.obs_offsets[c], o1 = b.obs_offsets[c + 1];
  struct Term { std::vector<uint8_t> p; double coeff; int obs; bool ztype; };
  std::vector<Term> terms;
  const uint64_t valid = nq >= 64 ? ~0ull : ((1ull << nq) - 1);
  for (int64_t o = o0; o < o1; ++o)
    for (int64_t t = b.term_offsets[o]; t < b.term_offsets[o + 1]; ++t) {
      const uint64_t x = b.term_x[t], z = b.term_z[t];
      if ((x | z) & ~valid) { out->status = BWQ_CIRC_BAD_QUBIT; return; }
      Term T;
      T.p.assign(n, 0);
      T.coeff = b.term_coeff[t];
      T.obs = (int)(o - o0);
      T.ztype = true;
      for (int q = 0; q < nq; ++q) {
        const int xb = (x >> q) & 1, zb = (z >> q) & 1;
        if (!xb && !zb) continue;
        if (bit_of[q] < 0) { if (xb) T.coeff = 0.0; continue; }  // idle qubit stays |0>
        T.p[bit_of[q]] = xb ? (zb ? 2 : 1) : 3;
        if (xb) T.ztype = false;
      }
      if (T.coeff == 0.0) { std::fill(T.p.begin(), T.p.end(), 0); T.ztype = true; }
      terms.push_back(std::move(T));
    }
  std::vector<std::vector<uint8_t>> fam_basis;
  std::vector<std::vector<int>> fam_terms;
  auto assign = [&](int ti) {
    const Term& T = terms[ti];
    for (size_t f = 0; f < fam_basis.size(); ++f) {
      bool ok = true;
      for (int q = 0; q < n && ok; ++q) ok = !T.p[q] || !fam_basis[f][q] || T.p[q] == fam_basis[f][q];
      if (!ok) continue;
      for (int q = 0; q < n; ++q) if (T.p[q]) fam_basis[f][q] = T.p[q];
      fam_terms[f].push_back(ti);
      return;
    }
    fam_basis.push_back(T.p);
    fam_terms.push_back({ti});
  };
  for (size_t i = 0; i < terms.size(); ++i) if (terms[i].ztype) assign((int)i);
  for (size_t i = 0; i < terms.size(); ++i) if (!terms[i].ztype) assign((int)i);

  // current measurement basis of every qubit (3 = Z = computational)
  std::vector<uint8_t> cur(n, 3);
  const double r = std::sqrt(0.5);
  const cd H[4] = {r, r, r, -r};
  const cd HSdg[4] = {r, cd(0, -r), r, cd(0, r)};  // H * Sdg
  auto rot = [&](uint8_t basis, cd* m) {
    if (basis == 1) std::memcpy(m, H, sizeof H);
    else if (basis == 2) std::memcpy(m, HSdg, sizeof HSdg);
    else { m[0] = 1; m[1] = 0; m[2] = 0; m[3] = 1; }
  };
  for (size_t f = 0; f < fam_basis.size(); ++f) {
    P.begin_stage();
    bool any = false;
    for (int q = 0; q < n; ++q) {
      const uint8_t want = fam_basis[f][q];
      if (!want || want == cur[q]) continue;
      cd a[4], bm[4], u[4];
      rot(want, a);
      rot(cur[q], bm);  // undo: bm^dagger
      for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j)
        u[i * 2 + j] = a[i * 2] * std::conj(bm[j * 2]) + a[i * 2 + 1] * std::conj(bm[j * 2 + 1]);
      P.gate1(q, u);
      cur[q] = want;
      any = true;
    }
    if (any) {
      P.end_gates();
      if (!P.pack_stage()) { out->status = BWQ_CIRC_BAD_OP; return; }
    }
    const int first = (int)out->zt_mask.size();
    for (int ti : fam_terms[f]) {
      const Term& T = terms[ti];
      uint32_t m = 0;
      for (int q = 0; q < n; ++q) if (T.p[q]) m |= 1u << P.phys[q];
      out->zt_mask.push_back(m);
      out->zt_coeff.push_back(T.coeff);
      out->zt_obs.push_back(T.obs);
    }
    out->segs.push_back(SvxSegment{SVSEG_EXPVAL, first, (int)fam_terms[f].size(), 0});
  }
}

}  // namespace bwq
