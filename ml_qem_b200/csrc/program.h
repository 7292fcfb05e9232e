// Internal types shared by the lowering stage (host C++) and the CUDA kernels.
//
// State layout ("Pauli-basis density matrix"): a circuit with n active qubits owns 4^n doubles
//   r[idx] = Tr(rho P_idx),  idx = sum_q digit_q * 4^pos(q),  digit: 0=I 1=X 2=Y 3=Z
// i.e. the density matrix expanded in the Pauli basis.  rho is Hermitian, so r is REAL and every
// gate / noise channel is a REAL transfer matrix; this is the same information as Aer's
// column-stacked complex vec(rho) in half the bytes, and Tr(rho P) is a lookup.
#pragma once
#include <cstdint>
#include <functional>
#include <vector>

#include "../../include/bwq.h"

namespace bwq {

// Register-pass macro-ops.  A pass owns two tile slots (a, b); every thread holds the 16 elements
// v[da + 4*db] of one (a, b) group in registers and runs the pass's macro-op list on them.  One
// macro-op = optional 1-qubit map on a, optional 1-qubit map on b, then an optional 2-qubit op.
enum PreKind : uint8_t {
  P_NONE = 0,
  P_ROT = 1,    // rz / phase as three shears: (tan(theta'/2), sin(theta'), sign)      4 doubles
  P_AFF = 2,    // trace-preserving 1-qubit channel: rows 1..3 of its 4x4 transfer matrix 12 doubles
  P_DENSE = 3   // general 4x4 (not trace preserving)                                   16 doubles
};
enum TwoKind : uint8_t {
  Q_NONE = 0,
  Q_CXN_AB = 1,   // CX (control a) fused with its diag+affine noise                   28 doubles
  Q_CXN_BA = 2,   // CX (control b) fused with its noise
  Q_CX_AB = 3,    // bare CX: signed permutation, no data
  Q_CX_BA = 4,
  Q_RELAX = 5,    // diag + affine noise alone, (q0,q1) = (a,b)                        28 doubles
  Q_RELAX_SW = 6, // (q0,q1) = (b,a)
  Q_DENSE = 7,    // 16x16 real matrix, (q0,q1) = (a,b)                               256 doubles
  Q_DENSE_SW = 8
};

// Host-side macro-op (lowering scratch): parameter offsets point into the noise table or, with
// kLocalMat set, into the circuit's own matrix scratch.
struct MacroOp {
  uint8_t pre_a, pre_b, twoq, pad;
  uint32_t off_a, off_b, off_2;
};
constexpr uint32_t kLocalMat = 0x80000000u;

// ---- device program: one contiguous, self-contained block per sweep -----------------------
// A CTA copies its sweep's block into shared memory (cp.async) while the tile streams in, so
// every descriptor and parameter the register passes need is one LDS away.
//   block := BlockHdr | PassHdr[n_passes] | BlockOp[...] | parameters (16-byte aligned)
struct BlockHdr {   // 16 B
  int32_t n_passes;
  int32_t ext_q16;  // TMA layout: offset (16-byte units) of the per-pass corner tables, 64 B per pass,
                    // followed by one more 64 B table: the 16 corner offsets IN THE STATE (elements) of the
                    // last pass when it stores its register groups straight to global memory (kPassStoreDirect)
  uint8_t slot_pos[8];  // TMA layout: digit position of tile slot s (box order), s < 6
};
struct PassHdr {    // 16 B
  uint16_t ops_q16; // offset of the pass's first BlockOp, in 16-byte units from the block start
  uint16_t n_ops;
  uint8_t sa, sb;   // tile slots of digits a and b
  uint8_t flags;    // kPassLoadDirect (first pass; mirrored in SweepDesc::pos[7]) | kPassStoreDirect (last pass)
  uint8_t sig;      // TMA layout: PassSig -- straight-line body the kernel runs for this op list
  uint8_t row;      // TMA layout: row of the thread-base table (slot pair index hi (hi - 1) / 2 + lo)
  uint8_t pad[3];
  uint32_t gofs;    // TMA layout: byte offset (XOR) of a thread's second register group
};
// Direct passes: the first pass of a sweep reads its register groups from global memory and the
// last one writes them back, instead of staging the tile through shared memory.  Requested when
// neither target slot is one of the two lowest slots (full 128-byte lines per warp access).
constexpr uint8_t kPassLoadDirect = 1, kPassStoreDirect = 2;
constexpr uint8_t kFirstDirect = 0x80;  // SweepDesc::pos[7] = kFirstDirect | sa | sb << 3 of the first pass
struct BlockOp {    // 16 B
  uint8_t pre_a, pre_b, twoq, pad0;
  uint16_t off_a, off_b, off_2;  // parameter offsets in 8-byte units from the block start
  uint16_t pad1[3];
};
struct SweepDesc {  // 16 B
  uint32_t blk_q16;     // block offset in 16-byte units from the program buffer base
  uint32_t blk_len_q16; // low 16 bits: block length in 16-byte units; high 16 bits (density
                        // matrix only): digit positions no pass has touched up to and including
                        // this sweep -- a tile with X or Y on such an outside digit is all zero
  uint8_t pos[8];       // digit positions resident in the tile, ascending; first n_tile (<= 7) valid;
                        // pos[7] (density matrix): first-pass descriptor, see kFirstDirect
};
constexpr int kBlockBytes = 8192;  // shared-memory program buffer per CTA

// ---- TMA tile layout (dm_sweep_tma_kernel; circuits wider than the 6-digit tile) ---------------
// The tile is the box of a rank-5 tensor map: dim 0 = the 16 contiguous elements of digits 0 and 1
// (128 bytes), dims 1..4 = the four other resident digits IN AN ORDER THE LOWERING CHOOSES (the
// "box order"; slot 2 + k of a pass header = box dim 1 + k).  Element with tile-local index j
// (2 bits per slot) lives at 8 * tswz(j) in shared memory -- CU_TENSOR_MAP_SWIZZLE_128B: byte
// address bits 4..6 ^= bits 7..9, i.e. (D0hi, D1lo, D1hi) ^= (D2lo, D2hi, D3lo).
// Lowering output: SweepDesc::pos[0..5] ascending positions, pos[6] = box order (2 bits per box
// dim: index into the ascending list of the four upper positions), pos[7] = kTmaSweep; the batch
// merge replaces pos[6..7] by the 16-bit index of the sweep's tensor map.
constexpr uint8_t kTmaSweep = 0x40;
constexpr int kTmaThreads = 128;
#ifdef __CUDACC__
#define BWQ_HD __host__ __device__
#else
#define BWQ_HD
#endif
BWQ_HD constexpr uint32_t tswz(uint32_t j) { return j ^ (((j >> 4) & 7u) << 1); }

// Register-pass work split of the TMA kernel: 128 threads x 32 elements.  A thread owns the 16
// (da, db) corners of TWO register groups.  Slot 0 not a target: the groups differ in the low bit
// of digit 0, i.e. one 16-byte shared-memory access brings corner i of both (beta = index bit 0).
// Slot 0 a target: they differ in the highest free index bit.  tbit[k] = tile index bit driven by
// thread-id bit k: a quarter warp (lane bits 0..2) must hit eight distinct 16-byte chunks, chunk
// bits = index bits (1,2,3) ^ (4,5,6); lane bit k-1 drives bit k, or bit k+3 when bit k belongs
// to a target digit -- impossible only for the slot pairs {0,2}, {1,2}, {1,3} (2-way conflict),
// which the lowering avoids through the box order.
struct TmaPassLayout {
  int tbit[7];
  int beta;
};
inline TmaPassLayout tma_pass_layout(int sa, int sb) {
  TmaPassLayout L{};
  bool is_target[12], used[12];
  for (int b = 0; b < 12; ++b) { is_target[b] = (b >> 1) == sa || (b >> 1) == sb; used[b] = is_target[b]; }
  const bool slot0 = sa == 0 || sb == 0;
  if (!slot0) { L.beta = 0; used[0] = true; }
  for (int k = 1; k <= 3; ++k) {
    int pick = -1;
    if (!used[k]) pick = k;
    else if (!used[k + 3]) pick = k + 3;
    L.tbit[k - 1] = pick;  // -1: no conflict-free choice, filled below
    if (pick >= 0) used[pick] = true;
  }
  if (slot0) {  // beta = highest free bit
    for (int b = 11; b >= 0; --b) if (!used[b]) { L.beta = b; used[b] = true; break; }
  }
  int next = 0;
  auto take = [&]() { while (used[next]) ++next; used[next] = true; return next; };
  for (int k = 0; k < 3; ++k) if (L.tbit[k] < 0) L.tbit[k] = take();
  for (int k = 3; k < 7; ++k) L.tbit[k] = take();
  return L;
}
inline int tma_pair_row(int sa, int sb) {
  const int lo = sa < sb ? sa : sb, hi = sa < sb ? sb : sa;
  return hi * (hi - 1) / 2 + lo;
}
constexpr int kTmaPairs = 15;
// thread-base table: row (lo, hi) x thread -> swizzled byte offset of the thread's first element
inline void fill_tma_a_table(uint32_t* t) {
  for (int hi = 1; hi < 6; ++hi)
    for (int lo = 0; lo < hi; ++lo) {
      const TmaPassLayout L = tma_pass_layout(lo, hi);
      for (int tid = 0; tid < kTmaThreads; ++tid) {
        uint32_t j = 0;
        for (int k = 0; k < 7; ++k) j |= uint32_t((tid >> k) & 1) << L.tbit[k];
        t[tma_pair_row(lo, hi) * kTmaThreads + tid] = 8u * tswz(j);
      }
    }
}

// Pass signatures with a straight-line body in the TMA kernel (no op dispatch inside the pass, so
// no register reconciliation at dispatch joins).  Macro-op (pre_a, pre_b, twoq) sequences:
enum PassSig : uint8_t {
  SIG_GENERIC = 0,
  SIG_AAC = 1,       // (AFF, AFF, CXN_AB)
  SIG_AAC_AA = 2,    // (AFF, AFF, CXN_AB) (AFF, AFF, -)
  SIG_AAC_RC = 3,    // (AFF, AFF, CXN_AB) (-, ROT, CXN_AB)         TFIM bond, first Trotter step
  SIG_C_RC = 4,      // (-, -, CXN_AB) (-, ROT, CXN_AB)             TFIM bond
  SIG_RAC = 5,       // (ROT, AFF, CXN_AB)
  SIG_0AC = 6,       // (-, AFF, CXN_AB)
  SIG_AAC_A0 = 7,    // (AFF, AFF, CXN_AB) (AFF, -, -)
  SIG_AAC_0A = 8,    // (AFF, AFF, CXN_AB) (-, AFF, -)
  SIG_COUNT
};

constexpr int kMaxTileQubits = 7;
constexpr int kMaxDmQubits = 16;   // 4^16 doubles = 34 GB
constexpr int kMaxSvQubits = 30;

// ---- host-side noise table -------------------------------------------------------------
struct NoiseEntry {
  uint8_t kind;
  int64_t off;
};
struct NoiseTable {
  std::vector<double> data;
  // key = opcode << 16 | q0 << 8 | q1   (q1 = 255 for 1-qubit, q0 = 255 for all-qubit default)
  std::vector<std::pair<uint32_t, NoiseEntry>> entries;  // sorted by key
  // 1-qubit gates (most of a transpiled circuit): resolved entry per (opcode < 32, qubit < 64),
  // all-qubit default included; -1 = no error, -2 = not tabulated (fall back to the search)
  std::vector<int32_t> one_q;
  // noisy transfer matrix (error x gate, 16 doubles) of every parameter-free 1-qubit gate that has
  // an error: index (opcode * 64 + qubit) * 16, valid where fixed_ok is set
  std::vector<double> fixed_ptm;
  std::vector<char> fixed_ok;
  void build_fixed();
  int set(const bwq_noise_table* t, char* err, size_t errlen);
  const NoiseEntry* find(uint16_t opcode, int q0, int q1) const;
  bool empty() const { return entries.empty(); }
};

// ---- lowered program of one circuit (indices relative to the circuit) -------------------
struct CircuitProgram {
  int32_t status = 0;
  int32_t n_digits = 0;                 // active qubits (>= 2, padded with idle digits)
  std::vector<int32_t> active;          // physical qubit of each digit (-1 = padding)
  std::vector<SweepDesc> sweeps;        // blk_q16 relative to this circuit's prog
  std::vector<uint64_t> prog;           // sweep blocks, 8-byte words (16-byte aligned blocks)
  int64_t n_passes = 0;
  std::vector<double> mats;             // lowering scratch (not uploaded)
  std::vector<int64_t> term_index;      // per Pauli term: element index or -1
  std::vector<double> term_coeff;
  int64_t n_gates = 0;
  bool needs_dense = false;             // uses P_DENSE / Q_DENSE* (selects the FULL kernel)
  bool tma = false;                     // sweeps are in the TMA tile layout
};

struct LowerOptions {
  int tile_qubits = 6;
  int low_qubits = 2;
  int direct = kPassLoadDirect | kPassStoreDirect;  // which direct passes the planner may request
  bool tma = false;  // emit the TMA tile layout for circuits wider than a 6-digit tile (tile_qubits 6, low_qubits 2)
  bool tma_direct_store = false;  // TMA layout: last pass stores to global memory itself (measured slower than the TMA store)
};

// Lowers circuit c of the batch.  Never throws; sets status on per-circuit failure.
void lower_dm_circuit(const NoiseTable& noise, const bwq_batch& b, int c, const LowerOptions& o,
                      CircuitProgram* out);

// Programs of circuit c at the ZNE noise factors folds[0..n): the gates are lowered once.  False when
// the circuit has 2-qubit gates other than cx (lower the expanded stream of bwq_variants instead).
bool lower_dm_circuit_folds(const NoiseTable& noise, const bwq_batch& b, int c, const LowerOptions& o, const int32_t* folds,
                            int n_folds, CircuitProgram* outs);

// ---- statevector lowering --------------------------------------------------------------
enum SvKind : int32_t { SV_U1 = 0, SV_CX = 1, SV_U2 = 2 };
struct SvOp {       // 16 B
  int32_t kind;
  uint8_t q0, q1;   // bit positions (compacted)
  uint8_t pad[2];
  int64_t off;      // into mats (doubles): U1 -> 8, U2 -> 32 (row-major complex, re/im)
};
struct SvProgram {
  int32_t status = 0;
  int32_t n_bits = 0;
  std::vector<int32_t> active;
  std::vector<SvOp> ops;
  std::vector<double> mats;
  // per term: x mask, z mask (compacted bits), y count; coeff = 0 marks a vanishing term
  std::vector<uint32_t> term_x, term_z;
  std::vector<int32_t> term_ny;
  std::vector<double> term_coeff;
  int64_t n_gates = 0;
};
void lower_sv_circuit(const bwq_batch& b, int c, SvProgram* out);

// ---- variant generation (variants.cpp): folds / twirls of a base batch on the flat gate stream
struct ExpandedBatch {
  int32_t n_variants = 1;               // variants per base circuit
  std::vector<int32_t> n_qubits, status; // status: per BASE circuit (BWQ_CIRC_BAD_OP: a gate has no inverse rule)
  std::vector<int64_t> op_offsets, obs_offsets, term_offsets;
  std::vector<bwq_op> ops;
  std::vector<double> params, term_coeff;
  std::vector<uint64_t> term_x, term_z;
  bwq_batch view{};                     // points into the vectors above
};
// runner(n, body): executes body(i) for i in [0, n) (in parallel); null = std::threads spawned per call
using ParallelRunner = std::function<void(int, const std::function<void(int)>&)>;
int expand_variants(const bwq_batch& base, const bwq_variants& v, ExpandedBatch* out, int threads = 1, const ParallelRunner* runner = nullptr);
uint32_t twirl_draw(uint64_t seed, uint64_t circuit, uint64_t twirl, uint64_t cx_index);

// gate library (host)
bool gate_is_2q(uint16_t opcode);
int gate_num_params(uint16_t opcode);
// 2x2 / 4x4 complex unitary, row-major, interleaved re/im.  Returns false for unknown opcode.
bool gate_unitary(uint16_t opcode, const double* params, double* u);
void ptm_from_unitary1(const double* u, double* r /*16*/);
void ptm_from_unitary2(const double* u, double* r /*256*/);

// ============================================================================================
// Statevector SWEEP program (wide circuits, batched or amplitude-sharded across GPUs).
//
// State: 2^n_local complex128 per GPU; with n_global > 0 the rank supplies the top index bits
// (physical bit n_local + i = bit i of the rank).  A sweep stages a tile of 2^K amplitudes in
// shared memory (the lowest L = max(0, K-8) physical bits always resident, plus up to 8 others)
// and runs register passes on slot pairs, like the density-matrix path.  Diagonal operations
// read the bits of the full physical index and therefore never need their qubits resident (or
// even local); controlled operations only need their TARGET resident (the control is a bit
// test).  Non-diagonal work on a global qubit is preceded by an EXCHANGE segment: physical bits
// [n_local-g, n_local) <-> [n_local, n_local+g), i.e. one all-to-all of 2^g contiguous blocks.
// Expectation values are always evaluated on Z-type strings: X/Y terms are grouped into
// qubit-wise commuting families and the planner appends the basis rotation as ordinary gates.
// ============================================================================================
enum SvOpKind : uint8_t {
  SVO_U1 = 1,    // 2x2 on the target slot                      8 doubles
  SVO_X = 2,     // Pauli X on the target slot (cx = conditional X)
  SVO_U2 = 3,    // 4x4 on (a,b), local index i_a + 2 i_b       32 doubles
  SVO_SWAP = 4,  // exchange slots a and b
  SVO_D1 = 5,    // phases by physical bit qa                    4 doubles (p0, p1)
  SVO_D2 = 6,    // phases by physical bits (qa,qb), b_qa+2b_qb  8 doubles
  // structured 1-qubit unitaries (global phase dropped -- it never reaches an expectation value):
  SVO_R1 = 7,    // real 2x2 (ry, h, ...): {m00, m01, m10, m11}                          4 doubles
  SVO_X1 = 8,    // real diagonal, imaginary off-diagonal (rx, sx, x): {d0, d1, o01, o10},
                 // u = [[d0, i o01], [i o10, d1]]                                       4 doubles
  // a whole layer of K commuting exp(-i t ZZ)-type bonds with ONE phase pair (p_even, p_odd): the
  // phase of an amplitude is table[w], w = number of bonds with odd parity =
  // sum_j popc((x ^ (x >> d_j)) & M_j) over the distinct bond distances d_j.
  // params: {u32 n_d, u32 K}, n_d x {u32 d, u32 M}, pad to 16 B, (K + 1) complex table
  SVO_DZZ = 9
};
enum : uint8_t { SVF_COND = 2, SVF_COND_VAL = 4 };
struct SvBlockOp {   // 8 B
  uint8_t kind, flags;
  uint8_t qa, qb;    // U1/X/R1/X1: qa = pass slot (0..3) of the target; U2/SWAP: pass slots of the (first, second)
                     // operand, matrix index i_first + 2 i_second; D1/D2: physical bit positions
  uint16_t off;      // parameter offset, 8-byte units from the block start
  uint8_t cond_bit;  // SVF_COND: physical bit tested against SVF_COND_VAL
  uint8_t pad;
};
// A register pass owns FOUR tile slots: every thread keeps the 16 amplitudes v[ia + 2 ib + 4 ic + 8 id]
// of one group (256 threads x 16 = one 2^12 tile) and runs the pass's op list on them.
// Fast passes (sig != SVS_GENERIC) have the shape  [diagonal ops] [one structured 1-qubit op of the
// SAME kind on pass slots 0 .. n-1] [diagonal ops]  and run as straight-line bodies.
enum SvSig : uint8_t {
  SVS_GENERIC = 0,
  SVS_X1 = 1,        // + (n - 1), n = 1..4 slot ops of kind SVO_X1
  SVS_R1 = 5,        // SVO_R1
  SVS_U1 = 9,        // SVO_U1 (unconditional)
  SVS_DIAG = 13,     // diagonal ops only
};
struct SvPassHdr {   // 96 B
  uint16_t ops_q8;   // first SvBlockOp, 8-byte units from the block start
  uint16_t n_ops;
  uint8_t s[4];      // tile slots of pass slots 0..3 (distinct)
  uint8_t sig;       // SvSig
  uint8_t n_pre;     // fast passes: diagonal ops before the slot ops (the rest follow them)
  uint8_t needs_index;  // some op reads the physical index (diagonal / conditional)
  uint8_t flags;     // kPassLoadDirect (first pass of the sweep) | kPassStoreDirect (last pass)
  uint8_t pp[4];     // physical bit positions of pass slots 0..3
  uint8_t tb[8];     // tile index bit driven by thread-id bit k (quarter warps hit 8 distinct 16-byte chunks)
  uint8_t pad[8];
  uint32_t cor[16];  // swizzled byte offsets of the 16 corners
};
static_assert(sizeof(SvPassHdr) == 96, "SvPassHdr layout");
// shared-memory swizzle of the statevector tile (16-byte amplitudes): the low three index bits are
// XOR-ed with bits 3..5, 6..8 and 9..11, so a lane bit may drive any of four index bits per chunk
// bit -- with four pass slots there is always a free one: conflict free for every slot choice.
BWQ_HD constexpr uint32_t svz12(uint32_t j) { return j ^ ((j >> 3) & 7u) ^ ((j >> 6) & 7u) ^ ((j >> 9) & 7u); }
// thread-id bit -> tile index bit for a pass on tile slots sl[0..3] (K tile bits)
inline void sv_thread_bits(const uint8_t sl[4], int K, uint8_t tb[8]) {
  bool used[16] = {};
  for (int i = 0; i < 4; ++i) used[sl[i]] = true;
  int n = 0;
  for (int k = 0; k < 3 && n < K - 4; ++k)
    for (int b = k; b < K; b += 3)
      if (!used[b]) { tb[n++] = (uint8_t)b; used[b] = true; break; }
  for (int b = 0; b < K && n < K - 4; ++b)
    if (!used[b]) { tb[n++] = (uint8_t)b; used[b] = true; }
  for (; n < 8; ++n) tb[n] = 31;  // tiles smaller than 2^12 (tests): unused thread bits
}
// Statevector sweeps mirror the first-pass descriptor in the high half of SweepDesc::blk_len_q16:
// kSvFirstDirect | sa | sb << 4 (the block length keeps the low 16 bits).  A pass is eligible when
// both slots are free slots (>= the always-resident low bits): lanes then walk the contiguous low
// bits, i.e. whole 256-byte runs.
constexpr uint32_t kSvFirstDirect = 0x8000u;
constexpr int kSvTileBitsDefault = 11;  // 32 KiB tiles, 128-thread CTAs, four per SM: 6 % faster than 2^12 tiles (two 256-thread CTAs per SM)
constexpr int kSvTileBitsMax = 12;
constexpr int kSvFreeSlots = 8;       // SweepDesc::pos holds the positions of slots L..K-1
constexpr int kSvSmallBits = 12;      // <= this: one CTA per circuit, state in shared memory

enum : int32_t { SVSEG_SWEEPS = 0, SVSEG_EXCHANGE = 1, SVSEG_EXPVAL = 2 };
struct SvxSegment {
  int32_t kind;
  int32_t first, count;  // SWEEPS: sweep range; EXPVAL: z-term range
  int32_t pad;
};
struct SvxProgram {
  int32_t status = 0;
  int32_t n_bits = 0;      // active qubits (padded so that n_local >= 2 + n_global)
  int32_t n_local = 0, n_global = 0, tile_bits = 0;
  std::vector<int32_t> active;     // physical (register) qubit of each logical bit, -1 = padding
  std::vector<SweepDesc> sweeps;
  std::vector<uint32_t> sweep_untouched;  // per sweep: physical bits no non-diagonal op has touched yet
  std::vector<uint64_t> prog;
  std::vector<SvxSegment> segs;
  std::vector<uint32_t> zt_mask;   // Z-type terms, physical masks under the mapping at that point
  std::vector<double> zt_coeff;    // coefficient (0 = vanishing term)
  std::vector<int32_t> zt_obs;     // observable index within the circuit
  int64_t n_gates = 0, n_passes = 0, n_exchanges = 0;
};
struct SvxOptions {
  int tile_bits = kSvTileBitsDefault;
  int n_global = 0;
  int direct = 1;  // let the planner request direct first / last passes
};
void lower_svx_circuit(const bwq_batch& b, int c, const SvxOptions& o, SvxProgram* out);

}  // namespace bwq
