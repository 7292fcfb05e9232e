// Internal types shared by the lowering stage (host C++) and the CUDA kernels.
//
// State layout ("Pauli-basis density matrix"): a circuit with n active qubits owns 4^n doubles
//   r[idx] = Tr(rho P_idx),  idx = sum_q digit_q * 4^pos(q),  digit: 0=I 1=X 2=Y 3=Z
// i.e. the density matrix expanded in the Pauli basis.  rho is Hermitian, so r is REAL and every
// gate / noise channel is a REAL transfer matrix; this is the same information as Aer's
// column-stacked complex vec(rho) in half the bytes, and Tr(rho P) is a lookup.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/bwq.h"

namespace bwq {

// Register-pass op kinds.  A pass owns two tile slots (a, b); every thread holds the 16 elements
// v[da + 4*db] of one (a, b) group in registers and runs the pass's op list on them.
enum OpKind : int32_t {
  K_DENSE1_A = 0,  // 4x4 real matrix on digit a          (16 doubles)
  K_DENSE1_B = 1,  // 4x4 real matrix on digit b          (16 doubles)
  K_CX_AB = 2,     // CX, control a, target b             (signed permutation, no data)
  K_CX_BA = 3,     // CX, control b, target a
  K_RELAX2 = 4,    // diag + affine noise, (q0,q1)=(a,b)  (25 doubles)
  K_RELAX2_SW = 5, // same with (q0,q1)=(b,a)
  K_DENSE2 = 6,    // 16x16 real matrix, (q0,q1)=(a,b)    (256 doubles)
  K_DENSE2_SW = 7, // same with (q0,q1)=(b,a)
  K_AFF1_A = 8,    // trace-preserving 1-qubit channel on digit a: rows 1..3 of its 4x4 (12 doubles)
  K_AFF1_B = 9,
  K_ROTZ_A = 10,   // rz / phase on digit a: (cos, sin) (2 doubles, padded to 4)
  K_ROTZ_B = 11,
  K_COUNT = 12
};

struct DevOp {      // 16 B
  int32_t kind;
  int32_t src;      // 0: batch matrix buffer, 1: ctx noise table
  int64_t off;      // offset in doubles
};
struct PassDesc {   // 16 B
  int32_t op_begin, op_end;
  uint8_t sa, sb;   // tile slots of digits a and b
  uint8_t pad[6];
};
struct SweepDesc {  // 16 B
  int32_t pass_begin, pass_end;
  uint8_t pos[8];   // digit positions resident in the tile, ascending; first n_tile valid
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
  int set(const bwq_noise_table* t, char* err, size_t errlen);
  const NoiseEntry* find(uint16_t opcode, int q0, int q1) const;
  bool empty() const { return entries.empty(); }
};

// ---- lowered program of one circuit (indices relative to the circuit) -------------------
struct CircuitProgram {
  int32_t status = 0;
  int32_t n_digits = 0;                 // active qubits (>= 2, padded with idle digits)
  std::vector<int32_t> active;          // physical qubit of each digit (-1 = padding)
  std::vector<SweepDesc> sweeps;
  std::vector<PassDesc> passes;
  std::vector<DevOp> ops;
  std::vector<double> mats;
  std::vector<int64_t> term_index;      // per Pauli term: element index or -1
  std::vector<double> term_coeff;
  int64_t n_gates = 0;
  bool needs_dense = false;             // uses K_DENSE1_* / K_DENSE2_* (selects the FULL kernel)
};

struct LowerOptions {
  int tile_qubits = 6;
  int low_qubits = 2;
};

// Lowers circuit c of the batch.  Never throws; sets status on per-circuit failure.
void lower_dm_circuit(const NoiseTable& noise, const bwq_batch& b, int c, const LowerOptions& o,
                      CircuitProgram* out);

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

// gate library (host)
bool gate_is_2q(uint16_t opcode);
int gate_num_params(uint16_t opcode);
// 2x2 / 4x4 complex unitary, row-major, interleaved re/im.  Returns false for unknown opcode.
bool gate_unitary(uint16_t opcode, const double* params, double* u);
void ptm_from_unitary1(const double* u, double* r /*16*/);
void ptm_from_unitary2(const double* u, double* r /*256*/);

}  // namespace bwq
