/*
 * bwq.h -- C ABI of the B200-native exact expectation-value engine ("blackwater-quantum").
 *
 * Drop-in boundary.  The reference (qiskit-community/ml-qem) reaches its
 * simulator through the Qiskit Estimator primitive:
 *     blackwater/data/utils.py:418-431   create_estimator_meas_data  (ideal + noisy AerEstimator)
 *     blackwater/data/utils.py:434-444   create_meas_data_from_estimators
 *     blackwater/library/learning/estimator.py:279-285  run(self, circuits=, observables=, ...)
 * and the arithmetic behind it is qiskit-aer's C++ controller (pybind, not in tree).  The entry
 * points below are what an FFI for that path binds: plain pointers and sizes, no torch / qiskit
 * types.  ml_qem_b200/engine.py is the ctypes binding; INTEGRATION.md shows the reference-side
 * stub.
 *
 * Conventions
 *   - qubit 0 is the least-significant bit of a basis-state index (Qiskit little-endian);
 *   - a Pauli term is (x_mask, z_mask, coeff): bit q of x_mask set for X or Y on qubit q, bit q
 *     of z_mask set for Z or Y; the value of a term is coeff * Tr(rho P) (real part returned);
 *   - the caller owns every host buffer for the duration of the call; the library owns all device
 *     memory inside bwq_ctx; every function returns 0 (BWQ_OK) or a negative error code and the
 *     message is available from bwq_last_error(); a ctx is not re-entrant (one per GPU);
 *   - per-circuit failures (unsupported op, too many qubits) are reported in out_status[] and do
 *     not void the rest of the batch.
 */
#ifndef BWQ_H_
#define BWQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BWQ_VERSION 100 /* 0.1.0 */

enum {
  BWQ_OK = 0,
  BWQ_ERR_ARG = -1,      /* malformed input */
  BWQ_ERR_CUDA = -2,     /* CUDA runtime error (message has the CUDA string) */
  BWQ_ERR_NOMEM = -3,    /* batch does not fit the device budget even one circuit at a time */
  BWQ_ERR_UNSUPPORTED = -4,
  BWQ_ERR_NO_DEVICE = -5 /* no CUDA device: there is NO CPU fallback */
};

/* per-circuit status */
enum { BWQ_CIRC_OK = 0, BWQ_CIRC_BAD_OP = 1, BWQ_CIRC_TOO_WIDE = 2, BWQ_CIRC_BAD_QUBIT = 3 };

/* Gate opcodes of the flat gate stream (Qiskit standard gates; names as in
 * blackwater/data/utils.py:19-49 plus the backend basis id/rz/sx/x/cx/reset). */
enum {
  BWQ_G_ID = 0, BWQ_G_X, BWQ_G_Y, BWQ_G_Z, BWQ_G_H, BWQ_G_S, BWQ_G_SDG, BWQ_G_T, BWQ_G_TDG,
  BWQ_G_SX, BWQ_G_SXDG, BWQ_G_RX, BWQ_G_RY, BWQ_G_RZ, BWQ_G_P, BWQ_G_U2, BWQ_G_U3,
  BWQ_G_RESET,                                   /* 1-qubit, non-unitary                  */
  BWQ_G_CX = 32, BWQ_G_CY, BWQ_G_CZ, BWQ_G_CH, BWQ_G_CRX, BWQ_G_CRY, BWQ_G_CRZ, BWQ_G_CP,
  BWQ_G_CU3, BWQ_G_SWAP, BWQ_G_ISWAP, BWQ_G_RZZ, BWQ_G_RXX, BWQ_G_RYY, BWQ_G_RZX, BWQ_G_ECR,
  BWQ_G_UNITARY1 = 64, /* params: 8 doubles, row-major 2x2 complex (re,im interleaved)   */
  BWQ_G_UNITARY2 = 65, /* params: 32 doubles, row-major 4x4 complex, index i_q0 + 2 i_q1 */
  BWQ_G_COUNT = 66
};

/* One gate of the stream.  q1 is ignored for 1-qubit gates.  param_idx indexes batch.params
 * (first of up to 3 consecutive angles; 8 / 32 doubles for UNITARY1 / UNITARY2). */
typedef struct {
  uint16_t opcode;
  uint8_t q0;
  uint8_t q1;
  uint32_t param_idx;
} bwq_op;

/* A batch of independent circuits with their observables, flat structure-of-arrays. */
typedef struct {
  int32_t n_circuits;
  const int32_t* n_qubits;    /* [n_circuits] register width (<= 64)                         */
  const int64_t* op_offsets;  /* [n_circuits+1] circuit c owns ops[op_offsets[c] .. [c+1])   */
  const bwq_op* ops;
  const double* params;
  int64_t n_params;
  const int64_t* obs_offsets; /* [n_circuits+1] circuit c owns observables [..)              */
  const int64_t* term_offsets;/* [n_observables+1] observable o owns Pauli terms [..)        */
  const uint64_t* term_x;
  const uint64_t* term_z;
  const double* term_coeff;   /* real coefficients                                           */
} bwq_batch;

/* Variants of every circuit of a batch, generated INSIDE the library (K0) from the base gate
 * stream: ZNE local folding of the 2-qubit gates and Pauli twirling of the cx gates -- what the
 * reference builds as new Python circuits per variant (docs/tutorials/zne_parallel.py:168-189,
 * docs/tutorials/derek_files/phase_diagram.ipynb:776,960).  Base circuit c yields
 * n_variants = max(1, n_folds) * max(1, n_twirls) circuits, variant index fold * n_twirls + twirl. */
typedef struct {
  int32_t n_folds;          /* 0 = no folding (factor 1)                                       */
  const int32_t* folds;     /* [n_folds] odd noise factors, e.g. {1, 3, 5}                     */
  int32_t n_twirls;         /* 0 = no twirling; else twirled instances per (circuit, fold)     */
  uint64_t seed;            /* the Pauli pairs are a pure function of (seed, circuit, twirl, cx index) */
} bwq_variants;

/* Noise table: one error per (opcode, ordered physical qubits), the lookup Aer's
 * NoiseModel.from_backend uses (reached from blackwater/data/utils.py:427).  Errors are given in
 * the real Pauli-transfer-matrix form R[i][j] = Tr(P_i E(P_j)) / 2^k, P in {I,X,Y,Z}, index
 * digit_q0 + 4*digit_q1.
 *   kind BWQ_NOISE_DENSE1 : 16 doubles, row-major 4x4
 *   kind BWQ_NOISE_DENSE2 : 256 doubles, row-major 16x16
 *   kind BWQ_NOISE_RELAX2 : 25 doubles d[16], ca[4], cb[4], cab -- the closed form of
 *        (thermal relaxation (x) thermal relaxation) o (any Pauli-diagonal channel):
 *        out[i] = d[i] in[i];  out[Z,b] += ca[b] in[I,b];  out[a,Z] += cb[a] in[a,I];
 *        out[Z,Z] += cab in[I,I]
 *   q1 = 255 marks a 1-qubit entry; q0 = 255 marks an all-qubit default for the opcode. */
enum { BWQ_NOISE_DENSE1 = 1, BWQ_NOISE_DENSE2 = 2, BWQ_NOISE_RELAX2 = 3 };
typedef struct {
  int32_t n_entries;
  const uint16_t* opcode;   /* [n_entries] */
  const uint8_t* q0;        /* [n_entries] */
  const uint8_t* q1;        /* [n_entries] */
  const uint8_t* kind;      /* [n_entries] */
  const int64_t* data_off;  /* [n_entries] offset into data (doubles) */
  const double* data;
  int64_t n_data;
} bwq_noise_table;

/* Tunables (0 = library default). */
typedef struct {
  int32_t tile_qubits;      /* Pauli digits per shared-memory tile, 2..7 (default 6)          */
  int32_t low_qubits;       /* lowest digits always resident in a tile (coalescing), default 2 */
  int64_t max_state_bytes;  /* device budget for resident states (default: 80% of free)        */
  int32_t chunk_circuits;   /* circuits simulated together (default: as many as fit)          */
  int32_t host_threads;     /* lowering threads (default: hardware concurrency)               */
  int32_t sv_tile_bits;     /* amplitudes per statevector tile = 2^this, 2..12 (default 12)    */
  int32_t flags;            /* BWQ_OPT_* bits (0 = defaults)                                  */
} bwq_options;
/* planner switches for kernel experiments: keep the first / last pass of a sweep on the staged path */
#define BWQ_OPT_NO_DIRECT_LOAD 1
#define BWQ_OPT_NO_DIRECT_STORE 2
#define BWQ_OPT_NO_PIPELINE 4   /* bwq_dm_run: lower the whole batch before the first launch */
#define BWQ_OPT_FORCE_PIPELINE 8 /* bwq_dm_run: pipeline segments even when the sweeps look too short to matter */
#define BWQ_OPT_PERSIST 32       /* density matrix: after the first sweep use the persistent double-buffered
                                   dm_sweep_tma_persistent_kernel (two resident CTAs per SM, producer warp) instead of one
                                   CTA per tile; measured slower (8 compute warps per SM), kept for A/B */
#define BWQ_OPT_TMA_DIRECT_STORE 64 /* TMA layout: the last pass of a sweep stores its register groups to global memory
                                   with 16-byte stores instead of scatter + TMA store; measured slower (0.635 vs 0.676), A/B */
#define BWQ_OPT_NO_ONCHIP 128   /* keep circuits of <= 5 active qubits on the lowering + tile-sweep path instead of
                                   dm_onchip_kernel (one warp interprets the raw gate stream of a circuit); A/B and parity tests */
#define BWQ_OPT_NO_TMA 16       /* density matrix: keep circuits wider than the tile on dm_sweep_kernel (LDG/STG tile
                                   movement) instead of dm_sweep_tma_kernel (TMA tensor-map tiles); A/B and parity tests */

/* Counters of the last *_run call (for the roofline: bytes = sweeps x 16 B x 4^n). */
typedef struct {
  int64_t n_sweep_launches;    /* kernel launches of the density-matrix sweep kernel     */
  int64_t n_state_sweeps;      /* (circuit, sweep) pairs executed = P of SURVEY 8(d)     */
  int64_t n_passes;            /* register passes (2-qubit groups) executed              */
  int64_t n_gates;             /* gates consumed from the stream                         */
  int64_t state_bytes_swept;   /* sum over state sweeps of 2 * 8 B * 4^n_active (the first
                                  sweep of a circuit only writes: 1 * 8 B * 4^n)           */
  int64_t n_other_launches;    /* init / expectation / statevector launches              */
  double lower_ms, h2d_ms, kernel_ms, d2h_ms; /* host wall / device-event times          */
  double sweep_kernel_ms;      /* CUDA-event time of the sweep launches only             */
  int64_t h2d_bytes, d2h_bytes;/* program upload / value download of the last run          */
  int64_t sv_state_bytes_swept;/* statevector sweeps: sum of 2 * 16 B * 2^n (+ 16 B * 2^n per
                                  expectation pass) of the last bwq_sv_* call                  */
  int64_t n_tma_sweep_launches;/* of n_sweep_launches: launches of dm_sweep_tma_kernel       */
  int64_t n_onchip_circuits;   /* (circuit, variant) pairs evolved by dm_onchip_kernel (no lowering,
                                  no sweeps: n_sweep_launches = 0 then)                        */
  double host_pre_ms;          /* pipelined bwq_dm_run / bwq_meas_data_run: wall time from entry to the first
                                  enqueue (checks, budget, lowering of the first segment)      */
  double call_wall_ms;         /* wall time of the whole call; call_wall_ms - host_pre_ms - kernel_ms is
                                  what the host adds after the last device event               */
} bwq_stats;

typedef struct bwq_ctx bwq_ctx;

int bwq_version(void);
/* Creates the engine on CUDA device `device`.  Fails with BWQ_ERR_NO_DEVICE when no GPU. */
int bwq_create(int device, bwq_ctx** out);
int bwq_destroy(bwq_ctx* ctx);
const char* bwq_last_error(const bwq_ctx* ctx); /* ctx may be NULL: last create() error */
int bwq_set_options(bwq_ctx* ctx, const bwq_options* opt);
/* Installs (copies) the noise table; NULL or n_entries == 0 => noise-free. */
int bwq_set_noise_table(bwq_ctx* ctx, const bwq_noise_table* table);

/* Noisy values: density-matrix evolution of every circuit under the installed noise table
 * (Aer Estimator, method=density_matrix, approximation=True, shots=None).
 * out_vals[n_observables] (order of term_offsets), out_status[n_circuits].  Batches of >= 256
 * circuits with enough sweep work (about 5 ms by a gate-count estimate) are cut into segments and
 * pipelined: the host lowers segment k+1 while the GPU sweeps segment k (BWQ_OPT_NO_PIPELINE /
 * BWQ_OPT_FORCE_PIPELINE override the choice). */
int bwq_dm_run(bwq_ctx* ctx, const bwq_batch* batch, double* out_vals, int32_t* out_status);
/* Split form of bwq_dm_run: prepare lowers the batch on the host and uploads the program (it
 * stays resident in HBM inside the ctx); execute runs the kernels and returns the values, and may
 * be repeated.  bwq_dm_run == prepare + execute. */
int bwq_dm_prepare(bwq_ctx* ctx, const bwq_batch* batch, int32_t* out_status);
int bwq_dm_execute(bwq_ctx* ctx, double* out_vals);
int bwq_dm_execute_device_out(bwq_ctx* ctx, double* d_out_vals);
/* Ideal values: statevector evolution, noise table ignored (qiskit.primitives.Estimator,
 * shots=None; docs/tutorials/h13_ising_data_gen_tomo.ipynb:811). */
int bwq_sv_run(bwq_ctx* ctx, const bwq_batch* batch, double* out_vals, int32_t* out_status);
int bwq_sv_prepare(bwq_ctx* ctx, const bwq_batch* batch, int32_t* out_status);
int bwq_sv_execute(bwq_ctx* ctx, double* out_vals);
/* Same as bwq_dm_run / bwq_sv_run but out_vals is a DEVICE pointer (e.g. a torch tensor's
 * data_ptr()) -- zero-copy label hand-off; returns after the work is enqueued and synchronised. */
int bwq_dm_run_device_out(bwq_ctx* ctx, const bwq_batch* batch, double* d_out_vals, int32_t* out_status);
/* Ideal AND noisy values of every circuit in one call: what create_estimator_meas_data
 * (blackwater/data/utils.py:418-431) returns per circuit, for the whole batch.  The statevector
 * side runs on a companion context (own stream and buffers) concurrently with the density-matrix
 * pipeline, so its lowering and launches hide behind the sweeps.  Equivalent to bwq_sv_run into
 * out_ideal / status_ideal and bwq_dm_run into out_noisy / status_noisy; bwq_get_stats reports the
 * density-matrix side. */
int bwq_meas_data_run(bwq_ctx* ctx, const bwq_batch* batch, double* out_ideal, double* out_noisy,
                      int32_t* status_ideal, int32_t* status_noisy);
/* (ideal, noisy) values of every VARIANT of every circuit: the noisy side (density matrix) runs all
 * n_circuits * n_variants circuits, the ideal side (statevector) the base circuits only -- folds
 * and twirls leave the ideal circuit unchanged.  out_noisy[(observable of circuit c, variant v)] is
 * laid out circuit-major, then variant, then the circuit's observables; out_ideal as bwq_sv_run.
 * status_noisy[n_circuits] (per base circuit: first failing variant), status_ideal[n_circuits]. */
int bwq_meas_data_run_variants(bwq_ctx* ctx, const bwq_batch* batch, const bwq_variants* variants, double* out_ideal,
                               double* out_noisy, int32_t* status_ideal, int32_t* status_noisy);
/* Noisy side alone: out_vals / out_status cover the n_circuits * n_variants expanded circuits
 * (circuit-major, then variant; out_status[n_circuits * n_variants]). */
int bwq_dm_run_variants(bwq_ctx* ctx, const bwq_batch* batch, const bwq_variants* variants, double* out_vals, int32_t* out_status);
/* Host-only view of the expansion (no GPU): sizes[0..3] = {n circuits, n ops, n params, variants per
 * circuit}; a second call with buffers copies op_offsets[n+1], ops[n_ops], params[n_params] out. */
int bwq_expand_variants(const bwq_batch* batch, const bwq_variants* variants, int64_t sizes[4], int64_t* op_offsets,
                        bwq_op* ops, double* params);
int bwq_get_stats(const bwq_ctx* ctx, bwq_stats* out);
int bwq_sync(bwq_ctx* ctx);

/* ---- host-only introspection of the lowering stage (no GPU needed; used by CPU tests) ---- */
typedef struct bwq_program bwq_program;
/* Lowers batch circuit `circuit` to its sweep program.  tile_qubits/low_qubits as in options. */
int bwq_lower_dm(const bwq_noise_table* table, const bwq_batch* batch, int32_t circuit,
                 int32_t tile_qubits, int32_t low_qubits, bwq_program** out);
/* Same with planner flags: bits 8..15 = ZNE noise factor (the fold-aware lowering bwq_*_variants uses for
 * cx-only circuits: gates lowered once, cx ops repeated); bit 1 = direct last-pass stores in the TMA layout; bit 0 = emit the TMA tile layout (what bwq_dm_run does by default for
 * circuits wider than the tile; see ml_qem_b200/csrc/program.h). */
int bwq_lower_dm_ex(const bwq_noise_table* table, const bwq_batch* batch, int32_t circuit,
                    int32_t tile_qubits, int32_t low_qubits, int32_t flags, bwq_program** out);
void bwq_program_free(bwq_program* p);
/* Sizes: [0]=n_active, [1]=n_sweeps, [2]=n_passes, [3]=n_prog (8-byte words), [4]=needs_dense,
 * [5]=status, [6]=n_terms, [7]=n_gates */
int bwq_program_sizes(const bwq_program* p, int64_t sizes[8]);
/* Copies the program out.  active_qubits[n_active] (physical qubit of digit d, -1 = padding);
 * sweeps: per sweep 10 int32 {block offset (16-byte units into prog), pos[0..7], block length
 * (16-byte units)}; prog[n_prog]: the sweep blocks (layout: ml_qem_b200/csrc/program.h);
 * term_index[n_terms] (element index, -1 = term vanishes); term_coeff. */
int bwq_program_read(const bwq_program* p, int32_t* active_qubits, int32_t* sweeps, uint64_t* prog,
                     int64_t* term_index, double* term_coeff);

/* ---- wide and amplitude-sharded statevector (ideal labels beyond 12 qubits) ----
 * Replaces qiskit.primitives.Estimator / Statevector.evolve for wide registers
 * (docs/tutorials/h13_ising_data_gen_tomo.ipynb:811 runs it at 16 qubits).  bwq_sv_run uses this
 * path internally for circuits wider than 12 active qubits; the entry points below expose it for
 * ONE circuit whose 2^n amplitudes are sharded over 2^n_global_bits GPUs (one process per GPU):
 * the program is a list of segments -- local tile sweeps, EXCHANGE (the caller swaps the top
 * n_global_bits local index bits with the rank bits: one all-to-all of 2^g contiguous blocks,
 * NCCL over NVLink), and EXPVAL (signed |amplitude|^2 sums accumulated into n_observables
 * partial values the caller all-reduces).  Device buffers are caller-owned. */
typedef struct bwq_svx_program bwq_svx_program;
/* Host-only planner.  tile_bits 0 = default (12). */
int bwq_svx_lower(const bwq_batch* batch, int32_t circuit, int32_t tile_bits, int32_t n_global_bits,
                  bwq_svx_program** out);
void bwq_svx_free(bwq_svx_program* p);
/* sizes: [0]=status [1]=n_bits [2]=n_local [3]=n_global [4]=tile_bits [5]=n_sweeps [6]=n_prog
 * (8-byte words) [7]=n_segments [8]=n_zterms [9]=n_passes [10]=n_exchanges [11]=n_observables */
int bwq_svx_sizes(const bwq_svx_program* p, int64_t sizes[12]);
/* active[n_bits]; sweeps: 10 int32 each {block offset, pos[0..7], block length} (16-byte units);
 * segs: 4 int32 each {kind (0 sweeps, 1 exchange, 2 expval), first, count, 0}. */
int bwq_svx_read(const bwq_svx_program* p, int32_t* active, int32_t* sweeps, uint64_t* prog, int32_t* segs,
                 uint32_t* zt_mask, double* zt_coeff, int32_t* zt_obs);
/* Algorithmic bytes this rank's kernels move for the whole program: 2 x 16 B per amplitude of
 * every live tile of every sweep (the first sweep only writes; tiles that are provably zero are
 * skipped) + 16 B per amplitude per expectation pass.  For the roofline. */
int64_t bwq_svx_bytes(const bwq_svx_program* p, int32_t rank);
/* Uploads the program to ctx's device (kept inside the handle). */
int bwq_svx_upload(bwq_ctx* ctx, bwq_svx_program* p);
/* Runs segment `segment` on this rank's shard d_state (2^n_local complex128, device memory).
 * SWEEPS segments update the shard in place (the first sweep of the program initialises
 * |0...0>); EXPVAL segments add this shard's contribution into d_obs[n_observables] (device);
 * EXCHANGE segments are the caller's job and are rejected here.  Work is enqueued on `stream`
 * (a cudaStream_t; 0 = the ctx's own non-blocking stream, (void*)1 = cudaStreamLegacy, i.e. the
 * default stream torch uses) and not synchronised. */
int bwq_svx_run_segment(bwq_ctx* ctx, const bwq_svx_program* p, int32_t segment, double* d_state,
                        int32_t rank, double* d_obs, void* stream);

/* A SWEEPS segment that is followed by an EXCHANGE, with the exchange FUSED into the store of its
 * last sweep: sv_sweep_kernel writes the swept tiles straight into the peers' NEW shards
 * (peer_dst[w] = rank w's new shard as mapped into this process, peer_dst[rank] the local one) --
 * the separate read + write of the whole shard by bwq_svx_exchange_* disappears and the NVLink
 * traffic overlaps the sweep's arithmetic.  d_state (the old shard) is left as the input of the
 * last sweep.  The caller skips the EXCHANGE segment and places the cross-rank barrier after this
 * call (all tiles have landed before anyone sweeps its new shard); world = 2^n_global <= 16. */
int bwq_svx_run_segment_push(bwq_ctx* ctx, const bwq_svx_program* p, int32_t segment, double* d_state, int32_t rank,
                             const uint64_t* peer_dst, int32_t world, void* stream);

/* EXCHANGE segment over NVLink peer memory (the engine's own replacement for the NCCL
 * all_to_all): the top log2(world) local index bits swap with the rank bits, so block b of the
 * new shard d_dst is block `rank` of rank b's old shard.  peer_src[w] = device pointer to rank
 * w's OLD shard as mapped into this process (symmetric memory / CUDA IPC; peer_src[rank] is the
 * local one).  The kernel pulls all 2^g blocks with P2P loads on `stream` (same convention as
 * bwq_svx_run_segment) and does not synchronise; the caller places a cross-rank barrier before
 * it (every peer has finished writing its old shard) and before the old shard is overwritten. */
int bwq_svx_exchange_pull(bwq_ctx* ctx, double* d_dst, const uint64_t* peer_src, int32_t world, int32_t rank,
                          int64_t n_local_amps, void* stream);
/* Same exchange with P2P stores: block b of the local OLD shard d_src is written into block `rank`
 * of rank b's NEW shard peer_dst[b].  The caller places the cross-rank barrier AFTER it (all
 * blocks have landed before anyone sweeps its new shard). */
int bwq_svx_exchange_push(bwq_ctx* ctx, const double* d_src, const uint64_t* peer_dst, int32_t world, int32_t rank,
                          int64_t n_local_amps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BWQ_H_ */
