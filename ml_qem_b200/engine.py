"""ctypes binding of the C ABI (include/bwq.h) and the flat batch encoder.

This is the only place Python talks to the CUDA library.  There is no CPU fallback: if
``libbwq.so`` is missing or no GPU is present, ``Engine()`` raises.
"""
import ctypes as C
import os
import threading

import numpy as np

from . import circuit as circuit_mod
from . import observable as observable_mod
from .gateset import NUM_PARAMS, OPCODES

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libbwq.so")
_lib = None


class EngineError(RuntimeError):
    pass


class _Op(C.Structure):
    _fields_ = [("opcode", C.c_uint16), ("q0", C.c_uint8), ("q1", C.c_uint8), ("param_idx", C.c_uint32)]


OP_DTYPE = np.dtype([("opcode", "<u2"), ("q0", "u1"), ("q1", "u1"), ("param_idx", "<u4")])


class _Batch(C.Structure):
    _fields_ = [
        ("n_circuits", C.c_int32), ("n_qubits", C.c_void_p), ("op_offsets", C.c_void_p), ("ops", C.c_void_p),
        ("params", C.c_void_p), ("n_params", C.c_int64), ("obs_offsets", C.c_void_p),
        ("term_offsets", C.c_void_p), ("term_x", C.c_void_p), ("term_z", C.c_void_p), ("term_coeff", C.c_void_p),
    ]


class _NoiseTable(C.Structure):
    _fields_ = [
        ("n_entries", C.c_int32), ("opcode", C.c_void_p), ("q0", C.c_void_p), ("q1", C.c_void_p),
        ("kind", C.c_void_p), ("data_off", C.c_void_p), ("data", C.c_void_p), ("n_data", C.c_int64),
    ]


class _Variants(C.Structure):
    _fields_ = [("n_folds", C.c_int32), ("folds", C.c_void_p), ("n_twirls", C.c_int32), ("seed", C.c_uint64)]


class Variants:
    """ZNE folds / Pauli twirls generated inside the library from the base gate stream (bwq_variants):
    base circuit c -> len(folds) * max(1, twirls) circuits, variant index fold * twirls + twirl."""

    def __init__(self, folds=(1,), twirls=0, seed=0):
        self.folds = tuple(int(f) for f in folds) or (1,)
        if any(f < 1 or f % 2 == 0 for f in self.folds):
            raise ValueError("noise factors of local folding must be odd positive integers")
        self.twirls = int(twirls)
        self.seed = int(seed)
        self._folds_arr = np.asarray(self.folds, dtype=np.int32)

    @property
    def n_variants(self):
        return len(self.folds) * max(1, self.twirls)

    def c_struct(self):
        return _Variants(len(self.folds), _ptr(self._folds_arr), self.twirls, self.seed)


class _Options(C.Structure):
    _fields_ = [("tile_qubits", C.c_int32), ("low_qubits", C.c_int32), ("max_state_bytes", C.c_int64),
                ("chunk_circuits", C.c_int32), ("host_threads", C.c_int32), ("sv_tile_bits", C.c_int32),
                ("flags", C.c_int32)]


class _Stats(C.Structure):
    _fields_ = [("n_sweep_launches", C.c_int64), ("n_state_sweeps", C.c_int64), ("n_passes", C.c_int64),
                ("n_gates", C.c_int64), ("state_bytes_swept", C.c_int64), ("n_other_launches", C.c_int64),
                ("lower_ms", C.c_double), ("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
                ("sweep_kernel_ms", C.c_double), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("sv_state_bytes_swept", C.c_int64), ("n_tma_sweep_launches", C.c_int64), ("n_onchip_circuits", C.c_int64), ("host_pre_ms", C.c_double),
                ("call_wall_ms", C.c_double)]


EXPORTS = [
    "bwq_version", "bwq_create", "bwq_destroy", "bwq_last_error", "bwq_set_options", "bwq_set_noise_table",
    "bwq_dm_run", "bwq_sv_run", "bwq_meas_data_run", "bwq_meas_data_run_variants", "bwq_dm_run_variants", "bwq_expand_variants", "bwq_dm_run_device_out", "bwq_dm_prepare", "bwq_dm_execute", "bwq_dm_execute_device_out", "bwq_sv_prepare", "bwq_sv_execute", "bwq_get_stats", "bwq_sync", "bwq_lower_dm", "bwq_lower_dm_ex",
    "bwq_program_free", "bwq_program_sizes", "bwq_program_read",
    "bwq_svx_lower", "bwq_svx_free", "bwq_svx_sizes", "bwq_svx_read", "bwq_svx_upload", "bwq_svx_run_segment",
    "bwq_svx_bytes", "bwq_svx_exchange_pull", "bwq_svx_exchange_push", "bwq_svx_run_segment_push",
]


def load_library(path=None):
    """Loads libbwq.so (built in-tree by ``python -m ml_qem_b200.build``); raises if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("BWQ_LIB") or _LIB_PATH  # BWQ_LIB: alternative build (kernel experiments)
    if not os.path.exists(p):
        raise EngineError(f"{p} not found: build it with `python -m ml_qem_b200.build` (needs nvcc); "
                          "this engine has no CPU fallback")
    lib = C.CDLL(p)
    lib.bwq_version.restype = C.c_int
    lib.bwq_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.bwq_destroy.argtypes = [C.c_void_p]
    lib.bwq_last_error.argtypes = [C.c_void_p]
    lib.bwq_last_error.restype = C.c_char_p
    lib.bwq_set_options.argtypes = [C.c_void_p, C.POINTER(_Options)]
    lib.bwq_set_noise_table.argtypes = [C.c_void_p, C.POINTER(_NoiseTable)]
    for f in (lib.bwq_dm_run, lib.bwq_sv_run, lib.bwq_dm_run_device_out):
        f.argtypes = [C.c_void_p, C.POINTER(_Batch), C.c_void_p, C.c_void_p]
    lib.bwq_meas_data_run.argtypes = [C.c_void_p, C.POINTER(_Batch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bwq_meas_data_run_variants.argtypes = [C.c_void_p, C.POINTER(_Batch), C.POINTER(_Variants), C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p]
    lib.bwq_dm_run_variants.argtypes = [C.c_void_p, C.POINTER(_Batch), C.POINTER(_Variants), C.c_void_p, C.c_void_p]
    lib.bwq_expand_variants.argtypes = [C.POINTER(_Batch), C.POINTER(_Variants), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bwq_dm_prepare.argtypes = [C.c_void_p, C.POINTER(_Batch), C.c_void_p]
    lib.bwq_dm_execute.argtypes = [C.c_void_p, C.c_void_p]
    lib.bwq_sv_prepare.argtypes = [C.c_void_p, C.POINTER(_Batch), C.c_void_p]
    lib.bwq_sv_execute.argtypes = [C.c_void_p, C.c_void_p]
    lib.bwq_dm_execute_device_out.argtypes = [C.c_void_p, C.c_void_p]
    lib.bwq_get_stats.argtypes = [C.c_void_p, C.POINTER(_Stats)]
    lib.bwq_sync.argtypes = [C.c_void_p]
    lib.bwq_lower_dm.argtypes = [C.POINTER(_NoiseTable), C.POINTER(_Batch), C.c_int32, C.c_int32, C.c_int32,
                                 C.POINTER(C.c_void_p)]
    lib.bwq_lower_dm_ex.argtypes = [C.POINTER(_NoiseTable), C.POINTER(_Batch), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.POINTER(C.c_void_p)]
    lib.bwq_program_free.argtypes = [C.c_void_p]
    lib.bwq_program_free.restype = None
    lib.bwq_program_sizes.argtypes = [C.c_void_p, C.c_void_p]
    lib.bwq_program_read.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    lib.bwq_svx_lower.argtypes = [C.POINTER(_Batch), C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    lib.bwq_svx_free.argtypes = [C.c_void_p]
    lib.bwq_svx_free.restype = None
    lib.bwq_svx_sizes.argtypes = [C.c_void_p, C.c_void_p]
    lib.bwq_svx_read.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    lib.bwq_svx_upload.argtypes = [C.c_void_p, C.c_void_p]
    lib.bwq_svx_bytes.argtypes = [C.c_void_p, C.c_int32]
    lib.bwq_svx_bytes.restype = C.c_int64
    lib.bwq_svx_run_segment.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.bwq_svx_run_segment_push.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    for f in (lib.bwq_svx_exchange_pull, lib.bwq_svx_exchange_push):
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]
    if path is None:
        _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class FlatBatch:
    """Flat structure-of-arrays batch (bwq_batch).  Build with ``encode_batch`` or directly from
    numpy arrays (the synthetic families in ml_qem_b200.families do the latter)."""

    def __init__(self, n_qubits, op_offsets, ops, params, obs_offsets, term_offsets, term_x, term_z, term_coeff):
        self.n_qubits = np.ascontiguousarray(n_qubits, dtype=np.int32)
        self.op_offsets = np.ascontiguousarray(op_offsets, dtype=np.int64)
        self.ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        self.params = np.ascontiguousarray(params, dtype=np.float64)
        self.obs_offsets = np.ascontiguousarray(obs_offsets, dtype=np.int64)
        self.term_offsets = np.ascontiguousarray(term_offsets, dtype=np.int64)
        self.term_x = np.ascontiguousarray(term_x, dtype=np.uint64)
        self.term_z = np.ascontiguousarray(term_z, dtype=np.uint64)
        self.term_coeff = np.ascontiguousarray(term_coeff, dtype=np.float64)
        self.n_circuits = len(self.n_qubits)
        if len(self.op_offsets) != self.n_circuits + 1 or len(self.obs_offsets) != self.n_circuits + 1:
            raise ValueError("offset arrays must have n_circuits + 1 entries")
        self.n_observables = int(self.obs_offsets[-1]) if self.n_circuits else 0
        if len(self.term_offsets) != self.n_observables + 1:
            raise ValueError("term_offsets must have n_observables + 1 entries")

    def c_struct(self):
        return _Batch(self.n_circuits, _ptr(self.n_qubits), _ptr(self.op_offsets), _ptr(self.ops), _ptr(self.params),
                      len(self.params), _ptr(self.obs_offsets), _ptr(self.term_offsets), _ptr(self.term_x),
                      _ptr(self.term_z), _ptr(self.term_coeff))

    def nbytes(self):
        return sum(a.nbytes for a in (self.n_qubits, self.op_offsets, self.ops, self.params, self.obs_offsets,
                                      self.term_offsets, self.term_x, self.term_z, self.term_coeff))

    def select(self, idx):
        """Sub-batch with the circuits ``idx`` (used to shard a batch across ranks).  Only the
        parameters the selected gates reference are copied (compacted per circuit), so batches whose
        variants share parameter slots at the end of the array (zne.twirl_batch / fold_batch) do not
        blow up."""
        idx = np.asarray(idx, dtype=np.int64)
        npar_of = _num_params_table()
        op_lo, op_hi = self.op_offsets[idx], self.op_offsets[idx + 1]
        op_off = np.concatenate([[0], np.cumsum(op_hi - op_lo)])
        gather = _ranges(op_lo, op_hi)
        ops = self.ops[gather].copy()
        # parameters: one run [param_idx, param_idx + n_params(opcode)) per gate, re-packed in gate order
        npar = npar_of[ops["opcode"]].astype(np.int64)
        new_idx = np.concatenate([[0], np.cumsum(npar)])
        params = self.params[_ranges(ops["param_idx"].astype(np.int64), ops["param_idx"].astype(np.int64) + npar)]
        if new_idx[-1] >= 2 ** 32:
            raise ValueError("sub-batch has too many parameters for 32-bit indices")
        ops["param_idx"] = new_idx[:-1].astype(np.uint32)
        ob_lo, ob_hi = self.obs_offsets[idx], self.obs_offsets[idx + 1]
        obs_off = np.concatenate([[0], np.cumsum(ob_hi - ob_lo)])
        obs = _ranges(ob_lo, ob_hi)
        t_lo, t_hi = self.term_offsets[obs], self.term_offsets[obs + 1]
        term_off = np.concatenate([[0], np.cumsum(t_hi - t_lo)])
        terms = _ranges(t_lo, t_hi)
        return FlatBatch(self.n_qubits[idx], op_off, ops, params, obs_off, term_off,
                         self.term_x[terms], self.term_z[terms], self.term_coeff[terms])


def _ranges(lo, hi):
    """Concatenation of arange(lo[i], hi[i]) for all i (vectorised)."""
    lo = np.asarray(lo, dtype=np.int64)
    n = np.asarray(hi, dtype=np.int64) - lo
    total = int(n.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64)
    starts = np.repeat(lo - np.concatenate([[0], np.cumsum(n)[:-1]]), n)
    return starts + np.arange(total, dtype=np.int64)


_NPAR_TABLE = None


def _num_params_table():
    global _NPAR_TABLE
    if _NPAR_TABLE is None:
        t = np.zeros(max(OPCODES.values()) + 1, dtype=np.int32)
        for name, code in OPCODES.items():
            t[code] = NUM_PARAMS.get(name, 0)
        _NPAR_TABLE = t
    return _NPAR_TABLE


def encode_batch(circuits, observables):
    """circuits[i] with its list of observables observables[i] -> FlatBatch.

    ``observables[i]`` is a list of Pauli observables evaluated on the SAME simulated state (the
    reference re-simulates the circuit once per observable, e.g. docs/tutorials/zne_parallel.py:259
    passes ``[circ] * 4``; here one evolution serves them all).  Vectorised: every circuit
    contributes its cached flat arrays (Circuit.flat), every observable its cached masks; lists of
    observables shared between circuits (``[obs] * n``) are converted once."""
    n_qubits, op_cnt, obs_cnt = [], [], []
    opc, q0s, q1s, npars, params = [], [], [], [], []
    tx, tz, tc, term_cnt = [], [], [], []
    obs_cache = {}
    for circ, obs_list in zip(circuits, observables):
        circ = circuit_mod.from_any(circ)
        o, a, b, k, p = circ.flat()
        n_qubits.append(circ.num_qubits)
        op_cnt.append(len(o))
        opc.append(o); q0s.append(a); q1s.append(b); npars.append(k); params.append(p)
        key = tuple(map(id, obs_list))  # the same observable objects in the same order
        conv = obs_cache.get(key)
        if conv is None:
            obs_conv = [observable_mod.from_any(ob) for ob in obs_list]
            masks = [ob.masks() for ob in obs_conv]
            conv = (obs_list, obs_conv,
                    np.concatenate([m[0] for m in masks]) if masks else np.zeros(0, dtype=np.uint64),
                    np.concatenate([m[1] for m in masks]) if masks else np.zeros(0, dtype=np.uint64),
                    np.concatenate([m[2] for m in masks]) if masks else np.zeros(0, dtype=np.float64),
                    [len(m[2]) for m in masks])
            obs_cache[key] = conv
        for ob in conv[1]:
            if ob.num_qubits != circ.num_qubits and len(ob):
                raise ValueError(f"observable acts on {ob.num_qubits} qubits, circuit has {circ.num_qubits}")
        tx.append(conv[2]); tz.append(conv[3]); tc.append(conv[4]); term_cnt.extend(conv[5])
        obs_cnt.append(len(conv[1]))
    cat = lambda xs, dt: np.concatenate(xs).astype(dt, copy=False) if xs else np.zeros(0, dtype=dt)
    n_ops = int(sum(op_cnt))
    ops = np.zeros(n_ops, dtype=OP_DTYPE)
    if n_ops:
        ops["opcode"], ops["q0"], ops["q1"] = cat(opc, np.uint16), cat(q0s, np.uint8), cat(q1s, np.uint8)
        k = cat(npars, np.int64)
        # param_idx = first parameter of the op (ops without parameters point at the running offset)
        ops["param_idx"] = (np.cumsum(k) - k).astype(np.uint32)
    off = lambda cnt: np.concatenate([[0], np.cumsum(np.asarray(cnt, dtype=np.int64))]) if len(cnt) else np.zeros(1, dtype=np.int64)
    return FlatBatch(n_qubits, off(op_cnt), ops, cat(params, np.float64), off(obs_cnt), off(term_cnt),
                     cat(tx, np.uint64), cat(tz, np.uint64), cat(tc, np.float64))


def _noise_struct(table):
    if table is None or len(table["opcode"]) == 0:
        return None, None
    keep = {k: np.ascontiguousarray(v) for k, v in table.items()}
    st = _NoiseTable(len(keep["opcode"]), _ptr(keep["opcode"]), _ptr(keep["q0"]), _ptr(keep["q1"]), _ptr(keep["kind"]),
                     _ptr(keep["data_off"]), _ptr(keep["data"]), len(keep["data"]))
    return st, keep


_KEEP_NOISE = object()  # run_dm(noise=...) default: keep the installed table

STATUS_TEXT = {0: "ok", 1: "unsupported or malformed operation", 2: "too many active qubits", 3: "qubit index out of range"}


class Engine:
    """One engine per GPU (bwq_ctx).  Not re-entrant: calls are serialised with a lock."""

    def __init__(self, device=0, **options):
        self._lib = load_library()
        self._ctx = C.c_void_p()
        rc = self._lib.bwq_create(int(device), C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.bwq_last_error(None).decode()
            self._ctx = None
            raise EngineError(f"bwq_create failed ({rc}): {msg}")
        self.device = int(device)
        self._lock = threading.RLock()
        self._noise_id = None
        self._noise_ref = self._noise_table_ref = None
        if options:
            self.set_options(**options)

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.bwq_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError(f"{what} failed ({rc}): {self._lib.bwq_last_error(self._ctx).decode()}")

    def set_options(self, tile_qubits=0, low_qubits=0, max_state_bytes=0, chunk_circuits=0, host_threads=0,
                    sv_tile_bits=0, flags=0):
        """flags: BWQ_OPT_* bits of include/bwq.h (1 = no direct load pass, 2 = no direct store pass,
        4 / 8 = never / always pipeline bwq_dm_run, 16 = no TMA kernel)."""
        o = _Options(tile_qubits, low_qubits, max_state_bytes, chunk_circuits, host_threads, sv_tile_bits, flags)
        self._check(self._lib.bwq_set_options(self._ctx, C.byref(o)), "bwq_set_options")

    def set_noise(self, model):
        """model: ml_qem_b200.noise.NoiseModel or None (noise-free)."""
        key = id(model) if model is not None else None
        with self._lock:
            table = model.to_table() if model is not None else None
            st, keep = _noise_struct(table)
            self._check(self._lib.bwq_set_noise_table(self._ctx, C.byref(st) if st is not None else None),
                        "bwq_set_noise_table")
            self._noise_id = key
            # the installed table: the model object and the table it was built from (a model edited
            # afterwards rebuilds its table, so it is installed again)
            self._noise_ref, self._noise_table_ref = model, table

    def _ensure_noise(self, model):
        """Installs ``model`` unless it already is the installed table (caller holds the lock)."""
        if model is _KEEP_NOISE:
            return
        table = model.to_table() if model is not None else None
        if model is not self._noise_ref or table is not self._noise_table_ref:
            self.set_noise(model)

    def _run(self, fn, batch, what, noise=None):
        vals = np.empty(batch.n_observables, dtype=np.float64)
        status = np.zeros(batch.n_circuits, dtype=np.int32)
        st = batch.c_struct()
        with self._lock:
            self._ensure_noise(_KEEP_NOISE if noise is None else noise[0])
            self._check(fn(self._ctx, C.byref(st), vals.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p)), what)
        return vals, status

    def run_dm(self, batch, noise=_KEEP_NOISE):
        """Noisy values (density matrix) -> (values, status).  ``noise``: the NoiseModel the batch
        must run under -- installed (only if it is not the installed table already) and used
        under ONE lock, so engines shared between estimators / threads never mix tables; default:
        whatever ``set_noise`` installed last."""
        return self._run(self._lib.bwq_dm_run, batch, "bwq_dm_run", None if noise is _KEEP_NOISE else (noise,))

    def prepare_dm(self, batch):
        """Lowers + uploads the batch; the program stays resident on the device.  -> status"""
        status = np.zeros(batch.n_circuits, dtype=np.int32)
        st = batch.c_struct()
        with self._lock:
            self._check(self._lib.bwq_dm_prepare(self._ctx, C.byref(st), status.ctypes.data_as(C.c_void_p)), "bwq_dm_prepare")
        self._prepared_obs = batch.n_observables
        return status

    def execute_dm(self):
        """Runs the prepared batch (kernels + D2H of the values) -> values."""
        vals = np.empty(self._prepared_obs, dtype=np.float64)
        with self._lock:
            self._check(self._lib.bwq_dm_execute(self._ctx, vals.ctypes.data_as(C.c_void_p)), "bwq_dm_execute")
        return vals

    def prepare_sv(self, batch):
        status = np.zeros(batch.n_circuits, dtype=np.int32)
        st = batch.c_struct()
        with self._lock:
            self._check(self._lib.bwq_sv_prepare(self._ctx, C.byref(st), status.ctypes.data_as(C.c_void_p)), "bwq_sv_prepare")
        self._prepared_sv_obs = batch.n_observables
        return status

    def execute_sv(self):
        vals = np.empty(self._prepared_sv_obs, dtype=np.float64)
        with self._lock:
            self._check(self._lib.bwq_sv_execute(self._ctx, vals.ctypes.data_as(C.c_void_p)), "bwq_sv_execute")
        return vals

    def run_sv(self, batch):
        """Ideal values (statevector) -> (values, status)."""
        return self._run(self._lib.bwq_sv_run, batch, "bwq_sv_run")

    def run_meas_data(self, batch, noise=_KEEP_NOISE):
        """(ideal, noisy) values of every circuit in one call -- the batch form of the reference's
        ``create_estimator_meas_data`` (blackwater/data/utils.py:418-431); the statevector side
        runs concurrently with the density-matrix pipeline.  -> (ideal, noisy, status_ideal, status_noisy)"""
        ideal = np.empty(batch.n_observables, dtype=np.float64)
        noisy = np.empty(batch.n_observables, dtype=np.float64)
        st_i = np.zeros(batch.n_circuits, dtype=np.int32)
        st_n = np.zeros(batch.n_circuits, dtype=np.int32)
        st = batch.c_struct()
        with self._lock:
            self._ensure_noise(noise)
            self._check(self._lib.bwq_meas_data_run(self._ctx, C.byref(st), ideal.ctypes.data_as(C.c_void_p),
                                                    noisy.ctypes.data_as(C.c_void_p), st_i.ctypes.data_as(C.c_void_p),
                                                    st_n.ctypes.data_as(C.c_void_p)), "bwq_meas_data_run")
        return ideal, noisy, st_i, st_n

    def run_dm_variants(self, batch, variants, noise=_KEEP_NOISE):
        """Noisy values of every library-generated variant -> (values[n_observables * n_variants]
        circuit-major then variant then the circuit's observables, status[n_circuits * n_variants])."""
        vals = np.empty(batch.n_observables * variants.n_variants, dtype=np.float64)
        status = np.zeros(batch.n_circuits * variants.n_variants, dtype=np.int32)
        st, vs = batch.c_struct(), variants.c_struct()
        with self._lock:
            self._ensure_noise(noise)
            self._check(self._lib.bwq_dm_run_variants(self._ctx, C.byref(st), C.byref(vs), vals.ctypes.data_as(C.c_void_p),
                                                      status.ctypes.data_as(C.c_void_p)), "bwq_dm_run_variants")
        return vals, status

    def run_meas_data_variants(self, batch, variants, noise=_KEEP_NOISE):
        """(ideal, noisy) values with the variants (ZNE folds, Pauli twirls) generated inside the
        library: noisy[n_circuits, n_variants, obs...] flattened circuit-major, ideal as run_sv
        (the base circuits: folds and twirls leave the ideal circuit unchanged).
        -> (ideal, noisy, status_ideal, status_noisy)"""
        ideal = np.empty(batch.n_observables, dtype=np.float64)
        noisy = np.empty(batch.n_observables * variants.n_variants, dtype=np.float64)
        st_i = np.zeros(batch.n_circuits, dtype=np.int32)
        st_n = np.zeros(batch.n_circuits, dtype=np.int32)
        st = batch.c_struct()
        vs = variants.c_struct()
        with self._lock:
            self._ensure_noise(noise)
            self._check(self._lib.bwq_meas_data_run_variants(self._ctx, C.byref(st), C.byref(vs), ideal.ctypes.data_as(C.c_void_p),
                                                             noisy.ctypes.data_as(C.c_void_p), st_i.ctypes.data_as(C.c_void_p),
                                                             st_n.ctypes.data_as(C.c_void_p)), "bwq_meas_data_run_variants")
        return ideal, noisy, st_i, st_n

    def svx_exchange(self, local_ptr, peer_ptrs, rank, n_local_amps, stream=0, push=False):
        """EXCHANGE of the sharded statevector through peer memory: pull (local = new shard,
        peers = old shards) or push (local = old shard, peers = new shards)."""
        peers = np.asarray(peer_ptrs, dtype=np.uint64)
        fn = self._lib.bwq_svx_exchange_push if push else self._lib.bwq_svx_exchange_pull
        self._check(fn(self._ctx, C.c_void_p(int(local_ptr)), peers.ctypes.data_as(C.c_void_p), len(peers), int(rank),
                       int(n_local_amps), C.c_void_p(int(stream)) if stream else None), "bwq_svx_exchange")

    def run_dm_into(self, batch, device_ptr, noise=_KEEP_NOISE):
        """Writes the values into device memory at ``device_ptr`` (e.g. torch ``tensor.data_ptr()``
        of a float64 CUDA tensor with n_observables elements) -- zero-copy label hand-off."""
        status = np.zeros(batch.n_circuits, dtype=np.int32)
        st = batch.c_struct()
        with self._lock:
            self._ensure_noise(noise)
            self._check(self._lib.bwq_dm_run_device_out(self._ctx, C.byref(st), C.c_void_p(int(device_ptr)),
                                                        status.ctypes.data_as(C.c_void_p)), "bwq_dm_run_device_out")
        return status

    def stats(self):
        s = _Stats()
        self._check(self._lib.bwq_get_stats(self._ctx, C.byref(s)), "bwq_get_stats")
        return {f: getattr(s, f) for f, _ in _Stats._fields_}

    def sync(self):
        self._check(self._lib.bwq_sync(self._ctx), "bwq_sync")


def expand_variants(batch, variants):
    """Host-only view of the library's variant generation (no GPU): the expanded FlatBatch
    (observables replicated per variant), e.g. to compare with circuits built in Python."""
    lib = load_library()
    bs, vs = batch.c_struct(), variants.c_struct()
    sizes = np.zeros(4, dtype=np.int64)
    rc = lib.bwq_expand_variants(C.byref(bs), C.byref(vs), sizes.ctypes.data_as(C.c_void_p), None, None, None)
    if rc != 0:
        raise EngineError(f"bwq_expand_variants failed ({rc})")
    n, n_ops, n_par, n_var = (int(x) for x in sizes)
    op_off = np.zeros(n + 1, dtype=np.int64)
    ops = np.zeros(n_ops, dtype=OP_DTYPE)
    params = np.zeros(n_par, dtype=np.float64)
    lib.bwq_expand_variants(C.byref(bs), C.byref(vs), sizes.ctypes.data_as(C.c_void_p), op_off.ctypes.data_as(C.c_void_p),
                            ops.ctypes.data_as(C.c_void_p), params.ctypes.data_as(C.c_void_p))
    rep = np.repeat(np.arange(batch.n_circuits), n_var)
    obs_cnt = np.diff(batch.obs_offsets)[rep]
    obs_src = _ranges(batch.obs_offsets[rep], batch.obs_offsets[rep + 1])
    term_src = _ranges(batch.term_offsets[obs_src], batch.term_offsets[obs_src + 1])
    return FlatBatch(batch.n_qubits[rep], op_off, ops, params, np.concatenate([[0], np.cumsum(obs_cnt)]),
                     np.concatenate([[0], np.cumsum(np.diff(batch.term_offsets)[obs_src])]),
                     batch.term_x[term_src], batch.term_z[term_src], batch.term_coeff[term_src])


def lower_dm(batch, circuit, noise_model=None, tile_qubits=0, low_qubits=0, tma=False, tma_direct_store=False, fold=1):
    """Host-only view of the lowering stage (no GPU): returns the sweep program of one circuit as a
    dict of numpy arrays (see bwq_program_read in include/bwq.h).  tma=True: the TMA tile layout
    the engine uses by default for circuits wider than the tile."""
    lib = load_library()
    st, keep = _noise_struct(noise_model.to_table() if noise_model is not None else None)
    prog = C.c_void_p()
    bs = batch.c_struct()
    rc = lib.bwq_lower_dm_ex(C.byref(st) if st is not None else None, C.byref(bs), circuit, tile_qubits, low_qubits,
                             (1 if tma else 0) | (2 if tma_direct_store else 0) | (int(fold) << 8), C.byref(prog))
    if rc != 0:
        raise EngineError(f"bwq_lower_dm failed ({rc}): {lib.bwq_last_error(None).decode()}")
    try:
        sizes = np.zeros(8, dtype=np.int64)
        lib.bwq_program_sizes(prog, sizes.ctypes.data_as(C.c_void_p))
        nd, nsw, nps, nprog, dense, status, nt, ng = (int(x) for x in sizes)
        out = {
            "n_digits": nd, "status": status, "n_gates": ng, "n_passes": nps, "needs_dense": bool(dense),
            "active": np.zeros(nd, dtype=np.int32), "sweeps": np.zeros((nsw, 10), dtype=np.int32),
            "prog": np.zeros(nprog, dtype=np.uint64), "term_index": np.zeros(nt, dtype=np.int64),
            "term_coeff": np.zeros(nt, dtype=np.float64),
        }
        lib.bwq_program_read(prog, *[out[k].ctypes.data_as(C.c_void_p) for k in
                                     ("active", "sweeps", "prog", "term_index", "term_coeff")])
        return out
    finally:
        lib.bwq_program_free(prog)


SEG_SWEEPS, SEG_EXCHANGE, SEG_EXPVAL = 0, 1, 2


class SvxProgram:
    """Handle of a lowered wide/sharded statevector program (bwq_svx_program).  ``info`` holds the
    host-side view (no GPU needed): segments, sweeps, program words, Z-type terms."""

    def __init__(self, batch, circuit=0, tile_bits=0, n_global_bits=0):
        self._lib = load_library()
        self._h = C.c_void_p()
        bs = batch.c_struct()
        rc = self._lib.bwq_svx_lower(C.byref(bs), circuit, tile_bits, n_global_bits, C.byref(self._h))
        if rc != 0:
            raise EngineError(f"bwq_svx_lower failed ({rc}): {self._lib.bwq_last_error(None).decode()}")
        sizes = np.zeros(12, dtype=np.int64)
        self._lib.bwq_svx_sizes(self._h, sizes.ctypes.data_as(C.c_void_p))
        (status, n_bits, n_local, n_global, tile, nsw, nprog, nseg, nzt, npass, nex, nobs) = (int(x) for x in sizes)
        self.info = {
            "status": status, "n_bits": n_bits, "n_local": n_local, "n_global": n_global, "tile_bits": tile,
            "n_passes": npass, "n_exchanges": nex, "n_observables": nobs,
            "active": np.zeros(n_bits, dtype=np.int32), "sweeps": np.zeros((nsw, 10), dtype=np.int32),
            "prog": np.zeros(nprog, dtype=np.uint64), "segs": np.zeros((nseg, 4), dtype=np.int32),
            "zt_mask": np.zeros(nzt, dtype=np.uint32), "zt_coeff": np.zeros(nzt, dtype=np.float64),
            "zt_obs": np.zeros(nzt, dtype=np.int32),
        }
        self._lib.bwq_svx_read(self._h, *[self.info[k].ctypes.data_as(C.c_void_p) for k in
                                          ("active", "sweeps", "prog", "segs", "zt_mask", "zt_coeff", "zt_obs")])

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bwq_svx_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def algorithmic_bytes(self, rank=0):
        return int(self._lib.bwq_svx_bytes(self._h, int(rank)))

    def upload(self, engine):
        engine._check(self._lib.bwq_svx_upload(engine._ctx, self._h), "bwq_svx_upload")

    def run_segment_push(self, engine, segment, state_ptr, rank, peer_ptrs, stream=0):
        """SWEEPS segment whose last sweep stores into the peers' new shards (the EXCHANGE that
        follows it is skipped by the caller): bwq_svx_run_segment_push."""
        peers = np.asarray(peer_ptrs, dtype=np.uint64)
        engine._check(self._lib.bwq_svx_run_segment_push(engine._ctx, self._h, int(segment), C.c_void_p(int(state_ptr)), int(rank),
                                                         peers.ctypes.data_as(C.c_void_p), len(peers),
                                                         C.c_void_p(int(stream)) if stream else None), "bwq_svx_run_segment_push")

    def run_segment(self, engine, segment, state_ptr, rank=0, obs_ptr=0, stream=0):
        engine._check(self._lib.bwq_svx_run_segment(engine._ctx, self._h, int(segment), C.c_void_p(int(state_ptr)),
                                                    int(rank), C.c_void_p(int(obs_ptr)) if obs_ptr else None,
                                                    C.c_void_p(int(stream)) if stream else None),
                      "bwq_svx_run_segment")
