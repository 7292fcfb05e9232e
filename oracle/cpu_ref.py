"""ctypes wrapper of oracle/cpu_ref.cpp (Aer-style C++/OpenMP restatement) -- test / baseline
infrastructure only.  Takes the same flat batch arrays as the C ABI (include/bwq.h: bwq_batch) and
the ORACLE's noise model (complex superoperators), so it shares no lowering code with the product."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libcpuref.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "cpu_ref.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s", "_build/libcpuref.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.cpuref_max_threads.restype = C.c_int
    return _lib


class _Batch(C.Structure):
    _fields_ = [("n_circuits", C.c_int32), ("n_qubits", C.c_void_p), ("op_offsets", C.c_void_p), ("ops", C.c_void_p),
                ("params", C.c_void_p), ("n_params", C.c_int64), ("obs_offsets", C.c_void_p),
                ("term_offsets", C.c_void_p), ("term_x", C.c_void_p), ("term_z", C.c_void_p), ("term_coeff", C.c_void_p)]


class _Noise(C.Structure):
    _fields_ = [("n_entries", C.c_int32), ("opcode", C.c_void_p), ("q0", C.c_void_p), ("q1", C.c_void_p),
                ("data_off", C.c_void_p), ("data", C.c_void_p)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def _batch_struct(fb):
    """fb: any object with the bwq_batch arrays as numpy attributes (ml_qem_b200.engine.FlatBatch)."""
    return _Batch(fb.n_circuits, _p(fb.n_qubits), _p(fb.op_offsets), _p(fb.ops), _p(fb.params), len(fb.params),
                  _p(fb.obs_offsets), _p(fb.term_offsets), _p(fb.term_x), _p(fb.term_z), _p(fb.term_coeff))


def noise_arrays(model, opcodes):
    """oracle.noise_model.NoiseModel -> arrays for cpuref_noise.  ``opcodes``: gate name -> opcode."""
    if model is None:
        return None
    opc, q0, q1, off, data = [], [], [], [], []
    pos = 0
    items = [((n, q), s) for (n, q), s in model.local.items()] + [((n, None), s) for n, s in model.default.items()]
    for (name, qubits), s in items:
        if name not in opcodes:
            continue
        opc.append(opcodes[name])
        q0.append(qubits[0] if qubits else 255)
        q1.append(qubits[1] if qubits and len(qubits) > 1 else 255)
        off.append(pos)
        flat = np.ascontiguousarray(s, dtype=np.complex128).reshape(-1)
        data.append(flat)
        pos += len(flat)
    keep = {"opcode": np.asarray(opc, dtype=np.uint16), "q0": np.asarray(q0, dtype=np.uint8),
            "q1": np.asarray(q1, dtype=np.uint8), "data_off": np.asarray(off, dtype=np.int64),
            "data": np.concatenate(data) if data else np.zeros(0, dtype=np.complex128)}
    return keep


_NAMES = None
_NPAR = None


def gate_table(fb):
    """Matrix of every op of the batch from oracle/gates.py (memoised per (gate, parameters)):
    -> (mats complex128 [total], off int64 [n_ops], -1 for reset / unsupported).  This is the ONLY gate
    library of the CPU restatement: the C++ side has none (no code shared with the product)."""
    global _NAMES, _NPAR
    from . import gates as G

    if _NAMES is None:
        from ml_qem_b200.gateset import NAMES, NUM_PARAMS  # opcode <-> name of the C ABI (include/bwq.h), data only

        _NAMES = dict(NAMES)
        _NPAR = {code: NUM_PARAMS.get(name, 0) for code, name in NAMES.items()}
    ops = fb.ops
    off = np.full(len(ops), -1, dtype=np.int64)
    cache, chunks, pos = {}, [], 0
    opcodes, pidx, params = ops["opcode"].tolist(), ops["param_idx"].tolist(), fb.params
    for g, (code, pi) in enumerate(zip(opcodes, pidx)):
        k = _NPAR.get(code, 0)
        key = (code, params[pi:pi + k].tobytes() if k else b"")
        hit = cache.get(key)
        if hit is None:
            name = _NAMES.get(code)
            p = params[pi:pi + k]
            if name is None or name == "reset":
                hit = -1
            else:
                if name == "unitary1":
                    m = (p[0::2] + 1j * p[1::2]).reshape(2, 2)
                elif name == "unitary2":
                    m = (p[0::2] + 1j * p[1::2]).reshape(4, 4)
                else:
                    m = G.gate_matrix(name, p)
                chunks.append(np.ascontiguousarray(m, dtype=np.complex128).reshape(-1))
                hit = pos
                pos += m.size
            cache[key] = hit
        off[g] = hit
    mats = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.complex128)
    return mats, off


def prepare(fb):
    """gate_table(fb), cached on the batch object (the table is not part of the timed region)."""
    tab = getattr(fb, "_cpuref_gates", None)
    if tab is None:
        tab = gate_table(fb)
        try:
            fb._cpuref_gates = tab
        except AttributeError:
            pass
    return tab


def run_dm(fb, noise_keep, threads=0, amplitude_parallel_qubits=0, fusion_threshold=0):
    lib = load()
    mats, goff = prepare(fb)
    out = np.zeros(fb.n_observables, dtype=np.float64)
    status = np.zeros(fb.n_circuits, dtype=np.int32)
    bs = _batch_struct(fb)
    ns = None
    if noise_keep is not None and len(noise_keep["opcode"]):
        ns = _Noise(len(noise_keep["opcode"]), _p(noise_keep["opcode"]), _p(noise_keep["q0"]), _p(noise_keep["q1"]),
                    _p(noise_keep["data_off"]), _p(noise_keep["data"]))
    lib.cpuref_dm_run(C.byref(bs), C.byref(ns) if ns is not None else None, _p(mats), _p(goff), _p(out), _p(status), int(threads),
                      int(amplitude_parallel_qubits), int(fusion_threshold))
    return out, status


def run_sv(fb, threads=0, amplitude_parallel_qubits=0):
    lib = load()
    out = np.zeros(fb.n_observables, dtype=np.float64)
    status = np.zeros(fb.n_circuits, dtype=np.int32)
    bs = _batch_struct(fb)
    mats, goff = prepare(fb)
    lib.cpuref_sv_run(C.byref(bs), _p(mats), _p(goff), _p(out), _p(status), int(threads), int(amplitude_parallel_qubits))
    return out, status


def max_threads():
    return load().cpuref_max_threads()
