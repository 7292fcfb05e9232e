"""Run-time probe for the REAL reference simulator -- test / baseline infrastructure.

qiskit / qiskit-aer are not in this image and cannot be installed offline, so everywhere else the
oracle is the numpy restatement (oracle/dm.py, oracle/sv.py).  A box that does have them (also
under a driver-provided ``baseline/_ref``) gets the genuine article: ``find()`` returns the module,
``estimate()`` evaluates circuits exactly the way the reference does at
blackwater/data/utils.py:422-430 (AerEstimator, density_matrix method, noise model attached,
``approximation=True``, ``shots=None``, ``skip_transpilation=True``; ideal values from
``qiskit.primitives.Estimator``).  Callers fall back to the restatement when ``find()`` is None
and label their numbers "port"; with Aer present they label them "reference".

The device noise model is rebuilt with Aer's OWN error constructors (thermal_relaxation_error,
depolarizing_error) from the same calibration dictionary, following basic_device_gate_errors
(SURVEY.md Appendix A.2), because a BackendV1 object cannot be made from a plain dict.
"""
import importlib
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find():
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.append(ref)
    try:
        if importlib.util.find_spec("qiskit") is None or importlib.util.find_spec("qiskit_aer") is None:
            return None
        return importlib.import_module("qiskit_aer")
    except Exception:  # noqa: BLE001 - a broken partial install counts as absent
        return None


def aer_noise_model(props_dict):
    """NoiseModel.from_backend semantics from a backend properties dict (the structure of
    tests/golden/backends.json): per gate depolarizing . thermal relaxation, see A.2."""
    from qiskit_aer.noise import NoiseModel, depolarizing_error, thermal_relaxation_error  # type: ignore
    from qiskit.quantum_info import average_gate_fidelity  # type: ignore

    from . import noise_model as onm

    nm = NoiseModel(basis_gates=["id", "rz", "sx", "x", "cx", "reset"])
    t1t2 = onm.qubit_relaxation(props_dict)
    for g in props_dict["gates"]:
        name, qubits = g["gate"], tuple(g["qubits"])
        par = g["parameters"]
        err_val = onm._val(par["gate_error"]) if "gate_error" in par else None
        length = onm._val(par["gate_length"]) if "gate_length" in par else 0.0
        relax = None
        if length and length > 0:
            for q in qubits:
                t1, t2 = t1t2[q]
                e = thermal_relaxation_error(t1, min(t2, 2 * t1), length)
                relax = e if relax is None else relax.expand(e)
        relax_fid = average_gate_fidelity(relax) if relax is not None else 1.0
        depol = None
        if err_val is not None and err_val > 1 - relax_fid:
            dim = 2 ** len(qubits)
            e = min(err_val, dim / (dim + 1))
            p = min(dim * (e - (1 - relax_fid)) / (dim * relax_fid - 1), 4 ** len(qubits) / (4 ** len(qubits) - 1))
            depol = depolarizing_error(p, len(qubits))
        combined = relax if depol is None else depol if relax is None else depol.compose(relax)
        if combined is not None:
            nm.add_quantum_error(combined, name, list(qubits))
    return nm


def to_qiskit(num_qubits, gate_ops):
    from qiskit import QuantumCircuit  # type: ignore

    qc = QuantumCircuit(num_qubits)
    for name, qubits, params in gate_ops:
        getattr(qc, name)(*params, *qubits)
    return qc


def estimate(num_qubits, gate_ops, observables, props_dict=None, threads=0):
    """values[len(observables)] from real Aer (noisy when ``props_dict`` is given, else ideal)."""
    import numpy as np
    from qiskit.quantum_info import SparsePauliOp  # type: ignore

    qc = to_qiskit(num_qubits, gate_ops)
    ops = [SparsePauliOp.from_list([(l, c) for l, c in ob]) for ob in observables]
    if props_dict is None:
        from qiskit.primitives import Estimator  # type: ignore

        est = Estimator()
        return np.asarray(est.run([qc] * len(ops), ops, shots=None).result().values, dtype=float)
    from qiskit_aer.primitives import Estimator as AerEstimator  # type: ignore

    opts = {"method": "density_matrix", "noise_model": aer_noise_model(props_dict)}
    if threads:
        opts["max_parallel_threads"] = int(threads)
    est = AerEstimator(backend_options=opts, run_options={"shots": None}, approximation=True, skip_transpilation=True)
    return np.asarray(est.run([qc] * len(ops), ops).result().values, dtype=float)
