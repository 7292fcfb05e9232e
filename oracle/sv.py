"""Ideal labels: statevector evolution and <psi|P|psi>, restated in numpy complex128.

[3P] qiskit.primitives.Estimator (terra 0.24.1, ``shots=None``) evolves
``Statevector(circuit)`` and returns ``expectation_value(observable)``; call sites
docs/tutorials/h13_ising_data_gen_tomo.ipynb:811, docs/tutorials/vqe_data_gen_parallel.py:31,
blackwater/data/utils.py:422-424 (ideal AerEstimator).  Same circuit format as oracle.dm.
"""
import numpy as np

from . import gates as G
from .dm import _parity, pauli_masks, strip_final_measurements


def apply_unitary(psi, n, qubits, u):
    k = len(qubits)
    t = psi.reshape((2,) * n)
    axes = [n - 1 - q for q in reversed(qubits)]
    t = np.moveaxis(t, axes, range(k))
    shp = t.shape
    t = (u @ t.reshape(2 ** k, -1)).reshape(shp)
    t = np.moveaxis(t, range(k), axes)
    return np.ascontiguousarray(t).reshape(-1)


def simulate(n, ops):
    psi = np.zeros(2 ** n, dtype=complex)
    psi[0] = 1.0
    for name, qubits, params in strip_final_measurements(list(ops)):
        name = name.lower()
        if name == "delay":
            continue
        if name == "reset":
            raise ValueError("oracle.sv: reset is not unitary; use oracle.dm")
        psi = apply_unitary(psi, n, tuple(qubits), G.gate_matrix(name, params))
    return psi


def expval_pauli(psi, n, label):
    x, z, ny = pauli_masks(label)
    c = np.arange(2 ** n, dtype=np.int64)
    sign = 1.0 - 2.0 * _parity(c & z)
    # P|c> = i^ny sign(c) |c^x>  =>  <psi|P|psi> = sum_c conj(psi[c^x]) i^ny sign(c) psi[c]
    return complex((1j) ** ny * np.sum(np.conj(psi[c ^ x]) * sign * psi))


def expval(psi, n, observable):
    tot = 0.0 + 0.0j
    for label, coeff in observable:
        tot += coeff * expval_pauli(psi, n, label)
    return np.real_if_close(tot)


def estimate(n, ops, observables):
    psi = simulate(n, ops)
    return np.array([np.real(expval(psi, n, o)) for o in observables], dtype=float)
