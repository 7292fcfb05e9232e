"""Aer ``density_matrix`` method + ``save_expectation_value``, restated in numpy complex128.

[3P] qiskit-aer DensityMatrix state (call sites blackwater/data/utils.py:422-430): the state
is vec(rho) column-stacked -- amplitude index r + c*2^n, qubit q's row bit is bit q and its
column bit is bit q+n.  A k-qubit superoperator acts on bits {q_i} U {q_i+n}; a unitary is
conj(U) (x) U; every QuantumError attached to (gate, qubits) is applied as ONE superoperator
right AFTER the gate (density_matrix method averages the error deterministically).
Exact expectation values: Tr(rho P) as in docs/tutorials/vqe_rf.py:57-83 (diag of rho for
Z-type strings), extended to X/Y by pairing rho[i, i^x].

Circuit format: iterable of (name: str, qubits: tuple[int], params: tuple[float]); qubit 0 is
the least-significant bit; Pauli labels are Qiskit strings (right-most char = qubit 0).
"""
import numpy as np

from . import gates as G
from .noise_model import RESET_SUPEROP, unitary_superop


def strip_final_measurements(ops):
    ops = [o for o in ops if o[0] != "barrier"]
    while ops and ops[-1][0] == "measure":
        ops.pop()
    if any(o[0] == "measure" for o in ops):
        raise ValueError("oracle: mid-circuit measurement is not supported in exact mode")
    return ops


def zero_state(n):
    v = np.zeros(4 ** n, dtype=complex)
    v[0] = 1.0
    return v


def apply_superop(v, n, qubits, s):
    """v <- S v on bits (q_0..q_{k-1}, q_0+n..q_{k-1}+n); local index bit i <-> position i."""
    k = len(qubits)
    bits = list(qubits) + [q + n for q in qubits]  # local bit j -> global bit bits[j]
    t = v.reshape((2,) * (2 * n))  # axis a <-> global bit 2n-1-a
    axes = [2 * n - 1 - b for b in reversed(bits)]  # most-significant local bit first
    t = np.moveaxis(t, axes, range(2 * k))
    shp = t.shape
    t = (s @ t.reshape(4 ** k, -1)).reshape(shp)
    t = np.moveaxis(t, range(2 * k), axes)
    return np.ascontiguousarray(t).reshape(-1)


def simulate(n, ops, noise=None):
    """Returns vec(rho) after the circuit; ``noise`` is an oracle.noise_model.NoiseModel or None."""
    v = zero_state(n)
    for name, qubits, params in strip_final_measurements(list(ops)):
        name = name.lower()
        qubits = tuple(qubits)
        if name in ("delay",):
            continue
        if name == "reset":
            v = apply_superop(v, n, qubits, RESET_SUPEROP)
        else:
            v = apply_superop(v, n, qubits, unitary_superop(G.gate_matrix(name, params)))
        if noise is not None:
            s = noise.get(name, qubits)
            if s is not None:
                v = apply_superop(v, n, qubits, s)
    return v


def pauli_masks(label):
    n = len(label)
    x = z = ny = 0
    for q in range(n):
        ch = label[n - 1 - q]
        if ch in "XY":
            x |= 1 << q
        if ch in "ZY":
            z |= 1 << q
        if ch == "Y":
            ny += 1
        if ch not in "IXYZ":
            raise ValueError(f"bad Pauli label {label!r}")
    return x, z, ny


def _parity(a):
    a = a.copy()
    for s in (32, 16, 8, 4, 2, 1):
        a ^= a >> s
    return a & 1


def expval_pauli(v, n, label):
    """Tr(rho P).  P|c> = i^{#Y} (-1)^{popcount(c & z)} |c ^ x>  =>  sum_c phase(c) rho[c, c^x]."""
    x, z, ny = pauli_masks(label)
    c = np.arange(2 ** n, dtype=np.int64)
    sign = 1.0 - 2.0 * _parity(c & z)
    vals = v[c + ((c ^ x) << n)]  # rho[r=c, c'=c^x] -> <c|rho|c^x>
    return complex((1j) ** ny * np.sum(sign * vals))


def expval_pauli_dense(v, n, label):
    rho = v.reshape(2 ** n, 2 ** n).T  # v[r + c 2^n] -> rho[r, c]
    p = np.array([[1.0]], dtype=complex)
    for ch in label:  # left-most char = highest qubit
        p = np.kron(p, G.PAULI[ch])
    return complex(np.trace(rho @ p))


def expval(v, n, observable):
    """observable: list of (label, coeff) -> np.real_if_close(sum_k c_k Tr(rho P_k))."""
    tot = 0.0 + 0.0j
    for label, coeff in observable:
        tot += coeff * expval_pauli(v, n, label)
    return np.real_if_close(tot)


def estimate(n, ops, observables, noise=None):
    """One circuit, many observables -> float array (Aer Estimator, approximation=True, shots=None)."""
    v = simulate(n, ops, noise)
    return np.array([np.real(expval(v, n, o)) for o in observables], dtype=float)
