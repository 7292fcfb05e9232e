"""Qiskit standard-gate matrices (little-endian), restated.

[3P] qiskit.circuit.library.standard_gates (qiskit-terra 0.24.1, requirements.txt:4-5); the
gate names are the ones the reference enumerates at blackwater/data/utils.py:19-49 and the
backend basis {id, rz, sx, x, cx, reset} (docs/tutorials/02_data_generation.ipynb cell 3).
Convention: for a gate on qargs (a, b) the local basis index is i_a + 2*i_b, so ``cx`` with
(a, b) = (control, target) is [[1,0,0,0],[0,0,0,1],[0,0,1,0],[0,1,0,0]].
"""
import cmath
import math

import numpy as np

I2 = np.eye(2, dtype=complex)
X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)
PAULI = {"I": I2, "X": X, "Y": Y, "Z": Z}


def u3(theta, phi, lam):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array(
        [[c, -cmath.exp(1j * lam) * s], [cmath.exp(1j * phi) * s, cmath.exp(1j * (phi + lam)) * c]],
        dtype=complex,
    )


def _controlled(u):
    """control = local qubit 0, target = local qubit 1 (index i_c + 2 i_t)."""
    m = np.eye(4, dtype=complex)
    for tr in range(2):
        for tc in range(2):
            m[1 + 2 * tr, 1 + 2 * tc] = u[tr, tc]
    return m


def _rot(p, theta):
    return math.cos(theta / 2) * np.eye(p.shape[0], dtype=complex) - 1j * math.sin(theta / 2) * p


def _kron2(b, a):
    """operator a on local qubit 0 and b on local qubit 1."""
    return np.kron(b, a)


def gate_matrix(name, params=()):
    n = name.lower()
    p = [float(x) for x in params]
    if n in ("unitary1", "unitary2"):
        # explicit matrix: row-major, (re, im) interleaved; 2-qubit index = i_q0 + 2 i_q1 (include/bwq.h:62-63)
        d = 2 if n == "unitary1" else 4
        return (np.array(p[0::2]) + 1j * np.array(p[1::2])).reshape(d, d)
    if n in ("id", "i"):
        return I2.copy()
    if n == "x":
        return X.copy()
    if n == "y":
        return Y.copy()
    if n == "z":
        return Z.copy()
    if n == "h":
        return np.array([[1, 1], [1, -1]], dtype=complex) / math.sqrt(2)
    if n == "s":
        return np.diag([1, 1j]).astype(complex)
    if n == "sdg":
        return np.diag([1, -1j]).astype(complex)
    if n == "t":
        return np.diag([1, cmath.exp(1j * math.pi / 4)]).astype(complex)
    if n == "tdg":
        return np.diag([1, cmath.exp(-1j * math.pi / 4)]).astype(complex)
    if n == "sx":
        return 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=complex)
    if n == "sxdg":
        return 0.5 * np.array([[1 - 1j, 1 + 1j], [1 + 1j, 1 - 1j]], dtype=complex)
    if n == "rx":
        return _rot(X, p[0])
    if n == "ry":
        return _rot(Y, p[0])
    if n == "rz":
        return np.diag([cmath.exp(-0.5j * p[0]), cmath.exp(0.5j * p[0])]).astype(complex)
    if n in ("p", "u1"):
        return np.diag([1, cmath.exp(1j * p[0])]).astype(complex)
    if n == "u2":
        return u3(math.pi / 2, p[0], p[1])
    if n in ("u3", "u"):
        return u3(p[0], p[1], p[2])
    if n in ("cx", "cnot"):
        return _controlled(X)
    if n == "cy":
        return _controlled(Y)
    if n == "cz":
        return _controlled(Z)
    if n == "ch":
        return _controlled(gate_matrix("h"))
    if n == "crx":
        return _controlled(_rot(X, p[0]))
    if n == "cry":
        return _controlled(_rot(Y, p[0]))
    if n == "crz":
        return _controlled(gate_matrix("rz", p))
    if n in ("cp", "cu1"):
        return _controlled(gate_matrix("p", p))
    if n == "cu3":
        return _controlled(u3(*p))
    if n == "swap":
        return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex)
    if n == "iswap":
        return np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], dtype=complex)
    if n == "rzz":
        return _rot(_kron2(Z, Z), p[0])
    if n == "rxx":
        return _rot(_kron2(X, X), p[0])
    if n == "ryy":
        return _rot(_kron2(Y, Y), p[0])
    if n == "rzx":
        # qiskit RZXGate(theta) = exp(-i theta/2 X(x)Z) in its (q1 (x) q0) matrix: Z on q0, X on q1
        return _rot(_kron2(X, Z), p[0])
    if n == "ecr":
        return np.array([[0, 1, 0, 1j], [1, 0, -1j, 0], [0, 1j, 0, 1], [-1j, 0, 1, 0]], dtype=complex) / math.sqrt(2)
    if n == "ccx":
        m = np.eye(8, dtype=complex)
        m[[3, 7]] = m[[7, 3]]
        return m
    if n == "cswap":
        m = np.eye(8, dtype=complex)
        m[[3, 5]] = m[[5, 3]]
        return m
    raise ValueError(f"oracle: unsupported gate {name!r}")


IGNORED = ("barrier", "measure", "delay")
