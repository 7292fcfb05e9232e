"""Pauli-transfer matrices (numpy, host side).

R[i, j] = Tr(P_i E(P_j)) / 2^k with P in {I, X, Y, Z} per qubit and index = digit_q0 + 4*digit_q1.
The engine stores a density matrix as its real Pauli-basis vector r[P] = Tr(rho P), so every
channel -- unitary, thermal relaxation, depolarizing, reset -- is a real matrix on r.
"""
import numpy as np

PAULIS = np.array([[[1, 0], [0, 1]], [[0, 1], [1, 0]], [[0, -1j], [1j, 0]], [[1, 0], [0, -1]]], dtype=complex)


def pauli_basis(k):
    if k == 1:
        return PAULIS
    # index i0 + 4 i1, matrix on basis index b0 + 2 b1 = kron(P_i1, P_i0)
    return np.array([np.kron(PAULIS[i1], PAULIS[i0]) for i1 in range(4) for i0 in range(4)])


def from_unitary(u):
    u = np.asarray(u, dtype=complex)
    k = {2: 1, 4: 2}[u.shape[0]]
    ps = pauli_basis(k)
    e = np.einsum("ab,jbc,dc->jad", u, ps, u.conj())  # U P_j U^dag
    return np.real(np.einsum("iab,jba->ij", ps, e)) / u.shape[0]


def from_kraus(kraus):
    kraus = [np.asarray(k, dtype=complex) for k in kraus]
    d = kraus[0].shape[0]
    ps = pauli_basis({2: 1, 4: 2}[d])
    e = sum(np.einsum("ab,jbc,dc->jad", k, ps, k.conj()) for k in kraus)
    return np.real(np.einsum("iab,jba->ij", ps, e)) / d


def tensor(r_q0, r_q1):
    """Channel r_q0 on local qubit 0 and r_q1 on local qubit 1."""
    return np.kron(r_q1, r_q0)


def embed(r, which):
    eye = np.eye(4)
    return tensor(r, eye) if which == 0 else tensor(eye, r)


def thermal_relaxation(t1, t2, time):
    """rho00 += p rho11, rho11 *= 1-p, coherences *= exp(-t/T2): I->I + p Z... in Pauli basis
    Z' = (1-p) Z + p I-coefficient, X' = e2 X, Y' = e2 Y."""
    p = 1.0 - np.exp(-time / t1) if np.isfinite(t1) else 0.0
    e2 = np.exp(-time / t2) if np.isfinite(t2) else 1.0
    r = np.diag([1.0, e2, e2, 1.0 - p])
    r[3, 0] = p
    return r


def depolarizing(p, k):
    r = np.eye(4 ** k) * (1.0 - p)
    r[0, 0] = 1.0
    return r


RESET = np.zeros((4, 4))
RESET[0, 0] = 1.0
RESET[3, 0] = 1.0


def process_fidelity(r):
    return float(np.trace(r)) / r.shape[0]


def average_gate_fidelity(r):
    d = int(round(np.sqrt(r.shape[0])))
    return (d * process_fidelity(r) + 1.0) / (d + 1.0)


def relax2_params(r, tol=0.0):
    """If the 16x16 PTM has the (relaxation (x) relaxation) o diagonal sparsity pattern return its
    25 parameters (d[16], ca[4], cb[4], cab) for the structured kernel op, else None."""
    r = np.asarray(r, dtype=float)
    d = np.diag(r).copy()
    ca = np.array([r[3 + 4 * b, 0 + 4 * b] for b in range(4)])
    cb = np.array([r[a + 12, a] for a in range(4)])
    cab = r[15, 0]
    rebuilt = np.diag(d)
    for b in range(4):
        rebuilt[3 + 4 * b, 4 * b] += ca[b]
    for a in range(4):
        rebuilt[a + 12, a] += cb[a]
    rebuilt[15, 0] += cab
    # ca[3]/cb[3] overlap with nothing else; (15,0) is cab only
    if np.max(np.abs(rebuilt - r)) > tol:
        return None
    return np.concatenate([d, ca, cb, [cab]])
