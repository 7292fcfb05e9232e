"""Digital ZNE and Pauli twirling on the flat gate stream.

The reference builds every noise-amplified / twirled variant as a new Python circuit and
re-simulates it: ``zne(BackendEstimator)`` with ``ZNEStrategy(noise_factors, LocalFoldingAmplifier(
gates_to_fold=2), PolynomialExtrapolator(degree))`` (docs/tutorials/zne_parallel.py:168-189,256-270,
docs/tutorials/h16_zne.ipynb:206-229, blackwater/library/learning/estimator.py:33-86) and
``add_pauli_twirls`` (docs/tutorials/derek_files/phase_diagram.ipynb:776).  Here the variants are
produced from the encoded base batch with vectorised numpy on the flat arrays of ``FlatBatch`` --
no per-variant circuit objects -- and all of them go to the GPU as ONE batch:

  fold_batch     every 2-qubit gate G -> G (G^dagger G)^((factor-1)/2)   (local folding)
  twirl_batch    a uniformly random 2-qubit Pauli before every cx and its CX-conjugate after it
  extrapolate    least-squares polynomial in the noise factor, evaluated at 0
  zne(cls)       Estimator decorator: run(..., zne_strategy=ZNEStrategy(...)) like prototype-zne
"""
from dataclasses import dataclass, field

import numpy as np

from .engine import OP_DTYPE, FlatBatch
from .gateset import NUM_PARAMS, OPCODES

_NAME = {v: k for k, v in OPCODES.items()}
_TWO_Q = np.array(sorted(c for n, c in OPCODES.items() if 32 <= c <= 47 or n == "unitary2"), dtype=np.uint16)
_SELF_INVERSE = np.array([OPCODES[n] for n in ("cx", "cy", "cz", "ch", "swap", "ecr")], dtype=np.uint16)
_NEGATE = np.array([OPCODES[n] for n in ("crx", "cry", "crz", "cp", "rzz", "rxx", "ryy", "rzx")], dtype=np.uint16)
_CX, _X, _RZ = OPCODES["cx"], OPCODES["x"], OPCODES["rz"]


def _segments(offsets):
    """circuit index of every op of a flat stream with the given [n+1] offsets."""
    return np.repeat(np.arange(len(offsets) - 1), np.diff(offsets))


def _rebuild(batch, ops, params, circ_of_op):
    n = batch.n_circuits
    op_off = np.zeros(n + 1, dtype=np.int64)
    np.add.at(op_off, circ_of_op + 1, 1)
    return FlatBatch(batch.n_qubits, np.cumsum(op_off), ops, params, batch.obs_offsets, batch.term_offsets,
                     batch.term_x, batch.term_z, batch.term_coeff)


def _inverse_params(batch, idx):
    """Appends the parameters of the inverse of ops[idx] (non self-inverse gates) to a copy of the
    parameter array; returns (params, param_idx of the inverses)."""
    ops = batch.ops[idx]
    extra, pidx = [], np.zeros(len(idx), dtype=np.uint32)
    base = len(batch.params)
    for k, op in enumerate(ops):
        code, p0 = int(op["opcode"]), int(op["param_idx"])
        name = _NAME[code]
        npar = NUM_PARAMS.get(name, 0)
        p = batch.params[p0:p0 + npar]
        if code in _NEGATE:
            q = -p
        elif name == "cu3":
            q = np.array([-p[0], -p[2], -p[1]])
        elif name == "unitary2":
            u = (p[0::2] + 1j * p[1::2]).reshape(4, 4).conj().T.reshape(-1)
            q = np.stack([u.real, u.imag], axis=1).reshape(-1)
        else:
            raise ValueError(f"fold_batch: no inverse rule for 2-qubit gate {name!r}")
        pidx[k] = base + sum(len(e) for e in extra)
        extra.append(q)
    params = np.concatenate([batch.params] + extra) if extra else batch.params
    return params, pidx


def fold_batch(batch, factor):
    """Local folding of every 2-qubit gate by an odd ``factor`` (LocalFoldingAmplifier(
    gates_to_fold=2)): G -> G (G^dagger G)^((factor-1)/2).  Self-inverse gates (cx, cz, ecr, ...)
    simply repeat ``factor`` times, each repetition carrying its own device error."""
    factor = int(factor)
    if factor < 1 or factor % 2 == 0:
        raise ValueError("noise factors of local folding must be odd positive integers")
    if factor == 1:
        return batch
    ops = batch.ops
    two = np.isin(ops["opcode"], _TWO_Q)
    rep = np.where(two, factor, 1)
    circ = _segments(batch.op_offsets)
    new_ops = np.repeat(ops, rep)
    params = batch.params
    need_inv = two & ~np.isin(ops["opcode"], _SELF_INVERSE)
    if need_inv.any():
        idx = np.nonzero(need_inv)[0]
        params, inv_pidx = _inverse_params(batch, idx)
        start = np.cumsum(rep) - rep  # first copy of every op in new_ops
        for k, i in enumerate(idx):   # copies 1, 3, 5, ... are the inverse
            sl = slice(start[i] + 1, start[i] + factor, 2)
            new_ops["param_idx"][sl] = inv_pidx[k]
            if _NAME[int(ops["opcode"][i])] == "unitary2":
                pass  # same opcode, conjugate-transposed matrix
    return _rebuild(batch, new_ops, params, np.repeat(circ, rep))


# CX conjugation of a 2-qubit Pauli (control, target), symplectic (x, z) per qubit, sign dropped:
# X_c -> X_c X_t, Z_t -> Z_c Z_t
_SYM = np.array([[0, 0], [1, 0], [1, 1], [0, 1]])          # I X Y Z -> (x, z)
_INV = {(0, 0): 0, (1, 0): 1, (1, 1): 2, (0, 1): 3}


def _cx_conjugate(pc, pt):
    xc, zc = _SYM[pc].T
    xt, zt = _SYM[pt].T
    code = np.array([[0, 3], [1, 2]])  # [x][z] -> Pauli index
    return code[xc, zc ^ zt], code[xt ^ xc, zt]


def twirl_batch(batch, n_twirls, rng):
    """Every circuit -> ``n_twirls`` Pauli-twirled instances (contiguous), observables replicated.
    Before each cx a uniformly random Pauli pair (P_c, P_t), after it CX (P_c P_t) CX, so the
    ideal circuit is unchanged while coherent cx errors average to Pauli noise.  Paulis are emitted
    in the backend basis as in ml_qem_b200.families (X = x, Y = rz(pi) x, Z = rz(pi); identity emits
    nothing, so no spurious ``id`` gate errors are added)."""
    n_twirls = int(n_twirls)
    ops = batch.ops
    n_ops = len(ops)
    circ = _segments(batch.op_offsets)
    params = np.concatenate([batch.params, [np.pi]])
    pi_idx = len(batch.params)
    is_cx = ops["opcode"] == _CX
    out_ops, out_circ = [], []
    n_cx = int(is_cx.sum())
    cx_pos = np.nonzero(is_cx)[0]
    for t in range(n_twirls):
        pc, pt = rng.integers(0, 4, size=n_cx), rng.integers(0, 4, size=n_cx)
        qc, qt = _cx_conjugate(pc, pt)
        # per cx up to 4 + 1 + 4 ops: [rz c][x c][rz t][x t] cx [rz c][x c][rz t][x t]
        slots = np.zeros((n_ops, 9), dtype=OP_DTYPE)
        keep = np.zeros((n_ops, 9), dtype=bool)
        slots[:, 4] = ops
        keep[:, 4] = True
        for col, (pauli, qfield) in enumerate(((pc, "q0"), (pt, "q1"), (qc, "q0"), (qt, "q1"))):
            base = 0 if col < 2 else 5
            off = base + 2 * (col % 2)
            q = ops[qfield][cx_pos]
            need_rz = (pauli == 2) | (pauli == 3)
            need_x = (pauli == 1) | (pauli == 2)
            slots["opcode"][cx_pos, off] = _RZ
            slots["q0"][cx_pos, off] = q
            slots["param_idx"][cx_pos, off] = pi_idx
            keep[cx_pos, off] = need_rz
            slots["opcode"][cx_pos, off + 1] = _X
            slots["q0"][cx_pos, off + 1] = q
            keep[cx_pos, off + 1] = need_x
        flat_keep = keep.reshape(-1)
        out_ops.append(slots.reshape(-1)[flat_keep])
        out_circ.append(np.repeat(circ * n_twirls + t, keep.sum(axis=1)))
    all_ops = np.concatenate(out_ops) if out_ops else np.zeros(0, dtype=OP_DTYPE)
    all_circ = np.concatenate(out_circ) if out_circ else np.zeros(0, dtype=np.int64)
    order = np.argsort(all_circ, kind="stable")
    all_ops, all_circ = all_ops[order], all_circ[order]
    n_new = batch.n_circuits * n_twirls
    op_off = np.zeros(n_new + 1, dtype=np.int64)
    np.add.at(op_off, all_circ + 1, 1)
    # observables: replicate per twirl
    obs_cnt = np.diff(batch.obs_offsets)
    term_cnt = np.diff(batch.term_offsets)
    new_obs_off = np.concatenate([[0], np.cumsum(np.repeat(obs_cnt, n_twirls))])
    obs_src = np.concatenate([np.tile(np.arange(batch.obs_offsets[c], batch.obs_offsets[c + 1]), n_twirls)
                              for c in range(batch.n_circuits)]) if batch.n_circuits else np.zeros(0, dtype=np.int64)
    new_term_off = np.concatenate([[0], np.cumsum(term_cnt[obs_src])]) if len(obs_src) else np.zeros(1, dtype=np.int64)
    term_src = np.concatenate([np.arange(batch.term_offsets[o], batch.term_offsets[o + 1]) for o in obs_src]) \
        if len(obs_src) else np.zeros(0, dtype=np.int64)
    return FlatBatch(np.repeat(batch.n_qubits, n_twirls), np.cumsum(op_off), all_ops, params, new_obs_off, new_term_off,
                     batch.term_x[term_src], batch.term_z[term_src], batch.term_coeff[term_src])


def average_twirls(values, n_twirls, obs_per_circuit):
    """values of a twirl_batch run -> mean over the twirls, shape [n_circuits * obs_per_circuit]
    (every circuit must carry ``obs_per_circuit`` observables)."""
    v = np.asarray(values, dtype=float).reshape(-1, n_twirls, obs_per_circuit)
    return v.mean(axis=1).reshape(-1)


def extrapolate(values, factors, degree=None):
    """Polynomial extrapolation to zero noise.  values[..., len(factors)] -> [...]; degree defaults
    to len(factors) - 1 (Richardson); degree 1 = linear least squares (prototype-zne's
    PolynomialExtrapolator / LinearExtrapolator)."""
    x = np.asarray(factors, dtype=float)
    y = np.asarray(values, dtype=float)
    degree = len(x) - 1 if degree is None else int(degree)
    if degree < 1 or degree > len(x) - 1:
        raise ValueError("need 1 <= degree <= len(factors) - 1")
    v = np.vander(x, degree + 1, increasing=True)          # [k, d+1]
    coef = np.linalg.pinv(v) @ y[..., None]                # least squares; [..., d+1, 1]
    return coef[..., 0, 0]


@dataclass
class PolynomialExtrapolator:
    degree: int = 1

    def __call__(self, values, factors):
        return extrapolate(values, factors, self.degree)


@dataclass
class ZNEStrategy:
    """Same knobs as prototype-zne's ZNEStrategy as the reference uses it (noise_factors, local
    folding of the 2-qubit gates, polynomial extrapolator)."""
    noise_factors: tuple = (1, 3)
    extrapolator: PolynomialExtrapolator = field(default_factory=PolynomialExtrapolator)


def zne(cls):
    """Estimator decorator, used like the reference's ``zne(BackendEstimator)``
    (docs/tutorials/zne_parallel.py:168): ``ZNEEstimator = zne(B200Estimator);
    ZNEEstimator(backend=...).run(circuits, observables, zne_strategy=ZNEStrategy(...))``.
    B200Estimator understands the ``zne_strategy`` run option natively (all circuits of a noise
    factor run as one GPU batch; metadata["zne"] carries the noisy values at every factor), so the
    decorator only has to produce the subclass the calling code expects."""
    if not hasattr(cls, "_call"):
        raise TypeError("zne() expects an Estimator class of this package")
    return type("ZNE" + cls.__name__, (cls,), {})
