"""CPU: ZNE folding / Pauli twirling on the flat gate stream against the circuit-level builders of
ml_qem_b200.families and against the oracle; extrapolation."""
import numpy as np
import pytest

import helpers
from ml_qem_b200 import engine, families as F, zne
from ml_qem_b200.circuit import Circuit
from ml_qem_b200.gateset import NAMES


def _ops(batch, c):
    return [(NAMES[int(o["opcode"])], int(o["q0"]), int(o["q1"])) for o in batch.ops[batch.op_offsets[c]:batch.op_offsets[c + 1]]]


def test_fold_batch_equals_circuit_level_folding():
    circs = [F.tfim_circuit(4, 3, 0.3, layout=[0, 1, 3, 4], num_physical=5, fold=1), F.brickwork_circuit(5, 2, np.random.default_rng(1))]
    obs = [F.single_z_observables([0, 1, 3, 4], 5), F.single_z_observables(list(range(5)), 5)]
    base = engine.encode_batch(circs, obs)
    for f in (3, 5):
        folded = zne.fold_batch(base, f)
        want = engine.encode_batch([F.tfim_circuit(4, 3, 0.3, layout=[0, 1, 3, 4], num_physical=5, fold=f),
                                    F.brickwork_circuit(5, 2, np.random.default_rng(1), fold=f)], obs)
        for c in range(2):
            assert _ops(folded, c) == _ops(want, c)
    with pytest.raises(ValueError):
        zne.fold_batch(base, 2)


def test_fold_non_self_inverse_gates_is_identity_on_ideal_values():
    rng = np.random.default_rng(2)
    c = Circuit(4)
    for _ in range(25):
        a, b = (int(x) for x in rng.choice(4, size=2, replace=False))
        k = int(rng.integers(0, 6))
        th = float(rng.uniform(-3, 3))
        if k == 0: c.append("rzz", (a, b), (th,))
        elif k == 1: c.append("crx", (a, b), (th,))
        elif k == 2: c.append("cu3", (a, b), (th, 0.4, -0.9))
        elif k == 3: c.append("ecr", (a, b))
        elif k == 4: c.append("u3", (a,), (th, 0.1, 0.2))
        else: c.append("cx", (a, b))
    obs = [[("XYZI", 1.0)], [("ZZZZ", 1.0)]]
    base = engine.encode_batch([c], [obs])
    ref = helpers.oracle_sv_values(c, obs)
    folded = zne.fold_batch(base, 3)
    # rebuild a Circuit from the folded stream and evaluate with the oracle: G G^dagger G == G
    c3 = Circuit(4)
    from ml_qem_b200.gateset import NUM_PARAMS
    for o in folded.ops:
        name = NAMES[int(o["opcode"])]
        npar = NUM_PARAMS.get(name, 0)
        qs = (int(o["q0"]), int(o["q1"])) if 32 <= int(o["opcode"]) <= 47 else (int(o["q0"]),)
        c3.append(name, qs, tuple(float(x) for x in folded.params[int(o["param_idx"]):int(o["param_idx"]) + npar]))
    assert len(c3.ops) > len(c.ops)
    assert np.max(np.abs(helpers.oracle_sv_values(c3, obs) - ref)) < 1e-12


def test_twirl_batch_preserves_ideal_values_and_layout():
    rng = np.random.default_rng(3)
    circs = [F.brickwork_circuit(4, 2, np.random.default_rng(5)), F.tfim_circuit(3, 2, 0.4, basis="X")]
    obs = [[[("ZIIZ", 1.0)], [("IXXI", 0.5), ("ZZZZ", 1.0)]], [[("ZZI", 1.0)]]]
    base = engine.encode_batch(circs, obs)
    tw = zne.twirl_batch(base, 6, rng)
    assert tw.n_circuits == 12 and tw.n_observables == 6 * 2 + 6 * 1
    from ml_qem_b200.gateset import NUM_PARAMS
    k = 0
    for c in range(2):
        ref = helpers.oracle_sv_values(circs[c], obs[c])
        for t in range(6):
            i = c * 6 + t
            cc = Circuit(int(tw.n_qubits[i]))
            n_cx = 0
            for o in tw.ops[tw.op_offsets[i]:tw.op_offsets[i + 1]]:
                name = NAMES[int(o["opcode"])]
                n_cx += name == "cx"
                npar = NUM_PARAMS.get(name, 0)
                qs = (int(o["q0"]), int(o["q1"])) if name == "cx" else (int(o["q0"]),)
                cc.append(name, qs, tuple(float(x) for x in tw.params[int(o["param_idx"]):int(o["param_idx"]) + npar]))
            assert n_cx == sum(1 for nme, *_ in circs[c].gate_ops() if nme == "cx")
            # twirling leaves the ideal circuit unchanged (up to a global phase)
            assert np.max(np.abs(helpers.oracle_sv_values(cc, obs[c]) - ref)) < 1e-12
    avg = zne.average_twirls(np.arange(12.0), 6, 1)
    assert np.allclose(avg, [2.5, 8.5])


def test_extrapolation():
    f = (1, 3, 5)
    y = np.array([[2.0 - 0.1 * x + 0.01 * x * x for x in f], [1.0 - 0.3 * x for x in f]])
    assert np.allclose(zne.extrapolate(y, f), [2.0, 1.0])
    assert np.allclose(zne.extrapolate(y[1], f, degree=1), 1.0)
    lin = zne.extrapolate(y[0], f, degree=1)  # least squares line through a parabola
    assert abs(lin - np.polyfit(f, y[0], 1)[1]) < 1e-12
    assert zne.PolynomialExtrapolator(2)(y, f).shape == (2,)
