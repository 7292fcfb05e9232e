"""Host-side regressions (no GPU): batch selection, observable phases, duck-typed unitary gates,
learning(skip_transpile=False)."""
import warnings

import numpy as np
import pytest

from ml_qem_b200 import circuit as circuit_mod
from ml_qem_b200 import engine, families as F, observable, zne
from ml_qem_b200.engine import _num_params_table


def _gates(b):
    t = _num_params_table()
    return [(int(o["opcode"]), int(o["q0"]), int(o["q1"]),
             tuple(b.params[int(o["param_idx"]):int(o["param_idx"]) + t[o["opcode"]]])) for o in b.ops]


def test_select_copies_only_referenced_parameters():
    circs, _, obs = F.config_brick10_twirl(n_base=2, n_twirls=3, seed=1)
    fb = engine.encode_batch(circs, [obs] * len(circs))
    sub = fb.select([4, 1, 5])
    ref = engine.encode_batch([circs[4], circs[1], circs[5]], [obs] * 3)
    for k in ("n_qubits", "op_offsets", "obs_offsets", "term_offsets", "term_x", "term_z", "term_coeff"):
        assert np.array_equal(getattr(sub, k), getattr(ref, k)), k
    assert _gates(sub) == _gates(ref)
    # twirled / folded batches share parameter slots appended at the END of the array: selecting
    # must not copy the whole array once per circuit (400x blow-up before)
    base = engine.encode_batch(circs[:2], [obs] * 2)
    tw = zne.twirl_batch(base, 50, np.random.default_rng(3))
    sel = tw.select(np.arange(tw.n_circuits))
    assert _gates(sel) == _gates(tw)
    assert len(sel.params) <= len(tw.ops)
    fo = zne.fold_batch(base, 3).select([1, 0])
    assert fo.n_circuits == 2 and len(fo.params) <= len(fo.ops)
    assert fb.select([]).n_circuits == 0


def test_pauli_phase_becomes_the_coefficient():
    assert observable.from_any("-ZI").terms == [("ZI", -1.0 + 0j)]
    assert observable.from_any("iXX").terms == [("XX", 1j)]

    class FakePauli:  # qiskit.quantum_info.Pauli duck type
        def to_label(self):
            return "-iZY"

    ob = observable.from_any(FakePauli())
    assert ob.terms == [("ZY", -1j)] and ob.is_complex()
    with pytest.raises(ValueError):
        ob.masks()  # complex coefficients never lose their imaginary part silently
    x, z, c = ob.imag_part().masks()
    assert c.tolist() == [-1.0]

    class FakeSparse:  # SparsePauliOp duck type: labels may carry a phase as well
        class paulis:
            @staticmethod
            def to_labels():
                return ["-XI", "IZ"]
        coeffs = np.array([2.0, 0.5 + 0.25j])

    ob = observable.from_any(FakeSparse())
    assert ob.terms == [("XI", -2.0 + 0j), ("IZ", 0.5 + 0.25j)]


def test_duck_typed_unitary_gate_reaches_the_unitary_branch():
    class Bit:
        pass

    class Op:
        def __init__(self, name, params):
            self.name, self.params = name, params

    class Inst:
        def __init__(self, op, qubits):
            self.operation, self.qubits = op, qubits

    class QC:
        def __init__(self):
            self.qubits = [Bit(), Bit()]
            self.num_qubits = 2
            h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
            cz = np.diag([1, 1, 1, -1]).astype(complex)
            self.data = [Inst(Op("unitary", [h]), [self.qubits[0]]),
                         Inst(Op("unitary", [cz]), [self.qubits[0], self.qubits[1]]),
                         Inst(Op("rz", [0.3]), [self.qubits[1]])]

    c = circuit_mod.from_any(QC())
    names = [n for n, _, _ in c.gate_ops()]
    assert names == ["unitary1", "unitary2", "rz"]
    assert len(c.gate_ops()[0][2]) == 8 and len(c.gate_ops()[1][2]) == 32


def test_learning_skip_transpile_false_warns_without_qiskit():
    from ml_qem_b200 import learning

    c = F.tfim_circuit(2, 1, 0.3)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out = learning._transpile_or_warn([c], backend=None)
    assert out == [c]
    try:
        import qiskit  # noqa: F401
    except Exception:  # noqa: BLE001
        assert any("no transpiler" in str(x.message) for x in w)


def test_basis_decompositions_equal_their_gates_up_to_phase():
    """families.BasisBuilder: rx / u3 / h / Paulis written in the backend basis (rz, sx, x) are the
    named gates up to a global phase (matrices from the oracle's gate table)."""
    from oracle import gates as G
    from ml_qem_b200.families import BasisBuilder

    def unitary(build):
        b = BasisBuilder(1)
        build(b)
        u = np.eye(2, dtype=complex)
        for name, _, params in b.done().gate_ops():
            u = G.gate_matrix(name, params) @ u
        return u

    def same_up_to_phase(a, b):
        k = np.argmax(np.abs(b))
        ph = a.reshape(-1)[k] / b.reshape(-1)[k]
        return abs(abs(ph) - 1) < 1e-12 and np.max(np.abs(a - ph * b)) < 1e-12

    for t in (-2.7, 0.3, 1.9):
        assert same_up_to_phase(unitary(lambda b: b.rx(t, 0)), G.gate_matrix("rx", (t,)))
        assert same_up_to_phase(unitary(lambda b: b.u3(t, 0.4, -1.1, 0)), G.gate_matrix("u3", (t, 0.4, -1.1)))
    assert same_up_to_phase(unitary(lambda b: b.h(0)), G.gate_matrix("h"))
    for p, name in ((1, "x"), (2, "y"), (3, "z")):
        assert same_up_to_phase(unitary(lambda b: b.pauli(p, 0)), G.gate_matrix(name))


def test_bound_view_encodes_like_bind_parameters():
    """Circuit.bound_view (what B200Estimator uses for run(batch * [ansatz], ..., parameter_values)):
    the parametrised circuit is walked once, every parameter set binds with one numpy expression;
    the encoded batch is identical to binding circuit by circuit."""
    from ml_qem_b200 import Circuit, engine
    from ml_qem_b200.circuit import Parameter

    th = [Parameter(f"t[{i}]") for i in range(11)]  # t[10] sorts after t[9], as ParameterVector does
    c = Circuit(3)
    for i, t in enumerate(th):
        c.ry(t, i % 3)
        if i % 3 == 2:
            c.cx(0, 1); c.cx(1, 2)
    c.rz(2 * th[0] + 0.5, 1); c.rx(-th[10], 2); c.u3(th[3] / 2, 0.1, th[4] - 1.0, 0); c.p(0.3, 0)
    rng = np.random.default_rng(3)
    vals = rng.uniform(-3, 3, size=(40, 11))
    obs = [[("ZZI", 1.0), ("XIX", 0.5)], [("IIY", 1.0)]]
    a = engine.encode_batch([c.bind_parameters(list(v)) for v in vals], [obs] * len(vals))
    b = engine.encode_batch([c.bound_view(v) for v in vals], [obs] * len(vals))
    for f in ("n_qubits", "op_offsets", "ops", "params", "obs_offsets", "term_offsets", "term_x", "term_z", "term_coeff"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    v = c.bound_view(vals[0])
    assert v.gate_ops() == c.bind_parameters(list(vals[0])).gate_ops() and v.num_parameters == 0 and v.size() == c.size()
    with pytest.raises(ValueError):
        c.bound_view(vals[0][:5])
