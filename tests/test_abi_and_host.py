"""CPU: the C-ABI library loads and exports every symbol include/bwq.h declares; host-side logic
(QASM reader, observables, Estimator validation, noise table packing).  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from ml_qem_b200 import Circuit, Parameter, PauliObservable, backends, engine, noise, parse_qasm, ptm
from ml_qem_b200.circuit import from_any

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "bwq.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|void|const char\*)\s+(bwq_[a-z_0-9]+)\s*\(", hdr, re.M))
    assert declared >= {"bwq_create", "bwq_dm_run", "bwq_sv_run", "bwq_set_noise_table", "bwq_lower_dm"}
    for name in declared:
        assert getattr(lib, name) is not None, name
    assert set(engine.EXPORTS) == declared
    assert lib.bwq_version() == 100


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = ctypes.c_void_p()
    rc = lib.bwq_create(0, ctypes.byref(ctx))
    assert rc == -5 and not ctx.value  # BWQ_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.bwq_last_error(None)
    with pytest.raises(engine.EngineError):
        engine.Engine(0)


def test_opcode_table_matches_header():
    hdr = open(os.path.join(ROOT, "include", "bwq.h")).read()
    body = hdr[hdr.index("BWQ_G_ID = 0"):hdr.index("BWQ_G_COUNT")]
    names = re.findall(r"BWQ_G_([A-Z0-9]+)(?:\s*=\s*(\d+))?", body)
    val = -1
    from ml_qem_b200.gateset import OPCODES
    for name, explicit in names:
        val = int(explicit) if explicit else val + 1
        assert OPCODES[name.lower()] == val, name


def test_qasm_reader_handles_reference_dataset_text():
    text = """OPENQASM 2.0;
include "qelib1.inc";
qreg q[5];
creg meas[4];
rz(pi/2) q[1];
sx q[1];
rz(-pi/4) q[3];
x q[4];
cx q[4],q[3];
rz(1.3504439735577733) q[3];
barrier q[2],q[1],q[3],q[4];
measure q[2] -> meas[0];
measure q[1] -> meas[1];
"""
    c = parse_qasm(text)
    assert c.num_qubits == 5
    ops = c.gate_ops()
    assert [o[0] for o in ops] == ["rz", "sx", "rz", "x", "cx", "rz"]
    assert ops[0][2] == (np.pi / 2,) and ops[4][1] == (4, 3) and abs(ops[2][2][0] + np.pi / 4) < 1e-15
    assert [o[1][0] for o in c.ops if o[0] == "measure"] == [2, 1]


def test_qasm_gate_definitions_broadcast_and_errors():
    text = """OPENQASM 2.0; include "qelib1.inc";
gate bell a,b { h a; cx a,b; }
gate rot(t) a { rz(t/2) a; rx(-t) a; }
qreg a[2]; qreg b[2];
bell a[0],b[1];
rot(pi) b;
u1(0.5) a[1];
"""
    c = parse_qasm(text)
    assert c.num_qubits == 4
    assert c.ops == [("h", (0,), ()), ("cx", (0, 3), ()), ("rz", (2,), (np.pi / 2,)), ("rx", (2,), (-np.pi,)),
                     ("rz", (3,), (np.pi / 2,)), ("rx", (3,), (-np.pi,)), ("p", (1,), (0.5,))]
    with pytest.raises(ValueError):
        parse_qasm('OPENQASM 2.0; qreg q[1]; creg c[1]; if(c==1) x q[0];')
    with pytest.raises(ValueError):
        parse_qasm("OPENQASM 2.0; qreg q[1]; frobnicate q[0];")
    mid = parse_qasm("OPENQASM 2.0; qreg q[1]; creg c[1]; measure q[0] -> c[0]; x q[0];")
    with pytest.raises(ValueError):
        mid.gate_ops()


def test_parameters_bind_in_sorted_order():
    th = [Parameter(f"t[{i}]") for i in (10, 2, 0)]
    c = Circuit(2)
    c.rz(th[0], 0); c.rx(2 * th[1] + 0.5, 1); c.ry(-th[2], 0); c.cx(0, 1)
    assert [p.name for p in c.parameters] == ["t[0]", "t[2]", "t[10]"]
    b = c.bind_parameters([0.1, 0.2, 0.3])
    assert b.ops[0][2] == (0.3,) and abs(b.ops[1][2][0] - 0.9) < 1e-15 and b.ops[2][2] == (-0.1,)
    with pytest.raises(ValueError):
        c.bind_parameters([0.1])


def test_duck_typed_qiskit_circuit():
    class Op:
        def __init__(self, name, params=()):
            self.name, self.params, self.condition = name, list(params), None

    class Inst:
        def __init__(self, op, qubits):
            self.operation, self.qubits = op, qubits

    class QC:
        num_qubits = 3
        name = "duck"

        def __init__(self):
            self.qubits = [object() for _ in range(3)]
            q = self.qubits
            self.data = [Inst(Op("rz", [0.25]), [q[2]]), Inst(Op("cx"), [q[2], q[0]]), Inst(Op("barrier"), q),
                         Inst(Op("measure"), [q[0]])]

    c = from_any(QC())
    assert c.gate_ops() == [("rz", (2,), (0.25,)), ("cx", (2, 0), ())]


def test_observable_inputs_and_masks():
    o = PauliObservable([("XIZY", 0.5), ("IIII", -1.0)])
    x, z, c = o.masks()
    assert (int(x[0]), int(z[0])) == (0b1001, 0b0011) and c.tolist() == [0.5, -1.0]

    class Paulis:
        def to_labels(self):
            return ["ZZ", "XI"]

    class SPO:
        paulis = Paulis()
        coeffs = np.array([1.0 + 0j, 0.25])

    from ml_qem_b200 import observable
    assert observable.from_any(SPO()).terms == [("ZZ", 1.0 + 0j), ("XI", 0.25 + 0j)]
    with pytest.raises(ValueError):
        PauliObservable([("ZQ", 1.0)])


def test_encode_batch_layout():
    c1 = Circuit(3); c1.rz(0.5, 1); c1.cx(1, 2); c1.u3(0.1, 0.2, 0.3, 0); c1.measure_all()
    c2 = Circuit(2); c2.sx(0)
    fb = engine.encode_batch([c1, c2], [[[("ZII", 1.0)], [("XXI", 2.0), ("III", 1.0)]], [[("IZ", 1.0)]]])
    assert fb.n_circuits == 2 and fb.n_observables == 3
    assert fb.op_offsets.tolist() == [0, 3, 4] and fb.obs_offsets.tolist() == [0, 2, 3]
    assert fb.term_offsets.tolist() == [0, 1, 3, 4]
    assert fb.ops["opcode"].tolist() == [13, 32, 16, 9] and fb.ops["param_idx"].tolist() == [0, 1, 1, 4]
    assert fb.params.tolist() == [0.5, 0.1, 0.2, 0.3]
    sub = fb.select([1])
    assert sub.n_circuits == 1 and sub.ops["opcode"].tolist() == [9] and sub.term_z.tolist() == [1]
    with pytest.raises(ValueError):
        engine.encode_batch([c2], [[[("ZZZ", 1.0)]]])


def test_noise_table_structure_detection():
    lima = backends.fake_lima()
    nm = noise.from_backend(lima)
    t = nm.to_table()
    kinds = {(int(o), int(a), int(b)): int(k) for o, a, b, k in zip(t["opcode"], t["q0"], t["q1"], t["kind"])}
    assert kinds[(32, 0, 1)] == noise.NOISE_RELAX2 and kinds[(9, 0, 255)] == noise.NOISE_DENSE1
    assert (13, 0, 255) not in kinds  # rz carries no error
    coh, _ = noise.add_coherent_noise(lima, theta=0.1, seed=0)
    t2 = coh.to_table()
    assert set(int(k) for o, k in zip(t2["opcode"], t2["kind"]) if o == 32) == {noise.NOISE_DENSE2}
    # product PTMs == oracle superoperators (independent implementations of the same model)
    from oracle import noise_model as onm
    import helpers
    om = onm.from_backend(helpers.golden("backends.json")["fakelima"])
    for (name, qubits), r in nm.local.items():
        s = om.get(name, qubits)
        k = len(qubits)
        ps = ptm.pauli_basis(k)
        d = 2 ** k
        ref = np.zeros_like(r)
        for j in range(4 ** k):
            out = (s @ ps[j].T.reshape(-1)).reshape(d, d).T  # column-stacked vec -> matrix
            for i in range(4 ** k):
                ref[i, j] = np.real(np.trace(ps[i] @ out)) / d
        assert np.max(np.abs(ref - r)) < 1e-15, (name, qubits)


def test_estimator_validation_without_gpu():
    from ml_qem_b200.estimator import B200Estimator
    est = B200Estimator()
    c = Circuit(2); c.h(0)
    with pytest.raises(ValueError, match="number of circuits"):
        est.run([c, c], ["ZZ"])
    with pytest.raises(ValueError, match="number of qubits"):
        est.run([c], ["ZZZ"])
    with pytest.raises(ValueError, match="number of values"):
        est.run([c], ["ZZ"], [[0.1]])
