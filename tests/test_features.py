"""CPU: feature encoders against the reference.

* encode_data: golden X, y produced by the reference's own function
  (tests/golden/make_golden_features.py imports blackwater/library/learning/mlp.py with stubbed
  plotting/qiskit modules), compared to float32 round-off;
* circuit_to_graph_data_json: the ``circuit_graph`` the reference stored next to the QASM text in
  docs/tutorials/data/mbd_datasets2 (node features exactly, wire edges as multisets);
* ExpValueEntry JSON round trip and tensors."""
import json
import os

import numpy as np
import pytest
import torch

import helpers
from ml_qem_b200 import backends, features as FT
from ml_qem_b200.circuit import Circuit


def _circuits(g):
    out = []
    for ops in g["ops"]:
        c = Circuit(5)
        for n, q, p in ops:
            c.append(n, q, p)
        out.append(c)
    return out


def test_encode_data_matches_reference_function():
    g = helpers.golden("encode_data.json")
    props = g["properties"]
    props["qubits_props"] = {int(k): v for k, v in props["qubits_props"].items()}  # the reference uses int keys
    circs = _circuits(g)
    X, y = FT.encode_data(circs, props, g["ideal"], g["noisy"], 4, meas_bases=g["meas_bases"])
    assert X.shape == (9, 61) and X.dtype == torch.float32
    assert torch.allclose(X, torch.tensor(g["X"]), atol=1e-7, rtol=0)
    assert torch.allclose(y, torch.tensor(g["y"]), atol=0, rtol=0)
    X1, y1 = FT.encode_data(circs, props, [v[0] for v in g["ideal"]], [[v[0]] for v in g["noisy"]], 1)
    assert torch.allclose(X1, torch.tensor(g["X_single"]), atol=1e-7, rtol=0)
    assert torch.allclose(y1, torch.tensor(g["y_single"]))
    # qubit 0 is left out of the t1/t2/readout means with integer keys (reference quirk) ...
    t1 = [props["qubits_props"][i]["t1"] for i in range(5)]
    assert abs(float(X[0, 6]) - 100 * np.mean(t1[1:])) < 1e-6
    # ... and included when the dict went through JSON (string keys)
    props_s = dict(props, qubits_props={str(k): v for k, v in props["qubits_props"].items()})
    Xs, _ = FT.encode_data(circs, props_s, g["ideal"], g["noisy"], 4, meas_bases=g["meas_bases"])
    assert abs(float(Xs[0, 6]) - 100 * np.mean(t1)) < 1e-6
    # QASM text input gives the same features as Circuit input
    qasm = "OPENQASM 2.0;\ninclude \"qelib1.inc\";\nqreg q[5];\nrz(0.3) q[1];\nsx q[1];\ncx q[1],q[0];\nrz(-2.0) q[0];\n"
    c = Circuit(5); c.append("rz", (1,), (0.3,)); c.append("sx", (1,)); c.append("cx", (1, 0)); c.append("rz", (0,), (-2.0,))
    Xa, _ = FT.encode_data([qasm], props, [[0.0] * 4], [[0.1] * 4], 4)
    Xb, _ = FT.encode_data([c], props, [[0.0] * 4], [[0.1] * 4], 4)
    assert torch.equal(Xa, Xb)


def _lima_properties(sample):
    raw = helpers.golden("backends.json")["fakelima"]
    ro = []
    for q in raw["qubits"]:
        q = {e["name"]: e for e in q} if isinstance(q, list) else q
        ro.append(float(q["readout_error"]["value"]) if "readout_error" in q else None)
    lima = backends.fake_lima()
    if any(r is None for r in ro):
        ro = None
    # the one-hot order of the stored dataset (hash order of a Python set at generation time):
    # recover it from the first sample
    nq, nc, ins = FT.qasm_instructions(sample["circuit"])
    order = {}
    for (name, *_), fv in zip(ins, sample["circuit_graph"]["nodes"]["DAGOpNode"]):
        order[name] = int(np.argmax(fv[3:11]))
    names = [None] * 6
    for name, pos in order.items():
        if pos < 6:
            names[pos] = name
    rest = [g for g in ("cx", "id", "reset", "rz", "sx", "x") if g not in names]
    names = [n if n is not None else rest.pop(0) for n in names]
    return FT.backend_properties_v1(lima, gates_set=names, readout_error=ro)


def test_graph_encoder_matches_stored_dataset_graphs():
    samples = helpers.golden("graph_sample.json")
    props = _lima_properties(samples[0])
    for s in samples:
        got = FT.circuit_to_graph_data_json(s["circuit"], props, use_gate_features=True, use_qubit_features=True)
        want = s["circuit_graph"]
        for kind in ("DAGOpNode", "DAGInNode", "DAGOutNode"):
            a, b = np.array(got["nodes"][kind], dtype=float), np.array(want["nodes"][kind], dtype=float)
            assert a.shape == b.shape, kind
            assert np.allclose(a, b, rtol=1e-12, atol=1e-15), kind
        assert set(got["edges"]) == set(want["edges"])
        for key in want["edges"]:
            def rows(e):
                idx = np.array(e["edge_index"]).T.tolist()
                return sorted((tuple(i), tuple(np.round(a, 15))) for i, a in zip(idx, e["edge_attr"]))
            assert rows(got["edges"][key]) == rows(want["edges"][key]), key


def test_exp_value_entry_round_trip(tmp_path):
    samples = helpers.golden("graph_sample.json")
    entries = [FT.ExpValueEntry.from_json(dict(s)) for s in samples]
    path = os.path.join(tmp_path, "entries.json")
    FT.save_entries(path, entries)
    back = FT.load_entries(path)
    assert len(back) == 3 and back[0].to_dict() == entries[0].to_dict()
    assert json.load(open(path))[0].keys() == samples[0].keys()
    e = FT.ExpValueEntry(circuit_graph=samples[0]["circuit_graph"], observable=[[1.0, 0.0]], ideal_exp_value=0.5,
                         noisy_exp_values=[0.4, 0.3], circuit_depth=7)
    t = e.to_tensors()
    assert t["x"].shape[1] == 22 and t["edge_index"].shape[0] == 2 and t["edge_attr"].shape[1] == 3
    assert t["y"].shape == (1, 1) and float(t["noisy_1"]) == pytest.approx(0.3) and float(t["circuit_depth"]) == 7.0
