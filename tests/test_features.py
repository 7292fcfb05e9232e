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


def test_encode_data_flat_equals_encode_data():
    """Batched feature rows from the flat gate stream == the reference-pinned per-circuit encoder."""
    from ml_qem_b200 import engine, families as F

    lima = backends.fake_lima()
    props = FT.backend_properties_v1(lima)
    rng = np.random.default_rng(3)
    circs = [F.tfim_circuit(4, 1 + i % 4, float(rng.uniform(0, 1)), layout=[0, 1, 3, 4], num_physical=5, basis="XYZ"[i % 3]) for i in range(7)]
    circs.append(F.random_basis_circuit(5, 60, rng, lima.coupling_map))
    c = Circuit(5); c.append("rz", (1,), (2 * np.pi,)); c.append("rz", (2,), (-2 * np.pi,)); c.append("rz", (0,), (7.0,)); c.append("reset", (3,))
    circs.append(c)  # angles on the outer bin edges and outside the histogram range
    n = len(circs)
    ideal = rng.uniform(-1, 1, size=(n, 4)).tolist()
    noisy = rng.uniform(-1, 1, size=(n, 4)).tolist()
    bases = [[1, 0, 0, 0] * 4 for _ in range(n)]
    obs = [F.single_z_observables([0, 1, 3, 4], 5)] * n
    fb = engine.encode_batch(circs, obs)
    X, y = FT.encode_data(circs, props, ideal, noisy, 4, meas_bases=bases)
    Xf, yf = FT.encode_data_flat(fb, props, ideal, noisy, 4, meas_bases=bases)
    assert Xf.shape == X.shape and torch.allclose(Xf, X, atol=1e-7) and torch.equal(yf, y)
    X1, _ = FT.encode_data(circs, props, ideal, noisy, 4)
    X1f, _ = FT.encode_data_flat(fb, props, torch.tensor(ideal), torch.tensor(noisy, dtype=torch.float64), 4)
    assert torch.allclose(X1f, X1, atol=1e-7)


def test_sharded_dataset_resume(tmp_path):
    """Chunked generation: one binary shard per chunk, a crashed run resumes at the first missing chunk."""
    from ml_qem_b200 import dataset, families as F

    rng = np.random.default_rng(0)
    circs = [F.tfim_circuit(3, 1 + i % 3, float(rng.uniform(0, 1))) for i in range(11)]
    obs = [F.single_z_observables([0, 1, 2], 3)] * len(circs)
    calls = []

    def evaluate(fb):
        calls.append(fb.n_circuits)
        if len(calls) == 3 and crash[0]:
            raise RuntimeError("simulated crash")
        v = np.arange(fb.n_observables, dtype=float) + 100 * len(calls)
        return v, -v

    crash = [True]
    out = str(tmp_path / "ds")
    with pytest.raises(RuntimeError):
        dataset.generate(circs, obs, evaluate, out, chunk_size=4)
    assert sorted(f for f in os.listdir(out) if f.startswith("chunk_")) == ["chunk_00000.npz", "chunk_00001.npz"]
    crash[0] = False
    calls.clear()
    m = dataset.generate(circs, obs, evaluate, out, chunk_size=4)
    assert calls == [3] and [s["reused"] for s in m["shards"]] == [True, True, False]
    d = dataset.load(out)
    assert d["ideal"].shape == (33,) and np.array_equal(d["noisy"], -d["ideal"]) and d["obs_offsets"][-1] == 33
    fb1 = dataset.shard_batch(out, 1)
    assert fb1.n_circuits == 4 and fb1.n_observables == 12
    t = dataset.load(out, as_torch=True)
    assert t["ideal"].dtype == torch.float64
