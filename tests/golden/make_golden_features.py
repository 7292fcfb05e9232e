"""Generates tests/golden/encode_data.json by running the REFERENCE's own encode_data
(/root/reference/blackwater/library/learning/mlp.py:149-203) on duck-typed circuits.  The module
imports qiskit / matplotlib / seaborn at the top for unrelated code; those imports are satisfied
with empty stub modules -- encode_data itself only needs torch, numpy, ``circuit.count_ops()`` and
``circuit.data``.  Also samples ``circuit_graph`` entries of the reference's stored datasets into
tests/golden/graph_sample.json.  Run in the build container (needs /root/reference):
    python tests/golden/make_golden_features.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference_mlp():
    for name in ("qiskit", "qiskit.circuit", "qiskit.circuit.random", "matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["qiskit"].QuantumCircuit = object
    spec = importlib.util.spec_from_file_location("ref_mlp", os.path.join(REF, "blackwater/library/learning/mlp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Instr:
    def __init__(self, name, params):
        self.name, self.params = name, list(params)


class DuckCircuit:
    """What encode_data touches of a QuantumCircuit: .data triples and count_ops()."""

    def __init__(self, ops):
        self.data = [(_Instr(n, p), list(q), []) for n, q, p in ops]

    def count_ops(self):
        out = {}
        for ins, _, _ in self.data:
            out[ins.name] = out.get(ins.name, 0) + 1
        return out


def main():
    from ml_qem_b200 import backends, families as F
    from ml_qem_b200.features import backend_properties_v1

    ref = import_reference_mlp()
    lima = backends.fake_lima()
    props = backend_properties_v1(lima, gates_set=["id", "rz", "cx", "reset", "sx", "x"])  # unsorted on purpose
    rng = np.random.default_rng(0)
    circs = [F.tfim_circuit(4, 1 + i % 4, float(rng.uniform(0, 1)), basis="XYZ"[i % 3], layout=[0, 1, 3, 4], num_physical=5)
             for i in range(6)]
    circs += [F.random_basis_circuit(5, 40, rng, lima.coupling_map) for _ in range(3)]
    ops = [[(n, list(q), [float(x) for x in p]) for n, q, p in c.ops] for c in circs]
    noisy = [[float(x) for x in rng.uniform(-1, 1, 4)] for _ in circs]
    ideal = [[float(x) for x in rng.uniform(-1, 1, 4)] for _ in circs]
    bases = [[float(b) for b in np.eye(3)[i % 3]] for i in range(len(circs))]
    X, y = ref.encode_data([DuckCircuit(o) for o in ops], props, ideal, noisy, 4, meas_bases=bases)
    X1, y1 = ref.encode_data([DuckCircuit(o) for o in ops], props, [v[0] for v in ideal], [[v[0]] for v in noisy], 1)
    out = {"properties": {k: (v if k not in ("qubits_props",) else {str(i): q for i, q in v.items()}) for k, v in props.items()},
           "qubits_props_int_keys": True, "ops": ops, "noisy": noisy, "ideal": ideal, "meas_bases": bases,
           "X": X.tolist(), "y": y.tolist(), "X_single": X1.tolist(), "y_single": y1.tolist()}
    json.dump(out, open(os.path.join(HERE, "encode_data.json"), "w"))
    print("encode_data.json", tuple(X.shape), tuple(X1.shape))

    # circuit_graph samples from the stored datasets (QASM + the graph the reference computed)
    src = os.path.join(REF, "docs/tutorials/data/mbd_datasets2/theta_0.05pi/val/step_1.json")
    entries = json.load(open(src))
    pick = [entries[i] for i in (0, 7, 42)]
    json.dump(pick, open(os.path.join(HERE, "graph_sample.json"), "w"))
    print("graph_sample.json", len(pick))


if __name__ == "__main__":
    main()
