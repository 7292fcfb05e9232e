"""Backend calibration data (T1/T2/readout, per-gate error and length) in a plain container.

The reference gets these from Qiskit fake backends -- ``FakeLima()`` in
tests/data/generators/test_exp_val_generator.py:17, ``backend.properties()`` in
blackwater/data/utils.py:139-175 -- and feeds them to ``AerSimulator.from_backend``
(blackwater/data/utils.py:427).  ``data/fake_backends.json`` holds the ibmq_lima / ibmq_belem /
ibmq_montreal snapshots the reference ships under docs/tutorials/device_params/ (converted by
tests/golden/make_golden.py); FakeGuadalupe's calibration is not in the reference tree, so
``synthetic_heavy_hex_chain`` samples a look-alike table from the Montreal snapshot.
"""
import json
import math
import os

import numpy as np

_UNIT = {"s": 1.0, "ms": 1e-3, "us": 1e-6, "µs": 1e-6, "ns": 1e-9, "ps": 1e-12, "": 1.0}
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "fake_backends.json")


def _val(entry):
    if isinstance(entry, dict):
        return float(entry["value"]) * _UNIT.get(entry.get("unit", ""), 1.0)
    return float(entry)


class BackendProps:
    """t1/t2 in seconds per qubit; gates[(name, qubits)] = (gate_error, gate_length_seconds)."""

    def __init__(self, name, t1, t2, gates, readout=None, coupling_map=None, basis_gates=None):
        self.name = name
        self.t1 = [float(x) for x in t1]
        self.t2 = [float(x) for x in t2]
        self.num_qubits = len(self.t1)
        self.gates = {(g, tuple(q)): (None if e is None else float(e), float(l)) for (g, q), (e, l) in gates.items()}
        self.readout = readout or {}
        if coupling_map is None:
            coupling_map = sorted({q for (_, q) in self.gates if len(q) == 2})
        self.coupling_map = [tuple(p) for p in coupling_map]
        self.basis_gates = basis_gates or sorted({g for (g, _) in self.gates})

    # -- constructors
    @classmethod
    def from_dict(cls, d, name=None):
        """BackendProperties.to_dict() layout (docs/demos/fake_backend_info.ipynb:51)."""
        t1, t2, ro = [], [], {}
        for i, q in enumerate(d["qubits"]):
            if isinstance(q, list):  # raw to_dict(): list of {name, value, unit}
                q = {e["name"]: e for e in q}
            a = _val(q["T1"]) if "T1" in q else math.inf
            t1.append(a)
            t2.append(_val(q["T2"]) if "T2" in q else 2 * a)
            if "prob_meas1_prep0" in q and "prob_meas0_prep1" in q:
                ro[i] = (_val(q["prob_meas0_prep1"]), _val(q["prob_meas1_prep0"]))
        gates = {}
        for g in d["gates"]:
            par = g["parameters"]
            if isinstance(par, list):
                par = {e["name"]: e for e in par}
            err = _val(par["gate_error"]) if "gate_error" in par else None
            length = _val(par["gate_length"]) if "gate_length" in par else 0.0
            gates[(g["gate"], tuple(g["qubits"]))] = (err, length)
        return cls(name or d.get("backend_name", "backend"), t1, t2, gates, ro)

    @classmethod
    def from_v1_dict(cls, d):
        """Output of blackwater.data.utils.get_backend_properties_v1 (utils.py:139-175):
        t1/t2 in seconds, gate_length in ns, gate keys 'cx_0_1'."""
        n = d["num_qubits"]
        qp = {int(k): v for k, v in d["qubits_props"].items()}
        t1 = [qp[i]["t1"] for i in range(n)]
        t2 = [qp[i]["t2"] for i in range(n)]
        gates = {}
        for key, g in d["gate_props"].items():
            parts = str(key).split("_")
            gates[(parts[0], tuple(int(x) for x in parts[1:]))] = (g.get("gate_error", 0.0), g.get("gate_length", 0.0) * 1e-9)
        return cls(d.get("name", "backend"), t1, t2, gates)

    @classmethod
    def from_backend(cls, backend):
        """Anything with ``properties().to_dict()`` (Qiskit BackendV1 / fake backends)."""
        if isinstance(backend, BackendProps):
            return backend
        if isinstance(backend, dict):
            return cls.from_v1_dict(backend) if "qubits_props" in backend else cls.from_dict(backend)
        props = backend.properties() if callable(getattr(backend, "properties", None)) else backend
        out = cls.from_dict(props.to_dict())
        conf = backend.configuration() if callable(getattr(backend, "configuration", None)) else None
        if conf is not None and getattr(conf, "coupling_map", None):
            out.coupling_map = [tuple(p) for p in conf.coupling_map]
        return out

    def to_dict(self):
        qubits = []
        for i in range(self.num_qubits):
            q = {"T1": {"value": self.t1[i], "unit": "s"}, "T2": {"value": self.t2[i], "unit": "s"}}
            if i in self.readout:
                q["prob_meas0_prep1"] = {"value": self.readout[i][0], "unit": ""}
                q["prob_meas1_prep0"] = {"value": self.readout[i][1], "unit": ""}
            qubits.append(q)
        gates = []
        for (g, qs), (err, length) in self.gates.items():
            par = {"gate_length": {"value": length, "unit": "s"}}
            if err is not None:
                par["gate_error"] = {"value": err, "unit": ""}
            gates.append({"gate": g, "qubits": list(qs), "parameters": par})
        return {"backend_name": self.name, "qubits": qubits, "gates": gates}


def _load(key):
    with open(_DATA) as f:
        return BackendProps.from_dict(json.load(f)[key])


def fake_lima():
    return _load("fakelima")


def fake_belem():
    return _load("fakebelem")


def fake_montreal():
    return _load("fakemontreal")


def synthetic_chain(n_qubits, seed=0, name=None, donor=None):
    """Linear-chain backend with basis {id, rz, sx, x, cx, reset}; T1/T2/errors/lengths are drawn
    (with replacement) from the donor snapshot's entries (default: ibmq_montreal)."""
    donor = donor or fake_montreal()
    rng = np.random.default_rng(seed)
    qi = rng.integers(0, donor.num_qubits, size=n_qubits)
    t1 = [donor.t1[i] for i in qi]
    t2 = [donor.t2[i] for i in qi]
    one = {g: [v for (name_, q), v in donor.gates.items() if name_ == g] for g in ("id", "sx", "x", "rz", "reset")}
    two = [v for (name_, q), v in donor.gates.items() if name_ == "cx"]
    gates = {}
    for q in range(n_qubits):
        k = int(rng.integers(0, len(one["sx"])))
        for g in ("id", "sx", "x"):
            gates[(g, (q,))] = one["sx"][k]
        gates[("rz", (q,))] = (0.0, 0.0)
        if one["reset"]:
            gates[("reset", (q,))] = one["reset"][int(rng.integers(0, len(one["reset"])))]
    for q in range(n_qubits - 1):
        err, length = two[int(rng.integers(0, len(two)))]
        gates[("cx", (q, q + 1))] = (err, length)
        gates[("cx", (q + 1, q))] = (err, length + 35.5e-9)
    ro = {q: (0.02, 0.01) for q in range(n_qubits)}
    return BackendProps(name or f"synthetic_chain_{n_qubits}", t1, t2, gates, ro)
