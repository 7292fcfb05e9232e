#!/usr/bin/env python
"""Regenerates tests/golden/*.json from the read-only reference checkout.

Runs only in the build container (needs /root/reference); the GPU box uses the committed
JSON files.  Nothing here imports qiskit: the calibration pickle is read with a stub
Unpickler and notebook outputs are parsed as text.

Sources (all under /root/reference):
  docs/tutorials/device_params/fakebackends_properties_record.json  (pickle of BackendProperties)
  docs/demos/fake_backend_info.ipynb   cells 4,5,8,9,10,11 (Aer noise-model dumps / infidelities)
  docs/tutorials/data/mbd_datasets2/theta_0.05pi/val/step_{1,2}.json (QASM + 10k-shot values)
  docs/tutorials/h2-hamiltonian-qubit-params.txt
"""
import json
import os
import pickle
import re
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, s):
        self.__dict__["state"] = s


class _U(pickle.Unpickler):
    def find_class(self, mod, name):
        if mod.startswith("qiskit"):
            return type(name, (_Stub,), {"__module__": mod})
        return super().find_class(mod, name)


def backends():
    path = os.path.join(REF, "docs/tutorials/device_params/fakebackends_properties_record.json")
    with open(path, "rb") as f:
        rec = _U(f).load()
    out = {}
    for key in ("fakelima", "fakebelem", "fakemontreal"):
        st = rec[key].state
        qubits = []
        for q in st["qubits"]:
            qubits.append({n.state["name"]: {"value": n.state["value"], "unit": n.state["unit"]} for n in q})
        gates = []
        for g in st["gates"]:
            gs = g.state
            params = {n.state["name"]: {"value": n.state["value"], "unit": n.state["unit"]} for n in gs["parameters"]}
            gates.append({"gate": gs["gate"], "qubits": list(gs["qubits"]), "parameters": params})
        out[key] = {
            "backend_name": st["backend_name"],
            "backend_version": st["backend_version"],
            "qubits": qubits,
            "gates": gates,
        }
    return out


def _nb_cells():
    nb = json.load(open(os.path.join(REF, "docs/demos/fake_backend_info.ipynb")))
    return nb["cells"]


def _out_text(cell, kind):
    for o in cell["outputs"]:
        if kind == "result" and o.get("output_type") == "execute_result":
            return "".join(o["data"]["text/plain"])
        if kind == "stdout" and o.get("name") == "stdout":
            return "".join(o["text"])
    raise KeyError(kind)


def aer_noise_dump():
    cells = _nb_cells()
    txt = _out_text(cells[4], "result")
    errs = eval(txt, {"array": np.array})  # python repr of to_dict()['errors'] (trusted: parsed as data)
    clean = []
    for e in errs:
        if e["type"] != "qerror":
            clean.append({"type": e["type"], "operations": e["operations"],
                          "gate_qubits": [list(q) for q in e["gate_qubits"]],
                          "probabilities": np.asarray(e["probabilities"]).tolist()})
            continue
        ins = []
        for circ in e["instructions"]:
            c2 = []
            for op in circ:
                p = []
                for m in op.get("params", []):
                    m = np.asarray(m)
                    if m.dtype.kind in "cf":
                        p.append({"re": np.real(m).tolist(), "im": np.imag(m).tolist()})
                    else:
                        p.append(m.tolist())
                c2.append({"name": op["name"], "qubits": list(op["qubits"]), "params": p})
            ins.append(c2)
        clean.append({
            "type": e["type"],
            "operations": e["operations"],
            "gate_qubits": [list(q) for q in e["gate_qubits"]],
            "probabilities": [float(x) for x in e["probabilities"]],
            "instructions": ins,
        })
    return clean


def kats():
    cells = _nb_cells()

    def three(cell, skip=0):
        lines = [l for l in _out_text(cell, "stdout").splitlines() if re.match(r"^\d\S* \d\S*$", l.strip())]
        return [[float(x) for x in l.split()] for l in lines]

    thetas_txt = _out_text(cells[10], "stdout")
    m = re.search(r"thetas \[(.*?)\]", thetas_txt, re.S)
    thetas = [float(x) for x in m.group(1).split()]
    return {
        "source": "docs/demos/fake_backend_info.ipynb cells 8-11 stdout (mean, std/len of 1-average_gate_fidelity)",
        "lima_infidelity_cx_x_sx": three(cells[8]),
        "belem_infidelity_cx_x_sx": three(cells[9]),
        "lima_coherent_infidelity_cx_x_sx": three(cells[10]),
        "belem_coherent_infidelity_cx_x_sx": three(cells[11]),
        "coherent_thetas_8digits": thetas,
        "lima_readout": _out_text(cells[5], "result"),
    }


def mbd_sample(per_file=40):
    out = []
    for step in (1, 2):
        rel = f"docs/tutorials/data/mbd_datasets2/theta_0.05pi/val/step_{step}.json"
        d = json.load(open(os.path.join(REF, rel)))
        for i, e in enumerate(d[:per_file]):
            out.append({
                "source": f"{rel}[{i}]",
                "qasm": e["circuit"],
                "ideal_exp_value": e["ideal_exp_value"],
                "noisy_exp_values": e["noisy_exp_values"],
            })
    return out


def h2():
    txt = open(os.path.join(REF, "docs/tutorials/h2-hamiltonian-qubit-params.txt")).read()
    out = []
    for blk in txt.strip().split("\n\n"):
        lines = blk.strip().splitlines()
        dist = float(lines[0].split()[0])
        fci = float(lines[1].split("=")[1])
        terms = []
        for l in lines[2:]:
            m = re.match(r"\s*([-\d.eE+]+)\s*\[([^\]]*)\]?", l)
            terms.append([float(m.group(1)), m.group(2).split()])
        out.append({"distance_A": dist, "fci": fci, "terms": terms})
    return out


def main():
    json.dump(backends(), open(os.path.join(OUT, "backends.json"), "w"), indent=0)
    json.dump(aer_noise_dump(), open(os.path.join(OUT, "aer_noise_lima.json"), "w"))
    json.dump(kats(), open(os.path.join(OUT, "kats.json"), "w"), indent=1)
    json.dump(mbd_sample(), open(os.path.join(OUT, "mbd_sample.json"), "w"))
    json.dump(h2(), open(os.path.join(OUT, "h2.json"), "w"), indent=0)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
