"""Feature encoders of the step right after the expectation-value hot path:
what the reference builds per circuit with Qiskit objects, restated on plain data so that the
engine's outputs flow into the sklearn / MLP / GNN models without qiskit or torch_geometric.

  backend_properties_v1       blackwater/data/utils.py:139-175 (get_backend_properties_v1)
  encode_data                 blackwater/library/learning/mlp.py:149-203 (circuit-level vector:
                              8 backend means x100, gate counts x0.01, 40 rotation-angle bins
                              x0.01, noisy <O> per qubit, optional measurement-basis one-hot)
  circuit_to_graph_data_json  blackwater/data/utils.py:198-389 (DAG node features + wire edges)
  ExpValueEntry               blackwater/data/generators/exp_val.py:32-61 (JSON schema of the
                              datasets under docs/tutorials/data, plus the QASM ``circuit`` field)

``encode_data`` is checked against the reference function itself (imported with stubbed qiskit /
plotting modules by tests/golden/make_golden_features.py); the graph encoder against the
``circuit_graph`` fields stored in the reference's datasets.
"""
import json
import re
from dataclasses import dataclass, field

import numpy as np

from . import circuit as circuit_mod
from .backends import BackendProps
from .gateset import IGNORED, canonical


# ----------------------------------------------------------------------------- backend dict
def backend_properties_v1(backend, gates_set=None, readout_error=None):
    """get_backend_properties_v1 layout: t1/t2 in seconds, gate_length in ns (the raw calibration
    value), gate keys 'cx_0_1'.  ``gates_set`` fixes the one-hot order (the reference takes it from
    a Python set, i.e. hash order -- utils.py:158; encode_data sorts it, the graph encoder does
    not); default: sorted."""
    b = BackendProps.from_backend(backend)
    names = sorted({g for (g, _) in b.gates})
    ro = readout_error
    if ro is None:
        ro = [0.5 * sum(b.readout[i]) if i in b.readout else 0.0 for i in range(b.num_qubits)]
    return {
        "name": b.name,
        "gates_set": list(gates_set) if gates_set is not None else names,
        "num_qubits": b.num_qubits,
        "qubits_props": {i: {"index": i, "t1": b.t1[i], "t2": b.t2[i], "readout_error": float(ro[i])}
                         for i in range(b.num_qubits)},
        "gate_props": {f"{g}_{'_'.join(str(q) for q in qs)}": {"index": f"{g}_{'_'.join(str(q) for q in qs)}",
                                                                 "gate_error": 0.0 if err is None else err,
                                                                 "gate_length": length * 1e9}
                       for (g, qs), (err, length) in b.gates.items()},
    }


def _recursive_collect(d, parent_key, target1, target2, out):
    for key, val in d.items():
        if isinstance(val, dict):
            _recursive_collect(val, key, target1, target2, out)
        elif parent_key and target1 in str(parent_key) and key == target2:
            out.append(val)   # NB truthiness: an integer parent key 0 (qubit 0) never matches -- the
    return out                # reference's behaviour, kept for feature parity


def _mean_of(properties, target1, target2):
    # mlp.py:136-145 (recursive_dict_loop): values of key target2 in sub-dicts whose key contains
    # target1; "out or 0." when nothing matches
    out = _recursive_collect(properties, None, target1, target2, [])
    return float(np.mean(out)) if out else 0.0


# ----------------------------------------------------------------------------- instruction lists
def _instructions(circ):
    """-> (num_qubits, [(name, qubits, params)]) keeping every instruction count_ops() would see."""
    if isinstance(circ, str):
        n, _, ins = qasm_instructions(circ)
        return n, [(name, qs, ps) for name, qs, _, ps in ins]
    if isinstance(circ, circuit_mod.Circuit):
        return circ.num_qubits, [(name, qs, tuple(float(p) for p in ps)) for name, qs, ps in circ.ops]
    # duck-typed qiskit circuit
    qindex = {q: i for i, q in enumerate(circ.qubits)} if hasattr(circ, "qubits") else {}
    ins = []
    for item in circ.data:
        op, qargs = (item.operation, item.qubits) if hasattr(item, "operation") else (item[0], item[1])
        ins.append((op.name, tuple(qindex.get(q, getattr(q, "index", q)) for q in qargs),
                    tuple(float(p) for p in getattr(op, "params", ()) if isinstance(p, (int, float)))))
    return circ.num_qubits, ins


_QREG = re.compile(r"^(qreg|creg)\s+([A-Za-z_][A-Za-z0-9_]*)\s*\[(\d+)\]$")


def qasm_instructions(text):
    """Flat OpenQASM-2 (the ``circuit`` field of the reference's datasets) -> (n_qubits, n_clbits,
    [(name, qubits, clbits, params)]) INCLUDING barriers and measurements, in program order."""
    text = re.sub(r"//[^\n]*", "", text)
    regs, sizes = {}, {"qreg": 0, "creg": 0}
    out = []
    for s in (x.strip() for x in text.split(";")):
        if not s or s.startswith("OPENQASM") or s.startswith("include"):
            continue
        m = _QREG.match(s)
        if m:
            regs[m.group(2)] = (m.group(1), sizes[m.group(1)], int(m.group(3)))
            sizes[m.group(1)] += int(m.group(3))
            continue

        def bits(arg):
            arg = arg.strip()
            mm = re.match(r"^([A-Za-z_][A-Za-z0-9_]*)\s*\[(\d+)\]$", arg)
            if mm:
                return [regs[mm.group(1)][1] + int(mm.group(2))]
            _, off, size = regs[arg]
            return list(range(off, off + size))

        if s.startswith("measure"):
            src, dst = s[len("measure"):].split("->")
            for q, c in zip(bits(src), bits(dst)):
                out.append(("measure", (q,), (c,), ()))
            continue
        m = circuit_mod._STMT.match(s)
        if not m:
            raise ValueError(f"QASM: cannot parse statement {s!r}")
        name, pstr, qstr = m.group(1), m.group(2), m.group(3)
        params = tuple(circuit_mod._eval_expr(p, None) for p in circuit_mod._split_args(pstr)) if pstr else ()
        qs = [q for a in circuit_mod._split_args(qstr) for q in bits(a)]
        out.append((name, tuple(qs), (), params))
    return sizes["qreg"], sizes["creg"], out


# ----------------------------------------------------------------------------- encode_data
def encode_data(circuits, properties, ideal_exp_vals, noisy_exp_vals, num_qubits, meas_bases=None):
    """Same arguments and result as the reference's encode_data: (X float32 [n, 8 + |gates_set| + 40
    + num_qubits + |basis|], y float32).  ``circuits``: Circuit / QASM text / qiskit-like objects."""
    import torch

    if isinstance(noisy_exp_vals[0], list) and len(noisy_exp_vals[0]) == 1:
        noisy_exp_vals = [x[0] for x in noisy_exp_vals]
    gates_set = sorted(properties["gates_set"])
    if meas_bases is None:
        meas_bases = [[]]
    vec = [_mean_of(properties, t1, t2) for t1, t2 in (("cx", "gate_error"), ("id", "gate_error"), ("sx", "gate_error"),
                                                       ("x", "gate_error"), ("rz", "gate_error"))]
    # target_key1 = '' matches every (truthy) parent key
    vec += [_mean_of(properties, "", k) for k in ("readout_error", "t1", "t2")]
    vec = torch.tensor(vec) * 100
    bin_size = 0.1 * np.pi
    n_bins = int(np.ceil(4 * np.pi / bin_size))
    bin_edges = np.arange(-2 * np.pi, 2 * np.pi + bin_size, bin_size)
    n = len(circuits)
    nv, ng = len(vec), len(gates_set)
    X = torch.zeros([n, nv + ng + n_bins + num_qubits + len(meas_bases[0])])
    X[:, :nv] = vec[None, :]
    col = {g: i for i, g in enumerate(gates_set)}
    for i, circ in enumerate(circuits):
        _, ins = _instructions(circ)
        counts = np.zeros(ng)
        angles = []
        for name, qs, ps in ins:
            if name in col:
                counts[col[name]] += 1
            if name in ("rx", "ry", "rz") and len(qs) == 1:
                angles.append(float(ps[0]))
        X[i, nv:nv + ng] = torch.tensor(counts) * 0.01
        hist, _ = np.histogram(angles, bins=bin_edges)
        X[i, nv + ng:nv + ng + n_bins] = torch.tensor(hist[:n_bins].tolist() if len(hist) >= n_bins else hist.tolist()) * 0.01
        if num_qubits > 1:
            assert len(noisy_exp_vals[i]) == num_qubits
        X[i, nv + ng + n_bins:nv + ng + n_bins + num_qubits] = torch.tensor(noisy_exp_vals[i])
    if meas_bases != [[]]:
        assert len(meas_bases) == n
        for i, basis in enumerate(meas_bases):
            X[i, nv + ng + n_bins + num_qubits:] = torch.tensor(basis)
    y = torch.tensor(ideal_exp_vals, dtype=torch.float32)
    return X, y


# ----------------------------------------------------------------------------- graph encoder
def circuit_to_graph_data_json(circuit, properties, use_gate_features=False, use_qubit_features=False):
    """DAG view of a circuit as the reference stores it in ``ExpValueEntry.circuit_graph``:
    nodes {DAGOpNode, DAGInNode, DAGOutNode} feature vectors and wire edges with (t1, t2,
    readout_error) attributes.  ``circuit``: OpenQASM-2 text (barriers / measurements kept, as
    circuit_to_dag sees them) or a Circuit.  Node order: op nodes in program order, in/out nodes
    qubits first then clbits; edges are emitted wire by wire (qubit wires only carry attributes,
    utils.py:325-346) -- the same multiset as the reference, whose edge ORDER comes from
    rustworkx iteration."""
    if isinstance(circuit, str):
        nq, nc, ins = qasm_instructions(circuit)
    else:
        nq, plain = _instructions(circuit)
        nc, ins = 0, []
        for name, qs, ps in plain:
            if name == "measure":
                ins.append((name, qs, (nc,), ps))
                nc += 1
            else:
                ins.append((name, qs, (), ps))
    types = list(properties["gates_set"]) + ["barrier", "measure"]
    tmap = {g: i for i, g in enumerate(types)}
    qp = {int(k): v for k, v in properties["qubits_props"].items()}
    op_nodes = []
    last = {("q", q): ("DAGInNode", q) for q in range(nq)}
    last.update({("c", c): ("DAGInNode", nq + c) for c in range(nc)})
    edges = {}

    def add_edge(src, dst, wire):
        if wire[0] != "q":
            return
        key = f"{src[0]}_wire_{dst[0]}"
        e = edges.setdefault(key, {"edge_index": [], "edge_attr": []})
        a = qp[wire[1]]
        e["edge_index"].append([src[1], dst[1]])
        e["edge_attr"].append([a["t1"], a["t2"], a["readout_error"]])

    for name, qs, cs, ps in ins:
        if name != "barrier" and len(qs) > 3:
            raise ValueError("Non barrier gate that has more than 3 qubits.")
        qprops = [qp[q] for q in qs] if name != "barrier" else []
        qprops += [{}] * (3 - len(qprops))
        qfeat = [v.get("t1", 0.0) for v in qprops] + [v.get("t2", 0.0) for v in qprops] + \
                [v.get("readout_error", 0.0) for v in qprops]
        gp = properties["gate_props"].get(f"{name}_{'_'.join(str(q) for q in qs)}", {})
        onehot = [0.0] * len(types)
        onehot[tmap[name]] = 1.0
        pf = [0.0, 0.0, 0.0]
        for i, p in enumerate(ps[:3]):
            pf[i] = float(p)
        fv = pf + onehot
        if use_qubit_features:
            fv += qfeat
        if use_gate_features:
            fv += [gp.get("gate_error", 0.0), gp.get("gate_length", 0.0)]
        idx = len(op_nodes)
        op_nodes.append(fv)
        for w in [("q", q) for q in qs] + [("c", c) for c in cs]:
            add_edge(last[w], ("DAGOpNode", idx), w)
            last[w] = ("DAGOpNode", idx)
    for q in range(nq):
        add_edge(last[("q", q)], ("DAGOutNode", q), ("q", q))
    data = {"nodes": {"DAGOpNode": op_nodes, "DAGInNode": [[0, 0] for _ in range(nq + nc)],
                      "DAGOutNode": [[0, 0] for _ in range(nq + nc)]}, "edges": {}}
    for key, d in edges.items():
        data["edges"][key] = {"edge_index": np.array(d["edge_index"]).T.tolist(), "edge_attr": d["edge_attr"]}
    return data


# ----------------------------------------------------------------------------- dataset entries
@dataclass
class ExpValueEntry:
    """One row of the reference's datasets (exp_val.py:32-61); ``circuit`` (QASM) and ``metadata``
    are the optional extra fields the tutorial notebooks add (loaders/exp_val.py:58-66 pops them)."""
    circuit_graph: dict
    observable: list
    ideal_exp_value: object
    noisy_exp_values: list
    circuit_depth: int = 0
    circuit: str = None
    metadata: dict = field(default=None)

    def to_dict(self):
        d = {"circuit_graph": self.circuit_graph, "observable": self.observable, "ideal_exp_value": self.ideal_exp_value,
             "noisy_exp_values": self.noisy_exp_values, "circuit_depth": self.circuit_depth}
        if self.circuit is not None:
            d["circuit"] = self.circuit
        if self.metadata is not None:
            d["metadata"] = self.metadata
        return d

    @classmethod
    def from_json(cls, dictionary):
        return cls(**dictionary)

    def to_tensors(self):
        """The tensors ``to_pyg_data`` puts into a PyG ``Data`` (exp_val.py:63-89), as a dict."""
        import torch

        key = "DAGOpNode_wire_DAGOpNode"
        g = self.circuit_graph
        out = {"x": torch.tensor(g["nodes"]["DAGOpNode"], dtype=torch.float),
               "edge_index": torch.tensor(g["edges"][key]["edge_index"], dtype=torch.long),
               "edge_attr": torch.tensor(g["edges"][key]["edge_attr"], dtype=torch.float),
               "y": torch.tensor([[self.ideal_exp_value]], dtype=torch.float),
               "observable": torch.tensor([self.observable], dtype=torch.float),
               "circuit_depth": torch.tensor([[self.circuit_depth]], dtype=torch.float)}
        for i, v in enumerate(self.noisy_exp_values):
            out[f"noisy_{i}"] = torch.tensor([[v]], dtype=torch.float)
        return out


def load_entries(path, num_samples=None):
    """Reads a dataset file written by the reference (a JSON list of entry dicts)."""
    with open(path) as f:
        data = json.load(f)
    if num_samples is not None:
        data = data[:num_samples]
    return [ExpValueEntry.from_json(e) for e in data]


def save_entries(path, entries):
    with open(path, "w") as f:
        json.dump([e.to_dict() for e in entries], f)


# ----------------------------------------------------------------------------- batched, from the flat gate stream
def encode_data_flat(batch, properties, ideal_exp_vals, noisy_exp_vals, num_qubits, meas_bases=None, device=None):
    """``encode_data`` for a whole FlatBatch at once, computed from the flat gate stream the engine
    consumes (no per-circuit objects, no Python loop over gates): gate counts by a scatter-add over
    (circuit, opcode), the 40 rotation-angle bins by one ``searchsorted``.  ``noisy_exp_vals`` /
    ``ideal_exp_vals`` may be torch tensors that already live on ``device`` (the engine's
    ``run_dm_into`` output): they are placed into X / y without a host round trip.  Same X, y as
    ``encode_data`` on the same circuits (tests/test_features.py)."""
    import torch

    from .gateset import NAMES, OPCODES

    gates_set = sorted(properties["gates_set"])
    vec = [_mean_of(properties, t1, t2) for t1, t2 in (("cx", "gate_error"), ("id", "gate_error"), ("sx", "gate_error"),
                                                       ("x", "gate_error"), ("rz", "gate_error"))]
    vec += [_mean_of(properties, "", k) for k in ("readout_error", "t1", "t2")]
    n = batch.n_circuits
    nv, ng = len(vec), len(gates_set)
    bin_size = 0.1 * np.pi
    n_bins = int(np.ceil(4 * np.pi / bin_size))
    edges = np.arange(-2 * np.pi, 2 * np.pi + bin_size, bin_size)
    nb = meas_bases if meas_bases is not None else [[]]
    width = nv + ng + n_bins + num_qubits + len(nb[0])
    X = np.zeros((n, width), dtype=np.float32)
    X[:, :nv] = (torch.tensor(vec) * 100).numpy()[None, :]
    ops = batch.ops
    circ_of = np.repeat(np.arange(n), np.diff(batch.op_offsets))
    # gate counts: opcode -> column of the sorted gates_set (names the backend does not list are not counted)
    col_of = np.full(max(NAMES) + 1, -1, dtype=np.int64)
    for j, g in enumerate(gates_set):
        code = OPCODES.get(canonical(g))
        if code is not None:
            col_of[code] = j
    cols = col_of[ops["opcode"]]
    keep = cols >= 0
    counts = np.zeros((n, ng))
    np.add.at(counts, (circ_of[keep], cols[keep]), 1.0)
    X[:, nv:nv + ng] = (torch.tensor(counts) * 0.01).to(torch.float32).numpy()
    # rotation angles of rx / ry / rz
    rot = np.isin(ops["opcode"], [OPCODES["rx"], OPCODES["ry"], OPCODES["rz"]])
    ang = batch.params[ops["param_idx"][rot].astype(np.int64)]
    idx = np.searchsorted(edges, ang, side="right") - 1
    idx[ang == edges[-1]] = len(edges) - 2           # numpy.histogram: the last bin is closed
    ok = (idx >= 0) & (idx < min(n_bins, len(edges) - 1))
    hist = np.zeros((n, n_bins))
    np.add.at(hist, (circ_of[rot][ok], idx[ok]), 1.0)
    X[:, nv + ng:nv + ng + n_bins] = (torch.tensor(hist) * 0.01).to(torch.float32).numpy()
    Xt = torch.from_numpy(X)
    if device is not None:
        Xt = Xt.to(device)
    noisy = noisy_exp_vals if torch.is_tensor(noisy_exp_vals) else torch.tensor(np.asarray(noisy_exp_vals, dtype=np.float64))
    Xt[:, nv + ng + n_bins:nv + ng + n_bins + num_qubits] = noisy.reshape(n, num_qubits).to(Xt.device, torch.float32)
    if meas_bases is not None:
        Xt[:, nv + ng + n_bins + num_qubits:] = torch.tensor(meas_bases, dtype=torch.float32).to(Xt.device)
    ideal = ideal_exp_vals if torch.is_tensor(ideal_exp_vals) else torch.tensor(np.asarray(ideal_exp_vals, dtype=np.float64))
    return Xt, ideal.to(Xt.device, torch.float32)


def graph_tensors_flat(batch, properties, use_gate_features=False, use_qubit_features=False):
    """``circuit_to_graph_data_json`` for a whole FlatBatch, computed from the flat gate stream (no
    per-circuit dicts, no Python loop over gates): the DAGOpNode feature matrix of every gate of
    every circuit and the DAGOpNode_wire_DAGOpNode edges (consecutive gates on a qubit wire), i.e.
    exactly what the graph model consumes (loaders/exp_val.py:63-89 keeps only these).  The gate
    stream carries no barriers / measurements (``Circuit.flat`` strips final measurements).

    Returns a dict of numpy arrays: ``x`` [n_ops, n_features] float32, ``edge_src`` / ``edge_dst``
    (global op indices, in the order circuit_to_graph_data_json emits them: by destination gate,
    then by the gate's qubit order), ``op_offsets`` [n_circuits + 1] and ``edge_offsets``
    [n_circuits + 1].  tests/test_features.py checks it against the per-circuit function."""
    from .gateset import NAMES, NUM_PARAMS, OPCODES

    types = [canonical(g) for g in properties["gates_set"]] + ["barrier", "measure"]
    ops = batch.ops
    n_ops = len(ops)
    opc = ops["opcode"].astype(np.int64)
    col_of = np.full(max(NAMES) + 1, -1, dtype=np.int64)
    for j, g in enumerate(types):
        if g in OPCODES:
            col_of[OPCODES[g]] = j
    cols = col_of[opc]
    if n_ops and cols.min() < 0:
        bad = NAMES[int(opc[np.argmin(cols)])]
        raise ValueError(f"gate {bad!r} is not in the backend's gate set {properties['gates_set']}")
    two = np.zeros(max(NAMES) + 1, dtype=bool)
    npar_of = np.zeros(max(NAMES) + 1, dtype=np.int64)
    for name, code in OPCODES.items():
        two[code] = (32 <= code < 64) or name == "unitary2"
        npar_of[code] = NUM_PARAMS.get(name, 0)
    is2 = two[opc]
    q0 = ops["q0"].astype(np.int64)
    q1 = ops["q1"].astype(np.int64)
    nt = len(types)
    nf = 3 + nt + (9 if use_qubit_features else 0) + (2 if use_gate_features else 0)
    x = np.zeros((n_ops, nf), dtype=np.float64)
    # up to three gate parameters
    npar = np.minimum(npar_of[opc], 3)
    pidx = ops["param_idx"].astype(np.int64)
    for k in range(3):
        m = npar > k
        x[m, k] = batch.params[pidx[m] + k]
    x[np.arange(n_ops), 3 + cols] = 1.0
    c = 3 + nt
    if use_qubit_features:
        qp = {int(k): v for k, v in properties["qubits_props"].items()}
        nq = max(qp) + 1 if qp else 0
        tab = np.zeros((nq + 1, 3))  # row nq = "no qubit"
        for q, v in qp.items():
            tab[q] = (v.get("t1", 0.0), v.get("t2", 0.0), v.get("readout_error", 0.0))
        if n_ops and (q0.max() >= nq or (is2.any() and q1[is2].max() >= nq)):
            raise KeyError("qubit without properties")
        qb = np.where(is2, q1, nq)
        for k in range(3):                       # layout: t1 x 3 qubits, t2 x 3, readout x 3 (third qubit never set)
            x[:, c + 3 * k + 0] = tab[q0, k]
            x[:, c + 3 * k + 1] = tab[qb, k]
        c += 9
    if use_gate_features:
        key = opc * 65536 + q0 * 256 + np.where(is2, q1, 255)
        uniq, inv = np.unique(key, return_inverse=True)
        vals = np.zeros((len(uniq), 2))
        for j, k in enumerate(uniq.tolist()):
            name, a, b = NAMES[k >> 16], (k >> 8) & 255, k & 255
            gp = properties["gate_props"].get(f"{name}_{a}" if b == 255 else f"{name}_{a}_{b}", {})
            vals[j] = (gp.get("gate_error", 0.0), gp.get("gate_length", 0.0))
        x[:, c:c + 2] = vals[inv]
    # wire edges: incidences (op, qubit) sorted by (circuit, qubit, op); neighbours on the same wire are linked
    n = batch.n_circuits
    op_offsets = batch.op_offsets.astype(np.int64)
    circ_of = np.repeat(np.arange(n, dtype=np.int64), np.diff(op_offsets))
    op_idx = np.arange(n_ops, dtype=np.int64)
    inc_op = np.concatenate([op_idx, op_idx[is2]])
    inc_q = np.concatenate([q0, q1[is2]])
    inc_slot = np.concatenate([np.zeros(n_ops, dtype=np.int64), np.ones(int(is2.sum()), dtype=np.int64)])
    wire = circ_of[inc_op] * 256 + inc_q
    order = np.lexsort((inc_op, wire))
    w_s, o_s, s_s = wire[order], inc_op[order], inc_slot[order]
    link = np.nonzero(w_s[1:] == w_s[:-1])[0]
    src, dst, slot = o_s[link], o_s[link + 1], s_s[link + 1]
    eorder = np.lexsort((slot, dst))
    src, dst = src[eorder], dst[eorder]
    edge_offsets = np.searchsorted(dst, op_offsets, side="left").astype(np.int64)
    return {"x": x.astype(np.float32), "edge_src": src, "edge_dst": dst, "op_offsets": op_offsets, "edge_offsets": edge_offsets}
