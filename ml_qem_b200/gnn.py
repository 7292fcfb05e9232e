"""Plain-torch graph model of the reference's GNN mitigator (BASELINE configs[4]) -- no torch_geometric.

The reference trains ``ExpValCircuitGraphModel_3`` (docs/tutorials/gnn.py:178-224): two
``TransformerConv`` layers, each followed by ``ASAPooling(ratio=0.5)``, a global mean pool, and the
``MLP3`` head (docs/tutorials/mlp.py:69-110) on [graph embedding, noisy values, circuit depth];
trained with Adam + ReduceLROnPlateau on MSE (gnn.py:282-378).  torch_geometric is not installable
here, so the three PyG layers are restated with index_add / scatter ops on the flat
(x, edge_index, batch) representation PyG uses:

  TransformerConv  multi-head dot-product attention over incoming edges, root (skip) weight, concat heads
  ASAPooling       master-query attention over each node's neighbourhood -> cluster features, LEConv
                   fitness score, top-ceil(ratio n) nodes per graph, coarsened adjacency S^T A S
  global_mean_pool mean of the node features per graph

The labels (ideal values) and the noisy inputs come straight from the engine: ``graph_batch`` takes
device tensors (e.g. filled by ``Engine.run_dm_into``) without a host round trip.
"""
import math

import numpy as np
import torch
from torch import nn


# ------------------------------------------------------------------------------------------ helpers
def scatter_softmax(src, index, n):
    """softmax of src[e] over the entries sharing index[e] (last dim of src = heads or nothing)."""
    shape = (n,) + tuple(src.shape[1:])
    mx = torch.full(shape, -float("inf"), dtype=src.dtype, device=src.device)
    mx = mx.scatter_reduce(0, index.view(-1, *([1] * (src.dim() - 1))).expand_as(src), src, reduce="amax", include_self=True)
    ex = torch.exp(src - mx[index])
    den = torch.zeros(shape, dtype=src.dtype, device=src.device).index_add_(0, index, ex)
    return ex / (den[index] + 1e-16)


def scatter_max(src, index, n):
    out = torch.full((n,) + tuple(src.shape[1:]), -float("inf"), dtype=src.dtype, device=src.device)
    return out.scatter_reduce(0, index.view(-1, 1).expand_as(src), src, reduce="amax", include_self=True)


def global_mean_pool(x, batch, n_graphs):
    out = torch.zeros(n_graphs, x.shape[1], dtype=x.dtype, device=x.device).index_add_(0, batch, x)
    cnt = torch.zeros(n_graphs, dtype=x.dtype, device=x.device).index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return out / cnt.clamp_min(1.0).unsqueeze(1)


def add_remaining_self_loops(edge_index, n):
    keep = edge_index[0] != edge_index[1]
    loops = torch.arange(n, device=edge_index.device)
    return torch.cat([edge_index[:, keep], torch.stack([loops, loops])], dim=1)


# ------------------------------------------------------------------------------------------- layers
class TransformerConv(nn.Module):
    """torch_geometric.nn.TransformerConv(in, out, heads, concat=True, beta=False, dropout, root_weight=True):
    out_i = W_skip x_i + ||_h sum_{j -> i} softmax_j((W_q x_i)^T (W_k x_j) / sqrt(out)) W_v x_j."""

    def __init__(self, in_channels, out_channels, heads=1, dropout=0.0):
        super().__init__()
        self.heads, self.out_channels, self.dropout = heads, out_channels, dropout
        self.lin_key = nn.Linear(in_channels, heads * out_channels)
        self.lin_query = nn.Linear(in_channels, heads * out_channels)
        self.lin_value = nn.Linear(in_channels, heads * out_channels)
        self.lin_skip = nn.Linear(in_channels, heads * out_channels)

    def forward(self, x, edge_index):
        n, h, c = x.shape[0], self.heads, self.out_channels
        src, dst = edge_index[0], edge_index[1]
        q = self.lin_query(x).view(n, h, c)[dst]
        k = self.lin_key(x).view(n, h, c)[src]
        v = self.lin_value(x).view(n, h, c)[src]
        alpha = scatter_softmax((q * k).sum(-1) / math.sqrt(c), dst, n)
        alpha = nn.functional.dropout(alpha, p=self.dropout, training=self.training)
        out = torch.zeros(n, h, c, dtype=x.dtype, device=x.device).index_add_(0, dst, v * alpha.unsqueeze(-1))
        return out.view(n, h * c) + self.lin_skip(x)


class LEConv(nn.Module):
    """torch_geometric.nn.LEConv: out_i = W3 x_i + sum_{j -> i} (W1 x_j - W2 x_i)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin1 = nn.Linear(in_channels, out_channels)
        self.lin2 = nn.Linear(in_channels, out_channels, bias=False)
        self.lin3 = nn.Linear(in_channels, out_channels)

    def forward(self, x, edge_index):
        src, dst = edge_index[0], edge_index[1]
        a, b = self.lin1(x), self.lin2(x)
        out = torch.zeros_like(a).index_add_(0, dst, a[src] - b[dst])
        return out + self.lin3(x)


class ASAPooling(nn.Module):
    """torch_geometric.nn.ASAPooling(in_channels, ratio) with its defaults (no intra-cluster GNN,
    negative_slope 0.2, no dropout): returns (x, edge_index, batch, perm) of the pooled graph."""

    def __init__(self, in_channels, ratio=0.5, negative_slope=0.2):
        super().__init__()
        self.ratio, self.negative_slope = ratio, negative_slope
        self.lin = nn.Linear(in_channels, in_channels)
        self.att = nn.Linear(2 * in_channels, 1)
        self.gnn_score = LEConv(in_channels, 1)

    def forward(self, x, edge_index, batch, n_graphs):
        n = x.shape[0]
        ei = add_remaining_self_loops(edge_index, n)
        src, dst = ei[0], ei[1]
        x_j = x[src]
        x_q = self.lin(scatter_max(x_j, dst, n))[dst]                       # master query of every cluster
        score = nn.functional.leaky_relu(self.att(torch.cat([x_q, x_j], dim=-1)).view(-1), self.negative_slope)
        score = scatter_softmax(score, dst, n)
        xc = torch.zeros_like(x).index_add_(0, dst, x_j * score.unsqueeze(1))  # cluster representations
        fitness = torch.sigmoid(self.gnn_score(xc, ei)).view(-1)
        # top ceil(ratio * n_g) clusters of every graph
        order = torch.argsort(fitness + 2.0 * (n_graphs - 1 - batch).to(fitness.dtype), descending=True)  # by graph, then fitness
        sorted_batch = batch[order]
        counts = torch.bincount(batch, minlength=n_graphs)
        keep_n = torch.ceil(self.ratio * counts.to(torch.float)).to(torch.long)
        start = torch.cumsum(counts, 0) - counts
        rank = torch.arange(n, device=x.device) - start[sorted_batch]
        perm = order[rank < keep_n[sorted_batch]]
        x_out = xc[perm] * fitness[perm].unsqueeze(1)
        # coarsened adjacency A' = S^T A S with S[j, cluster i] = score(j -> i), restricted to the kept
        # clusters.  Only its sparsity pattern is used downstream (TransformerConv takes no edge
        # weights), and PyG forms it with sparse products (structural non-zeros); here: two
        # sparse x sparse products of 0/1 matrices -- the dense n x n form costs 2 TFLOP per pooling
        # layer at 64 graphs x 200 nodes
        new_id = torch.full((n,), -1, dtype=torch.long, device=x.device)
        new_id[perm] = torch.arange(perm.numel(), device=x.device)
        m = perm.numel()
        new_batch = batch[perm]
        with torch.no_grad():
            sel = new_id[dst] >= 0
            one = lambda k: torch.ones(k, dtype=torch.float32, device=x.device)
            S = torch.sparse_coo_tensor(torch.stack([src[sel], new_id[dst][sel]]), one(int(sel.sum())), (n, m), check_invariants=False).coalesce()
            A = torch.sparse_coo_tensor(ei, one(ei.shape[1]), (n, n), check_invariants=False).coalesce()
            Ac = torch.sparse.mm(S.t().coalesce(), torch.sparse.mm(A, S)).coalesce()
            idx = Ac.indices()
            new_ei = idx[:, idx[0] != idx[1]].contiguous()   # row-major order, as nonzero() of the dense form
        return x_out, new_ei, new_batch, perm


class MLP3(nn.Module):
    """docs/tutorials/mlp.py:69-110."""

    def __init__(self, input_size, hidden_size, output_size, dropout_rate=0.3):
        super().__init__()
        self.fc1, self.bn1 = nn.Linear(input_size, hidden_size), nn.BatchNorm1d(hidden_size)
        self.fc2, self.bn2 = nn.Linear(hidden_size, hidden_size), nn.BatchNorm1d(hidden_size)
        self.fc3 = nn.Linear(hidden_size, hidden_size // 3)
        self.fc4 = nn.Linear(hidden_size // 3, output_size)
        self.drop = nn.Dropout(dropout_rate)

    def forward(self, x):
        x = self.drop(torch.relu(self.bn1(self.fc1(x))))
        x = self.drop(torch.relu(self.bn2(self.fc2(x))))
        x = self.drop(torch.relu(self.fc3(x)))
        return self.fc4(x)


class ExpValCircuitGraphModel(nn.Module):
    """ExpValCircuitGraphModel_3 of docs/tutorials/gnn.py:178-224 (same call signature)."""

    def __init__(self, num_node_features, hidden_channels, exp_value_size=4, dropout=0.3):
        super().__init__()
        self.transformer1 = TransformerConv(num_node_features, hidden_channels, heads=5, dropout=0.1)
        self.pooling1 = ASAPooling(hidden_channels * 5, 0.5)
        self.transformer2 = TransformerConv(hidden_channels * 5, hidden_channels, heads=3, dropout=0.1)
        self.pooling2 = ASAPooling(hidden_channels * 3, 0.5)
        self.body_seq = MLP3(hidden_channels * 3 + 1 + exp_value_size, hidden_channels * 5, exp_value_size, dropout)

    def forward(self, exp_value, observable, circuit_depth, nodes, edge_index, batch, n_graphs=None):
        n_graphs = int(batch.max().item()) + 1 if n_graphs is None else n_graphs
        g = self.transformer1(nodes, edge_index)
        g, edge_index, batch, _ = self.pooling1(g, edge_index, batch, n_graphs)
        g = self.transformer2(g, edge_index)
        g, edge_index, batch, _ = self.pooling2(g, edge_index, batch, n_graphs)
        g = global_mean_pool(g, batch, n_graphs)
        merge = torch.cat((g, exp_value.reshape(n_graphs, -1), circuit_depth.reshape(n_graphs, 1)), dim=1)
        return self.body_seq(merge)


# ------------------------------------------------------------------------------------- data + training
def graph_batch(entries, noisy=None, ideal=None, device="cpu"):
    """Collates dataset entries (features.ExpValueEntry or their dicts) into the flat PyG layout:
    dict(x, edge_index, batch, noisy_0, y, circuit_depth, observable, n_graphs).  ``noisy`` / ``ideal``:
    optional [n_graphs, exp_value_size] tensors already on ``device`` (the engine's zero-copy output)
    that replace the values stored in the entries."""
    xs, eis, bs, deps, obs, n0, ys = [], [], [], [], [], [], []
    off = 0
    for g, e in enumerate(entries):
        t = e.to_tensors() if hasattr(e, "to_tensors") else e
        x, ei = t["x"], t["edge_index"]
        loops = torch.arange(x.shape[0])
        ei = torch.cat([ei[:, ei[0] != ei[1]], torch.stack([loops, loops])], dim=1)  # AddSelfLoops transform of the loader
        xs.append(x); eis.append(ei + off); bs.append(torch.full((x.shape[0],), g, dtype=torch.long))
        deps.append(t["circuit_depth"].reshape(1, 1)); obs.append(t["observable"])
        n0.append(t["noisy_0"].reshape(1, -1)); ys.append(t["y"].reshape(1, -1))
        off += x.shape[0]
    out = {"x": torch.cat(xs).to(device), "edge_index": torch.cat(eis, dim=1).to(device), "batch": torch.cat(bs).to(device),
           "circuit_depth": torch.cat(deps).to(device), "observable": obs, "n_graphs": len(entries)}
    out["noisy_0"] = noisy.to(torch.float) if noisy is not None else torch.cat(n0).to(device)
    out["y"] = ideal.to(torch.float) if ideal is not None else torch.cat(ys).to(device)
    return out


def graph_batches_flat(flat, noisy, ideal, depth, batch_size, device="cpu", first=0, last=None, drop_last=True):
    """Mini-batches in the layout of ``graph_batch`` straight from ``features.graph_tensors_flat``
    (one vectorised pass over the gate stream of the whole dataset instead of a JSON graph, an
    entry object and five tensors per circuit).  ``noisy`` / ``ideal``: [n_circuits, exp_value_size]
    arrays or tensors (tensors on ``device`` are sliced in place -- the engine's zero-copy output);
    ``depth``: [n_circuits].  Circuits ``first`` .. ``last`` in chunks of ``batch_size``; equal to
    ``graph_batch`` on the same circuits (tests/test_gnn.py), self loops included."""
    x_all = torch.from_numpy(flat["x"]).to(device)
    oo, eo = flat["op_offsets"], flat["edge_offsets"]
    src_all, dst_all = torch.from_numpy(flat["edge_src"]).to(device), torch.from_numpy(flat["edge_dst"]).to(device)
    op_off = torch.from_numpy(oo).to(device)
    n = len(oo) - 1
    last = n if last is None else last
    as_t = lambda a: a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a), dtype=torch.float)
    noisy, ideal, depth = as_t(noisy).to(device, torch.float), as_t(ideal).to(device, torch.float), as_t(depth).to(device, torch.float)
    out = []
    for a in range(first, last, batch_size):
        b = min(a + batch_size, last)
        if b - a < batch_size and drop_last:
            break
        o0, o1, e0, e1 = int(oo[a]), int(oo[b]), int(eo[a]), int(eo[b])
        n_nodes = o1 - o0
        sizes = op_off[a + 1:b + 1] - op_off[a:b]
        batch = torch.repeat_interleave(torch.arange(b - a, device=device), sizes)
        src, dst = src_all[e0:e1] - o0, dst_all[e0:e1] - o0
        keep = src != dst
        src, dst = src[keep], dst[keep]
        loops = torch.arange(n_nodes, device=device)
        # per graph: its wire edges, then its self loops (the order graph_batch produces)
        ei = torch.stack([torch.cat([src, loops]), torch.cat([dst, loops])])
        key = torch.cat([2 * batch[dst], 2 * batch + 1])
        ei = ei[:, torch.argsort(key, stable=True)]
        out.append({"x": x_all[o0:o1], "edge_index": ei, "batch": batch, "circuit_depth": depth[a:b].reshape(-1, 1),
                    "observable": [torch.zeros(1, 0)] * (b - a), "n_graphs": b - a,
                    "noisy_0": noisy[a:b].reshape(b - a, -1), "y": ideal[a:b].reshape(b - a, -1)})
    return out


def train(model, train_batches, val_batches, epochs=100, lr=1e-3, patience=15, min_lr=1e-5, on_epoch=None):
    """The loop of docs/tutorials/gnn.py:282-378: Adam, ReduceLROnPlateau(factor 0.1), MSE; returns
    (train_losses, val_losses) per epoch."""
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, "min", factor=0.1, patience=patience, min_lr=min_lr)
    crit = nn.MSELoss()
    hist_t, hist_v = [], []

    def fwd(b):
        return model(b["noisy_0"], b["observable"], b["circuit_depth"], b["x"], b["edge_index"], b["batch"], b["n_graphs"])

    for epoch in range(epochs):
        model.train()
        tl = 0.0
        for b in train_batches:
            opt.zero_grad()
            loss = crit(fwd(b), b["y"])
            loss.backward()
            opt.step()
            tl += float(loss.item())
        model.eval()
        vl = 0.0
        with torch.no_grad():
            for b in val_batches:
                vl += float(crit(fwd(b), b["y"]).item())
        sched.step(vl)
        hist_t.append(tl / max(1, len(train_batches)))
        hist_v.append(vl / max(1, len(val_batches)))
        if on_epoch is not None:
            on_epoch(epoch, hist_t[-1], hist_v[-1])
    return hist_t, hist_v


# ------------------------------------------------------------------------------- ngem() consumer
class NgemJob:
    """blackwater/library/ngem/estimator.py:23-98: result() turns every (value, circuit, observable)
    into a graph sample and replaces the value by the model's prediction; ``result.metadata`` is passed
    through unchanged (:86).  The reference evaluates the model once per circuit in a Python loop
    (:49-84); here the samples of a job are collated and go through ONE forward pass."""

    def __init__(self, base_job, model, backend, circuits, observables, parameter_values, options=None):
        self._base_job, self._model, self._backend = base_job, model, backend
        self._circuits, self._observables, self._parameter_values = circuits, observables, parameter_values
        self._options = options

    def result(self):
        from . import features as FT, learning, observable as observable_mod
        from .estimator import EstimatorResult

        result = self._base_job.result()
        properties = self._backend if isinstance(self._backend, dict) and "gates_set" in self._backend else FT.backend_properties_v1(self._backend)
        entries = []
        for value, circuit, obs, params in zip(result.values, self._circuits, self._observables, self._parameter_values):
            try:
                observable_mod.from_any(obs)
            except TypeError as exc:  # BlackwaterException in the reference (:53-56)
                raise ValueError("Only `PauliSumOp` observables are supported by NGEM.") from exc
            bound = learning._bind(circuit, params)
            graph = FT.circuit_to_graph_data_json(bound, properties, use_qubit_features=True, use_gate_features=True)
            # (circuit_depth keeps the entry's default 0, as in the reference's NgemJob :71-76)
            entries.append(FT.ExpValueEntry(circuit_graph=graph, observable=learning.encode_pauli_sum_op(obs), ideal_exp_value=0.0,
                                            noisy_exp_values=[float(value)]))
        batch = graph_batch(entries, device=next(self._model.parameters()).device)
        was_training = self._model.training
        self._model.eval()
        with torch.no_grad():
            out = self._model(batch["noisy_0"], batch["observable"], batch["circuit_depth"], batch["x"], batch["edge_index"],
                              batch["batch"], batch["n_graphs"])
        self._model.train(was_training)
        return EstimatorResult(out.reshape(len(entries), -1)[:, 0].double().cpu().numpy(), result.metadata)

    def submit(self):
        return self._base_job.submit()

    def status(self):
        return self._base_job.status()

    def cancel(self):
        return self._base_job.cancel()

    def job_id(self):
        return self._base_job.job_id()

    def __repr__(self):
        return f"<NgemJob: {self._base_job.job_id()}>"


def ngem(cls, model, backend, options=None):
    """Decorator to turn an Estimator class into an NGEM estimator class (ngem/estimator.py:137-158):
    subclass, replace ``_run`` by a wrapper that calls the original with keyword arguments (:120-125)."""
    from functools import wraps

    run = cls._run

    @wraps(run)
    def ngem_run(self, circuits, observables, parameter_values, **run_options):
        job = run(self, circuits=circuits, observables=observables, parameter_values=parameter_values, **run_options)
        return NgemJob(job, model=model, backend=backend, circuits=circuits, observables=observables,
                       parameter_values=parameter_values, options=options)

    new_class = type(f"NGEM{cls.__name__}", (cls,), {})
    new_class._run = ngem_run
    return new_class
