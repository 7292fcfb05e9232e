"""CPU: plain-torch restatement of the reference's GNN mitigator (docs/tutorials/gnn.py:178-224;
torch_geometric is not installable here): the layers against dense formulas of what PyG's
TransformerConv / ASAPooling compute, the model on graphs built by the feature encoder, and the
training loop of gnn.py:282-378 on a tiny synthetic task."""
import math

import numpy as np
import pytest
import torch

import helpers
from ml_qem_b200 import backends, families as F, features as FT, gnn


def _graph(n, rng):
    src = rng.integers(0, n, size=3 * n)
    dst = rng.integers(0, n, size=3 * n)
    return torch.tensor(np.stack([src, dst]), dtype=torch.long)


def test_transformer_conv_equals_dense_attention():
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    n, cin, c, h = 9, 6, 4, 3
    x = torch.randn(n, cin)
    ei = torch.unique(_graph(n, rng), dim=1)  # simple graph: the dense formula has one entry per (j, i)
    conv = gnn.TransformerConv(cin, c, heads=h, dropout=0.0).eval()
    out = conv(x, ei)
    q = conv.lin_query(x).view(n, h, c); k = conv.lin_key(x).view(n, h, c); v = conv.lin_value(x).view(n, h, c)
    adj = torch.zeros(n, n, dtype=torch.bool)
    adj[ei[1], ei[0]] = True  # adj[i, j]: edge j -> i
    ref = torch.zeros(n, h, c)
    for i in range(n):
        js = torch.nonzero(adj[i]).view(-1)
        if len(js) == 0:
            continue
        a = torch.softmax((q[i].unsqueeze(0) * k[js]).sum(-1) / math.sqrt(c), dim=0)  # [deg, h]
        ref[i] = (a.unsqueeze(-1) * v[js]).sum(0)
    ref = ref.view(n, h * c) + conv.lin_skip(x)
    assert torch.allclose(out, ref, atol=1e-6)


def test_asapooling_keeps_half_of_every_graph_and_coarsens_inside_graphs():
    torch.manual_seed(1)
    rng = np.random.default_rng(1)
    sizes = [7, 4, 10]
    xs, eis, bs, off = [], [], [], 0
    for g, n in enumerate(sizes):
        xs.append(torch.randn(n, 5)); eis.append(_graph(n, rng) + off); bs.append(torch.full((n,), g)); off += n
    x, ei, batch = torch.cat(xs), torch.cat(eis, dim=1), torch.cat(bs)
    pool = gnn.ASAPooling(5, 0.5)
    xo, eo, bo, perm = pool(x, ei, batch, len(sizes))
    assert [int((bo == g).sum()) for g in range(3)] == [math.ceil(0.5 * n) for n in sizes]
    assert xo.shape == (len(perm), 5) and torch.equal(bo, batch[perm])
    assert (bo[eo[0]] == bo[eo[1]]).all() and (eo[0] != eo[1]).all()
    # the kept clusters are the fittest of their graph; fitness = sigmoid(LEConv(cluster features))
    ei_l = gnn.add_remaining_self_loops(ei, x.shape[0])
    x_q = pool.lin(gnn.scatter_max(x[ei_l[0]], ei_l[1], x.shape[0]))[ei_l[1]]
    sc = gnn.scatter_softmax(torch.nn.functional.leaky_relu(pool.att(torch.cat([x_q, x[ei_l[0]]], -1)).view(-1), 0.2), ei_l[1], x.shape[0])
    xc = torch.zeros_like(x).index_add_(0, ei_l[1], x[ei_l[0]] * sc.unsqueeze(1))
    fit = torch.sigmoid(pool.gnn_score(xc, ei_l)).view(-1)
    for g in range(3):
        mine = torch.nonzero(batch == g).view(-1)
        top = mine[torch.argsort(fit[mine], descending=True)[:math.ceil(0.5 * len(mine))]]
        assert set(top.tolist()) == set(perm[bo == g].tolist())
    # coarsened edges == the pattern of the dense S^T A S (S = assignment scores of the kept clusters)
    n, m = x.shape[0], len(perm)
    new_id = torch.full((n,), -1, dtype=torch.long)
    new_id[perm] = torch.arange(m)
    keep = new_id[ei_l[1]] >= 0
    S = torch.zeros(n, m).index_put_((ei_l[0][keep], new_id[ei_l[1]][keep]), sc[keep].detach(), accumulate=True)
    A = torch.zeros(n, n).index_put_((ei_l[0], ei_l[1]), torch.ones(ei_l.shape[1]), accumulate=True)
    Ac = S.t() @ A @ S
    Ac.fill_diagonal_(0.0)
    assert torch.equal(eo, torch.nonzero(Ac != 0).t())
    xo.sum().backward()  # gradients reach the attention and the score network
    assert pool.att.weight.grad is not None and pool.gnn_score.lin1.weight.grad is not None


def _entries(n, rng):
    lima = backends.fake_lima()
    props = FT.backend_properties_v1(lima)
    out = []
    for i in range(n):
        c = F.tfim_circuit(4, 1 + i % 3, float(rng.uniform(0, 1)), layout=[0, 1, 3, 4], num_physical=5)
        g = FT.circuit_to_graph_data_json(c, props, use_qubit_features=True, use_gate_features=True)
        ideal = rng.uniform(-1, 1, size=4)
        noisy = 0.8 * ideal + 0.02 * rng.normal(size=4)
        out.append(FT.ExpValueEntry(circuit_graph=g, observable=[], ideal_exp_value=ideal.tolist(),
                                    noisy_exp_values=[noisy.tolist()], circuit_depth=c.size()))
    return out


def test_model_forward_and_training_loop():
    torch.manual_seed(2)
    rng = np.random.default_rng(2)
    entries = _entries(48, rng)
    nf = len(entries[0].circuit_graph["nodes"]["DAGOpNode"][0])
    batches = [gnn.graph_batch(entries[i:i + 16]) for i in range(0, 32, 16)]
    val = [gnn.graph_batch(entries[32:])]
    model = gnn.ExpValCircuitGraphModel(num_node_features=nf, hidden_channels=8, exp_value_size=4, dropout=0.0)
    out = model(batches[0]["noisy_0"], batches[0]["observable"], batches[0]["circuit_depth"], batches[0]["x"],
                batches[0]["edge_index"], batches[0]["batch"])
    assert out.shape == (16, 4) and torch.isfinite(out).all()
    tl, vl = gnn.train(model, batches, val, epochs=30, lr=3e-3)
    assert len(tl) == 30 and tl[-1] < 0.5 * tl[0] and np.isfinite(vl).all()
    # labels / noisy inputs handed over as tensors (the engine's zero-copy output) replace the stored ones
    noisy = torch.zeros(16, 4); ideal = torch.ones(16, 4)
    b = gnn.graph_batch(entries[:16], noisy=noisy, ideal=ideal)
    assert torch.equal(b["noisy_0"], noisy) and torch.equal(b["y"], ideal)


def test_flat_graph_batches_equal_the_per_circuit_path():
    """features.graph_tensors_flat + gnn.graph_batches_flat (one vectorised pass over the flat gate
    stream) == circuit_to_graph_data_json + ExpValueEntry.to_tensors + graph_batch, tensor for tensor."""
    from ml_qem_b200 import engine

    rng = np.random.default_rng(5)
    lima = backends.fake_lima()
    props = FT.backend_properties_v1(lima)
    circs = [F.tfim_circuit(4, 1 + i % 3, float(rng.uniform(0, 1)), layout=[0, 1, 3, 4], num_physical=5) for i in range(20)]
    circs += [F.random_basis_circuit(5, int(rng.integers(12, 50)), rng, lima.coupling_map) for _ in range(12)]
    ideal = rng.uniform(-1, 1, size=(len(circs), 4))
    noisy = 0.8 * ideal
    entries = [FT.ExpValueEntry(circuit_graph=FT.circuit_to_graph_data_json(c, props, use_qubit_features=True, use_gate_features=True),
                                observable=[], ideal_exp_value=ideal[i].tolist(), noisy_exp_values=[noisy[i].tolist()],
                                circuit_depth=c.size()) for i, c in enumerate(circs)]
    fb = engine.encode_batch(circs, [[]] * len(circs))
    flat = FT.graph_tensors_flat(fb, props, use_gate_features=True, use_qubit_features=True)
    got = gnn.graph_batches_flat(flat, noisy, ideal, [c.size() for c in circs], batch_size=8)
    assert len(got) == 4
    for k, g in enumerate(got):
        ref = gnn.graph_batch(entries[8 * k:8 * k + 8])
        for key in ("x", "edge_index", "batch", "circuit_depth", "noisy_0", "y"):
            assert torch.equal(g[key], ref[key]), (k, key)
        assert g["n_graphs"] == ref["n_graphs"]
    # a gate outside the backend's gate set is an error, as in the per-circuit function
    from ml_qem_b200 import Circuit
    c = Circuit(2); c.h(0)
    with pytest.raises(ValueError):
        FT.graph_tensors_flat(engine.encode_batch([c], [[]]), props)
