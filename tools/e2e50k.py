"""BASELINE configs[4] end to end on one GPU: dataset generation (noisy + ideal values of N mixed
6-12-qubit circuits through the engine, sharded + resumable) -> graph encoding -> plain-torch GNN
training.  Prints one JSON line with the wall time of every stage.

    python tools/e2e50k.py [--n 50000] [--epochs 2] [--out /tmp/e2e50k]

The reference does the same with a serial Aer loop (docs/tutorials/h13_ising_data_gen_tomo.ipynb:790
reports 2.44 circuits/s), circuit_to_graph_data_json per circuit and the PyG model of
docs/tutorials/gnn.py:178-378."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_qem_b200 import backends, dataset, engine, families as F, features as FT, gnn, noise  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--chunk", type=int, default=5000)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--out", default="/tmp/e2e50k")
    ap.add_argument("--per-circuit-graphs", action="store_true", help="build the graphs circuit by circuit (the reference's way) instead of from the flat gate stream")
    args = ap.parse_args()
    t = {}
    t0 = time.perf_counter()
    circs, obs = F.config_mixed_dataset(n_circuits=args.n, seed=5)
    t["build_circuits_s"] = time.perf_counter() - t0
    be = backends.synthetic_chain(12, seed=12, name="synthetic_chain_12q")
    eng = engine.Engine(0)
    nm = noise.from_backend(be)
    t0 = time.perf_counter()
    manifest = dataset.generate(circs, obs, lambda fb: eng.run_meas_data(fb, noise=nm)[:2], args.out, chunk_size=args.chunk, resume=False,
                                meta={"workload": "mixed6_12_dataset"})
    t["generate_s"] = time.perf_counter() - t0
    d = dataset.load(args.out)
    # graph samples: 4 targets per circuit = <Z> on its first four active qubits (every circuit has >= 6)
    props = FT.backend_properties_v1(be)
    dev = torch.device("cuda", 0)
    n_train = int(0.9 * len(circs))
    first4 = (np.asarray(d["obs_offsets"])[:-1, None] + np.arange(4)[None, :]).astype(np.int64)
    if args.per_circuit_graphs:
        # the reference's way: one JSON graph + entry object per circuit (circuit_to_graph_data_json), then collate
        t0 = time.perf_counter()
        entries = []
        for i, c in enumerate(circs):
            a = int(d["obs_offsets"][i])
            g = FT.circuit_to_graph_data_json(c, props, use_qubit_features=True, use_gate_features=True)
            entries.append(FT.ExpValueEntry(circuit_graph=g, observable=[], ideal_exp_value=d["ideal"][a:a + 4].tolist(),
                                            noisy_exp_values=[d["noisy"][a:a + 4].tolist()], circuit_depth=c.size()))
        t["graph_encode_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        mk = lambda es: [gnn.graph_batch(es[i:i + args.batch], device=dev) for i in range(0, len(es) - args.batch + 1, args.batch)]
        train_b, val_b = mk(entries[:n_train]), mk(entries[n_train:])
        t["collate_s"] = time.perf_counter() - t0
        nf = len(entries[0].circuit_graph["nodes"]["DAGOpNode"][0])
    else:
        # one vectorised pass over the flat gate stream of the whole dataset (features.graph_tensors_flat)
        t0 = time.perf_counter()
        fb = engine.encode_batch(circs, [[]] * len(circs))
        flat = FT.graph_tensors_flat(fb, props, use_gate_features=True, use_qubit_features=True)
        t["graph_encode_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        depth = np.diff(fb.op_offsets)
        noisy4, ideal4 = d["noisy"][first4], d["ideal"][first4]
        train_b = gnn.graph_batches_flat(flat, noisy4, ideal4, depth, args.batch, device=dev, first=0, last=n_train)
        val_b = gnn.graph_batches_flat(flat, noisy4, ideal4, depth, args.batch, device=dev, first=n_train)
        torch.cuda.synchronize()
        t["collate_s"] = time.perf_counter() - t0
        nf = flat["x"].shape[1]
    torch.manual_seed(0)
    model = gnn.ExpValCircuitGraphModel(num_node_features=nf, hidden_channels=15, exp_value_size=4).to(dev)
    t0 = time.perf_counter()
    tl, vl = gnn.train(model, train_b, val_b, epochs=args.epochs)
    torch.cuda.synchronize()
    t["train_s"] = time.perf_counter() - t0
    base = float(np.mean((d["noisy"] - d["ideal"]) ** 2))
    print(json.dumps({"workload": "e2e50k (BASELINE configs[4])", "n_circuits": args.n, "chunks": manifest["n_chunks"],
                      "graphs": "per circuit (JSON entries)" if args.per_circuit_graphs else "flat gate stream (graph_tensors_flat)", "stages_s": t, "generate_circuits_per_s": args.n / t["generate_s"], "epochs": args.epochs,
                      "train_loss": tl, "val_loss": vl, "mse_noisy_vs_ideal_all_observables": base,
                      "total_s": sum(t.values())}))


if __name__ == "__main__":
    main()
