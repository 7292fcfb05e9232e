"""Sharded, resumable dataset generation on top of the engine.

The reference generates datasets with a serial Python loop and dumps one JSON list at the end
(blackwater/data/generators/exp_val.py:115-138, docs/tutorials/h13_ising_data_gen.ipynb:169-212):
a crash at circuit 49 000 of 50 000 loses everything.  Here the circuits go through the engine in
chunks; every finished chunk is one binary shard (``chunk_00012.npz``: ideal / noisy values, the flat
gate stream of the chunk) and a line in ``manifest.json``; a re-run skips the shards that are
already there (resume-by-chunk).  ``load`` gives the values back as numpy or torch tensors;
``entries`` rebuilds the reference's ``ExpValueEntry`` rows (JSON schema of loaders/exp_val.py:40-76)
for the consumers that want them.
"""
import json
import os

import numpy as np

from .engine import FlatBatch, encode_batch

MANIFEST = "manifest.json"


def _shard_path(out_dir, k):
    return os.path.join(out_dir, f"chunk_{k:05d}.npz")


def generate(circuits, observables, evaluate, out_dir, chunk_size=1000, resume=True, meta=None):
    """Evaluates ``circuits`` (any sequence; observables[i] = list of observables of circuit i) in
    chunks and writes one shard per chunk.  ``evaluate(flat_batch) -> (ideal, noisy)`` arrays of
    n_observables values each, e.g. ``lambda b: engine.run_meas_data(b)[:2]``.  Returns the manifest.
    With resume=True the chunks whose shard exists (and matches the chunk's circuit count) are not
    evaluated again."""
    os.makedirs(out_dir, exist_ok=True)
    n = len(circuits)
    n_chunks = (n + chunk_size - 1) // chunk_size
    manifest = {"n_circuits": n, "chunk_size": chunk_size, "n_chunks": n_chunks, "meta": meta or {}, "shards": []}
    for k in range(n_chunks):
        lo, hi = k * chunk_size, min(n, (k + 1) * chunk_size)
        path = _shard_path(out_dir, k)
        done = False
        if resume and os.path.exists(path):
            try:
                with np.load(path) as z:
                    done = int(z["n_circuits"]) == hi - lo and int(z["first"]) == lo
            except Exception:  # noqa: BLE001 - a truncated shard from a crashed run is simply redone
                done = False
        if not done:
            fb = encode_batch(circuits[lo:hi], observables[lo:hi])
            ideal, noisy = evaluate(fb)
            tmp = path + ".tmp.npz"
            np.savez(tmp, first=lo, n_circuits=hi - lo, ideal=np.asarray(ideal, dtype=np.float64),
                     noisy=np.asarray(noisy, dtype=np.float64), n_qubits=fb.n_qubits, op_offsets=fb.op_offsets, ops=fb.ops,
                     params=fb.params, obs_offsets=fb.obs_offsets, term_offsets=fb.term_offsets, term_x=fb.term_x,
                     term_z=fb.term_z, term_coeff=fb.term_coeff)
            os.replace(tmp, path)  # a shard appears atomically
        manifest["shards"].append({"file": os.path.basename(path), "first": lo, "n_circuits": hi - lo, "reused": bool(done)})
        with open(os.path.join(out_dir, MANIFEST), "w") as f:
            json.dump(manifest, f)
    return manifest


def load(out_dir, as_torch=False, device=None):
    """-> dict(ideal, noisy [n_observables total], obs_offsets [n_circuits + 1]) over all shards."""
    with open(os.path.join(out_dir, MANIFEST)) as f:
        manifest = json.load(f)
    ideal, noisy, obs_cnt = [], [], []
    for sh in manifest["shards"]:
        with np.load(os.path.join(out_dir, sh["file"])) as z:
            ideal.append(z["ideal"]); noisy.append(z["noisy"]); obs_cnt.append(np.diff(z["obs_offsets"]))
    out = {"ideal": np.concatenate(ideal), "noisy": np.concatenate(noisy),
           "obs_offsets": np.concatenate([[0], np.cumsum(np.concatenate(obs_cnt))])}
    if as_torch:
        import torch

        out = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) if device else torch.from_numpy(np.ascontiguousarray(v))
               for k, v in out.items()}
    return out


def shard_batch(out_dir, k):
    """The flat gate stream of shard k as a FlatBatch (e.g. for features.encode_data_flat)."""
    with np.load(_shard_path(out_dir, k)) as z:
        return FlatBatch(z["n_qubits"], z["op_offsets"], z["ops"], z["params"], z["obs_offsets"], z["term_offsets"],
                         z["term_x"], z["term_z"], z["term_coeff"])
