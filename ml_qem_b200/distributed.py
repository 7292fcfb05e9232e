"""Circuit-sharded multi-GPU driver: one process per GPU (torchrun), no data-path collective.

The density-matrix path partitions circuit-by-circuit: every circuit is an
independent unit, so ranks only meet for the final gather of the [n_circuits x n_observables]
values.  Circuits are dealt longest-first (cost = gates x 4^n_active for the noisy run) so the
per-rank work is balanced.  Works with the NCCL backend on GPUs and with gloo on CPU (tests).
"""
import numpy as np

from .gateset import is_two_qubit, NAMES


def circuit_costs(batch):
    """Estimated cost of each circuit of a FlatBatch: gates x 4^active_qubits."""
    costs = np.zeros(batch.n_circuits)
    for c in range(batch.n_circuits):
        ops = batch.ops[batch.op_offsets[c]:batch.op_offsets[c + 1]]
        if len(ops) == 0:
            continue
        used = set(int(q) for q in ops["q0"])
        two = np.isin(ops["opcode"], [k for k, v in NAMES.items() if is_two_qubit(v)])
        used |= set(int(q) for q in ops["q1"][two])
        costs[c] = len(ops) * 4.0 ** len(used)
    return costs


def deal_longest_first(costs, world_size):
    """-> list (per rank) of circuit indices; greedy longest-processing-time assignment."""
    order = np.argsort(-np.asarray(costs), kind="stable")
    loads = np.zeros(world_size)
    shards = [[] for _ in range(world_size)]
    for c in order:
        r = int(np.argmin(loads))
        shards[r].append(int(c))
        loads[r] += costs[c]
    return [sorted(s) for s in shards]


def run_sharded(batch, run_local, dist=None, device=None):
    """Shards ``batch`` across the ranks of the default process group, runs ``run_local(sub_batch)
    -> values[n_observables_of_sub_batch]`` on every rank and gathers the values on every rank in
    the original observable order.  With dist=None (or world size 1) runs everything locally."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(run_local(batch))
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    shards = deal_longest_first(circuit_costs(batch), world)
    mine = shards[rank]
    local = np.asarray(run_local(batch.select(mine)), dtype=np.float64) if mine else np.zeros(0)
    n_obs_per = [int(sum(batch.obs_offsets[c + 1] - batch.obs_offsets[c] for c in s)) for s in shards]
    width = max(n_obs_per) if n_obs_per else 0
    buf = torch.zeros(width, dtype=torch.float64, device=device)
    buf[:len(local)] = torch.from_numpy(local).to(buf.device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)  # the one collective of the path: the final gather
    out = np.empty(batch.n_observables, dtype=np.float64)
    for r, s in enumerate(shards):
        vals = parts[r].cpu().numpy()
        k = 0
        for c in s:
            a, b = int(batch.obs_offsets[c]), int(batch.obs_offsets[c + 1])
            out[a:b] = vals[k:k + (b - a)]
            k += b - a
    return out
