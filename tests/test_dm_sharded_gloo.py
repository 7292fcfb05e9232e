"""world_size-2 and -3 gloo runs of the circuit-sharded density-matrix driver
(ml_qem_b200.distributed.run_sharded): circuits dealt longest-first across the ranks, no data-path
collective, one all_gather of the values.  The local runs execute the lowered sweep programs with
the numpy emulator, so no GPU is needed; every rank must end with the full value vector, equal to
the oracle's."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _workload():
    from ml_qem_b200 import backends, families as F

    lima = backends.fake_lima()
    rng = np.random.default_rng(17)
    circs = [F.random_basis_circuit(5, int(rng.integers(0, 50)), rng, lima.coupling_map) for _ in range(11)]
    circs[4] = F.tfim_circuit(4, 2, 0.3, layout=[0, 1, 3, 4], num_physical=5, fold=3)
    obs = [[[("".join(rng.choice(list("IXYZ"), size=5)), float(rng.normal()))] for _ in range(1 + i % 3)] for i in range(len(circs))]
    return lima, circs, obs


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from program_emulator import expvals, run_program
        from ml_qem_b200 import engine, noise
        from ml_qem_b200.distributed import circuit_costs, deal_longest_first, run_sharded

        lima, circs, obs = _workload()
        nm = noise.from_backend(lima)
        batch = engine.encode_batch(circs, obs)

        def run_local(sub):
            out = []
            for c in range(sub.n_circuits):
                prog = engine.lower_dm(sub, c, nm)
                n_ob = int(sub.obs_offsets[c + 1] - sub.obs_offsets[c])
                t0 = int(sub.obs_offsets[c])
                counts = [int(sub.term_offsets[t0 + k + 1] - sub.term_offsets[t0 + k]) for k in range(n_ob)]
                out.append(expvals(prog, run_program(prog), counts))
            return np.concatenate(out) if out else np.zeros(0)

        vals = run_sharded(batch, run_local, dist)
        shards = deal_longest_first(circuit_costs(batch), world)
        assert sorted(c for s in shards for c in s) == list(range(len(circs)))
        np.save(os.path.join(out_dir, f"vals{rank}.npy"), vals)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_circuit_sharded_density_matrix_gloo(lib, tmp_path, world):
    import helpers

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    lima, circs, obs = _workload()
    on = helpers.oracle_noise("fakelima")
    ref = np.concatenate([helpers.oracle_dm_values(*helpers.compact(c, o, on)[:2], helpers.compact(c, o, on)[2]) for c, o in zip(circs, obs)])
    for r in range(world):
        vals = np.load(tmp_path / f"vals{r}.npy")
        assert vals.shape == ref.shape and np.max(np.abs(vals - ref)) <= 1e-12
