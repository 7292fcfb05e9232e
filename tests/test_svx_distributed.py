"""world_size-2 and -4 gloo runs of the amplitude-sharded statevector orchestration
(ml_qem_b200.statevector.ShardedStatevector): real all_to_all_single / all_reduce between
processes, local segments executed by the numpy emulator.  No GPU needed."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import helpers
        from svx_emulator import EmulatorExecutor
        from ml_qem_b200 import families as F
        from ml_qem_b200.statevector import ShardedStatevector

        sv = ShardedStatevector(EmulatorExecutor(), dist)
        n = 9
        rng = np.random.default_rng(3)
        cm = [(i, i + 1) for i in range(n - 1)] + [(i + 1, i) for i in range(n - 1)]
        cases = [(F.tfim_circuit(n, 3, 0.4, basis="Y"), F.tfim_observables(list(range(n)), n)),
                 (F.random_basis_circuit(n, 70, rng, cm), [[("XYZIIZYXI", 0.7), ("ZZZZZZZZZ", 1.0)], [("IIIIXIIII", 1.0)]])]
        err = 0.0
        for circ, obs in cases:
            vals = sv.estimate(circ, obs, tile_bits=5)
            ref = helpers.oracle_sv_values(circ, obs)
            err = max(err, float(np.max(np.abs(vals - ref))))
            assert sv.last_plan["n_exchanges"] >= 1
        np.save(os.path.join(out_dir, f"err{rank}.npy"), np.array([err]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_statevector_gloo(lib, tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / f"err{r}.npy")[0] <= 1e-12
