"""torchrun entry: amplitude-sharded statevector over N GPUs (NCCL).  --check compares against the
oracle at 16 qubits and against analytic/light-cone properties at --qubits (default 28)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--steps", type=int, default=4)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ml_qem_b200 import engine, families as F
    from ml_qem_b200.statevector import GpuExecutor, ShardedStatevector

    eng = engine.Engine(local)
    sv = ShardedStatevector(GpuExecutor(eng), dist)
    out = {"world": world}
    if args.check:
        import helpers
        n = 16
        rng = np.random.default_rng(1)
        cm = [(i, i + 1) for i in range(n - 1)] + [(i + 1, i) for i in range(n - 1)]
        worst = 0.0
        for circ, obs in ((F.tfim_circuit(n, 3, 0.45, basis="Y"), F.tfim_observables(list(range(n)), n)),
                          (F.random_basis_circuit(n, 200, rng, cm),
                           [[("XYZI" * 4, 0.7), ("Z" * n, 1.0)], [("IIIIXIIIIIIIIIII", 1.0)], [("Y" * n, 1.0)]])):
            vals = sv.estimate(circ, obs)
            ref = helpers.oracle_sv_values(circ, obs)
            worst = max(worst, float(np.max(np.abs(vals - ref))))
            assert sv.last_plan["n_exchanges"] >= 1, sv.last_plan
        out["oracle_max_abs_diff_16q"] = worst
        assert worst <= 1e-10, worst
    n, steps = args.qubits, args.steps
    h, dt = 1.0, 0.5
    circ = F.tfim_circuit(n, steps, 0.0, h=h, dt=dt)
    obs = F.tfim_observables(list(range(n)), n)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    vals = sv.estimate(circ, obs)
    torch.cuda.synchronize(); dist.barrier()
    dt_s = time.perf_counter() - t0
    z = np.cos(2 * h * dt * steps)
    err = max(float(np.max(np.abs(vals[:n] - z))), float(np.max(np.abs(vals[n:2 * n - 1] - z * z))),
              float(np.max(np.abs(vals[2 * n - 1:3 * n - 2]))), float(abs(vals[-1] - z ** n)))
    out.update({"qubits": n, "steps": steps, "product_state_max_abs_err": err, "seconds": dt_s, "plan": sv.last_plan})
    assert err <= 1e-10, err
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
