"""Runs one density-matrix workload a few times (for ncu / launch lists).  usage:
   python tools/profile_case.py tfim 12 [kq] [low] [reps]   |   brick 10 ..."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from ml_qem_b200 import backends, engine, families as F, noise  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "tfim"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
kq = int(sys.argv[3]) if len(sys.argv) > 3 else 0
low = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
eng = engine.Engine(0)
eng.set_options(tile_qubits=kq, low_qubits=low, flags=int(os.environ.get("BWQ_FLAGS", "0"), 0))
if kind == "tfim":
    be = backends.synthetic_chain(n, seed=n)
    circs, obs = F.config_tfim_dm(n=n, n_circuits=2, max_steps=3)
    obs = [obs] * len(circs)
else:
    be = backends.synthetic_chain(16, seed=2)
    circs, _, o = F.config_brick10_twirl(n_base=2, n_twirls=32, n=n)
    obs = [o] * len(circs)
eng.set_noise(noise.from_backend(be))
b = engine.encode_batch(circs, obs)
for _ in range(reps):
    t = time.time()
    v, s = eng.run_dm(b)
    dt = time.time() - t
st = eng.stats()
print(kind, n, "circ/s", len(circs) / dt, "kernel_ms", st["kernel_ms"], "GB/s", st["state_bytes_swept"] / st["kernel_ms"] / 1e6,
      "sweeps", st["n_state_sweeps"], "passes", st["n_passes"], "launches", st["n_sweep_launches"])
