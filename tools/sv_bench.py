"""Wide statevector timing on one GPU: python tools/sv_bench.py [n ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ml_qem_b200 import engine, families as F
eng = engine.Engine(0)
for n in [int(x) for x in sys.argv[1:]] or [16, 20, 24, 28]:
    steps = 10
    nc = max(1, min(256, 2 ** (26 - n))) if n < 26 else 1
    circs = [F.tfim_circuit(n, steps, 0.3 + 0.01 * i) for i in range(nc)]
    obs = F.tfim_observables(list(range(n)), n)
    b = engine.encode_batch(circs, [obs] * nc)
    for tb in (11, 12):
        eng.set_options(sv_tile_bits=tb)
        eng.prepare_sv(b)
        for _ in range(2):
            v = eng.execute_sv()
        s = eng.stats()
        print(f"sv n={n} circuits={nc} tile_bits={tb} kernel_ms={s['kernel_ms']:.3f} launches={s['n_other_launches']} "
              f"GB/s={s['sv_state_bytes_swept'] / s['kernel_ms'] / 1e6:.0f} circ/s={nc / s['kernel_ms'] * 1e3:.1f} norm={v[-1]:.3e}", flush=True)
