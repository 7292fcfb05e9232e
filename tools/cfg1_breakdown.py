"""Where the time of the cfg1 (tfim4_lima_zne) host-buffer calls goes: python tools/cfg1_breakdown.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ml_qem_b200 import backends, engine, families as F, noise
from ml_qem_b200.engine import Variants
noisy, base, obs = F.config_tfim4_lima_zne(n_base=2000, seed=0)
eng = engine.Engine(0)
nm = noise.from_backend(backends.fake_lima())
eng.set_noise(nm)
fb_all = engine.encode_batch(noisy, [obs] * len(noisy))
fb_base = engine.encode_batch(base, [obs] * len(base))
V = Variants(folds=(1, 3, 5))
def tm(f, n=5):
    f(); ts = []
    for _ in range(n):
        t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
    return 1e3 * min(ts), 1e3 * float(np.median(ts))
for name, f in (("dm_run 6000", lambda: eng.run_dm(fb_all)), ("sv_run 6000", lambda: eng.run_sv(fb_all)), ("sv_run 2000", lambda: eng.run_sv(fb_base)),
                ("dm_run_variants 2000x3", lambda: eng.run_dm_variants(fb_base, V)), ("meas_data 6000", lambda: eng.run_meas_data(fb_all)),
                ("meas_data_variants 2000x3", lambda: eng.run_meas_data_variants(fb_base, V))):
    lo, med = tm(f)
    s = eng.stats()
    print(f"{name:28s} min {lo:7.2f} ms  median {med:7.2f} ms | stats lower_ms {s['lower_ms']:.2f} h2d_ms {s['h2d_ms']:.2f} kernel_ms {s['kernel_ms']:.2f} d2h_ms {s['d2h_ms']:.2f} h2d_bytes {s['h2d_bytes']}")
for th in (4, 8, 16, 32):
    eng.set_options(host_threads=th)
    lo, med = tm(lambda: eng.run_dm_variants(fb_base, V))
    print("host_threads", th, "dm_run_variants min %.2f ms median %.2f" % (lo, med), "lower_ms", eng.stats()["lower_ms"])
