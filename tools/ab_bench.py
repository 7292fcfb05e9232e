"""A/B of alternative library builds (BWQ_LIB=...): sweep-kernel time of three density-matrix cases."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_qem_b200 import backends, engine, families as F, noise
eng = engine.Engine(0)
FLAGS = int(os.environ.get("BWQ_FLAGS", "0"), 0)
eng.set_options(flags=FLAGS)
out = [os.path.basename(os.environ.get("BWQ_LIB", "default")) + " flags=%#x" % FLAGS]
be = backends.synthetic_chain(16, seed=2); eng.set_noise(noise.from_backend(be))
tw, base, obs = F.config_brick10_twirl(n_base=4, n_twirls=50)
b = engine.encode_batch(tw, [obs] * len(tw))
for _ in range(4): eng.run_dm(b)
st = eng.stats(); out.append("brick10 %.3f ms %.0f GB/s" % (st["sweep_kernel_ms"], st["state_bytes_swept"] / st["sweep_kernel_ms"] / 1e6))
for n in (12, 13):
    be = backends.synthetic_chain(n, seed=n); eng.set_noise(noise.from_backend(be))
    circs, obs = F.config_tfim_dm(n=n, n_circuits=4, max_steps=4)
    b = engine.encode_batch(circs, [obs] * len(circs))
    for _ in range(3): eng.run_dm(b)
    st = eng.stats(); out.append("tfim%d %.3f ms %.0f GB/s" % (n, st["sweep_kernel_ms"], st["state_bytes_swept"] / st["sweep_kernel_ms"] / 1e6))
print(" | ".join(out))
