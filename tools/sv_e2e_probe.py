import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from ml_qem_b200 import backends, engine, families as F, noise
from ml_qem_b200.statevector import GpuExecutor, ShardedStatevector
from ml_qem_b200.engine import SvxProgram, encode_batch
eng = engine.Engine(0)
if len(sys.argv) > 1:   # occupy memory like the default bench does before the sub-workload
    be = backends.synthetic_chain(14, seed=14); eng.set_noise(noise.from_backend(be))
    circs, obs = F.config_tfim_dm(n=14, n_circuits=8, max_steps=3)
    eng.run_meas_data(encode_batch(circs, [obs] * len(circs)))
    print("after tfim14: free GB", torch.cuda.mem_get_info()[0] / 1e9)
sv = ShardedStatevector(GpuExecutor(eng), None)
n = 30; obs = F.tfim_observables(list(range(n)), n)
for i in range(4):
    c = F.tfim_circuit(n, 3 + i % 3, 0.3 + 0.1 * i, dt=0.25)
    t0 = time.perf_counter(); batch = encode_batch([c], [obs]); t1 = time.perf_counter()
    prog = SvxProgram(batch, 0, 0, 0); t2 = time.perf_counter()
    prog.upload(eng); t3 = time.perf_counter(); prog.close(); t4 = time.perf_counter()
    v = sv.estimate(c, obs, profile=True); t5 = time.perf_counter()
    print("encode %.1f lower %.1f upload %.1f close %.1f | estimate wall %.1f device %.1f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t5-t4), sv.last_plan["ms_total"]))
