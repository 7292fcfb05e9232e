"""Host-side view of the cfg2 host-buffer call: wall time against the engine's own counters for a
series of calls, then one call with the phase trace (BWQ_TRACE=1): python tools/e2e_trace.py [calls]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ml_qem_b200 import engine, noise
wl = bench.build_workload("brick10_guadalupe_twirl", 0, 1.0)
batch = engine.encode_batch(wl["circuits"], wl["observables"])
eng = engine.Engine(0); eng.set_noise(noise.from_backend(wl["backend"]))
for _ in range(3): eng.run_meas_data(batch)
rows = []
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
    t = time.perf_counter(); eng.run_meas_data(batch); w = 1e3 * (time.perf_counter() - t); s = eng.stats()
    rows.append((w, s["kernel_ms"], s["lower_ms"]))
print("wall / device(first launch..last D2H) / lowering(sum) ms per call:")
print("  " + "  ".join("%.1f/%.1f/%.1f" % r for r in rows))
print("nproc", os.cpu_count(), "mean wall %.1f  mean device %.1f" % (sum(r[0] for r in rows) / len(rows), sum(r[1] for r in rows) / len(rows)))
os.environ["BWQ_TRACE"] = "1"
eng.run_meas_data(batch)
