"""Where the end-to-end time of cfg2 goes: wall time of run_dm / run_sv and the engine's own counters."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ml_qem_b200 import engine, noise
wl = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else "brick10_guadalupe_twirl", 0, 1.0)
t = time.perf_counter(); batch = engine.encode_batch(wl["circuits"], wl["observables"]); print("encode_batch %.1f ms" % (1e3 * (time.perf_counter() - t)))
eng = engine.Engine(0); eng.set_noise(noise.from_backend(wl["backend"]))
for fl in (0, 4):
    eng.set_options(flags=fl)
    for rep in range(3):
        t = time.perf_counter(); eng.run_dm(batch); w = 1e3 * (time.perf_counter() - t); s = eng.stats()
        print("flags", fl, "run_dm wall %.1f ms | lower %.1f h2d %.2f kernel %.1f d2h %.2f" % (w, s["lower_ms"], s["h2d_ms"], s["kernel_ms"], s["d2h_ms"]))
for rep in range(3):
    t = time.perf_counter(); eng.run_sv(batch); w = 1e3 * (time.perf_counter() - t); s = eng.stats()
    print("run_sv wall %.1f ms | lower %.1f h2d %.2f kernel %.1f d2h %.2f" % (w, s["lower_ms"], s["h2d_ms"], s["kernel_ms"], s["d2h_ms"]))
eng.set_options()
for rep in range(4):
    t = time.perf_counter(); eng.run_meas_data(batch); w = 1e3 * (time.perf_counter() - t); s = eng.stats()
    print("run_meas_data wall %.1f ms | dm side: lower %.1f h2d %.2f kernel %.1f d2h %.2f | circuits %d" % (w, s["lower_ms"], s["h2d_ms"], s["kernel_ms"], s["d2h_ms"], batch.n_circuits))
