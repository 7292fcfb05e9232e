"""Randomised parity sweep on the GPU: random circuits of 2..9 qubits (every 1-qubit gate kind, cx and
other 2-qubit gates, resets, idle qubits, multi-Pauli observables) on random backends, engine (all
kernel paths: on-chip, single-tile, TMA sweeps, narrow statevector) against the numpy oracle.
    python tools/fuzz_parity.py [n_cases] [seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

from ml_qem_b200 import Circuit, backends, engine, noise  # noqa: E402
from oracle import dm, noise_model as onm, sv  # noqa: E402

ONE = ["id", "x", "y", "z", "h", "s", "sdg", "t", "tdg", "sx", "sxdg", "rx", "ry", "rz", "p", "u2", "u3"]
TWO = ["cx", "cx", "cx", "cz", "swap", "rzz", "crx", "cp", "ecr", "cy", "ch", "rxx", "iswap"]
NPAR = {"rx": 1, "ry": 1, "rz": 1, "p": 1, "u2": 2, "u3": 3, "rzz": 1, "crx": 1, "cp": 1, "rxx": 1}


def case(rng, cx_only):
    n = int(rng.integers(2, 10))
    be = backends.synthetic_chain(n, seed=int(rng.integers(0, 1000)))
    pairs = be.coupling_map
    c = Circuit(n)
    for _ in range(int(rng.integers(0, 70))):
        r = rng.random()
        if r < 0.3 and pairs:
            a, b = pairs[int(rng.integers(0, len(pairs)))]
            g = "cx" if cx_only else TWO[int(rng.integers(0, len(TWO)))]
            c.append(g, (a, b), tuple(float(x) for x in rng.uniform(-3.2, 3.2, size=NPAR.get(g, 0))))
        elif r < 0.34:
            c.reset(int(rng.integers(0, n)))
        else:
            g = ONE[int(rng.integers(0, len(ONE)))]
            c.append(g, (int(rng.integers(0, n)),), tuple(float(x) for x in rng.uniform(-3.2, 3.2, size=NPAR.get(g, 0))))
    obs = [[("".join(rng.choice(list("IXYZ"), size=n)), float(rng.normal())) for _ in range(int(rng.integers(1, 4)))]
           for _ in range(int(rng.integers(1, 5)))]
    return be, c, obs


def run(n_cases=120, seed=0, eng=None):
    rng = np.random.default_rng(seed)
    eng = eng or engine.Engine(0)
    worst = {"dm": 0.0, "sv": 0.0}
    paths = {}
    t0 = time.time()
    for k in range(n_cases):
        be, c, obs = case(rng, cx_only=bool(k % 2))
        fb = engine.encode_batch([c], [obs])
        v, st = eng.run_dm(fb, noise=noise.from_backend(be))
        s = eng.stats()
        path = "onchip" if s["n_onchip_circuits"] else ("tma" if s["n_tma_sweep_launches"] else ("sweep" if s["n_sweep_launches"] else "host"))
        paths[path] = paths.get(path, 0) + 1
        assert not st.any(), (k, st)
        ref = dm.estimate(c.num_qubits, c.gate_ops(), obs, onm.from_backend(be.to_dict()))
        e = float(np.max(np.abs(v - ref)))
        worst["dm"] = max(worst["dm"], e)
        assert e <= 1e-10, ("dm", k, path, e)
        if not any(g == "reset" for g, _, _ in c.gate_ops()):
            v, st = eng.run_sv(fb)
            assert not st.any()
            e = float(np.max(np.abs(v - sv.estimate(c.num_qubits, c.gate_ops(), obs))))
            worst["sv"] = max(worst["sv"], e)
            assert e <= 1e-10, ("sv", k, e)
    print("fuzz ok:", n_cases, "cases", "paths", paths, "worst", worst, "%.1f s" % (time.time() - t0))
    return paths, worst


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 120, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
