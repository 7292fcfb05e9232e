"""-m gpu: dm_onchip_kernel (ml_qem_b200/csrc/onchip.cuh) -- one warp interprets the raw gate stream
of a circuit of <= 5 active qubits.  Checked against the numpy oracle, against the lowering +
tile-sweep path (BWQ_OPT_NO_ONCHIP) and, for the variants, against variants built in Python."""
import numpy as np
import pytest

import helpers
from ml_qem_b200 import Circuit, backends, engine, families as F, noise
from ml_qem_b200.engine import Engine, Variants

pytestmark = pytest.mark.gpu
TOL = 1e-10
NO_ONCHIP = 128

ONE_Q = ["id", "x", "y", "z", "h", "s", "sdg", "t", "tdg", "sx", "sxdg", "rx", "ry", "rz", "p", "u2", "u3", "reset"]


def _random_circuit(rng, n, n_gates, pairs, gates=ONE_Q, unitary=False):
    c = Circuit(n)
    for _ in range(n_gates):
        if rng.random() < 0.3 and pairs:
            a, b = pairs[int(rng.integers(0, len(pairs)))]
            c.cx(a, b)
            continue
        g = gates[int(rng.integers(0, len(gates)))]
        q = int(rng.integers(0, n))
        npar = {"rx": 1, "ry": 1, "rz": 1, "p": 1, "u2": 2, "u3": 3}.get(g, 0)
        c.append(g, (q,), tuple(float(x) for x in rng.uniform(-3.2, 3.2, size=npar)))
        if unitary and rng.random() < 0.1:
            m = np.linalg.qr(rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2)))[0]
            c.ops.append(("unitary1", (q,), tuple(float(x) for z in m.reshape(-1) for x in (z.real, z.imag))))
    return c


def _labels(rng, n, k):
    return ["".join(rng.choice(list("IXYZ"), size=n)) for _ in range(k)]


def test_onchip_vs_oracle_and_tile_path_all_gates(engine_gpu):
    """Every 1-qubit gate kind + cx with FakeLima's errors, 2..5 qubits (idle qubits included):
    on-chip == oracle (1e-10) == lowering + sweeps (1e-12); the stats prove which path ran."""
    lima = backends.fake_lima()
    nm = noise.from_backend(lima)
    on = helpers.oracle_noise("fakelima")
    rng = np.random.default_rng(5)
    circs = []
    for k in range(24):
        pairs = [p for p in lima.coupling_map if rng.random() < 0.7] or [lima.coupling_map[0]]
        circs.append(_random_circuit(rng, 5, int(rng.integers(0, 60)), pairs, unitary=True))
    circs.append(Circuit(5))                        # no gate at all
    c1 = Circuit(5); c1.rz(0.3, 2); c1.sx(2)         # 1-qubit gates only
    circs.append(c1)
    obs = [[[(l, float(rng.normal()))] for l in _labels(rng, 5, 5)] + [[(l, float(rng.normal())) for l in _labels(rng, 5, 7)]] for _ in circs]
    fb = engine.encode_batch(circs, obs)
    vals, status = engine_gpu.run_dm(fb, noise=nm)
    assert engine_gpu.stats()["n_onchip_circuits"] == len(circs) and engine_gpu.stats()["n_sweep_launches"] == 0
    assert not status.any()
    ref = np.concatenate([helpers.oracle_dm_values(c, o, on) for c, o in zip(circs, obs)])
    assert np.max(np.abs(vals - ref)) <= TOL
    engine_gpu.set_options(flags=NO_ONCHIP)
    vals_t, status_t = engine_gpu.run_dm(fb, noise=nm)
    assert engine_gpu.stats()["n_onchip_circuits"] == 0
    engine_gpu.set_options()
    assert not status_t.any() and np.max(np.abs(vals - vals_t)) <= 1e-12
    # ideal side: the same kernel without the noise table == the statevector oracle
    # (a reset is not a unitary: status 1 and NaN, as on the statevector path proper)
    ideal, st = engine_gpu.run_sv(fb)
    assert engine_gpu.stats()["n_onchip_circuits"] == len(circs)
    has_reset = [any(g == "reset" for g, _, _ in c.gate_ops()) for c in circs]
    assert st.tolist() == [1 if r else 0 for r in has_reset] and any(has_reset) and not all(has_reset)
    off = np.cumsum([0] + [len(o) for o in obs])
    for i, r in enumerate(has_reset):
        if r:
            assert np.isnan(ideal[off[i]:off[i + 1]]).all()
        else:
            assert np.max(np.abs(ideal[off[i]:off[i + 1]] - helpers.oracle_sv_values(circs[i], obs[i]))) <= TOL, i
    engine_gpu.set_options(flags=NO_ONCHIP)
    ideal_t, st_t = engine_gpu.run_sv(fb)
    engine_gpu.set_options()
    assert st_t.tolist() == st.tolist() and np.nanmax(np.abs(ideal - ideal_t)) <= 1e-12
    both = engine_gpu.run_meas_data(fb, noise=nm)
    assert np.array_equal(both[0], ideal, equal_nan=True) and np.array_equal(both[1], vals)
    assert both[2].tolist() == st.tolist() and not both[3].any()


def test_onchip_status_codes_and_mixed_batches(engine_gpu):
    """Per-circuit failures keep lower_dm_circuit's codes (NaN values); a batch with one circuit the
    kernel does not cover (6 active qubits, or a cz) runs through the tile sweeps as a whole."""
    lima = backends.fake_lima()
    nm = noise.from_backend(lima)
    rng = np.random.default_rng(9)
    good = [_random_circuit(rng, 5, 30, lima.coupling_map) for _ in range(4)]
    obs = [[[(l, 1.0)] for l in _labels(rng, 5, 3)] for _ in range(5)]
    fb = engine.encode_batch(good[:1] + [good[0]] + good[1:], obs)
    # corrupt the first op of circuit 1 (not one of the circuits the host probes): qubit 9 of a 5-qubit register
    fb.ops = fb.ops.copy()
    g = int(fb.op_offsets[1])
    fb.ops["q0"][g] = 9
    vals, status = engine_gpu.run_dm(fb, noise=nm)
    assert engine_gpu.stats()["n_onchip_circuits"] == 5
    ok = np.r_[0:3, 6:15]
    assert status.tolist() == [0, 3, 0, 0, 0] and np.isnan(vals[3:6]).all() and not np.isnan(vals[ok]).any()
    engine_gpu.set_options(flags=NO_ONCHIP)
    vals_t, status_t = engine_gpu.run_dm(fb, noise=nm)
    engine_gpu.set_options()
    assert status_t.tolist() == status.tolist() and np.max(np.abs(vals[ok] - vals_t[ok])) <= 1e-12
    # mixed: a 6-qubit ladder next to small circuits; a cz next to small circuits
    be = backends.synthetic_chain(6, seed=3)
    nm6 = noise.from_backend(be)
    wide = F.tfim_circuit(6, 2, 0.4)
    small = [_random_circuit(rng, 6, 25, [(0, 1), (1, 2)]) for _ in range(4)]
    obs6 = [[[(l, 1.0)] for l in _labels(rng, 6, 3)] for _ in range(5)]
    from oracle import noise_model as onm
    on6 = onm.from_backend(be.to_dict())
    czc = Circuit(6); czc.h(0); czc.cz(0, 1); czc.sx(1)
    # position 1 is not among the circuits the host probes (0, N/2, N-1): the kernel itself reports it
    for odd in (wide, czc):
        for mixed in (small + [odd], small[:1] + [odd] + small[1:]):
            vals, status = engine_gpu.run_dm(engine.encode_batch(mixed, obs6), noise=nm6)
            assert engine_gpu.stats()["n_onchip_circuits"] == 0 and engine_gpu.stats()["n_sweep_launches"] > 0 and not status.any()
            ref = np.concatenate([helpers.oracle_dm_values(c, o, on6) for c, o in zip(mixed, obs6)])
            assert np.max(np.abs(vals - ref)) <= TOL


def test_onchip_variants_match_python_built_variants(lib):
    """Folds and twirls drawn inside the kernel (same counter-based generator as variants.cpp) ==
    the variants built circuit by circuit in Python and run on the tile-sweep path."""
    from test_variants import _Replay

    eng = Engine(0)
    lima = backends.fake_lima()
    nm = noise.from_backend(lima)
    seed = 1234
    steps = (1, 2, 3, 4)
    base = [F.tfim_circuit(4, s, 0.3 + 0.1 * s, layout=[0, 1, 3, 4], num_physical=5, basis="XYZ"[s % 3]) for s in steps]
    obs = [F.single_z_observables([0, 1, 3, 4], 5)] * len(base)
    fb = engine.encode_batch(base, obs)
    v = Variants(folds=(1, 3, 5), twirls=3, seed=seed)
    vals, st = eng.run_dm_variants(fb, v, noise=nm)
    assert eng.stats()["n_onchip_circuits"] == len(base) * 9 and not st.any()
    ref_circs = [F.tfim_circuit(4, s, 0.3 + 0.1 * s, layout=[0, 1, 3, 4], num_physical=5, basis="XYZ"[s % 3], fold=fold,
                                twirl_rng=_Replay(seed, c, t)) for c, s in enumerate(steps) for fold in (1, 3, 5) for t in range(3)]
    eng.set_options(flags=NO_ONCHIP)
    ref, st = eng.run_dm(engine.encode_batch(ref_circs, [obs[0]] * len(ref_circs)), noise=nm)
    vals_lib, st2 = eng.run_dm_variants(fb, v, noise=nm)   # host expansion + lowering + sweeps
    assert eng.stats()["n_onchip_circuits"] == 0
    eng.set_options()
    assert not st.any() and not st2.any()
    assert np.max(np.abs(vals - ref)) <= 1e-12 and np.max(np.abs(vals - vals_lib)) <= 1e-12
    on = helpers.oracle_noise("fakelima")
    for k in (7, 20, 35):
        assert np.max(np.abs(vals[4 * k:4 * k + 4] - helpers.oracle_dm_values(ref_circs[k], obs[0], on))) <= TOL
    # folds only (cfg1's shape) and the (ideal, noisy) call
    vf = Variants(folds=(1, 3, 5))
    ideal, noisy, st_i, st_n = eng.run_meas_data_variants(fb, vf, noise=nm)
    assert eng.stats()["n_onchip_circuits"] == len(base) * (3 + 1) and not st_i.any() and not st_n.any()  # + the ideal warps
    eng.set_options(flags=NO_ONCHIP)
    ideal_t, noisy_t, _, _ = eng.run_meas_data_variants(fb, vf, noise=nm)
    eng.set_options()
    assert np.max(np.abs(noisy - noisy_t)) <= 1e-12 and np.max(np.abs(ideal - ideal_t)) <= 1e-12
    eng.close()


def test_onchip_ranges_pipeline_gives_identical_values(lib, monkeypatch):
    """Large batches are cut into ranges of circuits (range r+1 is staged and uploaded while the kernel
    of range r runs); forced here on a small batch: values, statuses and twirl draws (a function of the
    BATCH index of a circuit) do not depend on the number of ranges."""
    lima = backends.fake_lima()
    nm = noise.from_backend(lima)
    rng = np.random.default_rng(21)
    circs = [_random_circuit(rng, 5, int(rng.integers(5, 60)), lima.coupling_map) for _ in range(37)]
    obs = [[[(l, 1.0)] for l in _labels(rng, 5, int(rng.integers(1, 4)))] for _ in circs]
    fb = engine.encode_batch(circs, obs)
    v = Variants(folds=(1, 3), twirls=3, seed=5)
    ref = None
    for ranges in ("1", "3", "8"):
        monkeypatch.setenv("BWQ_ONCHIP_RANGES", ranges)
        eng = Engine(0)
        a = eng.run_meas_data_variants(fb, v, noise=nm)
        assert eng.stats()["n_onchip_circuits"] == len(circs) * 7 and eng.stats()["n_other_launches"] == int(ranges)
        b = eng.run_dm(fb, noise=nm)
        c = eng.run_sv(fb)
        eng.close()
        got = [np.asarray(x) for x in (*a, *b, *c)]
        if ref is None:
            ref = got
        else:
            for x, y in zip(got, ref):
                assert np.array_equal(x, y, equal_nan=True), ranges
