"""-m gpu: parity cases the round-1 review found untested on the GPU -- all through the C ABI, against
the numpy oracle (or, at 12 qubits, the Aer-style C++ restatement).  Tolerance 1e-10."""
import threading

import numpy as np
import pytest

import helpers
from ml_qem_b200 import Circuit, backends, engine, families as F, noise
from ml_qem_b200.circuit import parse_qasm
from ml_qem_b200.estimator import B200Estimator

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _labels(rng, n, k):
    return ["".join(rng.choice(list("IXYZ"), size=n)) for _ in range(k)]


def test_all_qubit_default_noise_entries(engine_gpu):
    """The all-qubit default error (q0 = 255 in the noise table): modify_and_add_noise_to_model
    (docs/tutorials/mbd_utils.py:95-137) and AddNoise.add_coherent_noise(uniform=True)
    (docs/tutorials/noise_utils.py:116-123) against the oracle's restatement of both."""
    from oracle import noise_model as onm

    lima = backends.fake_lima()
    props = helpers.golden("backends.json")["fakelima"]
    rng = np.random.default_rng(21)
    circs = [F.random_basis_circuit(5, 60, rng, lima.coupling_map) for _ in range(6)]
    circs.append(F.tfim_circuit(4, 2, 0.4, basis="X", layout=[0, 1, 3, 4], num_physical=5))
    obs = [[[(l, float(rng.normal()))] for l in _labels(rng, 5, 5)] for _ in circs]
    batch = engine.encode_batch(circs, obs)
    cases = [
        (noise.modify_and_add_noise_to_model(lima, theta=np.pi / 8), onm.modify_and_add_noise_to_model(props, theta=np.pi / 8)),
        (noise.add_coherent_noise(lima, theta=0.04 * np.pi, uniform=True)[0],
         onm.add_coherent_noise(props, theta=0.04 * np.pi, uniform=True)),
        (noise.add_coherent_noise(lima, theta=0.04 * np.pi, uniform=True, add_depolarization=False)[0],
         onm.add_coherent_noise(props, theta=0.04 * np.pi, uniform=True, add_depolarization=False)),
    ]
    for nm, on in cases:
        on = on[0] if isinstance(on, tuple) else on
        assert nm.default and "cx" in nm.default and not any(k[0] == "cx" for k in nm.local)
        ref = np.concatenate([helpers.oracle_dm_values(c, o, on) for c, o in zip(circs, obs)])
        vals, status = engine_gpu.run_dm(batch, noise=nm)
        assert not status.any()
        assert np.max(np.abs(vals - ref)) <= TOL
    # a local entry for the exact ordered pair overrides the default (Aer's lookup order)
    nm = noise.modify_and_add_noise_to_model(lima, theta=np.pi / 8)
    on = onm.modify_and_add_noise_to_model(props, theta=np.pi / 8)
    full, ofull = noise.from_backend(lima), onm.from_backend(props)
    nm.add_quantum_error(full.get("cx", (0, 1)), "cx", (0, 1))
    on.local[("cx", (0, 1))] = ofull.local[("cx", (0, 1))]
    ref = np.concatenate([helpers.oracle_dm_values(c, o, on) for c, o in zip(circs, obs)])
    vals, status = engine_gpu.run_dm(batch, noise=nm)
    assert not status.any() and np.max(np.abs(vals - ref)) <= TOL


def test_run_dm_into_torch_tensor_equals_run_dm(engine_gpu):
    """bwq_dm_run_device_out (zero-copy label hand-off into a torch CUDA tensor), unsegmented and
    pipelined (>= 256 circuits of 7 qubits), and bwq_dm_execute_device_out."""
    import ctypes as C

    import torch

    be = backends.synthetic_chain(8, seed=4)
    nm = noise.from_backend(be)
    rng = np.random.default_rng(5)
    for n_circ, flags in ((6, 0), (300, 8)):  # 8 = BWQ_OPT_FORCE_PIPELINE
        circs = [F.tfim_circuit(7, 1 + i % 3, float(rng.uniform(0, 1)), basis="XYZ"[i % 3], num_physical=8) for i in range(n_circ)]
        circs[2] = Circuit(8)  # a gate-free circuit inside (value fixed up by the host)
        obs = [F.tfim_observables(list(range(7)), 8)[:5] for _ in circs]
        batch = engine.encode_batch(circs, obs)
        engine_gpu.set_options(flags=flags)
        ref, status = engine_gpu.run_dm(batch, noise=nm)
        assert not status.any()
        out = torch.full((batch.n_observables,), float("nan"), dtype=torch.float64, device="cuda")
        status = engine_gpu.run_dm_into(batch, out.data_ptr(), noise=nm)
        torch.cuda.synchronize()
        assert not status.any()
        assert np.array_equal(out.cpu().numpy(), ref)
    engine_gpu.set_options()
    # prepared program + device output
    st = engine_gpu.prepare_dm(batch)
    assert not st.any()
    out2 = torch.zeros(batch.n_observables, dtype=torch.float64, device="cuda")
    engine_gpu._check(engine_gpu._lib.bwq_dm_execute_device_out(engine_gpu._ctx, C.c_void_p(out2.data_ptr())), "bwq_dm_execute_device_out")
    torch.cuda.synchronize()
    assert np.array_equal(out2.cpu().numpy(), ref)


def test_reset_mid_circuit_on_gpu(engine_gpu):
    """reset (not trace-free in the Pauli basis: I -> I + Z) inside noisy circuits, 5 and 8 qubits
    (on chip and tiled), with the device's reset error attached."""
    from oracle import noise_model as onm

    lima = backends.fake_lima()
    on = helpers.oracle_noise("fakelima")
    rng = np.random.default_rng(31)
    circs = []
    for k in range(4):
        c = Circuit(5)
        for _ in range(30):
            r = int(rng.integers(0, 6))
            q = int(rng.integers(0, 5))
            if r == 0:
                c.rz(float(rng.uniform(-3, 3)), q)
            elif r == 1:
                c.sx(q)
            elif r == 2:
                c.x(q)
            elif r == 3:
                c.reset(q)
            else:
                a, b = lima.coupling_map[int(rng.integers(0, len(lima.coupling_map)))]
                c.cx(a, b)
        circs.append(c)
    obs = [[[(l, 1.0)] for l in _labels(rng, 5, 6)] for _ in circs]
    ref = np.concatenate([helpers.oracle_dm_values(c, o, on) for c, o in zip(circs, obs)])
    vals, status = engine_gpu.run_dm(engine.encode_batch(circs, obs), noise=noise.from_backend(lima))
    assert not status.any() and np.max(np.abs(vals - ref)) <= TOL
    be = backends.synthetic_chain(8, seed=8)
    c = F.tfim_circuit(8, 1, 0.6, basis="X")
    c.reset(3); c.sx(3); c.cx(3, 4); c.reset(0); c.cx(0, 1); c.reset(7); c.x(7); c.cx(6, 7)
    ob = F.tfim_observables(list(range(8)), 8)
    ref = helpers.oracle_dm_values(c, ob, onm.from_backend(be.to_dict()))
    for kq in (6, 4):
        engine_gpu.set_options(tile_qubits=kq)
        vals, status = engine_gpu.run_dm(engine.encode_batch([c], [ob]), noise=noise.from_backend(be))
        assert not status.any() and np.max(np.abs(vals - ref)) <= TOL, kq
    engine_gpu.set_options()


def test_stored_fakelima_qasm_circuits_through_the_estimator(lib):
    """The 80 FakeLima-transpiled QASM circuits the reference stores (docs/tutorials/data/
    mbd_datasets2, tests/golden/mbd_sample.json) through B200Estimator -- QASM text in, one run for
    ideal and one for noisy values -- against the oracle on the very same circuits."""
    from oracle import dm, sv

    lima = backends.fake_lima()
    on = helpers.oracle_noise("fakelima")
    entries = helpers.golden("mbd_sample.json")
    noisy_est, ideal_est = B200Estimator(backend=lima), B200Estimator()
    qasm, obs, ref_n, ref_i = [], [], [], []
    for e in entries:
        circ = parse_qasm(e["qasm"])
        meas = [o[1][0] for o in circ.ops if o[0] == "measure"]
        n = circ.num_qubits
        ops = circ.gate_ops()
        rho, psi = dm.simulate(n, ops, on), sv.simulate(n, ops)
        for q in meas:
            label = "".join("Z" if (n - 1 - i) == q else "I" for i in range(n))
            qasm.append(e["qasm"]); obs.append(label)
            ref_n.append(dm.expval_pauli(rho, n, label).real)
            ref_i.append(sv.expval_pauli(psi, n, label).real)
    assert len(qasm) >= 80
    res_n = noisy_est.run(qasm, obs).result()
    res_i = ideal_est.run(qasm, obs).result()
    assert len(res_n.metadata) == len(qasm) and res_n.metadata[0]["simulator_metadata"]["method"] == "density_matrix"
    assert np.max(np.abs(res_n.values - np.array(ref_n))) <= TOL
    assert np.max(np.abs(res_i.values - np.array(ref_i))) <= TOL
    # and they are the stored 10k-shot data up to shot noise once the readout confusion is applied
    # (tests/test_oracle_golden.py pins the oracle to them; here the GPU values take the same check)
    k, worst = 0, 0.0
    for e in entries:
        circ = parse_qasm(e["qasm"])
        meas = [o[1][0] for o in circ.ops if o[0] == "measure"]
        for j, q in enumerate(meas):
            a, b = on.readout[q][0, 1], on.readout[q][1, 0]
            z_meas = (1 - a - b) * res_n.values[k] + (b - a)
            worst = max(worst, abs(-z_meas - e["noisy_exp_values"][0][len(meas) - 1 - j]))
            k += 1
    assert worst < 5 / np.sqrt(10000)


def test_entangled_tfim12_dm_vs_cpu_restatement(engine_gpu):
    """12-qubit TFIM (3 Trotter steps: entangled across the whole chain, 134 MB state, 64 tiles per
    sweep) under chain noise against the Aer-style C++ restatement (oracle/cpu_ref.cpp, itself pinned
    to the numpy oracle in tests/test_cpu_ref.py); multi-Pauli observables incl. XX and Z^12."""
    from ml_qem_b200.gateset import OPCODES
    from oracle import cpu_ref, noise_model as onm

    cpu_ref.build()
    n = 12
    be = backends.synthetic_chain(n, seed=n)
    circs = [F.tfim_circuit(n, 3, 0.37, dt=0.25), F.tfim_circuit(n, 2, 0.81, dt=0.25, basis="Y")]
    obs = [F.tfim_observables(list(range(n)), n)] * 2
    batch = engine.encode_batch(circs, obs)
    ref, st = cpu_ref.run_dm(batch, cpu_ref.noise_arrays(onm.from_backend(be.to_dict()), OPCODES))
    assert not st.any()
    vals, status = engine_gpu.run_dm(batch, noise=noise.from_backend(be))
    assert not status.any()
    assert np.max(np.abs(vals - ref)) <= TOL
    assert np.max(np.abs(vals)) > 0.05  # not a trivially vanishing comparison


@pytest.mark.parametrize("world", [2, 4, 8])
def test_svx_exchange_kernel_same_device_peers(engine_gpu, world):
    """svx_exchange_kernel<PUSH> and <PULL> with all `world` shards on ONE GPU: the peer pointers
    are same-device buffers, so the kernel's block routing (block b of rank r's old shard becomes
    block r of rank b's new shard) is checked bit-exactly without a second GPU."""
    import torch

    n_local = 1 << 14
    blk = n_local // world
    gen = torch.Generator(device="cuda").manual_seed(world)
    old = [torch.randn(n_local, 2, dtype=torch.float64, device="cuda", generator=gen) for _ in range(world)]
    expect = [torch.cat([old[b][r * blk:(r + 1) * blk] for b in range(world)]) for r in range(world)]
    for push in (False, True):
        new = [torch.zeros_like(o) for o in old]
        for r in range(world):
            if push:   # local = rank r's OLD shard, peers = everybody's NEW shard
                engine_gpu.svx_exchange(old[r].data_ptr(), [t.data_ptr() for t in new], r, n_local, push=True)
            else:      # local = rank r's NEW shard, peers = everybody's OLD shard
                engine_gpu.svx_exchange(new[r].data_ptr(), [t.data_ptr() for t in old], r, n_local, push=False)
        engine_gpu.sync()
        for r in range(world):
            assert torch.equal(new[r], expect[r]), (push, r)


def test_estimators_sharing_an_engine_from_two_threads(lib):
    """Two estimators with DIFFERENT noise models on the shared per-device engine, run concurrently
    from two threads (and with ZNE's several runs per call): every value must come from its own
    estimator's noise table (the table is installed and used under one engine lock)."""
    from ml_qem_b200 import zne

    lima = backends.fake_lima()
    est_a = B200Estimator(backend=lima)
    est_b = B200Estimator(noise_model=noise.modify_and_add_noise_to_model(lima, theta=np.pi / 8))
    assert est_a._engine_handle() is est_b._engine_handle()
    on_a = helpers.oracle_noise("fakelima")
    from oracle import noise_model as onm
    on_b = onm.modify_and_add_noise_to_model(helpers.golden("backends.json")["fakelima"], theta=np.pi / 8)
    c = F.tfim_circuit(4, 2, 0.3, layout=[0, 1, 3, 4], num_physical=5)
    ob = [("ZIIIZ", 1.0)]
    ref_a = helpers.oracle_dm_values(c, [ob], on_a)[0]
    ref_b = helpers.oracle_dm_values(c, [ob], on_b)[0]
    assert abs(ref_a - ref_b) > 1e-3
    errs = []

    def worker(est, ref, use_zne):
        try:
            for _ in range(25):
                if use_zne:
                    r = est.run([c] * 3, [ob] * 3, zne_strategy=zne.ZNEStrategy(noise_factors=(1, 3))).result()
                    v = r.metadata[0]["zne"]["noise_amplification"]["values"][0]
                else:
                    v = est.run([c] * 3, [ob] * 3).result().values[1]
                if abs(v - ref) > TOL:
                    errs.append((use_zne, v, ref))
        except Exception as exc:  # noqa: BLE001
            errs.append(exc)

    ts = [threading.Thread(target=worker, args=(est_a, ref_a, True)), threading.Thread(target=worker, args=(est_b, ref_b, False))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs[:3]
    # engine options given to an estimator create a private engine instead of re-configuring the shared one
    est_c = B200Estimator(backend=lima, tile_qubits=4)
    assert est_c._engine_handle() is not est_a._engine_handle()
    assert abs(est_c.run(c, ob).result().values[0] - ref_a) <= TOL


def test_complex_coefficients_return_the_complex_sum(lib):
    """Aer returns np.real_if_close(sum_k c_k Tr(rho P_k)): complex coefficients give a complex value."""
    lima = backends.fake_lima()
    est = B200Estimator(backend=lima)
    c = F.tfim_circuit(4, 1, 0.5, layout=[0, 1, 3, 4], num_physical=5, basis="Y")
    ob = [("ZIIIZ", 0.5 + 0.25j), ("IIIZI", -1j), ("IXIXI", 2.0)]
    on = helpers.oracle_noise("fakelima")
    parts = helpers.oracle_dm_values(c, [[("ZIIIZ", 1.0)], [("IIIZI", 1.0)], [("IXIXI", 1.0)]], on)
    ref = (0.5 + 0.25j) * parts[0] - 1j * parts[1] + 2.0 * parts[2]
    res = est.run([c, c], [ob, "-IIIZI"]).result()
    assert np.iscomplexobj(res.values)
    assert abs(res.values[0] - ref) <= TOL and abs(res.values[1] - (-parts[1])) <= TOL


def test_tma_kernel_vs_classic_kernel_and_oracle(engine_gpu, monkeypatch):
    """dm_sweep_tma_kernel (default for circuits wider than the tile) against dm_sweep_kernel
    (BWQ_OPT_NO_TMA) and the oracle: chain circuits (straight-line pass bodies), random circuits
    (generic bodies, passes that target slot 0), coherent cx errors (dense ops: the FULL
    instantiation), several circuits per launch with different tile layouts, gate-free circuit."""
    from oracle import noise_model as onm

    rng = np.random.default_rng(77)
    n = 8
    be = backends.synthetic_chain(n, seed=21)
    nm, on = noise.from_backend(be), onm.from_backend(be.to_dict())
    chain = [(i, i + 1) for i in range(n - 1)] + [(i + 1, i) for i in range(n - 1)]
    circs = [F.tfim_circuit(n, 1 + k % 4, float(rng.uniform(0, 1)), basis="XYZ"[k % 3]) for k in range(6)]
    circs += [F.brickwork_circuit(n, 1 + k % 3, np.random.default_rng(k), twirl_rng=rng) for k in range(5)]
    circs += [F.random_basis_circuit(n, 120, rng, chain) for _ in range(5)]
    circs += [F.tfim_circuit(7, 2, 0.3, num_physical=n), Circuit(n)]
    obs = [[[(l, float(rng.normal()))] for l in _labels(rng, n, 5)] for _ in circs]
    batch = engine.encode_batch(circs, obs)
    ref = np.concatenate([helpers.oracle_dm_values(c, o, on) for c, o in zip(circs, obs)])
    engine_gpu.set_options()
    v_tma, st = engine_gpu.run_dm(batch, noise=nm)
    s_tma = engine_gpu.stats()
    assert not st.any() and s_tma["n_tma_sweep_launches"] > 0
    engine_gpu.set_options(flags=16)
    v_cls, st = engine_gpu.run_dm(batch, noise=nm)
    assert not st.any() and engine_gpu.stats()["n_tma_sweep_launches"] == 0
    for fl in (64, 32, 32 | 64):  # direct last-pass stores; persistent double-buffered kernel; both
        engine_gpu.set_options(flags=fl)
        v_alt, st = engine_gpu.run_dm(batch, noise=nm)
        assert not st.any() and np.max(np.abs(v_alt - v_tma)) <= 1e-13, fl
    engine_gpu.set_options()
    assert np.max(np.abs(v_tma - ref)) <= TOL and np.max(np.abs(v_cls - ref)) <= TOL
    assert np.max(np.abs(v_tma - v_cls)) <= 1e-13
    # several tensor-map windows (forced: normally 2^31 elements per map) and prepared/resident reruns
    monkeypatch.setenv("BWQ_TMA_WINDOW_LOG2", "1")
    st = engine_gpu.prepare_dm(batch)
    assert not st.any()
    assert np.array_equal(engine_gpu.execute_dm(), v_tma) and np.array_equal(engine_gpu.execute_dm(), v_tma)
    monkeypatch.delenv("BWQ_TMA_WINDOW_LOG2")
    # dense two-qubit ops (coherent cx error, non-basis gates) in the TMA layout: FULL instantiation
    lima7 = backends.synthetic_chain(7, seed=3)
    nm2 = noise.modify_and_add_noise_to_model(lima7, theta=np.pi / 8)
    on2 = onm.modify_and_add_noise_to_model(lima7.to_dict(), theta=np.pi / 8)
    c = F.tfim_circuit(7, 2, 0.55, basis="Y")
    c.rzz(0.4, 2, 3); c.ecr(5, 6); c.h(0); c.swap(0, 1)
    ob = F.tfim_observables(list(range(7)), 7)
    ref2 = helpers.oracle_dm_values(c, ob, on2)
    v2, st = engine_gpu.run_dm(engine.encode_batch([c], [ob]), noise=nm2)
    assert not st.any() and engine_gpu.stats()["n_tma_sweep_launches"] > 0
    assert np.max(np.abs(v2 - ref2)) <= TOL


def test_library_variants_on_gpu_match_python_built_variants(lib):
    """bwq_dm_run_variants / bwq_meas_data_run_variants / B200Estimator(variants=...): the folds and
    twirls generated inside the library give the values of the same variants built in Python (and
    of the oracle on one of them); the ideal side runs the base circuits only."""
    from test_variants import _Replay
    from ml_qem_b200 import zne
    from ml_qem_b200.engine import Engine, Variants

    eng = Engine(0)
    lima = backends.fake_lima()
    nm = noise.from_backend(lima)
    seed = 99
    base = [F.tfim_circuit(4, s, 0.3 + 0.1 * s, layout=[0, 1, 3, 4], num_physical=5, basis="XYZ"[s % 3]) for s in (1, 2, 3)]
    obs = [F.single_z_observables([0, 1, 3, 4], 5)] * 3
    fb = engine.encode_batch(base, obs)
    v = Variants(folds=(1, 3), twirls=4, seed=seed)
    vals, st = eng.run_dm_variants(fb, v, noise=nm)
    assert not st.any() and vals.shape == (3 * 8 * 4,)
    ref_circs = [F.tfim_circuit(4, s, 0.3 + 0.1 * s, layout=[0, 1, 3, 4], num_physical=5, basis="XYZ"[s % 3], fold=fold,
                                twirl_rng=_Replay(seed, c, t)) for c, s in enumerate((1, 2, 3)) for fold in (1, 3) for t in range(4)]
    ref, st = eng.run_dm(engine.encode_batch(ref_circs, [obs[0]] * len(ref_circs)), noise=nm)
    assert not st.any() and np.max(np.abs(vals - ref)) <= 1e-12
    on = helpers.oracle_noise("fakelima")
    assert np.max(np.abs(vals[5 * 4:6 * 4] - helpers.oracle_dm_values(ref_circs[5], obs[0], on))) <= TOL
    # folds only: the fold-aware lowering (gates lowered once per base circuit) == the folded circuits
    vf = Variants(folds=(1, 3, 5))
    vals_f, st = eng.run_dm_variants(fb, vf, noise=nm)
    ref_f = np.stack([eng.run_dm(zne.fold_batch(fb, f), noise=nm)[0].reshape(3, 4) for f in (1, 3, 5)], axis=1)
    assert not st.any() and np.max(np.abs(vals_f.reshape(3, 3, 4) - ref_f)) <= 1e-13
    ideal, noisy, st_i, st_n = eng.run_meas_data_variants(fb, v, noise=nm)
    assert not st_i.any() and not st_n.any() and np.array_equal(noisy, vals)
    assert np.max(np.abs(ideal - eng.run_sv(fb)[0])) == 0.0
    # estimator: twirl averages, ZNE over the fold means, every (fold, twirl) value in the metadata
    est = B200Estimator(backend=lima)
    pairs_c = [c for c in base for _ in range(4)]
    pairs_o = [o for _ in base for o in obs[0]]
    res = est.run(pairs_c, pairs_o, variants=v).result()
    grid = vals.reshape(3, 2, 4, 4)  # circuit, fold, twirl, observable
    assert np.max(np.abs(res.values - grid[:, 0].mean(axis=1).reshape(-1))) <= 1e-12
    assert np.allclose(res.metadata[5]["variants"]["values"], grid[1, :, :, 1])
    res_z = est.run(pairs_c, pairs_o, variants=v, zne_strategy=zne.ZNEStrategy(noise_factors=(1, 3))).result()
    m = grid.mean(axis=2)  # [circuit, fold, obs]
    assert np.max(np.abs(res_z.values - (1.5 * m[:, 0] - 0.5 * m[:, 1]).reshape(-1))) <= 1e-10
    eng.close()


def test_random_circuits_reach_every_density_matrix_path(engine_gpu):
    """tools/fuzz_parity.py: random circuits of 2..9 qubits (every gate kind, resets, idle qubits,
    multi-Pauli observables) on random synthetic backends; dm_onchip_kernel, the single-tile kernel and
    the TMA sweeps must each be exercised and agree with the oracle to 1e-10 (asserted inside)."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(os.path.dirname(__file__), "..", "tools", "fuzz_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    paths, worst = mod.run(n_cases=60, seed=11, eng=engine_gpu)
    assert all(paths.get(k, 0) > 0 for k in ("onchip", "sweep", "tma")), paths
    assert worst["dm"] <= TOL and worst["sv"] <= TOL
