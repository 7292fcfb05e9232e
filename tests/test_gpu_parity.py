"""-m gpu: CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Tolerance: |delta <O>| <= 1e-10 (BASELINE.json north_star)."""
import numpy as np
import pytest

import helpers
from ml_qem_b200 import backends, engine, families as F, noise

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _labels(rng, n, k):
    return ["".join(rng.choice(list("IXYZ"), size=n)) for _ in range(k)]


def test_dm_random_lima_all_tilings(engine_gpu):
    lima = backends.fake_lima()
    nm = noise.from_backend(lima)
    on = helpers.oracle_noise("fakelima")
    rng = np.random.default_rng(11)
    circs = [F.random_basis_circuit(5, int(rng.integers(1, 80)), rng, lima.coupling_map) for _ in range(24)]
    obs = [[[(l, float(rng.normal()))] for l in _labels(rng, 5, 6)] for _ in circs]
    batch = engine.encode_batch(circs, obs)
    ref = np.concatenate([helpers.oracle_dm_values(c, o, on) for c, o in zip(circs, obs)])
    engine_gpu.set_noise(nm)
    for kq, low in ((6, 2), (5, 2), (4, 2), (3, 1), (2, 1), (4, 1)):
        engine_gpu.set_options(tile_qubits=kq, low_qubits=low)
        vals, status = engine_gpu.run_dm(batch)
        assert not status.any()
        assert np.max(np.abs(vals - ref)) <= TOL, (kq, low)
    engine_gpu.set_options()


@pytest.mark.parametrize("n", [6, 7, 8])
def test_dm_tfim_chain_vs_oracle(engine_gpu, n):
    be = backends.synthetic_chain(n, seed=n)
    nm = noise.from_backend(be)
    from oracle import noise_model as onm
    on = onm.from_backend(be.to_dict())
    rng = np.random.default_rng(n)
    circs = [F.tfim_circuit(n, s, float(rng.uniform(0, 1)), basis="XYZ"[s % 3]) for s in (1, 2, 3)]
    obs = [F.tfim_observables(list(range(n)), n) for _ in circs]
    ref = np.concatenate([helpers.oracle_dm_values(c, o, on) for c, o in zip(circs, obs)])
    engine_gpu.set_noise(nm)
    for kq in (6, 7, 4):
        engine_gpu.set_options(tile_qubits=kq)
        vals, status = engine_gpu.run_dm(engine.encode_batch(circs, obs))
        assert not status.any()
        assert np.max(np.abs(vals - ref)) <= TOL, kq
    engine_gpu.set_options()


def test_dm_brickwork_twirled_vs_oracle(engine_gpu):
    n = 6
    be = backends.synthetic_chain(8, seed=3)
    nm = noise.from_backend(be)
    from oracle import noise_model as onm
    on = onm.from_backend(be.to_dict())
    rng = np.random.default_rng(5)
    circs = [F.brickwork_circuit(n, 2, np.random.default_rng(9), num_physical=8, twirl_rng=rng) for _ in range(4)]
    obs = [F.single_z_observables(list(range(n)), 8) for _ in circs]
    ref = []
    for c, o in zip(circs, obs):
        cc, oo, on2 = helpers.compact(c, o, on)
        ref.append(helpers.oracle_dm_values(cc, oo, on2))
    engine_gpu.set_noise(nm)
    vals, status = engine_gpu.run_dm(engine.encode_batch(circs, obs))
    assert not status.any()
    assert np.max(np.abs(vals - np.concatenate(ref))) <= TOL


def test_dm_coherent_noise_dense2(engine_gpu):
    lima = backends.fake_lima()
    nm, _ = noise.add_coherent_noise(lima, theta=0.04 * np.pi, seed=0)
    from oracle import noise_model as onm
    on = onm.add_coherent_noise(helpers.golden("backends.json")["fakelima"], theta=0.04 * np.pi, seed=0)
    rng = np.random.default_rng(2)
    circs = [F.random_basis_circuit(5, 50, rng, lima.coupling_map) for _ in range(8)]
    obs = [[[(l, 1.0)] for l in _labels(rng, 5, 5)] for _ in circs]
    ref = np.concatenate([helpers.oracle_dm_values(c, o, on) for c, o in zip(circs, obs)])
    engine_gpu.set_noise(nm)
    vals, status = engine_gpu.run_dm(engine.encode_batch(circs, obs))
    assert not status.any()
    assert np.max(np.abs(vals - ref)) <= TOL


def test_sv_vs_oracle(engine_gpu):
    rng = np.random.default_rng(4)
    circs, obs = [], []
    for n in (1, 2, 5, 9, 13):
        cm = [(i, i + 1) for i in range(n - 1)] + [(i + 1, i) for i in range(n - 1)] or [(0, 0)]
        c = F.random_basis_circuit(n, 60, rng, cm) if n > 1 else F.tfim_circuit(1, 2, 0.3)
        circs.append(c)
        obs.append([[(l, float(rng.normal()))] for l in _labels(rng, n, 5)])
    circs.append(F.tfim_circuit(10, 3, 0.4, basis="Y"))
    obs.append(F.tfim_observables(list(range(10)), 10))
    ref = np.concatenate([helpers.oracle_sv_values(c, o) for c, o in zip(circs, obs)])
    vals, status = engine_gpu.run_sv(engine.encode_batch(circs, obs))
    assert not status.any()
    assert np.max(np.abs(vals - ref)) <= TOL


def test_dm_equals_sv_without_noise(engine_gpu):
    rng = np.random.default_rng(8)
    n = 9
    circs = [F.tfim_circuit(n, 2, 0.7, basis="X")]
    obs = [F.tfim_observables(list(range(n)), n)]
    b = engine.encode_batch(circs, obs)
    engine_gpu.set_noise(None)
    v_dm, _ = engine_gpu.run_dm(b)
    v_sv, _ = engine_gpu.run_sv(b)
    assert np.max(np.abs(v_dm - v_sv)) <= TOL
