"""The C++/OpenMP Aer-style restatement (bench cpu_baseline) against the numpy oracle."""
import numpy as np
import pytest

import helpers
from ml_qem_b200 import backends, engine, families as F
from ml_qem_b200.gateset import OPCODES
from oracle import cpu_ref, noise_model as onm


@pytest.fixture(scope="module")
def built():
    cpu_ref.build()
    return True


def _labels(rng, n, k):
    return ["".join(rng.choice(list("IXYZ"), size=n)) for _ in range(k)]


@pytest.mark.parametrize("fusion_threshold", [99, 1])
def test_dm_matches_numpy_oracle(built, fusion_threshold):
    lima = backends.fake_lima()
    on = helpers.oracle_noise("fakelima")
    rng = np.random.default_rng(3)
    circs = [F.random_basis_circuit(5, int(rng.integers(0, 50)), rng, lima.coupling_map) for _ in range(10)]
    circs.append(F.tfim_circuit(4, 2, 0.3, basis="Y", layout=[0, 1, 3, 4], num_physical=5, fold=3, random_init_prefix=True))
    obs = [[[(l, float(rng.normal()))] for l in _labels(rng, 5, 5)] for _ in circs]
    fb = engine.encode_batch(circs, obs)
    ref = np.concatenate([helpers.oracle_dm_values(c, o, on) for c, o in zip(circs, obs)])
    vals, status = cpu_ref.run_dm(fb, cpu_ref.noise_arrays(on, OPCODES), threads=2, fusion_threshold=fusion_threshold)
    assert not status.any()
    assert np.max(np.abs(vals - ref)) < 1e-12


def test_dm_amplitude_parallel_and_non_basis_gates(built):
    from ml_qem_b200 import Circuit
    c = Circuit(4)
    c.h(0); c.cz(0, 1); c.u3(0.3, 0.2, -0.7, 2); c.swap(1, 2); c.rzz(0.4, 2, 3); c.ecr(3, 0); c.t(1); c.sdg(2)
    c.crx(0.9, 1, 3); c.reset(0); c.ry(0.5, 0); c.cp(1.1, 0, 2); c.y(3); c.rzx(0.6, 1, 2); c.iswap(0, 3); c.cy(2, 0)
    rng = np.random.default_rng(5)
    obs = [[(l, 1.0)] for l in _labels(rng, 4, 8)]
    fb = engine.encode_batch([c], [obs])
    ref = helpers.oracle_dm_values(c, obs, None)
    for apq, ft in ((1, 99), (1, 1), (99, 99)):
        vals, status = cpu_ref.run_dm(fb, None, threads=2, amplitude_parallel_qubits=apq, fusion_threshold=ft)
        assert not status.any() and np.max(np.abs(vals - ref)) < 1e-12


def test_sv_matches_numpy_oracle(built):
    rng = np.random.default_rng(9)
    circs = [F.tfim_circuit(8, 3, 0.4, basis="X"), F.brickwork_circuit(6, 2, rng, num_physical=8)]
    obs = [F.tfim_observables(list(range(8)), 8), F.single_z_observables(list(range(6)), 8)]
    fb = engine.encode_batch(circs, obs)
    ref = np.concatenate([helpers.oracle_sv_values(*helpers.compact(c, o)) for c, o in zip(circs, obs)])
    for apq in (1, 99):
        vals, status = cpu_ref.run_sv(fb, threads=2, amplitude_parallel_qubits=apq)
        assert not status.any() and np.max(np.abs(vals - ref)) < 1e-12
