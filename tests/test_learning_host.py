"""Host logic of ml_qem_b200.learning: the batched processors must reproduce the
reference's per-(circuit, Pauli term) loop (blackwater/library/learning/estimator.py:128-148,
:168-187, :220-247), and learning() must keep the decorator contract (:300-328)."""
import numpy as np
import pytest

from ml_qem_b200 import backends, families as F
from ml_qem_b200 import learning as L
from ml_qem_b200.estimator import EstimatorResult
from ml_qem_b200.features import backend_properties_v1, encode_data


class _Job:
    def __init__(self, values):
        self._values = values
        self.submitted = 0

    def job_id(self):
        return "job-7"

    def result(self):
        return EstimatorResult(np.array(self._values), [{"shots": None} for _ in self._values])

    def submit(self):
        self.submitted += 1

    def status(self):
        return "DONE"

    def cancel(self):
        return False


class _FakeEstimator:
    """Stands in for the engine's Estimator on a CPU-only box: fixed values, records the call."""

    def __init__(self, values):
        self.values = values
        self.calls = []

    def run(self, circuits, observables, parameter_values=None, **opts):
        return self._run(tuple(circuits), tuple(observables), tuple(parameter_values or [()] * len(circuits)), **opts)

    def _run(self, circuits, observables, parameter_values, **run_options):
        self.calls.append((circuits, observables, parameter_values, run_options))
        return _Job(self.values)


def _workload(rng, n_circ=9):
    lima = backends.fake_lima()
    circs = [F.random_basis_circuit(5, int(rng.integers(3, 40)), rng, lima.coupling_map) for _ in range(n_circ)]
    obs = []
    for _ in circs:
        k = int(rng.integers(1, 4))
        obs.append([("".join(rng.choice(list("IXYZ"), size=5)), float(rng.normal())) for _ in range(k)])
    vals = rng.uniform(-1, 1, size=n_circ)
    return lima, circs, obs, vals


def test_encode_pauli_sum_op_layout():
    rows = L.encode_pauli_sum_op([("XYZI", 0.5), ("IIZZ", -2.0)])
    assert rows[0] == [0.5, 0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 0]
    assert rows[1][0] == -2.0 and len(rows[1]) == 17


def test_scikit_processor_batch_equals_reference_loop():
    from sklearn.ensemble import RandomForestRegressor

    rng = np.random.default_rng(3)
    lima, circs, obs, vals = _workload(rng)
    props = backend_properties_v1(lima)
    width = encode_data([circs[0]], props, [[0.0]], [[0.1]], 1, L.encode_pauli_sum_op([("IIIIZ", 1.0)])).__getitem__(0).shape[1]
    model = RandomForestRegressor(n_estimators=8, random_state=0).fit(rng.normal(size=(64, width)), rng.normal(size=64))
    proc = L.ScikitLearningModelProcessor(model, lima)
    loop = np.array([proc.process(v, c, o, ()) for v, c, o in zip(vals, circs, obs)])
    batch = proc.process_batch(vals, circs, obs, [()] * len(circs))
    assert np.max(np.abs(loop - batch)) < 1e-12
    # the reference's arithmetic for one item, spelled out: sum_k coeff_k * model(features(term k))
    want = 0.0
    for label, coeff in obs[2]:
        X, _ = encode_data([circs[2]], props, [[0.0]], [[float(vals[2])]], 1, L.encode_pauli_sum_op([(label, 1.0)]))
        want += model.predict(X.numpy()).item() * coeff
    assert abs(want - batch[2]) < 1e-12


def test_torch_processor_batch_equals_loop_and_learning_decorator():
    import torch

    rng = np.random.default_rng(4)
    lima, circs, obs, vals = _workload(rng, n_circ=6)
    props = backend_properties_v1(lima)
    width = encode_data([circs[0]], props, [[0.0]], [[0.1]], 1, L.encode_pauli_sum_op([("IIIIZ", 1.0)]))[0].shape[1]
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(width, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
    proc = L.TorchLearningModelProcessor(model, lima)
    loop = np.array([proc.process(v, c, o, ()) for v, c, o in zip(vals, circs, obs)])
    assert np.max(np.abs(loop - proc.process_batch(vals, circs, obs, [()] * 6))) < 1e-6  # float32 model

    LearningEst = L.learning(_FakeEstimator, proc, skip_transpile=True, backend=lima)
    assert LearningEst.__name__ == "Learning_FakeEstimator" and issubclass(LearningEst, _FakeEstimator)
    est = LearningEst(vals)
    job = est.run(circs, obs)
    circuits, observables, parameter_values, _ = est.calls[0]  # the original _run saw keyword arguments
    assert len(circuits) == 6 and parameter_values == ((),) * 6
    assert job.job_id() == "job-7" and job.status() == "DONE" and job.cancel() is False
    res = job.result()
    assert np.max(np.abs(res.values - loop)) < 1e-6
    assert [m["original_value"] for m in res.metadata] == list(vals) and all("shots" in m for m in res.metadata)


def test_empty_processor_and_bad_observable():
    rng = np.random.default_rng(5)
    lima, circs, obs, vals = _workload(rng, n_circ=3)
    Est = L.learning(_FakeEstimator, L.EmptyProcessor())
    res = Est(vals).run(circs, obs).result()
    assert np.array_equal(res.values, vals)
    with pytest.raises(ValueError):
        Est(vals).run(circs, [object()] * 3).result()
