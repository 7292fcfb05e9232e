"""Learning-based mitigation on top of the engine's Estimator.

Restates blackwater/library/learning/estimator.py on plain data:

  learning(cls, processor, ...)     :300-328  subclass the estimator, wrap ``_run`` (patch_run :262-298)
  PostProcessedJob                  :194-256  result(): base values -> processor, metadata gets
                                              ``original_value``
  LearningMethodEstimatorProcessor  :20-30    process(expectation_value, circuits, observables, parameter_values)
  ScikitLearningModelProcessor      :88-148   per Pauli term: encode_data(num_qubits=1, noisy value,
                                              one-hot basis) -> model.predict -> x coeff, summed
  TorchLearningModelProcessor       :151-187  same with ``model(X)``
  EmptyProcessor                    :190-195
  encode_pauli_sum_op               blackwater/data/utils.py:447-474

The reference evaluates the model once per (circuit, Pauli term) inside a Python loop
(estimator.py:128-146, :220-247).  ``process_batch`` builds the feature rows of the WHOLE result
with one ``encode_data`` call and runs ONE ``model.predict`` / forward pass; the per-item
``process`` keeps the reference's semantics and is what the batched path is tested against.
Transpilation is the caller's business here (no qiskit in the image): circuits are taken as given,
i.e. the reference's ``skip_transpile=True`` behaviour; parameters are bound when the circuit
object knows how.
"""
from functools import wraps

import numpy as np

from . import observable as observable_mod
from .estimator import EstimatorResult
from .features import backend_properties_v1, encode_data

_ONE_HOT = {"X": [0, 0, 0, 1], "Y": [0, 0, 1, 0], "Z": [0, 1, 0, 0], "I": [1, 0, 0, 0]}


def encode_pauli_sum_op(op):
    """[[coeff, one-hot(char 0), one-hot(char 1), ...], ...] per Pauli term, label read left to
    right (utils.py:447-474).  ``op``: PauliObservable, label string, or [(label, coeff), ...]."""
    rows = []
    for label, coeff in observable_mod.from_any(op):
        row = [float(np.real(coeff))]
        for ch in label:
            row += _ONE_HOT.get(ch, [0, 0, 0, 0])
        rows.append(row)
    return rows


def _bind(circuit, params):
    if params is None or len(params) == 0:
        return circuit
    for name in ("assign_parameters", "bind_parameters"):
        if hasattr(circuit, name):
            return getattr(circuit, name)(list(params))
    raise ValueError("circuit has parameter values but cannot bind them")


class LearningMethodEstimatorProcessor:
    """Post-processing of expectation values (estimator.py:20-30)."""

    def process(self, expectation_value, circuits, observables, parameter_values):
        raise NotImplementedError

    def process_batch(self, values, circuits, observables, parameter_values):
        """All results of one job at once; default = the reference's per-item loop."""
        return np.array([self.process(v, c, o, p) for v, c, o, p in zip(values, circuits, observables, parameter_values)])


class EmptyProcessor(LearningMethodEstimatorProcessor):
    def process(self, expectation_value, circuits, observables, parameter_values):
        return expectation_value

    def process_batch(self, values, circuits, observables, parameter_values):
        return np.asarray(values)


class ZNEProcessor(LearningMethodEstimatorProcessor):
    """estimator.py:33-86: the post-processed value is the zero-noise extrapolation of the SAME circuit,
    obtained from a second estimator call with a ``zne_strategy`` (the reference: ``zne(BackendEstimator)``
    from prototype-zne; here any estimator of this package -- the folded variants run as one GPU batch).
    The reference transpiles the circuit for the backend and pads the observable to its five physical
    qubits for a 2-qubit measurement (estimator.py:52-80); circuits and observables are taken on the
    register they are given on here (skip_transpile semantics), so that step is the caller's.
    ``shots`` is accepted for signature compatibility: the engine is exact (shots=None)."""

    def __init__(self, zne_estimator, zne_strategy, backend=None, shots=None):
        self._zne_estimator = zne_estimator
        self._zne_strategy = zne_strategy
        self._backend = backend
        self._shots = shots

    def process(self, expectation_value, circuits, observables, parameter_values):
        job = self._zne_estimator.run(_bind(circuits, parameter_values), observables, zne_strategy=self._zne_strategy)
        return job.result().values[0]

    def process_batch(self, expectation_values, circuits, observables, parameter_values):
        """One estimator call for the whole job (the reference loops over the items, one ZNE job each)."""
        bound = [_bind(c, p) for c, p in zip(circuits, parameter_values)]
        return np.asarray(self._zne_estimator.run(bound, list(observables), zne_strategy=self._zne_strategy).result().values)


class _ModelProcessor(LearningMethodEstimatorProcessor):
    """Feature row per (circuit, Pauli term): the observable's noisy value is the single
    expectation feature and the term (coefficient 1) the measurement basis; the prediction is
    weighted with the term's coefficient and summed over the observable (estimator.py:128-148)."""

    def __init__(self, model, backend):
        self._model = model
        self._backend = backend
        self._properties = backend if isinstance(backend, dict) and "gates_set" in backend else backend_properties_v1(backend)

    def _predict(self, X):
        raise NotImplementedError

    def process(self, expectation_value, circuits, observables, parameter_values):
        total = 0.0
        for label, coeff in observable_mod.from_any(observables):
            X, _ = encode_data(circuits=[circuits], properties=self._properties, ideal_exp_vals=[[0.0]],
                               noisy_exp_vals=[[float(expectation_value)]], num_qubits=1,
                               meas_bases=encode_pauli_sum_op([(label, 1.0)]))
            total = total + float(np.asarray(self._predict(X)).reshape(-1)[0]) * np.real(coeff)
        return total

    def process_batch(self, values, circuits, observables, parameter_values):
        circs, noisy, bases, coeffs, owner = [], [], [], [], []
        for i, (v, c, o) in enumerate(zip(values, circuits, observables)):
            for label, coeff in observable_mod.from_any(o):
                circs.append(c)
                noisy.append([float(v)])
                bases.append(encode_pauli_sum_op([(label, 1.0)])[0])
                coeffs.append(np.real(coeff))
                owner.append(i)
        out = np.zeros(len(values))
        if not circs:
            return out
        widths = {len(b) for b in bases}
        if len(widths) != 1:  # observables of different widths: one model call per width
            for w in widths:
                sel = [k for k, b in enumerate(bases) if len(b) == w]
                X, _ = encode_data([circs[k] for k in sel], self._properties, [[0.0]] * len(sel), [noisy[k] for k in sel], 1,
                                   [bases[k] for k in sel])
                np.add.at(out, [owner[k] for k in sel], np.asarray(self._predict(X)).reshape(-1) * np.array([coeffs[k] for k in sel]))
            return out
        X, _ = encode_data(circs, self._properties, [[0.0]] * len(circs), noisy, 1, bases)
        np.add.at(out, owner, np.asarray(self._predict(X), dtype=float).reshape(-1) * np.array(coeffs))
        return out


class ScikitLearningModelProcessor(_ModelProcessor):
    """estimator.py:88-148 (``model.predict``)."""

    def _predict(self, X):
        return self._model.predict(X.numpy() if hasattr(X, "numpy") else X)


class TorchLearningModelProcessor(_ModelProcessor):
    """estimator.py:151-187 (``model(X)``)."""

    def _predict(self, X):
        import torch

        with torch.no_grad():
            return self._model(X).detach().cpu().numpy()


class PostProcessedJob:
    """estimator.py:194-256: forwards to the base job, post-processes in result()."""

    def __init__(self, base_job, processor, circuits, observables, parameter_values, skip_transpile=True,
                 backend=None, job_id=None, options=None):
        self._base_job = base_job
        self._processor = processor
        self._circuits = circuits
        self._observables = observables
        self._parameter_values = parameter_values
        self._skip_transpile = skip_transpile
        self._backend = backend
        self._job_id = job_id
        self._options = options

    def job_id(self):
        return self._job_id

    def backend(self):
        return self._backend

    def result(self):
        result = self._base_job.result()
        bound = [_bind(c, p) for c, p in zip(self._circuits, self._parameter_values)]
        if not self._skip_transpile:
            bound = _transpile_or_warn(bound, self._backend)
        for obs in self._observables:
            try:
                observable_mod.from_any(obs)
            except TypeError as exc:  # BlackwaterException in the reference (:226-229)
                raise ValueError("Only Pauli-sum observables are supported by learning primitive.") from exc
        mitigated = self._processor.process_batch(result.values, bound, self._observables, self._parameter_values)
        metadata = [{**meta, "original_value": value} for value, meta in zip(result.values, result.metadata)]
        return EstimatorResult(np.array(mitigated), metadata)

    def submit(self):
        return self._base_job.submit()

    def status(self):
        return self._base_job.status()

    def cancel(self):
        return self._base_job.cancel()

    def __repr__(self):
        return f"<LearningJob: {self._base_job.job_id()}>"


def _transpile_or_warn(circuits, backend):
    """skip_transpile=False (the reference's default, estimator.py:233-238: transpile at
    optimization_level=3 before encoding).  Uses qiskit's transpiler when it is importable;
    otherwise the circuits are encoded as given and the caller is told so."""
    try:
        from qiskit import transpile  # type: ignore
    except Exception:  # noqa: BLE001
        import warnings

        warnings.warn("learning(skip_transpile=False): qiskit is not importable, so no transpiler is available; "
                      "the circuits are encoded as given (pass circuits already in the backend basis)",
                      RuntimeWarning, stacklevel=3)
        return circuits
    return [transpile(c, backend, optimization_level=3) if not hasattr(c, "gate_ops") else c for c in circuits]


def patch_run(run, processor, skip_transpile=True, backend=None, options=None):
    """estimator.py:262-298: the original ``_run`` is called with keyword arguments."""

    @wraps(run)
    def patched_run(self, circuits, observables, parameter_values, **run_options):
        job = run(self, circuits=circuits, observables=observables, parameter_values=parameter_values, **run_options)
        return PostProcessedJob(job, processor, circuits, observables, parameter_values, skip_transpile=skip_transpile,
                                backend=backend, job_id=job.job_id(), options=options)

    return patched_run


def learning(cls, processor, skip_transpile=True, backend=None, options=None):
    """Decorator to turn an Estimator class into a LearningEstimator class (estimator.py:300-328)."""
    new_class = type(f"Learning{cls.__name__}", (cls,), {})
    new_class._run = patch_run(new_class._run, processor, skip_transpile, backend, options)
    return new_class
