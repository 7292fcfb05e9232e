"""Estimator primitive surface backed by the B200 engine.

Mirrors what the reference calls on its Aer-backed estimators:
  blackwater/data/utils.py:422-430, 442-443   estimator.run(circuits, observables).result().values[0]
  blackwater/library/learning/estimator.py:320-327  learning(): subclass + replace ``_run``; the
      original is invoked as run(self, circuits=..., observables=..., parameter_values=..., **opts)
      (:279-285) and the job must offer result()/job_id()/submit()/status()/cancel() (:215-256)
  blackwater/library/ngem/estimator.py:101-157  ngem(): same contract
  [3P] qiskit.primitives.BaseEstimator.run normalisation + validation

Modes: a noise model / backend given -> density-matrix (noisy) values, as
``AerEstimator(backend_options={"method": "density_matrix", "noise_model": NoiseModel.from_backend(b)},
approximation=True, skip_transpilation=True)`` with shots=None; none given -> ideal statevector
values, as ``qiskit.primitives.Estimator`` with shots=None.  When qiskit is importable the real
``EstimatorResult`` class is returned; otherwise a dataclass with the same fields.
"""
import threading
import uuid
from dataclasses import dataclass

import numpy as np

from . import circuit as circuit_mod
from . import noise as noise_mod
from . import observable as observable_mod
from .engine import Engine, STATUS_TEXT, encode_batch

try:  # pragma: no cover - qiskit is not installed in the build image
    from qiskit.primitives import EstimatorResult  # type: ignore
except Exception:  # noqa: BLE001

    @dataclass(frozen=True)
    class EstimatorResult:
        """Same fields as qiskit.primitives.EstimatorResult."""

        values: np.ndarray
        metadata: list


class B200Job:
    """Minimal JobV1-like handle (PrimitiveJob equivalent): runs eagerly on ``submit``."""

    def __init__(self, fn):
        self._fn = fn
        self._id = str(uuid.uuid4())
        self._result = None
        self._error = None
        self._done = threading.Event()

    def job_id(self):
        return self._id

    def submit(self):
        if self._done.is_set():
            return
        try:
            self._result = self._fn()
        except Exception as exc:  # noqa: BLE001 - re-raised from result()
            self._error = exc
        self._done.set()

    def result(self):
        self.submit()
        if self._error is not None:
            raise self._error
        return self._result

    def status(self):
        if not self._done.is_set():
            return "INITIALIZING"
        return "ERROR" if self._error is not None else "DONE"

    def done(self):
        return self._done.is_set() and self._error is None

    def running(self):
        return False

    def cancelled(self):
        return False

    def in_final_state(self):
        return self._done.is_set()

    def cancel(self):
        return False


_ENGINES = {}
_ENGINES_LOCK = threading.Lock()


def shared_engine(device=0):
    with _ENGINES_LOCK:
        if device not in _ENGINES:
            _ENGINES[device] = Engine(device)
        return _ENGINES[device]


class B200Estimator:
    """Exact (shots=None) Estimator on one B200.

    Args:
        backend: calibration source for the device noise model (BackendProps, properties dict,
            get_backend_properties_v1 dict or a Qiskit backend); ``None`` => ideal statevector.
        noise_model: explicit ml_qem_b200.noise.NoiseModel (overrides ``backend``).
        device: CUDA device index.
        options: default run options (kept and merged like BaseEstimator.options).
    """

    def __init__(self, backend=None, noise_model=None, device=0, options=None, engine=None, **engine_options):
        self._noise = noise_model if noise_model is not None else (
            noise_mod.from_backend(backend) if backend is not None else None)
        self._backend = backend
        self._device = device
        self._engine = engine
        self._engine_options = engine_options
        self._options = dict(options or {})

    # -- BaseEstimator surface
    @property
    def options(self):
        return dict(self._options)

    def set_options(self, **fields):
        self._options.update(fields)

    @property
    def noise_model(self):
        return self._noise

    def run(self, circuits, observables, parameter_values=None, **run_options):
        if not isinstance(circuits, (list, tuple)):
            circuits = [circuits]
        if isinstance(observables, (str,)) or not isinstance(observables, (list, tuple)) or (
                observables and isinstance(observables[0], tuple) and isinstance(observables[0][0], str)):
            observables = [observables]
        circuits = tuple(circuits)
        # the same observable / circuit objects usually repeat over the pairs (thousands of pairs for
        # small circuits): everything below works on the UNIQUE objects, the per-pair work is C-level
        # map/zip only
        ids_o = list(map(id, observables))
        conv = {i: observable_mod.from_any(o) for i, o in dict(zip(ids_o, observables)).items()}
        observables = tuple(map(conv.__getitem__, ids_o))
        if parameter_values is None:
            parameter_values = ((),) * len(circuits)
        else:
            parameter_values = list(parameter_values)
            if parameter_values and not isinstance(parameter_values[0], (list, tuple, np.ndarray)):
                parameter_values = [parameter_values]
            parameter_values = tuple(tuple(float(v) for v in pv) for pv in parameter_values)
        if len(circuits) != len(observables):
            raise ValueError(f"The number of circuits ({len(circuits)}) does not match the number of observables ({len(observables)}).")
        if len(circuits) != len(parameter_values):
            raise ValueError(f"The number of circuits ({len(circuits)}) does not match the number of parameter value sets ({len(parameter_values)}).")
        ids_c = list(map(id, circuits))
        cobj = dict(zip(ids_c, circuits))
        oobj = {id(o): o for o in conv.values()}
        shape = {}  # id(circuit) -> (number of parameters, number of qubits)
        for ic, io, npv in set(zip(ids_c, map(id, observables), map(len, parameter_values))):
            if ic not in shape:
                c = cobj[ic]
                shape[ic] = ((getattr(c, "num_parameters", 0), c.num_qubits) if not isinstance(c, str)
                             else (0, circuit_mod.from_any(c).num_qubits))
            npar, nq = shape[ic]
            o = oobj[io]
            if npv != npar or (len(o) and o.num_qubits != nq):
                # first offending pair, in pair order (the message names it)
                i = next(k for k, (a, b, pv) in enumerate(zip(ids_c, observables, parameter_values))
                         if a == ic and id(b) == io and len(pv) == npv)
                if npv != npar:
                    raise ValueError(f"The number of values ({npv}) does not match the number of parameters ({npar}) for the {i}-th circuit.")
                raise ValueError(f"The number of qubits of the {i}-th circuit ({nq}) does not match the number of qubits of the {i}-th observable ({o.num_qubits}).")
        opts = dict(self._options)
        opts.update(run_options)
        return self._run(circuits, observables, parameter_values, **opts)

    def _run(self, circuits, observables, parameter_values, **run_options):
        job = B200Job(lambda: self._call(circuits, observables, parameter_values, **run_options))
        job.submit()
        return job

    # -- the work
    def _engine_handle(self):
        """The engine of this estimator: the per-device shared one, or -- when engine options were
        given -- a private one, so the options never leak into other estimators."""
        if self._engine is None:
            self._engine = Engine(self._device, **self._engine_options) if self._engine_options else shared_engine(self._device)
        return self._engine

    def _call(self, circuits, observables, parameter_values, **run_options):
        shots = run_options.get("shots")
        if shots not in (None, 0):
            raise ValueError("B200Estimator is exact: run with shots=None")
        # unique (circuit, params) -> the pairs that use it: one evolution serves all its observables
        first = {}
        gidx = [first.setdefault(k, len(first)) for k in zip(map(id, circuits), parameter_values)]
        bound = []
        for (_, pv), i in zip(first, np.unique(np.asarray(gidx, dtype=np.int64), return_index=True)[1].tolist()):
            c = circuits[i]
            if pv:
                if isinstance(c, circuit_mod.Circuit):
                    c = c.bound_view(pv)  # the template is walked once for all parameter sets of the ansatz
                else:
                    c = c.assign_parameters(list(pv)) if hasattr(c, "assign_parameters") else c.bind_parameters(list(pv))
            bound.append(circuit_mod.from_any(c))
        g_arr = np.asarray(gidx, dtype=np.int64)
        order = np.argsort(g_arr, kind="stable")
        counts = np.bincount(g_arr, minlength=len(first))
        if len(first) and counts.min() == counts.max():
            groups = order.reshape(len(first), -1).tolist()
        else:
            groups = [g.tolist() for g in np.split(order, np.cumsum(counts)[:-1])] if len(first) else []
        # complex coefficients (Aer returns np.real_if_close of the complex sum): the C ABI takes
        # real coefficients, so such an observable is evaluated as <Re O> + i <Im O>
        is_cplx = {}
        for o in observables:
            if id(o) not in is_cplx:
                is_cplx[id(o)] = o.is_complex()
        cplx = [i for i, o in enumerate(observables) if is_cplx[id(o)]]
        if cplx:
            obs_lists = [[observables[i].real_part() if is_cplx[id(observables[i])] else observables[i] for i in g] +
                         [observables[i].imag_part() for i in g if is_cplx[id(observables[i])]] for g in groups]
        else:
            obs_lists = [[observables[i] for i in g] for g in groups]
        batch = encode_batch(bound, obs_lists)
        eng = self._engine_handle()
        noisy = self._noise is not None and not self._noise.is_ideal()
        method = "density_matrix" if noisy else "statevector"

        def evaluate(b):
            # the noise table is installed and used under one engine lock (shared engines)
            vals, status = eng.run_dm(b, noise=self._noise) if noisy else eng.run_sv(b)
            bad = np.nonzero(status)[0]
            if len(bad):
                c = int(bad[0])
                raise ValueError(f"circuit {groups[c][0]}: {STATUS_TEXT.get(int(status[c]), 'error')}")
            out = np.empty(len(circuits), dtype=complex if cplx else float)
            if not cplx:
                out[order] = vals  # group-major value order == the stable order of the pairs by group
                return out
            k = 0
            for g in groups:
                for i in g:
                    out[i] = vals[k]
                    k += 1
                for i in g:
                    if cplx and is_cplx[id(observables[i])]:
                        out[i] += 1j * vals[k]
                        k += 1
            return out

        sim_meta = {"method": method, "device": f"cuda:{self._device}"}
        meta = [{"simulator_metadata": sim_meta} for _ in circuits]
        strategy = run_options.get("zne_strategy")
        variants = run_options.get("variants")
        if variants is not None:
            # variants (ZNE folds x Pauli twirls) generated inside the library from the base gate stream
            # (docs/tutorials/zne_parallel.py:168-189, docs/tutorials/derek_files/phase_diagram.ipynb:776):
            # values[i] = twirl average at the first fold -- or its zero-noise extrapolation when an
            # extrapolator comes with the strategy; metadata[i]["variants"] keeps every (fold, twirl) value
            nf, nt = len(variants.folds), max(1, variants.twirls)
            per = []  # per unique circuit: [n_var, n_obs_of_group]
            if noisy:
                vals, status = eng.run_dm_variants(batch, variants, noise=self._noise)
                bad = np.nonzero(status)[0]
                if len(bad):
                    raise ValueError(f"circuit {groups[int(bad[0]) // (nf * nt)][0]}: {STATUS_TEXT.get(int(status[bad[0]]), 'error')}")
                k = 0
                for g in groups:
                    per.append(vals[k:k + nf * nt * len(g)].reshape(nf * nt, len(g)))
                    k += nf * nt * len(g)
            else:  # folds and twirls leave the ideal circuit unchanged
                base = evaluate(batch)
                per = [np.repeat(base[g][None, :], nf * nt, axis=0) for g in (np.asarray(g) for g in groups)]
            out = np.empty(len(circuits), dtype=float)
            gs = len(groups[0]) if groups else 0
            if groups and strategy is None and all(len(g) == gs for g in groups):
                # equal group sizes (the usual [circuit] * k pairs): no per-pair arithmetic in Python
                cube = np.stack(per).reshape(len(groups), nf, nt, gs)        # [group, fold, twirl, obs]
                out[order] = cube[:, 0].mean(axis=1).reshape(-1)
                views = list(np.ascontiguousarray(cube.transpose(0, 3, 1, 2)).reshape(-1, nf, nt))  # (group, obs) -> [fold, twirl]
                for i, vw in zip(order.tolist(), views):
                    meta[i]["variants"] = {"folds": variants.folds, "twirls": variants.twirls, "seed": variants.seed, "values": vw}
                return EstimatorResult(out, meta)
            for g, v in zip(groups, per):
                fold_means = v.reshape(nf, nt, len(g)).mean(axis=1)  # [fold, obs]
                for j, i in enumerate(g):
                    out[i] = fold_means[0, j]
                    if strategy is not None and nf > 1:
                        out[i] = strategy.extrapolator(fold_means[:, j], variants.folds)
                    meta[i]["variants"] = {"folds": variants.folds, "twirls": variants.twirls, "seed": variants.seed,
                                           "values": v[:, j].reshape(nf, nt)}
            return EstimatorResult(out, meta)
        if strategy is None:
            return EstimatorResult(np.real_if_close(evaluate(batch)), meta)
        # digital ZNE (docs/tutorials/zne_parallel.py:168-189): every noise factor is the same
        # encoded batch with its 2-qubit gates folded on the flat gate stream
        from . import zne as zne_mod

        factors = tuple(int(f) for f in strategy.noise_factors)
        vals = np.stack([evaluate(zne_mod.fold_batch(batch, f)) for f in factors], axis=-1)
        out = strategy.extrapolator(vals, factors)
        for m, v in zip(meta, vals):
            m["zne"] = {"noise_amplification": {"noise_factors": factors, "values": tuple(float(x) for x in v)},
                        "extrapolation": {"degree": getattr(strategy.extrapolator, "degree", None)}}
        return EstimatorResult(np.real_if_close(out), meta)
