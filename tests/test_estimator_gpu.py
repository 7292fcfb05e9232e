"""-m gpu: the Estimator primitive surface on the B200 engine, used the way the reference uses its
Aer-backed estimators (blackwater/data/utils.py:418-444) and wrapped the way
blackwater/library/learning/estimator.py:262-328 (learning / patch_run / PostProcessedJob) wraps
them; digital ZNE as in docs/tutorials/zne_parallel.py:168-189."""
import numpy as np
import pytest

import helpers
from ml_qem_b200 import B200Estimator, backends, families as F, noise
from ml_qem_b200.circuit import Circuit, Parameter
from ml_qem_b200.zne import PolynomialExtrapolator, ZNEStrategy, zne

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _pair(lib):
    lima = backends.fake_lima()
    return lima, B200Estimator(), B200Estimator(backend=lima)


def test_create_estimator_meas_data_drop_in(lib):
    """ideal + noisy value of one circuit/observable, exactly the calls of utils.py:422-430."""
    lima, ideal, noisy = _pair(lib)
    c = F.tfim_circuit(4, 2, 0.3, layout=[0, 1, 3, 4], num_physical=5, random_init_prefix=True)
    ob = [("ZIIIZ", 0.5), ("IXIXI", -1.0)]
    ideal_v = ideal.run([c], [ob]).result().values[0]
    noisy_v = noisy.run([c], [ob]).result().values[0]
    assert abs(ideal_v - helpers.oracle_sv_values(c, [ob])[0]) <= TOL
    assert abs(noisy_v - helpers.oracle_dm_values(c, [ob], helpers.oracle_noise("fakelima"))[0]) <= TOL
    assert abs(ideal_v - noisy_v) > 1e-4
    # create_meas_data_from_estimators (utils.py:434-444): a list of estimators, .values[0] of each
    out = [est.run(c, ob).result().values[0] for est in (ideal, noisy)]
    assert np.allclose(out, [ideal_v, noisy_v], atol=TOL)


def test_batched_run_qasm_and_parameters(lib):
    lima, ideal, noisy = _pair(lib)
    qasm = 'OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[5];\nrz(pi/2) q[0];\nsx q[0];\ncx q[0],q[1];\nx q[3];\ncx q[3],q[4];\n'
    th = Parameter("theta")
    pc = Circuit(5)
    pc.append("rx", (1,), (th,))
    pc.append("cx", (1, 3))
    circs = [qasm, pc, pc]
    obs = ["IIIZZ", "IZIZI", [("ZIIII", 2.0), ("IIIZI", 1.0)]]
    res = noisy.run(circs, obs, [(), (0.3,), (1.1,)]).result()
    assert res.values.shape == (3,) and len(res.metadata) == 3
    on = helpers.oracle_noise("fakelima")
    from ml_qem_b200.circuit import parse_qasm
    refs = [helpers.oracle_dm_values(parse_qasm(qasm), [[("IIIZZ", 1.0)]], on)[0],
            helpers.oracle_dm_values(pc.assign_parameters([0.3]), [[("IZIZI", 1.0)]], on)[0],
            helpers.oracle_dm_values(pc.assign_parameters([1.1]), [[("ZIIII", 2.0), ("IIIZI", 1.0)]], on)[0]]
    assert np.max(np.abs(res.values - refs)) <= TOL
    with pytest.raises(ValueError):
        noisy.run(circs, obs[:2])
    with pytest.raises(ValueError):
        noisy.run([pc], ["IZIZI"], [()])


def test_learning_style_decorator_wraps_the_estimator(lib):
    """The reference's learning(): subclass the estimator class, replace ``_run`` with a function
    that calls the original with KEYWORD arguments and wraps the job (estimator.py:272-296); the
    wrapped job zips result.values with result.metadata (:220-247)."""
    lima = backends.fake_lima()

    class PostProcessedJob:
        def __init__(self, base_job, processor, circuits, observables, parameter_values):
            self._base_job, self._processor = base_job, processor
            self._args = (circuits, observables, parameter_values)
            self._id = base_job.job_id()

        def result(self):
            result = self._base_job.result()
            vals, metas = [], []
            for value, circuit, obs, params, meta in zip(result.values, *self._args, result.metadata):
                vals.append(self._processor(value, circuit, obs, params))
                metas.append(dict(meta, original_value=value))
            return type(result)(np.array(vals), metas)

        def job_id(self):
            return self._id

        def status(self):
            return self._base_job.status()

    def learning(cls, processor):
        def patch_run(run):
            def patched(self, circuits, observables, parameter_values, **run_options):
                job = run(self, circuits=circuits, observables=observables, parameter_values=parameter_values, **run_options)
                return PostProcessedJob(job, processor, circuits, observables, parameter_values)
            return patched
        new_class = type("Learning" + cls.__name__, (cls,), {})
        new_class._run = patch_run(new_class._run)
        return new_class

    LearningEstimator = learning(B200Estimator, processor=lambda v, c, o, p: 2.0 * v + 1.0)
    est = LearningEstimator(backend=lima)
    c = F.tfim_circuit(4, 1, 0.5, layout=[0, 1, 3, 4], num_physical=5)
    obs = ["IIIIZ", "IIIZI", "ZIIII"]
    job = est.run([c] * 3, obs)
    assert isinstance(job.job_id(), str) and job.status() == "DONE"
    res = job.result()
    base = B200Estimator(backend=lima).run([c] * 3, obs).result().values
    assert np.allclose(res.values, 2.0 * base + 1.0, atol=TOL)
    assert [m["original_value"] for m in res.metadata] == list(base)
    assert all(m["simulator_metadata"]["method"] == "density_matrix" for m in res.metadata)


def test_learning_module_on_the_engine_estimator(lib):
    """ml_qem_b200.learning.learning() on the real engine estimator: the mitigated values are the
    model applied to the engine's noisy values, term by term (estimator.py:128-148), computed with
    one batched model call."""
    from sklearn.linear_model import Ridge

    from ml_qem_b200 import learning as L
    from ml_qem_b200.features import backend_properties_v1, encode_data

    lima = backends.fake_lima()
    props = backend_properties_v1(lima)
    rng = np.random.default_rng(2)
    circs = [F.tfim_circuit(4, s, 0.3 + 0.1 * s, layout=[0, 1, 3, 4], num_physical=5) for s in (1, 2, 3)]
    obs = [[("IIIIZ", 1.0), ("IIIZI", -0.5)], "ZIIII", [("ZIIIZ", 2.0)]]
    width = encode_data([circs[0]], props, [[0.0]], [[0.0]], 1, L.encode_pauli_sum_op("IIIIZ"))[0].shape[1]
    model = Ridge().fit(rng.normal(size=(40, width)), rng.normal(size=40))
    proc = L.ScikitLearningModelProcessor(model, lima)
    est = L.learning(B200Estimator, proc, skip_transpile=True, backend=lima)(backend=lima)
    res = est.run(circs, obs).result()
    base = B200Estimator(backend=lima).run(circs, obs).result().values
    want = [proc.process(v, c, o, ()) for v, c, o in zip(base, circs, obs)]
    assert np.allclose(res.values, want, atol=1e-12)
    assert [m["original_value"] for m in res.metadata] == list(base)


def test_zne_strategy_matches_folded_circuits_and_extrapolates(lib):
    lima = backends.fake_lima()
    ZNEEstimator = zne(B200Estimator)
    est = ZNEEstimator(backend=lima)
    c = F.tfim_circuit(4, 3, 0.4, basis="Y", layout=[0, 1, 3, 4], num_physical=5, random_init_prefix=True)
    obs = ["IIIIZ", "ZIIII"]
    strategy = ZNEStrategy(noise_factors=(1, 3, 5), extrapolator=PolynomialExtrapolator(degree=2))
    res = est.run([c, c], obs, zne_strategy=strategy).result()
    on = helpers.oracle_noise("fakelima")
    per_factor = []
    for f in (1, 3, 5):
        cf = F.tfim_circuit(4, 3, 0.4, basis="Y", layout=[0, 1, 3, 4], num_physical=5, random_init_prefix=True, fold=f)
        per_factor.append(helpers.oracle_dm_values(cf, [[(o, 1.0)] for o in obs], on))
    per_factor = np.array(per_factor).T  # [obs, factor]
    for k in range(2):
        got = res.metadata[k]["zne"]["noise_amplification"]["values"]
        assert np.max(np.abs(np.array(got) - per_factor[k])) <= TOL
        want = np.polyfit([1, 3, 5], per_factor[k], 2)[-1]
        assert abs(res.values[k] - want) <= 1e-9
    ideal = helpers.oracle_sv_values(c, [[(o, 1.0)] for o in obs])
    # the extrapolated value is closer to the ideal one than the unmitigated value
    assert np.all(np.abs(res.values - ideal) < np.abs(per_factor[:, 0] - ideal))


def test_twirled_batch_on_gpu_averages_to_the_untwirled_value_without_coherent_noise(lib):
    """Device noise here is Pauli/relaxation only, so every twirl instance has nearly the same
    noisy value; what is checked is that the flat-stream twirl generator feeds the engine."""
    from ml_qem_b200 import engine as E, zne as Z
    lima = backends.fake_lima()
    eng = E.Engine(0)
    eng.set_noise(noise.from_backend(lima))
    c = F.brickwork_circuit(4, 2, np.random.default_rng(2), layout=[0, 1, 3, 4], num_physical=5)
    obs = F.single_z_observables([0, 1, 3, 4], 5)
    base = E.encode_batch([c], [obs])
    tw = Z.twirl_batch(base, 16, np.random.default_rng(0))
    v_tw, st = eng.run_dm(tw)
    assert not st.any()
    v_sv, _ = eng.run_sv(tw)
    ideal = helpers.oracle_sv_values(c, obs)
    assert np.max(np.abs(v_sv.reshape(16, 4) - ideal)) <= TOL  # twirling never changes the ideal values
    avg = Z.average_twirls(v_tw, 16, 4)
    v0, _ = eng.run_dm(base)
    assert np.max(np.abs(avg - v0)) < 0.05


def test_ngem_consumer_passes_metadata_through(lib):
    """ngem(B200Estimator, model, backend) (blackwater/library/ngem/estimator.py:137-158): the wrapped
    ``_run`` is called with keyword arguments, the job's values are the GNN's predictions on the
    engine's noisy values, and ``result.metadata`` of the base estimator is passed through (:86)."""
    import torch

    from ml_qem_b200 import features as FT, gnn

    lima, ideal, noisy = _pair(lib)
    props = FT.backend_properties_v1(lima)
    circs = [F.tfim_circuit(4, s, 0.2 * s, layout=[0, 1, 3, 4], num_physical=5) for s in (1, 2, 3)]
    obs = ["IIIIZ", "ZIIII", [("IIIZI", 1.0)]]
    nf = len(FT.circuit_to_graph_data_json(circs[0], props, use_qubit_features=True, use_gate_features=True)["nodes"]["DAGOpNode"][0])
    torch.manual_seed(0)
    model = gnn.ExpValCircuitGraphModel(num_node_features=nf, hidden_channels=6, exp_value_size=1, dropout=0.0)
    NgemEstimator = gnn.ngem(B200Estimator, model, lima)
    est = NgemEstimator(backend=lima)
    job = est.run(circs, obs)
    res = job.result()
    base = noisy.run(circs, obs).result()
    assert res.values.shape == (3,) and len(res.metadata) == 3
    assert res.metadata == base.metadata and "NgemJob" in repr(job)
    # the model saw the engine's noisy values: recompute the prediction of circuit 1 by hand
    g = FT.circuit_to_graph_data_json(circs[1], props, use_qubit_features=True, use_gate_features=True)
    e = FT.ExpValueEntry(circuit_graph=g, observable=[], ideal_exp_value=0.0, noisy_exp_values=[float(base.values[1])])
    b = gnn.graph_batch([e])
    model.eval()
    with torch.no_grad():
        want = float(model(b["noisy_0"], b["observable"], b["circuit_depth"], b["x"], b["edge_index"], b["batch"], 1)[0, 0])
    assert abs(res.values[1] - want) < 1e-5


def test_zne_processor_in_the_learning_decorator(lib):
    """learning(B200Estimator, ZNEProcessor(zne estimator, strategy)) (blackwater/library/learning/
    estimator.py:33-86, 300-328): values = zero-noise extrapolation of the folded circuits, metadata
    keeps the unmitigated ``original_value``."""
    from ml_qem_b200 import learning, zne

    lima, ideal, noisy = _pair(lib)
    strategy = zne.ZNEStrategy(noise_factors=(1, 3), extrapolator=zne.PolynomialExtrapolator(degree=1))
    proc = learning.ZNEProcessor(zne.zne(B200Estimator)(backend=lima), strategy, backend=lima)
    est = learning.learning(B200Estimator, proc, backend=lima)(backend=lima)
    circs = [F.tfim_circuit(4, 2, 0.3, layout=[0, 1, 3, 4], num_physical=5), F.tfim_circuit(4, 3, 0.7, layout=[0, 1, 3, 4], num_physical=5, basis="X")]
    obs = ["IIIIZ", [("ZIIIZ", 0.5), ("IIIZI", 1.0)]]
    res = est.run(circs, obs).result()
    on = helpers.oracle_noise("fakelima")
    for k, (c, o) in enumerate(zip(circs, obs)):
        ob = [(o, 1.0)] if isinstance(o, str) else o
        v1 = helpers.oracle_dm_values(c, [ob], on)[0]
        c3 = F.tfim_circuit(4, 2 + k, 0.3 if k == 0 else 0.7, layout=[0, 1, 3, 4], num_physical=5, basis="Z" if k == 0 else "X", fold=3)
        v3 = helpers.oracle_dm_values(c3, [ob], on)[0]
        assert abs(res.values[k] - (1.5 * v1 - 0.5 * v3)) <= 1e-9
        assert abs(res.metadata[k]["original_value"] - v1) <= TOL
    assert abs(proc.process(0.0, circs[0], obs[0], ()) - res.values[0]) <= 1e-12


def test_vqe_style_parameter_sweep_of_one_ansatz(lib):
    """run(batch * [ansatz], batch * [op], parameter_values) as the (patched) VQE of the reference
    calls it (docs/tutorials/vqe_to_substitute*.py:260-269): 64 parameter sets of one parametrised
    ansatz with a multi-Pauli Hamiltonian, noisy and ideal, against the oracle (sample) and against
    circuits bound one by one."""
    lima, ideal, noisy = _pair(lib)
    th = [Parameter(f"t[{i}]") for i in range(8)]
    a = Circuit(5)
    for i, q in enumerate((0, 1, 3, 4)):
        a.ry(th[i], q)
    a.cx(0, 1); a.cx(1, 3); a.cx(3, 4)
    for i, q in enumerate((0, 1, 3, 4)):
        a.rz(2 * th[4 + i] - 0.25, q); a.sx(q)
    ham = [("ZZIII", -1.05), ("IZIZI", 0.39), ("XIIXI", 0.18), ("YYIII", -0.01), ("IIIII", 0.7)]
    rng = np.random.default_rng(8)
    vals = rng.uniform(-np.pi, np.pi, size=(64, 8))
    rn = noisy.run([a] * 64, [ham] * 64, vals).result()
    ri = ideal.run([a] * 64, [ham] * 64, [tuple(v) for v in vals]).result()
    bound = [a.bind_parameters(list(v)) for v in vals]
    rb = noisy.run(bound, [ham] * 64).result()
    assert np.array_equal(rn.values, rb.values)
    on = helpers.oracle_noise("fakelima")
    for k in (0, 31, 63):
        assert abs(rn.values[k] - helpers.oracle_dm_values(bound[k], [ham], on)[0]) <= TOL
        assert abs(ri.values[k] - helpers.oracle_sv_values(bound[k], [ham])[0]) <= TOL
