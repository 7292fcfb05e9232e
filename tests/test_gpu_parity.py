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
    for kq in (6, 7, 5, 4):
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


# ---------------------------------------------------------------------------------------------
# wide statevectors (tile-sweep path, > 12 active qubits)
# ---------------------------------------------------------------------------------------------
def _chain(n):
    return [(i, i + 1) for i in range(n - 1)] + [(i + 1, i) for i in range(n - 1)]


@pytest.mark.parametrize("tile_bits", [0, 12, 7])
def test_sv_wide_vs_oracle(engine_gpu, tile_bits):
    rng = np.random.default_rng(40 + tile_bits)
    circs, obs = [], []
    for n in (13, 14, 16, 13):
        circs.append(F.random_basis_circuit(n, 150, rng, _chain(n)))
        obs.append([[(l, float(rng.normal()))] for l in _labels(rng, n, 6)] + [[("Z" * n, 1.0), ("I" * n, 0.5)]])
    circs.append(F.tfim_circuit(15, 3, 0.4, basis="Y"))
    obs.append(F.tfim_observables(list(range(15)), 15))
    circs.append(F.brickwork_circuit(14, 2, np.random.default_rng(3)))
    obs.append(F.single_z_observables(list(range(14)), 14))
    # a narrow circuit in the same batch goes through the one-CTA kernel
    circs.append(F.tfim_circuit(6, 2, 0.3))
    obs.append(F.tfim_observables(list(range(6)), 6))
    ref = np.concatenate([helpers.oracle_sv_values(c, o) for c, o in zip(circs, obs)])
    engine_gpu.set_options(sv_tile_bits=tile_bits)
    vals, status = engine_gpu.run_sv(engine.encode_batch(circs, obs))
    engine_gpu.set_options()
    assert not status.any()
    assert np.max(np.abs(vals - ref)) <= TOL


def test_sv_wide_general_gates(engine_gpu):
    from ml_qem_b200.circuit import Circuit
    rng = np.random.default_rng(6)
    n = 13
    c = Circuit(n)
    for _ in range(120):
        a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
        k = int(rng.integers(0, 8))
        th = float(rng.uniform(-3, 3))
        if k == 0: c.append("cz", (a, b))
        elif k == 1: c.append("swap", (a, b))
        elif k == 2: c.append("crx", (a, b), (th,))
        elif k == 3: c.append("rzz", (a, b), (th,))
        elif k == 4: c.append("rxx", (a, b), (th,))
        elif k == 5: c.append("ecr", (a, b))
        elif k == 6: c.append("u3", (a,), (th, 0.3, -1.1))
        else: c.append("cx", (a, b))
    obs = [[(l, 1.0)] for l in _labels(rng, n, 8)]
    ref = helpers.oracle_sv_values(c, obs)
    vals, status = engine_gpu.run_sv(engine.encode_batch([c], [obs]))
    assert not status.any()
    assert np.max(np.abs(vals - ref)) <= TOL


def test_sharded_statevector_single_rank_gpu(engine_gpu):
    """ShardedStatevector on one GPU (no exchange): torch-owned state, segments through the C ABI."""
    from ml_qem_b200.statevector import GpuExecutor, ShardedStatevector
    sv = ShardedStatevector(GpuExecutor(engine_gpu))
    n = 18
    c = F.tfim_circuit(n, 3, 0.6, basis="X")
    obs = F.tfim_observables(list(range(n)), n)
    vals = sv.estimate(c, obs)
    ref = helpers.oracle_sv_values(c, obs)
    assert np.max(np.abs(vals - ref)) <= TOL
    e = F.tfim_circuit(3, 0, 0.1)  # no gates at all
    vals = sv.estimate(e, [[("ZZZ", 1.0)], [("XII", 1.0)]])
    assert np.allclose(vals, [1.0, 0.0], atol=TOL)


def test_sv_26q_product_state_and_light_cone(engine_gpu):
    """Size-independent properties at a width the oracle cannot reach (26 qubits = 1 GiB):
    J = 0 gives a product state with <Z> = cos(2 h dt steps); with J != 0 the bulk of a long
    chain equals the bulk of a short chain (light cone of 2 steps)."""
    n, steps, h, dt = 26, 3, 1.0, 0.5
    c = F.tfim_circuit(n, steps, 0.0, h=h, dt=dt)
    obs = F.tfim_observables(list(range(n)), n)
    vals, status = engine_gpu.run_sv(engine.encode_batch([c], [obs]))
    assert not status.any()
    z = np.cos(2 * h * dt * steps)
    assert np.max(np.abs(vals[:n] - z)) <= TOL
    assert np.max(np.abs(vals[n:2 * n - 1] - z * z)) <= TOL
    assert np.max(np.abs(vals[2 * n - 1:3 * n - 2])) <= TOL
    assert abs(vals[-1] - z ** n) <= TOL
    c = F.tfim_circuit(n, 2, 0.8)
    small = F.tfim_circuit(12, 2, 0.8)
    pick = lambda w, q: [[(F.pad_label({q: "Z"}, w), 1.0)], [(F.pad_label({q: "Z", q + 1: "Z"}, w), 1.0)],
                         [(F.pad_label({q: "X", q + 1: "X"}, w), 1.0)]]
    v_big, _ = engine_gpu.run_sv(engine.encode_batch([c], [pick(n, 13)]))
    ref = helpers.oracle_sv_values(small, pick(12, 5))
    assert np.max(np.abs(v_big - ref)) <= TOL


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_sharded_statevector_two_gpus(lib):
    """Amplitude-sharded over 2 GPUs (torchrun, NCCL all_to_all exchange) against the oracle."""
    import subprocess, sys, os
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for world in [w for w in (2, 4, 8) if w <= _n_gpus()]:
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                            "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
                            os.path.join(root, "tests", "svx_multi_gpu.py"), "--check"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("g", [1, 2, 3])
def test_svx_simulated_ranks_on_one_gpu(engine_gpu, g):
    """The amplitude-sharded program with all 2^g shards on ONE GPU: every local segment goes
    through the C ABI with its rank, the EXCHANGE is done with tensor copies (what
    all_to_all_single does between GPUs)."""
    import torch
    n = 15
    rng = np.random.default_rng(70 + g)
    cases = [(F.tfim_circuit(n, 3, 0.45, basis="Y"), F.tfim_observables(list(range(n)), n)),
             (F.random_basis_circuit(n, 200, rng, _chain(n)),
              [[("XYZ" * 5, 0.7), ("Z" * n, 1.0)], [("IIIIXIIIIIIIIII", 1.0)], [("Y" * n, 1.0)]])]
    G = 1 << g
    dev = torch.device("cuda", 0)
    for circ, obs in cases:
        prog = engine.SvxProgram(engine.encode_batch([circ], [obs]), 0, 0, g)
        info = prog.info
        assert info["status"] == 0 and info["n_exchanges"] >= 1
        prog.upload(engine_gpu)
        nl = info["n_local"]
        shards = [torch.empty(1 << nl, dtype=torch.complex128, device=dev) for _ in range(G)]
        vals = torch.zeros(info["n_observables"], dtype=torch.float64, device=dev)
        blk = 1 << (nl - g)
        ref = helpers.oracle_sv_values(circ, obs)
        # fused = True: the EXCHANGE rides on the store of the preceding sweep (bwq_svx_run_segment_push:
        # P2P stores into the peers' new shards -- here same-device "peers"); the new shards start as
        # NaN so a tile the pushed sweep failed to deliver (e.g. a known-zero one) cannot go unnoticed
        for fused in (False, True):
            vals.zero_()
            segs = info["segs"]
            skip = False
            n_fused = 0
            for seg, (kind, first, count, _) in enumerate(segs):
                if skip:
                    skip = False
                    continue
                if fused and kind == engine.SEG_SWEEPS and seg + 1 < len(segs) and segs[seg + 1][0] == engine.SEG_EXCHANGE:
                    new = [torch.full_like(s, float("nan")) for s in shards]
                    for r in range(G):
                        prog.run_segment_push(engine_gpu, seg, shards[r].data_ptr(), r, [t.data_ptr() for t in new])
                    engine_gpu.sync()
                    shards = new
                    skip = True
                    n_fused += 1
                elif kind == engine.SEG_EXCHANGE:
                    torch.cuda.synchronize()
                    new = [torch.empty_like(s) for s in shards]
                    for s in range(G):
                        for v in range(G):
                            new[v][s * blk:(s + 1) * blk] = shards[s][v * blk:(v + 1) * blk]
                    shards = new
                    torch.cuda.synchronize()
                else:
                    for r in range(G):
                        prog.run_segment(engine_gpu, seg, shards[r].data_ptr(), r, vals.data_ptr())
                    engine_gpu.sync()
            assert np.max(np.abs(vals.cpu().numpy() - ref)) <= TOL, fused
            assert not fused or n_fused >= 1
        prog.close()


def test_pipelined_run_and_meas_data_call(engine_gpu):
    """bwq_dm_run cuts batches of >= 256 circuits into pipelined segments (mixed widths, an empty
    circuit and a failing circuit inside): the values must be bit-identical to the unsegmented run
    and to the prepared/resident path, and bwq_meas_data_run == (bwq_sv_run, bwq_dm_run).
    A sample is checked against the oracle."""
    lima = backends.fake_lima()
    nm = noise.from_backend(lima)
    on = helpers.oracle_noise("fakelima")
    rng = np.random.default_rng(23)
    circs, obs = [], []
    for i in range(600):
        if i == 301:
            c = F.random_basis_circuit(5, 0, rng, lima.coupling_map)  # no gate at all: host-side value
        else:
            c = F.random_basis_circuit(5, int(rng.integers(1, 40)), rng, lima.coupling_map)
        circs.append(c)
        obs.append([[(l, float(rng.normal()))] for l in _labels(rng, 5, int(rng.integers(1, 4)))])
    batch = engine.encode_batch(circs, obs)
    engine_gpu.set_noise(nm)
    # 128 = BWQ_OPT_NO_ONCHIP: these 5-qubit circuits would otherwise run on dm_onchip_kernel (no segments)
    engine_gpu.set_options(flags=8 | 128)  # BWQ_OPT_FORCE_PIPELINE (5-qubit circuits are below the automatic threshold)
    v_pipe, st_pipe = engine_gpu.run_dm(batch)
    assert engine_gpu.stats()["n_sweep_launches"] >= 4  # one launch per segment at least
    engine_gpu.set_options(flags=4 | 128)  # BWQ_OPT_NO_PIPELINE
    v_one, st_one = engine_gpu.run_dm(batch)
    engine_gpu.set_options()
    assert not st_pipe.any() and not st_one.any()
    assert np.array_equal(v_pipe, v_one)
    assert not engine_gpu.prepare_dm(batch).any()
    assert np.array_equal(engine_gpu.execute_dm(), v_pipe)
    engine_gpu.set_options(flags=128)
    v_sv, st_sv = engine_gpu.run_sv(batch)
    engine_gpu.set_options(flags=8 | 128)
    ideal, noisy, st_i, st_n = engine_gpu.run_meas_data(batch)
    engine_gpu.set_options()
    assert not st_i.any() and not st_n.any()
    assert np.array_equal(ideal, v_sv) and np.array_equal(noisy, v_pipe)
    offs = np.cumsum([0] + [len(o) for o in obs])
    for i in (0, 150, 299, 300, 301, 302, 450, 599):
        ref = helpers.oracle_dm_values(circs[i], obs[i], on)
        assert np.max(np.abs(v_pipe[offs[i]:offs[i + 1]] - ref)) <= TOL, i


def test_dm_14q_full_size_factorised(engine_gpu):
    """BASELINE cfg3 size (14 qubits, 2.15 GB Pauli-basis state; the numpy oracle stops near 9).
    Size-independent property: a circuit acting only inside the pairs (0,1), (2,3), ... leaves a
    product state, so Tr(rho P) is the product of the pair expectations -- each one computed by the
    oracle on the 2-qubit sub-circuit with the device noise of those physical qubits."""
    from ml_qem_b200.circuit import Circuit

    n = 14
    be = backends.synthetic_chain(n, seed=14)
    nm = noise.from_backend(be)
    from oracle import noise_model as onm
    on = onm.from_backend(be.to_dict())
    rng = np.random.default_rng(14)
    full = Circuit(n)
    blocks = []
    for k in range(n // 2):
        a, b = 2 * k, 2 * k + 1
        sub = Circuit(n)
        for _ in range(int(rng.integers(6, 14))):
            r = int(rng.integers(0, 4))
            if r == 0:
                op = ("rz", (int(rng.choice([a, b])),), (float(rng.uniform(-3, 3)),))
            elif r == 1:
                op = ("sx", (int(rng.choice([a, b])),), ())
            else:
                op = ("cx", (a, b) if rng.integers(0, 2) else (b, a), ())
            sub.ops.append(op)
        blocks.append(sub)
    # interleave the blocks' gates (program order inside a block is kept)
    cursors = [0] * len(blocks)
    while any(c < len(bk.ops) for c, bk in zip(cursors, blocks)):
        k = int(rng.integers(0, len(blocks)))
        if cursors[k] < len(blocks[k].ops):
            full.ops.append(blocks[k].ops[cursors[k]])
            cursors[k] += 1
    labels = _labels(rng, n, 12) + ["Z" * n, "I" * n]
    obs = [[(l, 1.0)] for l in labels]
    engine_gpu.set_noise(nm)
    engine_gpu.set_options()
    vals, status = engine_gpu.run_dm(engine.encode_batch([full], [obs]))
    assert not status.any()
    ref = np.ones(len(labels))
    for k, sub in enumerate(blocks):
        a, b = 2 * k, 2 * k + 1
        sub_obs = []
        for l in labels:
            chars = ["I"] * n
            chars[n - 1 - a], chars[n - 1 - b] = l[n - 1 - a], l[n - 1 - b]
            sub_obs.append([("".join(chars), 1.0)])
        c2, o2, n2 = helpers.compact(sub, sub_obs, on)
        ref *= helpers.oracle_dm_values(c2, o2, n2)
    assert np.max(np.abs(vals - ref)) <= TOL


def test_mixed_widths_failing_circuit_and_empty_batch(engine_gpu):
    """One batch with widths 3..9 (chunks of different tile sizes, forced pipelining), a circuit the
    density-matrix path must reject (17 active qubits > 16) without voiding the batch, and the
    empty batch.  Values against the oracle on the active qubits."""
    from oracle import noise_model as onm

    width = 17
    be = backends.synthetic_chain(width, seed=17)
    nm = noise.from_backend(be)
    on = onm.from_backend(be.to_dict())
    rng = np.random.default_rng(31)
    cm = [(i, i + 1) for i in range(width - 1)] + [(i + 1, i) for i in range(width - 1)]
    circs, obs = [], []
    for i in range(280):
        n = 3 + i % 7
        circs.append(F.random_basis_circuit(n, int(rng.integers(2 * n, 8 * n)), rng, [p for p in cm if max(p) < n], width))
        obs.append([[("I" * (width - n) + "".join(rng.choice(list("IXYZ"), size=n)), float(rng.normal()))] for _ in range(2)])
    wide = F.tfim_circuit(width, 1, 0.3)
    bad = 137
    circs[bad], obs[bad] = wide, [[("Z" * width, 1.0)], [("I" * (width - 1) + "Z", 1.0)]]
    batch = engine.encode_batch(circs, obs)
    engine_gpu.set_noise(nm)
    engine_gpu.set_options(flags=8)
    ideal, noisy, st_i, st_n = engine_gpu.run_meas_data(batch)
    engine_gpu.set_options()
    assert not st_i.any()
    assert st_n[bad] != 0 and not np.delete(st_n, bad).any()
    offs = np.cumsum([0] + [len(o) for o in obs])
    assert np.isnan(noisy[offs[bad]:offs[bad + 1]]).all() and not np.isnan(np.delete(noisy, range(offs[bad], offs[bad + 1]))).any()
    for i in (0, 6, 69, 136, 138, 139, 279):
        c2, o2, n2 = helpers.compact(circs[i], obs[i], on)
        assert np.max(np.abs(noisy[offs[i]:offs[i + 1]] - helpers.oracle_dm_values(c2, o2, n2))) <= TOL, i
        assert np.max(np.abs(ideal[offs[i]:offs[i + 1]] - helpers.oracle_sv_values(c2, o2))) <= TOL, i
    empty = engine.encode_batch([], [])
    v, s = engine_gpu.run_dm(empty)
    assert len(v) == 0 and len(s) == 0
    i2, n2_, s1, s2 = engine_gpu.run_meas_data(empty)
    assert len(i2) == 0 and len(n2_) == 0
