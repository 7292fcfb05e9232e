"""CPU: lowering stage (host C++ in libbwq.so) + Pauli-transfer algebra against the oracle, by
executing the lowered sweep program with the numpy emulator.  No GPU needed."""
import numpy as np
import pytest

import helpers
from program_emulator import expvals, run_program
from ml_qem_b200 import backends, engine, families as F, noise

TOL = 1e-12


def _labels(rng, n, k):
    return ["".join(rng.choice(list("IXYZ"), size=n)) for _ in range(k)]


def _check(circ, obs, nm, on, tilings=((6, 2), (3, 1), (4, 2), (2, -1), (7, 2))):
    fb = engine.encode_batch([circ], [obs])
    if on is None:
        cc, oo = helpers.compact(circ, obs)
        ref = helpers.oracle_dm_values(cc, oo, None)
    else:
        cc, oo, on2 = helpers.compact(circ, obs, on)
        ref = helpers.oracle_dm_values(cc, oo, on2)
    for kq, low in tilings:
        prog = engine.lower_dm(fb, 0, nm, kq, low)
        assert prog["status"] == 0
        got = expvals(prog, run_program(prog), [len(o) for o in obs])
        assert np.max(np.abs(got - ref)) <= TOL, (kq, low)
    return prog


def test_lima_random_circuits(lib):
    lima = backends.fake_lima()
    nm = noise.from_backend(lima)
    on = helpers.oracle_noise("fakelima")
    rng = np.random.default_rng(1)
    for _ in range(12):
        c = F.random_basis_circuit(5, int(rng.integers(0, 60)), rng, lima.coupling_map)
        obs = [[(l, float(rng.normal()))] for l in _labels(rng, 5, 6)] + [[("ZZIII", 0.5), ("IXXII", -1.5), ("IIIII", 2.0)]]
        _check(c, obs, nm, on)


def test_tfim_zne_fold_on_lima_chain(lib):
    lima = backends.fake_lima()
    nm = noise.from_backend(lima)
    on = helpers.oracle_noise("fakelima")
    for fold in (1, 3, 5):
        c = F.tfim_circuit(4, 3, 0.37, basis="Y", layout=[0, 1, 3, 4], num_physical=5, fold=fold, random_init_prefix=True)
        prog = _check(c, F.single_z_observables([0, 1, 3, 4], 5), nm, on)
        assert prog["n_digits"] == 4 and len(prog["sweeps"]) == 1  # idle qubit 2 truncated, on-chip


def test_coherent_cx_noise_uses_dense_op(lib):
    lima = backends.fake_lima()
    nm, thetas = noise.add_coherent_noise(lima, theta=0.04 * np.pi, seed=0)
    from oracle import noise_model as onm
    on = onm.add_coherent_noise(helpers.golden("backends.json")["fakelima"], theta=0.04 * np.pi, seed=0)
    assert np.allclose(thetas, helpers.golden("kats.json")["coherent_thetas_8digits"], atol=1e-8)
    rng = np.random.default_rng(3)
    c = F.random_basis_circuit(5, 40, rng, lima.coupling_map)
    prog = _check(c, [[(l, 1.0)] for l in _labels(rng, 5, 5)], nm, on, tilings=((6, 2), (3, 1)))
    assert prog["needs_dense"]


def test_non_basis_gates_and_reset(lib):
    from ml_qem_b200 import Circuit
    c = Circuit(4)
    c.h(0); c.cz(0, 1); c.u3(0.3, 0.2, -0.7, 2); c.swap(1, 2); c.rzz(0.4, 2, 3); c.ecr(3, 0); c.t(1); c.sdg(2)
    c.crx(0.9, 1, 3); c.reset(0); c.ry(0.5, 0); c.cp(1.1, 0, 2); c.y(3); c.rzx(0.6, 1, 2); c.iswap(0, 3); c.cy(2, 0)
    rng = np.random.default_rng(5)
    obs = [[(l, 1.0)] for l in _labels(rng, 4, 8)]
    fb = engine.encode_batch([c], [obs])
    ref = helpers.oracle_dm_values(c, obs, None)
    for kq, low in ((6, 2), (2, -1), (3, 2)):
        prog = engine.lower_dm(fb, 0, None, kq, low)
        got = expvals(prog, run_program(prog), [1] * len(obs))
        assert np.max(np.abs(got - ref)) <= TOL


def test_idle_qubits_and_empty_circuit(lib):
    from ml_qem_b200 import Circuit
    c = Circuit(6)
    c.sx(4)
    obs = [[("IZIIII", 1.0)], [("IYIIII", 1.0)], [("ZIIIII", 1.0)], [("XIIIII", 1.0)], [("IIIIII", 3.0)]]
    fb = engine.encode_batch([c], [obs])
    prog = engine.lower_dm(fb, 0, None, 6, 2)
    got = expvals(prog, run_program(prog), [1] * 5)
    assert np.allclose(got, [0.0, -1.0, 1.0, 0.0, 3.0], atol=1e-15)
    empty = engine.lower_dm(engine.encode_batch([Circuit(3)], [[[("ZZZ", 1.0)]]]), 0, None, 6, 2)
    assert len(empty["sweeps"]) == 0 and empty["status"] == 0


def test_bad_input_sets_status(lib):
    import numpy as np
    from ml_qem_b200.engine import FlatBatch, OP_DTYPE
    ops = np.zeros(1, dtype=OP_DTYPE)
    ops["opcode"], ops["q0"] = 20, 0  # unknown opcode
    fb = FlatBatch([2], [0, 1], ops, [], [0, 0], [0], [], [], [])
    assert engine.lower_dm(fb, 0, None, 6, 2)["status"] == 1
    ops["opcode"], ops["q0"] = 1, 7  # qubit out of range
    fb = FlatBatch([2], [0, 1], ops, [], [0, 0], [0], [], [], [])
    assert engine.lower_dm(fb, 0, None, 6, 2)["status"] == 3


def test_sweep_packing_respects_tile_and_order(lib):
    be = backends.synthetic_chain(10, seed=1)
    nm = noise.from_backend(be)
    c = F.tfim_circuit(10, 3, 0.5)
    fb = engine.encode_batch([c], [F.single_z_observables(range(10), 10)])
    for kq, low in ((6, 2), (6, 1), (7, 2), (4, 2)):
        prog = engine.lower_dm(fb, 0, nm, kq, low)
        assert prog["n_digits"] == 10
        from program_emulator import decode_block
        total = 0
        for sw in prog["sweeps"]:
            pos = list(sw[1:1 + kq])
            m = min(max(low, 1), kq - 2)
            assert pos == sorted(set(pos)) and pos[:m] == list(range(m))
            assert sw[9] * 16 <= 8192
            passes, _ = decode_block(prog, sw)
            for sa, sb, ops in passes:
                assert sa < kq and sb < kq and sa != sb and len(ops) > 0
            total += len(passes)
        assert total == prog["n_passes"]


def test_direct_pass_flags_are_consistent(lib):
    """Direct passes (kernels.cuh): the planner may move a commuting pass to the front / back of a
    sweep and flag it to exchange its register groups with global memory.  The flagged passes must
    avoid the two lowest tile slots, the first-pass descriptor in SweepDesc::pos[7] must repeat the
    first pass header, and the reordered program must still match the oracle (run by the emulator
    in emitted order)."""
    from program_emulator import decode_block

    n = 9
    be = backends.synthetic_chain(n, seed=9)
    nm = noise.from_backend(be)
    from oracle import noise_model as onm
    on = onm.from_backend(be.to_dict())
    rng = np.random.default_rng(9)
    circ = F.brickwork_circuit(n, 3, rng, list(range(n)), n)
    obs = [[(l, 1.0)] for l in _labels(rng, n, 5)]
    prog = _check(circ, obs, nm, on, tilings=((6, 2),))
    n_first = n_last = 0
    for sw in prog["sweeps"]:
        passes, _ = decode_block(prog, sw)
        blk = prog["prog"][2 * int(sw[0]): 2 * (int(sw[0]) + int(sw[9]))].view(np.uint8)
        flags = [int(blk[16 * (1 + p) + 6]) for p in range(len(passes))]
        desc = int(sw[8])  # pos[7]
        assert all(f == 0 for f in flags[1:-1])
        if flags[0] & 1:
            sa, sb, _ = passes[0]
            assert min(sa, sb) >= 2 and desc == (0x80 | sa | (sb << 3))
            n_first += 1
        else:
            assert desc & 0x80 == 0
        if flags[-1] & 2:
            sa, sb, _ = passes[-1]
            assert min(sa, sb) >= 2
            n_last += 1
        if len(passes) > 1:
            assert not (flags[0] & 2) and not (flags[-1] & 1)
    assert n_first > 0 and n_last > 0


def test_tma_tile_layout_programs(lib):
    """Circuits wider than the 6-digit tile in the TMA tile layout (what the engine runs by default on
    the GPU): same values as the oracle through the emulator; every pass's host-computed addressing
    covers the tile exactly once (checked inside the emulator's decoder); the common op lists get a
    straight-line signature; the box order keeps the chain circuits free of bank conflicts."""
    import program_emulator as pe
    from oracle import noise_model as onm

    rng = np.random.default_rng(12)
    n = 8
    be = backends.synthetic_chain(n, seed=5)
    nm, on = noise.from_backend(be), onm.from_backend(be.to_dict())
    cases = [F.tfim_circuit(n, 3, 0.4, basis="Y"), F.brickwork_circuit(n, 2, rng, twirl_rng=np.random.default_rng(2)),
             F.random_basis_circuit(n, 150, rng, [(i, i + 1) for i in range(n - 1)] + [(i + 1, i) for i in range(n - 1)])]
    sig_hist = {}
    n_direct = [0]
    for c in cases:
        obs = [[(l, 1.0)] for l in _labels(rng, n, 6)]
        fb = engine.encode_batch([c], [obs])
        ref = helpers.oracle_dm_values(c, obs, on)
        prog = engine.lower_dm(fb, 0, nm, tma=True, tma_direct_store=True)
        assert prog["status"] == 0 and all(sw[8] == 0x40 for sw in prog["sweeps"])
        got = expvals(prog, run_program(prog), [1] * len(obs))
        assert np.max(np.abs(got - ref)) <= TOL
        # same state as the classic layout (the direct passes of the classic planner reorder
        # commuting passes, so the rounding differs in the last bits)
        classic = engine.lower_dm(fb, 0, nm)
        assert np.max(np.abs(run_program(classic) - run_program(prog))) <= 1e-14
        for sw in prog["sweeps"]:
            blk = prog["prog"][2 * int(sw[0]): 2 * (int(sw[0]) + int(sw[9]))]
            b = blk.view(np.uint8)
            for p in range(int(blk[:1].view(np.int32)[0])):
                h = b[16 * (1 + p): 16 * (2 + p)]
                sig_hist[int(h[7])] = sig_hist.get(int(h[7]), 0) + 1
                ext = int(blk[:1].view(np.int32)[1])
                worst = pe.check_tma_pass(h, b[16 * ext + 64 * p: 16 * ext + 64 * (p + 1)].view(np.uint32), int(h[4]), int(h[5]))
                if c is not cases[2]:
                    assert worst == 1, (int(h[4]), int(h[5]))  # chain circuits: conflict free by box order
                # slot positions in the block header = ascending positions permuted by the box order
                pos = [int(x) for x in sw[1:9]]
                slot_pos = pos[:2] + [pos[2 + ((pos[6] >> (2 * k)) & 3)] for k in range(4)]
                assert [int(x) for x in b[8:14]] == slot_pos
                if int(h[6]) & 2:  # last pass stores its groups directly: corner offsets in the state
                    n_p = int(blk[:1].view(np.int32)[0])
                    assert p == n_p - 1 and int(h[4]) >= 2 and int(h[5]) >= 2
                    gcor = b[16 * ext + 64 * n_p: 16 * ext + 64 * (n_p + 1)].view(np.uint32)
                    want = [((i & 3) << (2 * slot_pos[int(h[4])])) | ((i >> 2) << (2 * slot_pos[int(h[5])])) for i in range(16)]
                    assert [int(x) for x in gcor] == want
                    n_direct[0] += 1
    assert sum(v for k, v in sig_hist.items() if k != 0) > sum(sig_hist.values()) // 2
    assert n_direct[0] > 0
    # narrow circuits (one tile) keep the classic layout
    c6 = F.tfim_circuit(6, 2, 0.3)
    p6 = engine.lower_dm(engine.encode_batch([c6], [F.single_z_observables(list(range(6)), 6)]), 0, nm, tma=True)
    assert all(sw[8] != 0x40 for sw in p6["sweeps"])


def test_tma_pass_layout_is_conflict_free_except_three_pairs(lib):
    import program_emulator as pe
    for hi in range(1, 6):
        for lo in range(hi):
            tbit, beta = pe.tma_pass_layout(lo, hi)
            assert sorted(tbit + [beta] + [2 * lo, 2 * lo + 1, 2 * hi, 2 * hi + 1]) == list(range(12))
            if lo == 0:
                continue
            tid = np.arange(128)
            j = np.zeros(128, dtype=np.int64)
            for k in range(7):
                j |= ((tid >> k) & 1) << tbit[k]
            a = 8 * pe.tswz(j)
            degree = max(8 // len(set(((a[q:q + 8] >> 4) & 7).tolist())) for q in range(0, 128, 8))
            assert degree == (2 if (lo, hi) in ((1, 2), (1, 3)) else 1), (lo, hi, degree)
