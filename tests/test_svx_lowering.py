"""CPU: statevector sweep planner (host C++ in libbwq.so) against the oracle, by executing the
lowered program with the numpy emulator -- single shard and amplitude-sharded over 2, 4, 8
simulated ranks (EXCHANGE segments included).  No GPU needed."""
import numpy as np
import pytest

import helpers
from svx_emulator import run
from ml_qem_b200 import engine, families as F
from ml_qem_b200.circuit import Circuit

TOL = 1e-12


def _labels(rng, n, k):
    return ["".join(rng.choice(list("IXYZ"), size=n)) for _ in range(k)]


def _check(circ, obs, configs):
    fb = engine.encode_batch([circ], [obs])
    cc, oo = helpers.compact(circ, obs)
    ref = helpers.oracle_sv_values(cc, oo)
    out = []
    for tile_bits, g in configs:
        prog = engine.SvxProgram(fb, 0, tile_bits, g)
        assert prog.info["status"] == 0, (tile_bits, g)
        vals, _ = run(prog.info)
        assert np.max(np.abs(vals - ref)) <= TOL, (tile_bits, g, vals, ref)
        out.append(prog.info)
    return out


def test_random_basis_circuits_single_and_sharded(lib):
    rng = np.random.default_rng(21)
    for n in (3, 6, 9):
        cm = [(i, i + 1) for i in range(n - 1)] + [(i + 1, i) for i in range(n - 1)]
        for _ in range(3):
            c = F.random_basis_circuit(n, int(rng.integers(5, 90)), rng, cm)
            obs = [[(l, float(rng.normal()))] for l in _labels(rng, n, 5)] + [[("Z" * n, 0.5), ("X" * n, -1.0), ("I" * n, 2.0)]]
            _check(c, obs, [(11, 0), (4, 0), (3, 0), (5, 1), (4, 2)] + ([(4, 3)] if n >= 8 else []))


def test_tfim_fuses_bonds_to_diagonals_and_shards(lib):
    n = 10
    c = F.tfim_circuit(n, 3, 0.37, basis="X")
    obs = F.tfim_observables(list(range(n)), n)
    infos = _check(c, obs, [(11, 0), (6, 0), (6, 1), (5, 3)])
    # whole state resident: one sweep for the circuit; the XX family costs one more
    assert len(infos[0]["sweeps"]) <= 3
    assert infos[2]["n_exchanges"] >= 1 and infos[3]["n_exchanges"] >= 1


def test_general_two_qubit_gates_and_controls(lib):
    rng = np.random.default_rng(5)
    n = 7
    c = Circuit(n)
    for _ in range(60):
        a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
        kind = rng.integers(0, 9)
        if kind == 0:
            c.append("cz", (a, b))
        elif kind == 1:
            c.append("swap", (a, b))
        elif kind == 2:
            c.append("crx", (a, b), (float(rng.uniform(-3, 3)),))
        elif kind == 3:
            c.append("rzz", (a, b), (float(rng.uniform(-3, 3)),))
        elif kind == 4:
            c.append("rxx", (a, b), (float(rng.uniform(-3, 3)),))
        elif kind == 5:
            c.append("cp", (a, b), (float(rng.uniform(-3, 3)),))
        elif kind == 6:
            c.append("ecr", (a, b))
        elif kind == 7:
            c.append("u3", (a,), tuple(float(x) for x in rng.uniform(-3, 3, size=3)))
        else:
            c.append("cx", (a, b))
    obs = [[(l, 1.0)] for l in _labels(rng, n, 8)]
    _check(c, obs, [(11, 0), (4, 0), (5, 1), (3, 2)])


def test_idle_qubits_empty_circuit_and_padding(lib):
    c = Circuit(6)
    c.append("h", (1,))
    c.append("cx", (1, 4))
    obs = [[("IZIIZI", 1.0)], [("IXIIXI", 1.0)], [("ZIIIII", 1.0)], [("XIIIII", 1.0), ("IIIIII", 0.25)]]
    _check(c, obs, [(11, 0), (11, 1), (11, 2)])
    e = Circuit(3)
    fb = engine.encode_batch([e], [[[("ZZZ", 1.0)], [("XII", 1.0)]]])
    prog = engine.SvxProgram(fb, 0, 0, 0)
    vals, _ = run(prog.info)
    assert np.allclose(vals, [1.0, 0.0])


def test_exchange_count_tfim_chain_30q_plan(lib):
    """Planner only (no emulation): 30-qubit TFIM over 8 ranks needs about one exchange per
    Trotter step, and every bond became a diagonal op (no pass on a global qubit pair)."""
    n = 30
    c = F.tfim_circuit(n, 4, 0.5, basis="Z")
    obs = F.tfim_observables(list(range(n)), n)
    fb = engine.encode_batch([c], [obs])
    prog = engine.SvxProgram(fb, 0, 0, 3)
    info = prog.info
    assert info["status"] == 0 and info["n_local"] == 27 and info["n_global"] == 3
    assert 1 <= info["n_exchanges"] <= 12


def test_direct_pass_flags_of_statevector_sweeps(lib):
    """sv_sweep_kernel direct passes: the flagged first / last pass of a sweep uses free slots
    only (>= the always-resident low bits) and the high half of the sweep's block length carries
    the first-pass flag.  (The reordered programs are run by the emulator in the other tests.)"""
    from ml_qem_b200.engine import SvxProgram
    from svx_emulator import decode_block

    n = 20
    circ = F.tfim_circuit(n, 3, 0.4)
    prog = SvxProgram(engine.encode_batch([circ], [F.tfim_observables(list(range(n)), n)]), 0, 12, 0)
    info = prog.info
    LB = max(0, info["tile_bits"] - 8)
    n_first = n_last = 0
    for sw in info["sweeps"]:
        desc = (int(sw[9]) >> 16) & 0xffff
        passes, _ = decode_block(info, sw)
        flags = [p["flags"] for p in passes]
        assert all(f == 0 for f in flags[1:-1])
        if flags[0] & 1:
            assert all(x >= LB for x in passes[0]["s"]) and desc == 0x8000
            n_first += 1
        else:
            assert desc & 0x8000 == 0
        if flags[-1] & 2:
            assert all(x >= LB for x in passes[-1]["s"])
            n_last += 1
    assert n_first > 0 and n_last > 0


def test_four_slot_passes_and_fast_signatures(lib):
    """Register passes own four tile slots; the rx / h layers of the Trotter circuits become fast
    passes (straight-line kernel bodies) with up to four structured 1-qubit ops, the ZZ bonds of a
    layer one fused diagonal op."""
    from ml_qem_b200.engine import SvxProgram
    from svx_emulator import SVO_DZZ, SVO_X1, SVS_GENERIC, decode_block

    n = 24
    circ = F.tfim_circuit(n, 4, 0.4, dt=0.25)
    info = SvxProgram(engine.encode_batch([circ], [F.tfim_observables(list(range(n)), n)]), 0, 12, 0).info
    n_pass = n_fast = n_x1 = n_dzz = 0
    for sw in info["sweeps"]:
        for ph in decode_block(info, sw)[0]:
            n_pass += 1
            n_fast += ph["sig"] != SVS_GENERIC
            n_x1 += sum(o["kind"] == SVO_X1 for o in ph["ops"])
            n_dzz += sum(o["kind"] == SVO_DZZ for o in ph["ops"])
    assert n_x1 == 4 * n and n_dzz == 4
    assert n_fast == n_pass                      # no generic pass in a TFIM circuit
    assert n_pass <= (4 * n + n) / 4 + 8 + 4     # about four slot ops per pass
