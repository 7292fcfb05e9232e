"""Variant generation inside the library (bwq_variants: ZNE folds, Pauli twirls) against the
Python-level construction of the same variants -- host side, no GPU."""
import numpy as np
import pytest

from ml_qem_b200 import engine, families as F, zne
from ml_qem_b200.engine import Variants, _num_params_table

M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def twirl_draw(seed, c, t, k):
    """Restatement of the library's counter-based generator (csrc/variants.cpp: twirl_draw)."""
    h = splitmix64(splitmix64(splitmix64((seed ^ c) & M64) ^ t) ^ k)
    return h & 3, (h >> 2) & 3


class _Replay:
    """numpy Generator stand-in feeding families._twirl_pair the library's draws."""

    def __init__(self, seed, c, t):
        self.seed, self.c, self.t, self.k, self.pending = seed, c, t, 0, None

    def integers(self, lo, hi):
        if self.pending is None:
            pc, pt = twirl_draw(self.seed, self.c, self.t, self.k)
            self.k += 1
            self.pending = pt
            return pc
        pt, self.pending = self.pending, None
        return pt


def _gates(b, c):
    t = _num_params_table()
    out = []
    for o in b.ops[b.op_offsets[c]:b.op_offsets[c + 1]]:
        p = b.params[int(o["param_idx"]):int(o["param_idx"]) + t[o["opcode"]]]
        out.append((int(o["opcode"]), int(o["q0"]), int(o["q1"]) if o["opcode"] >= 32 else 0, tuple(np.round(p, 12))))
    return out


def test_library_twirls_and_folds_equal_the_python_built_circuits(lib):
    seed = 1234
    base = [F.tfim_circuit(4, 2, 0.37, layout=[0, 1, 3, 4], num_physical=5),
            F.brickwork_circuit(6, 2, np.random.default_rng(3), num_physical=8)]
    obs = [F.single_z_observables([0, 1, 3, 4], 5), F.single_z_observables(list(range(6)), 8)]
    fb = engine.encode_batch(base, obs)
    v = Variants(folds=(1, 3), twirls=3, seed=seed)
    ex = engine.expand_variants(fb, v)
    assert ex.n_circuits == 2 * 6 and ex.n_observables == 6 * (4 + 6)
    builders = [lambda fold, rng: F.tfim_circuit(4, 2, 0.37, layout=[0, 1, 3, 4], num_physical=5, fold=fold, twirl_rng=rng),
                lambda fold, rng: F.brickwork_circuit(6, 2, np.random.default_rng(3), num_physical=8, fold=fold, twirl_rng=rng)]
    for c in range(2):
        for fi, fold in enumerate((1, 3)):
            for t in range(3):
                ref = engine.encode_batch([builders[c](fold, _Replay(seed, c, t))], [obs[c]])
                got = _gates(ex, c * 6 + fi * 3 + t)
                # the Python builder merges consecutive rz on a qubit; compare per-qubit rz-merged streams
                assert _merge_rz(got) == _merge_rz(_gates(ref, 0)), (c, fold, t)
    # folds only == zne.fold_batch on the flat stream
    ex2 = engine.expand_variants(fb, Variants(folds=(1, 5)))
    f5 = zne.fold_batch(fb, 5)
    for c in range(2):
        assert _gates(ex2, 2 * c) == _gates(fb, c) and _gates(ex2, 2 * c + 1) == _gates(f5, c)
    with pytest.raises(ValueError):
        Variants(folds=(2,))


def _merge_rz(gates):
    """Canonical form: consecutive rz on one qubit merged (mod 4 pi), zero rotations dropped."""
    out, pend = [], {}

    def flush(q):
        a = pend.pop(q, None)
        if a is not None:
            a = float(np.remainder(a + 2 * np.pi, 4 * np.pi) - 2 * np.pi)
            if abs(a) > 1e-9 and abs(abs(a) - 4 * np.pi) > 1e-9:
                out.append(("rz", q, round(a, 9)))

    for op, q0, q1, p in gates:
        if op == 13:  # rz
            pend[q0] = pend.get(q0, 0.0) + p[0]
            continue
        flush(q0)
        if op >= 32:
            flush(q1)
        out.append((op, q0, q1, p))
    for q in sorted(pend):
        flush(q)
    return out


def test_folding_rules_for_parametrised_two_qubit_gates(lib):
    from ml_qem_b200 import Circuit

    c = Circuit(3)
    c.rzz(0.4, 0, 1); c.cp(1.1, 1, 2); c.cu3(0.3, 0.2, -0.7, 0, 2); c.cz(0, 1); c.h(2)
    fb = engine.encode_batch([c], [[[("ZZZ", 1.0)]]])
    ex = engine.expand_variants(fb, Variants(folds=(3,)))
    ref = zne.fold_batch(fb, 3)
    assert _gates(ex, 0) == _gates(ref, 0)


def test_fold_aware_lowering_equals_lowering_of_the_folded_circuit(lib):
    """bwq_*_variants lower the gates of a cx-only circuit once and repeat the cx ops per fold: the
    program must evolve the state exactly like the program of the explicitly folded circuit -- with the
    fused cx + relaxation error (device model) and with a cx followed by a separate dense error
    (coherent cx noise), on chip and tiled."""
    import helpers
    from program_emulator import run_program
    from ml_qem_b200 import backends, noise

    lima = backends.fake_lima()
    models = [noise.from_backend(lima), noise.add_coherent_noise(lima, theta=0.04 * np.pi, seed=0)[0]]
    c5 = F.tfim_circuit(4, 3, 0.37, basis="Y", layout=[0, 1, 3, 4], num_physical=5, random_init_prefix=True)
    o5 = F.single_z_observables([0, 1, 3, 4], 5)
    be8 = backends.synthetic_chain(8, seed=5)
    c8 = F.brickwork_circuit(8, 2, np.random.default_rng(1), twirl_rng=np.random.default_rng(2))
    o8 = F.single_z_observables(list(range(8)), 8)
    for circ, obs, nms, kw in ((c5, o5, models, {}), (c8, o8, [noise.from_backend(be8)], {"tma": True}), (c8, o8, [noise.from_backend(be8)], {})):
        fb = engine.encode_batch([circ], [obs])
        for nm in nms:
            for fold in (3, 5):
                folded = zne.fold_batch(fb, fold)
                ref = engine.lower_dm(folded, 0, nm, **kw)
                got = engine.lower_dm(fb, 0, nm, fold=fold, **kw)
                assert got["status"] == 0 and got["n_gates"] == ref["n_gates"]
                assert np.max(np.abs(run_program(got) - run_program(ref))) <= 1e-14
    # a circuit with another 2-qubit gate has no fold-aware program (the expanded stream is lowered)
    from ml_qem_b200 import Circuit
    c = Circuit(3); c.cx(0, 1); c.rzz(0.3, 1, 2)
    assert engine.lower_dm(engine.encode_batch([c], [[[("ZZZ", 1.0)]]]), 0, None, fold=3)["status"] == 1
