"""Shared test helpers: golden fixtures, oracle adapters (test infrastructure only)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def oracle_noise(backend_key):
    from oracle import noise_model as onm

    return onm.from_backend(golden("backends.json")[backend_key])


def oracle_dm_values(circ, observables, onoise):
    """circ: ml_qem_b200 Circuit; observables: list of [(label, coeff)...]."""
    from oracle import dm

    return dm.estimate(circ.num_qubits, circ.gate_ops(), observables, onoise)


def oracle_sv_values(circ, observables):
    from oracle import sv

    return sv.estimate(circ.num_qubits, circ.gate_ops(), observables)


def compact(circ, observables, onoise=None):
    """Drops untouched qubits (the oracle is exponential in the register width): returns an
    equivalent (circuit, observables[, oracle noise model]) on the active qubits only, with the
    noise model's physical-qubit keys renamed accordingly; X/Y on idle qubits zero a term."""
    from ml_qem_b200.circuit import Circuit

    used = sorted({q for _, qs, _ in circ.gate_ops() for q in qs})
    pos = {q: i for i, q in enumerate(used)}
    n = max(len(used), 1)
    c = Circuit(n)
    for name, qs, p in circ.gate_ops():
        c.ops.append((name, tuple(pos[q] for q in qs), p))
    out = []
    w = circ.num_qubits
    for ob in observables:
        terms = []
        for label, coeff in ob:
            chars = ["I"] * n
            dead = False
            for q in range(w):
                ch = label[w - 1 - q]
                if ch == "I":
                    continue
                if q in pos:
                    chars[n - 1 - pos[q]] = ch
                elif ch in "XY":
                    dead = True
            terms.append(("".join(chars), 0.0 if dead else coeff))
        out.append(terms)
    if onoise is None:
        return c, out
    from oracle.noise_model import NoiseModel
    m = NoiseModel()
    m.default = dict(onoise.default)
    for (name, qs), sup in onoise.local.items():
        if all(q in pos for q in qs):
            m.local[(name, tuple(pos[q] for q in qs))] = sup
    return c, out, m
