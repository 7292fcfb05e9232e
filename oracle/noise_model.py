"""Device noise model, restated from qiskit-aer (not in tree).

[3P] ``qiskit_aer.noise.NoiseModel.from_backend`` -> ``basic_device_gate_errors`` with the
defaults the reference uses (``AerSimulator.from_backend(backend)`` at
blackwater/data/utils.py:427; docs/tutorials/noise_utils.py:37,49,77,93;
docs/tutorials/mbd_utils.py:103,116): gate_error=True, thermal_relaxation=True,
temperature=0.  In-tree editors restated: docs/tutorials/noise_utils.py:36-144
(RemoveReadoutErrors, AddNoise) and docs/tutorials/mbd_utils.py:95-137.

Every error is a superoperator S (4^k x 4^k complex) in Aer's column-stacking convention:
vec(rho)[r + c*2^k], local basis index r = sum_i bit_i 2^i with i the position in the gate's
qubit tuple; Kraus {K} -> S = sum conj(K) (x) K.

Backend calibration input ("props") is the plain-dict form of BackendProperties.to_dict()
(docs/demos/fake_backend_info.ipynb:51): {"qubits": [ {name: {"value", "unit"}} ... ],
"gates": [ {"gate", "qubits", "parameters": {"gate_error": {...}, "gate_length": {...}}} ]}.
"""
import math

import numpy as np

from . import gates as G

_UNIT = {"s": 1.0, "ms": 1e-3, "us": 1e-6, "µs": 1e-6, "ns": 1e-9, "ps": 1e-12, "": 1.0,
         "GHz": 1e9, "MHz": 1e6, "kHz": 1e3, "Hz": 1.0}


def _val(entry):
    return float(entry["value"]) * _UNIT.get(entry.get("unit", ""), 1.0)


# ---------------------------------------------------------------- channel algebra
def kraus_to_superop(kraus):
    return sum(np.kron(np.conj(k), k) for k in kraus)


def unitary_superop(u):
    return np.kron(np.conj(u), u)


def tensor_superops(s_a, s_b):
    """Superop of channel a on local qubit 0 and channel b on local qubit 1."""
    a = s_a.reshape(2, 2, 2, 2)  # [c_a, r_a, c_a', r_a']
    b = s_b.reshape(2, 2, 2, 2)
    out = np.einsum("aAbB,cCdD->caCAdbDB", a, b)  # [c_b,c_a,r_b,r_a ; primed]
    return out.reshape(16, 16)


def embed_1q_in_2q(s, which):
    eye = np.eye(4, dtype=complex)
    return tensor_superops(s, eye) if which == 0 else tensor_superops(eye, s)


def depolarizing_superop(p, k):
    """[3P] qiskit_aer.noise.depolarizing_error(p, k): (1-p) rho + p Tr(rho) I/d."""
    d = 2 ** k
    vec_i = np.eye(d, dtype=complex).reshape(-1)
    return (1 - p) * np.eye(d * d, dtype=complex) + (p / d) * np.outer(vec_i, vec_i)


def depolarizing_probabilities(p, k):
    """Pauli-mixture probabilities [p_I, p_other...] Aer stores for depolarizing_error."""
    nt = 4 ** k
    return [1 - p * (nt - 1) / nt] + [p / nt] * (nt - 1)


def thermal_relaxation_params(t1, t2, time):
    p_reset = 1.0 - math.exp(-time / t1) if math.isfinite(t1) else 0.0
    e2 = math.exp(-time / t2) if math.isfinite(t2) else 1.0
    return p_reset, e2


def thermal_relaxation_superop(t1, t2, time):
    """[3P] thermal_relaxation_error(t1, t2, time, excited_state_population=0).

    rho00 += p_reset rho11; rho11 *= 1-p_reset; off-diagonals *= exp(-t/T2).  Aer emits a
    Kraus set (T2 > T1, via the Choi matrix) or the mixture {I, Z, reset} (T2 <= T1); both are
    this channel.
    """
    if t2 > 2 * t1 * (1 + 1e-12):
        raise ValueError("thermal_relaxation: T2 > 2 T1")
    pr, e2 = thermal_relaxation_params(t1, t2, time)
    s = np.zeros((4, 4), dtype=complex)
    s[0, 0] = 1.0
    s[0, 3] = pr
    s[1, 1] = e2
    s[2, 2] = e2
    s[3, 3] = 1.0 - pr
    return s


def thermal_relaxation_mixture(t1, t2, time):
    """Mixture probabilities [p_I, p_Z, p_reset] Aer stores when T2 <= T1 (else None)."""
    if t2 > t1:
        return None
    pr, e2 = thermal_relaxation_params(t1, t2, time)
    e1 = math.exp(-time / t1)
    p_z = (1 - pr) * (1 - e2 / e1) / 2
    return [1 - p_z - pr, p_z, pr]


def process_fidelity(s):
    d2 = s.shape[0]
    return float(np.real(np.trace(s))) / d2


def average_gate_fidelity(s):
    """[3P] qiskit.quantum_info.average_gate_fidelity(channel) against the identity."""
    d = int(round(math.sqrt(s.shape[0])))
    return (d * process_fidelity(s) + 1) / (d + 1)


RESET_SUPEROP = np.array([[1, 0, 0, 1], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]], dtype=complex)


# ---------------------------------------------------------------- the model object
class NoiseModel:
    """(gate name, physical qubit tuple) -> superop; optional all-qubit defaults per name.

    Lookup follows Aer: a local error for the exact ordered qubit tuple overrides the
    all-qubit default for that instruction.
    """

    def __init__(self):
        self.local = {}
        self.default = {}
        self.readout = {}
        self.info = {}  # (name, qubits) -> dict of intermediate quantities, for KAT tests

    def get(self, name, qubits):
        key = (name, tuple(qubits))
        if key in self.local:
            return self.local[key]
        return self.default.get(name)

    def copy(self):
        m = NoiseModel()
        m.local = dict(self.local)
        m.default = dict(self.default)
        m.readout = dict(self.readout)
        m.info = dict(self.info)
        return m


def qubit_relaxation(props):
    out = []
    for q in props["qubits"]:
        t1 = _val(q["T1"]) if "T1" in q else math.inf
        t2 = _val(q["T2"]) if "T2" in q else 2 * t1
        out.append((t1, min(t2, 2 * t1)))  # _truncate_t2_value
    return out


def from_backend(props):
    """NoiseModel.from_backend(backend) restated (Appendix A.2 of SURVEY.md)."""
    relax = qubit_relaxation(props)
    model = NoiseModel()
    for g in props["gates"]:
        name, qubits = g["gate"], tuple(g["qubits"])
        par = g["parameters"]
        gate_error = _val(par["gate_error"]) if "gate_error" in par else None
        gate_len = _val(par["gate_length"]) if "gate_length" in par else 0.0
        k = len(qubits)
        relax_s = None
        info = {"gate_error": gate_error, "gate_length": gate_len}
        if gate_len and gate_len > 0:
            per = [thermal_relaxation_superop(relax[q][0], relax[q][1], gate_len) for q in qubits]
            info["relax_mixtures"] = [thermal_relaxation_mixture(relax[q][0], relax[q][1], gate_len) for q in qubits]
            if k == 1:
                relax_s = per[0]
            elif k == 2:
                relax_s = tensor_superops(per[0], per[1])
            else:
                raise ValueError("from_backend: >2-qubit gate entries unsupported")
        relax_fid = average_gate_fidelity(relax_s) if relax_s is not None else 1.0
        relax_infid = 1.0 - relax_fid
        depol_s = None
        if gate_error is not None and gate_error > relax_infid:
            dim = 2 ** k
            err = min(gate_error, dim / (dim + 1))
            p = dim * (err - relax_infid) / (dim * relax_fid - 1)
            p = min(p, 4 ** k / (4 ** k - 1))
            depol_s = depolarizing_superop(p, k)
            info["depol_param"] = p
            info["depol_probabilities"] = depolarizing_probabilities(p, k)
        if depol_s is None and relax_s is None:
            continue
        if depol_s is None:
            s = relax_s
        elif relax_s is None:
            s = depol_s
        else:
            s = relax_s @ depol_s  # depol_error.compose(relax_error): depolarizing first
        model.local[(name, qubits)] = s
        model.info[(name, qubits)] = info
    for i, q in enumerate(props["qubits"]):
        if "prob_meas1_prep0" in q and "prob_meas0_prep1" in q:
            p10, p01 = _val(q["prob_meas1_prep0"]), _val(q["prob_meas0_prep1"])
            model.readout[i] = np.array([[1 - p10, p10], [p01, 1 - p01]])
    return model


def coupling_map_from_props(props):
    """Sorted directed pairs, the order FakeLima/FakeBelem's configuration().coupling_map uses."""
    return sorted({tuple(g["qubits"]) for g in props["gates"] if len(g["qubits"]) == 2})


def controlled_rx_error_unitary(theta):
    """docs/tutorials/noise_utils.py:97-101: (I(x)|0><0| + i RX(pi+theta)(x)|1><1|) @ CX."""
    up = 0.5 * (G.I2 + G.Z)
    down = 0.5 * (G.I2 - G.Z)
    over = np.kron(G.I2, up) + 1j * np.kron(G.gate_matrix("rx", [math.pi + theta]), down)
    return over @ G.gate_matrix("cx")


def _raw_cx_depol_therm(props, pair):
    """noise_utils.py:103-113: raw gate_error depolarizing + two 1-qubit thermal errors."""
    raw = {}
    for g in props["gates"]:
        if g["gate"] == "cx" and tuple(g["qubits"]) == tuple(pair):
            raw = g["parameters"]
    mag, t = _val(raw["gate_error"]), _val(raw["gate_length"])
    th = []
    for q in pair:
        t1, t2 = _val(props["qubits"][q]["T1"]), _val(props["qubits"][q]["T2"])
        th.append(thermal_relaxation_superop(t1, t2, t))
    return depolarizing_superop(mag, 2), th[0], th[1]


def add_coherent_noise(props, theta, uniform=False, add_depolarization=True, seed=None,
                       add_coherent=True, coupling_map=None):
    """docs/tutorials/noise_utils.py:69-144 (AddNoise.add_coherent_noise) restated.

    Quirk kept on purpose: ``coherent.compose(depol).compose(therm_q0)
    .compose(therm_q1)`` passes no qargs, so BOTH 1-qubit thermal errors land on local qubit 0.
    """
    if seed is not None:
        np.random.seed(seed)  # noise_utils.py:125 fix_random_seed
    model = from_backend(props)
    for key in [k for k in model.local if k[0] == "cx"]:
        del model.local[key]
    pairs = [tuple(p) for p in (coupling_map or coupling_map_from_props(props))]
    thetas = None

    def composite(pair, th, calib_pair):
        s = np.eye(16, dtype=complex)
        if th is not None:
            s = unitary_superop(controlled_rx_error_unitary(th))
        if add_depolarization or th is None:
            depol, t0, t1 = _raw_cx_depol_therm(props, calib_pair)
            s = embed_1q_in_2q(t1, 0) @ embed_1q_in_2q(t0, 0) @ depol @ s
        return s

    if add_coherent:
        if uniform:
            model.default["cx"] = composite(None, theta, pairs[0])
        else:
            thetas = np.random.uniform(0, theta, size=len(pairs))
            for pair, th in zip(pairs, thetas):
                model.local[("cx", pair)] = composite(pair, th, pair)
    else:  # restore_incoherent, noise_utils.py:138-144
        for pair in pairs:
            model.local[("cx", pair)] = composite(pair, None, pair)
    model.info["thetas"] = thetas
    return model


def modify_and_add_noise_to_model(props, theta=math.pi / 8):
    """docs/tutorials/mbd_utils.py:95-137: cx errors dropped, all-qubit coherent CX error."""
    model = from_backend(props)
    for key in [k for k in model.local if k[0] == "cx"]:
        del model.local[key]
    model.default["cx"] = unitary_superop(controlled_rx_error_unitary(theta))
    return model
