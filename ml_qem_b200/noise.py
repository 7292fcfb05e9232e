"""Device noise model in Pauli-transfer form -- the engine-side equivalent of
``NoiseModel.from_backend(backend)`` / ``AerSimulator.from_backend(backend)`` that the reference
builds at blackwater/data/utils.py:427 and edits in docs/tutorials/noise_utils.py:36-144 and
docs/tutorials/mbd_utils.py:95-137.

Semantics ([3P] qiskit-aer basic_device_gate_errors, defaults):
  per gate entry (name, qubits):  error = thermal_relaxation(T1, min(T2, 2 T1), gate_length) on
  each gate qubit, preceded by a depolarizing channel sized so the total average gate infidelity
  equals the reported gate_error (only when gate_error exceeds the relaxation infidelity).
  Errors are applied after their gate; readout errors never act in exact (shots=None) mode.
"""
import math

import numpy as np

from . import ptm
from .backends import BackendProps
from .gateset import OPCODES, canonical, is_two_qubit

NOISE_DENSE1, NOISE_DENSE2, NOISE_RELAX2 = 1, 2, 3


class NoiseModel:
    """(gate, physical qubits) -> PTM (4x4 or 16x16).  ``default[name]`` = all-qubit error."""

    def __init__(self, name="noise"):
        self.name = name
        self.local = {}
        self.default = {}
        self.readout = {}
        self._table = None

    def add_quantum_error(self, r, gate, qubits):
        self.local[(canonical(gate), tuple(qubits))] = np.asarray(r, dtype=float)
        self._table = None

    def add_all_qubit_quantum_error(self, r, gates):
        for g in ([gates] if isinstance(gates, str) else gates):
            self.default[canonical(g)] = np.asarray(r, dtype=float)
        self._table = None

    def remove(self, gate):
        gate = canonical(gate)
        for key in [k for k in self.local if k[0] == gate]:
            del self.local[key]
        self.default.pop(gate, None)
        self._table = None

    def get(self, gate, qubits):
        key = (canonical(gate), tuple(qubits))
        if key in self.local:
            return self.local[key]
        return self.default.get(key[0])

    def copy(self):
        m = NoiseModel(self.name)
        m.local, m.default, m.readout = dict(self.local), dict(self.default), dict(self.readout)
        return m

    def is_ideal(self):
        return not self.local and not self.default

    # -- C-ABI noise table (include/bwq.h: bwq_noise_table)
    def to_table(self):
        if self._table is not None:
            return self._table
        opcode, q0, q1, kind, off, data = [], [], [], [], [], []
        pos = 0

        def put(name, qubits, r):
            nonlocal pos
            if name not in OPCODES:
                return
            two = is_two_qubit(name)
            r = np.asarray(r, dtype=float)
            if r.shape != ((16, 16) if two else (4, 4)):
                raise ValueError(f"noise on {name}{qubits}: PTM shape {r.shape} does not match the gate")
            if two:
                rp = ptm.relax2_params(r)
                k, payload = (NOISE_RELAX2, rp) if rp is not None else (NOISE_DENSE2, r.reshape(-1))
            else:
                k, payload = NOISE_DENSE1, r.reshape(-1)
            opcode.append(OPCODES[name])
            q0.append(qubits[0] if qubits else 255)
            q1.append(qubits[1] if qubits and len(qubits) > 1 else 255)
            kind.append(k)
            off.append(pos)
            data.append(payload)
            pos += len(payload)

        for (name, qubits), r in self.local.items():
            put(name, qubits, r)
        for name, r in self.default.items():
            put(name, None, r)
        self._table = {
            "opcode": np.asarray(opcode, dtype=np.uint16), "q0": np.asarray(q0, dtype=np.uint8),
            "q1": np.asarray(q1, dtype=np.uint8), "kind": np.asarray(kind, dtype=np.uint8),
            "data_off": np.asarray(off, dtype=np.int64),
            "data": np.concatenate(data) if data else np.zeros(0),
        }
        return self._table


def from_backend(backend, gate_error=True, thermal_relaxation=True):
    """NoiseModel.from_backend(backend) -- accepts BackendProps, a properties dict, the dict of
    get_backend_properties_v1 (utils.py:139-175) or a Qiskit-like backend object."""
    props = BackendProps.from_backend(backend)
    model = NoiseModel(props.name)
    for (name, qubits), (err, length) in props.gates.items():
        k = len(qubits)
        if k > 2:
            raise ValueError("noise model: gates on more than 2 qubits are not supported")
        relax = None
        if thermal_relaxation and length and length > 0:
            per = [ptm.thermal_relaxation(props.t1[q], min(props.t2[q], 2 * props.t1[q]), length) for q in qubits]
            relax = per[0] if k == 1 else ptm.tensor(per[0], per[1])
        relax_fid = ptm.average_gate_fidelity(relax) if relax is not None else 1.0
        relax_infid = 1.0 - relax_fid
        depol = None
        if gate_error and err is not None and err > relax_infid:
            dim = 2 ** k
            e = min(err, dim / (dim + 1))
            p = dim * (e - relax_infid) / (dim * relax_fid - 1)
            p = min(p, 4 ** k / (4 ** k - 1))
            depol = ptm.depolarizing(p, k)
        if relax is None and depol is None:
            continue
        r = relax if depol is None else depol if relax is None else relax @ depol  # depolarizing first
        model.local[(canonical(name), tuple(qubits))] = r
    model.readout = dict(props.readout)
    return model


def remove_readout_errors(backend):
    """docs/tutorials/noise_utils.py:36-51.  Exact mode never samples a measurement, so this is
    ``from_backend`` with the readout table dropped."""
    m = from_backend(backend)
    m.readout = {}
    return m


def _controlled_rx_error(theta):
    """noise_utils.py:97-101: (I(x)|0><0| + i RX(pi+theta)(x)|1><1|) @ CX = controlled-RX(theta)."""
    up, down = np.diag([1.0, 0.0]).astype(complex), np.diag([0.0, 1.0]).astype(complex)
    a = (math.pi + theta) / 2
    rx = np.array([[math.cos(a), -1j * math.sin(a)], [-1j * math.sin(a), math.cos(a)]])
    cx = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=complex)
    return (np.kron(np.eye(2), up) + 1j * np.kron(rx, down)) @ cx


def add_coherent_noise(backend, theta, uniform=False, add_depolarization=True, seed=None, add_coherent=True):
    """AddNoise(backend).add_coherent_noise(...) of docs/tutorials/noise_utils.py:69-144.

    Reproduces the reference exactly, including that both 1-qubit thermal errors of the composite
    cx error act on the error's qubit 0 (``.compose`` without qargs, noise_utils.py:120,131,141).
    Returns (model, thetas)."""
    props = BackendProps.from_backend(backend)
    if seed is not None:
        np.random.seed(seed)  # noise_utils.py: fix_random_seed
    model = from_backend(props)
    model.remove("cx")
    pairs = list(props.coupling_map)

    def composite(theta_pair, calib):
        r = np.eye(16)
        if theta_pair is not None:
            r = ptm.from_unitary(_controlled_rx_error(theta_pair))
        if add_depolarization or theta_pair is None:
            err, length = props.gates[("cx", tuple(calib))]
            th0 = ptm.thermal_relaxation(props.t1[calib[0]], props.t2[calib[0]], length)
            th1 = ptm.thermal_relaxation(props.t1[calib[1]], props.t2[calib[1]], length)
            r = ptm.embed(th1, 0) @ ptm.embed(th0, 0) @ ptm.depolarizing(err, 2) @ r
        return r

    thetas = None
    if add_coherent:
        if uniform:
            model.add_all_qubit_quantum_error(composite(theta, pairs[0]), "cx")
        else:
            thetas = np.random.uniform(0, theta, size=len(pairs))
            for pair, th in zip(pairs, thetas):
                model.add_quantum_error(composite(th, pair), "cx", pair)
    else:
        for pair in pairs:
            model.add_quantum_error(composite(None, pair), "cx", pair)
    return model, thetas


def modify_and_add_noise_to_model(backend, theta=math.pi / 8):
    """docs/tutorials/mbd_utils.py:95-137: cx errors replaced by an all-qubit coherent over-rotation."""
    model = from_backend(backend)
    model.remove("cx")
    model.add_all_qubit_quantum_error(ptm.from_unitary(_controlled_rx_error(theta)), "cx")
    return model
