"""Synthetic circuit families of the reference's data-generation notebooks, emitted directly in
the backend basis {rz, sx, x, cx} on physical qubits (what ``transpile(circuit, backend)`` hands
to the simulator in the reference), plus the ZNE-folded and Pauli-twirled variants.

  TFIM Trotter     docs/tutorials/h13_ising_data_gen.ipynb:224-300 (IsingModel; rx layer, even
                   bonds cx-rz-cx, odd bonds, final X/Y/Z basis change), random-init prefix :404
  random brickwork docs/tutorials/mbd_utils.py:414-470 (completely_random=True: x on odd qubits,
                   per step cz + two u3 per bond on even then odd bonds, random phase per qubit)
  ZNE folding      docs/tutorials/zne_parallel.py:172-183 (LocalFoldingAmplifier(gates_to_fold=2):
                   every cx becomes cx^lambda, lambda odd)
  Pauli twirling   docs/tutorials/derek_files/phase_diagram.ipynb:776 (add_pauli_twirls: a random
                   2-qubit Pauli before each cx and its CX-conjugate after)

Basis decompositions (global phase dropped): rx(t) = rz(pi/2) sx rz(t+pi) sx rz(5pi/2);
u3(t,p,l) = rz(l) sx rz(t+pi) sx rz(p+3pi); h = rz(pi/2) sx rz(pi/2); cz = h_t cx h_t;
p(l) = rz(l); sdg = rz(-pi/2).  tests/test_host_regressions.py checks them against the oracle gate table.
"""
import math

import numpy as np

from .circuit import Circuit

PI = math.pi


class BasisBuilder:
    """Appends basis gates to a Circuit, merging consecutive rz on a qubit (as transpile does)."""

    def __init__(self, num_qubits):
        self.circ = Circuit(num_qubits)
        self._rz = {}

    def _flush(self, q):
        a = self._rz.pop(q, None)
        if a is not None:
            a = math.remainder(a, 4 * PI)
            if abs(a) > 1e-15:
                self.circ.append("rz", (q,), (a,))

    def rz(self, a, q):
        self._rz[q] = self._rz.get(q, 0.0) + a

    def sx(self, q):
        self._flush(q)
        self.circ.append("sx", (q,))

    def x(self, q):
        self._flush(q)
        self.circ.append("x", (q,))

    def cx(self, c, t, fold=1):
        self._flush(c)
        self._flush(t)
        for _ in range(fold):
            self.circ.append("cx", (c, t))

    def rx(self, a, q):
        self.rz(PI / 2, q); self.sx(q); self.rz(a + PI, q); self.sx(q); self.rz(5 * PI / 2, q)

    def u3(self, t, p, l, q):
        self.rz(l, q); self.sx(q); self.rz(t + PI, q); self.sx(q); self.rz(p + 3 * PI, q)

    def h(self, q):
        self.rz(PI / 2, q); self.sx(q); self.rz(PI / 2, q)

    def pauli(self, p, q):
        """p: 0=I 1=X 2=Y 3=Z."""
        if p == 1:
            self.x(q)
        elif p == 2:
            self.rz(PI, q); self.x(q)
        elif p == 3:
            self.rz(PI, q)

    def done(self):
        for q in list(self._rz):
            self._flush(q)
        return self.circ


def _twirl_pair(rng):
    """Random 2-qubit Pauli (pc, pt) and its image under CX conjugation (sign dropped)."""
    pc, pt = int(rng.integers(0, 4)), int(rng.integers(0, 4))
    sym = {0: (0, 0), 1: (1, 0), 2: (1, 1), 3: (0, 1)}
    inv = {v: k for k, v in sym.items()}
    xc, zc = sym[pc]
    xt, zt = sym[pt]
    return (pc, pt), (inv[(xc, zc ^ zt)], inv[(xt ^ xc, zt)])


class _Emitter(BasisBuilder):
    def __init__(self, num_qubits, fold=1, twirl_rng=None):
        super().__init__(num_qubits)
        self.fold = fold
        self.twirl_rng = twirl_rng

    def noisy_cx(self, c, t):
        if self.twirl_rng is None:
            self.cx(c, t, self.fold)
            return
        (pc, pt), (qc, qt) = _twirl_pair(self.twirl_rng)
        self.pauli(pc, c); self.pauli(pt, t)
        self.cx(c, t, self.fold)
        self.pauli(qc, c); self.pauli(qt, t)


def chain_layout(n, backend=None):
    """Physical qubits of an n-qubit line; on ibmq_lima/belem (T shape 0-1-{2,3}, 3-4) the 4-qubit
    chain is [0, 1, 3, 4]."""
    if backend is not None and backend.num_qubits == 5 and n == 4:
        return [0, 1, 3, 4]
    return list(range(n))


def tfim_circuit(n, steps, J, h=1.0, dt=0.5, basis="Z", layout=None, num_physical=None, fold=1,
                 twirl_rng=None, random_init_prefix=False):
    layout = list(layout) if layout is not None else list(range(n))
    e = _Emitter(num_physical or (max(layout) + 1), fold, twirl_rng)
    L = lambda q: layout[q]
    if random_init_prefix:  # fixed 7-gate prefix of h13_ising_data_gen.ipynb:404 (4 qubits)
        e.rz(0.0007186381718527407, L(1)); e.rz(2.4917901988569855, L(1)); e.rz(3.3854853863523835, L(3))
        e.rx(1.2846113715328817, L(3)); e.noisy_cx(L(3), L(0)); e.rx(4.212671608894216, L(2)); e.noisy_cx(L(2), L(3))
    allq = list(range(n))
    for _ in range(steps):
        for q in allq:
            e.rx(2 * h * dt, L(q))
        for bonds, targets in ((allq[0::2], allq[1::2]), (allq[1:-2:2], allq[2:-1:2])):
            bonds = [q for q in bonds if q + 1 < n]
            for q0 in bonds:
                e.noisy_cx(L(q0), L(q0 + 1))
            for q in targets:
                e.rz(-2 * J * dt, L(q))
            for q0 in bonds:
                e.noisy_cx(L(q0), L(q0 + 1))
    if basis == "X":
        for q in allq:
            e.h(L(q))
    elif basis == "Y":
        for q in allq:
            e.rz(-PI / 2, L(q)); e.h(L(q))
    elif basis != "Z":
        raise ValueError("basis must be X, Y or Z")
    return e.done()


def brickwork_circuit(n, steps, rng, layout=None, num_physical=None, fold=1, twirl_rng=None):
    layout = list(layout) if layout is not None else list(range(n))
    e = _Emitter(num_physical or (max(layout) + 1), fold, twirl_rng)
    L = lambda q: layout[q]
    par = lambda k: 8 * PI * rng.random(k) - 4 * PI  # mbd_utils.gen_random_param
    for q in range(n):
        if q % 2 == 1:
            e.x(L(q))
    for _ in range(steps):
        for start in (0, 1):
            for q in range(start, n - 1, 2):
                e.h(L(q + 1)); e.noisy_cx(L(q), L(q + 1)); e.h(L(q + 1))  # cz
                e.u3(*par(3), L(q)); e.u3(*par(3), L(q + 1))
        for q in range(n):
            e.rz(float(par(1)[0]), L(q))
    return e.done()


def random_basis_circuit(n, depth, rng, coupling_map, num_physical=None):
    """Random layers over {rz, sx, x, cx} on a coupling map (stand-in for
    transpile(random_circuit(n, depth), backend), blackwater/data/generators/exp_val.py:116-120)."""
    b = BasisBuilder(num_physical or n)
    pairs = [tuple(p) for p in coupling_map]
    for _ in range(depth):
        r = rng.integers(0, 5)
        if r == 0:
            b.rz(float(rng.uniform(-PI, PI)), int(rng.integers(0, n)))
        elif r == 1:
            b.sx(int(rng.integers(0, n)))
        elif r == 2:
            b.x(int(rng.integers(0, n)))
        else:
            c, t = pairs[int(rng.integers(0, len(pairs)))]
            b.cx(c, t)
    return b.done()


def pad_label(chars_by_qubit, width):
    """{physical qubit: 'X'|'Y'|'Z'} -> Qiskit label of the given width (right-most = qubit 0)."""
    s = ["I"] * width
    for q, ch in chars_by_qubit.items():
        s[width - 1 - q] = ch
    return "".join(s)


def single_z_observables(layout, width):
    return [[(pad_label({q: "Z"}, width), 1.0)] for q in layout]


def tfim_observables(layout, width):
    """All single-Z, nearest-neighbour ZZ and XX, and Z^(x)n  (3n - 1 observables)."""
    obs = single_z_observables(layout, width)
    for a, b in zip(layout[:-1], layout[1:]):
        obs.append([(pad_label({a: "Z", b: "Z"}, width), 1.0)])
    for a, b in zip(layout[:-1], layout[1:]):
        obs.append([(pad_label({a: "X", b: "X"}, width), 1.0)])
    obs.append([(pad_label({q: "Z" for q in layout}, width), 1.0)])
    return obs


# ------------------------------------------------------------------------ BASELINE.json configs
def config_tfim4_lima_zne(n_base=2000, seed=0, factors=(1, 3, 5), backend=None):
    """cfg1: 4-qubit TFIM on the Lima chain [0,1,3,4]; steps i mod 15, J~U(0,1), basis~{X,Y,Z};
    every base circuit at ZNE fold factors (1,3,5); 4 single-Z observables.
    Returns (noisy_circuits [n_base*len(factors)], ideal_circuits [n_base], observables)."""
    rng = np.random.default_rng(seed)
    layout, width = [0, 1, 3, 4], 5
    noisy, ideal = [], []
    for i in range(n_base):
        J, basis, steps = float(rng.uniform(0, 1)), "XYZ"[int(rng.integers(0, 3))], i % 15
        for f in factors:
            noisy.append(tfim_circuit(4, steps, J, basis=basis, layout=layout, num_physical=width, fold=f,
                                      random_init_prefix=True))
        ideal.append(noisy[-len(factors)])
    return noisy, ideal, single_z_observables(layout, width)


def config_brick10_twirl(n_base=20, n_twirls=100, seed=1, n=10, width=16):
    """cfg2: 10-qubit random brickwork (steps 1..5) on a 16-qubit chain table, 100 Pauli twirls per
    base circuit, 10 single-Z observables.  Returns (twirled circuits, base circuits, observables)."""
    rng = np.random.default_rng(seed)
    layout = list(range(n))
    twirled, base = [], []
    for i in range(n_base):
        steps = 1 + i % 5
        cseed = int(rng.integers(0, 2 ** 31))
        base.append(brickwork_circuit(n, steps, np.random.default_rng(cseed), layout, width))
        for _ in range(n_twirls):
            twirled.append(brickwork_circuit(n, steps, np.random.default_rng(cseed), layout, width,
                                             twirl_rng=rng))
    return twirled, base, single_z_observables(layout, width)


def config_tfim_dm(n=14, n_circuits=8, seed=2, max_steps=10):
    """cfg3: n-qubit TFIM (steps 1..max_steps cyclic), 3n-1 observables (Z, ZZ, XX, Z^n)."""
    rng = np.random.default_rng(seed)
    layout = list(range(n))
    circs = [tfim_circuit(n, 1 + i % max_steps, float(rng.uniform(0, 1)), dt=0.25, layout=layout) for i in range(n_circuits)]
    return circs, tfim_observables(layout, n)


def config_mixed_dataset(n_circuits=5000, seed=5, n_min=6, n_max=12, width=12):
    """cfg5 (generation part): circuits of n ~ U{n_min..n_max} active qubits on a ``width``-qubit
    chain table, families TFIM / brickwork / random basis layers in equal parts (SURVEY.md 8(d)
    `e2e50k`), single-Z observables on the active qubits.  Returns (circuits, observables per circuit)."""
    rng = np.random.default_rng(seed)
    coupling = [(i, i + 1) for i in range(width - 1)] + [(i + 1, i) for i in range(width - 1)]
    circs, obs = [], []
    for i in range(n_circuits):
        n = int(rng.integers(n_min, n_max + 1))
        layout = list(range(n))
        fam = i % 3
        if fam == 0:
            c = tfim_circuit(n, 1 + int(rng.integers(0, 6)), float(rng.uniform(0, 1)), dt=0.25, layout=layout, num_physical=width)
        elif fam == 1:
            c = brickwork_circuit(n, 1 + int(rng.integers(0, 5)), np.random.default_rng(int(rng.integers(0, 2 ** 31))), layout, width)
        else:
            c = random_basis_circuit(n, int(rng.integers(8 * n, 20 * n)), rng, [p for p in coupling if p[0] < n and p[1] < n], width)
        circs.append(c)
        obs.append(single_z_observables(layout, width))
    return circs, obs
