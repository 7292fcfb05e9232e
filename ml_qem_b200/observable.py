"""Pauli observables: list of (label, coeff); label is a Qiskit string (right-most = qubit 0).

Accepted inputs mirror what the reference hands to ``Estimator.run``: ``SparsePauliOp`` /
``PauliSumOp`` (blackwater/data/utils.py:477-491 generate_random_pauli_sum_op), plain labels,
and lists of (label, coeff) -- see [3P] BaseEstimator.run normalisation.
"""
import numpy as np


class PauliObservable:
    def __init__(self, terms):
        self.terms = [(str(l), complex(c)) for l, c in terms]
        widths = {len(l) for l, _ in self.terms}
        if len(widths) > 1:
            raise ValueError("all Pauli labels of an observable must have the same width")
        for l, _ in self.terms:
            if set(l) - set("IXYZ"):
                raise ValueError(f"bad Pauli label {l!r}")
        self.num_qubits = widths.pop() if widths else 0

    def __len__(self):
        return len(self.terms)

    def __iter__(self):
        return iter(self.terms)

    def masks(self):
        """(x_mask, z_mask, real coeff) arrays; complex coefficients keep their real part."""
        n = len(self.terms)
        x = np.zeros(n, dtype=np.uint64)
        z = np.zeros(n, dtype=np.uint64)
        c = np.zeros(n, dtype=np.float64)
        for k, (label, coeff) in enumerate(self.terms):
            xm = zm = 0
            w = len(label)
            for q in range(w):
                ch = label[w - 1 - q]
                if ch in "XY":
                    xm |= 1 << q
                if ch in "ZY":
                    zm |= 1 << q
            x[k], z[k], c[k] = xm, zm, coeff.real
        return x, z, c

    def __repr__(self):
        return f"PauliObservable({self.terms!r})"


def from_any(obj):
    if isinstance(obj, PauliObservable):
        return obj
    if isinstance(obj, str):
        return PauliObservable([(obj, 1.0)])
    if hasattr(obj, "primitive"):  # opflow PauliSumOp
        coeff = complex(getattr(obj, "coeff", 1.0))
        inner = from_any(obj.primitive)
        return PauliObservable([(l, c * coeff) for l, c in inner.terms])
    if hasattr(obj, "paulis") and hasattr(obj, "coeffs"):  # SparsePauliOp
        return PauliObservable(list(zip(obj.paulis.to_labels(), np.asarray(obj.coeffs))))
    if hasattr(obj, "to_label"):  # Pauli
        return PauliObservable([(obj.to_label().lstrip("+-i"), 1.0)])
    if isinstance(obj, (list, tuple)):
        if obj and isinstance(obj[0], str):
            return PauliObservable([(l, 1.0) for l in obj])
        return PauliObservable(list(obj))
    raise TypeError(f"cannot interpret {type(obj).__name__} as a Pauli observable")
