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

    def is_complex(self):
        """True when some coefficient has a non-zero imaginary part (the expectation value is then
        complex, as Aer's ``np.real_if_close`` of the complex sum would return it)."""
        return any(c.imag != 0.0 for _, c in self.terms)

    def imag_part(self):
        """Observable with the imaginary parts of the coefficients: <O> = <Re O> + i <Im O>."""
        return PauliObservable([(l, c.imag) for l, c in self.terms])

    def masks(self):
        """(x_mask, z_mask, real coeff) arrays.  The C ABI takes real coefficients: a complex
        observable is evaluated as its real part here plus ``imag_part()`` (the estimator does
        that); calling this directly on a complex observable drops nothing silently -- it raises."""
        if self.is_complex() and not getattr(self, "_real_only_ok", False):
            raise ValueError("complex Pauli coefficients: evaluate real_part() and imag_part() separately")
        cached = self.__dict__.get("_masks")
        if cached is not None:
            return cached
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
        self.__dict__["_masks"] = (x, z, c)
        return x, z, c

    def real_part(self):
        return PauliObservable([(l, c.real) for l, c in self.terms])

    def __repr__(self):
        return f"PauliObservable({self.terms!r})"


_PAULI_PHASE = {"": 1.0, "+": 1.0, "-": -1.0, "i": 1j, "+i": 1j, "-i": -1j}


def _split_phase(label):
    """'-iXZ' -> ('XZ', -1j): a qiskit Pauli label with its group phase prefix."""
    body = label.lstrip("+-i")
    prefix = label[:len(label) - len(body)]
    if prefix not in _PAULI_PHASE:
        raise ValueError(f"bad Pauli phase prefix {prefix!r}")
    return body, _PAULI_PHASE[prefix]


def from_any(obj):
    if isinstance(obj, PauliObservable):
        return obj
    if isinstance(obj, str):
        body, phase = _split_phase(obj)
        return PauliObservable([(body, phase)])
    if hasattr(obj, "primitive"):  # opflow PauliSumOp
        coeff = complex(getattr(obj, "coeff", 1.0))
        inner = from_any(obj.primitive)
        return PauliObservable([(l, c * coeff) for l, c in inner.terms])
    if hasattr(obj, "paulis") and hasattr(obj, "coeffs"):  # SparsePauliOp
        terms = []
        for label, c in zip(obj.paulis.to_labels(), np.asarray(obj.coeffs)):
            body, phase = _split_phase(label)
            terms.append((body, complex(c) * phase))
        return PauliObservable(terms)
    if hasattr(obj, "to_label"):  # Pauli: its phase is the coefficient (as SparsePauliOp(Pauli) keeps it)
        body, phase = _split_phase(obj.to_label())
        return PauliObservable([(body, phase)])
    if isinstance(obj, (list, tuple)):
        if obj and isinstance(obj[0], str):
            return PauliObservable([(l, 1.0) for l in obj])
        return PauliObservable(list(obj))
    raise TypeError(f"cannot interpret {type(obj).__name__} as a Pauli observable")
