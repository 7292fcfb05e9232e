"""Minimal circuit IR + OpenQASM-2 reader + duck-typed Qiskit adapter.

The reference passes ``qiskit.QuantumCircuit`` objects to ``Estimator.run``
(blackwater/data/utils.py:422-443, blackwater/library/learning/estimator.py:279-285) and its
datasets store circuits as OpenQASM 2 text (field ``circuit`` of the JSON entries written by
blackwater/data/generators/exp_val.py:48-61).  Qiskit is not installed on the GPU box, so the
engine accepts (a) this IR, (b) QASM-2 text, (c) anything that quacks like a QuantumCircuit
(``num_qubits``, ``data`` of instructions with ``operation.name/params`` and ``qubits``).
"""
import ast
import math
import operator
import re

import numpy as np

from .gateset import IGNORED, NUM_PARAMS, OPCODES, canonical


class Parameter:
    """Free circuit parameter; supports ``a * theta + b`` (what Trotter / ansatz circuits need)."""

    __slots__ = ("name", "scale", "shift", "_base")

    def __init__(self, name, scale=1.0, shift=0.0, _base=None):
        self.name, self.scale, self.shift = name, float(scale), float(shift)
        self._base = _base if _base is not None else self

    def _derive(self, scale, shift):
        return Parameter(self.name, scale, shift, self._base)

    def __mul__(self, k):
        return self._derive(self.scale * float(k), self.shift * float(k))

    __rmul__ = __mul__

    def __truediv__(self, k):
        return self * (1.0 / float(k))

    def __add__(self, k):
        return self._derive(self.scale, self.shift + float(k))

    __radd__ = __add__

    def __sub__(self, k):
        return self + (-float(k))

    def __rsub__(self, k):
        return (-self) + float(k)

    def __neg__(self):
        return self * -1.0

    def value(self, x):
        return self.scale * x + self.shift

    def __repr__(self):
        return f"Parameter({self.name!r}, scale={self.scale}, shift={self.shift})"


class Circuit:
    """Ordered list of (name, qubits, params) on ``num_qubits`` wires (qubit 0 = LSB)."""

    def __init__(self, num_qubits, name=None):
        self.num_qubits = int(num_qubits)
        self.name = name
        self.ops = []
        self.metadata = {}

    # -- construction
    def append(self, name, qubits, params=()):
        n = canonical(name)
        qubits = tuple(int(q) for q in (qubits if isinstance(qubits, (tuple, list)) else (qubits,)))
        if n in IGNORED:
            return self
        if n == "measure":
            self.ops.append((n, qubits, ()))
            return self
        if n not in OPCODES:
            raise ValueError(f"unsupported instruction {name!r}")
        for q in qubits:
            if not 0 <= q < self.num_qubits:
                raise ValueError(f"qubit {q} out of range for {self.num_qubits}-qubit circuit")
        params = tuple(params)
        if len(params) != NUM_PARAMS.get(n, 0):
            raise ValueError(f"{n} takes {NUM_PARAMS.get(n, 0)} parameters, got {len(params)}")
        self.ops.append((n, qubits, params))
        return self

    def __getattr__(self, item):
        if item.startswith("_") or canonical(item) not in OPCODES:
            raise AttributeError(item)
        name = canonical(item)
        npar = NUM_PARAMS.get(name, 0)

        def add(*args):
            return self.append(name, args[npar:], args[:npar])

        return add

    def barrier(self, *_):
        return self

    def measure(self, qubit, clbit=None):
        self.ops.append(("measure", (int(qubit),), ()))
        return self

    def measure_all(self):
        for q in range(self.num_qubits):
            self.measure(q)
        return self

    def copy(self):
        c = Circuit(self.num_qubits, self.name)
        c.ops = list(self.ops)
        c.metadata = dict(self.metadata)
        return c

    def remove_final_measurements(self):
        c = self.copy()
        while c.ops and c.ops[-1][0] == "measure":
            c.ops.pop()
        return c

    # -- parameters (bound in name order, ParameterVector-style "v[10]" after "v[9]")
    @property
    def parameters(self):
        seen = {}
        for _, _, params in self.ops:
            for p in params:
                if isinstance(p, Parameter):
                    seen.setdefault(p.name, p._base)
        return [seen[k] for k in sorted(seen, key=_param_sort_key)]

    @property
    def num_parameters(self):
        ops = self.ops
        key = (len(ops), id(ops[-1]) if ops else 0)
        cache = self.__dict__.get("_npar")
        if cache is None or cache[0] != key:
            cache = (key, len(self.parameters))
            self.__dict__["_npar"] = cache
        return cache[1]

    def bind_parameters(self, values):
        names = [p.name for p in self.parameters]
        if isinstance(values, dict):
            table = {(k.name if isinstance(k, Parameter) else k): float(v) for k, v in values.items()}
        else:
            values = list(values)
            if len(values) != len(names):
                raise ValueError(f"circuit has {len(names)} parameters, got {len(values)} values")
            table = dict(zip(names, (float(v) for v in values)))
        c = Circuit(self.num_qubits, self.name)
        c.metadata = dict(self.metadata)
        for name, qubits, params in self.ops:
            c.ops.append((name, qubits, tuple(p.value(table[p.name]) if isinstance(p, Parameter) else p for p in params)))
        return c

    assign_parameters = bind_parameters

    def gate_ops(self):
        """ops with trailing measurements stripped; raises on mid-circuit measurement."""
        ops = list(self.ops)
        while ops and ops[-1][0] == "measure":
            ops.pop()
        if any(o[0] == "measure" for o in ops):
            raise ValueError("mid-circuit measurement is not supported by the exact estimator")
        return ops

    def flat(self):
        """The gate stream as numpy arrays (opcode u16, q0 u8, q1 u8, parameters per op i32,
        parameters f64), cached until the op list changes -- what encode_batch concatenates, so a
        circuit object is walked gate by gate in Python once, not once per run."""
        ops = self.ops
        key = (len(ops), id(ops[-1]) if ops else 0)
        cache = self.__dict__.get("_flat")
        if cache is not None and cache[0] == key:
            return cache[1]
        gate_ops = self.gate_ops()
        n = len(gate_ops)
        opc = np.empty(n, dtype=np.uint16)
        q0 = np.empty(n, dtype=np.uint8)
        q1 = np.zeros(n, dtype=np.uint8)
        npar = np.zeros(n, dtype=np.int32)
        params = []
        for i, (name, qubits, pr) in enumerate(gate_ops):
            opc[i] = OPCODES[name]
            q0[i] = qubits[0]
            if len(qubits) > 1:
                q1[i] = qubits[1]
            k = NUM_PARAMS.get(name, 0)
            if k:
                npar[i] = k
                params.extend(float(p) for p in pr)
        out = (opc, q0, q1, npar, np.asarray(params, dtype=np.float64))
        self.__dict__["_flat"] = (key, out)
        return out

    def flat_template(self):
        """``flat()`` of a PARAMETRISED circuit, once for all parameter sets: the arrays of ``flat()``
        with zeros at the parametric positions, plus (position, parameter index, scale, shift) of every
        ``a * theta + b`` entry -- a parameter set then binds with one numpy expression instead of a
        Python walk over the gates (VQE-style ``run(batch * [ansatz], batch * [op], parameter_values)``,
        reference: docs/tutorials/vqe_to_substitute*.py:260-269)."""
        ops = self.ops
        key = (len(ops), id(ops[-1]) if ops else 0)
        cache = self.__dict__.get("_flat_tpl")
        if cache is not None and cache[0] == key:
            return cache[1]
        names = [p.name for p in self.parameters]
        index = {n: i for i, n in enumerate(names)}
        gate_ops = self.gate_ops()
        n = len(gate_ops)
        opc = np.empty(n, dtype=np.uint16)
        q0 = np.empty(n, dtype=np.uint8)
        q1 = np.zeros(n, dtype=np.uint8)
        npar = np.zeros(n, dtype=np.int32)
        const, pos, idx, scale, shift = [], [], [], [], []
        for i, (name, qubits, pr) in enumerate(gate_ops):
            opc[i] = OPCODES[name]
            q0[i] = qubits[0]
            if len(qubits) > 1:
                q1[i] = qubits[1]
            k = NUM_PARAMS.get(name, 0)
            if k:
                npar[i] = k
                for p in pr:
                    if isinstance(p, Parameter):
                        pos.append(len(const)); idx.append(index[p.name]); scale.append(p.scale); shift.append(p.shift)
                        const.append(0.0)
                    else:
                        const.append(float(p))
        out = (opc, q0, q1, npar, np.asarray(const, dtype=np.float64), np.asarray(pos, dtype=np.int64),
               np.asarray(idx, dtype=np.int64), np.asarray(scale, dtype=np.float64), np.asarray(shift, dtype=np.float64), len(names))
        self.__dict__["_flat_tpl"] = (key, out)
        return out

    def bound_view(self, values):
        """This circuit with its parameters (name order, as ``bind_parameters``) set to ``values``, as
        a view that shares the template: ``flat()`` costs one numpy expression."""
        return _BoundCircuit(self, values)

    def size(self):
        return len(self.ops)

    def count_ops(self):
        out = {}
        for n, _, _ in self.ops:
            out[n] = out.get(n, 0) + 1
        return out

    @staticmethod
    def from_qasm(text):
        return parse_qasm(text)


class _BoundCircuit(Circuit):
    """``Circuit.bound_view``: behaves like ``template.bind_parameters(values)``; the gate list is only
    materialised if somebody asks for ``ops``-based views (``gate_ops``, ``size`` ...)."""

    def __init__(self, template, values):
        tpl = template.flat_template()
        values = np.asarray(values, dtype=np.float64).reshape(-1)
        if len(values) != tpl[9]:
            raise ValueError(f"circuit has {tpl[9]} parameters, got {len(values)} values")
        self.num_qubits = template.num_qubits
        self.name = template.name
        self.metadata = template.metadata
        self._template, self._values = template, values

    @property
    def ops(self):
        bound = self.__dict__.get("_bound")
        if bound is None:
            bound = self.__dict__["_bound"] = self._template.bind_parameters(list(self._values))
        return bound.ops

    def flat(self):
        opc, q0, q1, npar, const, pos, idx, scale, shift, _ = self._template.flat_template()
        params = const.copy()
        if len(pos):
            params[pos] = scale * self._values[idx] + shift
        return opc, q0, q1, npar, params

    @property
    def num_parameters(self):
        return 0


def _param_sort_key(name):
    m = re.match(r"^(.*)\[(\d+)\]$", name)
    return (m.group(1), int(m.group(2))) if m else (name, -1)


# ------------------------------------------------------------------------------ OpenQASM 2
_BIN = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
        ast.Pow: operator.pow}
_FUN = {"sin": math.sin, "cos": math.cos, "tan": math.tan, "exp": math.exp, "ln": math.log, "sqrt": math.sqrt,
        "asin": math.asin, "acos": math.acos, "atan": math.atan}


def _eval_expr(expr, env=None):
    node = ast.parse(expr.replace("^", "**").strip(), mode="eval").body

    def ev(n):
        if isinstance(n, ast.Constant) and isinstance(n.value, (int, float)):
            return float(n.value)
        if isinstance(n, ast.Name):
            if n.id == "pi":
                return math.pi
            if env and n.id in env:
                return env[n.id]
            raise ValueError(f"unknown identifier {n.id!r} in QASM expression")
        if isinstance(n, ast.BinOp) and type(n.op) in _BIN:
            return _BIN[type(n.op)](ev(n.left), ev(n.right))
        if isinstance(n, ast.UnaryOp) and isinstance(n.op, (ast.USub, ast.UAdd)):
            v = ev(n.operand)
            return -v if isinstance(n.op, ast.USub) else v
        if isinstance(n, ast.Call) and isinstance(n.func, ast.Name) and n.func.id in _FUN and len(n.args) == 1:
            return _FUN[n.func.id](ev(n.args[0]))
        raise ValueError(f"unsupported QASM expression {expr!r}")

    return ev(node)


def _split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [a.strip() for a in out]


_STMT = re.compile(r"^([A-Za-z_][A-Za-z0-9_]*)\s*(?:\((.*)\))?\s*(.*)$", re.S)


def parse_qasm(text):
    text = re.sub(r"//[^\n]*", "", text)
    # user gate definitions: gate name(params) qargs { body }
    gate_defs = {}

    def grab_def(m):
        name, params, qargs, body = m.group(1), m.group(2), m.group(3), m.group(4)
        gate_defs[name] = ([p.strip() for p in params.split(",")] if params and params.strip() else [],
                           [q.strip() for q in qargs.split(",")], [s.strip() for s in body.split(";") if s.strip()])
        return ""

    text = re.sub(r"\bgate\s+([A-Za-z_][A-Za-z0-9_]*)\s*(?:\(([^)]*)\))?\s*([^{]*)\{([^}]*)\}", grab_def, text)
    stmts = [s.strip() for s in text.split(";") if s.strip()]
    qregs, offset = {}, 0
    body = []
    for s in stmts:
        if s.startswith("OPENQASM") or s.startswith("include") or s.startswith("creg"):
            continue
        m = re.match(r"^qreg\s+([A-Za-z_][A-Za-z0-9_]*)\s*\[(\d+)\]$", s)
        if m:
            qregs[m.group(1)] = (offset, int(m.group(2)))
            offset += int(m.group(2))
            continue
        body.append(s)
    circ = Circuit(offset)

    def resolve(arg, qenv):
        arg = arg.strip()
        if qenv is not None and arg in qenv:
            return [qenv[arg]]
        m = re.match(r"^([A-Za-z_][A-Za-z0-9_]*)\s*\[(\d+)\]$", arg)
        if m:
            off, size = qregs[m.group(1)]
            i = int(m.group(2))
            if i >= size:
                raise ValueError(f"QASM: index out of range in {arg!r}")
            return [off + i]
        if arg in qregs:
            off, size = qregs[arg]
            return list(range(off, off + size))
        raise ValueError(f"QASM: cannot resolve qubit argument {arg!r}")

    def emit(stmt, penv, qenv, depth=0):
        if depth > 32:
            raise ValueError("QASM: gate definitions nested too deeply")
        if stmt.startswith("measure"):
            src = stmt[len("measure"):].split("->")[0]
            for q in resolve(src, qenv):
                circ.ops.append(("measure", (q,), ()))
            return
        if stmt.startswith("if"):
            raise ValueError("QASM: classically conditioned operations are not supported")
        m = _STMT.match(stmt)
        if not m:
            raise ValueError(f"QASM: cannot parse statement {stmt!r}")
        name, pstr, qstr = m.group(1), m.group(2), m.group(3)
        params = [_eval_expr(p, penv) for p in _split_args(pstr)] if pstr else []
        qlists = [resolve(a, qenv) for a in _split_args(qstr)] if qstr.strip() else []
        if canonical(name) in IGNORED:
            return
        width = max((len(q) for q in qlists), default=1)
        for k in range(width):  # register broadcast
            qs = [q[k] if len(q) > 1 else q[0] for q in qlists]
            if name in gate_defs and canonical(name) not in OPCODES:
                pnames, qnames, gbody = gate_defs[name]
                for sub in gbody:
                    emit(sub, dict(zip(pnames, params)), dict(zip(qnames, qs)), depth + 1)
            elif canonical(name) == "u0":
                continue
            else:
                circ.append(name, qs, params)

    for s in body:
        emit(s, None, None)
    return circ


# ------------------------------------------------------------------------------ adapters
def from_any(obj):
    """Circuit | QASM text | Qiskit-like QuantumCircuit  ->  Circuit."""
    if isinstance(obj, Circuit):
        return obj
    if isinstance(obj, str):
        return parse_qasm(obj)
    if hasattr(obj, "data") and hasattr(obj, "num_qubits"):
        circ = Circuit(obj.num_qubits, getattr(obj, "name", None))
        index = {}
        for i, q in enumerate(getattr(obj, "qubits", [])):
            index[id(q)] = i
        for inst in obj.data:
            op = getattr(inst, "operation", None)
            qargs = getattr(inst, "qubits", None)
            if op is None:  # legacy (instruction, qargs, cargs) tuples
                op, qargs = inst[0], inst[1]
            qs = [index[id(q)] if id(q) in index else obj.find_bit(q).index for q in qargs]
            name = op.name
            if getattr(op, "condition", None) is not None:
                raise ValueError("classically conditioned operations are not supported")
            if canonical(name) in IGNORED:
                continue
            if canonical(name) == "measure":
                circ.ops.append(("measure", (qs[0],), ()))
                continue
            if canonical(name) == "unitary":  # UnitaryGate: its parameter is the matrix, not an angle
                mat = np.asarray(op.to_matrix() if hasattr(op, "to_matrix") else op.params[0], dtype=complex)
                if mat.shape not in ((2, 2), (4, 4)):
                    raise ValueError(f"unitary on {len(qs)} qubits: only 1- and 2-qubit unitaries are supported")
                flat = np.stack([mat.real, mat.imag], -1).reshape(-1)
                circ.ops.append(("unitary1" if mat.shape[0] == 2 else "unitary2", tuple(qs), tuple(flat)))
                continue
            params = []
            for p in op.params:
                try:
                    params.append(float(p))
                except TypeError as exc:
                    raise ValueError("unbound parameters: bind them (parameter_values) before lowering") from exc
            circ.append(name, qs, params)
        return circ
    raise TypeError(f"cannot interpret {type(obj).__name__} as a circuit")
