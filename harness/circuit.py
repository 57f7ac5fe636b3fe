"""Minimal QuantumCircuit-shaped container.

qiskit is not installable in the build image, so the host-side mirror of the ADAPT-AQC loop
(``adapt_aqc_b200.compiler``) and the tests drive the backends with this class.  It exposes
exactly the surface the reference touches on ``compiler.full_circuit``
(adaptaqc/utils/circuit_operations/circuit_operations_basic.py:50-99): ``.data`` (a mutable
list of instructions with ``.operation`` / ``.qubits`` / ``.clbits``), ``.qubits``,
``.num_qubits``, ``.copy()``; gates carry ``.name`` / ``.label`` / ``.params`` /
``.to_matrix()``.  The backends accept a real ``qiskit.QuantumCircuit`` the same way.
"""
import copy as _copy

import numpy as np

from adapt_aqc_b200 import gates as G

_SELF_INVERSE = {"x", "y", "z", "h", "cx", "cz", "swap", "id"}
_INVERSE_NAME = {"s": "sdg", "sdg": "s", "t": "tdg", "tdg": "t"}


class Gate:
    __slots__ = ("name", "params", "label", "num_qubits", "_matrix")

    def __init__(self, name, params=(), label=None, num_qubits=None, matrix=None):
        self.name = name
        self.params = list(params)
        self.label = label
        if num_qubits is None:
            num_qubits = G.GATE_TABLE[name][1] if name in G.GATE_TABLE else 1
        self.num_qubits = num_qubits
        self._matrix = None if matrix is None else np.asarray(matrix, dtype=np.complex128)

    def copy(self):
        return Gate(self.name, list(self.params), self.label, self.num_qubits, self._matrix)

    def to_mutable(self):
        return self.copy()

    def to_matrix(self):
        if self._matrix is not None:
            return self._matrix
        if self.num_qubits == 1:
            p = list(self.params) + [0.0, 0.0, 0.0]
            return G.matrix_of_entry((self.name, 0, -1, float(p[0]), float(p[1]), float(p[2]), None))
        if self.name == "cx":  # control = first qubit = least significant index bit
            return np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.complex128)
        if self.name == "cz":
            return np.diag([1, 1, 1, -1]).astype(np.complex128)
        if self.name == "swap":
            return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
        raise ValueError(f"no matrix for gate '{self.name}'")

    def inverse(self):
        n = self.name
        if self._matrix is not None:
            return Gate(n, [], self.label, self.num_qubits, self._matrix.conj().T)
        if n in _SELF_INVERSE:
            return self.copy()
        if n in _INVERSE_NAME:
            return Gate(_INVERSE_NAME[n], [], None, 1)
        if n in ("rx", "ry", "rz", "u1", "p"):
            return Gate(n, [-self.params[0]], self.label, 1)
        if n in ("u3", "u"):
            t, ph, lm = self.params
            return Gate(n, [-t, -lm, -ph], None, 1)
        if n == "u2":
            ph, lm = self.params
            return Gate("u3", [-np.pi / 2, -lm, -ph], None, 1)
        if n == "sx":
            return Gate("mat1", [], None, 1, self.to_matrix().conj().T)
        raise ValueError(f"cannot invert gate '{n}'")

    def __eq__(self, other):
        return (
            isinstance(other, Gate)
            and self.name == other.name
            and self.params == other.params
            and self.label == other.label
        )

    def __repr__(self):
        return f"Gate({self.name}, {self.params}, label={self.label})"


class MPSInstruction:
    """``set_matrix_product_state`` / ``save_matrix_product_state`` stand-in (not a Gate)."""

    __slots__ = ("name", "params", "label", "num_qubits")

    def __init__(self, name, mps, num_qubits):
        self.name = name
        self.params = [mps]
        self.label = None
        self.num_qubits = num_qubits

    def copy(self):
        return MPSInstruction(self.name, self.params[0], self.num_qubits)


class CircuitInstruction:
    __slots__ = ("operation", "qubits", "clbits")

    def __init__(self, operation, qubits, clbits=()):
        self.operation = operation
        self.qubits = tuple(qubits)
        self.clbits = tuple(clbits)

    def __iter__(self):  # legacy (op, qargs, cargs) unpacking
        return iter((self.operation, self.qubits, self.clbits))

    def __getitem__(self, i):
        return (self.operation, self.qubits, self.clbits)[i]


class Circuit:
    def __init__(self, num_qubits):
        self.num_qubits = int(num_qubits)
        self.qubits = list(range(self.num_qubits))
        self.clbits = []
        self.data = []

    # ---- container protocol ----
    def __len__(self):
        return len(self.data)

    def __iter__(self):
        return iter(self.data)

    def __eq__(self, other):
        return (
            isinstance(other, Circuit)
            and self.num_qubits == other.num_qubits
            and len(self.data) == len(other.data)
            and all(
                a.operation == b.operation and a.qubits == b.qubits for a, b in zip(self.data, other.data)
            )
        )

    def copy(self):
        c = Circuit(self.num_qubits)
        c.data = [CircuitInstruction(i.operation.copy(), i.qubits, i.clbits) for i in self.data]
        return c

    def append(self, operation, qubits):
        self.data.append(CircuitInstruction(operation, [int(q) for q in qubits]))
        return self

    def compose(self, other, qubits=None):
        out = self.copy()
        qubits = list(range(other.num_qubits)) if qubits is None else list(qubits)
        for inst in other.data:
            out.append(inst.operation.copy(), [qubits[q] for q in inst.qubits])
        return out

    def inverse(self):
        c = Circuit(self.num_qubits)
        for inst in reversed(self.data):
            c.data.append(CircuitInstruction(inst.operation.inverse(), inst.qubits))
        return c

    def count_ops(self):
        d = {}
        for inst in self.data:
            d[inst.operation.name] = d.get(inst.operation.name, 0) + 1
        return d

    def depth(self, filter_function=None):
        level = [0] * self.num_qubits
        for inst in self.data:
            if filter_function is not None and not filter_function(inst):
                continue
            d = 1 + max(level[q] for q in inst.qubits)
            for q in inst.qubits:
                level[q] = d
        return max(level) if level else 0

    # ---- builders (qiskit names) ----
    def _each(self, name, params, qubit, label=None):
        qs = qubit if isinstance(qubit, (list, tuple, range)) else [qubit]
        for q in qs:
            self.append(Gate(name, params, label), [q])
        return self

    def id(self, q): return self._each("id", [], q)
    def x(self, q): return self._each("x", [], q)
    def y(self, q): return self._each("y", [], q)
    def z(self, q): return self._each("z", [], q)
    def h(self, q): return self._each("h", [], q)
    def s(self, q): return self._each("s", [], q)
    def sdg(self, q): return self._each("sdg", [], q)
    def t(self, q): return self._each("t", [], q)
    def tdg(self, q): return self._each("tdg", [], q)
    def sx(self, q): return self._each("sx", [], q)
    def rx(self, theta, q, label=None): return self._each("rx", [theta], q, label)
    def ry(self, theta, q, label=None): return self._each("ry", [theta], q, label)
    def rz(self, theta, q, label=None): return self._each("rz", [theta], q, label)
    def p(self, lam, q): return self._each("p", [lam], q)
    def u1(self, lam, q): return self._each("u1", [lam], q)
    def u2(self, phi, lam, q): return self._each("u2", [phi, lam], q)
    def u3(self, theta, phi, lam, q): return self._each("u3", [theta, phi, lam], q)
    def u(self, theta, phi, lam, q): return self._each("u3", [theta, phi, lam], q)
    def cx(self, c, t): return self.append(Gate("cx"), [c, t])
    def cz(self, a, b): return self.append(Gate("cz"), [a, b])
    def swap(self, a, b): return self.append(Gate("swap"), [a, b])

    def ccx(self, a, b, c):
        """Toffoli, expanded to the standard 6-CX network (what unroll_to_basis_gates leaves)."""
        self.h(c); self.cx(b, c); self.tdg(c); self.cx(a, c); self.t(c); self.cx(b, c)
        self.tdg(c); self.cx(a, c); self.t(b); self.t(c); self.h(c); self.cx(a, b)
        self.t(a); self.tdg(b); self.cx(a, b)
        return self

    def unitary(self, matrix, qubits, label=None):
        qubits = list(qubits) if isinstance(qubits, (list, tuple)) else [qubits]
        m = np.asarray(matrix, dtype=np.complex128)
        name = "mat1" if len(qubits) == 1 else "mat2"
        return self.append(Gate(name, [], label, len(qubits), m), qubits)

    def set_matrix_product_state(self, mps):
        return self.append(MPSInstruction("set_matrix_product_state", mps, self.num_qubits), self.qubits)

    def save_matrix_product_state(self):
        return self.append(MPSInstruction("save_matrix_product_state", None, self.num_qubits), self.qubits)

    def to_gate_list(self):
        return [(i.operation.name, i.qubits, list(i.operation.params)) for i in self.data]


def deep_copy_instruction(inst):
    return CircuitInstruction(_copy.copy(inst.operation), inst.qubits, inst.clbits)
