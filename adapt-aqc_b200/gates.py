"""Gate-stream IR: the wire format of ``b200_gate`` (include/b200aqc.h) and the translator from
a QuantumCircuit-shaped object.

The translator is duck-typed on the handful of attributes the reference itself relies on
(``circuit.data[i].operation.name / .params / .to_matrix()``, ``circuit.data[i].qubits``,
``circuit.qubits``; see adaptaqc/utils/circuit_operations/circuit_operations_basic.py:70-132),
so it accepts a real ``qiskit.QuantumCircuit`` as well as ``adapt_aqc_b200.circuit.Circuit``.
"""
import numpy as np

GATE_DTYPE = np.dtype(
    [("op", "<i4"), ("q0", "<i4"), ("q1", "<i4"), ("aux", "<i4"), ("p", "<f8", (3,))]
)
assert GATE_DTYPE.itemsize == 40

OP_ID, OP_X, OP_Y, OP_Z, OP_H, OP_RX, OP_RY, OP_RZ, OP_U1, OP_U2, OP_U3 = range(11)
OP_CX, OP_CZ, OP_MAT1, OP_MAT2, OP_S, OP_SDG, OP_T, OP_TDG, OP_SX, OP_SWAP = range(11, 21)

# name -> (opcode, number of qubits, number of parameters)
GATE_TABLE = {
    "id": (OP_ID, 1, 0), "i": (OP_ID, 1, 0),
    "x": (OP_X, 1, 0), "y": (OP_Y, 1, 0), "z": (OP_Z, 1, 0), "h": (OP_H, 1, 0),
    "s": (OP_S, 1, 0), "sdg": (OP_SDG, 1, 0), "t": (OP_T, 1, 0), "tdg": (OP_TDG, 1, 0),
    "sx": (OP_SX, 1, 0),
    "rx": (OP_RX, 1, 1), "ry": (OP_RY, 1, 1), "rz": (OP_RZ, 1, 1),
    "u1": (OP_U1, 1, 1), "p": (OP_U1, 1, 1),
    "u2": (OP_U2, 1, 2), "u3": (OP_U3, 1, 3), "u": (OP_U3, 1, 3),
    "cx": (OP_CX, 2, 0), "cz": (OP_CZ, 2, 0), "swap": (OP_SWAP, 2, 0),
}

# instructions that carry no unitary action on the simulated state
IGNORED = {"barrier", "delay"}

ROTATIONS = ("rx", "ry", "rz")


class UnsupportedInstruction(ValueError):
    pass


def qubit_indices(circuit):
    """{qubit object: index}.  For Circuit (plain ints) the identity map."""
    qubits = circuit.qubits
    if len(qubits) and isinstance(qubits[0], (int, np.integer)):
        return None
    return {q: i for i, q in enumerate(qubits)}


def canonical_window(circuit, start=0, stop=None, qmap="auto"):
    """Translate circuit.data[start:stop] into a list of hashable tuples
    ``(name, q0, q1, p0, p1, p2, matrix_bytes_or_None)`` -- cheap to diff between evaluations."""
    if qmap == "auto":
        qmap = qubit_indices(circuit)
    data = circuit.data
    stop = len(data) if stop is None else stop
    out = []
    for i in range(start, stop):
        inst = data[i]
        op = inst.operation
        name = op.name
        if name in IGNORED:
            continue
        qs = inst.qubits
        if qmap is not None:
            qs = [qmap[q] for q in qs]
        ent = GATE_TABLE.get(name)
        if ent is not None:
            params = op.params
            npar = ent[2]
            p0 = float(params[0]) if npar > 0 else 0.0
            p1 = float(params[1]) if npar > 1 else 0.0
            p2 = float(params[2]) if npar > 2 else 0.0
            q1 = int(qs[1]) if ent[1] == 2 else -1
            out.append((name, int(qs[0]), q1, p0, p1, p2, None))
        else:
            if len(getattr(inst, "clbits", ())):
                raise UnsupportedInstruction(f"classical instruction '{name}' cannot be simulated")
            if len(qs) not in (1, 2) or not hasattr(op, "to_matrix"):
                raise UnsupportedInstruction(f"instruction '{name}' on {len(qs)} qubits is not supported")
            m = np.ascontiguousarray(np.asarray(op.to_matrix(), dtype=np.complex128))
            q1 = int(qs[1]) if len(qs) == 2 else -1
            out.append(("mat2" if len(qs) == 2 else "mat1", int(qs[0]), q1, 0.0, 0.0, 0.0, m.tobytes()))
    return out


def instruction_key(inst, qmap=None):
    """Value identity of one circuit instruction: (name, params, qubit indices), or None for instructions whose
    parameters are not plain numbers (e.g. a `unitary` carrying a matrix) -- those only match by object identity.

    qiskit >= 1.0 materialises a fresh ``CircuitInstruction`` on every ``QuantumCircuit.data[i]`` access, so the
    incremental translator diffs circuits on these values, never on ``id()`` alone."""
    op = inst.operation
    name = op.name
    if name not in GATE_TABLE and name not in IGNORED:
        return None
    qs = inst.qubits
    if qmap is not None:
        qs = [qmap[q] for q in qs]
    try:
        return (name, tuple(float(p) for p in op.params), tuple(int(q) for q in qs))
    except (TypeError, ValueError):        # unbound Parameter etc.
        return None


def window_index(circuit, start, gate_index):
    """Position in ``canonical_window(circuit, start, None)`` of the instruction at circuit index `gate_index`
    (the canonical window drops barriers / delays, so it is not always gate_index - start)."""
    data = circuit.data
    if not start <= gate_index < len(data):
        raise IndexError(f"gate index {gate_index} outside the window [{start}, {len(data)})")
    if data[gate_index].operation.name in IGNORED:
        raise ValueError(f"instruction {gate_index} ({data[gate_index].operation.name}) carries no gate")
    return sum(1 for i in range(start, gate_index) if data[i].operation.name not in IGNORED)


class GateStream:
    """Packed gate records + dense-matrix pool, ready for the C-ABI."""

    __slots__ = ("rec", "mats", "window")

    def __init__(self, rec=None, mats=None, window=None):
        self.rec = rec if rec is not None else np.zeros(0, dtype=GATE_DTYPE)
        self.mats = mats if mats is not None else np.zeros(0, dtype=np.float64)
        self.window = window          # the canonical entries the records were packed from (from_window)

    def __len__(self):
        return len(self.rec)

    @classmethod
    def from_window(cls, window):
        """window: list of canonical tuples (see canonical_window)."""
        n = len(window)
        rec = np.zeros(n, dtype=GATE_DTYPE)
        mats = []
        off = 0
        ops = rec["op"]; q0s = rec["q0"]; q1s = rec["q1"]; aux = rec["aux"]; ps = rec["p"]
        for k, (name, q0, q1, p0, p1, p2, mb) in enumerate(window):
            if mb is None:
                ops[k] = GATE_TABLE[name][0]
            else:
                ops[k] = OP_MAT2 if name == "mat2" else OP_MAT1
                m = np.frombuffer(mb, dtype=np.float64)
                aux[k] = off
                mats.append(m)
                off += m.size
            q0s[k] = q0
            q1s[k] = q1
            ps[k, 0] = p0; ps[k, 1] = p1; ps[k, 2] = p2
        return cls(rec, np.concatenate(mats) if mats else None, window)

    @classmethod
    def from_gates(cls, gates):
        """gates: [(name, qubits, params)] with params = angles, or the matrix for mat1/mat2."""
        window = []
        for name, qubits, params in gates:
            q0 = int(qubits[0])
            q1 = int(qubits[1]) if len(qubits) > 1 else -1
            if name in ("mat1", "mat2"):
                m = np.ascontiguousarray(np.asarray(params, dtype=np.complex128))
                window.append((name, q0, q1, 0.0, 0.0, 0.0, m.tobytes()))
            else:
                p = [float(x) for x in params] + [0.0, 0.0, 0.0]
                window.append((name, q0, q1, p[0], p[1], p[2], None))
        return cls.from_window(window)

    @classmethod
    def from_circuit(cls, circuit, start=0, stop=None):
        return cls.from_window(canonical_window(circuit, start, stop))

    def rec_ptr(self):
        return self.rec.ctypes.data if len(self.rec) else None

    def mats_ptr(self):
        return self.mats.ctypes.data if len(self.mats) else None


_SELF_INVERSE = ("id", "i", "x", "y", "z", "h", "cx", "cz", "swap")
_INVERSE_NAME = {"s": "sdg", "sdg": "s", "t": "tdg", "tdg": "t"}


def invert_window(window):
    """Canonical entries of the inverse circuit (reversed order, every gate inverted)."""
    out = []
    for ent in reversed(window):
        name, q0, q1, p0, p1, p2, mb = ent
        if mb is not None:
            d = 4 if q1 >= 0 else 2
            m = np.frombuffer(mb, dtype=np.complex128).reshape(d, d)
            out.append((name, q0, q1, 0.0, 0.0, 0.0, np.ascontiguousarray(m.conj().T).tobytes()))
        elif name in _SELF_INVERSE:
            out.append(ent)
        elif name in ROTATIONS or name in ("u1", "p"):
            out.append((name, q0, q1, -p0, p1, p2, None))
        elif name in _INVERSE_NAME:
            out.append((_INVERSE_NAME[name], q0, q1, p0, p1, p2, None))
        else:       # u2 / u3 / sx ...: as a dense matrix
            m = np.asarray(matrix_of_entry(ent), dtype=np.complex128)
            out.append(("mat1", q0, -1, 0.0, 0.0, 0.0, np.ascontiguousarray(m.conj().T).tobytes()))
    return out


_ROT_CACHE = {}


def one_qubit_matrix(name, theta):
    """2x2 matrix of rx/ry/rz(theta) (standard qiskit definitions).  Cached (read-only arrays): the optimiser
    asks for the same few shift angles of every gate (cost_minimiser.py:344-368)."""
    key = (name, float(theta))
    m = _ROT_CACHE.get(key)
    if m is not None:
        return m
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    if name == "rx":
        m = np.array([[c, -1j * s], [-1j * s, c]])
    elif name == "ry":
        m = np.array([[c, -s], [s, c]], dtype=np.complex128)
    elif name == "rz":
        m = np.array([[np.exp(-0.5j * theta), 0], [0, np.exp(0.5j * theta)]])
    else:
        raise ValueError(f"not a rotation: {name}")
    m.flags.writeable = False
    if len(_ROT_CACHE) > 8192:
        _ROT_CACHE.clear()
    _ROT_CACHE[key] = m
    return m


_CTUPLE_CACHE = {}


def complex_tuple(m):
    """(m00, m01, m10, m11) as Python complex of a 2x2 array; cached for the read-only rotation matrices of
    one_qubit_matrix (the optimiser's shift candidates)."""
    if not m.flags.writeable:
        t = _CTUPLE_CACHE.get(id(m))
        if t is not None and t[0] is m:
            return t[1]
        v = tuple(complex(x) for x in m.ravel())
        if len(_CTUPLE_CACHE) > 16384:
            _CTUPLE_CACHE.clear()
        _CTUPLE_CACHE[id(m)] = (m, v)
        return v
    return tuple(complex(x) for x in np.asarray(m).ravel())


def matrix_of_entry_complex(ent):
    """complex_tuple of a canonical 1-qubit window entry."""
    return complex_tuple(matrix_of_entry(ent))


_R2 = 1 / np.sqrt(2)


def matrix_of_entry(ent):
    """2x2 matrix of a canonical 1-qubit window entry."""
    name, _, _, p0, p1, p2, mb = ent
    if mb is not None:
        return np.frombuffer(mb, dtype=np.complex128).reshape(2, 2)
    if name in ROTATIONS:
        return one_qubit_matrix(name, p0)
    if name in ("id", "i"):
        return np.eye(2, dtype=np.complex128)
    if name == "x":
        return np.array([[0, 1], [1, 0]], dtype=np.complex128)
    if name == "y":
        return np.array([[0, -1j], [1j, 0]])
    if name == "z":
        return np.diag([1, -1]).astype(np.complex128)
    if name == "h":
        return np.array([[_R2, _R2], [_R2, -_R2]], dtype=np.complex128)
    if name == "s":
        return np.diag([1, 1j])
    if name == "sdg":
        return np.diag([1, -1j])
    if name == "t":
        return np.diag([1, np.exp(0.25j * np.pi)])
    if name == "tdg":
        return np.diag([1, np.exp(-0.25j * np.pi)])
    if name == "sx":
        return 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]])
    if name in ("u1", "p"):
        return np.diag([1, np.exp(1j * p0)])
    if name == "u2":
        return _R2 * np.array([[1, -np.exp(1j * p1)], [np.exp(1j * p0), np.exp(1j * (p0 + p1))]])
    if name in ("u3", "u"):
        c, s = np.cos(p0 / 2), np.sin(p0 / 2)
        return np.array([[c, -np.exp(1j * p2) * s], [np.exp(1j * p1) * s, np.exp(1j * (p1 + p2)) * c]])
    raise ValueError(f"not a 1-qubit gate: {name}")
