"""Objects shaped like qiskit 1.3's circuit API, as far as the backends touch it (gates.canonical_window /
instruction_key): ``QuantumCircuit.data[i]`` materialises a FRESH ``CircuitInstruction`` on every access, whose
``.operation`` is a fresh object with ``.params`` a fresh list; qubits are ``Qubit`` objects resolved through
``circuit.qubits`` / ``circuit.find_bit``.  Wraps a harness ``Circuit`` (qiskit itself is not installable here).
Used by tests/test_fake_qiskit.py to show that the incremental translator diffs circuits by VALUE."""


class Qubit:
    __slots__ = ("_index",)

    def __init__(self, index):
        self._index = index

    def __repr__(self):
        return f"Qubit(q, {self._index})"


class _BitLocation:
    def __init__(self, index):
        self.index = index
        self.registers = []


class FreshOperation:
    def __init__(self, op):
        self.name = op.name
        self.label = getattr(op, "label", None)
        self.params = list(op.params)
        self._op = op

    def to_matrix(self):
        return self._op.to_matrix()


class FreshInstruction:
    def __init__(self, inst, qubits):
        self._inst = inst
        self._qubits = qubits

    @property
    def operation(self):
        return FreshOperation(self._inst.operation)       # a new object per access

    @property
    def qubits(self):
        return tuple(self._qubits[q] for q in self._inst.qubits)

    @property
    def clbits(self):
        return ()


class _FreshData:
    def __init__(self, data, qubits):
        self._data, self._qubits = data, qubits

    def __len__(self):
        return len(self._data)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [FreshInstruction(x, self._qubits) for x in self._data[i]]
        return FreshInstruction(self._data[i], self._qubits)

    def __iter__(self):
        return (FreshInstruction(x, self._qubits) for x in self._data)


class FreshCircuitView:
    """A live view of a harness Circuit that behaves like a qiskit QuantumCircuit on read."""

    def __init__(self, circuit):
        self._circuit = circuit
        self.num_qubits = circuit.num_qubits
        self.qubits = [Qubit(i) for i in range(circuit.num_qubits)]

    @property
    def data(self):
        return _FreshData(self._circuit.data, self.qubits)

    def find_bit(self, bit):
        return _BitLocation(bit._index)

    def __len__(self):
        return len(self._circuit.data)


class ViewCompiler:
    """Proxy of a harness compiler whose ``full_circuit`` is a FreshCircuitView (persistent object, so that the
    backend's per-compiler caches stay keyed on one identity)."""

    def __init__(self, compiler):
        object.__setattr__(self, "_compiler", compiler)
        object.__setattr__(self, "_view", None)

    @property
    def full_circuit(self):
        c = self._compiler.full_circuit
        if self._view is None or self._view._circuit is not c:
            object.__setattr__(self, "_view", FreshCircuitView(c))
        return self._view

    def __getattr__(self, name):
        return getattr(self._compiler, name)


class ViewBackend:
    """Hands the wrapped backend a ViewCompiler instead of the harness compiler."""

    def __init__(self, inner):
        self.inner = inner
        self.kind = inner.kind
        self.simulator = inner.simulator
        self._proxies = {}

    def _proxy(self, compiler):
        p = self._proxies.get(id(compiler))
        if p is None or p._compiler is not compiler:
            p = self._proxies[id(compiler)] = ViewCompiler(compiler)
        return p

    def evaluate_global_cost(self, compiler):
        return self.inner.evaluate_global_cost(self._proxy(compiler))

    def evaluate_local_cost(self, compiler):
        return self.inner.evaluate_local_cost(self._proxy(compiler))

    def evaluate_circuit(self, compiler):
        return self.inner.evaluate_circuit(self._proxy(compiler))

    def measure_qubit_expectation_values(self, compiler):
        return self.inner.measure_qubit_expectation_values(self._proxy(compiler))

    def shift_costs(self, compiler, gate_index, candidates):
        return self.inner.shift_costs(self._proxy(compiler), gate_index, candidates)
