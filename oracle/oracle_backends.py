"""CPU ORACLE -- backend objects (test infrastructure, NOT product code).

``OracleSVBackend`` restates ``AerSVBackend`` (adaptaqc/backends/aer_sv_backend.py:19-59) line by
line on top of the C statevector oracle, *including the reference's redundancy*: every
evaluation re-simulates all gates of ``compiler.full_circuit`` from |0..0>, <Z_i> takes one
``probabilities([i])`` pass per qubit, and the pair heuristic re-runs the whole circuit once per
candidate pair through ``backend.simulator.run`` followed by a host partial trace
(adaptaqc/utils/circuit_operations/circuit_operations_running.py:58-63,
adaptaqc/utils/entanglement_measures.py:71-75, 325-340).  It is the checker for the parity
tests and the ``cpu_baseline`` / ``--impl reference`` arm of bench.py.  Only tests/,
__graft_entry__.smoke() and bench.py import it.

Parity status: pinned on the reference's known-answer tests (tests/test_oracle_kats.py); the
third-party simulators the reference calls (qiskit-aer 0.16, qiskit 1.3) are not installable
offline, so no output of the reference itself is available here.
"""
import numpy as np

from . import sv_oracle as orc


def circuit_to_gates(circuit):
    """[(name, qubits, params)] from a QuantumCircuit-shaped object (duck-typed)."""
    qubits = circuit.qubits
    qmap = None
    if len(qubits) and not isinstance(qubits[0], (int, np.integer)):
        qmap = {q: i for i, q in enumerate(qubits)}
    out = []
    for inst in circuit.data:
        op = inst.operation
        name = op.name
        if name in ("barrier", "delay"):
            continue
        qs = [int(q) if qmap is None else qmap[q] for q in inst.qubits]
        if name in orc.OPCODES and name not in ("mat1", "mat2"):
            out.append((name, qs, [float(p) for p in op.params]))
        else:
            out.append(("mat2" if len(qs) == 2 else "mat1", qs, np.asarray(op.to_matrix(), dtype=np.complex128)))
    return out


class OracleStatevector:
    """What Aer's ``result.get_statevector()`` offers to the reference (aer_sv_backend.py:29,52,56)."""

    def __init__(self, data):
        self.data = np.ascontiguousarray(data, dtype=np.complex128)
        self.num_qubits = int(np.log2(self.data.size))

    def __len__(self):
        return self.data.size

    def __getitem__(self, i):
        return self.data[i]

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)

    def probabilities(self, qargs):
        assert len(qargs) == 1
        return orc.probabilities(self.data, qargs[0])

    def partial_trace(self, a, b):
        """entanglement_measures.py:325-340"""
        if self.num_qubits == 2:
            return np.outer(self.data, self.data.conj())
        return orc.partial_trace(self.data, a, b)


class _Result:
    def __init__(self, sv):
        self._sv = sv

    def get_statevector(self):
        return self._sv


class _Job:
    def __init__(self, sv):
        self._sv = sv

    def result(self):
        return _Result(self._sv)


class OracleSVSimulator:
    """`Aer.get_backend("statevector_simulator")` stand-in: full simulation from |0..0> per run."""

    def __init__(self):
        self.runs = 0
        self.gates_applied = 0

    def run(self, circuit, **_options):
        gates = circuit_to_gates(circuit)
        self.runs += 1
        self.gates_applied += len(gates)
        return _Job(OracleStatevector(orc.evaluate_circuit(circuit.num_qubits, gates)))


class OracleSVBackend:
    kind = "sv"

    def __init__(self, simulator=None):
        self.simulator = simulator if simulator is not None else OracleSVSimulator()

    def evaluate_global_cost(self, compiler):
        if compiler.soften_global_cost:
            raise NotImplementedError("soften_global_cost is currently only implemented for AerMPSBackend")
        sv = self.evaluate_circuit(compiler)
        return 1 - (np.absolute(sv[0])) ** 2

    def evaluate_local_cost(self, compiler):
        e_vals = self.measure_qubit_expectation_values(compiler)
        return 0.5 * (1 - np.mean(e_vals))

    def evaluate_circuit(self, compiler):
        job = self.simulator.run(compiler.full_circuit, **compiler.backend_options, **compiler.execute_kwargs)
        return job.result().get_statevector()

    def measure_qubit_expectation_values(self, compiler):
        sv = self.evaluate_circuit(compiler)
        expectation_values = []
        for i in range(sv.num_qubits):
            [p0, p1] = sv.probabilities([i])
            expectation_values.append(p0 - p1)
        return expectation_values


class OracleMPSBackend:
    """Restates ``AerMPSBackend`` (adaptaqc/backends/aer_mps_backend.py:45-93) on the numpy MPS
    oracle, including its redundancy: every evaluation re-applies all un-absorbed gates to the
    target MPS (one SVD per 2-qubit gate) and the whole MPS is handed to host numpy."""

    kind = "mps"

    def __init__(self, simulator=None):
        from . import mps_oracle as mo
        self.mps_ops = mo
        self.simulator = simulator if simulator is not None else mo.OracleMPSSimulator()

    def evaluate_global_cost(self, compiler):
        mo = self.mps_ops
        circ_mps = self.evaluate_circuit(compiler)
        global_cost = 1 - np.absolute(mo.mps_dot(circ_mps, compiler.zero_mps, already_preprocessed=True)) ** 2
        if not compiler.soften_global_cost:
            return global_cost
        previous_cost = compiler.global_cost_history[-1] if len(compiler.global_cost_history) > 0 else 1
        alpha = abs(previous_cost - compiler.adapt_config.sufficient_cost)
        return global_cost - alpha * sum(self.evaluate_hamming_weight_one_overlaps(circ_mps))

    def evaluate_local_cost(self, compiler):
        evals = self.measure_qubit_expectation_values(compiler)
        return 0.5 * (1 - np.mean(evals))

    def evaluate_circuit(self, compiler):
        circ = compiler.full_circuit.copy()
        return self.mps_ops.mps_from_circuit(circ, return_preprocessed=True, sim=self.simulator)

    def measure_qubit_expectation_values(self, compiler):
        mps = self.evaluate_circuit(compiler)
        return [self.mps_ops.mps_expectation(mps, "Z", i, already_preprocessed=True)
                for i in range(compiler.full_circuit.num_qubits)]

    def evaluate_hamming_weight_one_overlaps(self, mps):
        return [abs(self.mps_ops.extract_amplitude(mps, 2 ** i, already_preprocessed=True)) ** 2
                for i in range(len(mps))]
