"""The backends against circuit objects shaped like qiskit >= 1.0 (tests/fake_qiskit.py): a fresh CircuitInstruction
and a fresh operation per ``data[i]`` access, Qubit objects instead of ints.  Round-1 finding: edit detection by
object identity flagged every gate as changed on every call, so the incremental fast paths never triggered on the real
reference.  The translator now diffs by value (gates.instruction_key); these tests pin (a) the results and (b) that the
device work (R moves, transfer passes) is THE SAME as with the identity-stable harness container."""
import numpy as np
import pytest

from adapt_aqc_b200 import gates as G
from adapt_aqc_b200.backends import B200SVBackend
from harness.circuit import Circuit, CircuitInstruction, Gate
from harness.compiler import AdaptCompiler
from harness.minimiser import replace_1q_gate
from adapt_aqc_b200.sv_engine import SVCostEvaluator
from oracle.oracle_backends import OracleSVBackend

from fake_qiskit import FreshCircuitView, ViewBackend
from helpers import FakeEngine, brickwork, thin_ansatz


@pytest.fixture
def fake_engine_backend(emu, monkeypatch):
    def _get_engine(self, num_qubits):
        if self._engine is None or self._engine.num_qubits != num_qubits:
            self._engine = FakeEngine(emu, num_qubits)
            self._evaluator = SVCostEvaluator(self._engine, None, None)
            self._state_version += 1
            self._last_run_key = None
        return self._engine
    monkeypatch.setattr(B200SVBackend, "_get_engine", _get_engine)
    return B200SVBackend


def test_canonical_window_and_keys_on_fresh_objects():
    c = Circuit(3)
    c.rx(0.3, 0, label="rx"); c.cx(0, 2); c.u3(0.1, 0.2, 0.3, 1); c.barrier() if hasattr(c, "barrier") else None
    v = FreshCircuitView(c)
    assert v.data[0] is not v.data[0] and v.data[0].operation is not v.data[0].operation
    assert G.canonical_window(v) == G.canonical_window(c)
    qmap = G.qubit_indices(v)
    assert G.instruction_key(v.data[0], qmap) == ("rx", (0.3,), (0,)) == G.instruction_key(c.data[0], None)
    assert G.instruction_key(v.data[1], qmap) == ("cx", (), (0, 2))
    c.unitary(np.eye(2), [1])
    assert G.instruction_key(c.data[-1], None) is None         # matrix-valued parameters: identity only


def test_fresh_instruction_objects_keep_the_incremental_fast_path(fake_engine_backend):
    n = 8
    target, trng = brickwork(n, 2, seed=11)
    ansatz = thin_ansatz(n, 6, trng)
    stats = {}
    costs = {}
    for kind in ("identity", "fresh"):
        inner = fake_engine_backend()
        backend = inner if kind == "identity" else ViewBackend(inner)
        comp = AdaptCompiler(target, backend=backend)
        comp.full_circuit.data.extend(ansatz.copy().data)
        out = [comp.evaluate_cost()]
        lo, hi = comp.variational_circuit_range()
        out.append(comp.minimizer._reduce_cost(True, (hi - 5, hi)))      # Rotoselect over the newest layer
        out.append(comp.minimizer._reduce_cost(False, (lo, hi)))         # one Rotosolve cycle
        out.append(comp.evaluate_cost())
        costs[kind] = out
        stats[kind] = dict(inner._evaluator.stats)
        runs = inner._engine.runs
        stats[kind]["engine_runs"] = runs
    np.testing.assert_allclose(costs["fresh"], costs["identity"], rtol=0, atol=1e-13)
    assert stats["fresh"] == stats["identity"], (stats["fresh"], stats["identity"])
    served_on_host = stats["fresh"]["host_evals"] + stats["fresh"].get("hot_evals", 0)
    assert served_on_host > 10 * (stats["fresh"]["t_passes"] + stats["fresh"]["t_gathers"])
    assert stats["fresh"].get("hot_evals", 0) > 0          # repeated values of one gate: four multiplications each
    # and the values are right
    ocomp = AdaptCompiler(target, backend=OracleSVBackend())
    ocomp.full_circuit.data.extend(ansatz.copy().data)
    assert abs(ocomp.evaluate_cost() - costs["fresh"][0]) < 1e-12


def test_facade_run_recognises_the_same_circuit_by_value(fake_engine_backend):
    """simulator.run(circuit) is called once per candidate pair with the same circuit
    (circuit_operations_running.py:58-63): only the first call simulates."""
    inner = fake_engine_backend()
    target, _ = brickwork(5, 2, seed=3)
    view = FreshCircuitView(target)
    sv0 = inner.simulator.run(view).result().get_statevector()
    runs = inner._engine.runs
    for _ in range(4):
        assert inner.simulator.run(view).result().get_statevector() is sv0
    assert inner._engine.runs == runs
    replace_1q_gate(target, 0, "rx", 0.5)
    assert inner.simulator.run(view).result().get_statevector() is not sv0
    assert inner._engine.runs == runs + 1


def test_shift_costs_maps_circuit_indices_past_barriers(fake_engine_backend):
    """gate_index is an index into full_circuit.data; the canonical window drops barriers, so the window position is
    not gate_index - lhs_gate_count (round-1 advisor finding: the wrong gate was replaced silently)."""
    n = 5
    target, trng = brickwork(n, 2, seed=5)
    ansatz = thin_ansatz(n, 3, trng)
    inner = fake_engine_backend()
    comp = AdaptCompiler(target, backend=inner)
    ocomp = AdaptCompiler(target, backend=OracleSVBackend())
    for c in (comp, ocomp):
        data = ansatz.copy().data
        c.full_circuit.data.extend(data[:5])
        c.full_circuit.data.append(CircuitInstruction(Gate("barrier"), list(range(n))))
        c.full_circuit.data.extend(data[5:])
    lo, hi = comp.variational_circuit_range()
    barrier_at = lo + 5
    assert comp.full_circuit.data[barrier_at].operation.name == "barrier"
    for idx in (lo + 1, barrier_at + 1, hi - 1):
        got = inner.shift_costs(comp, idx, [("ry", 0.4), ("rx", -0.2)])
        saved = ocomp.full_circuit.data[idx]
        ref = []
        for name, th in (("ry", 0.4), ("rx", -0.2)):
            replace_1q_gate(ocomp.full_circuit, idx, name, th)
            ref.append(ocomp.evaluate_cost())
        ocomp.full_circuit.data[idx] = saved
        np.testing.assert_allclose(got, ref, atol=1e-10)
    with pytest.raises(ValueError):
        inner.shift_costs(comp, barrier_at, [("ry", 0.4)])
    with pytest.raises(ValueError):
        inner.shift_costs(comp, lo + 2, [("ry", 0.4)])           # the cx of the first layer
