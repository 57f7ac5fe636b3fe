"""Pins the CPU oracle on the reference's own known-answer tests.

The third-party simulators the reference calls are not installable here, so the oracle is
anchored on the analytic values the reference's tests assert, plus an independent dense-matrix
simulator written here (full 2^n x 2^n unitaries via kron) for random circuits.
"""
import numpy as np
import pytest

from harness import measures as em
from harness.circuit import Circuit
from harness.compiler import AdaptCompiler
from oracle import sv_oracle as orc
from oracle.oracle_backends import OracleSVBackend, circuit_to_gates

from helpers import circuit_from_gates, random_gates


def dense_unitary(n, gate):
    """Independent check: embed one gate into the full 2^n x 2^n matrix (little-endian)."""
    name, qubits, params = gate
    c = Circuit(max(2, len(qubits)))
    if name in ("mat1", "mat2"):
        m = np.asarray(params, dtype=np.complex128)
    else:
        from harness.circuit import Gate
        m = Gate(name, params).to_matrix()
    k = len(qubits)
    dim = 1 << n
    u = np.zeros((dim, dim), dtype=np.complex128)
    for col in range(dim):
        sub = sum(((col >> q) & 1) << j for j, q in enumerate(qubits))
        rest = col
        for q in qubits:
            rest &= ~(1 << q)
        for row_sub in range(1 << k):
            row = rest
            for j, q in enumerate(qubits):
                row |= ((row_sub >> j) & 1) << q
            u[row, col] = m[row_sub, sub]
    return u


def dense_simulate(n, gates):
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1
    for g in gates:
        psi = dense_unitary(n, g) @ psi
    return psi


@pytest.mark.parametrize("n", [1, 2, 3, 5, 7])
def test_oracle_matches_dense_matrix_simulation(n):
    rng = np.random.default_rng(100 + n)
    for _ in range(3):
        gates = random_gates(n, 40, rng)
        np.testing.assert_allclose(orc.evaluate_circuit(n, gates), dense_simulate(n, gates), atol=1e-12)


def test_analytic_costs_of_simple_states():
    """test/recompilers/test_approximate_compiler.py:114-150."""
    analytic_costs = [0, 0, 1, 1 / 2, 1 / 2, 1 / 2, 15 / 16, 1 / 2]
    zero = Circuit(4)
    neel = Circuit(4); neel.x([0, 2])
    ghz = Circuit(4); ghz.h(0)
    for i in range(3):
        ghz.cx(0, i + 1)
    hadamard = Circuit(4); hadamard.h([0, 1, 2, 3])
    costs = []
    for circuit in [zero, neel, ghz, hadamard]:
        for optimise_local_cost in [False, True]:
            compiler = AdaptCompiler(circuit, backend=OracleSVBackend(), optimise_local_cost=optimise_local_cost)
            costs.append(compiler.evaluate_cost())
    np.testing.assert_allclose(costs, analytic_costs, atol=1e-14)


def test_sigma_z_expectations():
    """test/utils/test_utilityfunctions.py:86-95: X on q0, H on q1 -> [-1, 0, 1] to 15 decimals."""
    qc = Circuit(3); qc.x(0); qc.h(1)
    sv = orc.evaluate_circuit(3, circuit_to_gates(qc))
    np.testing.assert_array_almost_equal(orc.measure_qubit_expectation_values(sv), [-1.0, 0.0, 1.0], decimal=15)


def test_ghz5_amplitudes():
    """test/utils/circuit_operations/test_circuit_operations_running.py:48-61."""
    qc = Circuit(5); qc.h(0)
    for i in range(4):
        qc.cx(i, i + 1)
    sv = orc.evaluate_circuit(5, circuit_to_gates(qc))
    expect = np.zeros(32, dtype=np.complex128)
    expect[0] = expect[31] = 1 / np.sqrt(2)
    np.testing.assert_allclose(sv, expect, atol=1e-15)


def test_local_cost_not_above_global_cost():
    """test/recompilers/test_approximate_compiler.py:152-163."""
    rng = np.random.default_rng(5)
    for _ in range(5):
        qc = circuit_from_gates(4, random_gates(4, 30, rng, allow_mat=False))
        local = AdaptCompiler(qc, backend=OracleSVBackend(), optimise_local_cost=True).evaluate_cost()
        glob = AdaptCompiler(qc, backend=OracleSVBackend()).evaluate_cost()
        assert local <= glob + 1e-12


def test_concurrence_of_pure_two_qubit_state():
    """test/utils/test_entanglement_measures.py:47-51 restated: for a pure state (a,b,c,d) the
    Wootters concurrence is 2|ad - bc| (what qiskit.quantum_info.concurrence returns)."""
    rng = np.random.default_rng(0)
    for _ in range(5):
        v = rng.normal(size=4) + 1j * rng.normal(size=4)
        v /= np.linalg.norm(v)
        rho = np.outer(v, v.conj())
        assert abs(em.concurrence(rho) - 2 * abs(v[0] * v[3] - v[1] * v[2])) < 1e-7


def test_partial_trace_against_einsum():
    """entanglement_measures.py:325-340: RDM on (a,b), lower qubit least significant."""
    rng = np.random.default_rng(3)
    n = 6
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi /= np.linalg.norm(psi)
    t = psi.reshape([2] * n)  # axis k <-> qubit n-1-k
    for a, b in [(0, 1), (1, 4), (2, 5), (5, 3)]:
        lo, hi = min(a, b), max(a, b)
        moved = np.moveaxis(t, [n - 1 - hi, n - 1 - lo], [0, 1]).reshape(4, -1)  # index = 2*hi_bit + lo_bit
        expect = moved @ moved.conj().T
        np.testing.assert_allclose(orc.partial_trace(psi, a, b), expect, atol=1e-14)
    rho = orc.partial_trace(psi, 0, 3)
    assert abs(np.trace(rho) - 1) < 1e-14
    np.testing.assert_allclose(rho, rho.conj().T, atol=1e-15)


def test_toffoli_expansion_is_exact():
    """README 3-qubit example uses ccx(2,1,0); the harness expands it to the 6-CX network."""
    qc = Circuit(3); qc.ccx(2, 1, 0)
    u = np.eye(8, dtype=np.complex128)
    for g in circuit_to_gates(qc):
        u = dense_unitary(3, g) @ u
    expect = np.eye(8, dtype=np.complex128)
    expect[[6, 7]] = expect[[7, 6]]  # controls q2,q1 -> flip q0
    phase = u[0, 0]
    np.testing.assert_allclose(u / phase, expect, atol=1e-14)


def test_bell_pair_concurrence_through_the_reference_call_chain():
    """One full re-simulation + host partial trace per pair, as adapt_compiler.py:964-975."""
    qc = Circuit(3); qc.h(0); qc.cx(0, 1)
    backend = OracleSVBackend()
    compiler = AdaptCompiler(qc, backend=backend)
    ems = compiler._get_all_qubit_pair_entanglement_measures()
    assert compiler.coupling_map == [(0, 1), (1, 2), (0, 2)]
    np.testing.assert_allclose(ems, [1.0, 0.0, 0.0], atol=1e-7)
    assert backend.simulator.runs == 3  # the reference's redundancy is part of the restatement
