"""Pins the MPS oracle (oracle/mps_oracle.py).

Anchors: the reference's MPS known-answer tests, the statevector oracle (exact equality of the
represented state for untruncated simulation), and three of the reference's own 50-site chi=2
fixtures (tests/golden/random_mps_seed_*.npz, genuine qiskit-aer outputs)."""
import os

import numpy as np
import pytest

from harness.circuit import Circuit
from oracle import mps_oracle as mo
from oracle import sv_oracle as orc
from oracle.oracle_backends import circuit_to_gates

from helpers import circuit_from_gates, golden_seeds, load_golden_mps, random_gates


def test_golden_fixtures_are_normalised_and_canonical():
    """Gamma_i . lambda_i (right) gives <psi|psi> = 1 on the paper's targets; lambdas
    are descending with unit 2-norm (utilityfunctions.py:309-311)."""
    assert len(golden_seeds()) == 54          # every target the reference ships (paper/random_mps)
    for seed in golden_seeds():
        mps = load_golden_mps(seed)
        assert mo.check_mps(mps) and len(mps[0]) == 50
        assert abs(mo.mps_dot(mps, mps) - 1) < 1e-13
        for lam in mps[1]:
            assert np.all(np.diff(lam) <= 0) and abs(np.sum(lam ** 2) - 1) < 1e-13
        # Vidal canonical form of genuine Aer outputs: pins which side each lambda multiplies (A.5)
        for i, (a0, a1) in enumerate(mps[0]):
            if i < 49:        # sum_s (Gamma_s lambda_i)(Gamma_s lambda_i)^+ weighted on the left by lambda_{i-1}^2 ...
                lr = mps[1][i]
                ll = mps[1][i - 1] if i > 0 else np.ones(1)
                left = sum((ll[:, None] * a).conj().T @ (ll[:, None] * a) for a in (a0, a1))
                np.testing.assert_allclose(left, np.eye(len(lr)), atol=1e-7)
                right = sum((a * lr[None, :]) @ (a * lr[None, :]).conj().T for a in (a0, a1))
                np.testing.assert_allclose(right, np.eye(len(ll)), atol=1e-7)
        pp = mo._preprocess_mps(mps)
        assert pp[0].shape == (2, 1, 2) and pp[-1].shape == (2, 2, 1)
        z = [mo.mps_expectation(pp, "Z", q, already_preprocessed=True) for q in (0, 24, 49)]
        assert all(-1 <= v <= 1 for v in z)
        rho = mo.partial_trace(pp, [10, 11], already_preprocessed=True)
        assert abs(np.trace(rho) - 1) < 1e-13
        np.testing.assert_allclose(rho, rho.conj().T, atol=1e-14)


def test_set_mps_round_trip_on_golden_fixture():
    """test_utilityfunctions.py:317-338: set_matrix_product_state loads (Gamma, lambda) verbatim."""
    mps = load_golden_mps(17)
    qc = Circuit(50)
    qc.set_matrix_product_state(mps)
    out = mo.mps_from_circuit(qc)
    assert abs(abs(mo.mps_dot(out, mps)) ** 2 - 1) < 1e-10
    for (a0, a1), (b0, b1) in zip(mps[0], out[0]):
        np.testing.assert_allclose(a0, b0); np.testing.assert_allclose(a1, b1)
    for la, lb in zip(mps[1], out[1]):
        np.testing.assert_allclose(la, lb)


def test_preprocess_product_state_example():
    """test/utils/test_entanglement_measures.py:77-85."""
    one, zero = np.array([[1.0 + 0j]]), np.array([[0.0 + 0j]])
    mps = ([(one, zero), (one, zero)], [np.array([1.0])])
    pp = mo._preprocess_mps(mps)
    assert len(pp) == 2 and pp[0].shape == (2, 1, 1)
    assert mo.extract_amplitude(pp, 0, already_preprocessed=True) == 1
    np.testing.assert_allclose(mo.partial_trace(pp, [0, 1], already_preprocessed=True), np.diag([1, 0, 0, 0]))


def test_pauli_z_expectation_kats():
    """test/utils/test_utilityfunctions.py:201-211."""
    qc = Circuit(4)
    pp = mo.mps_from_circuit(qc.copy(), return_preprocessed=True)
    np.testing.assert_allclose([mo.mps_expectation(pp, "Z", i, True) for i in range(4)], [1, 1, 1, 1])
    qc.h([0, 1, 2, 3])
    pp = mo.mps_from_circuit(qc.copy(), return_preprocessed=True)
    np.testing.assert_allclose([mo.mps_expectation(pp, "Z", i, True) for i in range(4)], [0, 0, 0, 0], atol=1e-7)


def test_mps_from_circuit_appends_save_instruction_in_place():
    """test_utilityfunctions.py:186-193: callers must pass a copy."""
    qc = Circuit(3); qc.h(0)
    n0 = len(qc.data)
    mo.mps_from_circuit(qc)
    assert len(qc.data) == n0 + 1 and qc.data[-1].operation.name == "save_matrix_product_state"


@pytest.mark.parametrize("n", [2, 3, 5, 8, 10])
def test_untruncated_simulation_equals_statevector_oracle(n):
    rng = np.random.default_rng(300 + n)
    for _ in range(3):
        gates = random_gates(n, 40, rng)
        qc = circuit_from_gates(n, gates)
        sv = orc.evaluate_circuit(n, circuit_to_gates(qc))
        pp = mo.mps_from_circuit(qc.copy(), return_preprocessed=True)
        np.testing.assert_allclose(mo.mps_to_vector(pp, True), sv, atol=1e-10)
        assert abs(mo.mps_dot(pp, mo.zero_mps(n), True) - np.conj(sv[0])) < 1e-10
        np.testing.assert_allclose([mo.mps_expectation(pp, "Z", q, True) for q in range(n)],
                                   orc.measure_qubit_expectation_values(sv), atol=1e-10)
        for b in (1, (1 << n) - 1, 1 << (n - 1)):
            assert abs(mo.extract_amplitude(pp, b, True) - sv[b]) < 1e-10
        if n >= 3:
            for a, b in [(0, 1), (0, n - 1), (n - 1, 1)]:
                np.testing.assert_allclose(mo.partial_trace(pp, [a, b], True), orc.partial_trace(sv, a, b),
                                           atol=1e-10)


def test_readme_mps_example_costs():
    """README 50-qubit example (BASELINE config C2): Bell pairs (0,1), (2,3) + H on the rest.
    |<0|psi>|^2 = (1/2)^2 (1/2)^46."""
    n = 50
    qc = Circuit(n)
    qc.h(0); qc.cx(0, 1); qc.h(2); qc.cx(2, 3); qc.h(list(range(4, n)))
    pp = mo.mps_from_circuit(qc.copy(), return_preprocessed=True)
    assert max(g.shape[2] for g in pp) == 2
    amp = mo.mps_dot(pp, mo.zero_mps(n), True)
    assert abs(abs(amp) ** 2 - 0.5 ** 48) < 1e-25
    rho = mo.partial_trace(pp, [0, 1], True)
    bell = np.zeros((4, 4)); bell[0, 0] = bell[0, 3] = bell[3, 0] = bell[3, 3] = 0.5
    np.testing.assert_allclose(rho, bell, atol=1e-12)
    np.testing.assert_allclose(mo.partial_trace(pp, [1, 2], True), np.kron(np.eye(2) / 2, np.eye(2) / 2), atol=1e-12)


def test_truncation_rule():
    """Aer reduce_zeros in the default ("aer") reading -- chop sigma^2 <= 1e-16, cap at max chi, drop the smallest while
    the sum of squares stays below the threshold (count unchanged if that loop runs out), renormalise only if the
    count dropped below the chopped count -- and in the round-1 ("sigma") reading.  More cases: tests/test_chop_rule.py."""
    S = np.array([0.9, 0.4, 0.1, 1e-5, 1e-9, 1e-17])
    try:
        mo.set_chop_rule("aer")
        k, kept = mo.reduce_zeros(S, None, 1e-16)
        assert k == 4                                                   # 1e-9 and 1e-17 chopped (sigma^2 <= 1e-16) ...
        np.testing.assert_array_equal(kept, S[:4])                      # ... which does not renormalise
        k, kept = mo.reduce_zeros(S, None, 1e-8)
        assert k == 3 and abs(np.sum(kept ** 2) - 1) < 1e-15            # 1e-10 < 1e-8 dropped, + 1e-2 is not
        k, kept = mo.reduce_zeros(S, 2, 1e-16)
        assert k == 2 and abs(np.sum(kept ** 2) - 1) < 1e-15
        k, kept = mo.reduce_zeros(np.array([1.0, 0.5]), None, 10.0)
        assert k == 2                                                   # the drop loop ran out: count unchanged
        mo.set_chop_rule("sigma")
        k, kept = mo.reduce_zeros(S, None, 1e-16)
        assert k == 4 and abs(np.sum(kept ** 2) - 1) < 1e-15            # 1e-17 chopped, (1e-9)^2 < 1e-16 dropped
        k, kept = mo.reduce_zeros(np.array([1.0, 0.5]), None, 10.0)
        assert k == 1
    finally:
        mo.set_chop_rule("aer")
    k, kept = mo.reduce_zeros(np.array([0.8, 0.6]), None, 1e-16)
    np.testing.assert_array_equal(kept, [0.8, 0.6])                 # nothing dropped -> no renormalisation


def test_max_bond_dimension_is_honoured():
    n = 8
    rng = np.random.default_rng(2)
    qc = circuit_from_gates(n, random_gates(n, 80, rng, allow_mat=False))
    sim = mo.OracleMPSSimulator(1e-16, max_chi=4)
    mps = mo.mps_from_circuit(qc.copy(), sim=sim)
    assert max(len(l) for l in mps[1]) <= 4
