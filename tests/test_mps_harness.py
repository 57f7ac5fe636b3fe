"""MPS side of the compile-loop mirror on the CPU oracle backends (no GPU): pins the harness +
MPS oracle on the reference's own MPS tests (layer-absorption schedules, SV-vs-MPS agreement,
general-gradient analytic value)."""
import numpy as np
import pytest

from harness import gradients as gr
from harness import measures as em
from harness.circuit import Circuit
from harness.compiler import AdaptCompiler, AdaptConfig
from oracle import mps_oracle as mo
from oracle.oracle_backends import OracleMPSBackend, OracleSVBackend

from helpers import circuit_from_gates, load_golden_mps, random_gates


def _random_target(n, seed, gates=20):
    """Generic (non-Clifford) random state: u3 layers + cx ladder."""
    rng = np.random.default_rng(seed)
    c = Circuit(n)
    for layer in range(3):
        for q in range(n):
            c.u3(*rng.uniform(-np.pi, np.pi, 3), q)
        for q in range(layer % 2, n - 1, 2):
            c.cx(q, q + 1)
    return c


def test_sv_and_mps_costs_agree():
    """test/recompilers/test_approximate_compiler.py:78-112 (5 decimals)."""
    qc = _random_target(4, 1)
    for local in (False, True):
        c_sv = AdaptCompiler(qc, backend=OracleSVBackend(), optimise_local_cost=local).evaluate_cost()
        c_mps = AdaptCompiler(qc, backend=OracleMPSBackend(), optimise_local_cost=local).evaluate_cost()
        assert abs(c_sv - c_mps) < 1e-10


def test_sv_and_mps_entanglement_measures_agree():
    """test/utils/test_entanglement_measures.py:93-112 (atol 1e-6)."""
    qc = _random_target(3, 2)
    for method in (em.EM_TOMOGRAPHY_CONCURRENCE, em.EM_TOMOGRAPHY_NEGATIVITY, em.EM_TOMOGRAPHY_EOF):
        sv = AdaptCompiler(qc, entanglement_measure=method, backend=OracleSVBackend())
        mps = AdaptCompiler(qc, entanglement_measure=method, backend=OracleMPSBackend())
        np.testing.assert_allclose(sv._get_all_qubit_pair_entanglement_measures(),
                                   mps._get_all_qubit_pair_entanglement_measures(), atol=1e-6)


def test_layer_absorption_schedules():
    """test/recompilers/test_adapt_compiler.py:673-718."""
    qc = _random_target(4, 3)
    compiler = AdaptCompiler(qc, backend=OracleMPSBackend(),
                             adapt_config=AdaptConfig(rotosolve_frequency=4, max_layers_to_modify=3))
    got = []
    for i in range(13):
        compiler._add_layer(i)
        got.append(len(compiler.full_circuit.data) - 1)
    assert got == [0, 0, 5, 10, 0, 0, 5, 10, 0, 0, 5, 10, 0]
    compiler = AdaptCompiler(qc, backend=OracleMPSBackend(),
                             adapt_config=AdaptConfig(rotosolve_frequency=4, max_layers_to_modify=5))
    got = []
    for i in range(13):
        compiler._add_layer(i)
        got.append(len(compiler.full_circuit.data) - 1)
    assert got == [5, 10, 15, 20, 5, 10, 15, 20, 5, 10, 15, 20, 5]


def test_mps_compile_reaches_sufficient_cost_and_matches_sv_decisions():
    ghz = Circuit(4); ghz.h(0)
    for i in range(3):
        ghz.cx(i, i + 1)
    r_mps = AdaptCompiler(ghz, backend=OracleMPSBackend(), adapt_config=AdaptConfig(max_layers=8)).compile()
    r_sv = AdaptCompiler(ghz, backend=OracleSVBackend(), adapt_config=AdaptConfig(max_layers=8)).compile()
    assert r_mps.overlap > 1 - 1e-2
    assert r_mps.qubit_pair_history == r_sv.qubit_pair_history
    np.testing.assert_allclose(r_mps.global_cost_history, r_sv.global_cost_history, atol=1e-8)


def test_mps_target_from_golden_fixture_compiles():
    """A QiskitMPS target (paper/random_mps format) goes in through set_matrix_product_state
    (approximate_compiler.py:196-204).  8-site slice of the chi=2 fixture, renormalised."""
    gammas, lambdas = load_golden_mps(1)
    n = 6
    gam = [(a0.copy(), a1.copy()) for a0, a1 in gammas[:n]]
    gam[-1] = (gam[-1][0][:, :1].copy(), gam[-1][1][:, :1].copy())
    mps = (gam, [l.copy() for l in lambdas[:n - 1]])
    nrm = np.sqrt(abs(mo.mps_dot(mps, mps)))
    mps[0][0] = (mps[0][0][0] / nrm, mps[0][0][1] / nrm)
    comp = AdaptCompiler(mps, backend=OracleMPSBackend(), adapt_config=AdaptConfig(max_layers=10))
    assert comp.full_circuit.data[0].operation.name == "set_matrix_product_state"
    res = comp.compile()
    assert res.global_cost_history[-1] < res.global_cost_history[0] + 1e-12
    assert res.overlap > 0.9


def test_soften_global_cost_formula():
    """aer_mps_backend.py:58-70: C - alpha * sum_i |<2^i|psi>|^2, alpha = |prev - sufficient|."""
    qc = _random_target(4, 5)
    comp = AdaptCompiler(qc, backend=OracleMPSBackend(), soften_global_cost=True)
    comp.global_cost_history = []
    soft = comp.evaluate_cost()
    comp.soften_global_cost = False
    hard = comp.evaluate_cost()
    pp = comp.backend.evaluate_circuit(comp)
    hw1 = sum(abs(mo.extract_amplitude(pp, 1 << i, True)) ** 2 for i in range(4))
    assert abs(soft - (hard - abs(1 - comp.adapt_config.sufficient_cost) * hw1)) < 1e-12


def test_general_gradient_analytic_value():
    """test/utils/test_gradients.py:39-73: for |psi> = a|00> + b|01> + c|10> + d|11> and the
    identity-resolvable ansatz, the pair gradient is sqrt(Im(a* b)^2 + ... ) -- checked here in the
    reference's closed form for the thinly dressed CNOT: compare against direct evaluation of
    -Im(<s|G_k|psi><psi|U^+(0)|s>) with dense matrices."""
    rng = np.random.default_rng(7)
    v = rng.normal(size=4) + 1j * rng.normal(size=4)
    v /= np.linalg.norm(v)
    qc = Circuit(2); qc.unitary(_unitary_with_first_column(v), [0, 1])
    backend = OracleMPSBackend()
    comp = AdaptCompiler(qc, backend=backend, adapt_config=AdaptConfig(method="general_gradient"))
    grads = comp._get_all_qubit_pair_gradients()
    # dense restatement: psi = v (little-endian), |s> = |00>
    gens, degs = gr.get_generators_and_degeneracies(comp.layer_2q_gate, True, inverse=True)
    zero = np.zeros(4); zero[0] = 1
    u0_dag_s = _dense(comp.inverse_zero_ansatz) @ zero
    zero_overlap = np.vdot(v, u0_dag_s)
    tot = 0
    for g, d in zip(gens, degs):
        gs = _dense(g) @ zero
        tot += d * (-np.imag(np.vdot(gs, v) * zero_overlap)) ** 2
    assert abs(grads[0] - np.sqrt(tot)) < 1e-10
    # zero-gradient case (test_gradients.py:15-37): |00> has no gradient
    comp0 = AdaptCompiler(Circuit(2), backend=backend, adapt_config=AdaptConfig(method="general_gradient"))
    assert abs(comp0._get_all_qubit_pair_gradients()[0]) < 1e-12


def _unitary_with_first_column(v):
    m = np.eye(4, dtype=np.complex128)
    m[:, 0] = v
    q, r = np.linalg.qr(m)
    return q * (r[0, 0] / abs(r[0, 0]))


def _dense(circ):
    u = np.eye(4, dtype=np.complex128)
    for inst in circ.data:
        m = inst.operation.to_matrix()
        if len(inst.qubits) == 1:
            full = np.kron(np.eye(2), m) if inst.qubits[0] == 0 else np.kron(m, np.eye(2))
        else:
            full = m if tuple(inst.qubits) == (0, 1) else m[np.ix_([0, 2, 1, 3], [0, 2, 1, 3])]
        u = full @ u
    return u


def _dense_state(circ):
    from oracle import sv_oracle as orc
    from oracle.oracle_backends import circuit_to_gates
    return orc.evaluate_circuit(circ.num_qubits, circuit_to_gates(circ))


def _pair_transfers_dense(s, psi, pairs):
    """T_p[i][j] = <s|(|i><j| on pair p)|psi>, index = bit(pair[0]) + 2 bit(pair[1]), from dense little-endian vectors."""
    n = int(np.log2(psi.size))
    out = []
    for a, b in pairs:
        S = np.moveaxis(s.reshape([2] * n), [n - 1 - b, n - 1 - a], [0, 1]).reshape(4, -1)       # row = 2 bit(b) + bit(a)
        P = np.moveaxis(psi.reshape([2] * n), [n - 1 - b, n - 1 - a], [0, 1]).reshape(4, -1)
        out.append(S.conj() @ P.T)
    return np.array(out)


@pytest.mark.parametrize("rotoselect", [True, False])
@pytest.mark.parametrize("with_start", [False, True])
def test_pair_transfer_gradient_algebra_equals_the_reference_chain(rotoselect, with_start):
    """B200MPSBackend.general_grad_of_pairs gets every (pair, generator) overlap of
    adaptaqc/utils/gradients.py:23-124 as 4x4 algebra on T_p = <s|(|i><j|)_p|psi>.  Here T_p is built from dense
    vectors (the device read-out itself is GPU-tested) and the algebra is compared with the reference's chain -- one
    simulation + one mps_dot per (pair, generator) -- run on the oracle backend."""
    from adapt_aqc_b200.mps_backend import B200MPSBackend
    n = 5
    target = _random_target(n, 11)
    start = None
    if with_start:
        start = Circuit(n); start.x(1); start.h(3); start.cx(3, 4); start.ry(0.4, 0)
    cmap = [(0, 1), (1, 2), (2, 3), (3, 4), (0, 2), (4, 1), (3, 0)]           # incl. non-neighbours and (hi, lo) order
    comp = AdaptCompiler(target, backend=OracleMPSBackend(), coupling_map=cmap, starting_circuit=start,
                         use_rotoselect=rotoselect, adapt_config=AdaptConfig(method="general_gradient"))
    ref = comp._get_all_qubit_pair_gradients()
    psi = _dense_state(target)       # the chain uses full_circuit WITHOUT the inverse starting circuit (adapt_compiler.py:839-846)
    s = _dense_state(start) if start is not None else np.eye(1 << n)[0].astype(np.complex128)
    T = _pair_transfers_dense(s, psi, cmap)
    got = B200MPSBackend.gradients_from_pair_transfers(T, comp.inverse_zero_ansatz, comp.generators, comp.degeneracies)
    np.testing.assert_allclose(got, ref, atol=1e-10)
    if rotoselect or with_start:      # (rz-only generators on |0..0> give a real product: zero gradient)
        assert max(ref) > 1e-3
