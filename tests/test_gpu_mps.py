"""GPU parity tests of the MPS path (b200_mps_* through ctypes) against the numpy MPS oracle and
the statevector oracle.  Tolerance: BASELINE north_star asks for overlaps/costs within 1e-8 at the
same truncation threshold; untruncated amplitude-level checks use 1e-10."""
import numpy as np
import pytest

from harness import measures as em
from harness.circuit import Circuit
from harness.compiler import CMAP_LINEAR, AdaptCompiler, AdaptConfig, generate_coupling_map
from adapt_aqc_b200.gates import GateStream
from harness.minimiser import B200CostMinimiser, replace_1q_gate
from adapt_aqc_b200.mps_backend import B200MPSBackend, B200MPSSimulator, DeviceMPSView
from adapt_aqc_b200.mps_engine import MPSContext
from oracle import mps_oracle as mo
from oracle import sv_oracle as orc
from oracle.oracle_backends import OracleMPSBackend, OracleSVBackend, circuit_to_gates

from helpers import circuit_from_gates, golden_seeds, load_golden_mps, random_gates

pytestmark = pytest.mark.gpu

TOL = 1e-10
COST_TOL = 1e-8


@pytest.fixture(scope="module")
def ctx():
    c = MPSContext(0)
    yield c
    c.close()


def _generic_circuit(n, depth, seed, long_range=False):
    rng = np.random.default_rng(seed)
    c = Circuit(n)
    for layer in range(depth):
        for q in range(n):
            c.u3(*rng.uniform(-np.pi, np.pi, 3), q)
        for q in range(layer % 2, n - 1, 2):
            c.cx(q, q + 1)
        if long_range and n > 3:
            a, b = rng.choice(n, 2, replace=False)
            c.cz(int(a), int(b))
    return c


def _vector(mps):
    return mo.mps_to_vector(mps)


# ---- gate application: contraction + Jacobi SVD + truncation ------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 10])
def test_apply_all_opcodes_matches_statevector_oracle(ctx, n):
    rng = np.random.default_rng(7000 + n)
    m = ctx.new_mps(n)
    for _ in range(3):
        gates = random_gates(n, 40, rng)
        m.init_zero()
        m.apply(GateStream.from_gates(gates))
        sv = orc.evaluate_circuit(n, gates)
        np.testing.assert_allclose(_vector(m.get()), sv, atol=TOL)
        assert abs(m.amps([0])[0] - sv[0]) < TOL
    m.close()


def test_large_bond_dimension_uses_the_multi_cta_jacobi(ctx):
    """n = 14, deep circuit: the middle bond reaches 2^7 = 128 (> JACOBI_CTA_MAX_Q / 2 columns), so the
    per-round Jacobi kernels and 256 x 256 GEMMs are exercised; state equals the SV oracle."""
    n = 14
    c = _generic_circuit(n, 16, 11)
    m = ctx.new_mps(n)
    m.apply(GateStream.from_circuit(c))
    assert max(m.bond_dims()) == 128
    sv = orc.evaluate_circuit(n, circuit_to_gates(c))
    np.testing.assert_allclose(_vector(m.get()), sv, atol=1e-9)
    z, norm = m.expz()
    np.testing.assert_allclose(z, orc.measure_qubit_expectation_values(sv), atol=1e-9)
    assert abs(norm - 1) < 1e-9
    st = m.stats()
    assert st["svds"] > 0 and st["jacobi_sweeps"] > 0
    m.close()


def test_inverse_round_trip_and_long_range_gates(ctx):
    n = 9
    c = _generic_circuit(n, 5, 3, long_range=True)
    gs = GateStream.from_circuit(c)
    m = ctx.new_mps(n)
    m.apply(gs)
    sv = orc.evaluate_circuit(n, circuit_to_gates(c))
    np.testing.assert_allclose(_vector(m.get()), sv, atol=TOL)
    m.apply(gs, inverse=True)
    assert abs(abs(m.amps([0])[0]) - 1) < 1e-9
    m.close()


@pytest.fixture(params=["aer", "sigma"])
def chop_rule(request):
    """Both readings of Aer's reduce_zeros (include/b200aqc.h: B200_CHOP_AER default, B200_CHOP_SIGMA), product and
    oracle switched together."""
    from adapt_aqc_b200 import mps_engine as me
    me.set_chop_rule(request.param); mo.set_chop_rule(request.param)
    yield request.param
    me.set_chop_rule("aer"); mo.set_chop_rule("aer")


def test_truncation_rule_matches_oracle(ctx, chop_rule):
    n = 10
    c = _generic_circuit(n, 8, 5)
    # a weakly entangling tail: singular values between 1e-16 and 1e-8, where the two chop readings differ
    weak = c.copy()
    for q in range(0, n - 1, 2):
        weak.ry(3e-9, q); weak.cx(q, q + 1); weak.ry(-2e-9, q + 1)
    kept = {}
    for circ, thr, max_chi in [(c, 1e-6, None), (c, 1e-16, 6), (c, 1e-3, 8), (weak, 1e-16, None), (weak, 1e-20, None)]:
        m = ctx.new_mps(n, thr, max_chi)
        m.apply(GateStream.from_circuit(circ))
        ref = mo.mps_from_circuit(circ.copy(), sim=mo.OracleMPSSimulator(thr, max_chi))
        assert m.bond_dims() == [len(l) for l in ref[1]]
        got = m.get()
        for la, lb in zip(got[1], ref[1]):
            np.testing.assert_allclose(la, lb, atol=1e-9)
        assert abs(abs(mo.mps_dot(got, ref)) - abs(mo.mps_dot(ref, ref))) < 1e-8
        m.close()


# ---- set / get / read-outs ----------------------------------------------------------------------
def test_golden_fixture_round_trip_and_readouts(ctx):
    """ALL 54 of the reference's own 50-site chi=2 targets (tests/golden): set -> get is verbatim
    (test_utilityfunctions.py:317-338), <psi|psi> = 1, <Z>, amplitudes and RDMs equal the oracle's."""
    assert len(golden_seeds()) == 54
    for seed in golden_seeds():
        mps = load_golden_mps(seed)
        m = ctx.new_mps(50)
        m.set(mps)
        back = m.get()
        for (a0, a1), (b0, b1) in zip(mps[0], back[0]):
            np.testing.assert_array_equal(a0, b0); np.testing.assert_array_equal(a1, b1)
        for la, lb in zip(mps[1], back[1]):
            np.testing.assert_array_equal(la, lb)
        assert abs(m.dot(m) - 1) < 1e-12
        pp = mo._preprocess_mps(mps)
        z, norm = m.expz()
        np.testing.assert_allclose(z, [mo.mps_expectation(pp, "Z", q, True) for q in range(50)], atol=1e-12)
        assert abs(norm - 1) < 1e-12
        bits = [0, 1, 1 << 49, (1 << 50) - 1, 0x2AAAAAAAAAAAA]
        np.testing.assert_allclose(m.amps(bits), [mo.extract_amplitude(pp, b, True) for b in bits], atol=1e-15)
        pairs = [(0, 1), (10, 11), (3, 30), (49, 0), (24, 25)]
        for r, (a, b) in zip(m.pair_rdm(pairs), pairs):
            np.testing.assert_allclose(r, mo.partial_trace(pp, [a, b], True), atol=1e-12)
        m.close()


@pytest.mark.parametrize("n", [2, 4, 7])
def test_dot_transfer_and_rdm_match_oracle(ctx, n):
    ca, cb = _generic_circuit(n, 4, 20 + n), _generic_circuit(n, 3, 40 + n)
    a, b = ctx.new_mps(n), ctx.new_mps(n)
    a.apply(GateStream.from_circuit(ca)); b.apply(GateStream.from_circuit(cb))
    va = orc.evaluate_circuit(n, circuit_to_gates(ca)); vb = orc.evaluate_circuit(n, circuit_to_gates(cb))
    assert abs(a.dot(b) - np.vdot(va, vb)) < TOL
    for q in range(n):
        La = np.moveaxis(va.reshape([2] * n), n - 1 - q, 0).reshape(2, -1)
        Rb = np.moveaxis(vb.reshape([2] * n), n - 1 - q, 0).reshape(2, -1)
        np.testing.assert_allclose(a.transfer(b, [q]), La.conj() @ Rb.T, atol=TOL)
    for qa in range(n):
        for qb in range(n):
            if qa == qb:
                continue
            La = np.moveaxis(va.reshape([2] * n), [n - 1 - qb, n - 1 - qa], [0, 1]).reshape(4, -1)
            Rb = np.moveaxis(vb.reshape([2] * n), [n - 1 - qb, n - 1 - qa], [0, 1]).reshape(4, -1)
            np.testing.assert_allclose(a.transfer(b, [qa, qb]), La.conj() @ Rb.T, atol=TOL)
    if n >= 3:
        pairs = [(x, y) for x in range(n) for y in range(n) if x != y]
        for r, (x, y) in zip(a.pair_rdm(pairs), pairs):
            np.testing.assert_allclose(r, orc.partial_trace(va, x, y), atol=TOL)
    a.close(); b.close()


def test_errors(ctx):
    from adapt_aqc_b200.lib import B200Error
    m = ctx.new_mps(4)
    with pytest.raises(B200Error):
        m.apply(GateStream.from_gates([("cx", [0, 7], [])]))
    with pytest.raises(B200Error):
        m.pair_rdm([(1, 1)])
    with pytest.raises(B200Error):
        m.transfer(m, [0, 0])
    m.close()


# ---- backend level: same numbers and same decisions as the oracle-backed loop --------------------
def _targets():
    ghz = Circuit(4); ghz.h(0)
    for i in range(3):
        ghz.cx(i, i + 1)
    return {"ghz4": ghz, "generic4": _generic_circuit(4, 3, 9), "generic6": _generic_circuit(6, 2, 10)}


def test_mps_kats():
    """test/utils/test_utilityfunctions.py:201-211 and the README 50-qubit example (config C2)."""
    b = B200MPSBackend()
    comp = AdaptCompiler(Circuit(4), backend=b)
    np.testing.assert_allclose(b.measure_qubit_expectation_values(comp), [1, 1, 1, 1])
    had = Circuit(4); had.h([0, 1, 2, 3])
    comp = AdaptCompiler(had, backend=b)
    np.testing.assert_allclose(b.measure_qubit_expectation_values(comp), [0, 0, 0, 0], atol=1e-7)
    n = 50
    qc = Circuit(n)
    qc.h(0); qc.cx(0, 1); qc.h(2); qc.cx(2, 3); qc.h(list(range(4, n)))
    comp = AdaptCompiler(qc, backend=b)
    assert abs((1 - comp.evaluate_cost()) - 0.5 ** 48) < 1e-25
    ems = comp._get_all_qubit_pair_entanglement_measures()
    cm = comp.coupling_map
    assert abs(ems[cm.index((0, 1))] - 1) < 1e-7 and abs(ems[cm.index((2, 3))] - 1) < 1e-7
    assert max(e for p, e in zip(cm, ems) if p not in ((0, 1), (2, 3))) < 1e-7


@pytest.mark.parametrize("name", ["ghz4", "generic4", "generic6"])
def test_costs_and_heuristics_match_oracle_backend(name):
    target = _targets()[name]
    for local, soften in [(False, False), (True, False), (False, True)]:
        got = AdaptCompiler(target, backend=B200MPSBackend(), optimise_local_cost=local, soften_global_cost=soften)
        ref = AdaptCompiler(target, backend=OracleMPSBackend(), optimise_local_cost=local, soften_global_cost=soften)
        got.global_cost_history = ref.global_cost_history = []
        assert abs(got.evaluate_cost() - ref.evaluate_cost()) < COST_TOL
    got = AdaptCompiler(target, backend=B200MPSBackend())
    ref = AdaptCompiler(target, backend=OracleMPSBackend())
    np.testing.assert_allclose(got.backend.measure_qubit_expectation_values(got),
                               ref.backend.measure_qubit_expectation_values(ref), atol=COST_TOL)
    for method in (em.EM_TOMOGRAPHY_CONCURRENCE, em.EM_TOMOGRAPHY_NEGATIVITY):
        got.entanglement_measure_method = ref.entanglement_measure_method = method
        np.testing.assert_allclose(got._get_all_qubit_pair_entanglement_measures(),
                                   ref._get_all_qubit_pair_entanglement_measures(), atol=1e-7)
    sv = AdaptCompiler(target, backend=OracleSVBackend())
    assert abs(got.evaluate_cost() - sv.evaluate_cost()) < COST_TOL      # the reference's SV-vs-MPS bar is 1e-5


@pytest.mark.parametrize("name", ["ghz4", "generic4"])
@pytest.mark.parametrize("mode", ["incremental", "reference_order", "batched"])
def test_compile_decisions_match_oracle_backend(name, mode):
    target = _targets()[name]
    cfg = dict(max_layers=6)
    ref = AdaptCompiler(target, backend=OracleMPSBackend(), adapt_config=AdaptConfig(**cfg)).compile()
    backend = B200MPSBackend(incremental=(mode != "reference_order"))
    got = AdaptCompiler(target, backend=backend, adapt_config=AdaptConfig(**cfg),
                        minimiser_cls=B200CostMinimiser if mode == "batched" else None).compile()
    assert got.qubit_pair_history == ref.qubit_pair_history
    assert got.method_history == ref.method_history
    assert len(got.global_cost_history) == len(ref.global_cost_history)
    np.testing.assert_allclose(got.global_cost_history, ref.global_cost_history, atol=COST_TOL)
    two_q = lambda res: [(i.operation.name, i.qubits) for i in res.circuit.data if len(i.qubits) == 2]
    assert two_q(got) == two_q(ref)


def test_layer_absorption_schedule_on_device():
    """test/recompilers/test_adapt_compiler.py:673-718."""
    qc = _generic_circuit(4, 3, 4)
    compiler = AdaptCompiler(qc, backend=B200MPSBackend(),
                             adapt_config=AdaptConfig(rotosolve_frequency=4, max_layers_to_modify=3))
    got = []
    for i in range(9):
        compiler._add_layer(i)
        got.append(len(compiler.full_circuit.data) - 1)
    assert got == [0, 0, 5, 10, 0, 0, 5, 10, 0]


def test_general_gradient_pairs_match_oracle():
    target = _targets()["generic4"]
    cfg = AdaptConfig(method="general_gradient", max_layers=3)
    got = AdaptCompiler(target, backend=B200MPSBackend(), adapt_config=cfg)
    ref = AdaptCompiler(target, backend=OracleMPSBackend(), adapt_config=AdaptConfig(method="general_gradient", max_layers=3))
    np.testing.assert_allclose(got._get_all_qubit_pair_gradients(), ref._get_all_qubit_pair_gradients(), atol=COST_TOL)
    rg, rr = got.compile(), ref.compile()
    assert rg.qubit_pair_history == rr.qubit_pair_history


def test_golden_50_qubit_target_first_layers_match_oracle():
    """paper/random_mps target (50 sites, chi = 2): two ADAPT layers on a linear map, decisions and
    costs equal to the oracle-backed loop."""
    mps = load_golden_mps(1)
    cmap = generate_coupling_map(50, CMAP_LINEAR)
    cfg = dict(max_layers=2)
    got = AdaptCompiler(mps, backend=B200MPSBackend(), coupling_map=cmap, adapt_config=AdaptConfig(**cfg)).compile()
    ref = AdaptCompiler(mps, backend=OracleMPSBackend(), coupling_map=cmap, adapt_config=AdaptConfig(**cfg)).compile()
    assert got.qubit_pair_history == ref.qubit_pair_history
    np.testing.assert_allclose(got.global_cost_history, ref.global_cost_history, atol=COST_TOL)


def test_device_view_is_listlike_and_backend_pickles():
    import pickle
    b = B200MPSBackend(B200MPSSimulator(1e-12, 16))
    comp = AdaptCompiler(_targets()["generic4"], backend=b)
    view = b.evaluate_circuit(comp)
    assert isinstance(view, DeviceMPSView) and len(view) == 4 and view[0].shape[0] == 2
    ref = OracleMPSBackend().evaluate_circuit(AdaptCompiler(_targets()["generic4"], backend=OracleMPSBackend()))
    assert abs(abs(mo.mps_dot(list(view), ref, True)) - 1) < 1e-9
    b2 = pickle.loads(pickle.dumps(b))
    assert b2.simulator.options.matrix_product_state_truncation_threshold == 1e-12
    assert b2.simulator.options.matrix_product_state_max_bond_dimension == 16


def test_prefix_checkpoints_reproduce_full_reruns_bit_for_bit():
    """B200MPSSimulator.simulate resumes from the state after the last 2-qubit gate the previous run
    shares with the new gate list.  Under truncation (bond cap AND threshold) the resumed result must be
    IDENTICAL to a run from the base state by a simulator without history -- edits walk over every
    position of the un-absorbed window like Rotosolve does (cost_minimiser.py:344-368)."""
    n = 10
    rng = np.random.default_rng(3)
    base = _generic_circuit(n, 6, 11)
    target = mo.mps_from_circuit(base.copy(), sim=mo.OracleMPSSimulator(1e-16, None))
    for thr, max_chi in [(1e-16, 6), (1e-5, None)]:
        sim = B200MPSSimulator(thr, max_chi)
        qc = Circuit(n)
        qc.set_matrix_product_state(target)
        rot = []
        for layer in range(4):
            a = int(rng.integers(n - 1))
            for q in (a, a + 1):
                rot.append(len(qc.data)); qc.rz(float(rng.uniform(-3, 3)), q, label="rz")
            qc.cx(a, a + 1)
            for q in (a, a + 1):
                rot.append(len(qc.data)); qc.ry(float(rng.uniform(-3, 3)), q, label="ry")
        order = list(rot) + list(rng.permutation(rot))
        for idx in order:
            qc.data[idx].operation.params[0] = float(rng.uniform(-3, 3))
            h = sim.simulate(qc)
            got = h.get()
            sim._recycle(h)
            fresh = B200MPSSimulator(thr, max_chi)
            h2 = fresh.simulate(qc)
            ref = h2.get()
            h2.close()
            fresh.context().close()
            assert [len(l) for l in got[1]] == [len(l) for l in ref[1]]
            for (g0, g1), (r0, r1) in zip(got[0], ref[0]):
                assert np.array_equal(g0, r0) and np.array_equal(g1, r1)
            for la, lb in zip(got[1], ref[1]):
                assert np.array_equal(la, lb)
        assert sim.ckpt_stats["svds_skipped"] > 0 and sim.ckpt_stats["resumed_gates"] > 0


# ---- general-gradient heuristic from one batched read-out ---------------------------------------
def _dense(circ):
    from oracle import sv_oracle as orc
    return orc.evaluate_circuit(circ.num_qubits, circuit_to_gates(circ))


@pytest.mark.parametrize("n", [3, 6, 9])
def test_pair_transfer_matches_dense_contraction(ctx, n):
    """b200_mps_pair_transfer: T_p[i][j] = <s|(|i><j| on p)|psi> for all pairs from one left + one right sweep."""
    a, b = _generic_circuit(n, 3, 60 + n), _generic_circuit(n, 4, 80 + n)
    ma, mb = ctx.new_mps(n), ctx.new_mps(n)
    ma.apply(GateStream.from_circuit(a)); mb.apply(GateStream.from_circuit(b))
    va, vb = _dense(a), _dense(b)
    pairs = [(i, j) for i in range(n) for j in range(n) if i != j]            # both orders, all distances
    got = ma.pair_transfer(mb, pairs)
    for T, (p, q) in zip(got, pairs):
        A = np.moveaxis(va.reshape([2] * n), [n - 1 - q, n - 1 - p], [0, 1]).reshape(4, -1)
        B = np.moveaxis(vb.reshape([2] * n), [n - 1 - q, n - 1 - p], [0, 1]).reshape(4, -1)
        np.testing.assert_allclose(T, A.conj() @ B.T, atol=1e-12)
    ma.close(); mb.close()


@pytest.mark.parametrize("n,with_start,rotoselect", [(4, False, True), (6, True, True), (8, False, False), (10, True, False),
                                                     (10, False, True)])
def test_general_gradient_from_one_readout_equals_the_reference_chain(n, with_start, rotoselect):
    """B200MPSBackend.general_grad_of_pairs (one simulation + one batched read-out) against the reference's chain -- one
    simulator run + one mps_dot per (pair, generator), adaptaqc/utils/gradients.py:23-124 -- on the oracle backend AND on
    the device backend, all-to-all map."""
    from harness import gradients as gr
    target = _generic_circuit(n, 3, 100 + n)
    start = None
    if with_start:
        start = Circuit(n); start.x(0); start.h(n - 2); start.cx(n - 2, n - 1); start.ry(0.3, 1)
    cfg = dict(method="general_gradient")
    kw = dict(starting_circuit=start, use_rotoselect=rotoselect)
    dev = AdaptCompiler(target, backend=B200MPSBackend(), adapt_config=AdaptConfig(**cfg), **kw)
    ref = AdaptCompiler(target, backend=OracleMPSBackend(), adapt_config=AdaptConfig(**cfg), **kw)
    got = dev._get_all_qubit_pair_gradients()                       # device: batched read-out
    want = ref._get_all_qubit_pair_gradients()                      # oracle: the chain
    np.testing.assert_allclose(got, want, atol=COST_TOL)
    circuit = dev.full_circuit.copy()
    if start is not None:
        del circuit.data[len(circuit.data) - len(start.data):]
    chain_on_device = gr.general_grad_of_pairs(circuit, dev.inverse_zero_ansatz, dev.generators, dev.degeneracies,
                                               dev.coupling_map, start, dev.backend)
    np.testing.assert_allclose(got, chain_on_device, atol=COST_TOL)
    if rotoselect or with_start:
        assert max(want) > 1e-4


def test_general_gradient_analytic_value_and_zero_case():
    """test/utils/test_gradients.py:39-73 (ansatz rx on qubit 0, ry on qubit 1: gradient norm
    sqrt(Im(a* b)^2 + Re(a* c)^2) for psi = (a, b, c, d)) and :15-37 (no ansatz -> no gradient), through the device."""
    from harness import gradients as gr
    qc = _generic_circuit(2, 3, 5)
    a, b, c, _ = _dense(qc)
    expected = np.sqrt(np.imag(np.conj(a) * b) ** 2 + np.real(np.conj(a) * c) ** 2)
    ansatz = Circuit(2); ansatz.rx(0, 0); ansatz.ry(0, 1)
    gens, degs = gr.get_generators_and_degeneracies(ansatz, rotoselect=False, inverse=True)
    backend = B200MPSBackend()
    got = backend.general_grad_of_pairs(qc, ansatz.inverse(), gens, degs, [(0, 1)])[0]
    assert abs(got - expected) < 1e-10
    empty = Circuit(2)
    gens0, degs0 = gr.get_generators_and_degeneracies(empty)
    qc5 = _generic_circuit(5, 3, 6)
    start = Circuit(5); start.ry(0.7, 2); start.cx(2, 3)
    g0 = backend.general_grad_of_pairs(qc5, empty, gens0, degs0, [(0, 1), (1, 2), (2, 3), (3, 4)], start)
    np.testing.assert_allclose(g0, 0, atol=1e-12)


def test_batched_shift_costs_under_truncation_equal_the_one_scalar_path():
    """B200MPSBackend.shift_costs with REAL truncation (bond cap): the candidates are independent simulations run
    concurrently on worker contexts (own stream, own prefix checkpoints, one thread each); every value must be what the
    one-scalar-at-a-time path computes for the same circuit, and a Rotosolve cycle through the batched front end must
    leave the same circuit behind."""
    from harness.workloads import build_mps_workload
    n, chi = 16, 16
    target, ansatz = build_mps_workload(n, chi, 3)
    runs = {}
    for batched in (False, True):
        backend = B200MPSBackend(B200MPSSimulator(1e-16, max_chi=chi))
        backend.batch_truncating = True          # opt-in (measured slower than one scalar at a time on C4)
        comp = AdaptCompiler(target, backend=backend, minimiser_cls=B200CostMinimiser if batched else None)
        comp.full_circuit.data.extend(ansatz.copy().data)
        assert not backend._use_incremental(comp)
        lo, hi = comp.variational_circuit_range()
        rot = [i for i in range(lo, hi) if comp.full_circuit.data[i].operation.name == "rz"]
        if batched:
            cands = [("rx", 0.0), ("rx", np.pi / 2), ("ry", -np.pi / 2), ("rz", 0.7), ("ry", 1.1)]
            got = backend.shift_costs(comp, rot[2], cands)
            ref = []
            saved = comp.full_circuit.data[rot[2]]
            for name, th in cands:
                replace_1q_gate(comp.full_circuit, rot[2], name, th)
                ref.append(comp.evaluate_cost())
            comp.full_circuit.data[rot[2]] = saved
            np.testing.assert_allclose(got, ref, rtol=0, atol=1e-13)
        c = comp.minimizer._reduce_cost(False, (lo, hi))
        runs[batched] = (c, [(comp.full_circuit.data[i].operation.name, comp.full_circuit.data[i].operation.params[0]) for i in rot],
                         comp.cost_evaluation_counter)
    assert abs(runs[True][0] - runs[False][0]) < 1e-12
    for (na, ta), (nb, tb) in zip(runs[True][1], runs[False][1]):
        assert na == nb and abs(ta - tb) < 1e-9
