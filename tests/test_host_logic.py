"""Host logic without a GPU: the incremental evaluator, the backend adapter and the compile-loop
mirror are driven on a CPU stand-in engine (tests/helpers.FakeEngine = the product planner + the
kernels' thread bodies executed on the host) and compared with the oracle backend, which
re-simulates everything from |0..0> on every call like the reference does."""
import numpy as np
import pytest

from adapt_aqc_b200.backends import B200SVBackend
from harness.circuit import Circuit
from harness.compiler import AdaptCompiler, AdaptConfig
from harness.minimiser import B200CostMinimiser, replace_1q_gate
from adapt_aqc_b200.sv_engine import SVCostEvaluator
from oracle.oracle_backends import OracleSVBackend

from helpers import FakeEngine, brickwork, circuit_from_gates, compile_option_cases, random_gates, thin_ansatz


@pytest.fixture
def fake_backend(emu, monkeypatch, request):
    param = getattr(request, "param", None)
    compact_k, proj_k = param if isinstance(param, tuple) else (param, None)

    def _get_engine(self, num_qubits):
        if self._engine is None or self._engine.num_qubits != num_qubits:
            self._engine = FakeEngine(emu, num_qubits)
            compact = None
            if compact_k is not None and num_qubits > compact_k:
                compact = FakeEngine(emu, compact_k, n_slots=1)
            projected = None
            if isinstance(proj_k, (list, tuple)):      # nested projection levels
                projected = [FakeEngine(emu, k, n_slots=4) for k in proj_k if k < num_qubits]
            elif proj_k is not None and num_qubits >= proj_k + SVCostEvaluator.PROJECT_MIN_SAVING:
                projected = [FakeEngine(emu, proj_k, n_slots=4)]
            self._evaluator = SVCostEvaluator(self._engine, compact, projected)
            self._state_version += 1
            self._last_run_key = None
        return self._engine
    monkeypatch.setattr(B200SVBackend, "_get_engine", _get_engine)
    return B200SVBackend()


@pytest.fixture
def nested_levels(monkeypatch):
    """Registers of >= 8 qubits project onto engines ONE qubit smaller (production: >= 20 qubits), so that the small
    CPU stand-in engines exercise the multi-level nesting of the 28-qubit runs."""
    monkeypatch.setattr(SVCostEvaluator, "LARGE_QUBITS", 8)
    monkeypatch.setattr(SVCostEvaluator, "NEST_MIN_QUBITS", 6)


@pytest.mark.parametrize("fake_backend", [(None, (6, 8, 10, 11)), (5, (7, 9, 11))], indirect=True)
def test_nested_projection_levels_equal_full_resimulation(fake_backend, nested_levels):
    """The evaluator of a K-qubit projected engine projects its own tail onto the next smaller engine (one evaluator per
    engine, shared through the registry); edits walk across all levels and back; every cost equals the oracle's full
    re-simulation."""
    n = 12
    rng = np.random.default_rng(7)
    target, trng = brickwork(n, 3, seed=5)
    ansatz = thin_ansatz(n, 9, trng)
    comp = AdaptCompiler(target, backend=fake_backend)
    comp.full_circuit.data.extend(ansatz.data)
    ocomp = AdaptCompiler(target, backend=OracleSVBackend())
    ocomp.full_circuit.data.extend(ansatz.copy().data)
    rot = [i for i in range(*comp.variational_circuit_range())
           if comp.full_circuit.data[i].operation.name in ("rx", "ry", "rz")]
    order = rot + rot[::-1] + [rot[int(rng.integers(len(rot)))] for _ in range(30)]
    for step, idx in enumerate(order):
        for rep in range(1 + step % 2):
            name = ["rx", "ry", "rz"][int(rng.integers(3))]
            theta = float(rng.uniform(-np.pi, np.pi))
            for c in (comp, ocomp):
                replace_1q_gate(c.full_circuit, idx, name, theta)
            assert abs(comp.evaluate_cost() - ocomp.evaluate_cost()) < 1e-10, (step, idx)
    ev = fake_backend._evaluator
    levels = [e for e in ev._registry.values() if e is not ev and e.stats["evals"] > 0]
    assert len(levels) >= 3, [e.eng.num_qubits for e in levels]          # at least three nested levels did work
    assert sum(e.stats["projections"] for e in levels) > 0                # ... and projected further down themselves
    # the Rotosolve / Rotoselect stream of the bench step through the same levels
    lo, hi = comp.variational_circuit_range()
    olo, ohi = ocomp.variational_circuit_range()
    a = comp.minimizer._reduce_cost(True, (hi - 5, hi)); b = ocomp.minimizer._reduce_cost(True, (ohi - 5, ohi))
    assert abs(a - b) < 1e-9
    a = comp.minimizer._reduce_cost(False, (lo, hi)); b = ocomp.minimizer._reduce_cost(False, (olo, ohi))
    assert abs(a - b) < 1e-9


@pytest.mark.parametrize("fake_backend", [None, 6, (None, 8), (4, 8)], indirect=True)
@pytest.mark.parametrize("n", [4, 12])
@pytest.mark.parametrize("front", [True, False])
def test_incremental_evaluator_equals_full_resimulation(fake_backend, n, front, monkeypatch):
    monkeypatch.setattr(SVCostEvaluator, "front_mode", front)
    rng = np.random.default_rng(50 + n)
    target, trng = brickwork(n, 3, seed=n)
    ansatz = thin_ansatz(n, 5, trng)
    comp = AdaptCompiler(target, backend=fake_backend)
    comp.full_circuit.data.extend(ansatz.data)
    ocomp = AdaptCompiler(target, backend=OracleSVBackend())
    ocomp.full_circuit.data.extend(ansatz.copy().data)
    rot = [i for i in range(*comp.variational_circuit_range())
           if comp.full_circuit.data[i].operation.name in ("rx", "ry", "rz")]
    for step in range(40):
        idx = rot[int(rng.integers(len(rot)))] if step % 5 else rot[step % len(rot)]
        for rep in range(1 + step % 3):          # the optimiser asks for several values of ONE gate in a row
            name = ["rx", "ry", "rz"][int(rng.integers(3))]
            theta = float(rng.uniform(-np.pi, np.pi))
            for c in (comp, ocomp):
                replace_1q_gate(c.full_circuit, idx, name, theta)
            assert abs(comp.evaluate_cost() - ocomp.evaluate_cost()) < 1e-10
    st = fake_backend._evaluator.stats
    assert st["moves_R"] + st["rebuild_R"] + st.get("front_blocks", 0) + st.get("direct_projections", 0) > 0 and st["t_passes"] + st["t_gathers"] < st["evals"]
    if front and not fake_backend._evaluator.projected:
        assert st.get("front_blocks", 0) > 0
        if n >= 12:      # the bra's move and the transfer matrix came out of ONE pass (b200_sv_run_inner2)
            assert st.get("fused_T", 0) > 0
    elif fake_backend._evaluator.compact is not None and not fake_backend._evaluator.projected:
        assert st["t_gathers"] > 0 and st["compact_L"] > 0
    if fake_backend._evaluator.projected:
        assert st["projected_evals"] > 0 and st["projections"] > 0
        if n >= 12:      # bras whose tail was built on a smaller engine entered the register through an embedded-source sweep
            assert st.get("embedded_L", 0) >= st.get("scattered_L", 0) > 0      # (the tail's bra is kept between blocks)
            if front:    # ... and prefix|base> was never stored: the sweep kept the projected amplitudes only
                assert st.get("direct_projections", 0) > 0 and st["rebuild_R"] == 0


@pytest.mark.parametrize("fake_backend", [None, 5], indirect=True)
def test_interleaved_backend_calls_keep_the_caches_coherent(fake_backend):
    """evaluate_circuit / <Z> / shift_costs between cost evaluations consume circuit edits that the
    evaluator has not seen: the next cost must still equal a full re-simulation."""
    n = 7
    rng = np.random.default_rng(77)
    target, trng = brickwork(n, 2, seed=2)
    ansatz = thin_ansatz(n, 4, trng)
    comp = AdaptCompiler(target, backend=fake_backend)
    comp.full_circuit.data.extend(ansatz.data)
    ocomp = AdaptCompiler(target, backend=OracleSVBackend())
    ocomp.full_circuit.data.extend(ansatz.copy().data)
    rot = [i for i in range(*comp.variational_circuit_range())
           if comp.full_circuit.data[i].operation.name in ("rx", "ry", "rz")]
    for step in range(30):
        for _ in range(int(rng.integers(1, 4))):          # several edits, possibly in different layers
            idx = rot[int(rng.integers(len(rot)))]
            name, theta = ["rx", "ry", "rz"][int(rng.integers(3))], float(rng.uniform(-3, 3))
            for c in (comp, ocomp):
                replace_1q_gate(c.full_circuit, idx, name, theta)
        kind = step % 4
        if kind == 1:
            np.testing.assert_allclose(fake_backend.measure_qubit_expectation_values(comp),
                                       ocomp.backend.measure_qubit_expectation_values(ocomp), atol=1e-12)
            continue                                       # edits consumed without the evaluator seeing them
        if kind == 2:
            idx = rot[int(rng.integers(len(rot)))]
            got = fake_backend.shift_costs(comp, idx, [("ry", 0.3)])[0]
            replace_1q_gate(ocomp.full_circuit, idx, "ry", 0.3)
            ref = ocomp.evaluate_cost()
            replace_1q_gate(ocomp.full_circuit, idx, comp.full_circuit.data[idx].operation.name,
                            comp.full_circuit.data[idx].operation.params[0])
            assert abs(got - ref) < 1e-10
            continue
        assert abs(comp.evaluate_cost() - ocomp.evaluate_cost()) < 1e-10


def test_structure_change_falls_back_to_resimulation(fake_backend):
    n = 5
    target, trng = brickwork(n, 2, seed=9)
    comp = AdaptCompiler(target, backend=fake_backend)
    ocomp = AdaptCompiler(target, backend=OracleSVBackend())
    for layers in (1, 2, 3):
        ansatz = thin_ansatz(n, layers, np.random.default_rng(layers))
        for c in (comp, ocomp):
            del c.full_circuit.data[c.lhs_gate_count:]
            c.full_circuit.data.extend(ansatz.copy().data)
        assert abs(comp.evaluate_cost() - ocomp.evaluate_cost()) < 1e-12
        assert abs(comp.evaluate_cost() - ocomp.evaluate_cost()) < 1e-12   # unchanged circuit
    np.testing.assert_allclose(fake_backend.measure_qubit_expectation_values(comp),
                               ocomp.backend.measure_qubit_expectation_values(ocomp), atol=1e-12)


@pytest.mark.parametrize("fake_backend", [None, 2, (None, 2), (None, 3)], indirect=True)
@pytest.mark.parametrize("batched", [False, True])
def test_compile_decisions_match_oracle_backend(fake_backend, batched):
    ghz = Circuit(4); ghz.h(0)
    for i in range(3):
        ghz.cx(i, i + 1)
    rng = np.random.default_rng(3)
    rnd = circuit_from_gates(3, random_gates(3, 15, rng, allow_mat=False))
    # BASELINE config C1 (README.md:56-61): its Rotoselect steps contain exact ties between axes, so the pair
    # history is sensitive to the ORDER of the floating-point operations of an evaluation, not just its value
    readme = Circuit(3)
    readme.rx(1.23, 0); readme.cx(0, 1); readme.ry(2.5, 1); readme.rx(-1.6, 2); readme.ccx(2, 1, 0)
    for target in (ghz, rnd, readme):
        ref = AdaptCompiler(target, backend=OracleSVBackend(), adapt_config=AdaptConfig(max_layers=6)).compile()
        got = AdaptCompiler(target, backend=fake_backend, adapt_config=AdaptConfig(max_layers=6),
                            minimiser_cls=B200CostMinimiser if batched else None).compile()
        assert got.qubit_pair_history == ref.qubit_pair_history
        np.testing.assert_allclose(got.global_cost_history, ref.global_cost_history, atol=1e-9)
        assert got.cost_evaluations == ref.cost_evaluations
        assert abs(got.exact_overlap - got.overlap) < 1e-9


def test_bench_step_counts_220_evaluations(fake_backend):
    """bench.py's step: Rotoselect over the newest layer + one Rotosolve cycle (C3 shape, small n)."""
    import bench
    target, ansatz = bench.build_workload(6, 2, 16)
    for batched in (True, False):
        comp = bench.make_compiler(target, ansatz, fake_backend, batched)
        comp.evaluate_cost()
        before = comp.cost_evaluation_counter
        bench.one_step(comp)
        assert comp.cost_evaluation_counter - before == 4 * 7 + 64 * 3


@pytest.mark.parametrize("fake_backend", [(None, 6)], indirect=True)
def test_bench_step_device_work_is_the_same_through_both_front_ends(fake_backend):
    """The one-scalar-per-call reference interface and the batched minimiser must cost the device the same
    passes per step: one transfer pass per head block, bras prefetched or moved once, no thrash between a
    prefetched bra and the block that is still open (a regression once doubled the passes of one of them)."""
    import bench
    n = 10
    target, ansatz = bench.build_workload(n, 2, 16)
    per_mode = []
    for batched in (True, False):
        fake_backend.reset_cache()
        comp = bench.make_compiler(target, ansatz, fake_backend, batched)
        comp.evaluate_cost()
        bench.one_step(comp)                      # reach the steady state
        ev = fake_backend._evaluator
        s0 = dict(ev.stats)
        bench.one_step(comp)
        d = {k: ev.stats.get(k, 0) - s0.get(k, 0) for k in ev.stats}
        head_blocks = sum(1 for b in ev._blocks(ev.window or []) ) if ev.window else None
        per_mode.append(d)
        assert d["projected_evals"] > 0
        assert d["t_passes"] <= 16 and d["moves_L"] + d["rebuild_L"] + d.get("prefetched_L", 0) <= 17, d
    for k in ("t_passes", "moves_L", "rebuild_L"):
        assert abs(per_mode[0][k] - per_mode[1][k]) <= 1, (k, per_mode)


@pytest.mark.parametrize("fake_backend", [2, (2, 3)], indirect=True)
@pytest.mark.parametrize("case", compile_option_cases(), ids=lambda c: c[0])
def test_compile_options_make_the_same_decisions(fake_backend, case):
    name, target, kw, cfg = case
    ref = AdaptCompiler(target, backend=OracleSVBackend(), adapt_config=AdaptConfig(**cfg), **kw).compile()
    got = AdaptCompiler(target, backend=fake_backend, adapt_config=AdaptConfig(**cfg), **kw).compile()
    assert got.qubit_pair_history == ref.qubit_pair_history
    assert got.method_history == ref.method_history
    np.testing.assert_allclose(got.global_cost_history, ref.global_cost_history, atol=1e-9)
    if ref.local_cost_history is not None:
        np.testing.assert_allclose(got.local_cost_history, ref.local_cost_history, atol=1e-9)
    assert got.cost_evaluations == ref.cost_evaluations


def test_backend_shared_by_compilers_of_different_widths(fake_backend):
    """The reference's default-argument backends are singletons shared by every compiler
    (aer_sv_backend.py:20, adapt_compiler.py:59).  A wider compiler that is still alive must not leak
    its coupling map into the pair-RDM batch of a narrower one (entanglement_measures.py:71-75)."""
    wide_target, _ = brickwork(6, 2, seed=5)
    wide = AdaptCompiler(wide_target, backend=fake_backend)
    wide.evaluate_cost()
    qc = Circuit(3); qc.h(0); qc.cx(0, 1)
    narrow = AdaptCompiler(qc, backend=fake_backend)
    ems = narrow._get_all_qubit_pair_entanglement_measures()
    np.testing.assert_allclose(ems, [1.0, 0.0, 0.0], atol=1e-7)
    ref = AdaptCompiler(wide_target, backend=OracleSVBackend())._get_all_qubit_pair_entanglement_measures()
    np.testing.assert_allclose(wide._get_all_qubit_pair_entanglement_measures(), ref, atol=1e-9)


def test_mps_checkpoint_normal_form_is_the_same_circuit():
    """B200MPSSimulator._commute_phases_past_controls (checkpoint matching of the capped MPS path): every
    1-qubit gate as late as exact commutation allows.  Same unitary on random circuits; on two thinly dressed
    layers only the rotations on the cx TARGETS stay in front of a 2-qubit gate."""
    from adapt_aqc_b200.gates import canonical_window
    from adapt_aqc_b200.mps_backend import B200MPSSimulator
    from oracle import sv_oracle as orc
    norm = B200MPSSimulator._commute_phases_past_controls
    rng = np.random.default_rng(8)
    n = 5
    npar = {"rx": 1, "ry": 1, "rz": 1, "u1": 1, "p": 1, "u2": 2, "u3": 3, "u": 3}

    def to_gates(win):
        return [(e[0], [e[1]] + ([e[2]] if e[2] >= 0 else []), [e[3], e[4], e[5]][:npar.get(e[0], 0)]) for e in win]

    for trial in range(60):
        gates = []
        for _ in range(30):
            q = int(rng.integers(n)); q2 = int((q + 1 + rng.integers(n - 1)) % n)
            gates.append([("rz", [q], [float(rng.uniform(-3, 3))]), ("cx", [q, q2], []), ("cz", [q, q2], []), ("ry", [q], [0.3]),
                          ("t", [q], []), ("h", [q], []), ("u1", [q], [0.7]), ("swap", [q, q2], [])][int(rng.integers(8))])
        win = canonical_window(circuit_from_gates(n, gates))
        w2 = norm(win)
        assert sorted(map(repr, win)) == sorted(map(repr, w2))
        np.testing.assert_allclose(orc.evaluate_circuit(n, to_gates(w2)), orc.evaluate_circuit(n, to_gates(win)), atol=1e-12)
    a, b, c = 1, 2, 3
    thin = [("rz", a, -1, .1, 0, 0, None), ("rz", b, -1, .2, 0, 0, None), ("cx", a, b, 0, 0, 0, None), ("rz", a, -1, .3, 0, 0, None),
            ("rz", b, -1, .4, 0, 0, None), ("rz", b, -1, .5, 0, 0, None), ("rz", c, -1, .6, 0, 0, None), ("cx", b, c, 0, 0, 0, None),
            ("rz", b, -1, .7, 0, 0, None), ("rz", c, -1, .8, 0, 0, None)]
    w2 = norm(thin)
    last_2q = max(i for i, e in enumerate(w2) if e[2] >= 0)
    assert [e[3] for e in w2[:last_2q] if e[2] < 0] == [.2, .6]          # only the rotations on the cx targets


def test_resimulation_then_tail_edit_does_not_reuse_a_stale_projection(emu):
    """Sharded-register mode (dense_blocks=False): a head-gate edit is served by re-simulation, which leaves slot R /
    phi at the OLD head; the next tail-only edit must re-gather phi instead of taking the "only tail gates changed"
    shortcut (round-1 advisor finding: 2.7e-3 error in the order tail edit, head edit, tail edit)."""
    from adapt_aqc_b200.gates import canonical_window
    from oracle import sv_oracle as orc
    from oracle.oracle_backends import circuit_to_gates
    n = 12
    target, trng = brickwork(n, 3, seed=3)
    # head layers spread over the register (suffix touches > 8 qubits -> no compact / projected path), tail on 0..3
    ansatz = Circuit(n)
    pairs = [(0, 1), (4, 5), (8, 9), (10, 11), (2, 3), (6, 7), (1, 2), (0, 1), (2, 3)]
    for a, b in pairs:
        th = trng.uniform(-np.pi, np.pi, 4)
        ansatz.rz(th[0], a, label="rz"); ansatz.rz(th[1], b, label="rz"); ansatz.cx(a, b)
        ansatz.rz(th[2], a, label="rz"); ansatz.rz(th[3], b, label="rz")
    eng = FakeEngine(emu, n)
    ev = SVCostEvaluator(eng, None, [FakeEngine(emu, 8, n_slots=4)])
    ev.dense_blocks = False
    from adapt_aqc_b200.gates import GateStream
    ev.set_base("t", GateStream.from_circuit(target))
    base_gates = circuit_to_gates(target)

    def check(window, changed):
        got = ev.amp0(list(window), changed=changed)
        c = Circuit(n)
        c.data = list(ansatz.data)
        ref = orc.evaluate_circuit(n, base_gates + circuit_to_gates(c))[0]
        assert abs(got - ref) < 1e-10, (got, ref)

    window = canonical_window(ansatz)
    check(window, None)
    tail_idx, head_idx = len(window) - 1, 0
    for step, idx in enumerate([tail_idx, head_idx, tail_idx, tail_idx - 5, head_idx + 1, tail_idx]):
        replace_1q_gate(ansatz, idx, "rz", 0.37 * (step + 1))
        window[idx] = canonical_window(ansatz, idx, idx + 1)[0]
        check(window, [idx])
    assert ev.stats.get("resimulations", 0) >= 2 and ev.stats["projected_evals"] >= 3


@pytest.mark.parametrize("direct", [True, False])
def test_front_mode_head_edit_then_tail_edit_reprojects(emu, direct, monkeypatch):
    """Front mode serves head blocks from the base state and never touches slot R, so nothing on the ket side records
    that a head gate changed: the projection must be tied to the CONTENT of the prefix it was taken from.  Order: tail
    edit, head edit (evaluated, so the edit is no longer pending), tail edit -> the last one needs a fresh phi."""
    from adapt_aqc_b200.gates import GateStream, canonical_window
    from oracle import sv_oracle as orc
    from oracle.oracle_backends import circuit_to_gates
    monkeypatch.setattr(SVCostEvaluator, "project_direct", direct)
    n = 12
    target, trng = brickwork(n, 3, seed=5)
    ansatz = Circuit(n)
    pairs = [(0, 1), (2, 3), (8, 9), (10, 11), (4, 5), (6, 7), (5, 6), (4, 5), (6, 7), (9, 10)]     # tail: qubits 4..11
    for a, b in pairs:
        th = trng.uniform(-np.pi, np.pi, 4)
        ansatz.rz(th[0], a, label="rz"); ansatz.rz(th[1], b, label="rz"); ansatz.cx(a, b)
        ansatz.rz(th[2], a, label="rz"); ansatz.rz(th[3], b, label="rz")
    eng = FakeEngine(emu, n)
    ev = SVCostEvaluator(eng, None, [FakeEngine(emu, 8, n_slots=4)])
    ev.set_base("t", GateStream.from_circuit(target))
    base_gates = circuit_to_gates(target)

    def check(window, changed):
        got = ev.amp0(list(window), changed=changed)
        c = Circuit(n)
        c.data = list(ansatz.data)
        ref = orc.evaluate_circuit(n, base_gates + circuit_to_gates(c))[0]
        assert abs(got - ref) < 1e-10, (got, ref)

    window = canonical_window(ansatz)
    check(window, None)
    tail_idx, head_idx = len(window) - 1, 0
    order = [tail_idx, head_idx, tail_idx, head_idx + 6, tail_idx - 7, head_idx + 3, head_idx + 3, tail_idx, tail_idx]
    for step, idx in enumerate(order):
        replace_1q_gate(ansatz, idx, "rz", 0.41 * (step + 1))
        window[idx] = canonical_window(ansatz, idx, idx + 1)[0]
        check(window, [idx])
    st = ev.stats
    assert st.get("front_blocks", 0) >= 1 and st["projected_evals"] >= 4 and st["projections"] >= 3, st
    assert st.get("virtual_L", 0) >= 1, st         # head blocks: T from ONE read of the ket, the bra never written
    if direct:
        assert st.get("direct_projections", 0) >= 2 and st["rebuild_R"] == 0, st


@pytest.mark.parametrize("fake_backend", [(None, 8)], indirect=True)
def test_both_front_ends_issue_the_same_device_calls_per_step(fake_backend):
    """Registers large enough for the fused / embedded / projected-store passes (n = 12, tail engine of 8 qubits): the
    one-scalar-per-call interface must cost the device exactly what the batched front end costs -- the optimiser leaves two
    edits pending when it moves on, and opening the block of the WRONG one (cycle wrap-around, re-based nested level)
    once cost a projection and a bra rebuild per step."""
    import bench
    n = 12
    target, ansatz = bench.build_workload(n, 2, 16)
    per_mode = []
    for batched in (True, False):
        fake_backend.reset_cache()
        comp = bench.make_compiler(target, ansatz, fake_backend, batched)
        comp.evaluate_cost()
        for _ in range(2):
            bench.one_step(comp)                  # reach the steady state
        ev = fake_backend._evaluator
        engines = [ev.eng] + list(ev.projected)
        s0 = dict(ev.stats)
        c0 = [(e.runs, e.inners) for e in engines]
        bench.one_step(comp)
        d = {k: ev.stats.get(k, 0) - s0.get(k, 0) for k in ev.stats}
        calls = [(e.runs - r, e.inners - i) for e, (r, i) in zip(engines, c0)]
        per_mode.append((d, calls))
        assert d["projections"] == 1 and d.get("direct_projections", 0) == 1, d      # ONE projection per step
        assert d.get("virtual_L", 0) >= 1, d
    assert per_mode[0][1] == per_mode[1][1], per_mode
    for k in ("projections", "t_passes", "rebuild_L", "rebuild_R", "fused_T"):
        assert per_mode[0][0].get(k, 0) == per_mode[1][0].get(k, 0), (k, per_mode)


@pytest.mark.parametrize("switch", ["fused_passes", "lazy_bra", "front_mode", "project_direct"])
def test_every_evaluator_switch_off_gives_the_same_numbers(emu, switch, monkeypatch):
    """Each A/B switch of the evaluator selects a different sequence of device passes for the same
    amplitudes: with any one of them off, head / tail edits in every order still equal a full re-simulation."""
    from adapt_aqc_b200.gates import GateStream, canonical_window
    from oracle import sv_oracle as orc
    from oracle.oracle_backends import circuit_to_gates
    monkeypatch.setattr(SVCostEvaluator, switch, False)
    n = 12
    target, trng = brickwork(n, 3, seed=9)
    ansatz = Circuit(n)
    pairs = [(0, 1), (2, 3), (8, 9), (10, 11), (4, 5), (6, 7), (5, 6), (4, 5), (6, 7), (9, 10)]     # tail: qubits 4..11
    for a, b in pairs:
        th = trng.uniform(-np.pi, np.pi, 4)
        ansatz.rz(th[0], a, label="rz"); ansatz.rz(th[1], b, label="rz"); ansatz.cx(a, b)
        ansatz.rz(th[2], a, label="rz"); ansatz.rz(th[3], b, label="rz")
    ev = SVCostEvaluator(FakeEngine(emu, n), None, [FakeEngine(emu, 8, n_slots=4)])
    ev.set_base("t", GateStream.from_circuit(target))
    base_gates = circuit_to_gates(target)
    window = canonical_window(ansatz)
    rng = np.random.default_rng(17)
    rot = [i for i, e in enumerate(window) if e[2] < 0]
    order = [None] + [rot[int(rng.integers(len(rot)))] for _ in range(24)] + rot[:8] + rot[-8:]
    for step, idx in enumerate(order):
        if idx is not None:
            replace_1q_gate(ansatz, idx, ["rx", "ry", "rz"][step % 3], float(rng.uniform(-np.pi, np.pi)))
            window[idx] = canonical_window(ansatz, idx, idx + 1)[0]
        got = ev.amp0(list(window), changed=None if idx is None else [idx])
        c = Circuit(n)
        c.data = list(ansatz.data)
        ref = orc.evaluate_circuit(n, base_gates + circuit_to_gates(c))[0]
        assert abs(got - ref) < 1e-10, (switch, step, idx, got, ref)
