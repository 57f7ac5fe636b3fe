"""GPU parity tests of the statevector path: everything goes through the C-ABI of libb200aqc.so
(ctypes, adapt_aqc_b200.lib) and is compared with the CPU oracle on the same seeded inputs.

Tolerances: BASELINE north_star asks for overlap/cost within 1e-10 absolute (complex128); the
amplitude-level checks here use 1e-12.  Pair selection must be identical."""
import ctypes
import pickle

import numpy as np
import pytest

from adapt_aqc_b200 import lib as blib
from harness import measures as em
from adapt_aqc_b200.backends import B200SVBackend
from harness.circuit import Circuit
from harness.compiler import (CMAP_LINEAR, AdaptCompiler, AdaptConfig, generate_coupling_map)
from adapt_aqc_b200.gates import GateStream
from harness.minimiser import B200CostMinimiser
from adapt_aqc_b200.sv_engine import SLOT_BASE, SLOT_L, SLOT_R, SLOT_WORK, SVCostEvaluator, SVEngine, plan_detail
from oracle import sv_oracle as orc
from oracle.oracle_backends import OracleSVBackend, circuit_to_gates

from helpers import brickwork, circuit_from_gates, compile_option_cases, random_gates, thin_ansatz

pytestmark = pytest.mark.gpu

AMP_TOL = 1e-12
COST_TOL = 1e-10


@pytest.fixture(scope="module")
def backend():
    return B200SVBackend()


# ---- b200_sv_run / run_inverse ------------------------------------------------------------------
@pytest.fixture(params=["direct", "pipe"])
def sweep_kernel(request, monkeypatch):
    """Both sweep kernels (read at context creation): the direct-load default and the pipelined
    bulk-copy (TMA) variant -- same planner output, same per-thread op bodies."""
    monkeypatch.setenv("B200AQC_SWEEP", request.param)
    return request.param


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 11, 12, 13, 14, 17, 20, 22])
def test_run_matches_oracle_all_opcodes(n, sweep_kernel):
    if sweep_kernel == "pipe" and n <= 11:
        pytest.skip("n <= 11 runs in the single-CTA kernel whatever the sweep kernel")
    rng = np.random.default_rng(4000 + n)
    eng = SVEngine(n, n_slots=2)
    for trial in range(3):
        gates = random_gates(n, int(rng.integers(1, 120)), rng)
        ref = orc.evaluate_circuit(n, gates)
        eng.run(0, -1, GateStream.from_gates(gates))
        np.testing.assert_allclose(eng.download(0), ref, atol=AMP_TOL)
    eng.close()


@pytest.mark.parametrize("n", [4, 12, 16, 21])
def test_run_from_slot_and_inverse_round_trip(n, sweep_kernel):
    if sweep_kernel == "pipe" and n <= 11:
        pytest.skip("n <= 11 runs in the single-CTA kernel whatever the sweep kernel")
    rng = np.random.default_rng(4100 + n)
    eng = SVEngine(n, n_slots=3)
    g1 = random_gates(n, 60, rng)
    g2 = random_gates(n, 60, rng)
    eng.run(0, -1, GateStream.from_gates(g1))
    eng.run(1, 0, GateStream.from_gates(g2))       # out of place
    ref1 = orc.evaluate_circuit(n, g1)
    ref12 = orc.apply_gates(ref1, g2)
    np.testing.assert_allclose(eng.download(0), ref1, atol=AMP_TOL)
    np.testing.assert_allclose(eng.download(1), ref12, atol=AMP_TOL)
    eng.run(1, 1, GateStream.from_gates(g2), inverse=True)   # in place, inverse
    np.testing.assert_allclose(eng.download(1), ref1, atol=AMP_TOL)
    eng.run(1, 1, GateStream.from_gates(g1), inverse=True)
    e0 = np.zeros(1 << n, dtype=np.complex128); e0[0] = 1
    np.testing.assert_allclose(eng.download(1), e0, atol=AMP_TOL)
    assert abs(eng.amp(1, 0) - 1) < AMP_TOL
    eng.close()


def test_empty_stream_copy_upload_download():
    n = 13
    eng = SVEngine(n, n_slots=2)
    eng.run(0, -1, GateStream.from_gates([]))
    psi = eng.download(0)
    assert psi[0] == 1 and np.count_nonzero(psi) == 1
    rng = np.random.default_rng(1)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    eng.upload(1, v)
    eng.run(0, 1, GateStream.from_gates([]))    # empty stream = copy
    np.testing.assert_array_equal(eng.download(0), v)
    np.testing.assert_array_equal(eng.download(0, 100, 50), v[100:150])
    eng.copy(1, 0)
    assert eng.amp(1, 77) == v[77]
    eng.close()


# ---- read-outs ----------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 5, 11, 12, 13, 16, 19, 21])
def test_expz_matches_oracle(n):
    rng = np.random.default_rng(4200 + n)
    gates = random_gates(n, 60, rng)
    ref = orc.evaluate_circuit(n, gates)
    eng = SVEngine(n, n_slots=1)
    eng.run(0, -1, GateStream.from_gates(gates))
    z, norm = eng.expz(0)
    np.testing.assert_allclose(z, orc.measure_qubit_expectation_values(ref), atol=AMP_TOL)
    assert abs(norm - 1) < AMP_TOL
    eng.close()


@pytest.mark.parametrize("n", [2, 3, 4, 6, 12, 15, 18])
def test_pair_rdm_matches_oracle_for_every_pair(n):
    rng = np.random.default_rng(4300 + n)
    gates = random_gates(n, 80, rng)
    ref = orc.evaluate_circuit(n, gates)
    eng = SVEngine(n, n_slots=1)
    eng.run(0, -1, GateStream.from_gates(gates))
    pairs = [(a, b) for a in range(n) for b in range(n) if a != b]
    if n > 8:
        pairs = [pairs[i] for i in rng.choice(len(pairs), 40, replace=False)]
    rho = eng.pair_rdm(0, pairs)
    for r, (a, b) in zip(rho, pairs):
        if n == 2:
            lo_first = ref if a < b else ref  # 2 qubits: RDM of the whole (pure) state
            expect = np.outer(lo_first, lo_first.conj())
        else:
            expect = orc.partial_trace(ref, a, b)
        np.testing.assert_allclose(r, expect, atol=AMP_TOL)
        assert abs(np.trace(r) - 1) < AMP_TOL
    eng.close()


@pytest.mark.parametrize("n", [3, 12, 17])
def test_inner_matches_numpy(n):
    rng = np.random.default_rng(4400 + n)
    eng = SVEngine(n, n_slots=2)
    L = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    R = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    L /= np.linalg.norm(L); R /= np.linalg.norm(R)
    eng.upload(0, L); eng.upload(1, R)
    assert abs(eng.inner(0, 1, -1) - np.vdot(L, R)) < AMP_TOL
    for q in sorted({0, 1, n // 2, n - 1}):
        Lt = np.moveaxis(L.reshape([2] * n), n - 1 - q, 0).reshape(2, -1)
        Rt = np.moveaxis(R.reshape([2] * n), n - 1 - q, 0).reshape(2, -1)
        np.testing.assert_allclose(eng.inner(0, 1, q), Lt.conj() @ Rt.T, atol=AMP_TOL)
    for qa, qb in [(0, 1), (1, 0), (0, n - 1), (n - 1, n // 2), (n // 2, 1)]:
        if qa == qb:
            continue
        Lt = np.moveaxis(L.reshape([2] * n), [n - 1 - qb, n - 1 - qa], [0, 1]).reshape(4, -1)
        Rt = np.moveaxis(R.reshape([2] * n), [n - 1 - qb, n - 1 - qa], [0, 1]).reshape(4, -1)
        np.testing.assert_allclose(eng.inner2(0, 1, qa, qb), Lt.conj() @ Rt.T, atol=AMP_TOL)
    eng.close()


def test_inner2_gather_matches_dense_transfer():
    """Compact bra (K qubits of a second context) against the dense 4x4 transfer matrix."""
    n, K = 14, 6
    rng = np.random.default_rng(4450)
    eng, small = SVEngine(n, n_slots=2), SVEngine(K, n_slots=1)
    R = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    ell = rng.normal(size=1 << K) + 1j * rng.normal(size=1 << K)
    R /= np.linalg.norm(R); ell /= np.linalg.norm(ell)
    for qmap in ([0, 1, 2, 3, 4, 5], [13, 2, 7, 0, 9, 4], [5, 6, 8, 10, 11, 12]):
        c = np.arange(1 << K)
        x = np.zeros_like(c)
        for b, q in enumerate(qmap):
            x |= ((c >> b) & 1) << q
        L = np.zeros(1 << n, dtype=np.complex128); L[x] = ell
        eng.upload(0, L); eng.upload(1, R); small.upload(0, ell)
        small.sync()
        for qa, qb in [(qmap[0], qmap[1]), (qmap[3], qmap[1]), (qmap[5], qmap[2])]:
            dense = eng.inner2(0, 1, qa, qb)
            got = eng.inner2_gather(1, small, 0, qmap, qa, qb)
            np.testing.assert_allclose(got, dense, atol=AMP_TOL)
    eng.close(); small.close()


@pytest.mark.parametrize("n", [12, 14, 17, 22])
def test_fused_sweep_and_transfer_pass_equals_the_two_calls(n):
    """b200_sv_run_inner2 = b200_sv_run followed by b200_sv_inner2 (bra = the swept state): the swept state is bit-identical
    to the plain sweep's up to the folded-permutation rounding (1e-14), T matches the separate transfer pass and the oracle's
    numpy restatement, for pairs on lane / register / padding / untouched qubits, forward and inverse, empty, thin-layer and
    multi-sweep programs, in place and out of place."""
    rng = np.random.default_rng(8800 + n)
    dim = 1 << n
    eng = SVEngine(n, n_slots=4)
    for trial in range(8):
        psi0 = rng.normal(size=dim) + 1j * rng.normal(size=dim); psi0 /= np.linalg.norm(psi0)
        other = rng.normal(size=dim) + 1j * rng.normal(size=dim); other /= np.linalg.norm(other)
        if trial == 0:
            gates = []
        elif trial == 1:
            _, trng = brickwork(n, 1, seed=trial)
            gates = circuit_to_gates(thin_ansatz(n, 6, trng))
        else:
            gates = random_gates(n, int(rng.integers(1, 120 if trial < 6 else 400)), rng)
        qa, qb = [int(q) for q in rng.choice(n, size=2, replace=False)]
        if trial == 2:
            qa, qb = 0, 1
        if trial == 3:
            qa, qb = n - 1, 2
        inverse = bool(trial % 2)
        gs = GateStream.from_gates(gates)
        eng.upload(0, psi0); eng.upload(1, other)
        eng.run(2, 0, gs, inverse=inverse)
        ref_state = eng.download(2)
        ref_T = eng.inner2(2, 1, qa, qb)
        dst = 0 if trial % 3 == 0 else 3                 # in place / out of place
        T = eng.run_inner2(dst, 0, gs, 1, qa, qb, inverse=inverse)
        np.testing.assert_allclose(eng.download(dst), ref_state, rtol=0, atol=1e-14)
        np.testing.assert_allclose(T, ref_T, rtol=0, atol=1e-13)
        if n <= 17:
            Lt = np.moveaxis(ref_state.reshape([2] * n), [n - 1 - qb, n - 1 - qa], [0, 1]).reshape(4, -1)
            Rt = np.moveaxis(other.reshape([2] * n), [n - 1 - qb, n - 1 - qa], [0, 1]).reshape(4, -1)
            np.testing.assert_allclose(T, Lt.conj() @ Rt.T, rtol=0, atol=1e-12)
        np.testing.assert_array_equal(eng.download(1), other)     # `other` is only read
        # T only: one-sweep programs leave the destination alone (and say so), longer ones store as before
        eng.upload(0, psi0)
        T2, stored = eng.run_inner2(0, 0, gs, 1, qa, qb, inverse=inverse, store=False)
        np.testing.assert_allclose(T2, ref_T, rtol=0, atol=1e-13)
        if stored:
            np.testing.assert_allclose(eng.download(0), ref_state, rtol=0, atol=1e-14)
        else:
            np.testing.assert_array_equal(eng.download(0), psi0)
        if not gates or len(gates) < 20:
            assert not stored                      # short programs are one sweep
        if len(plan_detail(n, gs)) > 1:
            assert stored                          # (keeping the pair in every tile can only add sweeps)
    # |0..0> as the source (src = -1), as in b200_sv_run
    gs = GateStream.from_gates(random_gates(n, 60, rng))
    eng.upload(1, other)
    eng.run(2, -1, gs, inverse=True)
    ref_T = eng.inner2(2, 1, 3, n - 2)
    T = eng.run_inner2(0, -1, gs, 1, 3, n - 2, inverse=True)
    np.testing.assert_allclose(eng.download(0), eng.download(2), rtol=0, atol=1e-14)
    np.testing.assert_allclose(T, ref_T, rtol=0, atol=1e-13)
    with pytest.raises(blib.B200Error, match="destination"):
        eng.run_inner2(1, 0, GateStream.from_gates([]), 1, 0, 1)
    eng.close()
    small = SVEngine(8, n_slots=3)
    with pytest.raises(blib.B200Error, match="too small"):
        small.run_inner2(0, 0, GateStream.from_gates([]), 1, 0, 1)
    small.close()


@pytest.mark.parametrize("n", [12, 15, 22])
def test_embedded_source_sweep_equals_scatter_then_run(n):
    """b200_sv_run_embedded / b200_sv_run_embedded_inner2 = b200_sv_scatter + b200_sv_run (+ b200_sv_inner2): sorted and
    permuted qmaps, K from 2 to n, forward and inverse, empty program for the fused variant."""
    rng = np.random.default_rng(8900 + n)
    dim = 1 << n
    eng = SVEngine(n, n_slots=3)
    for trial in range(6):
        K = [2, 5, n - 1, n, 7, 9][trial]
        qmap = [int(q) for q in (rng.permutation(n)[:K] if trial % 2 else np.sort(rng.permutation(n)[:K]))]
        small = SVEngine(K, n_slots=1)
        phi = rng.normal(size=1 << K) + 1j * rng.normal(size=1 << K); phi /= np.linalg.norm(phi)
        other = rng.normal(size=dim) + 1j * rng.normal(size=dim); other /= np.linalg.norm(other)
        small.upload(0, phi); eng.upload(1, other)
        gates = random_gates(n, int(rng.integers(1, 100)), rng) if trial else []
        gs = GateStream.from_gates(gates)
        inverse = bool(trial % 2)
        eng.scatter(2, qmap, small, 0)
        eng.run(2, 2, gs, inverse=inverse)
        ref = eng.download(2)
        qa, qb = [int(q) for q in rng.choice(n, size=2, replace=False)]
        ref_T = eng.inner2(2, 1, qa, qb)
        assert eng.run_embedded(0, qmap, small, 0, gs, inverse=inverse) is None
        np.testing.assert_allclose(eng.download(0), ref, rtol=0, atol=1e-14)
        eng.init_zero(0)
        T = eng.run_embedded(0, qmap, small, 0, gs, inverse=inverse, fuse=(1, qa, qb))
        np.testing.assert_allclose(eng.download(0), ref, rtol=0, atol=1e-14)
        np.testing.assert_allclose(T, ref_T, rtol=0, atol=1e-13)
        # T only: the bra is never written when one sweep carries the program (one read of `other` is the whole pass)
        marker = rng.normal(size=dim) + 1j * rng.normal(size=dim)
        eng.upload(0, marker)
        T2, stored = eng.run_embedded(0, qmap, small, 0, gs, inverse=inverse, fuse=(1, qa, qb), store=False)
        np.testing.assert_allclose(T2, ref_T, rtol=0, atol=1e-13)
        if stored:
            np.testing.assert_allclose(eng.download(0), ref, rtol=0, atol=1e-14)
        else:
            np.testing.assert_array_equal(eng.download(0), marker)
        if not gates:
            assert not stored
        small.close()
    eng.close()


@pytest.mark.parametrize("n", [12, 15, 22])
def test_projected_store_sweep_equals_run_then_gather(n):
    """b200_sv_run_project = b200_sv_run + b200_sv_gather without storing the swept state: same 2^K amplitudes, the source
    untouched, the scratch slot untouched unless the program needs several sweeps."""
    rng = np.random.default_rng(9000 + n)
    dim = 1 << n
    eng = SVEngine(n, n_slots=3)
    for trial in range(7):
        K = [2, 5, n - 1, n, 7, 9, 8][trial]
        qmap = [int(q) for q in (rng.permutation(n)[:K] if trial % 2 else np.sort(rng.permutation(n)[:K]))]
        small = SVEngine(K, n_slots=2)
        psi0 = rng.normal(size=dim) + 1j * rng.normal(size=dim); psi0 /= np.linalg.norm(psi0)
        marker = rng.normal(size=dim) + 1j * rng.normal(size=dim)
        gates = [] if trial == 0 else random_gates(n, int(rng.integers(1, 100 if trial < 6 else 400)), rng)
        gs = GateStream.from_gates(gates)
        inverse = bool(trial % 2)
        eng.upload(0, psi0); eng.upload(1, marker)
        eng.run(2, 0, gs, inverse=inverse)
        eng.gather(2, qmap, small, 0)
        ref = small.download(0)
        used = eng.run_project(1, 0, gs, qmap, small, 1, inverse=inverse)
        np.testing.assert_allclose(small.download(1), ref, rtol=0, atol=1e-14)
        np.testing.assert_array_equal(eng.download(0), psi0)
        if not used:
            np.testing.assert_array_equal(eng.download(1), marker)
        if not gates:
            assert not used
        small.close()
    eng.close()


# ---- error behaviour ----------------------------------------------------------------------------
def test_errors_are_raised_not_swallowed():
    eng = SVEngine(4, n_slots=2)
    with pytest.raises(blib.B200Error, match="slot"):
        eng.init_zero(5)
    with pytest.raises(blib.B200Error, match="out of range"):
        eng.run(0, -1, GateStream.from_gates([("h", [9], [])]))
    with pytest.raises(blib.B200Error, match="out of range"):
        eng.amp(0, 1 << 4)
    with pytest.raises(blib.B200Error):
        eng.pair_rdm(0, [(1, 1)])
    eng.close()
    with pytest.raises(blib.B200Error):
        SVEngine(3, device=99)


def test_counters_and_timing():
    n = 16
    eng = SVEngine(n, n_slots=1)
    eng.set_timing(True)
    target, _ = brickwork(n, 4, 1)
    before = eng.counters()
    eng.run(0, -1, GateStream.from_circuit(target))
    ms = eng.last_ms()
    after = eng.counters()
    assert ms > 0
    assert after["sweeps"] > before["sweeps"] and after["launches"] > before["launches"]
    # algorithmic bytes: 32 * 2^n per sweep (read + write), except that the first sweep of a run from
    # |0..0> has no read pass (the kernel synthesises the source)
    assert after["bytes"] - before["bytes"] == ((after["sweeps"] - before["sweeps"]) * 32 - 16) * (1 << n)
    mid = eng.counters()
    eng.run(0, 0, GateStream.from_circuit(target))
    end = eng.counters()
    assert end["bytes"] - mid["bytes"] == (end["sweeps"] - mid["sweeps"]) * 32 * (1 << n)
    eng.close()


# ---- the reference's known-answer tests through the B200 backend ---------------------------------
def test_analytic_costs_of_simple_states(backend):
    """test/recompilers/test_approximate_compiler.py:114-150."""
    analytic = [0, 0, 1, 1 / 2, 1 / 2, 1 / 2, 15 / 16, 1 / 2]
    zero = Circuit(4)
    neel = Circuit(4); neel.x([0, 2])
    ghz = Circuit(4); ghz.h(0)
    for i in range(3):
        ghz.cx(0, i + 1)
    had = Circuit(4); had.h([0, 1, 2, 3])
    costs = []
    for circuit in [zero, neel, ghz, had]:
        for local in [False, True]:
            costs.append(AdaptCompiler(circuit, backend=backend, optimise_local_cost=local).evaluate_cost())
    np.testing.assert_allclose(costs, analytic, atol=1e-14)


def test_sigma_z_and_ghz_kats(backend):
    """test/utils/test_utilityfunctions.py:86-95; test_circuit_operations_running.py:48-61."""
    qc = Circuit(3); qc.x(0); qc.h(1)
    comp = AdaptCompiler(qc, backend=backend)
    np.testing.assert_array_almost_equal(backend.measure_qubit_expectation_values(comp), [-1.0, 0.0, 1.0], 15)
    ghz = Circuit(5); ghz.h(0)
    for i in range(4):
        ghz.cx(i, i + 1)
    sv = backend.simulator.run(ghz).result().get_statevector()
    expect = np.zeros(32, dtype=np.complex128); expect[0] = expect[31] = 1 / np.sqrt(2)
    np.testing.assert_allclose(np.asarray(sv), expect, atol=1e-15)
    assert len(sv) == 32 and sv.num_qubits == 5
    np.testing.assert_allclose(sv.probabilities([2]), [0.5, 0.5], atol=1e-15)


def test_soften_global_cost_raises_like_the_reference(backend):
    """aer_sv_backend.py:24-27"""
    qc = Circuit(2); qc.h(0)
    comp = AdaptCompiler(qc, backend=backend, soften_global_cost=True)
    with pytest.raises(NotImplementedError):
        comp.evaluate_cost()


# ---- incremental evaluator == full re-simulation ------------------------------------------------
@pytest.mark.parametrize("n", [4, 12, 15])
def test_evaluator_tracks_rotosolve_edits(n):
    """Every cost the incremental evaluator serves equals the oracle's full re-simulation of the
    edited circuit (what AerSVBackend does on each call, aer_sv_backend.py:37-47)."""
    rng = np.random.default_rng(4500 + n)
    target, trng = brickwork(n, 3, seed=n)
    ansatz = thin_ansatz(n, 6, trng)
    backend = B200SVBackend()
    comp = AdaptCompiler(target, backend=backend)
    comp.full_circuit.data.extend(ansatz.data)
    oracle_comp = AdaptCompiler(target, backend=OracleSVBackend())
    oracle_comp.full_circuit.data.extend([d for d in ansatz.copy().data])
    rot = [i for i in range(*comp.variational_circuit_range())
           if comp.full_circuit.data[i].operation.name in ("rx", "ry", "rz")]
    from harness.minimiser import replace_1q_gate
    for step in range(60):
        idx = rot[int(rng.integers(len(rot)))] if step % 7 else rot[step % len(rot)]
        name = ["rx", "ry", "rz"][int(rng.integers(3))]
        theta = float(rng.uniform(-np.pi, np.pi))
        for c in (comp, oracle_comp):
            replace_1q_gate(c.full_circuit, idx, name, theta)
        assert abs(comp.evaluate_cost() - oracle_comp.evaluate_cost()) < COST_TOL
    st = backend._evaluator.stats
    assert st["moves_R"] + st.get("front_blocks", 0) + st.get("direct_projections", 0) > 0 and st["t_passes"] + st["t_gathers"] < st["evals"]


def test_shift_costs_equal_individual_evaluations(backend):
    n = 13
    target, trng = brickwork(n, 2, seed=3)
    ansatz = thin_ansatz(n, 4, trng)
    comp = AdaptCompiler(target, backend=backend)
    comp.full_circuit.data.extend(ansatz.data)
    ocomp = AdaptCompiler(target, backend=OracleSVBackend())
    ocomp.full_circuit.data.extend(ansatz.copy().data)
    from harness.minimiser import replace_1q_gate
    idx = comp.variational_circuit_range()[0] + 8
    h = np.pi / 2
    cands = [("rx", 0.0)] + [(g, s) for g in ("rx", "ry", "rz") for s in (h, -h)]
    got = backend.shift_costs(comp, idx, cands)
    for (name, theta), c in zip(cands, got):
        replace_1q_gate(ocomp.full_circuit, idx, name, theta)
        assert abs(c - ocomp.evaluate_cost()) < COST_TOL


# ---- whole compile loop: identical decisions ----------------------------------------------------
def _targets():
    readme = Circuit(3)
    readme.rx(1.23, 0); readme.cx(0, 1); readme.ry(2.5, 1); readme.rx(-1.6, 2); readme.ccx(2, 1, 0)
    ghz = Circuit(5); ghz.h(0)
    for i in range(4):
        ghz.cx(i, i + 1)
    rng = np.random.default_rng(11)
    rnd4 = circuit_from_gates(4, random_gates(4, 25, rng, allow_mat=False))
    return {"readme3": readme, "ghz5": ghz, "random4": rnd4}


@pytest.mark.parametrize("name", ["readme3", "ghz5", "random4"])
@pytest.mark.parametrize("batched", [False, True])
def test_compile_makes_the_same_decisions_as_the_oracle_backend(name, batched):
    """BASELINE north_star: identical chosen qubit pairs, ansatz structure and layer count;
    costs within 1e-10."""
    target = _targets()[name]
    cfg = dict(max_layers=12)
    ref = AdaptCompiler(target, backend=OracleSVBackend(), adapt_config=AdaptConfig(**cfg)).compile()
    comp = AdaptCompiler(target, backend=B200SVBackend(), adapt_config=AdaptConfig(**cfg),
                         minimiser_cls=B200CostMinimiser if batched else None)
    got = comp.compile()
    assert got.qubit_pair_history == ref.qubit_pair_history
    assert got.method_history == ref.method_history
    assert len(got.global_cost_history) == len(ref.global_cost_history)
    np.testing.assert_allclose(got.global_cost_history, ref.global_cost_history, atol=1e-9)
    two_q = lambda res: [(i.operation.name, i.qubits) for i in res.circuit.data if len(i.qubits) == 2]
    assert two_q(got) == two_q(ref)
    assert abs(got.overlap - ref.overlap) < 1e-9
    if name == "readme3":
        # ghz5 / random4 compile to overlap 1 exactly; at an exact optimum some rotations are pure
        # gauge (flat cost: e.g. rz on a |0> qubit) and several Rotoselect axes tie to the last bit,
        # so roundoff -- not the backend -- picks among equivalent 1-qubit gates.
        assert [i.operation.name for i in got.circuit.data] == [i.operation.name for i in ref.circuit.data]
        assert [i.qubits for i in got.circuit.data] == [i.qubits for i in ref.circuit.data]
        assert got.cost_evaluations == ref.cost_evaluations
    # independent check of the answer: |<target|compiled>|^2 on the device
    assert abs(got.exact_overlap - got.overlap) < 1e-9
    if name != "random4":
        assert got.overlap > 1 - 1e-2


@pytest.mark.parametrize("case", compile_option_cases(), ids=lambda c: c[0])
def test_compile_options_make_the_same_decisions(case):
    """Every compile option of the reference that changes what the backend is asked for."""
    name, target, kw, cfg = case
    ref = AdaptCompiler(target, backend=OracleSVBackend(), adapt_config=AdaptConfig(**cfg), **kw).compile()
    got = AdaptCompiler(target, backend=B200SVBackend(), adapt_config=AdaptConfig(**cfg), **kw).compile()
    assert got.qubit_pair_history == ref.qubit_pair_history
    assert got.method_history == ref.method_history
    np.testing.assert_allclose(got.global_cost_history, ref.global_cost_history, atol=1e-9)
    if ref.local_cost_history is not None:
        np.testing.assert_allclose(got.local_cost_history, ref.local_cost_history, atol=1e-9)
    assert got.cost_evaluations == ref.cost_evaluations


def test_compile_12_qubits_linear_map_matches_oracle():
    """Tiled kernels inside the full loop (n > 11), local cost + expectation method included."""
    n = 12
    target, _ = brickwork(n, 2, seed=5)
    cmap = generate_coupling_map(n, CMAP_LINEAR)
    for kw in (dict(method="ISL"), dict(method="expectation")):
        cfg = dict(max_layers=4, **kw)
        ref = AdaptCompiler(target, backend=OracleSVBackend(), coupling_map=cmap,
                            adapt_config=AdaptConfig(**cfg)).compile()
        got = AdaptCompiler(target, backend=B200SVBackend(), coupling_map=cmap,
                            adapt_config=AdaptConfig(**cfg)).compile()
        assert got.qubit_pair_history == ref.qubit_pair_history
        np.testing.assert_allclose(got.global_cost_history, ref.global_cost_history, atol=1e-9)


def test_entanglement_measures_through_the_facade(backend):
    """run_circuit_without_transpilation -> partial_trace -> concurrence per pair
    (entanglement_measures.py:71-75); Bell pair gives [1,0,0]."""
    qc = Circuit(3); qc.h(0); qc.cx(0, 1)
    comp = AdaptCompiler(qc, backend=backend)
    ems = comp._get_all_qubit_pair_entanglement_measures()
    np.testing.assert_allclose(ems, [1.0, 0.0, 0.0], atol=1e-7)
    for method in (em.EM_TOMOGRAPHY_EOF, em.EM_TOMOGRAPHY_NEGATIVITY, em.EM_TOMOGRAPHY_LOG_NEGATIVITY):
        comp.entanglement_measure_method = method
        o = AdaptCompiler(qc, backend=OracleSVBackend(), entanglement_measure=method)
        np.testing.assert_allclose(comp._get_all_qubit_pair_entanglement_measures(),
                                   o._get_all_qubit_pair_entanglement_measures(), atol=1e-9)


def test_backend_is_picklable_for_checkpointing(tmp_path):
    """adapt_compiler.py:484-506 pickles the whole compiler, backend included."""
    ghz = _targets()["ghz5"]
    comp = AdaptCompiler(ghz, backend=B200SVBackend(), adapt_config=AdaptConfig(max_layers=3))
    comp.compile(checkpoint_every=1, checkpoint_dir=str(tmp_path))
    with open(tmp_path / "1.pkl", "rb") as f:
        resumed = pickle.load(f)
    assert resumed.resume_from_layer == 2
    resumed.adapt_config.max_layers = 12
    res = resumed.compile()
    ref = AdaptCompiler(ghz, backend=OracleSVBackend(), adapt_config=AdaptConfig(max_layers=12)).compile()
    assert res.qubit_pair_history == ref.qubit_pair_history
    assert abs(res.overlap - ref.overlap) < 1e-9


# ---- full-size properties (BASELINE config C3: 28 qubits, 4 GiB per state) -----------------------
def test_c3_size_properties():
    """At 2^28 amplitudes the oracle is too slow for routine tests; check size-independent
    properties instead: U^+ U |0> = |0>, norm preservation, <Z> of a product layer, and the
    cost of [target | target^+] = 0."""
    n = 28
    target, trng = brickwork(n, 8, seed=1234)
    eng = SVEngine(n, n_slots=2)
    gs = GateStream.from_circuit(target)
    eng.run(0, -1, gs)
    z, norm = eng.expz(0)
    assert abs(norm - 1) < 1e-10
    assert np.all(np.abs(z) <= 1 + 1e-12)
    eng.run(1, 0, gs, inverse=True)
    a0 = eng.amp(1, 0)
    assert abs(a0 - 1) < 1e-10
    z1, norm1 = eng.expz(1)
    np.testing.assert_allclose(z1, np.ones(n), atol=1e-10)
    assert abs(eng.inner(0, 0, -1) - 1) < 1e-10
    # one layer of ry(theta_q): <Z_q> = cos(theta_q) exactly
    th = trng.uniform(-np.pi, np.pi, n)
    eng.run(1, -1, GateStream.from_gates([("ry", [q], [th[q]]) for q in range(n)]))
    zq, _ = eng.expz(1)
    np.testing.assert_allclose(zq, np.cos(th), atol=1e-12)
    # Bell pairs on (q, q+14): pair RDM is the Bell projector, others are product states
    gates = []
    for q in range(0, 14):
        gates += [("h", [q], []), ("cx", [q, q + 14], [])]
    eng.run(1, -1, GateStream.from_gates(gates))
    rho = eng.pair_rdm(1, [(0, 14), (13, 27), (0, 1)])
    bell = np.zeros((4, 4)); bell[0, 0] = bell[0, 3] = bell[3, 0] = bell[3, 3] = 0.5
    np.testing.assert_allclose(rho[0], bell, atol=1e-12)
    np.testing.assert_allclose(rho[1], bell, atol=1e-12)
    np.testing.assert_allclose(rho[2], np.eye(4) / 4, atol=1e-12)
    eng.close()


def test_c3_evaluator_equals_device_resimulation():
    """BASELINE config C3 at full size: the incremental evaluator (block transfer matrices, compact bra,
    projected tail on the 16/20/24-qubit contexts) against a from-scratch simulation of the whole
    circuit on the device -- itself parity-tested against the oracle at the sizes the oracle can run.
    Edits walk over head layers (dense path) and tail layers (projected path) and back."""
    n = 28
    _, trng = brickwork(n, 1, seed=1234)
    layers = thin_ansatz(n, 16, trng)
    target = Circuit(n)
    for q in range(n):
        target.ry(float(trng.uniform(0.5, 2.5)), q)            # dense superposition on every qubit
    target.data.extend(layers.copy().data)
    ansatz = target.inverse()            # exact solution: after small edits the costs are of order 0.1
    backend = B200SVBackend()
    comp = AdaptCompiler(target, backend=backend)
    comp.full_circuit.data.extend(ansatz.copy().data)
    rng = np.random.default_rng(5)
    lo, hi = comp.variational_circuit_range()
    rot = [i for i in range(lo, hi) if comp.full_circuit.data[i].operation.name in ("rx", "ry", "rz")]
    from harness.minimiser import replace_1q_gate
    visits = [rot[0], rot[1], rot[-1], rot[-2], rot[len(rot) // 2], rot[5], rot[-9], rot[2], rot[-1]]
    for idx in visits:
        for rep in range(2):
            replace_1q_gate(comp.full_circuit, idx, ["rx", "ry", "rz"][int(rng.integers(3))], float(rng.uniform(-0.6, 0.6)))
            got = comp.evaluate_cost()
        sv = backend.simulator.run(comp.full_circuit).result().get_statevector()      # full re-simulation
        ref = 1 - abs(sv[0]) ** 2
        assert abs(got - ref) < 1e-10, (idx, got, ref)
        assert 1e-3 < ref < 1 - 1e-6, ref          # the costs stay away from both trivial values
    st = backend._evaluator.stats
    assert st["projected_evals"] > 0 and st["t_passes"] + st["t_gathers"] > 0
    backend._engine.close()


@pytest.mark.parametrize("n,K", [(14, 12), (18, 13)])
def test_gather_and_scatter_match_numpy(n, K):
    """b200_sv_gather / b200_sv_scatter: projection onto |0> of the qubits outside qmap and its adjoint."""
    rng = np.random.default_rng(60 + n)
    big, small = SVEngine(n, n_slots=2), SVEngine(K, n_slots=1)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    big.upload(0, psi)
    qmap = sorted(int(q) for q in rng.choice(n, K, replace=False))
    qmap = qmap[3:] + qmap[:3]                      # not sorted: compact bit b <-> register qubit qmap[b]
    c = np.arange(1 << K)
    x = np.zeros_like(c)
    for b, q in enumerate(qmap):
        x |= ((c >> b) & 1) << q
    big.gather(0, qmap, small, 0)
    np.testing.assert_array_equal(small.download(0), psi[x])
    big.scatter(1, qmap, small, 0)
    ref = np.zeros(1 << n, dtype=np.complex128)
    ref[x] = psi[x]
    np.testing.assert_array_equal(big.download(1), ref)
    big.close(); small.close()


def test_pair_rdm_parts_sum_to_the_undivided_call_bit_for_bit():
    """b200_sv_pair_rdm_part: the shares of 2 / 3 / 8 ranks sum, bit for bit, to b200_sv_pair_rdm, every
    pair is produced by exactly one share, and each share launches its fraction of the read passes."""
    n = 14
    rng = np.random.default_rng(9)
    eng = SVEngine(n, n_slots=1)
    eng.run(0, -1, GateStream.from_gates(random_gates(n, 150, rng)))
    pairs = [(a, b) for a in range(n) for b in range(a + 1, n)]                # all-to-all: 91 pairs
    l0 = eng.counters()["launches"]
    whole = eng.pair_rdm(0, pairs)
    whole_launches = eng.counters()["launches"] - l0
    for n_parts in (2, 3, 8):
        acc = np.zeros_like(whole)
        owners = np.zeros(len(pairs), dtype=int)
        launches = []
        for part in range(n_parts):
            l0 = eng.counters()["launches"]
            share = eng.pair_rdm(0, pairs, part=part, n_parts=n_parts)
            launches.append(eng.counters()["launches"] - l0)
            owners += np.any(share.reshape(len(pairs), -1) != 0, axis=1)
            acc += share
        assert np.array_equal(acc, whole)
        assert np.all(owners == 1)
        assert sum(launches) == whole_launches and max(launches) <= whole_launches / n_parts + 2
    ref = orc.evaluate_circuit(n, random_gates(n, 150, np.random.default_rng(9)))
    for r, (a, b) in zip(whole, pairs):
        np.testing.assert_allclose(r, orc.partial_trace(ref, a, b), atol=AMP_TOL)
    eng.close()


@pytest.mark.parametrize("world,positions", [(2, [9]), (4, [11, 6]), (8, [7, 12, 10]), (4, None)])
def test_peer_swap_kernels_on_one_gpu(world, positions):
    """b200_sv_peer_swap / b200_sv_peer_swap_strided: the `world` slots of one engine stand in for the ranks' slices
    (every amplitude pair is owned by exactly one rank's kernel, so running the ranks' kernels one after the other on one
    GPU gives what the concurrent launches give over NVLink).  positions = None: contiguous chunks (top local bits)."""
    nl = 13
    g = int(np.log2(world))
    rng = np.random.default_rng(world)
    eng = SVEngine(nl, n_slots=world)
    before = [rng.normal(size=1 << nl) + 1j * rng.normal(size=1 << nl) for _ in range(world)]
    for r in range(world):
        eng.upload(r, before[r])
    ptrs = [eng.device_ptr(r) for r in range(world)]
    for r in range(world):
        eng.peer_swap(r, [None if p == r else ptrs[p] for p in range(world)], r, positions)
    eng.sync()
    pos = positions if positions is not None else [nl - g + j for j in range(g)]
    idx = np.arange(1 << nl)
    chunk_of = np.zeros_like(idx)
    for j, p in enumerate(pos):
        chunk_of |= ((idx >> p) & 1) << j
    for r in range(world):
        got = eng.download(r)
        want = np.empty_like(got)
        for p in range(world):
            # the amplitudes of rank r whose victim bits spell p come from rank p's set that spells r (same other bits)
            sel = chunk_of == p
            src_idx = idx[sel]
            for j, b in enumerate(pos):
                src_idx = (src_idx & ~(1 << b)) | (((r >> j) & 1) << b)
            want[sel] = before[p][src_idx]
        np.testing.assert_array_equal(got, want)
    eng.close()
