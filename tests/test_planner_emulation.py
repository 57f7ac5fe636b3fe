"""Planner + kernel thread bodies, run on the CPU (tests/emu), against the oracle.

This is the `not gpu` safety net for the sm_100a sweep kernel: the *same* __host__ __device__
functions the CUDA kernel calls are executed thread by thread on the host, driven by the same
planner, and compared with the oracle on seeded random circuits."""
import numpy as np
import pytest

from adapt_aqc_b200.gates import GateStream
from adapt_aqc_b200.sv_engine import plan_stats
from oracle import sv_oracle as orc
from oracle.oracle_backends import circuit_to_gates

from helpers import brickwork, emu_run, random_gates, thin_ansatz

TOL = 1e-12


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 10, 11, 12, 13, 15])
def test_random_circuits_all_opcodes(emu, n):
    rng = np.random.default_rng(1000 + n)
    for trial in range(3):
        gates = random_gates(n, int(rng.integers(1, 150)), rng)
        ref = orc.evaluate_circuit(n, gates)
        got, stats = emu_run(emu, n, gates)
        np.testing.assert_allclose(got, ref, atol=TOL)
        assert stats[3] == (1 if n <= 11 else 0)


@pytest.mark.parametrize("n", [3, 11, 12, 14])
def test_inverse_round_trip(emu, n):
    rng = np.random.default_rng(2000 + n)
    gates = random_gates(n, 80, rng)
    psi, _ = emu_run(emu, n, gates)
    back, _ = emu_run(emu, n, gates, psi0=psi, inverse=True)
    e0 = np.zeros(1 << n, dtype=np.complex128); e0[0] = 1
    np.testing.assert_allclose(back, e0, atol=TOL)


def test_empty_and_identity_streams(emu):
    for n in (2, 12):
        got, _ = emu_run(emu, n, [])
        assert got[0] == 1 and np.count_nonzero(got) == 1
        got, _ = emu_run(emu, n, [("id", [0], []), ("id", [n - 1], [])])
        assert got[0] == 1 and np.count_nonzero(got) == 1


def test_brickwork_plus_ansatz_matches_oracle(emu):
    n = 16
    target, rng = brickwork(n, 4, seed=1234)
    ansatz = thin_ansatz(n, 8, rng)
    gates = circuit_to_gates(target) + circuit_to_gates(ansatz)
    ref = orc.evaluate_circuit(n, gates)
    got, stats = emu_run(emu, n, gates)
    np.testing.assert_allclose(got, ref, atol=TOL)
    # fused sweeps: far fewer passes over the state than gates
    assert stats[0] <= len(gates) // 6


def test_lane_qubit_mixing_is_served_by_shuffles_or_shared_memory_rounds(emu):
    """Rounds that touch HBM keep qubits 0..2 on the warp lanes.  Up to two mixing gates there go
    through shuffles in the same round; more of them (or a dense 2-qubit gate) get shared-memory
    rounds between a load round and a store round.  Qubits 3 and 4 can be register qubits."""
    n = 13
    gates = [("h", [0], []), ("cx", [0, 1], []), ("ry", [3], [0.4])]
    ref = orc.evaluate_circuit(n, gates)
    got, stats = emu_run(emu, n, gates)
    np.testing.assert_allclose(got, ref, atol=TOL)
    assert stats[0] == 1 and stats[1] == 1
    gates = [("h", [0], []), ("h", [1], []), ("h", [2], []), ("cx", [2, 0], []), ("ry", [1], [0.4]), ("cx", [9, 1], [])]
    ref = orc.evaluate_circuit(n, gates)
    got, stats = emu_run(emu, n, gates)
    np.testing.assert_allclose(got, ref, atol=TOL)
    assert stats[0] == 1 and stats[1] >= 2


@pytest.mark.parametrize("n", [12, 14])
def test_lane_ops_every_control_kind(emu, n):
    """X / dense 1q on each lane qubit with no control, a lane control, a register control and a
    thread-level control; diagonals with register / thread-level / mixed qubits."""
    rng = np.random.default_rng(77 + n)
    for t in range(3):
        for c in (None, (t + 1) % 3, 3, 4, 7, n - 1):
            gates = [("h", [q], []) for q in range(n)] + [("rz", [q], [0.1 + 0.3 * q]) for q in range(n)]
            gates += [("cx", [c, t], [])] if c is not None else [("x", [t], [])]
            gates += [("u3", [t], list(rng.uniform(-3, 3, 3))), ("cz", [t, 5], []), ("cz", [3, 8], []), ("cz", [0, 1], [])]
            gates += [("cx", [t, 6], []), ("rz", [6], [0.7]), ("cx", [n - 1, 4], [])]
            ref = orc.evaluate_circuit(n, gates)
            got, _ = emu_run(emu, n, gates)
            np.testing.assert_allclose(got, ref, atol=TOL)


def test_thin_layer_is_one_cx_and_one_diagonal():
    """Diagonal fusion: rz rz cx rz rz -> cx + one 2-qubit phase, 1 round; the cx leads its round, so it is folded into
    the round's load addressing and ONE op (the phase) is executed."""
    n = 16
    gates = [("rz", [9], [0.3]), ("rz", [10], [-0.2]), ("cx", [9, 10], []), ("rz", [9], [1.1]), ("rz", [10], [0.5])]
    st = plan_stats(n, GateStream.from_gates(gates))
    assert tuple(st[:3]) == (1, 1, 1), st


def test_diagonal_and_control_qubits_do_not_need_tile_slots(emu):
    """rz / cz / cx-controls on 12 different high qubits still fit one sweep: only mixing targets
    occupy tile slots."""
    n = 20
    gates = [("h", [19], [])]
    for q in range(5, 19):
        gates += [("rz", [q], [0.1 * q]), ("cz", [q, 19], [])]
    gates += [("cx", [7, 19], []), ("cx", [12, 19], [])]
    gs = GateStream.from_gates(gates)
    assert plan_stats(n, gs)[0] == 1
    ref = orc.evaluate_circuit(n, gates)
    got, _ = emu_run(emu, n, gates)
    np.testing.assert_allclose(got, ref, atol=TOL)


def test_plan_for_c3_workload_is_a_few_sweeps():
    """BASELINE config C3: 28-qubit depth-8 brickwork target, 16 thinly dressed layers."""
    n = 28
    target, rng = brickwork(n, 8, seed=1234)
    ansatz = thin_ansatz(n, 16, rng)
    t_stats = plan_stats(n, GateStream.from_circuit(target))
    a_stats = plan_stats(n, GateStream.from_circuit(ansatz))
    assert t_stats[0] <= 24, t_stats
    assert a_stats[0] <= 3, a_stats


@pytest.mark.parametrize("n", [3, 12])
def test_invert_window_is_the_inverse_circuit(emu, n):
    """gates.invert_window (used by the sharded engine, which has no inverse entry point of its own)
    against the library's inverse canonicalisation, on every opcode."""
    from adapt_aqc_b200.gates import canonical_window, invert_window
    from helpers import circuit_from_gates
    rng = np.random.default_rng(300 + n)
    gates = random_gates(n, 70, rng)
    window = canonical_window(circuit_from_gates(n, gates))
    psi, _ = emu_run(emu, n, GateStream.from_window(window))
    back, _ = emu_run(emu, n, GateStream.from_window(invert_window(window)), psi0=psi)
    via_flag, _ = emu_run(emu, n, GateStream.from_window(window), psi0=psi, inverse=True)
    e0 = np.zeros(1 << n, dtype=np.complex128); e0[0] = 1
    np.testing.assert_allclose(back, e0, atol=TOL)
    np.testing.assert_allclose(via_flag, e0, atol=TOL)


@pytest.mark.parametrize("n", [12, 13, 16])
def test_folded_permutations_match_oracle(emu, n):
    """X / CX that commute to the front or back of their round are folded into the round's load / store addressing
    (sv_plan.h PFold): register-controlled CX chains (GF(2)-linear relabelling resolved on the host), thread-level
    controls and unconditional X (per-thread XOR masks), from |0..0> (implicit source) and from a given state,
    forwards and inverted.  The direct (folded) and pipelined (executed) bodies must agree bit for bit (emu_run)."""
    rng = np.random.default_rng(900 + n)
    for trial in range(12):
        gates = []
        if trial % 2:
            gates += [("h", [q], []) for q in range(n)]
        for _ in range(int(rng.integers(5, 60))):
            k = int(rng.integers(6))
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            gates.append([("x", [a], []), ("cx", [a, b], []), ("cx", [a, b], []), ("cz", [a, b], []),
                          ("rz", [a], [float(rng.uniform(-3, 3))]), ("u3", [a], list(rng.uniform(-3, 3, 3)))][k if trial % 3 else min(k, 4)])
        ref = orc.evaluate_circuit(n, gates)
        got, _ = emu_run(emu, n, gates)
        np.testing.assert_allclose(got, ref, atol=TOL)
        psi0 = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        psi0 /= np.linalg.norm(psi0)
        got2, _ = emu_run(emu, n, gates, psi0=psi0)
        np.testing.assert_allclose(got2, orc.apply_gates(psi0, gates), atol=TOL)
        back, _ = emu_run(emu, n, gates, psi0=got2, inverse=True)
        np.testing.assert_allclose(back, psi0, atol=TOL)


def test_ansatz_sweeps_execute_only_their_diagonals():
    """C3 ansatz (16 thin layers, brickwall order): after diagonal fusion every layer is cx + one 2-qubit phase, and every
    cx whose target is a register qubit is folded into the addressing of its round -- the sweeps execute the 16 phases + the
    cx whose targets are warp-lane bits of an HBM round (folded into the global address when they lead / trail it, else served
    by shuffles), instead of 32 ops."""
    n = 28
    _, rng = brickwork(n, 8, seed=1234)
    ansatz = thin_ansatz(n, 16, rng)
    st = plan_stats(n, GateStream.from_circuit(ansatz))
    assert st[0] <= 3 and st[2] <= 17, st          # 16 phases + at most one cx on a lane qubit


@pytest.mark.parametrize("n", [12, 14, 17])
def test_phase_heavy_rounds_match_oracle(emu, n):
    """Rounds dominated by phase ops: 1- and 2-qubit diagonals on register, thread-level and mixed qubits, interleaved
    with cx (folded or executed) and with NON-unitary diagonals (|ratio| != 1, zero entries: the generic slow path)."""
    rng = np.random.default_rng(1200 + n)
    for trial in range(10):
        gates = [("h", [q], []) for q in range(n)] if trial % 2 == 0 else [("ry", [q], [0.3 + 0.1 * q]) for q in range(n)]
        for _ in range(int(rng.integers(10, 70))):
            k = int(rng.integers(8))
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            th = float(rng.uniform(-3, 3))
            if k == 0: gates.append(("rz", [a], [th]))
            elif k == 1: gates.append(("cz", [a, b], []))
            elif k == 2: gates.append(("u1", [a], [th]))
            elif k == 3: gates.append(("mat2", [a, b], np.diag(np.exp(1j * rng.uniform(-3, 3, 4)))))      # generic 2-qubit phase
            elif k == 4: gates.append(("t", [a], []))
            elif k == 5: gates.append(("cx", [a, b], []))
            elif k == 6 and trial % 3 == 0: gates.append(("mat1", [a], np.diag([1.0, 0.5 * np.exp(1j * th)])))   # not unitary
            elif k == 6 and trial % 3 == 1: gates.append(("mat1", [a], np.diag([1.0, 0.0])))                    # projector
            else: gates.append(("ry", [a], [th]))
        ref = orc.evaluate_circuit(n, gates)
        got, _ = emu_run(emu, n, gates)
        np.testing.assert_allclose(got, ref, atol=TOL)


def _transfer(L, R, n, qa, qb):
    """T[i][j] = sum_rest conj(L[i,rest]) R[j,rest], index = bit(qa) + 2 bit(qb) (numpy restatement of b200_sv_inner2)."""
    Lt = np.moveaxis(L.reshape([2] * n), [n - 1 - qb, n - 1 - qa], [0, 1]).reshape(4, -1)
    Rt = np.moveaxis(R.reshape([2] * n), [n - 1 - qb, n - 1 - qa], [0, 1]).reshape(4, -1)
    return Lt.conj() @ Rt.T


@pytest.mark.parametrize("n", [12, 13, 15])
def test_fused_sweep_and_transfer_pass_matches_the_two_separate_steps(emu, n):
    """sv_sweep_inner2_kernel's bodies on the CPU: the swept state equals the plain sweep's, T equals the transfer matrix of
    (swept state, other) -- for pairs on lane qubits, register qubits, tile padding and qubits the gates never touch, for
    empty programs, multi-sweep programs and X-heavy rounds whose trailing folds must not leak into the stored tile."""
    from helpers import emu_run_inner2
    rng = np.random.default_rng(4200 + n)
    dim = 1 << n
    for trial in range(8):
        psi0 = rng.normal(size=dim) + 1j * rng.normal(size=dim); psi0 /= np.linalg.norm(psi0)
        other = rng.normal(size=dim) + 1j * rng.normal(size=dim); other /= np.linalg.norm(other)
        if trial == 0:
            gates = []
        elif trial == 1:      # thin layers: cx + rz (folded permutations and lane-qubit X)
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
        ref, _ = emu_run(emu, n, gates, psi0=psi0, inverse=inverse)
        got, T, stats = emu_run_inner2(emu, n, gates, psi0, other, qa, qb, inverse=inverse)
        np.testing.assert_allclose(got, ref, atol=1e-13)
        np.testing.assert_allclose(T, _transfer(ref, other, n, qa, qb), atol=1e-12)
        # without the write-back the state stays as it was when the program is a single sweep
        if stats[0] == 1:
            keep, T2, _ = emu_run_inner2(emu, n, gates, psi0, other, qa, qb, inverse=inverse, write_back=False)
            np.testing.assert_array_equal(keep, psi0)
            np.testing.assert_array_equal(T2, T)


@pytest.mark.parametrize("n", [12, 14])
def test_embedded_source_sweep_matches_scatter_then_run(emu, n):
    """The first sweep reads its tiles from a compact 2^K-amplitude array (b200_sv_run_embedded): same state as embedding
    first and sweeping afterwards, same T from the fused variant -- for sorted and permuted qmaps, K from 2 to n."""
    from helpers import emu_run_embedded
    rng = np.random.default_rng(4300 + n)
    dim = 1 << n
    for trial in range(6):
        K = [2, 5, n - 1, n, 7, 9][trial]
        qmap = [int(q) for q in (rng.permutation(n)[:K] if trial % 2 else np.sort(rng.permutation(n)[:K]))]
        phi = rng.normal(size=1 << K) + 1j * rng.normal(size=1 << K); phi /= np.linalg.norm(phi)
        other = rng.normal(size=dim) + 1j * rng.normal(size=dim); other /= np.linalg.norm(other)
        gates = random_gates(n, int(rng.integers(1, 100)), rng)
        c = np.arange(1 << K)
        x = np.zeros_like(c)
        for b, q in enumerate(qmap):
            x |= ((c >> b) & 1) << q
        psi = np.zeros(dim, dtype=np.complex128); psi[x] = phi
        inverse = bool(trial % 2)
        ref, _ = emu_run(emu, n, gates, psi0=psi, inverse=inverse)
        got, none = emu_run_embedded(emu, n, gates, phi, qmap, inverse=inverse)
        assert none is None
        np.testing.assert_allclose(got, ref, atol=1e-13)
        qa, qb = [int(q) for q in rng.choice(n, size=2, replace=False)]
        got2, T = emu_run_embedded(emu, n, gates if trial else [], phi, qmap, inverse=inverse, fuse=(qa, qb), other=other)
        ref2 = ref if trial else psi
        np.testing.assert_allclose(got2, ref2, atol=1e-13)
        np.testing.assert_allclose(T, _transfer(ref2, other, n, qa, qb), atol=1e-12)


@pytest.mark.parametrize("n", [12, 14])
def test_projected_store_sweep_matches_run_then_gather(emu, n):
    """The last round keeps only the amplitudes with |0> on every qubit outside qmap (b200_sv_run_project): same 2^K numbers
    as sweeping the whole state and gathering them afterwards, every entry of phi written exactly once."""
    from helpers import emu_run_project
    rng = np.random.default_rng(4400 + n)
    dim = 1 << n
    for trial in range(7):
        K = [2, 5, n - 1, n, 7, 9, 8][trial]
        qmap = [int(q) for q in (rng.permutation(n)[:K] if trial % 2 else np.sort(rng.permutation(n)[:K]))]
        psi0 = rng.normal(size=dim) + 1j * rng.normal(size=dim); psi0 /= np.linalg.norm(psi0)
        gates = [] if trial == 0 else random_gates(n, int(rng.integers(1, 100 if trial < 6 else 400)), rng)
        inverse = bool(trial % 2)
        ref, _ = emu_run(emu, n, gates, psi0=psi0, inverse=inverse)
        c = np.arange(1 << K)
        x = np.zeros_like(c)
        for b, q in enumerate(qmap):
            x |= ((c >> b) & 1) << q
        phi, sweeps = emu_run_project(emu, n, gates, psi0, qmap, inverse=inverse)
        np.testing.assert_allclose(phi, ref[x], atol=1e-13)
