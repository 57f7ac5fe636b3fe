"""Parity at the BENCH sizes, against the oracle (not against the device itself).

Round-1 review: at full size the statevector path was only compared with a from-scratch simulation on the same device,
and the MPS path had no check beyond chi = 128.  Here the exact circuits bench.py times (harness.workloads, same seeds)
are evaluated once by the CPU oracle and compared with what the backends return through the reference interface:

  C3  28-qubit brickwork(depth 8, seed 1234) + 16 thin layers: the whole state (up to a global phase), amplitude 0 /
      global cost, all <Z_q>, three pair RDMs; then the incremental evaluator after a full optimiser step
      (Rotoselect + Rotosolve cycle = 220 evaluations on the device) against a second oracle re-simulation.
      Tolerance 1e-10 (BASELINE north_star, complex128 statevector).
  C4  50-qubit random Vidal MPS at chi = 256 + 2 un-absorbed layers: costs of 3 edited circuits in the capped
      (max bond 256: reference contraction order, 512x512 SVD + truncation per CNOT) and default (threshold 1e-16, block
      transfer matrices) modes, bond dimensions equal, costs within 1e-8 (north_star, MPS at the same truncation).

The oracle needs ~25 s per 28-qubit evaluation on the box's host cores and 4 GiB for the state."""
import os

import numpy as np
import pytest

from adapt_aqc_b200.backends import B200SVBackend
from adapt_aqc_b200.mps_backend import B200MPSBackend, B200MPSSimulator
from harness.compiler import AdaptCompiler
from harness.minimiser import B200CostMinimiser, replace_1q_gate
from harness.workloads import build_mps_workload, build_workload
from oracle import mps_oracle as mo
from oracle import sv_oracle as orc
from oracle.oracle_backends import OracleMPSBackend, circuit_to_gates

pytestmark = pytest.mark.gpu

SV_TOL = 1e-10
MPS_TOL = 1e-8


def _oracle_state(n, target, window_circuit):
    orc.lib().orc_set_num_threads(os.cpu_count() or 1)
    return orc.evaluate_circuit(n, circuit_to_gates(target) + circuit_to_gates(window_circuit))


def test_c3_bench_circuit_at_28_qubits_against_the_oracle():
    n = 28
    target, ansatz = build_workload(n, 8, 16)                       # bench.py's default workload
    backend = B200SVBackend()
    comp = AdaptCompiler(target, backend=backend, minimiser_cls=B200CostMinimiser)
    comp.full_circuit.data.extend(ansatz.copy().data)
    lo, hi = comp.variational_circuit_range()

    def window():
        w = comp.full_circuit.copy()
        del w.data[:lo]
        return w

    # ---- the circuit as benchmarked ----
    psi = _oracle_state(n, target, window())
    cost = comp.evaluate_cost()                                     # incremental evaluator (projected tail)
    assert abs(cost - (1 - abs(psi[0]) ** 2)) < SV_TOL
    sv = backend.evaluate_circuit(comp)
    assert abs(abs(sv[0]) - abs(psi[0])) < SV_TOL
    dev = sv.data                                                   # 4 GiB download: the WHOLE state
    k = int(np.argmax(np.abs(psi[:1 << 20])))
    phase = (dev[k] / psi[k]) / abs(dev[k] / psi[k])                # global phase is not observable
    err = 0.0
    for s in range(0, 1 << n, 1 << 24):
        err = max(err, float(np.max(np.abs(dev[s:s + (1 << 24)] - phase * psi[s:s + (1 << 24)]))))
    assert err < 1e-12
    del dev
    np.testing.assert_allclose(backend.measure_qubit_expectation_values(comp), orc.measure_qubit_expectation_values(psi),
                               rtol=0, atol=SV_TOL)
    dsv = backend.simulator.run(comp.full_circuit).result().get_statevector()       # the facade the reference calls per pair
    for a, b in [(0, 1), (13, 14), (5, 27)]:
        np.testing.assert_allclose(dsv.partial_trace(a, b), orc.partial_trace(psi, a, b), rtol=0, atol=SV_TOL)
    del psi

    # ---- after one optimiser step on the device (the bench step: 220 evaluations) ----
    e0 = comp.cost_evaluation_counter
    comp.minimizer._reduce_cost(True, (hi - 5, hi))
    predicted = comp.minimizer._reduce_cost(False, (lo, hi))
    assert comp.cost_evaluation_counter - e0 == 220
    psi2 = _oracle_state(n, target, window())
    ref = 1 - abs(psi2[0]) ** 2
    assert abs(comp.evaluate_cost() - ref) < SV_TOL                 # evaluator caches after 220 edits
    assert abs(predicted - ref) < 1e-9                              # the closed-form minimum the optimiser returned
    st = backend._evaluator.stats
    assert st["projected_evals"] > 0 and st["host_evals"] + st["projected_evals"] >= 220
    for e in backend.engines():
        e.close()


@pytest.mark.parametrize("mode", ["capped", "default"])
def test_c4_chi_256_costs_and_bonds_against_the_oracle(mode):
    n, chi, layers = 50, 256, 2
    cap = chi if mode == "capped" else None
    target, ansatz = build_mps_workload(n, chi, layers)             # bench.py's C4 workload
    comp = AdaptCompiler(target, backend=B200MPSBackend(B200MPSSimulator(1e-16, max_chi=cap)))
    ocomp = AdaptCompiler(target, backend=OracleMPSBackend(mo.OracleMPSSimulator(1e-16, cap)))
    for c in (comp, ocomp):
        c.full_circuit.data.extend(ansatz.copy().data)
    rng = np.random.default_rng(5)
    rot = [i for i in range(*comp.variational_circuit_range()) if comp.full_circuit.data[i].operation.name == "rz"]
    for trial in range(4):
        if trial:                                                   # trial 0: the circuit as benchmarked
            for idx in rng.choice(rot, size=2, replace=False):
                name, theta = ["rx", "ry", "rz"][int(rng.integers(3))], float(rng.uniform(-np.pi, np.pi))
                for c in (comp, ocomp):
                    replace_1q_gate(c.full_circuit, int(idx), name, theta)
        got, ref = comp.evaluate_cost(), ocomp.evaluate_cost()
        assert abs(got - ref) < MPS_TOL, (mode, trial, got, ref)
        assert 0 <= ref <= 1
        # a random 50-qubit state has |<0|psi>|^2 ~ 2^-50, so the cost alone says little: compare the overlap itself,
        # relatively, and O(1) observables of the same state below
        view = comp.backend.evaluate_circuit(comp)
        ref_mps = ocomp.backend.evaluate_circuit(ocomp)
        a_dev = abs(view.handle.amps([0])[0]) ** 2
        a_ref = abs(mo.mps_dot(ref_mps, ocomp.zero_mps, already_preprocessed=True)) ** 2
        assert abs(a_dev - a_ref) <= 1e-6 * a_ref, (mode, trial, a_dev, a_ref)
        dims = view.handle.bond_dims()
        assert dims == [g.shape[2] for g in ref_mps[:-1]]
        assert max(dims) == (chi if mode == "capped" else 2 * chi)
    z, norm = view.handle.expz()
    assert abs(norm - 1) < MPS_TOL
    for q in (0, 23, 24, 25, 26, 49):
        assert abs(z[q] - mo.mps_expectation(ref_mps, "Z", q, already_preprocessed=True)) < MPS_TOL
    rho = comp.backend.mps_ops.partial_trace(view, [24, 25], already_preprocessed=True)
    np.testing.assert_allclose(rho, mo.partial_trace(ref_mps, [24, 25], already_preprocessed=True), rtol=0, atol=MPS_TOL)
