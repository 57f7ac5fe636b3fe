"""bench.py's CPU reference arm (no GPU): the 2-qubit block fusion that stands in for Aer's fusion pass leaves the state
unchanged, and the timed chunks of `--impl reference` really are consecutive pieces of complete evaluations."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from harness.workloads import build_workload  # noqa: E402
from oracle import sv_oracle as orc  # noqa: E402
from oracle.oracle_backends import circuit_to_gates  # noqa: E402

from helpers import random_gates  # noqa: E402


def test_two_qubit_block_fusion_preserves_the_state():
    n = 10
    target, ansatz = build_workload(n, 8, 16)
    gates = circuit_to_gates(target) + circuit_to_gates(ansatz)
    fused = bench.fuse_two_qubit_blocks(n, gates)
    assert len(fused) < len(gates) / 2.5 and all(g[0] in ("mat1", "mat2") for g in fused)
    np.testing.assert_allclose(orc.evaluate_circuit(n, fused), orc.evaluate_circuit(n, gates), atol=1e-12)
    rng = np.random.default_rng(1)
    for trial in range(20):
        n = int(rng.integers(2, 8))
        gates = random_gates(n, int(rng.integers(1, 80)), rng)
        np.testing.assert_allclose(orc.evaluate_circuit(n, bench.fuse_two_qubit_blocks(n, gates)),
                                   orc.evaluate_circuit(n, gates), atol=1e-12)


def test_c3_fusion_count():
    target, ansatz = build_workload(28, 8, 16)
    gates = circuit_to_gates(target) + circuit_to_gates(ansatz)
    assert len(gates) == 404 and len(bench.fuse_two_qubit_blocks(28, gates)) == 124


def test_reference_arm_times_complete_evaluations():
    target, ansatz = build_workload(12, 8, 16)
    psi = orc.evaluate_circuit(12, circuit_to_gates(target) + circuit_to_gates(ansatz))
    for fused in (True, False):
        ref = bench.CpuReference(12, target, ansatz, fused=fused)
        per = -(-ref.G // 7)
        for _ in range(7):
            assert ref.chunk(per) > 0
        assert len(ref.evaluations) == 1 and abs(ref.evaluations[0] - (1 - abs(psi[0]) ** 2)) < 1e-12
    env = dict(os.environ, OMP_NUM_THREADS="1")          # what torchrun exports to its workers
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--qubits", "14", "--steps", "6",
                          "--warmup", "2"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["extrapolated"] is False
    assert line["cpu_baseline"]["cores"] == bench.host_threads()          # set explicitly, not inherited
    assert abs(line["steps"] * line["ms_per_step"] * 1e-3 - 1.0 / line["value"] * line["cpu_baseline"]["evaluations_timed"]) < 1e-6
    assert line["wall_s"] >= line["steps"] * line["ms_per_step"] * 1e-3
