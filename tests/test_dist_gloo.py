"""Multi-process (gloo, CPU) tests of the sharded statevector's host logic: qubit layout
bookkeeping, in-place chunked all-to-all exchange, resolution of global control / diagonal qubits,
cross-rank reductions of the read-outs.  The GPU path swaps gloo for NCCL and the emulated local
engine for SVEngine; everything else is the same code."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,n,port", [(2, 7, 29611), (4, 8, 29612)])
def test_sharded_statevector_matches_oracle(emu, world, n, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), str(n)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "dist ok" in res.stdout


@pytest.mark.parametrize("world,n,port", [(2, 8, 29613), (4, 9, 29614)])
def test_compile_on_a_sharded_register_matches_oracle(emu, world, n, port):
    """B200ShardedSVBackend: the ADAPT loop over a register sharded by global qubits -- projected-tail
    evaluations (ranked gather + all-reduce), re-simulation fallback, sharded <Z> / pair-RDM read-outs --
    makes the decisions of the oracle backend."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"),
           str(n), "cpu", "compile"]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "dist compile ok" in res.stdout


@pytest.mark.parametrize("world,n,port", [(2, 6, 29615), (4, 7, 29616)])
def test_pair_rdm_passes_divided_over_replicas_make_the_same_decisions(emu, world, n, port):
    """ranks holding replicas of the state divide the ISL pair-RDM passes and all-reduce the P x 16
    complex results (B200SVBackend(pair_comm=...)); pair history, EM values and costs equal the undivided compile."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"),
           str(n), "cpu", "pairsplit"]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "dist pairsplit ok" in res.stdout
