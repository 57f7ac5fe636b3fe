"""NCCL version of tests/test_dist_gloo.py: the sharded statevector on >= 2 real GPUs (skipped on a
one-GPU box).  Same worker, same oracle checks; the local engine is SVEngine on torch-owned
device buffers and the exchange goes over NVLink."""
import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    from adapt_aqc_b200.lib import load
    n = ctypes.c_int(0)
    return n.value if load().b200_device_count(ctypes.byref(n)) == 0 else 0


@pytest.mark.parametrize("world,n,port,exchange", [(2, 15, 29621, "peer"), (2, 15, 29624, "nccl"), (4, 16, 29622, "peer"),
                                                   (8, 17, 29623, "peer"), (8, 17, 29625, "nccl")])
def test_sharded_statevector_on_nccl(world, n, port, exchange):
    """exchange = "peer": global qubits are swapped by b200_sv_peer_swap over NVLink peer memory (CUDA IPC);
    "nccl": send/recv through a staging buffer."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, B200AQC_EXCHANGE=exchange)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"),
           str(n), "gpu"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "dist ok" in res.stdout
    assert f"exchange={exchange}" in res.stdout, res.stdout[-500:]


@pytest.mark.parametrize("world,n,port", [(2, 20, 29631), (8, 22, 29632)])
def test_compile_on_a_sharded_register(world, n, port):
    """B200ShardedSVBackend on real GPUs: same pairs, costs and evaluation count as the single-GPU backend."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"),
           str(n), "gpu", "compile"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "dist compile ok" in res.stdout


@pytest.mark.parametrize("world,n,port", [(2, 18, 29641), (8, 20, 29642)])
def test_pair_rdm_passes_divided_over_replicas(world, n, port):
    """B200SVBackend(pair_comm=...) on real GPUs: b200_sv_pair_rdm_part shares + NCCL all-reduce give, bit for bit, the
    entanglement measures and therefore the pair history of the single-GPU compile."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"),
           str(n), "gpu", "pairsplit"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "dist pairsplit ok" in res.stdout
