"""Worker of tests/test_dist_gloo.py: one rank of a gloo process group checking the sharded
statevector (adapt_aqc_b200.dist_sv) against the oracle.  CPU only: the local engine is the
emulated kernel path (tests/helpers.FakeEngine)."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import adapt_aqc_b200  # noqa: E402,F401
from adapt_aqc_b200.dist_sv import ShardedStatevector, TorchComm  # noqa: E402
from adapt_aqc_b200.gates import GateStream  # noqa: E402
from helpers import FakeEngine, random_gates  # noqa: E402
from oracle import sv_oracle as orc  # noqa: E402


def load_emu():
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libb200aqc_emu.so"))
    lib.emu_last_error.restype = ctypes.c_char_p
    dp = ctypes.POINTER(ctypes.c_double)
    lib.emu_sv_run.argtypes = [ctypes.c_int, dp, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                               ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]
    lib.emu_sv_run_inner2.argtypes = [ctypes.c_int, dp, dp, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, ctypes.POINTER(ctypes.c_int32)]
    lib.emu_sv_run_project.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]
    lib.emu_sv_run_embedded.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp,
                                        ctypes.POINTER(ctypes.c_int32)]
    return lib


def gather_state(sv, slot):
    """Full logical-order statevector on every rank (test only)."""
    comm = sv.comm
    if hasattr(sv.eng, "download"):
        host = sv.eng.download(slot)
    else:
        host = sv.eng.slots[slot].copy()
    local = torch.from_numpy(np.ascontiguousarray(host).view(np.float64)).to(comm.device)
    parts = [torch.empty_like(local) for _ in range(comm.world)]
    dist.all_gather(parts, local)
    phys = np.concatenate([p.cpu().numpy().view(np.complex128) for p in parts])      # physical order
    n, perm = sv.n, sv.perm[slot]
    idx = np.arange(1 << n)
    pidx = np.zeros_like(idx)
    for l in range(n):
        pidx |= ((idx >> l) & 1) << perm[l]
    return phys[pidx]


def compile_main(n, on_gpu):
    """mode `compile`: the ADAPT loop on a sharded register (B200ShardedSVBackend) makes the same decisions
    as on the oracle backend -- pair RDMs and <Z> are sharded read-outs, cost evaluations run in the
    projected tail (gather + all-reduce) or by re-simulation on the sharded register."""
    from harness.compiler import AdaptCompiler, AdaptConfig
    from adapt_aqc_b200.dist_sv import B200ShardedSVBackend, ShardedEngine
    from adapt_aqc_b200.sv_engine import SVCostEvaluator
    from helpers import brickwork
    from oracle.oracle_backends import OracleSVBackend
    if on_gpu:
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        backend = B200ShardedSVBackend(local_rank)
        proj_k = None
    else:
        dist.init_process_group("gloo")
        emu = load_emu()
        proj_k = 4

        class CpuSharded(B200ShardedSVBackend):
            def _make_engines(self, num_qubits):
                comm = TorchComm()
                nl = num_qubits - int(np.log2(comm.world))
                eng = FakeEngine(emu, nl, n_slots=4)
                tensors = [torch.from_numpy(s.view(np.float64)) for s in eng.slots]
                staging = torch.empty(max(8, (2 << nl) // comm.world // 3), dtype=torch.float64)
                sv = ShardedStatevector(num_qubits, eng, comm, tensors, staging)
                return ShardedEngine(sv), [FakeEngine(emu, 5, n_slots=2)], [FakeEngine(emu, proj_k, n_slots=4)]
        backend = CpuSharded()
    target, _ = brickwork(n, 2, seed=21)
    cfg = dict(max_layers=4 if not on_gpu else 5)
    got = AdaptCompiler(target, backend=backend, adapt_config=AdaptConfig(**cfg)).compile()
    st = dict(backend._evaluator.stats)
    if n <= 16 or on_gpu:
        # reference: the oracle on the CPU; at GPU sizes the single-GPU backend (parity-tested against the oracle)
        from adapt_aqc_b200.backends import B200SVBackend
        ref_backend = B200SVBackend(device=int(os.environ.get("LOCAL_RANK", "0"))) if on_gpu else OracleSVBackend()
        ref = AdaptCompiler(target, backend=ref_backend, adapt_config=AdaptConfig(**cfg)).compile()
        assert got.qubit_pair_history == ref.qubit_pair_history, (got.qubit_pair_history, ref.qubit_pair_history)
        assert np.allclose(got.global_cost_history, ref.global_cost_history, atol=1e-9)
        assert got.cost_evaluations == ref.cost_evaluations
    assert st["projected_evals"] > 0, st
    if on_gpu:
        backend._engine.close()
    if dist.get_rank() == 0:
        print(f"dist compile ok: world={dist.get_world_size()} n={n} layers={len(got.qubit_pair_history)} "
              f"evals={got.cost_evaluations} cost={got.global_cost_history[-1]:.6f} stats={st}")
    dist.destroy_process_group()


def pairsplit_main(n, on_gpu):
    """mode `pairsplit`: every rank runs the SAME compile on a replica of the state; the pair-RDM read
    passes of the ISL heuristic (adapt_compiler.py:955-976) are divided among the ranks and summed with one small
    all-reduce.  Decisions must be those of the undivided run (and of the oracle at CPU sizes)."""
    from adapt_aqc_b200.backends import B200SVBackend
    from adapt_aqc_b200.sv_engine import SVCostEvaluator
    from harness.compiler import AdaptCompiler, AdaptConfig
    from helpers import brickwork
    from oracle.oracle_backends import OracleSVBackend
    if on_gpu:
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = TorchComm(torch.device("cuda", local_rank))
        make = lambda pc: B200SVBackend(device=local_rank, pair_comm=pc)   # noqa: E731
    else:
        dist.init_process_group("gloo")
        comm = TorchComm()
        emu = load_emu()

        class CpuBackend(B200SVBackend):
            def _get_engine(self, num_qubits):
                if self._engine is None or self._engine.num_qubits != num_qubits:
                    self._engine = FakeEngine(emu, num_qubits)
                    self._evaluator = SVCostEvaluator(self._engine, None, None)
                    self._state_version += 1
                    self._last_run_key = None
                return self._engine
        make = lambda pc: CpuBackend(pair_comm=pc)   # noqa: E731
    target, _ = brickwork(n, 2, seed=33)
    cfg = dict(max_layers=4)
    split_backend = make(comm)
    got = AdaptCompiler(target, backend=split_backend, adapt_config=AdaptConfig(**cfg)).compile()      # all-to-all map
    whole = AdaptCompiler(target, backend=make(None), adapt_config=AdaptConfig(**cfg)).compile()
    assert got.qubit_pair_history == whole.qubit_pair_history, (got.qubit_pair_history, whole.qubit_pair_history)
    for a, b in zip(got.entanglement_measures_history, whole.entanglement_measures_history):
        assert a == b, "divided pair passes must reproduce the undivided values bit for bit"
    assert got.global_cost_history == whole.global_cost_history and got.cost_evaluations == whole.cost_evaluations
    if n <= 14:
        ref = AdaptCompiler(target, backend=OracleSVBackend(), adapt_config=AdaptConfig(**cfg)).compile()
        assert got.qubit_pair_history == ref.qubit_pair_history
        assert np.allclose(got.global_cost_history, ref.global_cost_history, atol=1e-9)
    if on_gpu:      # each rank launched ~1/world of the RDM passes
        eng = split_backend._engine
        # (counted through the profile classes would need profiling on; the launch counter is enough here)
        mine = eng.counters()["launches"]
        tot = comm.allreduce_sum(np.array([float(mine)]))[0]
        assert mine < 0.75 * tot or comm.world == 1
        for e in split_backend.engines():
            e.close()
    if comm.rank == 0:
        print(f"dist pairsplit ok: world={comm.world} n={n} pairs={got.qubit_pair_history}")
    dist.destroy_process_group()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    on_gpu = len(sys.argv) > 2 and sys.argv[2] == "gpu"
    if len(sys.argv) > 3 and sys.argv[3] == "compile":
        return compile_main(n, on_gpu)
    if len(sys.argv) > 3 and sys.argv[3] == "pairsplit":
        return pairsplit_main(n, on_gpu)
    if on_gpu:
        from adapt_aqc_b200.dist_sv import make_gpu_sharded
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # a staging buffer of 1/3 chunk forces multi-piece exchanges here too
        sv = make_gpu_sharded(n, n_slots=2, local_rank=local_rank,
                              staging_bytes=8 * max(8, (2 << (n - int(np.log2(dist.get_world_size())))) // dist.get_world_size() // 3))
        comm = sv.comm
    else:
        dist.init_process_group("gloo")
        comm = TorchComm()
        g = int(np.log2(comm.world))
        nl = n - g
        emu = load_emu()
        eng = FakeEngine(emu, nl, n_slots=2)
        tensors = [torch.from_numpy(s.view(np.float64)) for s in eng.slots]
        staging = torch.empty(max(8, (2 << nl) // comm.world // 3), dtype=torch.float64)   # forces multi-piece exchanges
        sv = ShardedStatevector(n, eng, comm, tensors, staging)

    rng = np.random.default_rng(123)              # same stream on every rank
    for trial in range(4):
        gates = random_gates(n, 60, rng)
        window = GateStream  # noqa: F841
        from adapt_aqc_b200.gates import canonical_window
        from helpers import circuit_from_gates
        circ = circuit_from_gates(n, gates)
        win = canonical_window(circ)
        sv.run(0, -1, win)
        ref = orc.evaluate_circuit(n, gates)
        got = gather_state(sv, 0)
        assert np.max(np.abs(got - ref)) < 1e-12, (trial, np.max(np.abs(got - ref)))
        # read-outs
        for index in (0, 1, (1 << n) - 1, 1 << (n - 1), 37 % (1 << n)):
            assert abs(sv.amp(0, index) - ref[index]) < 1e-12
        z, norm = sv.expz(0)
        assert np.allclose(z, orc.measure_qubit_expectation_values(ref), atol=1e-12) and abs(norm - 1) < 1e-12
        pairs = [(a, b) for a in range(n) for b in range(n) if a != b]
        rho = sv.pair_rdm(0, pairs)
        for r, (a, b) in zip(rho, pairs):
            assert np.allclose(r, orc.partial_trace(ref, a, b), atol=1e-12), (a, b)
        # second segment applied to the (possibly permuted) state, out of place
        more = random_gates(n, 30, rng)
        sv.run(1, 0, canonical_window(circuit_from_gates(n, more)))
        ref2 = orc.apply_gates(orc.evaluate_circuit(n, gates), more)
        assert np.max(np.abs(gather_state(sv, 1) - ref2)) < 1e-12
        assert abs(sv.inner(1, 1) - 1) < 1e-12
    assert sv.stats["exchanges"] > 0, "the test circuits must exercise the global-qubit exchange"
    if on_gpu:
        sv.close()
    if comm.rank == 0:
        print(f"dist ok: world={comm.world} n={n} exchanges={sv.stats['exchanges']} bytes_sent={comm.bytes_sent} "
              f"exchange={getattr(sv, 'exchange_mode', 'nccl')}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
