#!/usr/bin/env python
"""bench.py -- cost evaluations / second of the ADAPT-AQC hot path on B200 (BASELINE.json metric).

Workload: 28-qubit brickwork target (depth 8, seed 1234) + 16 thinly
dressed CNOT layers in brickwall order.  One STEP = the evaluation stream the reference optimiser
issues for one layer: Rotoselect over the newest layer (4 rotations x 7 costs,
cost_minimiser.py:318-342) followed by one Rotosolve cycle over all 64 rotations (3 costs each,
cost_minimiser.py:344-368) = 220 cost evaluations.

  value   : evals/s with everything resident in HBM, batched front end (B200CostMinimiser: one
            transfer-matrix launch serves all shift values of a gate).
  e2e     : evals/s through the reference-facing interface -- the unmodified CostMinimiser asking
            backend.evaluate_global_cost(compiler) for ONE scalar per call; the circuit lives on
            the host, every call ships gate records / plans to the device and reads the
            amplitude back (bytes counted by the library).
  roofline: the fused gate-sweep kernel (sv_sweep_kernel), 32 * 2^n algorithmic bytes per
            launch, timed per launch with CUDA events on the library's stream.

`--impl reference` times the CPU restatement of the reference path (oracle/, full re-simulation of
every gate from |0..0> per evaluation, aer_sv_backend.py:37-47) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RESULT_OUT = sys.stdout
METRIC = "cost_evals_per_sec"
UNIT = "evals/s"


from harness.workloads import build_mps_workload, build_workload  # noqa: E402


def make_compiler(target, ansatz, backend, batched):
    from harness.compiler import AdaptCompiler
    from harness.minimiser import B200CostMinimiser
    comp = AdaptCompiler(target, backend=backend, minimiser_cls=B200CostMinimiser if batched else None)
    comp.full_circuit.data.extend(ansatz.copy().data)
    return comp


def one_step(comp, n_layer_gates=5):
    """Rotoselect over the newest layer, then one Rotosolve cycle over the whole ansatz."""
    lo, hi = comp.variational_circuit_range()
    comp.minimizer._reduce_cost(True, (hi - n_layer_gates, hi))
    comp.minimizer._reduce_cost(False, (lo, hi))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback"


def measured_traffic(n, kernel="sv_sweep_kernel"):
    """DRAM bytes per launch of the sweep kernel (or of the fused sweep + transfer pass) from the committed
    `ncu --set full` capture (profiles/sweep_traffic.json, taken at n = 28); None for other sizes / no capture."""
    p = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    try:
        t = json.load(open(p))
        if kernel != "sv_sweep_kernel":
            t = t[kernel]
        per_amp = 48 if kernel.endswith("(store)") else 32
        if int(t["algorithmic_bytes_per_launch"]) == per_amp * (1 << n):
            return float(t["traffic_bytes_per_launch"])
    except Exception:  # noqa: BLE001
        pass
    return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def fuse_two_qubit_blocks(n, gates):
    """Greedy 2-qubit block fusion of an oracle gate list [(name, qubits, params)] -> list of mat2 / mat1 gates with the
    same product: 1-qubit gates are absorbed into the neighbouring 2-qubit gate on their qubit, consecutive 2-qubit gates
    on the same pair are multiplied together.  qiskit-aer fuses gates before applying them (fusion is on by default from
    14 qubits up), so timing one sweep per ORIGINAL gate would overstate the reference's cost; this is the baseline's
    stand-in for that pass (brickwork C3: 404 gates -> 124 dense 4x4 sweeps)."""
    from oracle.mps_oracle import gate_matrix
    pending = [None] * n            # 2x2 waiting for a 2-qubit gate on that qubit
    block_of = [None] * n           # index into `out` of the open 2-qubit block whose LAST op on this qubit it is
    out = []                        # [qubits(a, b), 4x4] with index bit(a) + 2 bit(b)

    def kron_on(m, pos):            # 2x2 on bit `pos` of a 2-qubit index
        return np.kron(np.eye(2), m) if pos == 0 else np.kron(m, np.eye(2))

    for name, qubits, params in gates:
        m = np.asarray(params, dtype=np.complex128).reshape(2 ** len(qubits), -1) if name in ("mat1", "mat2") \
            else gate_matrix(name, params)
        if len(qubits) == 1:
            q = qubits[0]
            b = block_of[q]
            if b is not None:
                pos = out[b][0].index(q)
                out[b][1] = kron_on(m, pos) @ out[b][1]
            else:
                pending[q] = m if pending[q] is None else m @ pending[q]
            continue
        a, c = qubits
        ba, bc = block_of[a], block_of[c]
        if ba is not None and ba == bc:                       # same open block on the same pair: multiply in
            qa, _ = out[ba][0]
            mm = m if qa == a else m[np.ix_([0, 2, 1, 3], [0, 2, 1, 3])]
            out[ba][1] = mm @ out[ba][1]
            continue
        pre = np.eye(4, dtype=np.complex128)
        for q, pos in ((a, 0), (c, 1)):
            if pending[q] is not None:
                pre = kron_on(pending[q], pos) @ pre
                pending[q] = None
        out.append([(a, c), m @ pre])
        block_of[a] = block_of[c] = len(out) - 1
    fused = [("mat2", list(q), mat) for q, mat in out]
    fused += [("mat1", [q], pending[q]) for q in range(n) if pending[q] is not None]
    return fused


class CpuReference:
    """CPU restatement of the reference statevector path (oracle/sv_oracle.c): every cost evaluation re-simulates ALL
    G gates of full_circuit from |0..0>, one sweep over the 2^n amplitudes per gate, no gate fusion
    (aer_sv_backend.py:37-47 asks Aer for exactly that; Aer's own fusion pass would merge some neighbouring gates).
    The thread count is set explicitly (torchrun exports OMP_NUM_THREADS=1 to its workers)."""

    def __init__(self, n, target, ansatz, fused=True):
        from oracle import sv_oracle as orc
        from oracle.oracle_backends import circuit_to_gates
        self.orc, self.n = orc, n
        self.gates = circuit_to_gates(target) + circuit_to_gates(ansatz)
        self.circuit_gates = len(self.gates)
        self.fused = fused
        if fused:
            self.gates = fuse_two_qubit_blocks(n, self.gates)
        self.G = len(self.gates)
        self.threads = host_threads()
        orc.lib().orc_set_num_threads(self.threads)
        self.cores = orc.num_threads()
        self.psi = np.zeros(1 << n, dtype=np.complex128)
        self.pos = 0                 # gates of the current evaluation already applied
        self.evaluations = []        # cost of every COMPLETED evaluation

    def _apply(self, lo, hi):
        orc = self.orc
        rec, mats = orc.pack_gates(self.gates[lo:hi])
        dp = orc.ctypes.POINTER(orc.ctypes.c_double)
        rc = orc.lib().orc_sv_apply(self.n, self.psi.view(np.float64).ctypes.data_as(dp), rec.ctypes.data, len(rec),
                                    mats.ctypes.data_as(dp))
        assert rc == 0

    def chunk(self, n_gates):
        """Apply the next n_gates gates of the evaluation stream (|0..0> is re-initialised when an evaluation starts,
        inside the timed region like in the reference); returns seconds."""
        t0 = time.perf_counter()
        left = n_gates
        while left > 0:
            if self.pos == 0:
                self.psi[:] = 0
                self.psi[0] = 1
            k = min(left, self.G - self.pos)
            self._apply(self.pos, self.pos + k)
            self.pos += k
            left -= k
            if self.pos == self.G:
                self.evaluations.append(float(1 - abs(self.psi[0]) ** 2))
                self.pos = 0
        return time.perf_counter() - t0

    def restart(self):
        self.pos = 0


def cpu_baseline_sample(n, target, ansatz, unfused_too=True):
    """cpu_baseline of the main arm: ONE complete evaluation, timed (404 gates at n = 28: ~25 s on 16 host threads)."""
    ref = CpuReference(n, target, ansatz, fused=True)
    ref.chunk(4)                     # page in the 4 GiB state, spin up the thread pool
    ref.restart()
    dt = ref.chunk(ref.G)
    out = {"value": 1.0 / dt, "unit": UNIT, "cores": ref.cores, "kind": "port", "extrapolated": False,
           "gate_fusion": "2-qubit blocks",
           "sample": f"1 complete evaluation of full_circuit from |0..0> at n={n}: {ref.circuit_gates} gates fused into {ref.G} "
                     f"dense 2-qubit sweeps (stand-in for Aer's fusion pass), {dt:.1f} s, timed, not extrapolated; cost "
                     f"{ref.evaluations[-1]:.12f}"}
    if unfused_too:
        raw = CpuReference(n, target, ansatz, fused=False)
        k = min(raw.G, 60)
        dt_raw = raw.chunk(k) * raw.G / k
        out["unfused"] = {"value": 1.0 / dt_raw, "unit": UNIT, "extrapolated": k < raw.G,
                          "sample": f"{k} of {raw.G} gates, one sweep per original gate (round-1 baseline)"}
    return out


# ---- config C4: 50-qubit random MPS at bond dimension 256 ----------------------------------------
FP64_FMA_PEAK_TFLOPS = 33.9     # measured on this pool with scripts/micro/sweep_probe.cu (profiles/probe_r01k.txt): 58.3 FMA/clk/SM


def measure_zgemm_tflops(device, n=4096, reps=5):
    """Denominator for the DMMA contractions: cuBLAS Zgemm n^3 through torch.matmul on complex128, best of
    `reps`, CUDA events, 8 n^3 real flop.  Library call used ONLY as the roofline denominator."""
    try:
        import torch
        dev = torch.device("cuda", device)
        a = torch.randn(n, n, dtype=torch.complex128, device=dev)
        b = torch.randn(n, n, dtype=torch.complex128, device=dev)
        torch.matmul(a, b)
        torch.cuda.synchronize(dev)
        best = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        del a, b
        torch.cuda.empty_cache()
        return 8.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:  # noqa: BLE001
        return None


def mps_step(comp):
    lo, hi = comp.variational_circuit_range()
    comp.minimizer._reduce_cost(False, (lo, hi))


def bench_mps(args, device, with_cpu=True):
    """evals/s of the MPS path on config C4, in two truncation settings:

    capped   : max_bond_dimension = chi (Aer's cap), reference contraction order -- every evaluation
               re-applies the un-absorbed gates to the chi=256 target (one 512x512 complex SVD per
               CNOT, truncated back to chi), the same work per evaluation as
               AerMPSBackend.evaluate_global_cost (aer_mps_backend.py:49-57).
    default  : the reference's default simulator (threshold 1e-16, no cap, aer_mps_backend.py:27):
               truncation is at roundoff level, so the block transfer-matrix evaluator applies -- SVDs
               only when the optimiser moves to another layer."""
    from harness.compiler import AdaptCompiler
    from adapt_aqc_b200.mps_backend import B200MPSBackend, B200MPSSimulator
    n, chi, layers = args.mps_qubits, args.mps_chi, args.mps_layers
    t0 = time.perf_counter()
    target, ansatz = build_mps_workload(n, chi, layers)
    gen_s = time.perf_counter() - t0
    out = {"workload": f"C4: {n}-qubit random Vidal MPS, chi={chi}, {layers} un-absorbed thinly-dressed CNOT layers; "
                       "step = one Rotosolve cycle over their rotations", "target_generation_s": gen_s}
    zgemm_peak = measure_zgemm_tflops(device)
    out["denominators"] = {"fp64_fma_tflops": FP64_FMA_PEAK_TFLOPS, "fp64_fma_source": "scripts/micro/sweep_probe.cu on this pool (profiles/probe_r01k.txt)",
                           "cublas_zgemm_4096_tflops": zgemm_peak, "zgemm_source": "torch.matmul complex128 4096^3, best of 5, measured in this run"}
    from harness.minimiser import B200CostMinimiser
    for mode, cap in (("capped", chi), ("capped_batched", chi), ("default", None)):
        sim = B200MPSSimulator(1e-16, max_chi=cap, device=device)
        backend = B200MPSBackend(sim)
        backend.batch_truncating = mode == "capped_batched"      # (off by default: measured slower, see mps_backend.py)
        # capped_batched: the batched front end under real truncation -- the 3 shift values of a gate are independent
        # simulations run concurrently on worker contexts (B200MPSBackend._shift_costs_truncating)
        comp = AdaptCompiler(target, backend=backend, minimiser_cls=B200CostMinimiser if mode == "capped_batched" else None)
        comp.full_circuit.data.extend(ansatz.copy().data)
        comp.evaluate_cost()
        ctx = sim.context()
        for _ in range(max(1, args.warmup // 2)):
            mps_step(comp)
        ctx.sync()
        c0, e0 = ctx.counters(), comp.cost_evaluation_counter
        f0 = sum(m.stats()["svd_flops"] for m in ctx._live)
        ctx.profile(True)
        ctx.mark(0)
        steps = max(1, args.steps // 2) if mode.startswith("capped") else max(2, args.steps)
        t_wall = time.perf_counter()
        for _ in range(steps):
            mps_step(comp)
        ctx.mark(1)
        ms = ctx.elapsed_ms()
        if mode == "capped_batched":
            # the candidate simulations run on the worker contexts' own streams (every shift_costs call returns after
            # reading their amplitudes back, so the wall clock around the loop covers all device work)
            ms = 1e3 * (time.perf_counter() - t_wall)
        prof = ctx.profile_read()
        ctx.profile(False)
        c1 = ctx.counters()
        evals = comp.cost_evaluation_counter - e0
        res = {
            "truncation": {"threshold": 1e-16, "max_bond_dimension": cap},
            "evaluation": {"capped": "reference contraction order, one scalar per call",
                           "capped_batched": "reference contraction order, the shift values of a gate simulated concurrently",
                           "default": "block transfer matrices"}[mode],
            "value": evals / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "evals_per_step": evals // steps,
            "gpu_launches": int(c1["launches"] - c0["launches"]),
            "kernel_ms": prof["mps"][0] + prof["svd"][0] + prof["gemm"][0],
            "kernel_launches": int(prof["mps"][1] + prof["svd"][1] + prof["gemm"][1]),
            "kernel_ms_by_class": {"jacobi_svd": prof["svd"][0], "dmma_gemm": prof["gemm"][0], "other": prof["mps"][0]},
            "dmma_flops": int(c1["tensor_flops"] - c0["tensor_flops"]),
            "h2d_bytes_per_step": (c1["h2d_bytes"] - c0["h2d_bytes"]) / steps,
            "d2h_bytes_per_step": (c1["d2h_bytes"] - c0["d2h_bytes"]) / steps,
        }
        if mode == "capped_batched":
            res["timing"] = "wall clock (device work is spread over worker contexts); kernel classes / rooflines: see `capped`"
            res["workers"] = backend.SHIFT_WORKERS
            out[mode] = res
            continue
        svd_flops = sum(m.stats()["svd_flops"] for m in ctx._live) - f0
        svd_ms, gemm_ms = prof["svd"][0], prof["gemm"][0]
        # the dominant kernel of this mode and the roof that bounds it (FP64 FMA pipe for the Jacobi SVD, FP64 tensor
        # cores for the contractions; both against measured denominators)
        rl_svd = {"bound": "fp64", "kernel": "jacobi_block_kernel / jacobi_cta_kernel", "unit": "TFLOP/s",
                  "achieved": svd_flops / (svd_ms * 1e-3) / 1e12 if svd_ms > 0 else None, "peak": FP64_FMA_PEAK_TFLOPS,
                  "flop_model": "(40 p + 28 q) flop per column pair per sweep, counted by the library", "launches": int(prof["svd"][1]),
                  "share_of_kernel_time": svd_ms / max(1e-9, res["kernel_ms"])}
        rl_svd["frac"] = rl_svd["achieved"] / rl_svd["peak"] if rl_svd["achieved"] else None
        rl_gemm = {"bound": "tensor", "kernel": "zgemm_dmma_kernel / zgemm_dmma64_kernel", "unit": "TFLOP/s",
                   "achieved": res["dmma_flops"] / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None, "peak": zgemm_peak,
                   "peak_kind": "measured cuBLAS Zgemm 4096^3", "launches": int(prof["gemm"][1]),
                   "share_of_kernel_time": gemm_ms / max(1e-9, res["kernel_ms"])}
        rl_gemm["frac"] = rl_gemm["achieved"] / zgemm_peak if (rl_gemm["achieved"] and zgemm_peak) else None
        res["roofline"] = rl_svd if svd_ms >= gemm_ms else rl_gemm
        res["roofline_other"] = rl_gemm if svd_ms >= gemm_ms else rl_svd
        if backend._engine is not None:
            res["max_bond"] = max(max(m.bond_dims()) for m in backend._engine.slots)
            res["svd"] = {k: int(sum(m.stats()[k] for m in backend._engine.slots)) for k in ("svds", "jacobi_sweeps")}
        if with_cpu:
            from oracle.oracle_backends import OracleMPSBackend
            from oracle import mps_oracle as mo
            ocomp = AdaptCompiler(target, backend=OracleMPSBackend(mo.OracleMPSSimulator(1e-16, cap)))
            ocomp.full_circuit.data.extend(ansatz.copy().data)
            t0 = time.perf_counter()
            k = 0
            while k < 2 or (time.perf_counter() - t0 < 6 and k < 20):
                ocomp.evaluate_cost()
                k += 1
            dt = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": k / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"{k} full evaluations (numpy/LAPACK MPS oracle, {dt:.1f} s)"}
        out[mode] = res
        if backend._engine is not None:
            backend._engine.close()
    return out


# ---- compile wall-time (BASELINE metric, second half): a whole ADAPT-AQC compile of C3 -----------
def bench_compile_converging(args, device, pair_comm=None, cpu_s_per_fused_sweep=None):
    """Compile WALL-TIME on a target ADAPT-AQC actually compiles (harness.workloads.compilable_target): the run ends by
    reaching the reference's sufficient cost 1e-2 (adapt_compiler.py:368-393), default AdaptConfig, default all-to-all
    coupling map (P = 378 pair RDMs per layer).  pair_comm: divide the pair-RDM passes over the ranks."""
    from adapt_aqc_b200.backends import B200SVBackend
    from harness.compiler import AdaptCompiler, AdaptConfig
    from harness.workloads import compilable_target
    n = args.qubits
    target = compilable_target(n, args.converging_layers)
    backend = B200SVBackend(device=device, pair_comm=pair_comm)
    comp = AdaptCompiler(target, backend=backend, adapt_config=AdaptConfig(max_layers=args.converging_max_layers))
    comp.evaluate_cost()
    eng = backend._engine
    eng.sync()
    l0 = sum(e.counters()["launches"] for e in backend.engines())
    eng.profile(True)
    t0 = time.perf_counter()
    res = comp.compile()
    eng.sync()
    wall = time.perf_counter() - t0
    prof = eng.profile_read()
    eng.profile(False)
    l1 = sum(e.counters()["launches"] for e in backend.engines())
    out = {"workload": f"{n}-qubit compilable target (ry layer + {args.converging_layers} thin layers, seed 1234), default AdaptConfig, "
                       f"all-to-all coupling map (P={n * (n - 1) // 2})" + (f", pair-RDM passes divided over {pair_comm.world} ranks" if pair_comm else ""),
           "wall_s": wall, "layers": len(res.qubit_pair_history), "cost_evaluations": int(comp.cost_evaluation_counter),
           "final_global_cost": float(res.global_cost_history[-1]), "overlap": float(res.overlap),
           "exact_overlap": float(res.exact_overlap), "converged": bool(res.global_cost_history[-1] < comp.adapt_config.sufficient_cost),
           "evals_per_s": comp.cost_evaluation_counter / wall, "gpu_launches": int(l1 - l0),
           "kernel_ms": {k: round(v[0], 2) for k, v in prof.items() if v[1]},
           "kernel_launches": {k: int(v[1]) for k, v in prof.items() if v[1]},
           "pair_history_head": [list(p) for p in res.qubit_pair_history[:8]]}
    if cpu_s_per_fused_sweep:
        # what the reference's algorithm would cost on the host cores: one full re-simulation per evaluation and one per
        # candidate pair per layer (adapt_compiler.py:964-975), each = the fused 2-qubit sweeps of the circuit at that point
        from oracle.oracle_backends import circuit_to_gates
        layers = len(res.qubit_pair_history)
        n_target = len(fuse_two_qubit_blocks(n, circuit_to_gates(target)))
        sims = int(comp.cost_evaluation_counter) + layers * len(comp.coupling_map)
        avg_sweeps = n_target + layers / 2.0                     # the ansatz grows by one fused block per layer
        out["cpu_estimate_s"] = sims * avg_sweeps * cpu_s_per_fused_sweep
        out["cpu_estimate_note"] = (f"ESTIMATE, not timed: {sims} full re-simulations (evaluations + one per candidate pair per layer) x "
                                    f"~{avg_sweeps:.0f} fused 2-qubit sweeps x {cpu_s_per_fused_sweep * 1e3:.0f} ms per sweep measured on this box "
                                    "(cpu_baseline)")
    for e in backend.engines():
        e.close()
    return out


def bench_compile_split(args, local_rank, world):
    """N > 1: the C3 all-to-all compile with every rank running the same compiler on a replica of U|0> and the ISL
    pair-RDM read passes divided over the ranks (b200_sv_pair_rdm_part + one all-reduce of P x 16 complex per layer)."""
    import torch
    import torch.distributed as dist
    from adapt_aqc_b200.backends import B200SVBackend
    from adapt_aqc_b200.dist_sv import TorchComm
    from harness.compiler import AdaptCompiler, AdaptConfig
    n = args.qubits
    target, _ = build_workload(n, args.depth, 0)
    out = {}
    for name, comm in (("divided", TorchComm(torch.device("cuda", local_rank))), ("undivided", None)):
        backend = B200SVBackend(device=local_rank, pair_comm=comm)
        comp = AdaptCompiler(target, backend=backend, adapt_config=AdaptConfig(max_layers=args.compile_layers, method="ISL"))
        comp.evaluate_cost()
        eng = backend._engine
        eng.sync()
        dist.barrier()
        eng.profile(True)
        t0 = time.perf_counter()
        res = comp.compile()
        eng.sync()
        dist.barrier()
        wall = time.perf_counter() - t0
        prof = eng.profile_read()
        eng.profile(False)
        t = torch.tensor([wall, prof["rdm"][0]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = {"wall_s": float(t[0]), "rdm_kernel_ms_max_over_ranks": float(t[1]), "rdm_launches_this_rank": int(prof["rdm"][1]),
                     "layers": len(res.qubit_pair_history), "cost_evaluations": int(comp.cost_evaluation_counter),
                     "pair_history": [list(p) for p in res.qubit_pair_history]}
        for e in backend.engines():
            e.close()
    out["same_pairs"] = out["divided"]["pair_history"] == out["undivided"]["pair_history"]
    out["workload"] = (f"C3 compile, {n}-qubit brickwork(depth={args.depth}) target, AdaptConfig(max_layers={args.compile_layers}, "
                       f"method='ISL'), all-to-all coupling map (P={n * (n - 1) // 2}); every rank runs the same compile on a replica")
    return out


def bench_readme_configs(args, device, with_cpu=True):
    """BASELINE configs C1 (README 3-qubit statevector example) and C2 (README 50-qubit MPS example): whole
    AdaptCompiler.compile() runs with the reference's defaults, TIMED on the B200 backends and -- same harness, same
    circuits -- on the CPU oracle backends (the restatement of the Aer path).  These are the parity configs: launch-latency
    bound, the GPU is not expected to win; pair histories and final costs must agree."""
    from adapt_aqc_b200.backends import B200SVBackend
    from adapt_aqc_b200.mps_backend import B200MPSBackend
    from harness.circuit import Circuit
    from harness.compiler import AdaptCompiler
    c1 = Circuit(3)
    c1.rx(1.23, 0); c1.cx(0, 1); c1.ry(2.5, 1); c1.rx(-1.6, 2); c1.ccx(2, 1, 0)
    n = 50
    c2 = Circuit(n)
    c2.h(0); c2.cx(0, 1); c2.h(2); c2.cx(2, 3); c2.h(list(range(4, n)))
    out = {}
    for name, qc, make_gpu, make_cpu in (
            ("c1_readme_3q_sv", c1, lambda: B200SVBackend(device=device), "sv"),
            ("c2_readme_50q_mps", c2, lambda: B200MPSBackend(device=device), "mps")):
        r = {}
        backends = [("b200", make_gpu)]
        if with_cpu:
            from oracle.oracle_backends import OracleMPSBackend, OracleSVBackend
            backends.append(("cpu_oracle", OracleSVBackend if make_cpu == "sv" else OracleMPSBackend))
        hist = {}
        for tag, mk in backends:
            comp = AdaptCompiler(qc, backend=mk())
            t0 = time.perf_counter()
            res = comp.compile()
            wall = time.perf_counter() - t0
            r[tag] = {"wall_s": wall, "layers": len(res.qubit_pair_history), "cost_evaluations": int(res.cost_evaluations),
                      "final_global_cost": float(res.global_cost_history[-1]), "evals_per_s": res.cost_evaluations / wall}
            hist[tag] = (res.qubit_pair_history, res.global_cost_history)
        if with_cpu:
            r["same_pairs"] = hist["b200"][0] == hist["cpu_oracle"][0]
            r["max_cost_difference"] = float(max(abs(a - b) for a, b in zip(hist["b200"][1], hist["cpu_oracle"][1])))
            r["cpu_cores"] = host_threads()
        out[name] = r
    out["note"] = ("whole compiles with default AdaptConfig and the default all-to-all coupling map; cpu_oracle = the same compile loop on "
                   "the CPU restatement of the Aer path (oracle/), timed on this box")
    return out


def bench_compile(args, device, cpu_evals_per_s=None):
    """AdaptCompiler.compile() (adapt_compiler.py:246) on the C3 target: ISL pair selection from pair
    RDMs on a linear coupling map (P = n - 1 pairs), Rotoselect on each new layer, Rotosolve over the
    window, `--compile-layers` layers.  Wall time includes every host-side step (circuit edits, plan
    building, 4x4 measures); the final exact overlap is computed on the device."""
    from adapt_aqc_b200.backends import B200SVBackend
    from harness.compiler import AdaptCompiler, AdaptConfig
    from harness.minimiser import B200CostMinimiser
    n = args.qubits
    target, _ = build_workload(n, args.depth, 0)
    linear = [(i, i + 1) for i in range(n - 1)]
    out = {"workload": f"C3 compile: {n}-qubit brickwork(depth={args.depth}) target, AdaptConfig(max_layers={args.compile_layers}, "
                       f"method='ISL'); coupling map linear (P={n - 1}) or the reference default all-to-all (P={n * (n - 1) // 2})"}
    for name, mcls, cmap in (("linear_batched", B200CostMinimiser, linear), ("linear_reference_minimiser", None, linear),
                             ("all_to_all_reference_minimiser", None, None)):
        backend = B200SVBackend(device=device)
        comp = AdaptCompiler(target, backend=backend, coupling_map=cmap, minimiser_cls=mcls,
                             adapt_config=AdaptConfig(max_layers=args.compile_layers, method="ISL"))
        comp.evaluate_cost()              # U|0> once (also the reference's first call, adapt_compiler.py:334)
        eng = backend._engine
        eng.sync()
        c0 = eng.counters()
        l0 = sum(e.counters()["launches"] for e in backend.engines())
        eng.profile(True)
        t0 = time.perf_counter()
        res = comp.compile()
        eng.sync()
        wall = time.perf_counter() - t0
        prof = eng.profile_read()
        eng.profile(False)
        c1 = eng.counters()
        l1 = sum(e.counters()["launches"] for e in backend.engines())
        layers = len(res.qubit_pair_history)
        evals = int(comp.cost_evaluation_counter)
        r = {"wall_s": wall, "layers": layers, "cost_evaluations": evals, "final_global_cost": float(res.global_cost_history[-1]),
             "overlap": float(res.overlap), "evals_per_s": evals / wall, "gpu_launches": int(l1 - l0),
             "sweeps": int(c1["sweeps"] - c0["sweeps"]),
             "kernel_ms": {k: round(v[0], 2) for k, v in prof.items() if v[1]},
             "kernel_launches": {k: int(v[1]) for k, v in prof.items() if v[1]}}
        if cpu_evals_per_s:
            # the reference re-simulates everything per evaluation and once more per candidate pair per layer
            sims = evals + layers * len(comp.coupling_map)
            r["cpu_estimate_s"] = sims / cpu_evals_per_s
            r["cpu_estimate_note"] = (f"{sims} full re-simulations (evaluations + one per candidate pair per layer, "
                                      "adapt_compiler.py:964-975) at the measured CPU rate of this box; estimate, not timed")
        out[name] = r
        backend._engine.close()
    return out


# ---- config C5: statevector sharded over the ranks by global qubits --------------------------------
def bench_sharded(args, local_rank, world):
    """One cost evaluation = full_circuit (brickwork target + thin layers) applied to |0..0> on a
    state of n = sharded_local_qubits + log2(world) qubits, + amplitude 0.  Local gates run in the
    fused sweep kernels; global qubits are exchanged with NCCL send/recv over NVLink."""
    import torch
    import torch.distributed as dist
    from adapt_aqc_b200.dist_sv import make_gpu_sharded
    from adapt_aqc_b200.gates import canonical_window
    from adapt_aqc_b200.lib import B200Error
    from adapt_aqc_b200.sv_engine import SVEngine
    g = int(np.log2(world))
    # BASELINE config 5: 34 qubits at 2 / 4 / 8 GPUs = 128 / 64 / 32 GiB per GPU (one slot, exchanged in place over
    # NVLink peer memory).  The allocation is agreed on collectively: if any rank cannot hold its slice, all fall back
    # to one qubit less (reported in "qubits").
    n = args.sharded_qubits if args.sharded_local_qubits is None else args.sharded_local_qubits + g
    sv = None
    while sv is None:
        try:
            local_engine = SVEngine(n - g, device=local_rank, n_slots=1)
            ok = 1.0
        except B200Error:
            local_engine, ok = None, 0.0
        flag = torch.tensor([ok], device="cuda", dtype=torch.float64)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if float(flag[0]) > 0:
            sv = make_gpu_sharded(n, n_slots=1, local_rank=local_rank, engine=local_engine)
        else:
            if local_engine is not None:
                local_engine.close()
            n -= 1
            if n - g < 26:
                raise RuntimeError("no sharded register size fits")
    target, ansatz = build_workload(n, args.sharded_depth, args.sharded_layers)
    window = canonical_window(target) + canonical_window(ansatz)
    eng, comm = sv.eng, sv.comm
    sv.run(0, -1, window)                       # warm-up (also JIT of NCCL channels)
    sv.amp(0, 0)
    eng.sync(); torch.cuda.synchronize()
    dist.barrier()
    comm.bytes_sent, comm.exchange_ms = 0, 0.0
    sv.stats["exchanges"] = 0
    c0 = eng.counters()
    reps = max(1, args.steps)
    eng.profile(True)
    t0 = time.perf_counter()
    for _ in range(reps):
        sv.run(0, -1, window)
        a0 = sv.amp(0, 0)
    eng.sync(); torch.cuda.synchronize()
    dist.barrier()
    wall = time.perf_counter() - t0
    prof = eng.profile_read()
    eng.profile(False)
    c1 = eng.counters()
    z, norm = sv.expz(0)
    t = torch.tensor([wall, comm.exchange_ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, xms = float(t[0]), float(t[1])
    sweeps = (c1["sweeps"] - c0["sweeps"]) / reps
    local_bytes = 16 * (1 << (n - g))
    out = {
        "workload": f"C5: {n}-qubit brickwork(depth={args.sharded_depth}) + {args.sharded_layers} thin layers from |0..0>, "
                    f"{n - g} local + {g} global qubits per rank",
        "qubits": n, "exchange": getattr(sv, "exchange_mode", "nccl"), "state_bytes_per_gpu": local_bytes, "value": reps / wall, "unit": UNIT, "s_per_eval": wall / reps,
        "sweeps_per_eval": sweeps, "exchanges_per_eval": sv.stats["exchanges"] / reps,
        "nvlink_bytes_sent_per_gpu_per_eval": comm.bytes_sent / reps,
        "exchange_ms_per_eval": xms / reps,
        "nvlink_GBps_per_direction": (comm.bytes_sent / reps) / (xms / reps * 1e-3) / 1e9 if xms > 0 else None,
        "sweep_ms_avg": prof["sweep"][0] / max(1, prof["sweep"][1]),
        "sweep_GBps": 32.0 * (1 << (n - g)) / (prof["sweep"][0] / max(1, prof["sweep"][1]) * 1e-3) / 1e9 if prof["sweep"][1] else None,
        "norm": norm, "amp0_abs2": abs(a0) ** 2,
    }
    sv.close()
    del sv
    torch.cuda.empty_cache()
    if not args.no_compile:
        # the compile keeps 4 sharded slots (WORK, BASE, L, R): 34 qubits at 8 GPUs (4 x 32 GiB), 33 at 4, 32 at 2
        out["compile"] = bench_sharded_compile(args, local_rank, min(n, 31 + g))
    return out


def bench_sharded_compile(args, local_rank, n):
    """AdaptCompiler.compile() on the sharded register (B200ShardedSVBackend): U|0> resident across the ranks,
    ISL pair selection from sharded pair-RDM passes on a linear map, Rotoselect / Rotosolve in the projected
    tail (one gather + all-reduce of 2^K amplitudes per projection, then a K-qubit engine per rank)."""
    import torch.distributed as dist
    from harness.compiler import AdaptCompiler, AdaptConfig
    from adapt_aqc_b200.dist_sv import B200ShardedSVBackend
    target, _ = build_workload(n, args.sharded_depth, 0)
    backend = B200ShardedSVBackend(local_rank)
    comp = AdaptCompiler(target, backend=backend, coupling_map=[(i, i + 1) for i in range(n - 1)],
                         adapt_config=AdaptConfig(max_layers=args.sharded_compile_layers, method="ISL"))
    comp.evaluate_cost()
    backend._engine.sync()
    dist.barrier()
    sv = backend._engine.sv
    sv.stats["exchanges"] = 0
    t0 = time.perf_counter()
    res = comp.compile()
    backend._engine.sync()
    dist.barrier()
    wall = time.perf_counter() - t0
    st = dict(backend._evaluator.stats)
    out = {"workload": f"{n}-qubit brickwork(depth={args.sharded_depth}) target, AdaptConfig(max_layers={args.sharded_compile_layers}, "
                       f"method='ISL'), linear coupling map (P={n - 1}), register sharded over the ranks",
           "wall_s": wall, "layers": len(res.qubit_pair_history), "cost_evaluations": int(comp.cost_evaluation_counter),
           "final_global_cost": float(res.global_cost_history[-1]), "exchanges": int(sv.stats["exchanges"]),
           "projections": st.get("projections"), "projected_evals": st.get("projected_evals"),
           "resimulations": st.get("resimulations", 0)}
    for e in backend.engines():
        e.close()
    return out


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU algorithm for the path (full re-simulation per evaluation) on the host
    cores, as TIMED work: the K timed steps are consecutive chunks of ceil(G/K) gates of real evaluations, so that
    together they cover at least one complete G-gate evaluation; value = evaluations completed per second of timed
    work (fractions count by gates), nothing is extrapolated unless K * chunk < G (reported)."""
    if rank != 0:
        return
    n = args.qubits
    target, ansatz = build_workload(n, args.depth, args.layers)
    t_all = time.perf_counter()
    ref = CpuReference(n, target, ansatz)
    per_step = max(1, -(-ref.G // max(1, args.steps)))
    for _ in range(args.warmup):
        ref.chunk(min(per_step, 8))              # warm-up: page in the state, spin up the threads
    ref.restart()
    ref.evaluations.clear()
    times = [ref.chunk(per_step) for _ in range(args.steps)]
    total = float(sum(times))
    gates_done = per_step * args.steps
    v = gates_done / ref.G / total
    extrapolated = gates_done < ref.G
    base = {"value": v, "unit": UNIT, "cores": ref.cores, "kind": "port", "extrapolated": extrapolated,
            "extrapolation_factor": ref.G / gates_done if extrapolated else 1.0,
            "evaluations_timed": gates_done / ref.G, "gate_fusion": "2-qubit blocks" if ref.fused else "none",
            "sample": f"{args.steps} timed steps x {per_step} fused gates = {gates_done / ref.G:.2f} complete evaluations of "
                      f"full_circuit ({ref.circuit_gates} gates fused into {ref.G} dense 2-qubit sweeps, the stand-in for Aer's "
                      f"fusion pass) at n={n} ({total:.1f} s of timed work); completed-evaluation costs "
                      f"{[round(c, 12) for c in ref.evaluations[:3]]}"}
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
        "config": workload_config(args), "cpu_baseline": base,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
        "step_definition": f"one step = {per_step} consecutive gates of a full re-simulation ({per_step / ref.G:.4f} evaluations)",
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


def workload_config(args):
    return {"workload": f"C3: {args.qubits}-qubit brickwork(depth={args.depth}, seed=1234) target + {args.layers} "
                        "thinly-dressed CNOT layers; step = Rotoselect(last layer) + Rotosolve cycle(all rotations)",
            "qubits": args.qubits, "target_depth": args.depth, "ansatz_layers": args.layers,
            "evals_per_step": 220 if args.layers == 16 else None,
            "l2": "inputs larger than L2 (each statevector is 16*2^n bytes)",
            "parallelism": f"replicas x{args.gpus}" if args.gpus > 1 else "single GPU"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--layers", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mps", action="store_true", help="skip the secondary C4 (MPS) measurement")
    ap.add_argument("--mps-only", action="store_true", help="run only the C4 (MPS) measurement and print it")
    ap.add_argument("--sharded-qubits", type=int, default=34, help="register size of the sharded C5 leg (N > 1)")
    ap.add_argument("--sharded-local-qubits", type=int, default=None, help="override: qubits per rank of the sharded C5 leg")
    ap.add_argument("--sharded-depth", type=int, default=4)
    ap.add_argument("--sharded-layers", type=int, default=4)
    ap.add_argument("--no-sharded", action="store_true")
    ap.add_argument("--no-compile", action="store_true", help="skip the compile wall-time leg")
    ap.add_argument("--compile-layers", type=int, default=6)
    ap.add_argument("--converging-layers", type=int, default=6, help="thin layers of the compilable target")
    ap.add_argument("--converging-max-layers", type=int, default=400)
    ap.add_argument("--no-converging", action="store_true", help="skip the converging compile leg")
    ap.add_argument("--no-readme", action="store_true", help="skip the C1 / C2 README compile legs")
    ap.add_argument("--sharded-compile-layers", type=int, default=4)
    ap.add_argument("--mps-qubits", type=int, default=50)
    ap.add_argument("--mps-chi", type=int, default=256)
    ap.add_argument("--mps-layers", type=int, default=2)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # stdout carries exactly ONE JSON line (rank 0): native libraries that write to fd 1 (NCCL prints its
    # version banner there) are sent to stderr, the result line goes to a private copy of the real stdout
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import adapt_aqc_b200  # noqa: F401
    from adapt_aqc_b200.backends import B200SVBackend

    if args.mps_only:
        print(json.dumps(bench_mps(args, local_rank, with_cpu=not args.no_cpu_baseline)), file=RESULT_OUT, flush=True)
        return

    dist = None
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL's own messages (version banner, warnings) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    n = args.qubits
    target, ansatz = build_workload(n, args.depth, args.layers)

    # ---- value leg: batched front end, HBM-resident ---------------------------------------------
    backend = B200SVBackend(device=local_rank)
    comp = make_compiler(target, ansatz, backend, batched=True)
    comp.evaluate_cost()                     # builds U|0> (slot BASE) and allocates the engine
    eng = backend._engine
    for _ in range(args.warmup):
        one_step(comp)
    eng.sync()
    sampler = ClockSampler(local_rank)

    def all_launches():
        return sum(e.counters()["launches"] for e in backend.engines())

    c0 = eng.counters()
    l0 = all_launches()
    e0 = comp.cost_evaluation_counter
    barrier()
    sampler.start()
    eng.profile(True)
    eng.mark(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step(comp)
    eng.mark(1)
    dev_ms = eng.elapsed_ms()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    prof = eng.profile_read()
    eng.profile(False)
    clocks = sampler.stop()
    c1 = eng.counters()
    evals = comp.cost_evaluation_counter - e0
    launches = all_launches() - l0       # register engine + the compact / projected engines

    # ---- e2e leg: the reference-facing one-scalar-per-call interface ----------------------------
    backend2 = backend            # same engine / same cached U|0>: the interface is what changes
    comp2 = make_compiler(target, ansatz, backend2, batched=False)
    comp2.evaluate_cost()
    for _ in range(args.warmup):
        one_step(comp2)
    eng.sync()
    d0 = eng.counters()
    f0 = comp2.cost_evaluation_counter
    barrier()
    eng.mark(0)
    for _ in range(args.steps):
        one_step(comp2)
    eng.mark(1)
    e2e_ms = eng.elapsed_ms()
    d1 = eng.counters()
    e2e_evals = comp2.cost_evaluation_counter - f0

    # ---- max over ranks ------------------------------------------------------------------------
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])
        cnt = torch.tensor([evals, e2e_evals, launches], device="cuda", dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        evals, e2e_evals, launches = float(cnt[0]), float(cnt[1]), int(cnt[2])

    value = evals / (dev_ms * 1e-3)
    e2e_value = e2e_evals / (e2e_ms * 1e-3)

    peak, peak_kind = measured_peak_hbm()

    def hbm_roofline(cls, kernel, bytes_per_amp):
        ms, cnt = prof[cls]
        alg = float(bytes_per_amp) * (1 << n)
        ach = alg / (ms / max(1, cnt) * 1e-3) / 1e9 if cnt else None
        return {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": ach / peak if ach else None, "traffic": measured_traffic(n, kernel), "bytes_per_launch": alg,
                "launches": int(cnt), "avg_launch_ms": ms / max(1, cnt), "share_of_step": ms / dev_ms if world == 1 else None}

    # the register-sized kernels of the step: the gate sweep (read + write: 32 B per amplitude) and the fused sweep +
    # transfer pass in its three forms (read two states, write one: 48 B; T only: two reads, 32 B; from an embedded source:
    # one read, one write, 32 B; from an embedded source and T only: ONE read, 16 B) and the sweep that keeps only the projected
    # amplitudes (one read, 16 B).  The one with the most device time leads, the others follow under `also`.
    cands = [hbm_roofline("sweep", "sv_sweep_kernel", 32),
             hbm_roofline("fused", "sv_sweep_inner2_kernel (store)", 48),
             hbm_roofline("fused_read", "sv_sweep_inner2_kernel (T only)", 32),
             hbm_roofline("fused_embed", "sv_sweep_inner2_kernel (embedded source)", 32),
             hbm_roofline("project", "sv_sweep_project_kernel", 16),
             hbm_roofline("fused_embed_read", "sv_sweep_inner2_kernel (embedded source, T only)", 16)]
    for c in cands:
        if c["kernel"].endswith("(embedded source, T only)"):
            # nominal bytes = one read of the ket; tiles of the embedded bra that are zero as a whole are neither formed nor
            # read, so the DRAM traffic of this class is LOWER than that (half of it for three of the four head blocks of
            # C3): `achieved` / `frac` are effective rates on the nominal bytes, not a bandwidth utilisation
            c["note"] = "effective rate on the nominal bytes (one read of the ket); all-zero tiles are skipped, DRAM traffic is lower"
    cands = [c for c in cands if c["launches"]] or cands[:1]
    cands.sort(key=lambda c: -(c["avg_launch_ms"] * c["launches"]))
    roofline = dict(cands[0])
    if len(cands) > 1:
        roofline["also"] = cands[1:]
    roofline["other_kernels_ms"] = {k: round(v[0], 3) for k, v in prof.items()
                                    if k not in ("sweep", "fused", "fused_read", "fused_embed", "project", "fused_embed_read") and v[1]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex128", "data": "synthetic", "config": workload_config(args),
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": (d1["h2d_bytes"] - d0["h2d_bytes"]) / args.steps,
                "d2h_bytes_per_step": (d1["d2h_bytes"] - d0["d2h_bytes"]) / args.steps,
                "ms_per_step": e2e_ms / args.steps, "evals_per_step": e2e_evals / args.steps / max(1, world)},
        "roofline": roofline, "host_wall_ms_per_step": wall_ms / args.steps,
        "evaluator_stats": dict(backend._evaluator.stats),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(n, target, ansatz)
    if rank == 0 and world == 1 and not args.no_compile:
        backend._engine.close()
        try:
            cpu_rate = line.get("cpu_baseline", {}).get("value")
            line["compile_c3"] = bench_compile(args, local_rank, cpu_rate)
        except Exception as exc:  # noqa: BLE001
            line["compile_c3"] = {"error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_compile and not args.no_converging:
        try:
            cb = line.get("cpu_baseline") or {}
            per_sweep = (1.0 / cb["value"] / 124.0) if (cb.get("value") and args.qubits == 28 and args.depth == 8 and args.layers == 16) else None
            line["compile_converging"] = bench_compile_converging(args, local_rank, cpu_s_per_fused_sweep=per_sweep)
        except Exception as exc:  # noqa: BLE001
            line["compile_converging"] = {"error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_compile and not args.no_readme:
        try:
            line["compile_c1_c2"] = bench_readme_configs(args, local_rank, with_cpu=not args.no_cpu_baseline)
        except Exception as exc:  # noqa: BLE001
            line["compile_c1_c2"] = {"error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_mps:
        if backend._engine is not None:
            backend._engine.close()          # free the 16 GiB of statevector slots first
        try:
            line["mps_c4"] = bench_mps(args, local_rank, with_cpu=not args.no_cpu_baseline)
        except Exception as exc:  # noqa: BLE001 - the secondary measurement must not hide the main line
            line["mps_c4"] = {"error": repr(exc)}
    if dist is not None and not args.no_compile:
        for e in backend.engines():
            e.close()
        try:
            line["compile_c3_pairs_divided"] = bench_compile_split(args, local_rank, world)
        except Exception as exc:  # noqa: BLE001
            line["compile_c3_pairs_divided"] = {"error": repr(exc)}
    if dist is not None and not args.no_sharded:
        for e in backend.engines():
            e.close()                    # free the replica's statevector slots first
        try:
            sharded = bench_sharded(args, local_rank, world)
        except Exception as exc:  # noqa: BLE001
            sharded = {"error": repr(exc)}
        line["sharded_c5"] = sharded
    if dist is not None:
        # the paths that actually USE several GPUs for one problem, side by side (the headline `value` of an N > 1 line is
        # N independent replicas of the C3 stream: throughput, not scaling)
        sh, pd = line.get("sharded_c5") or {}, line.get("compile_c3_pairs_divided") or {}
        line["multi_gpu_paths"] = {
            "c5_qubits": sh.get("qubits"), "c5_s_per_eval": sh.get("s_per_eval"), "c5_sweeps_per_eval": sh.get("sweeps_per_eval"),
            "c5_exchanges_per_eval": sh.get("exchanges_per_eval"), "c5_nvlink_GBps_per_direction": sh.get("nvlink_GBps_per_direction"),
            "c5_compile_wall_s": (sh.get("compile") or {}).get("wall_s"), "c5_compile_qubits": (sh.get("compile") or {}).get("workload", "")[:8],
            "c3_all_to_all_compile_wall_s": {k: (pd.get(k) or {}).get("wall_s") for k in ("divided", "undivided")},
            "c3_rdm_kernel_ms_max_over_ranks": {k: (pd.get(k) or {}).get("rdm_kernel_ms_max_over_ranks") for k in ("divided", "undivided")},
            "c3_same_pairs": pd.get("same_pairs"),
        }
    if rank == 0:
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
