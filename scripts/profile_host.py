"""cProfile of the host side of the bench step (reference-facing one-scalar-per-call interface and the batched front end)
on a real GPU.  python scripts/profile_host.py [steps]"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adapt_aqc_b200  # noqa: E402,F401
import bench  # noqa: E402
from adapt_aqc_b200.backends import B200SVBackend  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
target, ansatz = bench.build_workload(28, 8, 16)
backend = B200SVBackend()
for batched in (False, True):
    comp = bench.make_compiler(target, ansatz, backend, batched=batched)
    comp.evaluate_cost()
    for _ in range(3):
        bench.one_step(comp)
    backend._engine.sync()
    e0 = comp.cost_evaluation_counter
    t0 = time.perf_counter()
    for _ in range(steps):
        bench.one_step(comp)
    backend._engine.sync()
    dt = time.perf_counter() - t0
    print(f"batched={batched}: {(comp.cost_evaluation_counter - e0) / dt:.0f} evals/s, {1e3 * dt / steps:.2f} ms/step (wall, unprofiled)")
    # device time per engine and kernel class over the same steps (CUDA events around every launch)
    for e in backend.engines():
        e.profile(True)
    t0 = time.perf_counter()
    for _ in range(steps):
        bench.one_step(comp)
    backend._engine.sync()
    dt = time.perf_counter() - t0
    tot = 0.0
    for e in backend.engines():
        p = e.profile_read(); e.profile(False)
        used = {k: (round(v[0] / steps, 3), v[1] // steps) for k, v in p.items() if v[1]}
        tot += sum(v[0] for v in p.values()) / steps
        print(f"  engine {e.num_qubits:2d} qubits: (ms, launches) per step by class {used}")
    print(f"  device total {tot:.2f} ms per step of {1e3 * dt / steps:.2f} ms wall (event-timed launches)")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(steps):
        bench.one_step(comp)
    backend._engine.sync()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
    print(s.getvalue())
