"""Probe: how long does ADAPT-AQC take to COMPILE a 28-qubit target on one B200?  python scripts/converge_probe.py [n] [target layers] [max layers]
Prints per-layer progress (wall, cost, evaluations) so that bench.py's converging-compile leg can be sized."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adapt_aqc_b200  # noqa: E402,F401
from adapt_aqc_b200.backends import B200SVBackend  # noqa: E402
from harness.compiler import AdaptCompiler, AdaptConfig  # noqa: E402
from harness.workloads import compilable_target  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
tl = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ml = int(sys.argv[3]) if len(sys.argv) > 3 else 150
cmap = sys.argv[4] if len(sys.argv) > 4 else "full"
target = compilable_target(n, tl)
backend = B200SVBackend()
comp = AdaptCompiler(target, backend=backend, coupling_map=None if cmap == "full" else [(i, i + 1) for i in range(n - 1)],
                     adapt_config=AdaptConfig(max_layers=ml))
orig = comp._add_layer
t0 = time.perf_counter()


def logged(index):
    cost = orig(index)
    print(f"layer {index:3d} pair {comp.qubit_pair_history[-1]} cost {cost:.6f} evals {comp.cost_evaluation_counter} "
          f"wall {time.perf_counter() - t0:7.2f} s", flush=True)
    return cost


comp._add_layer = logged
res = comp.compile()
print("done: layers", len(res.qubit_pair_history), "evals", res.cost_evaluations, "cost", res.global_cost_history[-1],
      "overlap", res.overlap, "exact", res.exact_overlap, "wall", time.perf_counter() - t0, dict(backend._evaluator.stats))
