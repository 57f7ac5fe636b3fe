"""(B200AQC_SWEEP=direct|pipe selects the sweep kernel.)  Micro-benchmark of the fused sweep kernel on the C3 circuits: per-sweep device time next to the
plan's (rounds, ops, dense ops, contiguous low qubits).  python scripts/sweep_micro.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import adapt_aqc_b200  # noqa: E402,F401
from adapt_aqc_b200.gates import GateStream, canonical_window  # noqa: E402
from adapt_aqc_b200.sv_engine import SVEngine, plan_detail  # noqa: E402
from helpers import brickwork, thin_ansatz  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
target, rng = brickwork(n, 8, 1234)
ansatz = thin_ansatz(n, 16, rng)
eng = SVEngine(n, n_slots=2)
streams = {"target": GateStream.from_circuit(target), "ansatz": GateStream.from_circuit(ansatz),
           "layer(13,14)": GateStream.from_window(canonical_window(ansatz)[30:35]),
           "layer(0,1)": GateStream.from_window(canonical_window(ansatz)[0:5])}
only = sys.argv[3].split(";") if len(sys.argv) > 3 else None      # e.g. "ansatz" or "layer(13,14);layer(0,1)" (for ncu captures)
if only:
    eng.run(0, -1, streams["layer(13,14)"])      # any normalised state will do for timing
    streams = {k: v for k, v in streams.items() if k in only}
else:
    eng.run(0, -1, streams["target"])
for name, gs in streams.items():
    detail = plan_detail(n, gs)
    for rep in range(reps):
        eng.profile(True)
        eng.run(1, 0, gs)
        p = eng.profile_read()
        per = eng.profile_sweeps()
        eng.profile(False)
    print(f"{name:14s} sweeps={len(detail):2d} total={p['sweep'][0]:8.3f} ms  avg={p['sweep'][0] / max(1, p['sweep'][1]):7.3f} ms  "
          f"GB/s={32 * 2 ** n * p['sweep'][1] / (p['sweep'][0] * 1e-3) / 1e9:7.1f}")
    for d, ms in zip(detail, per):
        print(f"    (rounds, ops, dense, c)={d}  {ms:7.3f} ms  {32 * 2 ** n / (ms * 1e-3) / 1e9:7.1f} GB/s")
eng.close()
