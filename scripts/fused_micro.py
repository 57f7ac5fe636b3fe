"""Micro-benchmark of the fused sweep + transfer pass (b200_sv_run_inner2) against the two separate calls, on the C3
register: device time per class (CUDA events around each launch) and the HBM rate each achieves on its ALGORITHMIC bytes
(sweep 32 * 2^n, transfer pass 32 * 2^n, fused 48 * 2^n).  python scripts/fused_micro.py [n] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import adapt_aqc_b200  # noqa: E402,F401
from adapt_aqc_b200.gates import GateStream, canonical_window, invert_window  # noqa: E402
from adapt_aqc_b200.sv_engine import SVEngine  # noqa: E402
from helpers import brickwork, thin_ansatz  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
only = sys.argv[3] if len(sys.argv) > 3 else None       # "fused" / "separate" / "read" (T only, no store): one side only (ncu captures)
target, rng = brickwork(n, 8, 1234)
ansatz = thin_ansatz(n, 16, rng)
win = canonical_window(ansatz)
eng = SVEngine(n, n_slots=3)
eng.run(0, -1, GateStream.from_window(win[:40]))
eng.run(1, -1, GateStream.from_window(win[40:80]))
# the optimiser's step from one block to the next: one 5-gate block leaves the bra, its neighbour (inverted) enters
cases = {"middle(10 gates)": (GateStream.from_window(list(win[30:35]) + invert_window(win[35:40])), 13, 14),
         "layer(5 gates)": (GateStream.from_window(win[0:5]), int(win[0][1]), int(win[2][1] if win[2][2] < 0 else win[2][2])),
         "empty": (GateStream.from_window([]), 3, 17)}
for name, (gs, qa, qb) in cases.items():
    if qa == qb:
        qb = (qa + 1) % n
    rows = []
    T0 = T1 = None
    for rep in range(reps):
        if only == "read":
            eng.profile(True)
            T1, stored = eng.run_inner2(2, 0, gs, 1, qa, qb, store=False)
            assert not stored
            p = eng.profile_read(); eng.profile(False)
            rows.append((0.0, 0.0, 0.0, p["fused_read"][0], p["reduce"][0]))
            continue
        if only != "fused":
            eng.profile(True)
            eng.run(2, 0, gs)
            T0 = eng.inner2(2, 1, qa, qb)
            p = eng.profile_read(); eng.profile(False)
            sep = (p["sweep"][0] + p["fill"][0], p["inner"][0], p["reduce"][0])
        else:
            sep, T0 = (0.0, 0.0, 0.0), None
        if only != "separate":
            eng.profile(True)
            T1 = eng.run_inner2(2, 0, gs, 1, qa, qb)
            p = eng.profile_read(); eng.profile(False)
            fus = (p["fused"][0], p["reduce"][0])
        else:
            fus, T1 = (0.0, 0.0), None
        rows.append(sep + fus)
    if T0 is not None and T1 is not None:
        assert np.allclose(T0, T1, rtol=0, atol=1e-13), abs(T0 - T1).max()
    r = np.median(np.array(rows), axis=0)
    gb = 2.0 ** n / 1e9
    if only == "read":
        print(f"{name:18s} T only (two reads, no store) {r[3]:6.3f} ms ({32 * gb / max(r[3], 1e-9) * 1e3:6.0f} GB/s)")
        continue
    print(f"{name:18s} sweep {r[0]:6.3f} ms ({32 * gb / max(r[0], 1e-9) * 1e3:6.0f} GB/s)  transfer {r[1]:6.3f} ms ({32 * gb / max(r[1], 1e-9) * 1e3:6.0f} GB/s)  "
          f"separate total {r[0] + r[1] + r[2]:6.3f} ms | fused {r[3]:6.3f} ms ({48 * gb / max(r[3], 1e-9) * 1e3:6.0f} GB/s)  total {r[3] + r[4]:6.3f} ms")
# the bra's rebuild: tail built on a 24-qubit engine, head gates at full size, T against the ket -- one pass
if only is None and n >= 26:
    K = 24
    small = SVEngine(K, n_slots=1)
    small.run(0, -1, GateStream.from_window([e for e in win if max(e[1], e[2]) < K][:40]))
    small.sync()
    for label, outside in (("outside = 4 highest qubits", list(range(n - 4, n))), ("outside = qubits 0, 5, 13, n-1", [0, 5, 13, n - 1])):
        qmap = [q for q in range(n) if q not in outside]
        gs = GateStream.from_window(win[30:40])
        rows = []
        for rep in range(reps):
            eng.profile(True)
            eng.scatter(2, qmap, small, 0)
            eng.run(2, 2, gs, inverse=True)
            T0 = eng.inner2(2, 1, 13, 14)
            p = eng.profile_read(); eng.profile(False)
            sep = p["sweep"][0] + p["fill"][0] + p["inner"][0] + p["reduce"][0]
            eng.profile(True)
            T1 = eng.run_embedded(2, qmap, small, 0, gs, inverse=True, fuse=(1, 13, 14))
            p = eng.profile_read(); eng.profile(False)
            rows.append((sep, p["fused_embed"][0]))
        assert np.allclose(T0, T1, rtol=0, atol=1e-13)
        r = np.median(np.array(rows), axis=0)
        print(f"embedded bra ({label}): zero fill + scatter + sweep + transfer {r[0]:6.3f} ms | one pass {r[1]:6.3f} ms "
              f"({32 * 2.0 ** n / 1e9 / r[1] * 1e3:6.0f} GB/s on read `other` + write)")
    small.close()
eng.close()
