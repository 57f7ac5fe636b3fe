"""Randomised soak of SVCostEvaluator against the oracle on the CPU emulation of the kernels (tests/emu): random ansatz
structures, every engine configuration (compact / projected / nested sizes), random single and double edits in the
optimiser's patterns (final write + first shift of the next gate, cycle wrap-around), known and unknown `changed`.
Development tool (imports tests/ and oracle/): python scripts/soak_evaluator.py [seeds] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import adapt_aqc_b200  # noqa: E402,F401
from adapt_aqc_b200 import gates as G  # noqa: E402
from adapt_aqc_b200.gates import GateStream, canonical_window  # noqa: E402
from adapt_aqc_b200.sv_engine import SVCostEvaluator  # noqa: E402
from harness.circuit import Circuit  # noqa: E402
from harness.minimiser import replace_1q_gate  # noqa: E402
from oracle import sv_oracle as orc  # noqa: E402
from oracle.oracle_backends import circuit_to_gates  # noqa: E402
from helpers import FakeEngine, brickwork  # noqa: E402
from dist_worker import load_emu  # noqa: E402

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
emu = load_emu()
bad = 0
for seed in range(seeds):
    rng = np.random.default_rng(seed)
    n = 12
    target, trng = brickwork(n, 2, seed=seed)
    ansatz = Circuit(n)
    for _ in range(int(rng.integers(6, 14))):
        if rng.random() < 0.5:
            a, b = [int(x) for x in rng.choice(n, 2, replace=False)]
        else:
            a = int(rng.integers(n)); b = (a + 1) % n
        th = trng.uniform(-np.pi, np.pi, 4)
        ansatz.rz(th[0], a, label="rz"); ansatz.rz(th[1], b, label="rz"); ansatz.cx(a, b)
        ansatz.rz(th[2], a, label="rz"); ansatz.rz(th[3], b, label="rz")
    cfg = seed % 4
    compact = [FakeEngine(emu, 6, n_slots=1)] if cfg in (1, 3) else None
    proj = ([FakeEngine(emu, 8, n_slots=4)] if cfg in (2, 3)
            else ([FakeEngine(emu, 5, n_slots=4), FakeEngine(emu, 9, n_slots=4)] if cfg == 0 else None))
    ev = SVCostEvaluator(FakeEngine(emu, n), compact, proj)
    ev.set_base("t", GateStream.from_circuit(target))
    base_gates = circuit_to_gates(target)
    window = canonical_window(ansatz)
    rot = [i for i, e in enumerate(window) if e[2] < 0]
    for step in range(steps):
        idxs = sorted(set(int(rot[int(rng.integers(len(rot)))]) for _ in range(int(rng.integers(1, 3)))))
        if step % 7 == 3:       # the optimiser's pattern: final write of one gate + first shift of the next
            j = int(rng.integers(len(rot) - 1)); idxs = [rot[j], rot[j + 1]]
        if step % 11 == 5:      # cycle wrap-around
            idxs = [rot[0], rot[-1]]
        for idx in idxs:
            replace_1q_gate(ansatz, idx, ["rx", "ry", "rz"][int(rng.integers(3))], float(rng.uniform(-np.pi, np.pi)))
            window[idx] = canonical_window(ansatz, idx, idx + 1)[0]
        changed = idxs if rng.random() < 0.85 else None
        if step % 3 == 1:       # the batched front end: every shift value of ONE gate from the same call
            k = idxs[-1]
            cands = [G.one_qubit_matrix(["rx", "ry", "rz"][int(rng.integers(3))], float(t)) for t in (0.0, np.pi / 2, -np.pi / 2)]
            gots = ev.shift_amplitudes(list(window), k, cands, changed=changed)
            refs = []
            for cm in cands:
                w2 = circuit_to_gates(ansatz)
                w2[k] = ("mat1", w2[k][1], cm)
                refs.append(orc.evaluate_circuit(n, base_gates + w2)[0])
            err = max(abs(a - b) for a, b in zip(gots, refs))
        else:
            got = ev.amp0(list(window), changed=changed)
            c = Circuit(n); c.data = list(ansatz.data)
            err = abs(got - orc.evaluate_circuit(n, base_gates + circuit_to_gates(c))[0])
        if err > 1e-10:
            bad += 1
            print("MISMATCH seed", seed, "cfg", cfg, "step", step, idxs, changed, err)
            break
print("soak done:", seeds, "seeds x", steps, "edits, mismatches:", bad)
sys.exit(1 if bad else 0)
