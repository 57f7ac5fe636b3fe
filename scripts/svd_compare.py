"""Comparator for the on-device Jacobi SVD (K9): cuSOLVER through torch.linalg.svd on the same
(2 chi x 2 chi) complex128 two-site matrices, next to one b200_mps two-qubit gate (contraction +
Jacobi SVD + truncation + site update).  python scripts/svd_compare.py [chi ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import adapt_aqc_b200  # noqa: E402,F401
from adapt_aqc_b200.gates import GateStream  # noqa: E402
from adapt_aqc_b200.mps_engine import MPSContext  # noqa: E402
from helpers import random_vidal_mps  # noqa: E402

chis = [int(x) for x in sys.argv[1:]] or [64, 128, 256]
ctx = MPSContext(0)
for chi in chis:
    n = 2 * int(np.log2(chi)) + 6
    target = random_vidal_mps(n, chi, 1)
    m, base = ctx.new_mps(n, 1e-16, chi), ctx.new_mps(n, 1e-16, chi)
    base.set(target)
    gs = GateStream.from_gates([("cx", [n // 2 - 1, n // 2], [])])
    times = []
    for rep in range(4):
        m.copy_from(base)
        ctx.sync()
        ctx.mark(0); m.apply(gs); ctx.mark(1)
        times.append(ctx.elapsed_ms())
    st = m.stats()
    a = torch.randn(2 * chi, 2 * chi, dtype=torch.complex128, device="cuda")
    res = {}
    for drv in ("gesvdj", "gesvd"):
        ts = []
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            torch.linalg.svd(a, full_matrices=False, driver=drv)
            torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
        res[drv] = min(ts)
    t0 = time.perf_counter(); np.linalg.svd(a.cpu().numpy(), full_matrices=False); cpu = 1e3 * (time.perf_counter() - t0)
    print(f"chi={chi:4d} matrix {2*chi}x{2*chi}: b200 2q gate {min(times[1:]):8.2f} ms (sweeps/svd {st['jacobi_sweeps']/max(1,st['svds']):.1f})  "
          f"cusolver gesvdj {res['gesvdj']:8.2f} ms  gesvd {res['gesvd']:8.2f} ms  numpy/LAPACK {cpu:8.2f} ms")
    m.close(); base.close()
ctx.close()
