"""Generates the committed fixtures under tests/golden/ (run in the build container only; the GPU
box has no /root/reference).

random_mps_seed_<s>.npz : ALL 54 of the reference's random 50-site chi=2 MPS targets
    (/root/reference/paper/random_mps/target_seed_<s>.pkl, QiskitMPS format
    adaptaqc/utils/constants.py:17), re-packed as plain arrays: g<i> = (2, chi_l, chi_r) Gamma of
    site i, l<i> = lambda of bond i.  They are data produced BY the reference's authors with
    qiskit-aer, i.e. genuine outputs of the third-party simulator the oracle restates.
"""
import glob
import os
import pickle
import re

import numpy as np

SRC = "/root/reference/paper/random_mps"
HERE = os.path.dirname(os.path.abspath(__file__))

for seed in sorted(int(re.search(r"seed_(\d+)", f).group(1)) for f in glob.glob(os.path.join(SRC, "target_seed_*.pkl"))):
    gammas, lambdas = pickle.load(open(os.path.join(SRC, f"target_seed_{seed}.pkl"), "rb"))
    arrays = {}
    for i, (a0, a1) in enumerate(gammas):
        arrays[f"g{i}"] = np.stack([np.asarray(a0), np.asarray(a1)])
    for i, lam in enumerate(lambdas):
        arrays[f"l{i}"] = np.asarray(lam)
    np.savez_compressed(os.path.join(HERE, f"random_mps_seed_{seed}.npz"), **arrays)
    print(seed, len(gammas), max(a.shape[2] for a in arrays.values() if a.ndim == 3))
