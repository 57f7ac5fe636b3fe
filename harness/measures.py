"""Pair-entanglement measures: the 4x4 post-processing of the reduced density matrices.

The RDMs themselves come from the device (``b200_sv_pair_rdm`` / ``b200_mps_pair_rdm``); what
is left is O(1) host algebra on 4x4 matrices, kept on the host with the same scipy calls as the
reference so that pair selection ties break the same way
(adaptaqc/utils/entanglement_measures.py:262-308, 343-370).
"""
import itertools

import numpy as np
from scipy import linalg

EM_TOMOGRAPHY_EOF = "EM_TOMOGRAPHY_EOF"
EM_TOMOGRAPHY_CONCURRENCE = "EM_TOMOGRAPHY_CONCURRENCE"
EM_TOMOGRAPHY_NEGATIVITY = "EM_TOMOGRAPHY_NEGATIVITY"
EM_TOMOGRAPHY_LOG_NEGATIVITY = "EM_TOMOGRAPHY_LOG_NEGATIVITY"

_YY = np.array([[0, 0, 0, -1], [0, 0, 1, 0], [0, 1, 0, 0], [-1, 0, 0, 0]], dtype=np.complex128)


def concurrence(rho):
    """Wootters concurrence (PhysRevLett.80.2245); entanglement_measures.py:278-296."""
    rho = np.asarray(rho, dtype=np.complex128)
    rho_tilde = _YY @ rho.conjugate() @ _YY
    eigenvalues = linalg.eig(rho @ rho_tilde, left=False, right=False)
    if not np.allclose(np.imag(eigenvalues), 0):
        return 0
    lambdas = sorted(np.sqrt(np.real(eigenvalues).clip(min=0)), reverse=True)
    return np.max([0, lambdas[0] - lambdas[1] - lambdas[2] - lambdas[3]])


def eof(rho):
    """Entanglement of formation = binary entropy of (1+sqrt(1-C^2))/2; :262-275."""
    c = concurrence(rho)
    if c == 0:
        return 0
    x = 0.5 * (1 + np.sqrt(1 - c**2))
    return (-x * np.log2(x)) - ((1 - x) * np.log2(1 - x))


def partial_transpose(rho, wrt=1):
    """:343-356"""
    rho = np.asarray(rho)
    tp = np.array(rho, copy=True)
    for ja, ka, jb, kb in itertools.product(range(2), repeat=4):
        if wrt == 1:
            tp[ka * 2 + jb][ja * 2 + kb] = rho[ja * 2 + jb][ka * 2 + kb]
        else:
            tp[ja * 2 + kb][ka * 2 + jb] = rho[ja * 2 + jb][ka * 2 + kb]
    return tp


def trace_norm(m):
    """:359-370"""
    return np.real(np.trace(linalg.sqrtm(np.matmul(m, np.conjugate(m).transpose()))))


def negativity(rho):
    return (trace_norm(partial_transpose(rho)) - 1) / 2


def log_negativity(rho):
    return np.log2(trace_norm(partial_transpose(rho)))


def measure_from_rho(method, rho):
    """Dispatch of entanglement_measures.py:89-98."""
    if method == EM_TOMOGRAPHY_EOF:
        return eof(rho)
    if method == EM_TOMOGRAPHY_CONCURRENCE:
        return concurrence(rho)
    if method == EM_TOMOGRAPHY_NEGATIVITY:
        return negativity(rho)
    if method == EM_TOMOGRAPHY_LOG_NEGATIVITY:
        return log_negativity(rho)
    raise ValueError("Invalid entanglement measure method")
