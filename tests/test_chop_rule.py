"""Aer's MPS truncation rule (reduce_zeros / num_of_SV of qiskit-aer 0.16 svd.cpp) on hand-built singular-value
vectors, product (C-ABI, host function) against oracle, in BOTH readings of the two ambiguous details
(include/b200aqc.h: B200_CHOP_AER = default, B200_CHOP_SIGMA = round-1 behaviour).  The Aer source is not vendored
under /root/reference, so neither reading can be executed against Aer here; DESIGN.md section 5 says which is assumed."""
import numpy as np
import pytest

from adapt_aqc_b200 import mps_engine as me
from oracle import mps_oracle as mo


@pytest.fixture(params=["aer", "sigma"])
def rule(request):
    me.set_chop_rule(request.param)
    mo.set_chop_rule(request.param)
    yield request.param
    me.set_chop_rule("aer")
    mo.set_chop_rule("aer")


def _norm(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.sqrt(np.sum(v * v))


CASES = [
    # (S, max_chi, thr, kept count under "aer", kept count under "sigma")
    (_norm([1.0, 0.5, 0.25]), None, 1e-16, 3, 3),
    ([1.0, 1e-9, 1e-12], None, 1e-16, 1, 1),               # same count, different route: chopped (aer) / summed away (sigma)
    ([1.0, 3e-8, 1e-9], None, 1e-16, 2, 2),
    ([1.0, 2e-8, 9e-9, 9e-9], None, 1e-16, 2, 3),          # sigma^2 <= 1e-16 is chopped by Aer even if the SUM would not fit
    ([1.0, 1e-20], None, 1e-16, 1, 1),
    (_norm([1.0, 0.5, 0.25, 0.125]), 2, 1e-16, 2, 2),      # bond cap
    (_norm([1.0, 1e-3, 1e-5, 1e-6]), None, 1e-8, 2, 2),    # threshold drops the tail, loop breaks at 1e-3
    ([1.0, 1e-5, 1e-6], None, 1e-8, 3, 1),                 # loop runs out: Aer leaves the count unchanged
    ([1.0, 1e-5, 1e-6], 2, 1e-8, 2, 1),                    # ... at the capped count
    ([1.0], None, 1e-16, 1, 1),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_hand_built_vectors(rule, case):
    S, cap, thr, k_aer, k_sigma = CASES[case]
    S = np.asarray(S, dtype=np.float64)
    k_ref, kept_ref = mo.reduce_zeros(S, cap, thr)
    k, kept = me.reduce_zeros(S, cap, thr)
    assert k == k_ref == (k_aer if rule == "aer" else k_sigma)
    np.testing.assert_array_equal(kept, kept_ref)
    counted = int(np.count_nonzero((S * S if rule == "aer" else S) > 1e-16))
    if k < counted:
        assert abs(np.sum(kept ** 2) - 1) < 1e-15          # renormalised only when something was dropped
    else:
        np.testing.assert_array_equal(kept, S[:k])


def test_random_vectors_agree_with_the_oracle(rule):
    rng = np.random.default_rng(3)
    for _ in range(300):
        n = int(rng.integers(1, 40))
        S = np.sort(10.0 ** rng.uniform(-19, 0, n))[::-1].copy()
        S[0] = 1.0
        cap = None if rng.random() < 0.5 else int(rng.integers(1, n + 1))
        thr = float(10.0 ** rng.uniform(-17, -3))
        k_ref, kept_ref = mo.reduce_zeros(S, cap, thr)
        k, kept = me.reduce_zeros(S, cap, thr)
        assert k == k_ref
        np.testing.assert_allclose(kept, kept_ref, rtol=1e-15, atol=0)


def test_default_rule_is_aer():
    me.set_chop_rule("aer")
    assert me.reduce_zeros([1.0, 1e-9])[0] == 1
    with pytest.raises(Exception):
        me.set_chop_rule(7)
