"""MPS engine: device-resident matrix-product states over the ``b200_mps_*`` C-ABI.

``DeviceMPS`` wraps one ``b200_mps`` handle (Vidal form in HBM).  ``MPSEngine`` gives the four
"slots" interface (WORK, BASE, L, R) that ``sv_engine.SVCostEvaluator`` drives, so the block
transfer-matrix evaluation of Rotosolve / Rotoselect is shared verbatim between the statevector
and the MPS path: |R> = prefix applied to the target MPS, <L| = suffix applied to <0|, and one
environment sweep (``b200_mps_transfer``) yields the 4x4 matrix T that serves every evaluation of
a layer.

Wire format on the host = the reference's ``QiskitMPS`` (adaptaqc/utils/constants.py:17):
``(gammas, lambdas)``, ``gammas[i] = (A0, A1)`` each (chi_{i-1}, chi_i), ``lambdas[i]`` real.
"""
import ctypes

import numpy as np

from . import gates as G
from .lib import check, dptr, load


def pack_qiskit_mps(mps):
    """QiskitMPS tuple -> (bond_dims int32[n-1], gammas float64 flat, lambdas float64 flat)."""
    gammas, lambdas = mps
    n = len(gammas)
    bond = np.array([len(np.atleast_1d(l)) for l in lambdas], dtype=np.int32)
    gs = []
    for i, (a0, a1) in enumerate(gammas):
        cl = 1 if i == 0 else int(bond[i - 1])
        cr = 1 if i == n - 1 else int(bond[i])
        g = np.stack([np.asarray(a0, dtype=np.complex128).reshape(cl, cr),
                      np.asarray(a1, dtype=np.complex128).reshape(cl, cr)])
        gs.append(g.reshape(-1))
    gflat = np.ascontiguousarray(np.concatenate(gs)).view(np.float64)
    lflat = (np.ascontiguousarray(np.concatenate([np.asarray(l, dtype=np.float64).reshape(-1) for l in lambdas]))
             if n > 1 else np.zeros(1))
    return bond, gflat, lflat


def pack_preprocessed(pp):
    """Preprocessed list [(2, chi_l, chi_r)] -> same triple with unit lambdas (Gamma.lambda is
    already folded into the tensors)."""
    n = len(pp)
    bond = np.array([pp[i].shape[2] for i in range(n - 1)], dtype=np.int32)
    gflat = np.ascontiguousarray(np.concatenate([np.asarray(g, dtype=np.complex128).reshape(-1) for g in pp])).view(np.float64)
    lflat = np.ones(max(1, int(bond.sum())))
    return bond, gflat, lflat


CHOP_AER, CHOP_SIGMA = 0, 1      # include/b200aqc.h: B200_CHOP_AER / B200_CHOP_SIGMA


def set_chop_rule(rule):
    """Process-wide reading of Aer's reduce_zeros (see include/b200aqc.h, b200_mps_set_chop_rule)."""
    check(load().b200_mps_set_chop_rule({"aer": CHOP_AER, "sigma": CHOP_SIGMA}.get(rule, rule)))


def reduce_zeros(singular_values, max_bond_dimension=None, truncation_threshold=1e-16):
    """(kept count, kept values) of the library's truncation rule on a descending vector (host only, no GPU)."""
    s = np.ascontiguousarray(singular_values, dtype=np.float64)
    out = np.zeros(len(s))
    k = ctypes.c_int(0)
    check(load().b200_mps_reduce_zeros(dptr(s), len(s), int(max_bond_dimension or 0), float(truncation_threshold),
                                       ctypes.byref(k), dptr(out)))
    return k.value, out[:k.value].copy()


class MPSContext:
    """One GPU context for MPS work (one CUDA stream)."""

    def __init__(self, device=0):
        self._lib = load()
        self._ctx = ctypes.c_void_p()
        check(self._lib.b200_ctx_create(int(device), ctypes.byref(self._ctx)))
        self.device = int(device)
        self._live = []

    def new_mps(self, num_qubits, truncation_threshold=1e-16, max_bond_dimension=None):
        return DeviceMPS(self, num_qubits, truncation_threshold, max_bond_dimension)

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            for m in list(self._live):
                m.close()
            self._lib.b200_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # bookkeeping shared with SVEngine
    def sync(self):
        check(self._lib.b200_ctx_sync(self._ctx))

    def counters(self):
        out = (ctypes.c_uint64 * 8)()
        check(self._lib.b200_ctx_counters(self._ctx, out))
        return {"launches": out[0], "sweeps": out[1], "gates": out[2], "bytes": out[3],
                "h2d_bytes": out[4], "d2h_bytes": out[5], "calls": out[6], "tensor_flops": out[7]}

    def mark(self, which):
        check(self._lib.b200_ctx_mark(self._ctx, int(which)))

    def elapsed_ms(self):
        ms = ctypes.c_double()
        check(self._lib.b200_ctx_elapsed_ms(self._ctx, ctypes.byref(ms)))
        return ms.value

    def profile(self, enable=True):
        check(self._lib.b200_ctx_profile(self._ctx, 1 if enable else 0))

    def profile_read(self):
        from .sv_engine import SVEngine
        names = SVEngine.PROF_CLASSES          # one table for both engines: B200_PROF_CLASSES entries (include/b200aqc.h)
        ms = (ctypes.c_double * len(names))()
        cnt = (ctypes.c_uint64 * len(names))()
        check(self._lib.b200_ctx_profile_read(self._ctx, ms, cnt))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(names)}


class DeviceMPS:
    def __init__(self, context, num_qubits, truncation_threshold=1e-16, max_bond_dimension=None):
        self.context = context
        self._lib = context._lib
        self.num_qubits = int(num_qubits)
        self._h = ctypes.c_void_p()
        check(self._lib.b200_mps_create(context._ctx, self.num_qubits, float(truncation_threshold),
                                        int(max_bond_dimension or 0), ctypes.byref(self._h)))
        context._live.append(self)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.b200_mps_destroy(self._h)
            self._h = ctypes.c_void_p()
            if self in self.context._live:
                self.context._live.remove(self)

    # ---- state ----
    def set_truncation(self, truncation_threshold, max_bond_dimension=None):
        check(self._lib.b200_mps_set_truncation(self._h, float(truncation_threshold), int(max_bond_dimension or 0)))

    def init_zero(self):
        check(self._lib.b200_mps_init_zero(self._h))

    def set(self, qiskit_mps):
        bond, g, l = pack_qiskit_mps(qiskit_mps)
        self._set_packed(bond, g, l)

    def set_preprocessed(self, pp):
        bond, g, l = pack_preprocessed(pp)
        self._set_packed(bond, g, l)

    def _set_packed(self, bond, g, l):
        if len(bond) != self.num_qubits - 1:
            raise ValueError("MPS has the wrong number of sites")
        check(self._lib.b200_mps_set(self._h, bond.ctypes.data if len(bond) else None, g.ctypes.data, l.ctypes.data))

    def bond_dims(self):
        out = np.zeros(max(1, self.num_qubits - 1), dtype=np.int32)
        check(self._lib.b200_mps_bond_dims(self._h, out.ctypes.data))
        return [int(x) for x in out[:self.num_qubits - 1]]

    def get(self):
        """Download as a QiskitMPS tuple."""
        bond = self.bond_dims()
        n = self.num_qubits
        dims = [(1 if i == 0 else bond[i - 1], 1 if i == n - 1 else bond[i]) for i in range(n)]
        g = np.zeros(sum(2 * a * b for a, b in dims), dtype=np.complex128)
        l = np.zeros(max(1, sum(bond)), dtype=np.float64)
        check(self._lib.b200_mps_get(self._h, g.ctypes.data, l.ctypes.data))
        gammas, lambdas, go, lo = [], [], 0, 0
        for i, (a, b) in enumerate(dims):
            t = g[go:go + 2 * a * b].reshape(2, a, b)
            gammas.append((t[0].copy(), t[1].copy()))
            go += 2 * a * b
            if i < n - 1:
                lambdas.append(l[lo:lo + b].copy())
                lo += b
        return (gammas, lambdas)

    def copy_from(self, other):
        check(self._lib.b200_mps_copy(self._h, other._h))

    def apply(self, stream, inverse=False):
        fn = self._lib.b200_mps_apply_inverse if inverse else self._lib.b200_mps_apply
        check(fn(self._h, stream.rec_ptr(), len(stream.rec), stream.mats_ptr(), len(stream.mats)))

    # ---- read-outs ----
    def amps(self, bitstrings):
        bits = np.ascontiguousarray(np.asarray(bitstrings, dtype=np.uint64))
        out = np.zeros(2 * len(bits))
        if len(bits):
            check(self._lib.b200_mps_amps(self._h, bits.ctypes.data, len(bits), dptr(out)))
        return out.view(np.complex128).copy()

    def dot(self, other):
        """<self|other>"""
        out = np.zeros(2)
        check(self._lib.b200_mps_dot(self._h, other._h, dptr(out)))
        return complex(out[0], out[1])

    def transfer(self, other, qubits):
        """T[i][j] = <self|(|i><j| on qubits)|other>; 2x2 or 4x4."""
        q = np.asarray(qubits, dtype=np.int32)
        d = 1 << len(q)
        out = np.zeros(2 * d * d)
        check(self._lib.b200_mps_transfer(self._h, other._h, q.ctypes.data, len(q), dptr(out)))
        return out.view(np.complex128).reshape(d, d).copy()

    def pair_transfer(self, other, pairs):
        """T_p[i][j] = <self|(|i><j| on pair p)|other> for every pair, index = bit(pair[0]) + 2 bit(pair[1]):
        all pairs from one left + one right environment sweep."""
        pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        out = np.zeros((len(pairs), 16), dtype=np.complex128)
        if len(pairs):
            check(self._lib.b200_mps_pair_transfer(self._h, other._h, pairs.ctypes.data, len(pairs), dptr(out.view(np.float64))))
        return out.reshape(-1, 4, 4)

    def expz(self):
        out = np.zeros(self.num_qubits + 1)
        check(self._lib.b200_mps_expz(self._h, dptr(out)))
        return out[:-1], float(out[-1])

    def pair_rdm(self, pairs):
        pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        out = np.zeros((len(pairs), 16), dtype=np.complex128)
        if len(pairs):
            check(self._lib.b200_mps_pair_rdm(self._h, pairs.ctypes.data, len(pairs), dptr(out.view(np.float64))))
        return out.reshape(-1, 4, 4)

    def stats(self):
        out = (ctypes.c_uint64 * 4)()
        check(self._lib.b200_mps_stats(self._h, out))
        return {"svds": out[0], "jacobi_sweeps": out[1], "max_bond": out[2], "svd_flops": out[3]}


SLOT_WORK, SLOT_BASE, SLOT_L, SLOT_R = 0, 1, 2, 3


class MPSEngine:
    """The slot interface ``SVCostEvaluator`` expects, backed by four device MPS."""

    def __init__(self, num_qubits, context=None, device=0, truncation_threshold=1e-16, max_bond_dimension=None,
                 n_slots=4):
        self.context = context if context is not None else MPSContext(device)
        self._owns = context is None
        self.num_qubits = int(num_qubits)
        self.slots = [self.context.new_mps(num_qubits, truncation_threshold, max_bond_dimension) for _ in range(n_slots)]

    def close(self):
        for m in self.slots:
            m.close()
        if self._owns:
            self.context.close()

    def set_truncation(self, thr, max_chi):
        for m in self.slots:
            m.set_truncation(thr, max_chi)

    def run(self, dst, src, stream, inverse=False):
        d = self.slots[dst]
        if src < 0:
            d.init_zero()
        elif src != dst:
            d.copy_from(self.slots[src])
        if len(stream):
            d.apply(stream, inverse=inverse)

    def amp(self, slot, index=0):
        return complex(self.slots[slot].amps([index])[0])

    def inner(self, l_slot, r_slot, q=-1):
        if q < 0:
            return self.slots[l_slot].dot(self.slots[r_slot])
        return self.slots[l_slot].transfer(self.slots[r_slot], [q])

    def inner2(self, l_slot, r_slot, qa, qb):
        return self.slots[l_slot].transfer(self.slots[r_slot], [qa, qb])

    def expz(self, slot):
        return self.slots[slot].expz()

    def pair_rdm(self, slot, pairs):
        return self.slots[slot].pair_rdm(pairs)

    def sync(self):
        self.context.sync()

    def counters(self):
        return self.context.counters()
