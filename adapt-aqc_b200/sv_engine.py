"""Statevector engine: device-resident states + the incremental cost evaluator.

``SVEngine`` is a thin object wrapper over the ``b200_sv_*`` C-ABI (include/b200aqc.h).
``SVCostEvaluator`` implements what ``AerSVBackend.evaluate_global_cost`` computes
(adaptaqc/backends/aer_sv_backend.py:23-47) without re-simulating the whole circuit on every
call:

  full_circuit = [ target U (fixed) | variational window W_0 .. W_{m-1} (+ fixed rhs) ]
  cost         = 1 - |<0| W_{m-1} ... W_0 U |0>|^2

  * U|0> is simulated once and kept in HBM (slot BASE).
  * For a window that differs from the previous one only in 1-qubit gate k (what Rotosolve /
    Rotoselect do, adaptaqc/utils/cost_minimiser.py:318-368) the evaluator keeps
        |R_k> = W_{k-1}..W_0 U|0>      (slot R)        |L_k> = (W_{m-1}..W_{k+1})^+ |0>   (slot L)
    and one launch of the inner-product kernel gives the 2x2 matrix
        M[i][j] = <L_k| (|i><j| on qubit q_k) |R_k>,   <0|psi> = sum_ij G[i][j] M[i][j]
    for ANY gate G at position k: all shift angles and all three axes of that gate come from a
    single launch.  Moving to another gate applies / un-applies only the gates in between.
  * Anything else (structure change, many gates changed) falls back to re-applying the window to
    the cached U|0> with the fused sweep kernels.
"""
import ctypes

import numpy as np

from . import gates as G
from .lib import B200Error, check, dptr, load

SLOT_WORK, SLOT_BASE, SLOT_L, SLOT_R = 0, 1, 2, 3


class SVEngine:
    """One GPU context holding `n_slots` statevectors of `num_qubits` qubits."""

    def __init__(self, num_qubits, device=0, n_slots=4):
        self._lib = load()
        self._ctx = ctypes.c_void_p()
        check(self._lib.b200_ctx_create(int(device), ctypes.byref(self._ctx)))
        self.device = int(device)
        self.num_qubits = int(num_qubits)
        self.n_slots = int(n_slots)
        try:
            check(self._lib.b200_sv_alloc(self._ctx, self.num_qubits, self.n_slots))
        except B200Error:
            self.close()
            raise

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.b200_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state manipulation ----
    def init_zero(self, slot):
        check(self._lib.b200_sv_init_zero(self._ctx, slot))

    def copy(self, dst, src):
        check(self._lib.b200_sv_copy(self._ctx, dst, src))

    def run(self, dst, src, stream, inverse=False):
        """dst <- stream applied to src (src = -1: |0..0>).  `stream` is a GateStream."""
        fn = self._lib.b200_sv_run_inverse if inverse else self._lib.b200_sv_run
        check(fn(self._ctx, dst, src, stream.rec_ptr(), len(stream.rec), stream.mats_ptr(), len(stream.mats)))

    def run_gates(self, dst, src, gates, inverse=False):
        self.run(dst, src, G.GateStream.from_gates(gates), inverse)

    # ---- read-outs ----
    def amp(self, slot, index=0):
        out = np.zeros(2)
        check(self._lib.b200_sv_amp(self._ctx, slot, int(index), dptr(out)))
        return complex(out[0], out[1])

    def expz(self, slot):
        """(<Z_q> for all q, norm^2)"""
        out = np.zeros(self.num_qubits + 1)
        check(self._lib.b200_sv_expz(self._ctx, slot, dptr(out)))
        return out[:-1], float(out[-1])

    def pair_rdm(self, slot, pairs):
        pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        out = np.zeros((len(pairs), 16), dtype=np.complex128)
        if len(pairs):
            check(self._lib.b200_sv_pair_rdm(self._ctx, slot, pairs.ctypes.data, len(pairs), dptr(out.view(np.float64))))
        return out.reshape(-1, 4, 4)

    def inner(self, l_slot, r_slot, q=-1):
        out = np.zeros(8)
        check(self._lib.b200_sv_inner(self._ctx, l_slot, r_slot, int(q), dptr(out)))
        m = out.view(np.complex128)
        return complex(m[0]) if q < 0 else m.reshape(2, 2).copy()

    def download(self, slot, offset=0, count=None):
        count = (1 << self.num_qubits) - offset if count is None else count
        host = np.empty(count, dtype=np.complex128)
        check(self._lib.b200_sv_download(self._ctx, slot, int(offset), int(count), host.ctypes.data))
        return host

    def upload(self, slot, host, offset=0):
        host = np.ascontiguousarray(host, dtype=np.complex128)
        check(self._lib.b200_sv_upload(self._ctx, slot, int(offset), host.size, host.ctypes.data))

    def device_ptr(self, slot):
        p = ctypes.c_void_p()
        check(self._lib.b200_sv_device_ptr(self._ctx, slot, ctypes.byref(p)))
        return p.value

    # ---- bookkeeping ----
    def sync(self):
        check(self._lib.b200_ctx_sync(self._ctx))

    def counters(self):
        out = (ctypes.c_uint64 * 8)()
        check(self._lib.b200_ctx_counters(self._ctx, out))
        return {"launches": out[0], "sweeps": out[1], "gates": out[2], "bytes": out[3],
                "h2d_bytes": out[4], "d2h_bytes": out[5], "calls": out[6]}

    def mark(self, which):
        """Record bench event 0 (start) / 1 (stop) on the context's stream."""
        check(self._lib.b200_ctx_mark(self._ctx, int(which)))

    def elapsed_ms(self):
        ms = ctypes.c_double()
        check(self._lib.b200_ctx_elapsed_ms(self._ctx, ctypes.byref(ms)))
        return ms.value

    PROF_CLASSES = ("sweep", "small", "expz", "rdm", "inner", "fill", "reduce", "mps")

    def profile(self, enable=True):
        check(self._lib.b200_ctx_profile(self._ctx, 1 if enable else 0))

    def profile_read(self):
        """{class: (total_ms, launches)} of the kernels launched since profile(True)."""
        ms = (ctypes.c_double * 8)()
        cnt = (ctypes.c_uint64 * 8)()
        check(self._lib.b200_ctx_profile_read(self._ctx, ms, cnt))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(self.PROF_CLASSES)}

    def set_timing(self, enable=True):
        check(self._lib.b200_ctx_set_timing(self._ctx, 1 if enable else 0))

    def last_ms(self):
        ms = ctypes.c_double()
        check(self._lib.b200_ctx_last_ms(self._ctx, ctypes.byref(ms)))
        return ms.value


def plan_stats(num_qubits, stream):
    """(sweeps, rounds, fused ops, small_path) the planner produces -- no GPU needed."""
    out = (ctypes.c_int32 * 4)()
    check(load().b200_sv_plan_stats(int(num_qubits), stream.rec_ptr(), len(stream.rec), stream.mats_ptr(),
                                    len(stream.mats), out))
    return tuple(out)


def _is_1q(ent):
    return ent[2] < 0


class SVCostEvaluator:
    """Global-cost evaluation of ``[prefix | window]`` circuits with HBM-resident caches."""

    REFRESH_MOVES = 256  # rebuild L/R from scratch after this many incremental moves

    def __init__(self, engine: SVEngine):
        self.eng = engine
        self.base_key = None          # identity of the cached prefix state
        self.window = None            # canonical window the L/R/M caches refer to
        self.pivot = None             # index k of the open gate (L/R exclude it)
        self.M = None                 # 2x2 transfer matrix at the pivot
        self.lr_valid = False
        self.moves = 0
        self.stats = {"resim": 0, "pivot_build": 0, "pivot_move": 0, "m_hits": 0, "evals": 0}

    # ---- prefix (target) state ----
    def set_base(self, key, prefix_stream):
        """Simulate the fixed prefix once into slot BASE (key identifies it)."""
        if key is not None and key == self.base_key:
            return
        self.eng.run(SLOT_BASE, -1, prefix_stream)
        self.base_key = key
        self.invalidate()

    def invalidate(self):
        self.window = None
        self.pivot = None
        self.M = None
        self.lr_valid = False

    # ---- evaluation ----
    def amp0(self, window):
        """<0| W |base> for the canonical window `window` (one scalar, as the reference asks)."""
        self.stats["evals"] += 1
        old = self.window
        if old is not None and len(old) == len(window):
            diff = [i for i in range(len(window)) if window[i] != old[i]]
            k = self._choose_pivot(diff, old, window)
            if k is not None:
                self.prepare_pivot(window, k)
                return complex(np.sum(G.matrix_of_entry(window[k]) * self.M))
        return self._resim(window)

    def _choose_pivot(self, diff, old, new):
        if not diff:
            return self.pivot if (self.pivot is not None and self.lr_valid) else None
        same_1q = all(_is_1q(old[i]) and _is_1q(new[i]) and old[i][1] == new[i][1] for i in diff)
        if not same_1q or len(diff) > 2:
            return None
        if len(diff) == 2 and self.pivot not in diff:
            return None
        others = [i for i in diff if i != self.pivot]
        return others[0] if others else self.pivot

    def _resim(self, window):
        eng = self.eng
        eng.run(SLOT_WORK, SLOT_BASE, G.GateStream.from_window(window))
        self.stats["resim"] += 1
        self.window = list(window)
        self.pivot = None
        self.M = None
        self.lr_valid = False
        return eng.amp(SLOT_WORK, 0)

    def prepare_pivot(self, window, k):
        """Make the transfer matrix M valid for position k of `window` (window[k] is irrelevant:
        L/R exclude it).  One launch of the inner kernel serves every gate put at k afterwards."""
        old = self.window
        ok = old is not None and len(old) == len(window) and self.lr_valid and self.pivot is not None
        if ok:
            kp = self.pivot
            for i in range(len(window)):
                if i != k and window[i] != old[i]:
                    # besides k, only the previous pivot may have changed (same qubit, 1-qubit gate)
                    if not (i == kp and _is_1q(old[i]) and _is_1q(window[i]) and old[i][1] == window[i][1]):
                        ok = False
                        break
        if ok and self.pivot == k:
            self.stats["m_hits"] += 1
        else:
            self._move_pivot(k, old if ok else None, window)
        self.window = list(window)

    def _move_pivot(self, k, old, new):
        """Make L/R/M valid for pivot k of window `new`; `old` = window the caches refer to."""
        eng = self.eng
        kp = self.pivot
        if old is None or self.moves >= self.REFRESH_MOVES:
            eng.run(SLOT_R, SLOT_BASE, G.GateStream.from_window(new[:k]))
            eng.run(SLOT_L, -1, G.GateStream.from_window(new[k + 1:]), inverse=True)
            self.moves = 0
            self.stats["pivot_build"] += 1
        elif k > kp:
            # R_k = new[k-1]..new[kp] R_kp ;  L_k = old[k]..old[kp+1] L_kp
            eng.run(SLOT_R, SLOT_R, G.GateStream.from_window(new[kp:k]))
            eng.run(SLOT_L, SLOT_L, G.GateStream.from_window(old[kp + 1:k + 1]))
            self.moves += 1
            self.stats["pivot_move"] += 1
        else:
            # R_k = (old[kp-1]..old[k])^+ R_kp ;  L_k = (new[kp]..new[k+1])^+ L_kp
            eng.run(SLOT_R, SLOT_R, G.GateStream.from_window(old[k:kp]), inverse=True)
            eng.run(SLOT_L, SLOT_L, G.GateStream.from_window(new[k + 1:kp + 1]), inverse=True)
            self.moves += 1
            self.stats["pivot_move"] += 1
        self.pivot = k
        self.lr_valid = True
        self.M = eng.inner(SLOT_L, SLOT_R, new[k][1])

    # ---- batched API (K6): every shift value of one gate from one launch ----
    def shift_amplitudes(self, window, k, candidates):
        """<0|psi> for each replacement 2x2 matrix in `candidates` at window position k."""
        self.prepare_pivot(window, k)
        self.stats["evals"] += len(candidates)
        return [complex(np.sum(np.asarray(c) * self.M)) for c in candidates]
