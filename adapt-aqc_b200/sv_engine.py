"""Statevector engine: device-resident states + the incremental cost evaluator.

``SVEngine`` is a thin object wrapper over the ``b200_sv_*`` C-ABI (include/b200aqc.h).
``SVCostEvaluator`` implements what ``AerSVBackend.evaluate_global_cost`` computes
(adaptaqc/backends/aer_sv_backend.py:23-47) without re-simulating the whole circuit on every
call:

  full_circuit = [ target U (fixed) | variational window W_0 .. W_{m-1} (+ fixed rhs) ]
  cost         = 1 - |<0| W_{m-1} ... W_0 U |0>|^2

  * U|0> is simulated once and kept in HBM (slot BASE).
  * The window is partitioned into BLOCKS: maximal runs of consecutive gates supported on at most
    two qubits (qa, qb).  An ADAPT layer (rz rz cx rz rz on one pair) is exactly one block.
  * For the block [a0, a1) the evaluator keeps, in HBM,
        |R> = W_{a0-1}..W_0 U|0>   (slot R)        |L> = (W_{m-1}..W_{a1})^+ |0>   (slot L)
    and one read pass over both (``b200_sv_inner2``) gives the 4x4 transfer matrix
        T[i][j] = <L| (|i><j| on (qa,qb)) |R>,      <0|psi> = sum_ij O[i][j] T[i][j]
    for ANY operator O on (qa,qb) -- i.e. for any angles and any rx/ry/rz choice of every
    rotation in the block.  All Rotoselect / Rotosolve evaluations of a layer
    (adaptaqc/utils/cost_minimiser.py:267-368), over any number of optimiser cycles, are then
    4x4 host algebra; the device is touched only when the optimiser moves to another layer
    (two fused gate sweeps + one transfer pass) or the circuit structure changes.
"""
import ctypes
import os

import numpy as np

from . import gates as G
from .lib import B200Error, check, dptr, load

SLOT_WORK, SLOT_BASE, SLOT_L, SLOT_R = 0, 1, 2, 3


class SVEngine:
    """One GPU context holding `n_slots` statevectors of `num_qubits` qubits."""

    def __init__(self, num_qubits, device=0, n_slots=4, external_memory=False):
        """external_memory=True: slots get their memory from attach() (caller-owned device
        buffers, e.g. torch tensors that NCCL sends from / receives into)."""
        self._lib = load()
        self._ctx = ctypes.c_void_p()
        check(self._lib.b200_ctx_create(int(device), ctypes.byref(self._ctx)))
        self.device = int(device)
        self.num_qubits = int(num_qubits)
        self.n_slots = int(n_slots)
        try:
            if external_memory:
                check(self._lib.b200_sv_reserve_slots(self._ctx, self.num_qubits, self.n_slots))
            else:
                check(self._lib.b200_sv_alloc(self._ctx, self.num_qubits, self.n_slots))
        except B200Error:
            self.close()
            raise

    def attach(self, slot, device_ptr):
        check(self._lib.b200_sv_attach(self._ctx, int(slot), ctypes.c_void_p(int(device_ptr))))

    # ---- NVLink peer memory (sharded statevector, dist_sv) ----
    def ipc_export(self, slot):
        """64-byte CUDA IPC handle of a library-owned slot (to be opened by the other ranks)."""
        buf = ctypes.create_string_buffer(64)
        check(self._lib.b200_sv_ipc_export(self._ctx, int(slot), buf))
        return buf.raw

    def ipc_open(self, handle):
        p = ctypes.c_void_p()
        check(self._lib.b200_sv_ipc_open(self._ctx, bytes(handle), ctypes.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        check(self._lib.b200_sv_ipc_close(self._ctx, ctypes.c_void_p(ptr)))

    def peer_swap(self, slot, peer_ptrs, rank, positions=None):
        """In-place exchange of chunk[p] with rank p's chunk[rank] for all p (one kernel over NVLink).  positions: local
        bit positions that trade places with the rank bits (default: the top log2(world) local bits = contiguous chunks)."""
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[ctypes.c_void_p(p or 0) for p in peer_ptrs])
        if positions is None:
            check(self._lib.b200_sv_peer_swap(self._ctx, int(slot), arr, len(peer_ptrs), int(rank)))
        else:
            pos = np.ascontiguousarray(np.asarray(positions, dtype=np.int32))
            check(self._lib.b200_sv_peer_swap_strided(self._ctx, int(slot), arr, len(peer_ptrs), int(rank), pos.ctypes.data))

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

    def pair_rdm(self, slot, pairs, part=0, n_parts=1):
        """4x4 RDMs of `pairs`.  n_parts > 1: only this share of the read passes is launched (pairs owned by other
        shares come back as zeros; the sum over the shares is bit-identical to the undivided call)."""
        pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        out = np.zeros((len(pairs), 16), dtype=np.complex128)
        if len(pairs):
            check(self._lib.b200_sv_pair_rdm_part(self._ctx, slot, pairs.ctypes.data, len(pairs), int(part), int(n_parts),
                                                  dptr(out.view(np.float64))))
        return out.reshape(-1, 4, 4)

    def inner(self, l_slot, r_slot, q=-1):
        out = np.zeros(8)
        check(self._lib.b200_sv_inner(self._ctx, l_slot, r_slot, int(q), dptr(out)))
        m = out.view(np.complex128)
        return complex(m[0]) if q < 0 else m.reshape(2, 2).copy()

    def inner2(self, l_slot, r_slot, qa, qb):
        """4x4 transfer matrix T[i][j] = <L|(|i><j| on (qa,qb))|R>, index = bit(qa) + 2 bit(qb)."""
        out = np.zeros(32)
        check(self._lib.b200_sv_inner2(self._ctx, l_slot, r_slot, int(qa), int(qb), dptr(out)))
        return out.view(np.complex128).reshape(4, 4).copy()

    FUSED_MIN_QUBITS = 12

    def run_inner2(self, dst, src, stream, other, qa, qb, inverse=False, store=True):
        """run(dst, src, stream) and inner2(dst, other, qa, qb) in one pass over the register (b200_sv_run_inner2).
        store=False: only T is wanted -- returns (T, stored): dst is left alone when one sweep carries the program."""
        out = np.zeros(32)
        stored = ctypes.c_int(1 if store else 0)
        check(self._lib.b200_sv_run_inner2(self._ctx, dst, src, stream.rec_ptr(), len(stream.rec), stream.mats_ptr(),
                                           len(stream.mats), 1 if inverse else 0, int(other), int(qa), int(qb), dptr(out),
                                           ctypes.byref(stored)))
        T = out.view(np.complex128).reshape(4, 4).copy()
        return T if store else (T, bool(stored.value))

    def run_project(self, scratch, src, stream, qmap, dst_engine, dst_slot, inverse=False):
        """run + gather without storing the swept state (b200_sv_run_project): slot `dst_slot` of `dst_engine` <- the
        projection of (stream applied to slot src) onto |0> of every qubit outside qmap.  Synchronous like ``gather``.
        Returns True if slot `scratch` was overwritten (programs of several sweeps)."""
        qm = np.ascontiguousarray(np.asarray(qmap, dtype=np.int32))
        if len(qm) != dst_engine.num_qubits:
            raise ValueError("qmap must name one qubit per qubit of the destination engine")
        dst_engine.sync()
        used = ctypes.c_int(0)
        check(self._lib.b200_sv_run_project(self._ctx, int(scratch), int(src), stream.rec_ptr(), len(stream.rec), stream.mats_ptr(),
                                            len(stream.mats), 1 if inverse else 0, ctypes.c_void_p(dst_engine.device_ptr(dst_slot)),
                                            len(qm), qm.ctypes.data, ctypes.byref(used)))
        self.sync()
        return bool(used.value)

    def run_embedded(self, dst, qmap, src_engine, src_slot, stream, inverse=False, fuse=None, store=True):
        """dst <- stream applied to the embedded state (slot `src_slot` of `src_engine` on the qubits qmap, |0> elsewhere):
        scatter + run without the zero fill and the read pass.  fuse = (other slot, qa, qb): also returns the transfer
        matrix of (dst, other) from the same pass (b200_sv_run_embedded_inner2); store=False (with fuse): T is all that is
        wanted -- returns (T, stored), dst is left alone when one sweep carries the program."""
        qm = np.ascontiguousarray(np.asarray(qmap, dtype=np.int32))
        src_engine.sync()
        ptr = ctypes.c_void_p(src_engine.device_ptr(src_slot))
        if fuse is None:
            check(self._lib.b200_sv_run_embedded(self._ctx, int(dst), ptr, len(qm), qm.ctypes.data, stream.rec_ptr(), len(stream.rec),
                                                 stream.mats_ptr(), len(stream.mats), 1 if inverse else 0))
            return None
        out = np.zeros(32)
        stored = ctypes.c_int(1 if store else 0)
        check(self._lib.b200_sv_run_embedded_inner2(self._ctx, int(dst), ptr, len(qm), qm.ctypes.data, stream.rec_ptr(),
                                                    len(stream.rec), stream.mats_ptr(), len(stream.mats), 1 if inverse else 0,
                                                    int(fuse[0]), int(fuse[1]), int(fuse[2]), dptr(out), ctypes.byref(stored)))
        T = out.view(np.complex128).reshape(4, 4).copy()
        return T if store else (T, bool(stored.value))

    def inner2_gather(self, r_slot, compact_engine, compact_slot, qmap, qa, qb):
        """Same T with the bra given compactly: slot `compact_slot` of `compact_engine` (K qubits,
        same device) holds L on the qubits qmap[0..K)."""
        qm = np.ascontiguousarray(np.asarray(qmap, dtype=np.int32))
        out = np.zeros(32)
        check(self._lib.b200_sv_inner2_gather(self._ctx, r_slot, ctypes.c_void_p(compact_engine.device_ptr(compact_slot)),
                                              len(qm), qm.ctypes.data, int(qa), int(qb), dptr(out)))
        return out.view(np.complex128).reshape(4, 4).copy()

    def gather(self, slot, qmap, dst_engine, dst_slot):
        """dst[c] = slot[deposit(c, qmap)]: the state projected onto |0> of every qubit outside qmap, into
        slot `dst_slot` of the len(qmap)-qubit engine `dst_engine` (same device).  Synchronous."""
        qm = np.ascontiguousarray(np.asarray(qmap, dtype=np.int32))
        if len(qm) != dst_engine.num_qubits:
            raise ValueError("qmap must name one qubit per qubit of the destination engine")
        dst_engine.sync()
        check(self._lib.b200_sv_gather(self._ctx, int(slot), qm.ctypes.data, len(qm),
                                       ctypes.c_void_p(dst_engine.device_ptr(dst_slot))))
        self.sync()

    def scatter(self, slot, qmap, src_engine, src_slot):
        """slot[x] = src[extract(x, qmap)] where all qubits outside qmap are 0, and 0 elsewhere (inverse of gather)."""
        qm = np.ascontiguousarray(np.asarray(qmap, dtype=np.int32))
        src_engine.sync()
        check(self._lib.b200_sv_scatter(self._ctx, int(slot), qm.ctypes.data, len(qm),
                                        ctypes.c_void_p(src_engine.device_ptr(src_slot))))

    def gather_ranked(self, slot, qmap, rank_bits, rank, dst_engine, dst_slot):
        """One rank's part of ``gather`` on a register sharded over 2^rank_bits ranks (dist_sv): qmap entries
        >= num_qubits name rank bits.  Asynchronous on this engine's stream (the caller synchronises)."""
        qm = np.ascontiguousarray(np.asarray(qmap, dtype=np.int32))
        dst_engine.sync()
        check(self._lib.b200_sv_gather_ranked(self._ctx, int(slot), qm.ctypes.data, len(qm), int(rank_bits), int(rank),
                                              ctypes.c_void_p(dst_engine.device_ptr(dst_slot))))

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

    PROF_CLASSES = ("sweep", "small", "expz", "rdm", "inner", "fill", "reduce", "mps", "svd", "gemm", "fused", "fused_embed", "fused_read", "project", "fused_embed_read")

    def profile(self, enable=True):
        check(self._lib.b200_ctx_profile(self._ctx, 1 if enable else 0))

    def profile_read(self):
        """{class: (total_ms, launches)} of the kernels launched since profile(True)."""
        ms = (ctypes.c_double * len(self.PROF_CLASSES))()
        cnt = (ctypes.c_uint64 * len(self.PROF_CLASSES))()
        check(self._lib.b200_ctx_profile_read(self._ctx, ms, cnt))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(self.PROF_CLASSES)}

    def profile_sweeps(self, max_entries=4096):
        """Per-launch device times (ms) of the sweep kernel since profile(True), in launch order."""
        ms = (ctypes.c_double * max_entries)()
        n = ctypes.c_int()
        check(self._lib.b200_ctx_profile_sweeps(self._ctx, ms, max_entries, ctypes.byref(n)))
        return [ms[i] for i in range(min(n.value, max_entries))]

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


_I2 = np.eye(2, dtype=np.complex128)
_SWAP_BITS = [0, 2, 1, 3]
_M4 = {
    "cx": np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.complex128),
    "cz": np.diag([1, 1, 1, -1]).astype(np.complex128),
    "swap": np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128),
}


def plan_detail(num_qubits, stream, max_sweeps=256):
    """[(rounds, ops, dense ops, contiguous low qubits)] per sweep -- no GPU needed."""
    out = np.zeros(4 * max_sweeps, dtype=np.int32)
    ns = ctypes.c_int32()
    check(load().b200_sv_plan_detail(int(num_qubits), stream.rec_ptr(), len(stream.rec), stream.mats_ptr(),
                                     len(stream.mats), out.ctypes.data, max_sweeps, ctypes.byref(ns)))
    return [tuple(int(x) for x in out[4 * k:4 * k + 4]) for k in range(min(ns.value, max_sweeps))]


def _support(ent):
    return (ent[1],) if ent[2] < 0 else (ent[1], ent[2])


def partition_blocks(window):
    """[(start, stop, support)]: greedy maximal runs of consecutive gates on <= 2 qubits.  Depends
    only on the qubits of the gates, so Rotosolve / Rotoselect edits never move a boundary."""
    blocks = []
    start, supp = 0, ()
    for i, ent in enumerate(window):
        s = tuple(sorted(set(supp) | set(_support(ent))))
        if len(s) > 2:
            blocks.append((start, i, supp))
            start, s = i, tuple(sorted(_support(ent)))
        supp = s
    if len(window) > start:
        blocks.append((start, len(window), supp))
    return blocks


def _embed_1q(m, pos, npair):
    """2x2 matrix `m` acting on bit `pos` of a 1- or 2-qubit index, without np.kron."""
    if npair == 1:
        return m
    out = np.zeros((4, 4), dtype=np.complex128)
    if pos == 0:
        out[0:2, 0:2] = m
        out[2:4, 2:4] = m
    else:
        out[0::2, 0::2] = m
        out[1::2, 1::2] = m
    return out


_EMBED_CACHE = {}


def embed_entry(ent, pair, override=None):
    """Matrix of a canonical window entry on the qubits `pair` (index = bit(pair[0]) + 2 bit(pair[1]);
    2x2 when len(pair) == 1).  `override`: replacement 2x2 matrix for a 1-qubit entry."""
    if ent[2] < 0:
        pos = 0 if ent[1] == pair[0] else 1
        if override is not None:
            return _embed_1q(np.asarray(override, dtype=np.complex128), pos, len(pair))
        key = (ent, pos, len(pair))
        m = _EMBED_CACHE.get(key)
        if m is None:
            if len(_EMBED_CACHE) > 4096:
                _EMBED_CACHE.clear()
            m = _EMBED_CACHE[key] = _embed_1q(G.matrix_of_entry(ent), pos, len(pair))
        return m
    m4 = _M4.get(ent[0])
    if m4 is None:
        m4 = np.frombuffer(ent[6], dtype=np.complex128).reshape(4, 4)
    if (ent[1], ent[2]) != tuple(pair):
        m4 = m4[np.ix_(_SWAP_BITS, _SWAP_BITS)]
    return m4


class SVCostEvaluator:
    """Global-cost evaluation of ``[prefix | window]`` circuits with HBM-resident caches.

    R (slot R) and L are tracked independently:
      rwin : the gates R currently contains after the base (R = rwin . base)
      lwin : the gates the dense L (slot L) contains (L = lwin^+ |0>), or None
      lkey : identity of the compact L (a 2^K state in the compact engine), or None
    `compact` (optional): a second, K-qubit engine on the same device.  <L| = suffix^+ <0| is
    supported only on the qubits the suffix touches; when there are at most K of them L is built
    from scratch in the compact engine (microseconds) and T comes from a GATHER of 2^K amplitudes
    of R (``inner2_gather``) instead of a pass over 2^n.
    """

    REFRESH_MOVES = 256  # rebuild R / dense L from scratch after this many incremental moves
    PROJECT_MIN_SAVING = 3  # project only onto engines at least this many qubits smaller than the register ...
    LARGE_QUBITS = 20       # ... or at least ONE qubit smaller once the register has this many qubits (see min_saving)
    NEST_MIN_QUBITS = 26    # the evaluator of a projected engine of at least this size projects its own tail further down

    def __init__(self, engine, compact=None, projected=None, registry=None):
        """projected: optional K-qubit engines (4 slots each, same device) for the PROJECTED TAIL: once the
        gates of window[m:] act on at most K qubits S, <0|W|base> = <0_S| W[m:] |phi> with
        |phi> = <0_rest| W[:m] |base> (a gather of 2^K amplitudes of R, ``gather``), so every evaluation
        of the blocks in window[m:] -- R/L moves, transfer passes -- runs on the 2^K-amplitude engine
        through a nested evaluator, and no sweep over the 2^n register is needed while the optimiser
        stays in the tail.  The nested evaluator projects its own tail onto the next smaller engine, and so on: every
        thin layer adds at most two qubits to the support of the gates behind it, so each level keeps one or two
        blocks and the passes over 2^n, 2^(n-1), 2^(n-2), 2^(n-4) ... amplitudes add up to a geometric series.
        registry: {id(engine): evaluator} shared by all levels -- ONE evaluator per engine, whoever projects into it."""
        self.eng = engine
        # False (sharded registers, dist_sv.ShardedEngine): no dense bra / transfer passes over the register --
        # blocks outside the projected tail are evaluated by re-simulating the window from the base state
        self.dense_blocks = True
        self.prefetch_L = True        # see _prefetch_next_L
        self.projected = sorted(projected or [], key=lambda e: e.num_qubits)
        self._registry = registry if registry is not None else {}
        self._registry.setdefault(id(engine), self)
        self._proj_state = None       # (engine id, m, qmap, base key, token) of the phi currently held by that engine
        self._split_key, self._split = None, None
        self._split_info = None       # (split, (engine, qmap, position of each qubit, nested evaluator) | None)
        self._tail_cache = None       # (proj key, remapped tail window)
        # one or several compact engines of increasing size; the smallest that fits is used
        if compact is None:
            self.compacts = []
        elif isinstance(compact, (list, tuple)):
            self.compacts = sorted(compact, key=lambda e: e.num_qubits)
        else:
            self.compacts = [compact]
        self.compact = self.compacts[-1] if self.compacts else None
        self.base_key = None          # identity of the cached prefix state
        self.window = None            # last window evaluated
        self.cut = None               # (b0, b1) of the block T is open on
        self.pair = None              # qubits the transfer matrix T is open on
        self.T = None
        self.rwin = None
        self.r_slot = SLOT_R          # slot that holds R (SLOT_BASE while the prefix is empty)
        self.lwin = None
        self.lkey = None
        self.r_moves = self.l_moves = 0
        self._part_key, self._part = None, None
        self.stats = {"rebuild_R": 0, "rebuild_L": 0, "moves_R": 0, "moves_L": 0, "compact_L": 0, "t_passes": 0,
                      "t_gathers": 0, "host_evals": 0, "evals": 0, "projections": 0, "projected_evals": 0}

    # ---- prefix (target) state ----
    def set_base(self, key, prefix_stream):
        """Simulate the fixed prefix once into slot BASE (key identifies it)."""
        if key is not None and key == self.base_key:
            return
        self.eng.run(SLOT_BASE, -1, prefix_stream)
        self.base_key = key
        self.invalidate()

    def min_saving(self):
        """Project only onto engines at least this many qubits smaller than the register: halving a pass over >= 2^20
        amplitudes pays for the gather; small registers are launch-latency bound anyway."""
        return 1 if self.eng.num_qubits >= self.LARGE_QUBITS else self.PROJECT_MIN_SAVING

    def invalidate(self):
        self._proj_state = None       # (a nested evaluator is re-based, with a fresh token, by the next projection)
        self._hot, self._hot_dirty = None, False
        self.window = None
        self.cut = None
        self.pair = None
        self.T = None
        self._gw = None
        self.rwin = None
        self.lwin = None
        self.lkey = None

    # ---- block bookkeeping ----
    def _blocks(self, window, changed=None):
        """Block partition of `window`, cached on the qubits of its gates.  `changed` (indices that differ
        from the window of the previous call): only those entries are checked against the cached structure."""
        pk = self._part_key
        if (changed is not None and pk is not None and len(pk) == len(window)
                and all(pk[i] == (window[i][1], window[i][2]) for i in changed)):
            return self._part
        key = tuple((e[1], e[2]) for e in window)
        if key != pk:
            self._part_key, self._part = key, partition_blocks(window)
        return self._part

    _prev_k = None      # index of the gate whose value the previous amp0 call served (hot path), else None

    def _current_gate(self, changed):
        """The gate the optimiser is working on, out of the pending edits.  One scalar per call, the reference's optimiser
        leaves TWO edits pending when it moves on (cost_minimiser.py:344-368): the final angle of the gate it has just
        finished -- whose shifted values the previous calls served -- and the first shift of the next gate.  The next gate
        is usually the later one, but not when a cycle wraps around (last gate of the window -> first gate): opening the
        block of the finished gate there costs a projection and a bra rebuild for a single value."""
        if not changed:
            return None
        if len(changed) == 2 and self._prev_k is not None and self._prev_k in changed:
            return changed[0] if changed[1] == self._prev_k else changed[1]
        return max(changed)

    def _select_block(self, window, focus, changed=None):
        blocks = self._blocks(window)
        old, target = self.window, None
        if old is not None and self.cut is not None and len(old) == len(window):
            a0, a1 = self.cut
            diff = [i for i in range(len(window)) if window[i] != old[i]]
            outside = [i for i in diff if not a0 <= i < a1]
            if not outside:
                for b in blocks:
                    if b[0] == a0 and b[1] == a1:
                        return b
            target = self._current_gate(outside if outside else diff)
        if target is None and changed:
            target = self._current_gate(changed)       # no previous window to diff against (e.g. after a projected phase)
        if target is None:
            target = len(window) - 1 if focus is None else min(max(focus, 0), len(window) - 1)
        for b in blocks:
            if b[0] <= target < b[1]:
                return b
        raise AssertionError("block partition does not cover the window")

    # ---- projected tail ----
    def _tail_split(self, window, changed=None):
        """(m, qubits): window[m:] is the longest block-aligned tail whose support fits the largest
        projected engine; None if there is no such engine or no saving."""
        if not self.projected:
            return None
        blocks = self._blocks(window, changed)
        if self._split_key is not self._part_key:
            kmax = self.projected[-1].num_qubits
            supp, m = set(), len(window)
            for (s0, _, sp) in reversed(blocks):
                new = supp | set(sp)
                if len(new) > kmax:
                    break
                supp, m = new, s0
            ok = m < len(window) and len(supp) + self.min_saving() <= self.eng.num_qubits
            self._split_key, self._split = self._part_key, ((m, sorted(supp)) if ok else None)
        return self._split

    def _proj_info(self, split):
        """(split, (engine, qmap, position of each qubit, nested evaluator) | None) for a tail split, cached."""
        m, supp = split
        info = self._split_info
        if info is None or info[0] is not split:
            fits = [e for e in self.projected if e.num_qubits >= len(supp) and
                    e.num_qubits + self.min_saving() <= self.eng.num_qubits]
            if not fits:
                info = self._split_info = (split, None)
                return info
            peng = fits[0]
            used = set(supp)
            free = [q for q in range(self.eng.num_qubits) if q not in used]
            qmap = supp + free[:peng.num_qubits - len(supp)]     # padded qubits carry no gate: <0| projects them out
            sub = self._sub_for(peng)
            info = self._split_info = (split, (peng, tuple(qmap), {q: c for c, q in enumerate(qmap)}, sub))
        return info

    def _sub_for(self, peng):
        """The evaluator of a projected engine: one per engine, shared by every level that may project into it; its own
        projected tail goes to the smaller engines (no compact-bra engines: they belong to the top level)."""
        sub = self._registry.get(id(peng))
        if sub is None:
            nested = [e for e in self.projected if e.num_qubits < peng.num_qubits] if peng.num_qubits >= self.NEST_MIN_QUBITS else []
            sub = SVCostEvaluator(peng, None, nested, registry=self._registry)
        return sub

    def _bra_into(self, slot, window, fuse=None, store=True, split=False):
        """slot <- window^+ |0..0> on this engine.  The longest tail of `window` that fits a smaller engine is built THERE
        (recursively, and kept: the optimiser asks for several bras with the same tail in a row); the head gates run at
        this size in a sweep that reads its tiles straight from the small engine's slot (b200_sv_run_embedded: no zero
        fill, no scatter pass, no read pass).  fuse = (other slot, qa, qb): the transfer matrix of (slot, other) comes out
        of that same sweep and is returned (else None).  store=False (with fuse): T is all that is wanted -- returns
        (T | None, stored): when one sweep carries the head gates the bra is never written to the register at all."""
        eng, stream = self.eng, G.GateStream.from_window
        T = None
        if split is False:                              # (the caller may already have it)
            split = self._embed_split(window)
        if split is not None:
            m, peng, qmap, tail = split
            sub = self._sub_for(peng)
            if sub._work_tail != tail:                  # slot WORK of that engine still holds this tail's bra otherwise
                sub._work_tail = None
                sub._bra_into(SLOT_WORK, tail)
                sub._work_tail = tail
                self.stats["scattered_L"] = self.stats.get("scattered_L", 0) + 1
            if self.fused_passes and hasattr(eng, "run_embedded") and eng.num_qubits >= getattr(eng, "FUSED_MIN_QUBITS", 1 << 30) \
                    and (m > 0 or fuse is not None):
                self.stats["embedded_L"] = self.stats.get("embedded_L", 0) + 1
                if fuse is not None:
                    self.stats["fused_T"] = self.stats.get("fused_T", 0) + 1
                if fuse is not None and not store:
                    return eng.run_embedded(slot, list(qmap), peng, SLOT_WORK, stream(window[:m]), inverse=True, fuse=fuse,
                                            store=False)
                T = eng.run_embedded(slot, list(qmap), peng, SLOT_WORK, stream(window[:m]), inverse=True, fuse=fuse)
                return T if store else (T, True)
            eng.scatter(slot, list(qmap), peng, SLOT_WORK)
            if m > 0:
                eng.run(slot, slot, stream(window[:m]), inverse=True)
            return None if store else (None, True)
        if fuse is not None and self.fused_passes and eng.num_qubits >= getattr(eng, "FUSED_MIN_QUBITS", 1 << 30):
            self.stats["fused_T"] = self.stats.get("fused_T", 0) + 1
            T = eng.run_inner2(slot, -1, stream(window), fuse[0], fuse[1], fuse[2], inverse=True)
            return T if store else (T, True)
        eng.run(slot, -1, stream(window), inverse=True)
        return None if store else (None, True)

    _work_tail = None       # the gate list whose bra slot WORK of this engine holds for the level above (see _bra_into)

    _es_struct = None       # cache of _embed_split: (structure key, m, engine, qmap, pos)
    _es_tail = None         # ... and (the tail's entries, their translation): unchanged entries are the same objects

    def _embed_split(self, window):
        """(m, engine, qmap, tail in that engine's numbering): window[m:] is the longest block-aligned tail whose support
        fits a smaller engine (its bra is built there and embedded); None if there is none.  The split depends only on
        which qubits the gates act on and the translated tail only on the tail's entries: both are remembered, the
        optimiser asks for the same ones several times in a row."""
        eng = self.eng
        if not (self.projected and hasattr(eng, "scatter") and len(window)):
            return None
        key = [(e[1], e[2]) for e in window]
        st = self._es_struct
        if st is None or st[0] != key:
            kmax = max((e.num_qubits for e in self.projected if e.num_qubits + self.min_saving() <= eng.num_qubits), default=0)
            supp, m = set(), len(window)
            for (s0, _, sp) in reversed(partition_blocks(window)):
                new = supp | set(sp)
                if len(new) > kmax:
                    break
                supp, m = new, s0
            if m >= len(window):
                st = self._es_struct = (key, None, None, None, None)
            else:
                peng = [e for e in self.projected if e.num_qubits >= len(supp)][0]
                supp = sorted(supp)
                used = set(supp)
                qmap = supp + [q for q in range(eng.num_qubits) if q not in used][:peng.num_qubits - len(supp)]
                st = self._es_struct = (key, m, peng, qmap, {q: c for c, q in enumerate(qmap)})
            self._es_tail = None
        _, m, peng, qmap, pos = st
        if m is None:
            return None
        src = window[m:]
        tc = self._es_tail
        if tc is not None and tc[0] == src:
            tail = tc[1]
        else:
            tail = [(e[0], pos[e[1]], pos[e[2]] if e[2] >= 0 else -1) + tuple(e[3:]) for e in src]
            self._es_tail = (src, tail)
        return m, peng, qmap, tail

    def _projected(self, window, target, changed):
        """If window[target] lies in the projected tail: (nested evaluator, tail window in the engine's qubit
        numbering, changed indices relative to the tail | None, m); else None.  Makes phi valid."""
        split = self._tail_split(window, changed)
        if split is None or target < split[0]:
            return None
        m, supp = split
        info = self._proj_info(split)
        if info[1] is None:
            return None
        peng, qmap, pos, sub = info[1]
        ps = self._proj_state
        if ps is not None and sub.base_key is not ps[4]:
            ps = self._proj_state = None      # another level has projected into that engine since
        # phi is valid while it was projected from the same prefix of the same base into the same engine / qubit map
        pwin = self._pwin            # the prefix phi was projected from (slot R itself need not hold it: see below)
        state = (id(peng), m, qmap, self.base_key)
        # (compared by CONTENT: head blocks served in front mode never touch slot R, and a head edit that has already been
        # evaluated is no longer in `changed` -- nothing else records it.  Unchanged entries are the same tuple objects, so
        # the comparison is m identity checks.)
        prefix_same = m == 0 or (pwin is not None and len(pwin) == m and pwin == window[:m])
        if not (ps is not None and state == ps[:4] and prefix_same):
            # Nothing but the projection needs prefix|base> here (front mode serves the head blocks from the base state):
            # unless slot R already holds it, the sweep that would build it keeps only the projected amplitudes
            # (b200_sv_run_project: half the traffic of sweep + gather; slot R stays as it is)
            direct = (m > 0 and self.project_direct and self.front_mode and self.fused_passes and hasattr(self.eng, "run_project")
                      and self.eng.num_qubits >= getattr(self.eng, "FUSED_MIN_QUBITS", 1 << 30)
                      and not (self.rwin is not None and self.rwin == window[:m]))
            if direct:
                if self.eng.run_project(SLOT_R, SLOT_BASE, G.GateStream.from_window(window[:m]), list(qmap), peng, SLOT_BASE):
                    self.rwin, self.r_slot, self.r_moves = None, SLOT_R, 0      # slot R served as scratch
                self.stats["direct_projections"] = self.stats.get("direct_projections", 0) + 1
            else:
                if m > 0:
                    self._update_R(window, m)
                self.eng.gather(self.r_slot if m > 0 else SLOT_BASE, list(qmap), peng, SLOT_BASE)
            self._pwin = list(window[:m])
            token = ("projected", id(self), self.stats["projections"])
            sub.base_key = token
            sub.invalidate()
            self._proj_state = state + (token,)
            self.stats["projections"] += 1
            self._tail_cache = None
        # the dense evaluator's T / window no longer describe the circuit once the tail is edited here
        self.T = None
        self.window = None
        tc = self._tail_cache
        sub_changed = None
        state = self._proj_state
        if (tc is not None and tc[0] == state and len(tc[1]) == len(window) - m and changed is not None
                and all(i >= m for i in changed)):
            tail = tc[1]
            for i in changed:
                e = window[i]
                tail[i - m] = (e[0], pos[e[1]], pos[e[2]] if e[2] >= 0 else -1) + tuple(e[3:])
            sub_changed = [i - m for i in changed]
        else:
            tail = [(e[0], pos[e[1]], pos[e[2]] if e[2] >= 0 else -1) + tuple(e[3:]) for e in window[m:]]
            self._tail_cache = (state, tail)
        return sub, tail, sub_changed, m

    # ---- R: prefix applied to the base ----
    def _update_R(self, new, b0):
        """Returns True if slot R changed."""
        eng, stream = self.eng, G.GateStream.from_window
        old = self.rwin
        if old is not None and old == new[:b0]:
            return False
        if b0 == 0:
            # empty prefix: R IS the base state -- read slot BASE instead of copying it (self.r_slot)
            self.rwin, self.r_slot, self.r_moves = [], SLOT_BASE, 0
            return True
        if old is not None and self.r_moves < self.REFRESH_MOVES and self.r_slot == SLOT_R:
            a0 = len(old)
            m = min(a0, b0)
            # moving costs |b0 - a0| gates, rebuilding from the base b0 gates: take the cheaper one
            if abs(b0 - a0) <= b0 and old[:m] == new[:m]:
                if b0 > a0:
                    eng.run(SLOT_R, SLOT_R, stream(new[a0:b0]))
                elif b0 < a0:
                    eng.run(SLOT_R, SLOT_R, stream(old[b0:a0]), inverse=True)
                self.rwin = list(new[:b0])
                self.r_moves += 1
                self.stats["moves_R"] += 1
                return True
        eng.run(SLOT_R, SLOT_BASE, stream(new[:b0]))
        self.rwin = list(new[:b0])
        self.r_slot = SLOT_R
        self.r_moves = 0
        self.stats["rebuild_R"] += 1
        return True

    # ---- L: suffix^+ applied to |0..0> ----
    def _update_L_dense(self, new, b1):
        return self._update_L_to(new[b1:])

    MIDDLE_MAX_GATES = int(os.environ.get("B200AQC_MIDDLE_MAX", "32"))
    LAZY_MAX_GATES = int(os.environ.get("B200AQC_LAZY_MAX", "32"))
    lazy_bra = os.environ.get("B200AQC_LAZY", "1") != "0"

    fused_passes = os.environ.get("B200AQC_FUSED", "1") != "0"

    def _move_L(self, gate_stream, inverse, fuse, store=True):
        """One in-place sweep of slot L; fuse = (other slot, qa, qb): the transfer matrix against that slot is taken in the
        same pass over the register (b200_sv_run_inner2) and left in self._fused_T.  store=False (fused only): T is all
        that is wanted -- returns False if slot L was left as it was."""
        eng = self.eng
        if fuse is not None and self.fused_passes and eng.num_qubits >= getattr(eng, "FUSED_MIN_QUBITS", 1 << 30):
            self.stats["fused_T"] = self.stats.get("fused_T", 0) + 1
            if store:
                self._fused_T = eng.run_inner2(SLOT_L, SLOT_L, gate_stream, fuse[0], fuse[1], fuse[2], inverse=inverse)
                return True
            self._fused_T, stored = eng.run_inner2(SLOT_L, SLOT_L, gate_stream, fuse[0], fuse[1], fuse[2], inverse=inverse,
                                                   store=False)
            return stored
        eng.run(SLOT_L, SLOT_L, gate_stream, inverse=inverse)
        return True

    def _update_L_to(self, sfx, fuse=None, next_gates=0):
        """Make slot L hold sfx^+ |0..0> (sfx: gate list applied in order to a ket).  Returns True if the slot changed.
        fuse: see _move_L (self._fused_T is None afterwards unless the update was one fused sweep).  next_gates: size of the
        block the optimiser will open next from this bra (0: none at this level), see the lazy-bra rule below."""
        eng, stream = self.eng, G.GateStream.from_window
        old = self.lwin
        self._fused_T = None
        if old is not None and old == sfx:
            return False
        split = False
        if (fuse is not None and self.lazy_bra and self.fused_passes and self.dense_blocks and hasattr(eng, "run_embedded")
                and eng.num_qubits >= getattr(eng, "FUSED_MIN_QUBITS", 1 << 30)):
            split = self._embed_split(sfx)
        if split:
            # T is all that is wanted and the tail of this bra lives on a smaller engine: ONE read of the ket, the bra is
            # formed tile by tile in shared memory from that engine's slot and never written (cheaper than any move of
            # the stored bra, which costs a second read of the register)
            T, stored = self._bra_into(SLOT_L, sfx, fuse=fuse, store=False, split=split)
            if T is not None:
                self._fused_T = T
                if stored:
                    self.lwin = list(sfx); self.l_moves = 0
                    self.stats["rebuild_L"] += 1
                else:
                    self.stats["virtual_L"] = self.stats.get("virtual_L", 0) + 1
                return True
            if stored:                      # (the bra was built without a transfer matrix: the caller takes it)
                self.lwin = list(sfx); self.l_moves = 0
                self.stats["rebuild_L"] += 1
                return True
        if old is not None and self.l_moves < self.REFRESH_MOVES:
            so, sn = len(old), len(sfx)
            if sn <= so and so - sn <= sn and old[so - sn:] == sfx:
                self._move_L(stream(old[:so - sn]), False, fuse)
                self.lwin = list(sfx); self.l_moves += 1; self.stats["moves_L"] += 1
                return True
            if sn > so and sn - so <= sn and sfx[sn - so:] == old:
                self._move_L(stream(sfx[:sn - so]), True, fuse)
                self.lwin = list(sfx); self.l_moves += 1; self.stats["moves_L"] += 1
                return True
            # middle replacement: old = P + X + S, new = P + Y + S  =>  L_new = P^+ Y^+ X P L_old = Y^+ X L_old when the
            # gates of P act on none of the qubits of X and Y (front mode: one block leaves the bra, its neighbour enters)
            p = 0
            lim = min(so, sn)
            while p < lim and old[p] == sfx[p]:
                p += 1
            q = 0
            while q < lim - p and old[so - 1 - q] == sfx[sn - 1 - q]:
                q += 1
            X, Y = old[p:so - q], sfx[p:sn - q]
            if 0 < len(X) + len(Y) <= self.MIDDLE_MAX_GATES:
                xy = set()
                for e in X + Y:
                    xy.update(_support(e))
                if all(xy.isdisjoint(_support(e)) for e in old[:p]):
                    # a short replacement whose T is all that is wanted leaves slot L where it is: the next block is
                    # then reached from the same stored bra with a (longer) replacement of its own, and each block costs
                    # two reads of the register instead of two reads and a write
                    # ... unless the NEXT block's replacement from that stored bra would no longer fit (it would fall back to a
                    # rebuild from |0..0> plus a separate transfer pass): then this pass stores and the chain starts afresh
                    lazy = (fuse is not None and self.lazy_bra and len(X) + len(Y) <= self.LAZY_MAX_GATES
                            and not (next_gates and len(X) + len(Y) + 2 * next_gates > self.MIDDLE_MAX_GATES))
                    stored = self._move_L(stream(list(X) + G.invert_window(Y)), False, fuse, store=not lazy)
                    self.stats["middle_L"] = self.stats.get("middle_L", 0) + 1
                    if stored:
                        self.lwin = list(sfx); self.l_moves += 1
                    else:
                        self.stats["lazy_L"] = self.stats.get("lazy_L", 0) + 1
                    return True
        # rebuild: suffix^+ |0> is supported on the qubits the suffix touches -- built on the smaller engines as far as it
        # fits them (cheap sweeps), embedded level by level, only the remaining head gates are applied at this size
        if self.dense_blocks:
            self._fused_T = self._bra_into(SLOT_L, sfx, fuse=fuse, split=split)
        else:
            eng.run(SLOT_L, -1, stream(sfx), inverse=True)
        self.lwin = list(sfx)
        self.l_moves = 0
        self.stats["rebuild_L"] += 1
        return True

    def _compact_map(self, sfx, pair):
        """Qubits the compact bra must carry, or None if it does not fit the compact engine."""
        if not self.compacts or len(pair) != 2:
            return None
        touched = set(pair)
        for e in sfx:
            touched.add(e[1])
            if e[2] >= 0:
                touched.add(e[2])
        fits = [c for c in self.compacts if c.num_qubits >= len(touched)]
        if not fits:
            return None
        self.compact = fits[0]
        qmap = sorted(touched)
        free = [q for q in range(self.eng.num_qubits) if q not in touched]
        return qmap + free[:self.compact.num_qubits - len(qmap)]   # pad: unused compact bits stay |0>

    def _update_L_compact(self, new, b1, qmap):
        sfx = new[b1:]
        key = (tuple(sfx), tuple(qmap))
        if key == self.lkey:
            return False
        pos = {q: c for c, q in enumerate(qmap)}
        remapped = [(e[0], pos[e[1]], pos[e[2]] if e[2] >= 0 else -1) + tuple(e[3:]) for e in sfx]
        self.compact.run(0, -1, G.GateStream.from_window(remapped), inverse=True)
        self.lkey = key
        self.stats["compact_L"] += 1
        return True

    def _open_block_ok(self, window, changed):
        """The edits stay inside the block T is open on (the fast path of amp0 applies)."""
        return (changed is not None and self.T is not None and self.window is not None and len(self.window) == len(window)
                and all(self.cut[0] <= i < self.cut[1] for i in changed))

    def _compact_ok(self, window, block):
        """dense_blocks = False: a block can be served only with a compact bra (gather of R)."""
        b0, b1, supp = block
        return len(supp) == 2 and self._compact_map(window[b1:], tuple(supp)) is not None

    front_mode = os.environ.get("B200AQC_FRONT", "1") != "0"       # see _prepare_block
    FRONT_MAX_PREFIX = 96   # gates

    def _front_ok(self, window, b0, pair):
        """The block at [b0, ...) on `pair` commutes to the front of the window: every gate before it avoids its qubits."""
        if not (self.front_mode and self.dense_blocks) or len(pair) != 2 or not 0 < b0 <= self.FRONT_MAX_PREFIX:
            return False
        a, b = pair
        for e in window[:b0]:
            if e[1] == a or e[1] == b or e[2] == a or e[2] == b:
                return False
        return True

    def _prefetch_next_L(self, window, b1):
        """T of the open block has been read back, so slot L is free: enqueue (asynchronously) the bra of the
        NEXT block -- Rotosolve walks the blocks in order (cost_minimiser.py:267-316) -- while the host
        evaluates the open one.  Only bookkeeping (lwin) records it: if the optimiser goes elsewhere the
        usual move / rebuild logic starts from this state."""
        nxt = next((b for b in self._blocks(window) if b[0] == b1), None)
        if nxt is None or self.lwin is None or self.l_moves >= self.REFRESH_MOVES:
            return
        if self.lwin != list(window[b1:]):
            return          # T came from a pass that left slot L at an earlier bra (lazy middle replacement)
        n0, n1, supp = nxt
        if len(supp) == 2 and self._front_ok(window, n0, tuple(supp)):
            return          # the next block commutes to the front: its bra is one middle replacement away from this one
        if len(supp) == 2 and self._compact_map(window[n1:], tuple(supp)) is not None:
            return          # the next block will use a compact bra
        split = self._tail_split(window)
        if split is not None and n0 >= split[0]:
            return          # ... or the projected tail
        self.eng.run(SLOT_L, SLOT_L, G.GateStream.from_window(window[n0:n1]))
        self.lwin = list(window[n1:])
        self.l_moves += 1
        self.stats["prefetched_L"] = self.stats.get("prefetched_L", 0) + 1

    def _prepare_block(self, window, block):
        """Make R, L and T valid for `block` of `window`."""
        eng = self.eng
        b0, b1, supp = block
        n = eng.num_qubits
        pair = tuple(supp) if (len(supp) == 2 or n == 1) else (supp[0], (supp[0] + 1) % n)
        if self._front_ok(window, b0, pair):
            # FRONT MODE: every gate in front of the block acts on other qubits, so the block commutes to the front of the
            # window:  <0| W[b1:] B W[:b0] |base> = <0| W[b1:] W[:b0] B |base>.  The ket side is the base state itself (no
            # R move); the bra carries W[:b0] + W[b1:], and going from one such block to the next replaces ONE block in
            # the middle of that list (_update_L_to): one sweep per block instead of an R move plus a bra move.
            sfx = list(window[:b0]) + list(window[b1:])
            tkey = (b0, b1, pair, "front")
            if self.T is not None and self._tkey == tkey and self._t_sfx == sfx:
                self.cut, self.pair = (b0, b1), pair
                self.window = list(window)
                self._gw = None
                return
            # (the block that follows, if the optimiser will meet it at this level in front mode too)
            nxt = next((b for b in self._blocks(window) if b[0] == b1), None)
            next_gates = 0
            if nxt is not None and len(nxt[2]) == 2 and self._front_ok(window, nxt[0], tuple(nxt[2])):
                split = self._tail_split(window) if self.projected else None
                if split is None or nxt[0] < split[0]:
                    next_gates = nxt[1] - nxt[0]
            self._update_L_to(sfx, fuse=(SLOT_BASE, pair[0], pair[1]), next_gates=next_gates)
            self.T = self._fused_T if self._fused_T is not None else eng.inner2(SLOT_L, SLOT_BASE, *pair)
            self._fused_T = None
            self.stats["t_passes"] += 1
            self.stats["front_blocks"] = self.stats.get("front_blocks", 0) + 1
            self._tkey, self._t_sfx = tkey, sfx
            self.cut, self.pair = (b0, b1), pair
            self.window = list(window)
            self._gw = None
            return
        r_changed = self._update_R(window, b0)
        if (not r_changed and self.T is not None and self._tkey is not None and self._tkey[:3] == (b0, b1, pair)
                and self._t_sfx == window[b1:]):
            # T is still valid for this block (slot L may already hold the prefetched bra of the next one)
            self.cut, self.pair = (b0, b1), pair
            self.window = list(window)
            self._gw = None
            return
        qmap = self._compact_map(window[b1:], pair)
        if qmap is not None:
            l_changed = self._update_L_compact(window, b1, qmap)
            mode = "compact"
        else:
            self._fused_T = None
            l_changed = self._update_L_to(window[b1:], fuse=(self.r_slot, pair[0], pair[1]) if len(pair) == 2 else None)
            mode = "dense"
        tkey = (b0, b1, pair, mode)
        if mode == "dense" and self._fused_T is not None:
            self.T, self._fused_T = self._fused_T, None       # the bra moved in one sweep: T came out of the same pass
            self.stats["t_passes"] += 1
            self._tkey = tkey
            self._t_sfx = list(window[b1:])
            self._gw = None
            if self.prefetch_L:
                self._prefetch_next_L(window, b1)
        elif r_changed or l_changed or self.T is None or tkey != self._tkey:
            if mode == "compact":
                self.compact.sync()
                self.T = eng.inner2_gather(self.r_slot, self.compact, 0, qmap, *pair)
                self.stats["t_gathers"] += 1
            else:
                self.T = (eng.inner(SLOT_L, self.r_slot, pair[0]) if len(pair) == 1
                          else eng.inner2(SLOT_L, self.r_slot, *pair))
                self.stats["t_passes"] += 1
            self._tkey = tkey
            self._t_sfx = list(window[b1:])
            self._gw = None
            if mode == "dense" and self.prefetch_L:
                self._prefetch_next_L(window, b1)
        self.cut, self.pair = (b0, b1), pair
        self.window = list(window)
        self._gw = None

    _pwin = None
    project_direct = os.environ.get("B200AQC_PROJECT_DIRECT", "1") != "0"
    _tkey = None
    _t_sfx = None
    _gw = None          # (k, w): gate-context cache of the open block, see _gate_context
    _hot = None         # (k, w, window length, qubit): the gate the optimiser is working on, see amp0
    _hot_dirty = False  # a hot-path value was served: the nested levels have not seen that gate's latest entry

    def _gate_context(self, window, k):
        """(w00, w01, w10, w11), Python complex, such that for ANY 2x2 matrix m placed at window position k of the open
        block  <0|psi> = sum_ab m[a][b] w_ab.  With pre / post = the products of the block's other gates before / after
        position k:  <0|psi> = sum_ij (post E(m) pre)_ij T_ij = tr(E(m) . pre T^T post), so w is the partial trace of
        W = pre T^T post over the block's other qubit.  The optimiser asks for 3-7 values of the SAME gate in a row
        (cost_minimiser.py:318-368): they all share w, and each costs four complex multiplications on the host.
        Cached until T, or another gate of the block, changes."""
        gw = self._gw
        if gw is not None and gw[0] == k:
            return gw[1]
        b0, b1 = self.cut
        pair = self.pair
        pre = post = None
        for i in range(b0, k):
            mm = embed_entry(window[i], pair)
            pre = mm if pre is None else mm @ pre
        for i in range(k + 1, b1):
            mm = embed_entry(window[i], pair)
            post = mm if post is None else mm @ post
        W = self.T.T
        if pre is not None:
            W = pre @ W
        if post is not None:
            W = W @ post
        if len(pair) == 1:
            w = (complex(W[0, 0]), complex(W[1, 0]), complex(W[0, 1]), complex(W[1, 1]))
        elif window[k][1] == pair[0]:       # the gate's qubit is the low index bit: idx(a, o) = a + 2 o
            w = tuple(complex(W[b, a] + W[b + 2, a + 2]) for a in (0, 1) for b in (0, 1))
        else:                               # high index bit: idx(a, o) = o + 2 a
            w = tuple(complex(W[2 * b, 2 * a] + W[2 * b + 1, 2 * a + 1]) for a in (0, 1) for b in (0, 1))
        self._gw = (k, w)
        return w

    def _operator(self, window, override_index=None, override=None):
        b0, b1 = self.cut
        op = None
        for i in range(b0, b1):
            m = embed_entry(window[i], self.pair, override if i == override_index else None)
            op = m if op is None else m @ op
        return op

    # ---- evaluation ----
    def amp0(self, window, focus=None, changed=None, prev_k=None):
        """<0| W |base> for the canonical window `window` (one scalar, as the reference asks).
        `focus`: index of the gate the optimiser is most likely to edit next (block choice when the
        structure changed).  `changed`: indices at which `window` differs from the previous call's
        window (None = unknown), so that edits inside the open block skip all bookkeeping."""
        self.stats["evals"] += 1
        # HOT PATH: the optimiser asks for another value of the gate it has just been given a value for
        # (cost_minimiser.py:356-363: theta = 0, pi/2, -pi/2 of ONE gate; 7 values in Rotoselect) and nothing else
        # changed: the gate context w computed by whichever nested level owns that gate still holds, the value is four
        # complex multiplications -- no block bookkeeping, no descent through the projection levels.  The levels below
        # have then not seen the latest entry of that gate: it is added to `changed` when the hot path is left.
        hot = self._hot
        self._prev_k = hot[0] if hot is not None else prev_k       # (prev_k: the same hint from the level above)
        if hot is not None:
            if (changed is not None and len(changed) == 1 and changed[0] == hot[0] and len(window) == hot[2]
                    and window[hot[0]][2] < 0 and window[hot[0]][1] == hot[3]):
                self._hot_dirty = True
                self.stats["hot_evals"] = self.stats.get("hot_evals", 0) + 1
                w = hot[1]
                m = G.matrix_of_entry_complex(window[hot[0]])
                return m[0] * w[0] + m[1] * w[1] + m[2] * w[2] + m[3] * w[3]
            if self._hot_dirty and changed is not None:
                changed = sorted(set(changed) | {hot[0]})
            self._hot, self._hot_dirty = None, False
        if len(window) == 0:
            self.invalidate()
            return self.eng.amp(SLOT_BASE, 0)
        if self.projected:
            target = self._current_gate(changed)
            if target is None:
                target = len(window) - 1 if focus is None else min(max(focus, 0), len(window) - 1)
            pj = self._projected(window, target, changed)
            if pj is not None:
                sub, tail, sub_changed, m = pj
                self.stats["projected_evals"] += 1
                pk = self._prev_k
                # (the nested level may have nothing to diff against -- it was re-based by the projection: its block
                # choice then follows `focus`, which is the gate the optimiser is working on whenever that is known)
                sub_focus = target - m if changed else (None if focus is None else max(focus - m, 0))
                out = sub.amp0(tail, focus=sub_focus, changed=sub_changed,
                               prev_k=pk - m if pk is not None and pk >= m else None)
                sh = sub._hot
                if sh is not None and not sub._hot_dirty:
                    e = window[sh[0] + m]
                    self._hot = (sh[0] + m, sh[1], len(window), e[1])      # the same context, in this level's numbering
                    sub._hot = None                                          # (only the top level takes the hot path)
                return out
        if not self.dense_blocks and not self._open_block_ok(window, changed) \
                and not self._compact_ok(window, self._select_block(window, focus, changed)):
            self.stats["resimulations"] = self.stats.get("resimulations", 0) + 1
            self.T = None
            self.window = None
            # slot R / phi were NOT advanced to this window: a later tail-only edit must not take the
            # "only tail gates changed since phi was gathered" shortcut of _projected
            self._proj_state = None
            self._work_tail = None
            self.eng.run(SLOT_WORK, SLOT_BASE, G.GateStream.from_window(window))
            return self.eng.amp(SLOT_WORK, 0)
        if (changed is not None and self.T is not None and self.window is not None and len(self.window) == len(window)
                and all(self.cut[0] <= i < self.cut[1] and window[i][1] == self.window[i][1]
                        and window[i][2] == self.window[i][2] for i in changed)):
            for i in changed:
                self.window[i] = window[i]
            # pivot = the gate the optimiser is working on (the last one it touched): all 3-7 values it asks for in a
            # row are then computed with the SAME arithmetic (one gate context w, four multiplications each), exactly
            # like the batched front end -- Rotoselect's exact ties between axes break the same way in both
            gw = self._gw
            k = max(changed) if changed else (gw[0] if gw is not None else None)
            if k is not None and window[k][2] < 0:
                if gw is not None and (gw[0] != k or len(changed) > 1):
                    self._gw = None            # another gate of the block changed: its context is stale
                self.stats["host_evals"] += 1
                w = self._gate_context(window, k)
                self._hot = (k, w, len(window), window[k][1])
                m = G.matrix_of_entry_complex(window[k])
                return m[0] * w[0] + m[1] * w[1] + m[2] * w[2] + m[3] * w[3]
            self._gw = None
        else:
            self._prepare_block(window, self._select_block(window, focus, changed))
            # same arithmetic as the fast path / the batched front end whenever a pivot gate is known
            b0, b1 = self.cut
            inside = [i for i in (changed or ()) if b0 <= i < b1]
            k = max(inside) if inside else None
            if k is not None and window[k][2] < 0:
                self.stats["host_evals"] += 1
                w = self._gate_context(window, k)
                self._hot = (k, w, len(window), window[k][1])
                m = G.matrix_of_entry_complex(window[k])
                return m[0] * w[0] + m[1] * w[1] + m[2] * w[2] + m[3] * w[3]
        self.stats["host_evals"] += 1
        return complex(np.sum(self._operator(window) * self.T))

    # ---- batched API (K6): every shift value of one gate from one transfer pass ----
    def shift_amplitudes(self, window, k, candidates, changed=None):
        """<0|psi> for each replacement 2x2 matrix in `candidates` at window position k.  `changed`: as in amp0."""
        if self._hot is not None:
            if self._hot_dirty and changed is not None:
                changed = sorted(set(changed) | {self._hot[0]})
            self._hot, self._hot_dirty = None, False
        if self.projected:
            pj = self._projected(window, k, changed)
            if pj is not None:
                sub, tail, sub_changed, m = pj
                self.stats["evals"] += len(candidates)
                self.stats["projected_evals"] += len(candidates)
                return sub.shift_amplitudes(tail, k - m, candidates, sub_changed)
        if not self.dense_blocks and not any(b[0] <= k < b[1] and self._compact_ok(window, b) for b in self._blocks(window, changed)):
            out = []
            for cand in candidates:
                w = list(window)
                w[k] = ("mat1", window[k][1], -1, 0.0, 0.0, 0.0, np.ascontiguousarray(cand, dtype=np.complex128).tobytes())
                out.append(self.amp0(w))
            return out
        for b in self._blocks(window, changed):
            if b[0] <= k < b[1]:
                self._prepare_block(window, b)
                break
        self.stats["evals"] += len(candidates)
        self.stats["host_evals"] += len(candidates)
        # Every candidate shares the gate context w of position k (_gate_context): the same four complex multiplications
        # per value as the one-scalar-at-a-time path (amp0), so both front ends produce bit-identical costs.
        self._gw = None if (self._gw is not None and self._gw[0] != k) else self._gw
        w = self._gate_context(window, k)
        out = []
        for c in candidates:
            m = G.complex_tuple(c)
            out.append(m[0] * w[0] + m[1] * w[1] + m[2] * w[2] + m[3] * w[3])
        return out
