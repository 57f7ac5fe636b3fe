"""Statevectors too large for one GPU: sharding by global qubits (BASELINE config C5).

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  The 2^n amplitudes are
split by the top g = log2(world) *physical* index bits: rank r holds the 2^(n-g) amplitudes whose
global bits equal r.  A logical -> physical qubit permutation is kept per state:

  * gates whose mixing targets are all local run through the single-GPU fused sweep kernels on the
    local slice (``SVEngine``), unchanged; a control or diagonal qubit that is global only selects
    or phases the whole slice (decided on the host from the rank's bits);
  * before a gate that mixes a global qubit, an all-to-all exchange swaps the g global qubits with
    g local ones (the local qubits used furthest in the future): every rank sends the chunk with
    top-local bits = p to peer p and receives the peer's chunk -- (1 - 1/G) of the slice crosses
    NVLink once.  The exchange is done in place through a small staging buffer so that a slot can
    fill most of the 180 GB.

Read-outs reduce across ranks with one small all-reduce: amplitude 0, all <Z_q>, pair RDMs.

The reference has no distributed path at all; this module is what lets the same
backend interface reach 34 qubits.  Host-side logic is tested on CPU with gloo, world_size 2
(tests/test_dist_gloo.py).
"""
import numpy as np

from . import gates as G

# gates that act diagonally on every qubit they touch
_DIAG_1Q = {"z", "s", "sdg", "t", "tdg", "rz", "u1", "p", "id", "i"}


def mixing_qubits(ent):
    """Qubits on which a canonical window entry acts non-diagonally (they must be local)."""
    name, q0, q1 = ent[0], ent[1], ent[2]
    if q1 < 0:
        if name in _DIAG_1Q:
            return ()
        if name == "mat1":
            m = np.frombuffer(ent[6], dtype=np.complex128)
            return () if (m[1] == 0 and m[2] == 0) else (q0,)
        return (q0,)
    if name == "cx":
        return (q1,)
    if name == "cz":
        return ()
    if name == "mat2":
        m = np.frombuffer(ent[6], dtype=np.complex128).reshape(4, 4)
        if np.count_nonzero(m - np.diag(np.diag(m))) == 0:
            return ()
    return (q0, q1)


class TorchComm:
    """Thin wrapper over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device if device is not None else torch.device("cpu")
        self.bytes_sent = 0
        self.exchange_ms = 0.0

    def allreduce_engine_slot(self, engine, slot):
        """Sum slot `slot` of `engine` (a whole small register, replicated per rank) over the ranks, in place."""
        if hasattr(engine, "slots"):                      # CPU stand-in engine (gloo tests)
            a = engine.slots[slot]
            a[...] = self.allreduce_sum(a.view(np.float64)).view(np.complex128)
            return
        engine.sync()
        t = self.torch.as_tensor(_CudaAlias(engine.device_ptr(slot), 2 << engine.num_qubits), device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        self.torch.cuda.synchronize(self.device)

    def allreduce_sum(self, arr):
        t = self.torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64)).to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def exchange_chunks(self, state, staging):
        """In-place all-to-all of the `world` equal chunks of the 1-D tensor `state`: chunk[p] goes
        to rank p, whose chunk[rank] comes back into the same place.  `staging`: scratch tensor."""
        torch, dist = self.torch, self.dist
        Gw = self.world
        chunk = state.numel() // Gw
        timing = state.is_cuda
        if timing:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        # Double-buffered staging: the transfer of piece k+1 is issued (it runs on NCCL's stream) before
        # this stream copies piece k back into place, so the copy-back hides behind the next transfer.
        half = max(1, staging.numel() // 2)
        piece = min(chunk, half)
        bufs = (staging[:piece], staging[half:half + piece]) if 2 * piece <= staging.numel() else (staging[:piece],)
        jobs = [(self.rank ^ step, off) for step in range(1, Gw) for off in range(0, chunk, piece)]

        def issue(k):
            peer, off = jobs[k]
            cnt = min(piece, chunk - off)
            mine = state[peer * chunk + off: peer * chunk + off + cnt]
            buf = bufs[k % len(bufs)][:cnt]
            works = dist.batch_isend_irecv([dist.P2POp(dist.isend, mine, peer), dist.P2POp(dist.irecv, buf, peer)])
            return works, mine, buf

        pending = issue(0) if jobs else None
        for k in range(len(jobs)):
            works, mine, buf = pending
            pending = issue(k + 1) if (k + 1 < len(jobs) and len(bufs) > 1) else None
            for w in works:
                w.wait()
            mine.copy_(buf)
            self.bytes_sent += mine.numel() * state.element_size()
            if pending is None and k + 1 < len(jobs):
                pending = issue(k + 1)
        if timing:
            e1.record()
            e1.synchronize()
            self.exchange_ms += e0.elapsed_time(e1)


class ShardedStatevector:
    """`n_slots` sharded states of `num_qubits` qubits over `comm.world` ranks.

    engine: local engine over n_local = num_qubits - g qubits (SVEngine on the GPU; any object with
            run/amp/expz/pair_rdm/inner/copy for the CPU tests);
    slot_tensors: one flat torch tensor (float64, 2 * 2^n_local) per slot, aliasing the engine's
            slot memory -- what the communicator sends from / receives into."""

    def __init__(self, num_qubits, engine, comm, slot_tensors, staging, sync=None):
        self.n = int(num_qubits)
        self.comm = comm
        self.g = int(np.log2(comm.world))
        if (1 << self.g) != comm.world:
            raise ValueError("world size must be a power of two")
        self.nl = self.n - self.g
        if self.nl < self.g or engine.num_qubits != self.nl:
            raise ValueError("local engine must hold num_qubits - log2(world) qubits")
        self.eng = engine
        self.slot_tensors = slot_tensors
        self.staging = staging
        self._sync = sync or (lambda: None)
        # perm[slot][logical] = physical position; positions >= nl are rank bits
        self.perm = [list(range(self.n)) for _ in slot_tensors]
        self.stats = {"exchanges": 0, "local_runs": 0}
        self.peer_ptrs = None      # [slot][rank] -> device pointer into that rank's slot (CUDA IPC), or None: NCCL path
        self.strided_exchange = True    # peer path: exchange the victims in place (no SWAP-localisation sweep)
        self.reorder = True        # run(): commutation-aware scheduling that postpones gates on global qubits

    def close(self):
        """Collective: every rank unmaps its peers' slots (CUDA IPC) before any rank frees its own."""
        if self.peer_ptrs is not None:
            self._sync()
            self.comm.dist.barrier()
            for row in self.peer_ptrs:
                for ptr in row:
                    if ptr:
                        self.eng.ipc_close(ptr)
            self.peer_ptrs = None
            self.comm.dist.barrier()
        self.eng.close()

    # ---- layout ----
    def _rank_bit(self, phys):
        return (self.comm.rank >> (phys - self.nl)) & 1

    def _is_global(self, slot, logical):
        return self.perm[slot][logical] >= self.nl

    def _exchange(self, slot, victims):
        """Swap the g global qubits with the local logical qubits `victims` (len g)."""
        perm = self.perm[slot]
        nl, g = self.nl, self.g
        inv = {p: l for l, p in enumerate(perm)}
        if self.peer_ptrs is not None and self.strided_exchange:
            # peer-memory path: the victims stay where they are -- chunk p is the strided set of amplitudes whose bits
            # at the victims' positions spell p (b200_sv_peer_swap_strided): no SWAP-localisation sweep
            positions = [perm[v] for v in victims]
            self._sync()
            self._peer_exchange(slot, positions)
            self._sync()
            for j, v in enumerate(victims):
                glob = inv[nl + j]
                perm[glob], perm[v] = positions[j], nl + j
            self.stats["exchanges"] += 1
            return
        # 1. move the victims to the top-g local physical positions with local swap gates
        swaps = []
        for j, v in enumerate(victims):
            want = nl - g + j
            have = perm[v]
            if have != want:
                other = inv[want]
                swaps.append(("swap", have, want, 0.0, 0.0, 0.0, None))
                perm[v], perm[other] = want, have
                inv[want], inv[have] = v, other
        if swaps:
            self.eng.run(slot, slot, G.GateStream.from_window(swaps))
            self.stats["local_runs"] += 1
            self.stats["localisation_sweeps"] = self.stats.get("localisation_sweeps", 0) + 1
        # 2. all-to-all: chunk index (top-g local bits) <-> rank bits
        self._sync()
        if self.peer_ptrs is not None:
            self._peer_exchange(slot)
        else:
            self.comm.exchange_chunks(self.slot_tensors[slot], self.staging)
        self._sync()
        for j in range(g):
            a, b = inv[nl - g + j], inv[nl + j]
            perm[a], perm[b] = nl + j, nl - g + j
        self.stats["exchanges"] += 1

    def _peer_exchange(self, slot, positions=None):
        """One fused kernel per rank over NVLink peer memory (b200_sv_peer_swap[_strided]): no staging, no copy-back.
        Ranks meet before (every slice is final) and after (every remote store has landed)."""
        import time
        comm = self.comm
        comm.dist.barrier()
        t0 = time.perf_counter()
        self.eng.peer_swap(slot, self.peer_ptrs[slot], comm.rank, positions)
        self.eng.sync()
        comm.exchange_ms += 1e3 * (time.perf_counter() - t0)
        comm.dist.barrier()
        comm.bytes_sent += (comm.world - 1) * (16 << self.nl) // comm.world

    def ensure_local(self, slot, needed, next_use=None):
        """Make every logical qubit in `needed` local; victims = local qubits not needed, used
        furthest in the future (next_use[logical], larger = later)."""
        perm = self.perm[slot]
        if not any(perm[q] >= self.nl for q in needed):
            return
        cands = [l for l in range(self.n) if perm[l] < self.nl and l not in needed]
        if len(cands) < self.g:
            raise ValueError("too many qubits must be local at once for this sharding")
        if self.peer_ptrs is not None and self.strided_exchange:
            # strided exchange: a victim at bit position p leaves runs of 2^p contiguous amplitudes (16 B each);
            # keep the runs >= 1 KB when there is a choice
            high = [l for l in cands if perm[l] >= 6]
            if len(high) >= self.g:
                cands = high
        cands.sort(key=lambda l: (-(next_use[l] if next_use is not None else 0), -perm[l]))
        self._exchange(slot, cands[:self.g])

    # ---- gate application ----
    def _localise(self, slot, ent):
        """Translate one canonical entry to the local physical qubits of this rank; returns a list
        of entries (possibly empty) -- global control / diagonal qubits are resolved here."""
        perm, nl = self.perm[slot], self.nl
        name, q0, q1 = ent[0], ent[1], ent[2]
        p0 = perm[q0]
        if q1 < 0:
            if p0 < nl:
                return [(name, p0, -1) + ent[3:]]
            m = G.matrix_of_entry(ent)                      # diagonal on a global qubit: a scalar
            ph = m[1, 1] if self._rank_bit(p0) else m[0, 0]
            return [] if ph == 1 else [("mat1", 0, -1, 0.0, 0.0, 0.0, np.diag([ph, ph]).astype(np.complex128).tobytes())]
        p1 = perm[q1]
        if p0 < nl and p1 < nl:
            return [(name, p0, p1) + ent[3:]]
        if name == "cx":                                    # control is global (target is local by construction)
            return [("x", p1, -1, 0.0, 0.0, 0.0, None)] if self._rank_bit(p0) else []
        if name == "cz":
            if p0 >= nl and p1 >= nl:
                both = self._rank_bit(p0) and self._rank_bit(p1)
                return [("mat1", 0, -1, 0.0, 0.0, 0.0, np.diag([-1, -1]).astype(np.complex128).tobytes())] if both else []
            loc, glob = (p1, p0) if p0 >= nl else (p0, p1)
            return [("z", loc, -1, 0.0, 0.0, 0.0, None)] if self._rank_bit(glob) else []
        # diagonal mat2 with a global qubit
        d = np.diag(np.frombuffer(ent[6], dtype=np.complex128).reshape(4, 4))     # index bit(q0) + 2 bit(q1)
        if p0 >= nl and p1 >= nl:
            ph = d[self._rank_bit(p0) + 2 * self._rank_bit(p1)]
            return [("mat1", 0, -1, 0.0, 0.0, 0.0, np.diag([ph, ph]).astype(np.complex128).tobytes())]
        if p0 >= nl:
            b = self._rank_bit(p0)
            return [("mat1", p1, -1, 0.0, 0.0, 0.0, np.diag([d[b], d[b + 2]]).astype(np.complex128).tobytes())]
        b = self._rank_bit(p1)
        return [("mat1", p0, -1, 0.0, 0.0, 0.0, np.diag([d[2 * b], d[2 * b + 1]]).astype(np.complex128).tobytes())]

    def run(self, dst, src, window):
        """dst <- window applied to src (src = -1: |0..0>).  `window`: canonical entries on logical
        qubits (gates.canonical_window)."""
        eng = self.eng
        if src < 0:
            self.perm[dst] = list(range(self.n))
            if self.comm.rank == 0:
                eng.run(dst, -1, G.GateStream.from_window([]))
            else:
                eng.run(dst, -1, G.GateStream.from_window([("mat1", 0, -1, 0.0, 0.0, 0.0,
                                                               np.zeros((2, 2), dtype=np.complex128).tobytes())]))
        elif src != dst:
            eng.copy(dst, src)
            self.perm[dst] = list(self.perm[src])
        if not self.reorder:
            return self._run_in_order(dst, window)
        # Commutation-aware scheduling.  Gates are taken in circuit order whenever all their mixing qubits are local; a
        # gate that mixes a GLOBAL qubit is postponed, and later gates may overtake the postponed ones only if they
        # commute with them (no postponed gate touches their mixing qubits; no postponed gate mixes a qubit they touch
        # at all).  When nothing more can be taken, ONE exchange brings the global qubits in -- the victims are the local
        # qubits whose next pending use is furthest away -- and the scan restarts.  On a brickwork target this applies
        # every gate outside the light cone of the global qubits, through all layers, before the first exchange.
        pending = list(window)
        mix = [mixing_qubits(e) for e in pending]
        touch = [tuple(q for q in (e[1], e[2]) if q >= 0) for e in pending]
        while pending:
            batch, rest, rest_mix, rest_touch = [], [], [], []
            blocked_any, blocked_mix = set(), set()
            for e, m, t in zip(pending, mix, touch):
                free = not (set(m) & blocked_any) and not (set(t) & blocked_mix)
                if free and not any(self._is_global(dst, q) for q in m):
                    batch.extend(self._localise(dst, e))
                else:
                    rest.append(e); rest_mix.append(m); rest_touch.append(t)
                    blocked_any.update(t); blocked_mix.update(m)
            if batch:
                eng.run(dst, dst, G.GateStream.from_window(batch))
                self.stats["local_runs"] += 1
            pending, mix, touch = rest, rest_mix, rest_touch
            if not pending:
                break
            # every remaining gate waits (directly or through a postponed gate) on a global qubit: exchange.
            INF = len(pending) + 1
            nxt = [INF] * self.n
            for i in range(len(pending) - 1, -1, -1):
                for q in mix[i]:
                    nxt[q] = i
            needed = {q for q in range(self.n) if self._is_global(dst, q) and nxt[q] < INF}
            if not needed:       # cannot happen: a postponed gate mixes a global qubit
                raise AssertionError("sharded scheduler is stuck without a pending global qubit")
            # the first postponed gate must be executable after the exchange: its local qubits are not victims
            self.ensure_local(dst, needed | set(mix[0]), nxt)

    def _run_in_order(self, dst, window):
        """Circuit order (round-1 behaviour, reorder = False): exchange as soon as a gate mixes a global qubit."""
        eng = self.eng
        mix = [mixing_qubits(e) for e in window]
        # next use (as a mixing target) of every logical qubit, scanning from the back
        INF = len(window) + 1
        nxt = [INF] * self.n
        next_use_at = [None] * len(window)
        for i in range(len(window) - 1, -1, -1):
            for q in mix[i]:
                nxt[q] = i
            next_use_at[i] = list(nxt)
        batch = []

        def flush():
            if batch:
                eng.run(dst, dst, G.GateStream.from_window(batch))
                self.stats["local_runs"] += 1
                batch.clear()

        for i, ent in enumerate(window):
            if any(self._is_global(dst, q) for q in mix[i]):
                flush()
                # bring in this gate's qubits and as many upcoming mixing qubits as are global now
                needed = set(mix[i])
                self.ensure_local(dst, needed, next_use_at[i])
            batch.extend(self._localise(dst, ent))
        flush()

    # ---- read-outs (every rank returns the same values) ----
    def amp(self, slot, index=0):
        perm, nl = self.perm[slot], self.nl
        phys = 0
        for l in range(self.n):
            if (index >> l) & 1:
                phys |= 1 << perm[l]
        owner, local = phys >> nl, phys & ((1 << nl) - 1)
        v = self.eng.amp(slot, local) if owner == self.comm.rank else 0j
        out = self.comm.allreduce_sum(np.array([v.real, v.imag]))
        return complex(out[0], out[1])

    def expz(self, slot):
        perm, nl = self.perm[slot], self.nl
        z_loc, norm_loc = self.eng.expz(slot)             # unnormalised sums over the local slice
        vec = np.zeros(self.n + 1)
        vec[:nl] = z_loc
        for p in range(nl, self.n):
            vec[p] = (1 - 2 * self._rank_bit(p)) * norm_loc
        vec[self.n] = norm_loc
        tot = self.comm.allreduce_sum(vec)
        return np.array([tot[perm[l]] for l in range(self.n)]), float(tot[self.n])

    def pair_rdm(self, slot, pairs):
        """4x4 RDMs (lower logical qubit least significant) for logical `pairs`."""
        pairs = [tuple(p) for p in pairs]
        out = {}
        remaining = list(dict.fromkeys(tuple(sorted(p)) for p in pairs))
        while remaining:
            perm = self.perm[slot]
            ready = [p for p in remaining if perm[p[0]] < self.nl and perm[p[1]] < self.nl]
            if ready:
                phys = [(perm[a], perm[b]) for a, b in ready]
                rho = self.eng.pair_rdm(slot, phys)
                flat = self.comm.allreduce_sum(np.ascontiguousarray(rho).view(np.float64).reshape(-1))
                rho = flat.view(np.complex128).reshape(-1, 4, 4)
                for (a, b), (pa, pb), r in zip(ready, phys, rho):
                    if pa > pb:       # kernel convention: lower PHYSICAL qubit is the LSB -> swap index bits
                        r = r[np.ix_([0, 2, 1, 3], [0, 2, 1, 3])]
                    out[(a, b)] = r
                remaining = [p for p in remaining if p not in out]
            if remaining:
                # make the first blocked pair local; keep the qubits of other remaining pairs if possible
                weight = [0] * self.n
                for a, b in remaining:
                    weight[a] += 1; weight[b] += 1
                self.ensure_local(slot, set(remaining[0]), next_use=[-w for w in weight])
        return np.array([out[tuple(sorted(p))] for p in pairs]).reshape(-1, 4, 4)

    def inner(self, l_slot, r_slot):
        """<L|R>; the two slots must share a layout (true for states produced from the same run)."""
        if self.perm[l_slot] != self.perm[r_slot]:
            raise ValueError("inner product needs identical qubit layouts")
        v = self.eng.inner(l_slot, r_slot, -1)
        out = self.comm.allreduce_sum(np.array([v.real, v.imag]))
        return complex(out[0], out[1])


class ShardedEngine:
    """SVEngine-shaped view of a ShardedStatevector, so that the backend / the incremental evaluator
    (sv_engine.SVCostEvaluator with dense_blocks = False) can drive a register that spans several GPUs:
    gate streams, amplitude / <Z> / pair-RDM read-outs, and the PROJECTION of the register onto |0> of all
    but K qubits (each rank gathers the amplitudes it owns, one all-reduce of 2^K amplitudes assembles them
    on every rank), after which the optimisation of the tail runs on a K-qubit engine replicated per rank.
    Every rank must issue the same calls in the same order."""

    def __init__(self, sharded):
        self.sv = sharded
        self.num_qubits = sharded.n

    def run(self, dst, src, stream, inverse=False):
        window = stream.window
        if window is None:
            raise ValueError("the sharded engine needs gate streams built with GateStream.from_window / from_circuit")
        self.sv.run(dst, src, G.invert_window(window) if inverse else window)

    def copy(self, dst, src):
        self.sv.eng.copy(dst, src)
        self.sv.perm[dst] = list(self.sv.perm[src])

    def amp(self, slot, index=0):
        return self.sv.amp(slot, index)

    def expz(self, slot):
        return self.sv.expz(slot)

    def pair_rdm(self, slot, pairs):
        return self.sv.pair_rdm(slot, pairs)

    def gather(self, slot, qmap, dst_engine, dst_slot):
        sv = self.sv
        phys = [sv.perm[slot][q] for q in qmap]            # positions >= nl are rank bits
        sv.eng.gather_ranked(slot, phys, sv.g, sv.comm.rank, dst_engine, dst_slot)
        sv._sync()
        sv.comm.allreduce_engine_slot(dst_engine, dst_slot)

    def inner2_gather(self, r_slot, compact_engine, compact_slot, qmap, qa, qb):
        """Transfer matrix against a compact bra: the 2^K amplitudes of the sharded R on qmap are gathered
        (ranked gather + all-reduce) into a spare slot of the compact engine, then the ordinary transfer
        pass runs between the two 2^K-amplitude states."""
        if getattr(compact_engine, "n_slots", len(getattr(compact_engine, "slots", [0, 0]))) < 2:
            raise ValueError("the compact engine of a sharded backend needs two slots")
        spare = 1 if compact_slot == 0 else 0
        qmap = list(qmap)
        self.gather(r_slot, qmap, compact_engine, spare)
        return compact_engine.inner2(compact_slot, spare, qmap.index(qa), qmap.index(qb))

    def download(self, slot, offset=0, count=None):
        raise MemoryError("a sharded statevector is not downloaded to one host")

    def sync(self):
        self.sv._sync()

    def counters(self):
        return self.sv.eng.counters()

    def profile(self, enable=True):
        self.sv.eng.profile(enable)

    def profile_read(self):
        return self.sv.eng.profile_read()

    def close(self):
        self.sv.close()


class _CudaAlias:
    """Lets torch view library-owned device memory (torch.as_tensor over __cuda_array_interface__)."""

    def __init__(self, ptr, n_doubles):
        self.__cuda_array_interface__ = {"shape": (int(n_doubles),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def make_gpu_sharded(num_qubits, n_slots=2, staging_bytes=1 << 30, local_rank=0, exchange=None, engine=None):
    """Builds a ShardedStatevector over the default process group (NCCL), one rank per GPU.

    exchange: "peer" (default) = the slots are opened on every rank through CUDA IPC and global qubits are
    exchanged by b200_sv_peer_swap over NVLink peer memory; "nccl" = send/recv through a staging buffer
    (also the fallback when IPC is not available).  Env B200AQC_EXCHANGE overrides."""
    import os
    import torch
    import torch.distributed as dist
    from .sv_engine import SVEngine
    dev = torch.device("cuda", local_rank)
    comm = TorchComm(dev)
    g = int(np.log2(comm.world))
    nl = num_qubits - g
    exchange = os.environ.get("B200AQC_EXCHANGE", exchange or "peer")
    # engine: a pre-allocated local engine (lets the caller agree on the allocation collectively before any rank enters
    # the collectives below)
    eng = engine if engine is not None else SVEngine(nl, device=local_rank, n_slots=n_slots)
    if eng.num_qubits != nl or eng.n_slots < n_slots:
        raise ValueError("pre-allocated engine does not match the sharding")
    tensors = [torch.as_tensor(_CudaAlias(eng.device_ptr(s), 2 << nl), device=dev) for s in range(n_slots)]
    staging = None
    peer_ptrs = None
    if exchange == "peer" and comm.world > 1:
        try:
            mine = [eng.ipc_export(s) for s in range(n_slots)]
            ok = True
        except Exception:  # noqa: BLE001
            mine, ok = None, False
        table = [None] * comm.world
        dist.all_gather_object(table, mine)
        if ok and all(t is not None for t in table):
            try:
                peer_ptrs = [[None if r == comm.rank else eng.ipc_open(table[r][s]) for r in range(comm.world)]
                             for s in range(n_slots)]
            except Exception:  # noqa: BLE001
                peer_ptrs = None
        flags = [None] * comm.world
        dist.all_gather_object(flags, peer_ptrs is not None)
        if not all(flags):
            peer_ptrs = None          # every rank must take the same path
    if peer_ptrs is None:
        staging = torch.empty(min(staging_bytes // 8, (2 << nl) // comm.world), dtype=torch.float64, device=dev)

    def sync():
        eng.sync()
        torch.cuda.synchronize(dev)

    sv = ShardedStatevector(num_qubits, eng, comm, tensors, staging, sync)
    sv.peer_ptrs = peer_ptrs
    sv.exchange_mode = "peer" if peer_ptrs is not None else "nccl"
    sv._keepalive = (tensors, staging)
    return sv


def _sharded_backend_class():
    from .backends import B200SVBackend, project_sizes
    from .sv_engine import SVCostEvaluator, SVEngine

    class B200ShardedSVBackend(B200SVBackend):
        """``B200SVBackend`` for registers larger than one GPU: the statevector is sharded by global qubits over
        the ranks of the default process group (one process per GPU; every rank runs the same compiler and
        issues the same backend calls).  Cost evaluations of the blocks in the projected tail run on a
        K-qubit engine replicated per rank (one gather + all-reduce of 2^K amplitudes per projection); blocks
        outside it re-simulate the window from the resident U|0> (fused sweeps + peer-memory exchanges);
        <Z> and pair RDMs are sharded read-outs with one small all-reduce."""

        def __init__(self, local_rank=0, exchange=None):
            super().__init__(device=local_rank)
            self.exchange = exchange

        def __getstate__(self):
            return {"device": self.device, "exchange": self.exchange}

        def __setstate__(self, state):
            self.__init__(local_rank=state.get("device", 0), exchange=state.get("exchange"))

        def _make_engines(self, num_qubits):
            """(ShardedEngine, [compact-bra engines, 2 slots], [projected engines, 4 slots]) -- overridden by the CPU tests."""
            from .backends import COMPACT_QUBITS
            sv = make_gpu_sharded(num_qubits, n_slots=4, local_rank=self.device, exchange=self.exchange)
            sizes = project_sizes(num_qubits, limit=min(sv.nl, 26))     # replicated per rank: 4 slots of at most 1 GiB
            csizes = sorted({min(k, sv.nl) for k in COMPACT_QUBITS})
            return (ShardedEngine(sv), [SVEngine(k, device=self.device, n_slots=2) for k in csizes],
                    [SVEngine(k, device=self.device, n_slots=4) for k in sizes])

        def overlap_between_circuits(self, circuit1, circuit2):
            """|<psi1|psi2>|^2 = |<0| U1^+ U2 |0>|^2: both circuits on ONE sharded slot (two independently
            simulated slots would end up in different qubit layouts), then amplitude 0."""
            eng = self._get_engine(circuit1.num_qubits)
            from .sv_engine import SLOT_WORK
            eng.run(SLOT_WORK, -1, G.GateStream.from_circuit(circuit2))
            eng.run(SLOT_WORK, SLOT_WORK, G.GateStream.from_circuit(circuit1), inverse=True)
            self._state_version += 1
            return np.absolute(eng.amp(SLOT_WORK, 0)) ** 2

        def _get_engine(self, num_qubits):
            if self._engine is None or self._engine.num_qubits != num_qubits:
                for e in [self._engine] + list(getattr(self, "_projected", None) or []) + list(self._compact or []):
                    if e is not None:
                        e.close()
                self._engine, self._compact, self._projected = self._make_engines(num_qubits)
                self._evaluator = SVCostEvaluator(self._engine, self._compact, self._projected)
                self._evaluator.dense_blocks = False
                self._state_version += 1
                self._last_run_key = None
                self._last_run_insts = None
            return self._engine

    return B200ShardedSVBackend


def __getattr__(name):          # B200ShardedSVBackend is built on first use (keeps `import dist_sv` light)
    if name == "B200ShardedSVBackend":
        cls = _sharded_backend_class()
        globals()[name] = cls
        return cls
    raise AttributeError(name)
