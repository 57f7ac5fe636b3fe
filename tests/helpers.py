"""Shared test helpers: random circuits, the benchmark circuit families, emulator driver."""
import ctypes

import numpy as np

from harness.circuit import Circuit, Gate
from harness.workloads import brickwall_pairs, brickwork, random_vidal_mps, thin_ansatz  # noqa: F401
from adapt_aqc_b200.gates import GateStream


def random_unitary(dim, rng):
    a = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def random_gates(n, ng, rng, allow_mat=True):
    """[(name, qubits, params)] over every opcode the C-ABI knows."""
    names1 = ["x", "y", "z", "h", "s", "sdg", "t", "tdg", "sx", "rx", "ry", "rz", "u1", "u2", "u3", "id"]
    names2 = ["cx", "cz", "swap"]
    if allow_mat:
        names1.append("mat1")
        names2.append("mat2")
    npar = {"rx": 1, "ry": 1, "rz": 1, "u1": 1, "u2": 2, "u3": 3}
    out = []
    for _ in range(ng):
        if n < 2 or rng.random() < 0.55:
            nm = names1[rng.integers(len(names1))]
            q = [int(rng.integers(n))]
            if nm == "mat1":
                out.append((nm, q, random_unitary(2, rng)))
            else:
                out.append((nm, q, list(rng.uniform(-np.pi, np.pi, npar.get(nm, 0)))))
        else:
            nm = names2[rng.integers(len(names2))]
            q = [int(x) for x in rng.choice(n, 2, replace=False)]
            out.append((nm, q, random_unitary(4, rng) if nm == "mat2" else []))
    return out


def circuit_from_gates(n, gates):
    c = Circuit(n)
    for name, qubits, params in gates:
        if name in ("mat1", "mat2"):
            c.unitary(params, qubits)
        else:
            c.append(Gate(name, params), qubits)
    return c


def golden_seeds():
    """Seeds of every committed paper/random_mps fixture (all 54 of the reference's targets)."""
    import glob
    import os
    import re
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return sorted(int(re.search(r"seed_(\d+)", f).group(1)) for f in glob.glob(os.path.join(here, "random_mps_seed_*.npz")))


def load_golden_mps(seed):
    """tests/golden/random_mps_seed_<seed>.npz -> QiskitMPS tuple (constants.py:17)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"random_mps_seed_{seed}.npz"))
    n = sum(1 for k in z.files if k.startswith("g"))
    gammas = [(z[f"g{i}"][0], z[f"g{i}"][1]) for i in range(n)]
    lambdas = [z[f"l{i}"] for i in range(n - 1)]
    return (gammas, lambdas)


def emu_run(emu, n, gates, psi0=None, inverse=False):
    gs = gates if isinstance(gates, GateStream) else GateStream.from_gates(gates)
    outs = []
    for variant in (0, 1):      # direct-load kernel body, pipelined (bulk-copy) kernel body
        st = np.zeros(1 << n, dtype=np.complex128) if psi0 is None else np.array(psi0, dtype=np.complex128)
        stats = (ctypes.c_int32 * 4)()
        emu.emu_set_variant(variant)
        rc = emu.emu_sv_run(n, st.view(np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                            1 if psi0 is None else 0, gs.rec_ptr(), len(gs), gs.mats_ptr(), len(gs.mats),
                            int(inverse), stats)
        assert rc == 0, emu.emu_last_error()
        outs.append(st)
    # Same arithmetic up to where pending phases are flushed: the direct body folds X-type ops on lane qubits into its
    # addressing, the pipelined body executes them as shuffles (which flush the pending phase first), so a product of two
    # phases may be rounded once instead of twice -- a few ulp, never more.
    assert np.allclose(outs[0], outs[1], rtol=0, atol=1e-14), "direct and pipelined sweep bodies disagree"
    return st, tuple(stats)


def emu_run_inner2(emu, n, gates, psi0, other, qa, qb, inverse=False, write_back=True):
    """CPU execution of the fused sweep + transfer-pass kernel bodies: (swept state, T[i][j] with index = bit(qa) + 2 bit(qb),
    planner stats)."""
    gs = gates if isinstance(gates, GateStream) else GateStream.from_gates(gates)
    st = np.array(psi0, dtype=np.complex128)
    oth = np.ascontiguousarray(other, dtype=np.complex128)
    t = np.zeros(32)
    stats = (ctypes.c_int32 * 4)()
    dp = ctypes.POINTER(ctypes.c_double)
    rc = emu.emu_sv_run_inner2(n, st.view(np.float64).ctypes.data_as(dp), oth.view(np.float64).ctypes.data_as(dp),
                               gs.rec_ptr(), len(gs), gs.mats_ptr(), len(gs.mats), int(inverse), int(qa), int(qb),
                               1 if write_back else 0, t.ctypes.data_as(dp), stats)
    assert rc == 0, emu.emu_last_error()
    T = t.view(np.complex128).reshape(4, 4)
    if qa > qb:     # kernel index: bit 0 = the lower qubit
        sw = [0, 2, 1, 3]
        T = T[np.ix_(sw, sw)]
    return st, T.copy(), tuple(stats)


def emu_run_embedded(emu, n, gates, phi, qmap, inverse=False, fuse=None, other=None):
    """CPU execution of the embedded-source sweep (and of its fused variant): (state, T or None).  Falls back to the
    product's scatter + run order when there is no tiled sweep to ride on (small register, empty program)."""
    gs = gates if isinstance(gates, GateStream) else GateStream.from_gates(gates)
    qm = np.ascontiguousarray(np.asarray(qmap, dtype=np.int32))
    phi = np.ascontiguousarray(phi, dtype=np.complex128)
    dp = ctypes.POINTER(ctypes.c_double)
    if n <= 11 or (len(gs) == 0 and fuse is None):
        c = np.arange(len(phi))
        x = np.zeros_like(c)
        for b, q in enumerate(qm):
            x |= ((c >> b) & 1) << int(q)
        psi = np.zeros(1 << n, dtype=np.complex128); psi[x] = phi
        out, _ = emu_run(emu, n, gs, psi0=psi, inverse=inverse)
        return out, None
    st = np.zeros(1 << n, dtype=np.complex128)
    t = np.zeros(32)
    stats = (ctypes.c_int32 * 4)()
    oth = np.ascontiguousarray(other, dtype=np.complex128) if fuse is not None else np.zeros(1, dtype=np.complex128)
    qa, qb = (int(fuse[0]), int(fuse[1])) if fuse is not None else (-1, -1)
    rc = emu.emu_sv_run_embedded(n, st.view(np.float64).ctypes.data_as(dp), phi.view(np.float64).ctypes.data_as(dp), len(qm),
                                 qm.ctypes.data, gs.rec_ptr(), len(gs), gs.mats_ptr(), len(gs.mats), int(inverse), qa, qb,
                                 oth.view(np.float64).ctypes.data_as(dp), t.ctypes.data_as(dp), stats)
    assert rc == 0, emu.emu_last_error()
    emu_run_embedded.last_sweeps = int(stats[0])
    if fuse is None:
        return st, None
    T = t.view(np.complex128).reshape(4, 4)
    if qa > qb:
        sw = [0, 2, 1, 3]
        T = T[np.ix_(sw, sw)]
    return st, T.copy()


def emu_run_project(emu, n, gates, psi0, qmap, inverse=False):
    """CPU execution of the projected-store sweep: (phi, number of sweeps)."""
    gs = gates if isinstance(gates, GateStream) else GateStream.from_gates(gates)
    qm = np.ascontiguousarray(np.asarray(qmap, dtype=np.int32))
    st = np.array(psi0, dtype=np.complex128)
    phi = np.full(1 << len(qm), np.nan + 0j, dtype=np.complex128)        # every entry must be written
    stats = (ctypes.c_int32 * 4)()
    dp = ctypes.POINTER(ctypes.c_double)
    rc = emu.emu_sv_run_project(n, st.view(np.float64).ctypes.data_as(dp), phi.view(np.float64).ctypes.data_as(dp), len(qm),
                                qm.ctypes.data, gs.rec_ptr(), len(gs), gs.mats_ptr(), len(gs.mats), int(inverse), stats)
    assert rc == 0, emu.emu_last_error()
    return phi, int(stats[0])


class FakeEngine:
    """CPU stand-in for SVEngine used by the `not gpu` host-logic tests: gate application goes
    through the product planner + kernel thread bodies (tests/emu), read-outs through the oracle.
    It exists only to exercise the evaluator / backend / compile-loop logic without a GPU."""

    def __init__(self, emu, num_qubits, n_slots=4):
        self.emu = emu
        self.num_qubits = num_qubits
        self.slots = [np.zeros(1 << num_qubits, dtype=np.complex128) for _ in range(n_slots)]
        self.runs = 0
        self.inners = 0

    def close(self):
        pass

    def run(self, dst, src, stream, inverse=False):
        psi0 = None if src < 0 else self.slots[src]
        out, _ = emu_run(self.emu, self.num_qubits, stream, psi0=psi0, inverse=inverse)
        self.slots[dst][...] = out          # in place: the sharded tests alias slot memory with torch tensors
        self.runs += 1

    FUSED_MIN_QUBITS = 12

    def run_inner2(self, dst, src, stream, other, qa, qb, inverse=False, store=True):
        if src < 0:                                # |0..0> source: always stored
            zero = np.zeros(1 << self.num_qubits, dtype=np.complex128); zero[0] = 1
            out, T, _ = emu_run_inner2(self.emu, self.num_qubits, stream, zero, self.slots[other], qa, qb, inverse=inverse)
            self.slots[dst][...] = out
            self.runs += 1
            self.inners += 1
            return T if store else (T, True)
        out, T, stats = emu_run_inner2(self.emu, self.num_qubits, stream, self.slots[src], self.slots[other], qa, qb, inverse=inverse)
        self.runs += 1
        self.inners += 1
        stored = store or stats[0] != 1          # b200_sv_run_inner2: T only when one sweep carries the program
        if stored:
            self.slots[dst][...] = out
        else:
            keep, T2, _ = emu_run_inner2(self.emu, self.num_qubits, stream, self.slots[src], self.slots[other], qa, qb,
                                         inverse=inverse, write_back=False)
            assert np.array_equal(keep, self.slots[src]) and np.array_equal(T2, T)
        return T if store else (T, stored)

    def run_project(self, scratch, src, stream, qmap, dst_engine, dst_slot, inverse=False):
        phi, sweeps = emu_run_project(self.emu, self.num_qubits, stream, self.slots[src], qmap, inverse=inverse)
        dst_engine.slots[dst_slot][...] = phi
        self.runs += 1
        if sweeps > 1:
            self.slots[scratch][...] = np.nan      # the product leaves an intermediate state there: nobody may rely on it
        return sweeps > 1

    def run_embedded(self, dst, qmap, src_engine, src_slot, stream, inverse=False, fuse=None, store=True):
        other = self.slots[fuse[0]] if fuse is not None else None
        out, T = emu_run_embedded(self.emu, self.num_qubits, stream, src_engine.slots[src_slot], qmap, inverse=inverse,
                                  fuse=None if fuse is None else (fuse[1], fuse[2]), other=other)
        self.runs += 1
        stored = store or fuse is None or emu_run_embedded.last_sweeps != 1     # b200_sv_run_embedded_inner2's `stored`
        if stored:
            self.slots[dst][...] = out
        return T if store else (T, stored)

    def copy(self, dst, src):
        self.slots[dst][...] = self.slots[src]

    def sync(self):
        pass

    def amp(self, slot, index=0):
        return complex(self.slots[slot][index])

    def expz(self, slot):
        from oracle import sv_oracle as orc
        psi = self.slots[slot]
        return np.array(orc.measure_qubit_expectation_values(psi)), float(np.vdot(psi, psi).real)

    def pair_rdm(self, slot, pairs, part=0, n_parts=1):
        """part / n_parts: emulates b200_sv_pair_rdm_part with one "pass" per pair -- pairs of other parts are zeros."""
        from oracle import sv_oracle as orc
        return np.array([orc.partial_trace(self.slots[slot], a, b) if i % n_parts == part else np.zeros((4, 4), dtype=np.complex128)
                         for i, (a, b) in enumerate(pairs)]).reshape(-1, 4, 4)

    def inner(self, l_slot, r_slot, q=-1):
        self.inners += 1
        L, R = self.slots[l_slot], self.slots[r_slot]
        if q < 0:
            return complex(np.vdot(L, R))
        n = self.num_qubits
        Lt = np.moveaxis(L.reshape([2] * n), n - 1 - q, 0).reshape(2, -1)
        Rt = np.moveaxis(R.reshape([2] * n), n - 1 - q, 0).reshape(2, -1)
        return Lt.conj() @ Rt.T

    def inner2(self, l_slot, r_slot, qa, qb):
        self.inners += 1
        n = self.num_qubits
        L, R = self.slots[l_slot].reshape([2] * n), self.slots[r_slot].reshape([2] * n)
        Lt = np.moveaxis(L, [n - 1 - qb, n - 1 - qa], [0, 1]).reshape(4, -1)   # index = 2*bit(qb) + bit(qa)
        Rt = np.moveaxis(R, [n - 1 - qb, n - 1 - qa], [0, 1]).reshape(4, -1)
        return Lt.conj() @ Rt.T

    def inner2_gather(self, r_slot, compact_engine, compact_slot, qmap, qa, qb):
        """numpy restatement of b200_sv_inner2_gather: scatter the compact bra into the full register."""
        self.inners += 1
        n, K = self.num_qubits, compact_engine.num_qubits
        ell = compact_engine.slots[compact_slot]
        c = np.arange(1 << K)
        x = np.zeros_like(c)
        for b, q in enumerate(qmap):
            x |= ((c >> b) & 1) << q
        L = np.zeros(1 << n, dtype=np.complex128)
        L[x] = ell
        saved = self.slots[0].copy()
        self.slots[0][...] = L
        out = self.inner2(0, r_slot, qa, qb) if r_slot != 0 else None
        self.slots[0][...] = saved
        self.inners -= 1
        return out

    def gather(self, slot, qmap, dst_engine, dst_slot):
        """numpy restatement of b200_sv_gather."""
        c = np.arange(1 << len(qmap))
        x = np.zeros_like(c)
        for b, q in enumerate(qmap):
            x |= ((c >> b) & 1) << q
        dst_engine.slots[dst_slot][...] = self.slots[slot][x]

    def scatter(self, slot, qmap, src_engine, src_slot):
        """numpy restatement of b200_sv_scatter."""
        c = np.arange(1 << len(qmap))
        x = np.zeros_like(c)
        for b, q in enumerate(qmap):
            x |= ((c >> b) & 1) << q
        self.slots[slot][...] = 0
        self.slots[slot][x] = src_engine.slots[src_slot]

    def gather_ranked(self, slot, qmap, rank_bits, rank, dst_engine, dst_slot):
        """numpy restatement of b200_sv_gather_ranked."""
        n = self.num_qubits
        c = np.arange(1 << len(qmap))
        x = np.zeros_like(c)
        rb = np.zeros_like(c)
        covered = 0
        for b, q in enumerate(qmap):
            if q < n:
                x |= ((c >> b) & 1) << q
            else:
                rb |= ((c >> b) & 1) << (q - n)
                covered |= 1 << (q - n)
        ok = (rb == (rank & covered)) & ((rank & ~covered) == 0)
        dst_engine.slots[dst_slot][...] = np.where(ok, self.slots[slot][x], 0)

    def device_ptr(self, slot):
        return 0

    def download(self, slot, offset=0, count=None):
        return self.slots[slot][offset:None if count is None else offset + count].copy()

    def upload(self, slot, host, offset=0):
        self.slots[slot][offset:offset + len(host)] = host


def compile_option_cases():
    """(name, target, AdaptCompiler kwargs, AdaptConfig kwargs): the reference's compile options that
    change what the backend is asked for (test/recompilers/test_adapt_compiler.py: local cost :142-152,
    custom layer gate :815-818, starting circuit :795-802, initial single-qubit layer, pair-selection
    methods :206-237, brickwall :1465-1543, rotosolve options)."""
    rng = np.random.default_rng(21)
    n = 4
    target = Circuit(n)
    for layer in range(2):
        for q in range(n):
            target.u3(*rng.uniform(-np.pi, np.pi, 3), q)
        for q in range(layer % 2, n - 1, 2):
            target.cx(q, q + 1)
    start = Circuit(n); start.x(0); start.h(2)
    cz_layer = Circuit(2)
    cz_layer.ry(0, 0, label="ry"); cz_layer.ry(0, 1, label="ry"); cz_layer.cz(0, 1)
    cz_layer.ry(0, 0, label="ry"); cz_layer.ry(0, 1, label="ry")
    return [
        ("default", target, {}, dict(max_layers=4)),
        ("local_cost", target, dict(optimise_local_cost=True), dict(max_layers=3)),
        ("starting_circuit", target, dict(starting_circuit=start), dict(max_layers=3)),
        ("isql", target, dict(initial_single_qubit_layer=True), dict(max_layers=3)),
        ("custom_cz_layer", target, dict(custom_layer_2q_gate=cz_layer), dict(max_layers=3)),
        ("no_rotoselect", target, dict(use_rotoselect=False), dict(max_layers=3)),
        ("expectation", target, {}, dict(max_layers=3, method="expectation")),
        ("basic", target, {}, dict(max_layers=3, method="basic")),
        ("brickwall", target, {}, dict(max_layers=4, method="brickwall")),
        ("rotosolve_every_2", target, {}, dict(max_layers=4, rotosolve_frequency=2, max_layers_to_modify=2)),
        ("reuse_qubit_mode", target, {}, dict(max_layers=4, reuse_exponent=1, reuse_priority_mode="qubit")),
    ]
