"""B200 MPS backend behind ADAPT-AQC's backend interface.

``B200MPSBackend`` is the drop-in for ``AerMPSBackend`` (adaptaqc/backends/aer_mps_backend.py:45-93):
same four methods + ``evaluate_hamming_weight_one_overlaps``, same argument (the compiler), same
formulas.  ``backend.simulator`` is an ``AerSimulator(method="matrix_product_state")``-shaped facade
(``.options.matrix_product_state_truncation_threshold`` is read by the reference at
adaptaqc/compilers/approximate_compiler.py:223-226).  ``backend.mps_ops`` mirrors the
``aqc_research.mps_operations`` functions the reference imports (aer_mps_backend.py:14-19,
entanglement_measures.py:16,77, adapt_compiler.py:19,1129, gradients.py:14) with the same
signatures, operating on device-resident states: an MPS never travels to the host unless host code
indexes into it.

Everything is evaluated on the GPU through libb200aqc.so; there is no CPU path.
"""
import numpy as np

from . import gates as G
from . import backends as _sv
from .mps_engine import SLOT_BASE, SLOT_L, SLOT_R, SLOT_WORK, DeviceMPS, MPSContext, MPSEngine
from .sv_engine import SVCostEvaluator

try:  # pragma: no cover - the reference needs qiskit + qiskit-aer + aqc_research
    from adaptaqc.backends.aer_mps_backend import AerMPSBackend as _MPSBase
except Exception:  # noqa: BLE001
    _MPSBase = _sv._SVBase if not _sv.HAVE_REFERENCE else object

MPS_INSTRUCTIONS = ("set_matrix_product_state", "save_matrix_product_state")


class DeviceMPSView:
    """A *preprocessed* MPS (what ``mps_from_circuit(..., return_preprocessed=True)`` returns) that
    lives in HBM.  List-like for host code that insists on indexing (downloads once, lazily)."""

    def __init__(self, sim, handle):
        self._sim = sim
        self.handle = handle
        self.num_qubits = handle.num_qubits
        self._host = None
        self._rdm = {}
        self._expz = None
        self._pair_hint = None

    def __len__(self):
        return self.num_qubits

    def _materialise(self):
        if self._host is None:
            gammas, lambdas = self.handle.get()
            out = []
            for i, (a0, a1) in enumerate(gammas):
                g = np.stack([a0, a1])
                if i < self.num_qubits - 1:
                    g = g * lambdas[i].reshape(1, 1, -1)
                out.append(g)
            self._host = out
        return self._host

    def __getitem__(self, i):
        return self._materialise()[i]

    def __iter__(self):
        return iter(self._materialise())

    def __del__(self):
        try:
            self._sim._recycle(self.handle)
        except Exception:  # noqa: BLE001
            pass


class _Op:
    def __init__(self, name, params):
        self.name, self.params, self.label = name, params, name


class _Inst:
    def __init__(self, operation, qubits):
        self.operation, self.qubits, self.clbits = operation, qubits, ()


class _CircuitView:
    """full_circuit with ONE instruction replaced (a shift candidate), without copying or touching the circuit."""

    def __init__(self, circuit, data, index, inst):
        self.num_qubits = circuit.num_qubits
        self.qubits = circuit.qubits
        self.data = data[:index] + [inst] + data[index + 1:]


class _Options:
    def __init__(self, thr, max_chi):
        self.matrix_product_state_truncation_threshold = thr
        self.matrix_product_state_max_bond_dimension = max_chi


class _MPSResult:
    def __init__(self, data):
        self._data = data

    def data(self, _experiment=0):
        return self._data


class _MPSJob:
    def __init__(self, data):
        self._data = data

    def result(self):
        return _MPSResult(self._data)


class B200MPSSimulator:
    """``mps_sim_with_args`` replacement (aer_mps_backend.py:27-42)."""

    name = "b200_matrix_product_state"

    def __init__(self, mps_truncation_threshold=1e-16, max_chi=None, mps_log_data=False, device=0):
        self.options = _Options(mps_truncation_threshold, max_chi)
        self.device = device
        self._context = None
        self._pool = {}            # num_qubits -> [DeviceMPS]
        self._uploaded = {}        # id(host mps object) -> (host object, DeviceMPS): set_mps cache
        self.runs = 0
        # prefix checkpoints of the most recent run: states right after each 2-qubit gate (= each SVD)
        self._ckpt_key = None      # (n, truncation settings)
        self._ckpt_base = None     # the host target MPS object the checkpoints start from (None: |0..0>)
        self._ckpt_window = []     # canonical window of the run that produced them
        self._ckpts = []           # [(prefix length, DeviceMPS)], ascending
        self.ckpt_stats = {"resumed_gates": 0, "applied_gates": 0, "svds_skipped": 0}

    def __getstate__(self):
        return {"thr": self.options.matrix_product_state_truncation_threshold,
                "max_chi": self.options.matrix_product_state_max_bond_dimension, "device": self.device}

    def __setstate__(self, st):
        self.__init__(st["thr"], st["max_chi"], device=st.get("device", 0))

    @property
    def ops(self):
        """The aqc_research.mps_operations-shaped functions bound to this simulator (registration.install_mps)."""
        if getattr(self, "_ops", None) is None:
            self._ops = MPSOps(self)
        return self._ops

    # ---- handles ----
    def context(self):
        if self._context is None:
            self._context = MPSContext(self.device)
        return self._context

    def _acquire(self, n):
        pool = self._pool.setdefault(n, [])
        m = pool.pop() if pool else self.context().new_mps(n)
        m.set_truncation(self.options.matrix_product_state_truncation_threshold,
                         self.options.matrix_product_state_max_bond_dimension)
        return m

    def _recycle(self, handle):
        if handle is not None and handle._h.value and len(self._pool.setdefault(handle.num_qubits, [])) < 8:
            self._pool[handle.num_qubits].append(handle)
        elif handle is not None:
            handle.close()

    def device_copy_of(self, host_mps):
        """Device copy of a host QiskitMPS object, cached on object identity (the target MPS sits
        inside the set_matrix_product_state instruction and is re-used on every evaluation)."""
        ent = self._uploaded.get(id(host_mps))
        if ent is not None and ent[0] is host_mps:
            return ent[1]
        m = self.context().new_mps(len(host_mps[0]))
        m.set(host_mps)
        if len(self._uploaded) > 16:
            _, old = self._uploaded.pop(next(iter(self._uploaded)))
            old.close()
        self._uploaded[id(host_mps)] = (host_mps, m)
        return m

    # ---- simulation ----
    def simulate(self, circuit, use_checkpoints=True):
        """Run ``[set_matrix_product_state?] gates... [save_matrix_product_state?]`` and return a
        DeviceMPS handle owned by the caller.  use_checkpoints=False: a side computation (e.g. the starting state of
        the gradient heuristic) that must not evict the prefix checkpoints of the circuit being optimised."""
        self.runs += 1
        n = circuit.num_qubits
        out = self._acquire(n)
        data = circuit.data
        start = 0
        if len(data) and data[0].operation.name == "set_matrix_product_state":
            out.copy_from(self.device_copy_of(data[0].operation.params[0]))
            start = 1
        else:
            out.init_zero()
        stop = len(data)
        while stop > start and data[stop - 1].operation.name == "save_matrix_product_state":
            stop -= 1
        window = G.canonical_window(circuit, start, stop)
        if window and use_checkpoints:
            self._apply_with_checkpoints(out, window, data[0].operation.params[0] if start else None)
        elif window:
            out.apply(G.GateStream.from_window(window))
        return out

    MAX_CHECKPOINTS = 24
    _DIAGONAL_1Q = ("rz", "z", "s", "sdg", "t", "tdg", "u1", "p", "id", "i")

    @classmethod
    def _commute_phases_past_controls(cls, window):
        """Normal form for checkpoint matching: every 1-qubit gate is moved as LATE as exact commutation allows
        -- past gates on other qubits, and, if it is diagonal, past a cx it controls or a cz it takes part in.
        (One right-to-left pass; gates on the same qubit never overtake each other.)  A run that only changed
        such a rotation then resumes from a checkpoint behind the 2-qubit gates it commutes with: no new SVD.
        The circuit is unchanged (exact commutations); moving a unitary diagonal across an SVD leaves the
        singular values and the kept subspace untouched."""
        out = list(window)
        if len(out) > 512:
            return out
        for j in range(len(out) - 2, -1, -1):
            g = out[j]
            if g[2] >= 0:
                continue
            q = g[1]
            diag = g[0] in cls._DIAGONAL_1Q and g[6] is None
            k = j
            while k + 1 < len(out):
                h = out[k + 1]
                if h[2] < 0:
                    ok = h[1] != q
                elif q not in (h[1], h[2]):
                    ok = True
                else:
                    ok = diag and ((h[0] == "cx" and h[1] == q) or h[0] == "cz")
                if not ok:
                    break
                out[k], out[k + 1] = h, g
                k += 1
        return out

    def _drop_checkpoints(self, keep=0):
        while len(self._ckpts) > keep:
            self._recycle(self._ckpts.pop()[1])

    def _apply_with_checkpoints(self, out, window, base_obj):
        """Apply `window` to `out` (which holds the base state), resuming from the longest prefix the
        previous run shares with it.  The reference re-applies every un-absorbed gate on every evaluation
        (aer_mps_backend.py:76-78) although the optimiser changed ONE rotation: the state right after
        each 2-qubit gate (each SVD + truncation) of the previous run is kept on the device, and a run
        whose gate list starts with the same gates resumes from it.  Same gates in the same order with the
        same truncation rule give the same state, so the result is identical to a full re-run."""
        o = self.options
        window = self._commute_phases_past_controls(window)
        key = (out.num_qubits, o.matrix_product_state_truncation_threshold, o.matrix_product_state_max_bond_dimension)
        if key != self._ckpt_key or base_obj is not self._ckpt_base:      # (the host target object is kept alive here)
            self._drop_checkpoints()
            self._ckpt_key, self._ckpt_base, self._ckpt_window = key, base_obj, []
        old = self._ckpt_window
        common = 0
        while common < len(old) and common < len(window) and old[common] == window[common]:
            common += 1
        self._drop_checkpoints(sum(1 for plen, _ in self._ckpts if plen <= common))
        pos = 0
        if self._ckpts:
            pos, h = self._ckpts[-1]
            out.copy_from(h)
            self.ckpt_stats["resumed_gates"] += pos
            self.ckpt_stats["svds_skipped"] += len(self._ckpts)
        two_q = [i for i in range(pos, len(window)) if window[i][2] >= 0]
        for i in two_q:
            out.apply(G.GateStream.from_window(window[pos:i + 1]))
            self.ckpt_stats["applied_gates"] += i + 1 - pos
            pos = i + 1
            if len(self._ckpts) < self.MAX_CHECKPOINTS:
                h = self._acquire(out.num_qubits)
                h.copy_from(out)
                self._ckpts.append((pos, h))
        if pos < len(window):
            out.apply(G.GateStream.from_window(window[pos:]))
            self.ckpt_stats["applied_gates"] += len(window) - pos
        self._ckpt_window = list(window)

    def run(self, circuit, **_options):
        """Aer-shaped entry: ``sim.run(qc, shots=1).result().data(0)[label]`` gives the QiskitMPS."""
        handle = self.simulate(circuit)
        mps = handle.get()
        self._recycle(handle)
        label = "matrix_product_state"
        for inst in circuit.data:
            if inst.operation.name == "save_matrix_product_state":
                label = getattr(inst.operation, "label", None) or getattr(inst.operation, "_label", None) or label
        return _MPSJob({label: mps, "matrix_product_state": mps, "my_mps": mps})


def mps_sim_with_args(mps_truncation_threshold=1e-16, max_chi=None, mps_log_data=False, device=0):
    return B200MPSSimulator(mps_truncation_threshold, max_chi, mps_log_data, device)


class MPSOps:
    """``aqc_research.mps_operations`` with the signatures the reference uses, on device states."""

    def __init__(self, default_sim):
        self._default_sim = default_sim
        self._sims_by_threshold = {}      # trunc_thr -> B200MPSSimulator sharing the default simulator's device context

    # -- conversions --
    @staticmethod
    def check_mps(x):
        return (isinstance(x, tuple) and len(x) == 2 and isinstance(x[0], list) and isinstance(x[1], list)
                and len(x[0]) == len(x[1]) + 1)

    @staticmethod
    def _preprocess_mps(mps):
        gammas, lambdas = mps
        n = len(gammas)
        out = []
        for i, (a0, a1) in enumerate(gammas):
            g = np.stack([np.asarray(a0, dtype=np.complex128), np.asarray(a1, dtype=np.complex128)])
            if i < n - 1:
                g = g * np.asarray(lambdas[i], dtype=np.float64).reshape(1, 1, -1)
            out.append(g)
        return out

    def _device(self, mps, already_preprocessed, sim=None):
        """(handle, temporary?)"""
        sim = sim or self._default_sim
        if isinstance(mps, DeviceMPSView):
            return mps.handle, False
        if isinstance(mps, DeviceMPS):
            return mps, False
        if self.check_mps(mps):
            h = sim._acquire(len(mps[0]))
            h.set(mps)
            return h, True
        h = sim._acquire(len(mps))
        h.set_preprocessed(list(mps))
        return h, True

    # -- the functions --
    def mps_from_circuit(self, qc, trunc_thr=1e-16, print_log_data=False, return_preprocessed=False, sim=None,
                         out_state=None):
        """Appends the save instruction to `qc` in place like aqc_research does (callers pass
        copies, aer_mps_backend.py:77), runs the circuit on the device."""
        if sim is None:
            sim = self._default_sim
            if trunc_thr != sim.options.matrix_product_state_truncation_threshold:
                # one cached simulator per threshold, on the default simulator's context (no new device context per call)
                sim = self._sims_by_threshold.get(trunc_thr)
                if sim is None:
                    sim = self._sims_by_threshold[trunc_thr] = B200MPSSimulator(trunc_thr, device=self._default_sim.device)
                    sim._context = self._default_sim.context()
        if hasattr(qc, "save_matrix_product_state"):
            qc.save_matrix_product_state()
        handle = sim.simulate(qc)
        if return_preprocessed:
            return DeviceMPSView(sim, handle)
        mps = handle.get()
        sim._recycle(handle)
        return mps

    def mps_dot(self, mps1, mps2, already_preprocessed=False):
        a, ta = self._device(mps1, already_preprocessed)
        b, tb = self._device(mps2, already_preprocessed)
        val = a.dot(b)
        if ta:
            self._default_sim._recycle(a)
        if tb:
            self._default_sim._recycle(b)
        return val

    def mps_expectation(self, mps, pauli, qubit, already_preprocessed=False):
        if pauli != "Z":
            raise NotImplementedError("only Pauli Z expectations are on the reference's hot path")
        if isinstance(mps, DeviceMPSView):
            if mps._expz is None:
                mps._expz = mps.handle.expz()     # all qubits from one pair of sweeps
            return float(mps._expz[0][qubit])
        h, tmp = self._device(mps, already_preprocessed)
        z, _ = h.expz()
        if tmp:
            self._default_sim._recycle(h)
        return float(z[qubit])

    def extract_amplitude(self, mps, bitstring, already_preprocessed=False):
        h, tmp = self._device(mps, already_preprocessed)
        val = complex(h.amps([int(bitstring)])[0])
        if tmp:
            self._default_sim._recycle(h)
        return val

    def partial_trace(self, mps, qubits, already_preprocessed=False, pair_hint=None):
        key = tuple(sorted(int(q) for q in qubits))
        if isinstance(mps, DeviceMPSView):
            if key not in mps._rdm:
                need = [key]
                pair_hint = pair_hint or mps._pair_hint
                if pair_hint:
                    need = sorted({tuple(sorted(p)) for p in pair_hint} | {key})
                    need = [p for p in need if p not in mps._rdm]
                for p, r in zip(need, mps.handle.pair_rdm(need)):
                    mps._rdm[p] = r
            return mps._rdm[key]
        h, tmp = self._device(mps, already_preprocessed)
        rho = h.pair_rdm([key])[0]
        if tmp:
            self._default_sim._recycle(h)
        return rho

    def mps_to_vector(self, mps, already_preprocessed=False):
        pp = mps if already_preprocessed else (self._preprocess_mps(mps) if self.check_mps(mps) else mps)
        v = np.ones((1, 1), dtype=np.complex128)
        for g in pp:
            v = np.einsum("kx,sxy->sky", v, g).reshape(-1, g.shape[2])
        return v[:, 0]


class MPSCostEvaluator(SVCostEvaluator):
    """Block transfer-matrix evaluator (sv_engine.SVCostEvaluator) on MPS slots: the base state is a
    loaded MPS instead of a simulated prefix."""

    def set_base_handle(self, key, handle):
        if key is not None and key == self.base_key:
            return
        self.eng.slots[SLOT_BASE].copy_from(handle)
        self.base_key = key
        self.invalidate()


class B200MPSBackend(_MPSBase):
    kind = "mps"  # what isinstance(backend, AerMPSBackend) decides in the reference

    def __init__(self, simulator=None, device=0, incremental=True):
        self.simulator = simulator if simulator is not None else B200MPSSimulator(device=device)
        self.mps_ops = self.simulator.ops
        self.incremental = incremental
        self._engine = None
        self._evaluator = None

    def __getstate__(self):          # (the worker simulators and their thread pool are rebuilt on demand)
        return {"simulator": self.simulator, "incremental": self.incremental}

    def __setstate__(self, st):
        self.__init__(simulator=st["simulator"], incremental=st.get("incremental", True))

    # ---- incremental evaluation machinery ----
    def _use_incremental(self, compiler):
        """The transfer-matrix shortcut re-orders the contraction, which is exact only up to the
        truncation error; use it when truncation is at roundoff level (the reference's default
        1e-16, aer_mps_backend.py:27), otherwise reproduce the reference's contraction order."""
        o = self.simulator.options
        return (self.incremental and not compiler.soften_global_cost
                and o.matrix_product_state_truncation_threshold <= 1e-12
                and o.matrix_product_state_max_bond_dimension is None)

    def _get_evaluator(self, n):
        o = self.simulator.options
        if self._engine is None or self._engine.num_qubits != n:
            if self._engine is not None:
                self._engine.close()
            self._engine = MPSEngine(n, context=self.simulator.context(),
                                     truncation_threshold=o.matrix_product_state_truncation_threshold,
                                     max_bond_dimension=o.matrix_product_state_max_bond_dimension)
            self._evaluator = MPSCostEvaluator(self._engine)
        self._engine.set_truncation(o.matrix_product_state_truncation_threshold, o.matrix_product_state_max_bond_dimension)
        return self._evaluator

    def _base_and_window(self, compiler):
        """(device base state | None, cache key, canonical window, circuit index of the window's first gate)"""
        circuit = compiler.full_circuit
        data = circuit.data
        if len(data) and data[0].operation.name == "set_matrix_product_state":
            host = data[0].operation.params[0]
            base = self.simulator.device_copy_of(host)
            key = ("mps", id(host))
            window, start = G.canonical_window(circuit, 1, None), 1
        else:   # circuit target that was not converted (not produced by the reference's prepare_circuit)
            base, key, window, start = None, ("zero", circuit.num_qubits), G.canonical_window(circuit, 0, None), 0
        return base, key, window, start

    def amp0(self, compiler, with_start=False):
        ev = self._get_evaluator(compiler.full_circuit.num_qubits)
        base, key, window, start = self._base_and_window(compiler)
        if base is not None:
            ev.set_base_handle(key, base)
        elif ev.base_key != key:
            self._engine.slots[SLOT_BASE].init_zero()
            ev.base_key = key
            ev.invalidate()
        return (ev, window, start) if with_start else (ev, window)

    # ---- the backend methods (aer_mps_backend.py:49-93) ----
    def evaluate_global_cost(self, compiler):
        if self._use_incremental(compiler):
            ev, window = self.amp0(compiler)
            amp = ev.amp0(window, focus=len(window) - compiler.rhs_gate_count - 1)
            return 1 - np.absolute(amp) ** 2
        circ_mps = self.evaluate_circuit(compiler)
        n = compiler.full_circuit.num_qubits
        if not compiler.soften_global_cost:
            return 1 - np.absolute(circ_mps.handle.amps([0])[0]) ** 2
        # <0|psi> and the n Hamming-weight-one amplitudes from one launch
        amps = circ_mps.handle.amps([0] + [1 << i for i in range(n)])
        global_cost = 1 - np.absolute(amps[0]) ** 2
        previous_cost = compiler.global_cost_history[-1] if len(compiler.global_cost_history) > 0 else 1
        alpha = abs(previous_cost - compiler.adapt_config.sufficient_cost)
        return global_cost - alpha * sum(np.absolute(amps[1:]) ** 2)

    def evaluate_local_cost(self, compiler):
        evals = self.measure_qubit_expectation_values(compiler)
        return 0.5 * (1 - np.mean(evals))

    def evaluate_circuit(self, compiler):
        circ = compiler.full_circuit.copy()
        view = self.mps_ops.mps_from_circuit(circ, return_preprocessed=True, sim=self.simulator)
        view._pair_hint = getattr(compiler, "coupling_map", None)   # all candidate pairs in one batch
        return view

    def measure_qubit_expectation_values(self, compiler):
        mps = self.evaluate_circuit(compiler)
        return [self.mps_ops.mps_expectation(mps, "Z", i, already_preprocessed=True)
                for i in range(compiler.full_circuit.num_qubits)]

    def evaluate_hamming_weight_one_overlaps(self, mps):
        h, tmp = self.mps_ops._device(mps, True)
        out = list(np.absolute(h.amps([1 << i for i in range(len(mps))])) ** 2)
        if tmp:
            self.simulator._recycle(h)
        return out

    # ---- general-gradient pair heuristic ----
    @staticmethod
    def _two_qubit_unitary(circuit):
        """4x4 matrix (index = bit(qubit 0) + 2 bit(qubit 1)) of a 2-qubit circuit of basis gates."""
        from .sv_engine import embed_entry
        if circuit.num_qubits != 2:
            raise ValueError("generators / ansatz of general_grad_of_pairs act on two qubits")
        u = np.eye(4, dtype=np.complex128)
        for ent in G.canonical_window(circuit):
            u = embed_entry(ent, (0, 1)) @ u
        return u

    def general_grad_of_pairs(self, circuit, inverse_zero_ansatz, generators, degeneracies, coupling_map,
                              starting_circuit=None):
        """adaptaqc/utils/gradients.py:23-124 with the same arguments and the same return value
        (g_pair = sqrt(sum_k deg_k Im(<s|G_k|psi><psi|U^+(0)|s>)^2) for every pair), from ONE simulation of |psi> and ONE
        batched read-out instead of P x (#generators + 1) simulator runs + mps_dot calls.

        Every operator involved acts on the pair only, so with T_p[i][j] = <s|(|i><j| on p)|psi>
            <s|G_k|psi>     = sum_ij (M_k^+)[i][j] T_p[i][j]        (generators[k] is the circuit of G_k^+, matrix M_k)
            <psi|U^+(0)|s>  = conj(sum_ij (M_0^+)[i][j] T_p[i][j])  (inverse_zero_ansatz has matrix M_0)
        For |s> = |0..0> (no starting circuit) T_p[0][j] is the amplitude of |psi> on the bitstring that has j on the
        pair and zeros elsewhere: all 4 P amplitudes come from ONE kernel launch (b200_mps_amps).  With a starting
        circuit T_p comes from one left + one right environment sweep of <s|psi> (b200_mps_pair_transfer)."""
        sim = self.simulator
        pairs = [(int(c), int(t)) for c, t in coupling_map]
        psi = sim.simulate(circuit)
        try:
            if starting_circuit is None or len(starting_circuit.data) == 0:
                bits = [((j & 1) << c) | ((j >> 1) << t) for c, t in pairs for j in range(4)]
                amps = psi.amps(bits).reshape(len(pairs), 4)
                T = np.zeros((len(pairs), 4, 4), dtype=np.complex128)
                T[:, 0, :] = amps
            else:
                s_state = sim.simulate(starting_circuit, use_checkpoints=False)
                try:
                    T = s_state.pair_transfer(psi, pairs)
                finally:
                    sim._recycle(s_state)
        finally:
            sim._recycle(psi)
        return self.gradients_from_pair_transfers(T, inverse_zero_ansatz, generators, degeneracies)

    @classmethod
    def gradients_from_pair_transfers(cls, T, inverse_zero_ansatz, generators, degeneracies):
        """The 4x4 host algebra of general_grad_of_pairs: T[p][i][j] = <s|(|i><j| on pair p)|psi>."""
        m0 = cls._two_qubit_unitary(inverse_zero_ansatz).conj().T
        gens = [cls._two_qubit_unitary(g).conj().T for g in generators]
        gradients = []
        for Tp in T:
            zero_overlap = np.conj(np.sum(m0 * Tp))
            total = 0
            for mk, deg in zip(gens, degeneracies):
                overlap = np.sum(mk * Tp)
                total += (-1 * np.imag(overlap * zero_overlap)) ** 2 * deg
            gradients.append(np.sqrt(total))
        return gradients

    # ---- batched extension (B200CostMinimiser) ----
    SHIFT_WORKERS = 2       # concurrent candidate simulations (the blocked Jacobi SVD occupies 64 of the 148 SMs)

    # Under real truncation the shift values of a gate can be simulated concurrently (_shift_costs_truncating).  Measured
    # on C4 (profiles/bench_r2n_mps.json) that is SLOWER than one scalar at a time -- 49.2 vs 58.4 evals/s: only 9 of the 24
    # evaluations of a Rotosolve cycle need an SVD at all (prefix checkpoints), two cooperative Jacobi launches do not
    # overlap enough to pay for the second context's checkpoint misses -- so the batched front end uses it only on request.
    batch_truncating = False

    def supports_shift_costs(self, compiler):
        if compiler.soften_global_cost or getattr(compiler, "optimise_local_cost", False):
            return False
        return self._use_incremental(compiler) or self.batch_truncating

    def _shift_costs_truncating(self, compiler, gate_index, candidates):
        """Reference contraction order (real truncation: bond cap or a threshold above roundoff).  The candidates are
        INDEPENDENT simulations that share every gate before `gate_index`; each worker simulator has its own device
        context (own stream, own scratch, own prefix checkpoints) and is driven by its own thread -- ctypes releases the
        GIL, so the SVDs of different candidates overlap on the GPU.  Every candidate's value is what the one-scalar
        path computes for that circuit: same gates, same order, same truncation rule."""
        import concurrent.futures
        o = self.simulator.options
        if getattr(self, "_shift_pool", None) is None:
            self._shift_pool = concurrent.futures.ThreadPoolExecutor(max_workers=self.SHIFT_WORKERS)
            self._shift_sims = []
        key = (o.matrix_product_state_truncation_threshold, o.matrix_product_state_max_bond_dimension)
        if getattr(self, "_shift_key", None) != key:
            self._shift_sims = [B200MPSSimulator(key[0], key[1], device=self.simulator.device) for _ in range(self.SHIFT_WORKERS)]
            self._shift_key = key
        circuit = compiler.full_circuit
        data = list(circuit.data)
        if data[gate_index].operation.name not in G.ROTATIONS and getattr(data[gate_index].operation, "label", None) not in G.ROTATIONS:
            raise ValueError(f"full_circuit index {gate_index} is not a rotation gate")
        qubits = data[gate_index].qubits

        def one(worker, jobs):
            sim = self._shift_sims[worker]
            out = []
            for j in jobs:
                name, theta = candidates[j]
                view = _CircuitView(circuit, data, gate_index, _Inst(_Op(name, [float(theta)]), qubits))
                handle = sim.simulate(view)
                out.append((j, 1 - np.absolute(handle.amps([0])[0]) ** 2))
                sim._recycle(handle)
            return out

        jobs = [list(range(w, len(candidates), self.SHIFT_WORKERS)) for w in range(self.SHIFT_WORKERS)]
        futures = [self._shift_pool.submit(one, w, jb) for w, jb in enumerate(jobs) if jb]
        costs = [None] * len(candidates)
        for f in futures:
            for j, c in f.result():
                costs[j] = c
        return costs

    def shift_costs(self, compiler, gate_index, candidates):
        if not self._use_incremental(compiler):
            if not self.supports_shift_costs(compiler):
                raise NotImplementedError("batched shifts are not defined for the softened / local cost")
            return self._shift_costs_truncating(compiler, gate_index, candidates)
        ev, window, start = self.amp0(compiler, with_start=True)
        # the window starts right after the set_matrix_product_state instruction (index 1), or at 0 for a circuit
        # target that was not converted -- not necessarily at lhs_gate_count
        k = gate_index - start
        if len(window) != len(compiler.full_circuit.data) - start:   # barriers / delays were dropped from the window
            k = G.window_index(compiler.full_circuit, start, gate_index)
        if not 0 <= k < len(window) or window[k][2] >= 0:
            raise ValueError(f"full_circuit index {gate_index} is not a 1-qubit gate")
        mats = [G.one_qubit_matrix(name, theta) for name, theta in candidates]
        return [1 - np.absolute(a) ** 2 for a in ev.shift_amplitudes(window, k, mats)]
