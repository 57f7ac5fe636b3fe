"""Host-side mirror of the ADAPT-AQC compile loop (test / bench harness).

The real ``adaptaqc.compilers.AdaptCompiler`` needs qiskit, which cannot be installed in the
build image, so the loop that *drives* the backend is restated here on the qiskit-free circuit
container, following the reference line by line:

  ApproximateCompiler.__init__ / _prepare_full_circuit / evaluate_cost
        adaptaqc/compilers/approximate_compiler.py:74-163, 435-527
  AdaptCompiler.compile / _add_layer / pair selection / reuse priorities / MPS layer caching
        adaptaqc/compilers/adapt/adapt_compiler.py:246-482, 585-1163
  AdaptConfig                 adaptaqc/compilers/adapt/adapt_config.py:16-97
  gate pruning                adaptaqc/utils/circuit_operations/circuit_operations_optimisation.py:31-204
  coupling maps               adaptaqc/utils/constants.py:34-110

It talks to the backend only through the reference's interface (the four AQCBackend methods,
``backend.simulator.run`` for the per-pair statevector, the mps_operations-shaped helpers), so
the same harness runs on the CPU oracle backends (oracle/) and on the B200 backends, and the two
can be compared decision by decision (chosen pairs, layer count, costs).
"""
import logging
import os
import pickle
import timeit
from pathlib import Path

import numpy as np

from . import measures as em
from . import minimiser as mini
from .circuit import Circuit, CircuitInstruction, Gate
from .minimiser import ALG_ROTOSELECT, ALG_ROTOSOLVE, CostMinimiser, has_stopped_improving

logger = logging.getLogger(__name__)

CMAP_FULL, CMAP_LINEAR, CMAP_LADDER = "CMAP_FULL", "CMAP_LINEAR", "CMAP_LADDER"
DEFAULT_SUFFICIENT_COST = 1e-2
MINIMUM_ROTATION_ANGLE = 1e-3


# ---- constants.py:34-110 ----------------------------------------------------------------------
def generate_coupling_map(num_qubits, map_kind, both_dir=False, loop=False):
    c_map = []
    if map_kind == CMAP_FULL:
        for i in range(1, num_qubits):
            for j in range(num_qubits - i):
                c_map.append((j, j + i))
    elif map_kind == CMAP_LINEAR:
        for j in range(num_qubits - 1):
            c_map.append((j, j + 1))
        if loop:
            c_map.append((num_qubits - 1, 0))
    elif map_kind == CMAP_LADDER:
        j = 0
        while j + 1 <= num_qubits - 1:
            c_map.append((j, j + 1))
            j += 2
        j = 1
        if loop and num_qubits % 2 == 1:
            c_map.append((num_qubits - 1, 0))
        while j + 1 <= num_qubits - 1:
            c_map.append((j, j + 1))
            j += 2
        if loop and num_qubits % 2 == 0:
            c_map.append((num_qubits - 1, 0))
    else:
        raise ValueError(f"Invalid coupling map type {map_kind}")
    if both_dir:
        c_map += [(t, s) for (s, t) in c_map]
    return c_map


def remove_permutations_from_coupling_map(coupling_map):
    seen, unique = set(), []
    for pair in coupling_map:
        if tuple(sorted(pair)) not in seen:
            seen.add(tuple(sorted(pair)))
            unique.append(pair)
    return unique


class AdaptConfig:
    """adapt_config.py:16-97 (same names, same defaults)."""

    def __init__(self, max_layers=int(1e5), sufficient_cost=DEFAULT_SUFFICIENT_COST, max_2q_gates=1e4,
                 cost_improvement_num_layers=10, cost_improvement_tol=1e-2, max_layers_to_modify=100,
                 method="ISL", bad_qubit_pair_memory=10, reuse_exponent=0, reuse_priority_mode="pair",
                 rotosolve_frequency=1, rotoselect_tol=1e-5, rotosolve_tol=1e-3, entanglement_threshold=1e-8):
        self.bad_qubit_pair_memory = bad_qubit_pair_memory
        self.max_layers = max_layers
        self.sufficient_cost = sufficient_cost
        self.max_2q_gates = max_2q_gates
        self.cost_improvement_tol = cost_improvement_tol
        self.cost_improvement_num_layers = int(cost_improvement_num_layers)
        self.max_layers_to_modify = max_layers_to_modify
        self.method = method
        self.rotosolve_frequency = rotosolve_frequency
        self.rotoselect_tol = rotoselect_tol
        self.rotosolve_tol = rotosolve_tol
        self.entanglement_threshold = entanglement_threshold
        self.reuse_exponent = reuse_exponent
        self.reuse_priority_mode = reuse_priority_mode.lower()


class AdaptResult:
    """adapt_result.py:14-70 (the fields the harness fills)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


# ---- circuit editing helpers (circuit_operations_*.py) ----------------------------------------
def add_to_circuit(circuit, other, location=None, qubit_subset=None):
    """Insert `other`'s gates into circuit.data at `location` (default: end)."""
    location = len(circuit.data) if location is None else location
    qs = list(range(other.num_qubits)) if qubit_subset is None else list(qubit_subset)
    new = [CircuitInstruction(i.operation.copy(), [qs[q] for q in i.qubits]) for i in other.data]
    circuit.data[location:location] = new


def add_dressed_cnot(circuit, control, target, thinly_dressed=False, v1=True, v2=True, v3=True, v4=True):
    """circuit_operations_basic.py:135-189; rotations are labelled rz/ry so Rotoselect sees them."""
    def dress(q):
        circuit.append(mini.create_1q_gate("rz", 0), [q])
        if not thinly_dressed:
            circuit.append(mini.create_1q_gate("ry", 0), [q])
            circuit.append(mini.create_1q_gate("rz", 0), [q])
    if v1: dress(control)
    if v2: dress(target)
    circuit.append(Gate("cx"), [control, target])
    if v3: dress(control)
    if v4: dress(target)


def circuit_by_inverting_circuit(circuit):
    """circuit_operations_full_circuit.py:364-382: labelled rotations keep their label."""
    new = Circuit(circuit.num_qubits)
    for inst in circuit.data[::-1]:
        gate = inst.operation
        if gate.label not in ("rx", "ry", "rz"):
            inv = gate.inverse()
        else:
            inv = gate.copy()
            inv.params[0] *= -1
        inv.label = gate.label
        new.data.append(CircuitInstruction(inv, inst.qubits))
    return new


def extract_inner_circuit(circuit, gate_range):
    new = Circuit(circuit.num_qubits)
    new.data = [CircuitInstruction(i.operation.copy(), i.qubits) for i in circuit.data[gate_range[0]:gate_range[1]]]
    return new


def find_num_gates(circuit, gate_range=None):
    """(num_2q, num_1q) in range; circuit_operations_full_circuit.py:283-309."""
    rng = range(*gate_range) if gate_range is not None else range(len(circuit.data))
    n1 = sum(1 for i in rng if len(circuit.data[i].qubits) == 1 and isinstance(circuit.data[i].operation, Gate))
    n2 = sum(1 for i in rng if len(circuit.data[i].qubits) == 2 and isinstance(circuit.data[i].operation, Gate))
    return n2, n1


def multi_qubit_gate_depth(circuit):
    return circuit.depth(filter_function=lambda inst: len(inst.qubits) > 1 and isinstance(inst.operation, Gate))


def find_previous_gate_on_qubit(circuit, gate_index):
    """circuit_operations_circuit_division.py: previous instruction sharing a qubit."""
    qubits = set(circuit.data[gate_index].qubits)
    for i in range(gate_index - 1, -1, -1):
        if qubits & set(circuit.data[i].qubits):
            return circuit.data[i].operation, i
    return None, None


def zyz_angles(mat):
    """(theta, phi, lam) with mat = e^{i g} Rz(phi) Ry(theta) Rz(lam); qiskit's
    OneQubitEulerDecomposer().angles, used at circuit_operations_optimisation.py:153."""
    det_arg = np.angle(np.linalg.det(mat))
    theta = 2.0 * np.arctan2(abs(mat[1, 0]), abs(mat[0, 0]))
    ang1, ang2 = np.angle(mat[1, 1]), np.angle(mat[1, 0])
    return theta, ang1 + ang2 - det_arg, ang1 - ang2


def remove_unnecessary_1q_gates_from_circuit(circuit, remove_zero_gates=True, remove_small_gates=False,
                                             gate_range=None, min_rotation_angle=MINIMUM_ROTATION_ANGLE):
    """circuit_operations_optimisation.py:77-166."""
    if gate_range is None:
        gate_range = (0, len(circuit.data))
    to_remove, dealt_with = [], []

    def removable(g):
        return (remove_zero_gates and g.params[0] == 0) or (
            remove_small_gates and np.absolute(g.params[0]) < min_rotation_angle)

    for gate_index in range(gate_range[1] - 1, gate_range[0] - 1, -1):
        gate = circuit.data[gate_index].operation
        if gate_index in to_remove or gate_index in dealt_with or not mini.is_supported_1q_gate(gate):
            continue
        if removable(gate):
            to_remove.append(gate_index)
            continue
        matrix = gate.to_matrix()
        prev_idx = [gate_index]
        prev_gate, prev_gate_index = find_previous_gate_on_qubit(circuit, gate_index)
        while prev_gate is not None and mini.is_supported_1q_gate(prev_gate) and prev_gate_index >= gate_range[0]:
            if removable(prev_gate):
                to_remove.append(prev_gate_index)
            else:
                prev_idx.append(prev_gate_index)
                matrix = np.matmul(matrix, prev_gate.to_matrix())
            prev_gate, prev_gate_index = find_previous_gate_on_qubit(circuit, prev_gate_index)
        if len(prev_idx) > 3:
            theta, phi, lam = zyz_angles(matrix)
            mini.replace_1q_gate(circuit, prev_idx[0], "rz", phi)
            mini.replace_1q_gate(circuit, prev_idx[1], "ry", theta)
            mini.replace_1q_gate(circuit, prev_idx[2], "rz", lam)
            dealt_with += [prev_idx[1], prev_idx[2]]
            to_remove += prev_idx[3:]
        else:
            dealt_with += prev_idx
    for index in sorted(set(to_remove), reverse=True):
        del circuit.data[index]


def remove_unnecessary_2q_gates_from_circuit(circuit, gate_range=None):
    """circuit_operations_optimisation.py:169-204."""
    if gate_range is None:
        gate_range = (0, len(circuit.data))
    to_remove = []
    for gate_index in range(gate_range[1] - 1, gate_range[0] - 1, -1):
        inst = circuit.data[gate_index]
        if inst.operation.name not in ("cx", "cy", "cz") or gate_index in to_remove:
            continue
        prev_gate, prev_gate_index = find_previous_gate_on_qubit(circuit, gate_index)
        if prev_gate is None or prev_gate.name != inst.operation.name or prev_gate_index < gate_range[0]:
            continue
        if prev_gate_index in to_remove:
            continue
        if circuit.data[prev_gate_index].qubits == inst.qubits:
            to_remove += [gate_index, prev_gate_index]
    for index in sorted(to_remove, reverse=True):
        del circuit.data[index]


def remove_unnecessary_gates_from_circuit(circuit, remove_zero_gates=True, remove_small_gates=False,
                                          gate_range=None):
    """circuit_operations_optimisation.py:31-74."""
    gate_range = [0, len(circuit.data)] if gate_range is None else list(gate_range)
    last_len = len(circuit.data)
    i = 0
    while True:
        if i == 0:
            remove_unnecessary_1q_gates_from_circuit(circuit, remove_zero_gates, remove_small_gates, gate_range)
            i = 1
        else:
            remove_unnecessary_2q_gates_from_circuit(circuit, gate_range)
            i = 0
        new_len = len(circuit.data)
        if new_len != last_len:
            gate_range[1] -= last_len - new_len
            last_len = new_len
        elif i == 0:
            return


# ---- entanglement_measures.py:39-98 / circuit_operations_running.py:44-69 ----------------------
def run_circuit_without_transpilation(circuit, backend, backend_options=None, execute_kwargs=None,
                                      return_statevector=False):
    job = backend.simulator.run(circuit, **(backend_options or {}), **(execute_kwargs or {}))
    result = job.result()
    if not return_statevector:
        raise NotImplementedError("counts are only produced by the sampling backend (out of scope)")
    return result.get_statevector()


def partial_trace(statevector, a, b):
    """entanglement_measures.py:325-340.  The statevector object owns the arithmetic: a
    DeviceStatevector runs the RDM kernel, the oracle's statevector its C restatement."""
    return statevector.partial_trace(a, b)


def calculate_entanglement_measure(method, circuit, qubit_1, qubit_2, backend, backend_options=None,
                                   execute_kwargs=None, mps=None):
    if backend.kind == "sv":
        statevector = run_circuit_without_transpilation(circuit, backend, return_statevector=True)
        rho = partial_trace(statevector, qubit_1, qubit_2)
    elif backend.kind == "mps":
        rho = backend.mps_ops.partial_trace(mps, [qubit_1, qubit_2], already_preprocessed=True)
    else:
        raise NotImplementedError("tomography on sampling backends is out of scope")
    return em.measure_from_rho(method, rho)


class AdaptCompiler:
    """Mirror of ApproximateCompiler + AdaptCompiler for circuit / MPS targets, |0..0> input."""

    def __init__(self, target, entanglement_measure=em.EM_TOMOGRAPHY_CONCURRENCE, backend=None,
                 execute_kwargs=None, coupling_map=None, adapt_config=None, custom_layer_2q_gate=None,
                 starting_circuit=None, use_roto_algos=True, use_rotoselect=True, rotosolve_fraction=1.0,
                 optimise_local_cost=False, soften_global_cost=False, initial_single_qubit_layer=False,
                 minimiser_cls=None):
        if backend is None:
            raise ValueError("a backend is required")
        if not use_roto_algos:
            raise NotImplementedError("only the Rotosolve/Rotoselect optimisers are on the hot path")
        self.target = target
        self.backend = backend
        self.is_statevector_backend = backend.kind == "sv"
        self.is_aer_mps_backend = backend.kind == "mps"
        target_is_mps = isinstance(target, tuple)
        if target_is_mps and not self.is_aer_mps_backend:
            raise Exception("Aer MPS backend must be used when target is an Aer MPS")
        self.circuit_to_compile = self.prepare_circuit()
        self.execute_kwargs = dict(execute_kwargs or {})
        self.execute_kwargs.setdefault("shots", 1)
        self.execute_kwargs.setdefault("optimization_level", 0)
        self.backend_options = {"method": "automatic"}
        self.total_num_qubits = self.circuit_to_compile.num_qubits
        self.qubit_subset_to_compile = list(range(self.total_num_qubits))
        self.general_initial_state = False
        self.starting_circuit = starting_circuit
        if self.is_aer_mps_backend:
            self.zero_mps = backend.mps_ops.mps_from_circuit(Circuit(self.total_num_qubits), return_preprocessed=True)
        self.optimise_local_cost = optimise_local_cost
        self.soften_global_cost = soften_global_cost
        self.full_circuit, self.lhs_gate_count, self.rhs_gate_count = self._prepare_full_circuit()
        if not 0 < rotosolve_fraction <= 1:
            raise ValueError("rotosolve_fraction must be in the range (0,1]")
        if minimiser_cls is None or minimiser_cls is CostMinimiser:
            self.minimizer = CostMinimiser(self.evaluate_cost, self.variational_circuit_range, self.full_circuit,
                                           rotosolve_fraction)
        else:
            self.minimizer = minimiser_cls(self, rotosolve_fraction)
        self.cost_evaluation_counter = 0
        self.compiling_finished = False

        # ---- AdaptCompiler.__init__ (adapt_compiler.py:121-220) ----
        self.entanglement_measure_method = entanglement_measure
        self.adapt_config = adapt_config if adapt_config is not None else AdaptConfig()
        if coupling_map is None:
            coupling_map = generate_coupling_map(self.total_num_qubits, CMAP_FULL, False, False)
        self.remove_unnecessary_gates_during_adapt = custom_layer_2q_gate is None
        self.use_rotoselect = use_rotoselect
        self.layer_2q_gate = self.construct_layer_2q_gate(custom_layer_2q_gate)
        self.coupling_map = remove_permutations_from_coupling_map(coupling_map)
        self.qubit_pair_history = []
        self.bad_qubit_pairs = []
        self.pair_selection_method_history = []
        self.entanglement_measures_history = []
        self.e_val_history = []
        self.general_gradient_history = []
        self.time_taken = None
        self.initial_single_qubit_layer = initial_single_qubit_layer
        if self.is_aer_mps_backend:
            self.layers_saved_to_mps = self.full_circuit.copy()
            del self.layers_saved_to_mps.data[1:]
        self.layers_as_gates = []
        self.resume_from_layer = None
        self.prev_checkpoint_time_taken = None
        if self.adapt_config.method == "general_gradient":
            if not self.is_aer_mps_backend:
                raise ValueError("general_gradient method is only implemented for Aer MPS backend")
            from . import gradients as gr
            self.generators, self.degeneracies = gr.get_generators_and_degeneracies(
                self.layer_2q_gate, use_rotoselect, inverse=True)
            self.inverse_zero_ansatz = self.layer_2q_gate.inverse()
        if self.soften_global_cost and self.optimise_local_cost:
            raise ValueError("soften_global_cost must be False when optimising local cost")

    # ---- ApproximateCompiler ------------------------------------------------------------------
    def prepare_circuit(self):
        """approximate_compiler.py:165-217: MPS backends embed the target as one
        set_matrix_product_state instruction."""
        if isinstance(self.target, tuple):
            c = Circuit(len(self.target[0]))
            c.set_matrix_product_state(self.target)
            return c
        prepared = self.target.copy()
        if self.is_aer_mps_backend:
            target_mps = self.backend.mps_ops.mps_from_circuit(prepared, sim=self.backend.simulator)
            c = Circuit(prepared.num_qubits)
            c.set_matrix_product_state(target_mps)
            return c
        return prepared

    def _prepare_full_circuit(self):
        """approximate_compiler.py:435-512 for initial_state=None, general_initial_state=False."""
        qc = Circuit(self.total_num_qubits)
        add_to_circuit(qc, self.circuit_to_compile)
        lhs_gate_count = len(qc.data)
        if self.starting_circuit is not None:
            add_to_circuit(qc, self.starting_circuit.inverse())
        return qc, lhs_gate_count, len(qc.data) - lhs_gate_count

    def variational_circuit_range(self, circuit=None):
        circuit = self.full_circuit if circuit is None else circuit
        return self.lhs_gate_count, len(circuit.data) - self.rhs_gate_count

    def evaluate_cost(self):
        """approximate_compiler.py:514-527."""
        self.cost_evaluation_counter += 1
        if self.optimise_local_cost:
            return self.backend.evaluate_local_cost(self)
        return self.backend.evaluate_global_cost(self)

    def get_compiled_circuit(self):
        """approximate_compiler.py:385-433 without the register bookkeeping."""
        compiled = circuit_by_inverting_circuit(extract_inner_circuit(self.full_circuit, self.variational_circuit_range()))
        if self.starting_circuit is not None:
            add_to_circuit(compiled, self.starting_circuit, 0)
        return compiled

    # ---- AdaptCompiler ------------------------------------------------------------------------
    def construct_layer_2q_gate(self, custom_layer_2q_gate):
        if custom_layer_2q_gate is None:
            qc = Circuit(2)
            add_dressed_cnot(qc, 0, 1, True)
            return qc
        for inst in custom_layer_2q_gate.data:
            gate = inst.operation
            if gate.label is None and gate.name in mini.SUPPORTED_1Q_GATES:
                gate.label = gate.name
        return custom_layer_2q_gate

    def get_layer_2q_gate(self, layer_index):
        return self.layer_2q_gate.copy()

    def compile(self, checkpoint_every=0, checkpoint_dir="checkpoint/", delete_prev_chkpt=False):
        """adapt_compiler.py:246-482."""
        start_time = timeit.default_timer()
        if self.resume_from_layer is None:
            self.time_taken = 0
            start_point = 0
            self.cost_evaluation_counter = 0
            self.global_cost, self.local_cost = None, None
            self.cnot_depth = None
            self.global_cost_history = []
            if self.optimise_local_cost:
                self.local_cost_history = []
            self.cnot_depth_history = []
            self.g_range = self.variational_circuit_range
            self.original_lhs_gate_count = self.lhs_gate_count
        else:
            start_point = self.resume_from_layer
            self.time_taken = self.prev_checkpoint_time_taken
        if checkpoint_every > 0:
            Path(checkpoint_dir).mkdir(parents=True, exist_ok=True)

        for layer_count in range(start_point, self.adapt_config.max_layers):
            if self.optimise_local_cost:
                self.local_cost = self._add_layer(layer_count)
                self.global_cost = self.backend.evaluate_global_cost(self)
                self.local_cost_history.append(self.local_cost)
            else:
                self.global_cost = self._add_layer(layer_count)
            self.global_cost_history.append(self.global_cost)
            self.record_cnot_depth()

            if self.remove_unnecessary_gates_during_adapt and not self.is_aer_mps_backend:
                remove_unnecessary_gates_from_circuit(self.full_circuit, False, False, gate_range=self.g_range())

            ref = self.ref_circuit_as_gates if self.is_aer_mps_backend else self.full_circuit
            num_2q_gates, num_1q_gates = find_num_gates(
                ref, self.g_range(self.ref_circuit_as_gates if self.is_aer_mps_backend else None))

            cinl = self.adapt_config.cost_improvement_num_layers
            cit = self.adapt_config.cost_improvement_tol
            if len(self.global_cost_history) >= cinl and has_stopped_improving(self.global_cost_history[-cinl:], cit):
                logger.warning("ADAPT-AQC stopped improving")
                self.compiling_finished = True
                break
            if self.global_cost < self.adapt_config.sufficient_cost:
                self.compiling_finished = True
                break
            elif num_2q_gates >= self.adapt_config.max_2q_gates:
                self.minimizer.minimize_cost(algorithm_kind=ALG_ROTOSOLVE, max_cycles=10, tol=1e-5,
                                             stop_val=self.adapt_config.sufficient_cost)
                self.compiling_finished = True
                break
            if checkpoint_every > 0 and layer_count % checkpoint_every == 0:
                self.checkpoint(checkpoint_every, checkpoint_dir, delete_prev_chkpt, layer_count, start_time)

        if self.is_aer_mps_backend:
            self.full_circuit = self.ref_circuit_as_gates
        else:
            self.lhs_gate_count = self.original_lhs_gate_count
        remove_unnecessary_gates_from_circuit(self.full_circuit, True, True, gate_range=self.g_range())

        if self.soften_global_cost:
            self.soften_global_cost = False
            final_global_cost = self.backend.evaluate_global_cost(self)
            self.soften_global_cost = True
        else:
            final_global_cost = self.backend.evaluate_global_cost(self)
        self.global_cost_history.append(final_global_cost)
        compiled_circuit = self.get_compiled_circuit()
        num_2q_gates, num_1q_gates = find_num_gates(compiled_circuit)
        self.cnot_depth_history.append(multi_qubit_gate_depth(compiled_circuit))

        exact_overlap = "Not computable without SV backend"
        if self.is_statevector_backend and hasattr(self.backend, "overlap_between_circuits"):
            # device-side replacement of calculate_overlap_between_circuits (adapt_compiler.py:447-452)
            exact_overlap = self.backend.overlap_between_circuits(self.circuit_to_compile, compiled_circuit)

        return AdaptResult(
            circuit=compiled_circuit, overlap=1 - final_global_cost, exact_overlap=exact_overlap,
            num_1q_gates=num_1q_gates, num_2q_gates=num_2q_gates, cnot_depth_history=self.cnot_depth_history,
            global_cost_history=self.global_cost_history,
            local_cost_history=self.local_cost_history if self.optimise_local_cost else None,
            entanglement_measures_history=self.entanglement_measures_history, e_val_history=self.e_val_history,
            qubit_pair_history=self.qubit_pair_history, method_history=self.pair_selection_method_history,
            time_taken=self.time_taken + (timeit.default_timer() - start_time),
            cost_evaluations=self.cost_evaluation_counter, coupling_map=self.coupling_map,
        )

    def checkpoint(self, checkpoint_every, checkpoint_dir, delete_prev_chkpt, layer_count, start_time):
        """adapt_compiler.py:484-506: the whole compiler, backend included, is pickled."""
        self.resume_from_layer = layer_count + 1
        self.prev_checkpoint_time_taken = self.time_taken + (timeit.default_timer() - start_time)
        with open(os.path.join(checkpoint_dir, f"{layer_count}.pkl"), "wb") as f:
            pickle.dump(self, f)
        if delete_prev_chkpt:
            try:
                os.remove(os.path.join(checkpoint_dir, f"{layer_count - checkpoint_every}.pkl"))
            except FileNotFoundError:
                pass

    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop("g_range", None)  # bound method; restored below
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.g_range = self.variational_circuit_range
        self.minimizer.cost_finder = self.evaluate_cost
        self.minimizer.variational_circuit_range = self.variational_circuit_range
        self.minimizer.full_circuit = self.full_circuit

    def _add_layer(self, index):
        """adapt_compiler.py:585-689."""
        ansatz_start_index = self.variational_circuit_range()[0]
        if self.initial_single_qubit_layer and index == 0:
            layer_idx = self._add_rotation_to_all_qubits()
        else:
            layer_idx = self._add_entangling_layer(index)
        stop_val = 0 if self.optimise_local_cost else self.adapt_config.sufficient_cost

        alg = ALG_ROTOSELECT if (self.use_rotoselect or (self.initial_single_qubit_layer and index == 0)) else ALG_ROTOSOLVE
        cost = self.minimizer.minimize_cost(algorithm_kind=alg, tol=self.adapt_config.rotoselect_tol,
                                            stop_val=stop_val, indexes_to_modify=layer_idx)
        if (self.adapt_config.rotosolve_frequency != 0 and index > 0
                and index % self.adapt_config.rotosolve_frequency == 0):
            multi = self._calculate_multi_layer_optimisation_indices(ansatz_start_index)
            cost = self.minimizer.minimize_cost(algorithm_kind=ALG_ROTOSOLVE, tol=self.adapt_config.rotosolve_tol,
                                                stop_val=stop_val, indexes_to_modify=multi)

        if self.is_aer_mps_backend:
            self.layers_as_gates.append(index)
            num_layers_to_absorb = self._calculate_num_layers_to_absorb(index)
            if num_layers_to_absorb > 0:
                includes_isql = self.layers_as_gates[0] == 0 and self.initial_single_qubit_layer
                num_gates = self._get_num_gates_to_cache(num_layers_to_absorb, includes_isql)
                gates_absorbed = self._absorb_n_gates_into_mps(num_gates)
                add_to_circuit(self.layers_saved_to_mps, gates_absorbed)
                del self.layers_as_gates[:num_layers_to_absorb]
            self._update_reference_circuit()
        return cost

    def _calculate_num_layers_to_absorb(self, index):
        """:691-705"""
        f = self.adapt_config.rotosolve_frequency
        next_rotosolve_layer = index + (f - index % f)
        lowest_index = next_rotosolve_layer - self.adapt_config.max_layers_to_modify + 1
        return len([i for i in self.layers_as_gates if i < lowest_index])

    def _update_reference_circuit(self):
        """:707-715"""
        not_saved = self.full_circuit.copy()
        del not_saved.data[0]
        self.ref_circuit_as_gates = self.layers_saved_to_mps.copy()
        add_to_circuit(self.ref_circuit_as_gates, not_saved)

    def _calculate_multi_layer_optimisation_indices(self, ansatz_start_index):
        """:717-741"""
        isql = int(self.initial_single_qubit_layer)
        num_entangling_layers = self.adapt_config.max_layers_to_modify - isql
        n_first = self.full_circuit.num_qubits * isql
        end = self.variational_circuit_range()[1]
        start = max(ansatz_start_index, end - len(self.layer_2q_gate.data) * num_entangling_layers - n_first)
        first_layer_end = ansatz_start_index + n_first
        if ansatz_start_index < start < first_layer_end:
            start = first_layer_end
        return start, end

    def _add_entangling_layer(self, index):
        """:743-760"""
        control, target = self._find_appropriate_qubit_pair()
        add_to_circuit(self.full_circuit, self.get_layer_2q_gate(index), self.variational_circuit_range()[1],
                       qubit_subset=[control, target])
        self.qubit_pair_history.append((control, target))
        end = self.variational_circuit_range()[1]
        return end - len(self.layer_2q_gate.data), end

    def _add_rotation_to_all_qubits(self):
        """:762-773"""
        first_layer = Circuit(self.full_circuit.num_qubits)
        for q in range(self.full_circuit.num_qubits):
            first_layer.append(Gate("ry", [0]), [q])
        add_to_circuit(self.full_circuit, first_layer, self.variational_circuit_range()[1])
        self.entanglement_measures_history.append([None])
        self.e_val_history.append(None)
        self.general_gradient_history.append(None)
        self.qubit_pair_history.append((None, None))
        self.pair_selection_method_history.append(None)
        end = self.variational_circuit_range()[1]
        return end - self.full_circuit.num_qubits, end

    def _find_appropriate_qubit_pair(self):
        """:775-830"""
        method = self.adapt_config.method
        if method == "random":
            self.pair_selection_method_history.append("random")
            return self.coupling_map[np.random.randint(len(self.coupling_map))]
        if method == "basic":
            self.pair_selection_method_history.append("basic")
            return self.coupling_map[np.argmax(self._get_all_qubit_pair_reuse_priorities(1))]
        if method == "expectation":
            return self._find_best_expectation_qubit_pair()
        if method == "ISL":
            ems = self._get_all_qubit_pair_entanglement_measures()
            self.entanglement_measures_history.append(ems)
            return self._find_best_entanglement_qubit_pair(ems)
        if method == "general_gradient":
            gradients = self._get_all_qubit_pair_gradients()
            self.general_gradient_history.append(gradients)
            self.pair_selection_method_history.append("general_gradient")
            priorities = self._get_all_qubit_pair_reuse_priorities(self.adapt_config.reuse_exponent)
            return self.coupling_map[np.argmax(np.multiply(gradients, priorities))]
        if method == "brickwall":
            n = self.full_circuit.num_qubits
            if n < 2:
                raise ValueError("Cannot pick a pair if there are fewer than two qubits")
            if len(self.qubit_pair_history) == 0 or n == 2 or self.qubit_pair_history[-1][0] is None:
                return (0, 1)
            prev = self.qubit_pair_history[-1]
            nxt = (prev[0] + 2, prev[1] + 2)
            n_odd = n % 2
            if nxt == (n, n + 1):
                return (1 - n_odd, 2 - n_odd)
            if nxt == (n - 1, n):
                return (0 + n_odd, 1 + n_odd)
            return nxt
        raise ValueError(
            f"Invalid compiling method {method}. "
            f"Method must be one of ISL, expectation, random, basic, general_gradient, brickwall")

    def _get_all_qubit_pair_gradients(self):
        """:839-856"""
        from . import gradients as gr
        end = len(self.full_circuit) - (len(self.starting_circuit) if self.starting_circuit is not None else 0)
        circuit = extract_inner_circuit(self.full_circuit, (0, end))
        # same dispatch as adapt_aqc_b200.registration.install_gradients performs on a real adaptaqc installation
        if hasattr(self.backend, "general_grad_of_pairs"):
            return self.backend.general_grad_of_pairs(circuit, self.inverse_zero_ansatz, self.generators, self.degeneracies,
                                                      self.coupling_map, self.starting_circuit)
        return gr.general_grad_of_pairs(circuit, self.inverse_zero_ansatz, self.generators, self.degeneracies,
                                        self.coupling_map, self.starting_circuit, self.backend)

    def _find_best_entanglement_qubit_pair(self, entanglement_measures):
        """:858-921"""
        cfg = self.adapt_config
        reuse_priorities = self._get_all_qubit_pair_reuse_priorities(cfg.reuse_exponent)
        if len(self.entanglement_measures_history) >= 2 + int(self.initial_single_qubit_layer):
            prev_qp_index = self.coupling_map.index(self.qubit_pair_history[-1])
            pre_em = self.entanglement_measures_history[-2][prev_qp_index]
            post_em = self.entanglement_measures_history[-1][prev_qp_index]
            if post_em >= pre_em:
                self.bad_qubit_pairs.append(self.coupling_map[prev_qp_index])
            if len(self.bad_qubit_pairs) > cfg.bad_qubit_pair_memory:
                del self.bad_qubit_pairs[0]
        filtered_ems = [e * p for (e, p) in zip(entanglement_measures, reuse_priorities)]
        for qp in set(self.bad_qubit_pairs):
            reps = len([x for x in self.qubit_pair_history[-1 * cfg.bad_qubit_pair_memory:] if x == qp])
            if reps >= 1:
                filtered_ems[self.coupling_map.index(qp)] = -1
        if max(filtered_ems) <= cfg.entanglement_threshold:
            return self._find_best_expectation_qubit_pair()
        self.pair_selection_method_history.append("ISL")
        self.e_val_history.append(None)
        return self.coupling_map[np.argmax(filtered_ems)]

    def _find_best_expectation_qubit_pair(self):
        """:923-953"""
        reuse_priorities = self._get_all_qubit_pair_reuse_priorities(self.adapt_config.reuse_exponent)
        e_vals = self.backend.measure_qubit_expectation_values(self)
        self.e_val_history.append(e_vals)
        e_val_sums = [e_vals[c] + e_vals[t] for c, t in self.coupling_map]
        e_val_priorities = [2 - e for e in e_val_sums]
        combined = [e * p for (e, p) in zip(e_val_priorities, reuse_priorities)]
        self.pair_selection_method_history.append("expectation")
        return self.coupling_map[np.argmax(combined)]

    def _get_all_qubit_pair_entanglement_measures(self):
        """:955-976"""
        self.circ_mps = self.backend.evaluate_circuit(self) if self.is_aer_mps_backend else None
        return [
            calculate_entanglement_measure(self.entanglement_measure_method, self.full_circuit, control, target,
                                           self.backend, self.backend_options, self.execute_kwargs, self.circ_mps)
            for control, target in self.coupling_map
        ]

    def _get_all_qubit_pair_reuse_priorities(self, k):
        """:984-1000"""
        if not len(self.qubit_pair_history):
            return [1 for _ in range(len(self.coupling_map))]
        mode = self.adapt_config.reuse_priority_mode
        if mode == "pair":
            return [self._get_pair_reuse_priority(qp, k) for qp in self.coupling_map]
        if mode == "qubit":
            return [self._get_qubit_reuse_priority(qp, k) for qp in self.coupling_map]
        raise ValueError(f"Reuse priority mode must be one of: {['pair', 'qubit']}")

    def _is_last_pair(self, qubit_pair):
        return (len(self.qubit_pair_history) > 0 + int(self.initial_single_qubit_layer)
                and qubit_pair == self.qubit_pair_history[-1])

    def _get_qubit_reuse_priority(self, qubit_pair, k):
        """:1008-1038"""
        if self._is_last_pair(qubit_pair):
            return -1
        if k == 0:
            return 1
        rev = self.qubit_pair_history[::-1]

        def last_use(q):
            for i, tup in enumerate(rev):
                if q in tup:
                    return i
            return np.inf
        return np.min([1 - np.exp2(-(last_use(q) + 1) / k) for q in qubit_pair])

    def _get_pair_reuse_priority(self, qubit_pair, k):
        """:1040-1065"""
        if self._is_last_pair(qubit_pair):
            return -1
        if k == 0:
            return 1
        rev = self.qubit_pair_history[::-1]
        try:
            return 1 - np.exp2(-rev.index(qubit_pair) / k)
        except ValueError:
            return 1

    def _get_num_gates_to_cache(self, n, includes_isql=False):
        """:1092-1095"""
        return len(self.layer_2q_gate) * (n - int(includes_isql)) + self.full_circuit.num_qubits * int(includes_isql)

    def _absorb_n_gates_into_mps(self, n):
        """:1097-1145: fold the first n ansatz gates into the set_matrix_product_state instruction."""
        num_gates_to_absorb = n + 1
        circ_to_absorb = self.full_circuit.copy()
        del circ_to_absorb.data[num_gates_to_absorb:]
        gates_absorbed = circ_to_absorb.copy()
        del gates_absorbed.data[0]
        absorbed_mps = self.backend.mps_ops.mps_from_circuit(circ_to_absorb, sim=self.backend.simulator)
        mps_circuit = Circuit(self.full_circuit.num_qubits)
        mps_circuit.set_matrix_product_state(absorbed_mps)
        num_not_absorbed = len(self.full_circuit.data) - num_gates_to_absorb
        if num_not_absorbed != 0:
            del self.full_circuit.data[:-num_not_absorbed]
        else:
            del self.full_circuit.data[:]
        self.full_circuit.data.insert(0, mps_circuit.data[0])
        return gates_absorbed

    def record_cnot_depth(self):
        """:1147-1163"""
        if self.is_aer_mps_backend:
            ansatz = extract_inner_circuit(self.ref_circuit_as_gates, (1, len(self.ref_circuit_as_gates)))
        else:
            ansatz = extract_inner_circuit(self.full_circuit,
                                           (self.original_lhs_gate_count, self.variational_circuit_range()[1]))
        self.cnot_depth = multi_qubit_gate_depth(ansatz)
        self.cnot_depth_history.append(self.cnot_depth)
