"""Host-side mirror of the reference optimiser + the batched B200 front end.

``CostMinimiser`` restates the Rotosolve / Rotoselect logic of
adaptaqc/utils/cost_minimiser.py:52-106, 267-368 and the closed forms of
adaptaqc/utils/utilityfunctions.py:34-57, 272-278 on the qiskit-free circuit container.  It asks
the backend for ONE scalar per call, exactly like the reference, and is what the parity tests
and ``bench.py`` drive (the real ``adaptaqc`` package cannot be imported without qiskit).

``B200CostMinimiser`` is the batched variant: the same decisions,
but the 3 (Rotosolve) or 7 (Rotoselect) costs of one gate come from one ``shift_costs`` call,
i.e. one inner-product kernel launch.  It is installed with ``compiler.minimizer = ...``
(attribute created at adaptaqc/compilers/approximate_compiler.py:151-156).
"""
import logging
import random

import numpy as np

from .circuit import CircuitInstruction, Gate

logger = logging.getLogger(__name__)

ALG_ROTOSOLVE = "rotosolve"
ALG_ROTOSELECT = "rotoselect"
SUPPORTED_1Q_GATES = ["rx", "ry", "rz"]


# ---- closed forms (utilityfunctions.py:34-57, 105-119, 272-278) -------------------------------
def normalized_angle(angle):
    while (angle > np.pi) or (angle < -np.pi):
        if angle > np.pi:
            angle -= 2 * np.pi
        elif angle < np.pi:
            angle += 2 * np.pi
    return angle


def minimum_of_sinusoidal(value_0, value_pi_by_2, value_minus_pi_by_2):
    """argmin / min of a*sin(x+b)+c sampled at 0, pi/2, -pi/2."""
    theta_min = -(np.pi / 2) - np.arctan2(
        2 * value_0 - value_pi_by_2 - value_minus_pi_by_2, value_pi_by_2 - value_minus_pi_by_2
    )
    theta_min = normalized_angle(theta_min)
    intercept_c = 0.5 * (value_pi_by_2 + value_minus_pi_by_2)
    value_pi = (value_pi_by_2 + value_minus_pi_by_2) - value_0
    amplitude_a = 0.5 * (((value_0 - value_pi) ** 2 + (value_pi_by_2 - value_minus_pi_by_2) ** 2) ** 0.5)
    return theta_min, intercept_c - amplitude_a


def has_stopped_improving(cost_history, rel_tol=1e-2):
    try:
        poly_fit_res = np.polyfit(list(range(len(cost_history))), cost_history, 1)
        grad = poly_fit_res[0] / np.absolute(np.mean(cost_history))
        return grad > -1 * rel_tol
    except np.linalg.LinAlgError:
        return False


# ---- circuit helpers (circuit_operations_basic.py:19-132) -------------------------------------
def create_1q_gate(gate_name, angle):
    if gate_name not in SUPPORTED_1Q_GATES:
        raise ValueError(f"Unsupported gate {gate_name}")
    return Gate(gate_name, [angle], label=gate_name)


def replace_1q_gate(circuit, gate_index, gate_name, angle):
    if gate_name is None:
        return
    inst = circuit.data[gate_index]
    circuit.data[gate_index] = CircuitInstruction(create_1q_gate(gate_name, angle), inst.qubits, inst.clbits)


def is_supported_1q_gate(gate):
    if not isinstance(gate, Gate):
        return False
    gate_name = gate.label if gate.label is not None else gate.name
    return gate_name in SUPPORTED_1Q_GATES


def find_rotation_indices(circuit, indices):
    return [i for i in indices if is_supported_1q_gate(circuit.data[i].operation)]


class CostMinimiser:
    """cost_minimiser.py:32-106, 267-368 (Rotosolve / Rotoselect only)."""

    def __init__(self, cost_finder, variational_circuit_range, full_circuit, rotosolve_fraction=1.0):
        self.cost_finder = cost_finder
        self.variational_circuit_range = variational_circuit_range
        self.full_circuit = full_circuit
        self.rotosolve_fraction = rotosolve_fraction

    def minimize_cost(self, algorithm_kind=ALG_ROTOSOLVE, max_cycles=1000, stop_val=-np.inf, tol=1e-10,
                      indexes_to_modify=None, **_unused):
        if algorithm_kind not in (ALG_ROTOSOLVE, ALG_ROTOSELECT):
            raise NotImplementedError(f"optimiser '{algorithm_kind}' is outside the hot path")
        cost_history = []
        cost = self.cost_finder()
        cycles = 0
        while cost > stop_val and cycles < max_cycles:
            cost = self._reduce_cost(algorithm_kind == ALG_ROTOSELECT, indexes_to_modify)
            cycles += 1
            cost_history.append(cost)
            if len(cost_history) > 3 and has_stopped_improving(cost_history[-3:], tol):
                break
        return cost

    def _sample(self, change_1q_gate_kind, indexes_to_modify):
        vrange = self.variational_circuit_range()
        if indexes_to_modify is None:
            indexes_to_modify = vrange
        else:
            indexes_to_modify = (max(indexes_to_modify[0], vrange[0]), min(indexes_to_modify[1], vrange[1]))
        if self.rotosolve_fraction < 1.0 and not change_1q_gate_kind:
            idx = find_rotation_indices(self.full_circuit, list(range(*indexes_to_modify)))
            sample = random.sample(idx, int(np.ceil(self.rotosolve_fraction * len(idx))))
            sample.sort()
            return sample
        return list(range(*indexes_to_modify))

    def _reduce_cost(self, change_1q_gate_kind=False, indexes_to_modify=None):
        cost = 1
        for index in self._sample(change_1q_gate_kind, indexes_to_modify):
            old_gate = self.full_circuit.data[index].operation
            if change_1q_gate_kind and is_supported_1q_gate(old_gate):
                cost = self.replace_with_best_1q_gate(index)
            elif is_supported_1q_gate(old_gate):
                angle, cost = self.find_best_angle(index, old_gate.label)
                replace_1q_gate(self.full_circuit, index, old_gate.label, angle)
        return cost

    def replace_with_best_1q_gate(self, gate_index):
        replace_1q_gate(self.full_circuit, gate_index, "rx", 0)
        cost_identity = self.cost_finder()
        best_gate_name, best_gate_angle, best_gate_cost = None, None, 1
        for gate_name in SUPPORTED_1Q_GATES:
            min_angle, cost = self.find_best_angle(gate_index, gate_name, cost_identity)
            if cost < best_gate_cost:
                best_gate_name, best_gate_angle, best_gate_cost = gate_name, min_angle, cost
        replace_1q_gate(self.full_circuit, gate_index, best_gate_name, best_gate_angle)
        return best_gate_cost

    def find_best_angle(self, gate_index, gate_name, cost_for_identity=None):
        circ_instr = self.full_circuit.data[gate_index]
        costs = []
        angles_to_run = [0, np.pi / 2, -np.pi / 2]
        if cost_for_identity is not None:
            costs.append(cost_for_identity)
            angles_to_run.remove(0)
        for theta in angles_to_run:
            replace_1q_gate(self.full_circuit, gate_index, gate_name, theta)
            costs.append(self.cost_finder())
        theta_min, cost_min = minimum_of_sinusoidal(costs[0], costs[1], costs[2])
        self.full_circuit.data[gate_index] = circ_instr
        return theta_min, cost_min


class B200CostMinimiser(CostMinimiser):
    """Same decisions as CostMinimiser; all shift values of a gate from one device launch.

    `compiler.backend` must provide ``shift_costs(compiler, gate_index, candidates)``
    (B200SVBackend / B200MPSBackend).  ``compiler.cost_evaluation_counter`` advances by the
    number of costs produced so that AdaptResult.cost_evaluations keeps its meaning
    (approximate_compiler.py:522).
    """

    def __init__(self, compiler, rotosolve_fraction=1.0):
        super().__init__(compiler.evaluate_cost, compiler.variational_circuit_range, compiler.full_circuit,
                         rotosolve_fraction)
        self.compiler = compiler

    def _batched_ok(self):
        c = self.compiler
        if not hasattr(c.backend, "shift_costs") or c.optimise_local_cost or c.soften_global_cost:
            return False
        supports = getattr(c.backend, "_use_incremental", None)
        return True if supports is None else bool(supports(c))

    def _shift(self, gate_index, candidates):
        costs = self.compiler.backend.shift_costs(self.compiler, gate_index, candidates)
        self.compiler.cost_evaluation_counter += len(candidates)
        return costs

    def replace_with_best_1q_gate(self, gate_index):
        if not self._batched_ok():
            return super().replace_with_best_1q_gate(gate_index)
        h = np.pi / 2
        cands = [("rx", 0.0)] + [(g, s) for g in SUPPORTED_1Q_GATES for s in (h, -h)]
        costs = self._shift(gate_index, cands)
        cost_identity = costs[0]
        best_gate_name, best_gate_angle, best_gate_cost = None, None, 1
        for j, gate_name in enumerate(SUPPORTED_1Q_GATES):
            min_angle, cost = minimum_of_sinusoidal(cost_identity, costs[1 + 2 * j], costs[2 + 2 * j])
            if cost < best_gate_cost:
                best_gate_name, best_gate_angle, best_gate_cost = gate_name, min_angle, cost
        # reference semantics: the gate was set to rx(0) first, and stays so if nothing beats cost 1
        replace_1q_gate(self.full_circuit, gate_index, "rx", 0)
        replace_1q_gate(self.full_circuit, gate_index, best_gate_name, best_gate_angle)
        return best_gate_cost

    def find_best_angle(self, gate_index, gate_name, cost_for_identity=None):
        if not self._batched_ok():
            return super().find_best_angle(gate_index, gate_name, cost_for_identity)
        h = np.pi / 2
        if cost_for_identity is None:
            c0, cp, cm = self._shift(gate_index, [(gate_name, 0.0), (gate_name, h), (gate_name, -h)])
        else:
            c0 = cost_for_identity
            cp, cm = self._shift(gate_index, [(gate_name, h), (gate_name, -h)])
        return minimum_of_sinusoidal(c0, cp, cm)
