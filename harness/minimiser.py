"""TEST / BENCH HARNESS (not product code): host-side mirror of the reference optimiser.

``CostMinimiser`` restates the Rotosolve / Rotoselect logic of
adaptaqc/utils/cost_minimiser.py:52-106, 267-368 and the closed forms of
adaptaqc/utils/utilityfunctions.py:34-57, 272-278 on the qiskit-free circuit container.  It asks
the backend for ONE scalar per call, exactly like the reference, and is what the parity tests
and ``bench.py`` drive (the real ``adaptaqc`` package cannot be imported without qiskit).

``B200CostMinimiser`` here = the product's batched front end (adapt_aqc_b200.minimiser.make_b200_minimiser)
bound to this harness ``CostMinimiser``; on a real installation it binds to adaptaqc's own class.
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


# The batched front end is PRODUCT code (adapt_aqc_b200.minimiser); here it is bound to this harness optimiser
from adapt_aqc_b200.minimiser import make_b200_minimiser  # noqa: E402

B200CostMinimiser = make_b200_minimiser(CostMinimiser, replace_1q_gate, minimum_of_sinusoidal)
