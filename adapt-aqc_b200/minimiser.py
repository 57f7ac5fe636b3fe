"""Batched optimiser front end -- product code.

``make_b200_minimiser(base_cls, replace_1q_gate, minimum_of_sinusoidal)`` derives ``B200CostMinimiser`` from a
Rotosolve / Rotoselect optimiser class with the reference's interface
(adaptaqc/utils/cost_minimiser.py:32-106, 267-368): the same decisions, but the 3 (Rotosolve) or 7 (Rotoselect)
costs of one gate come from ONE ``backend.shift_costs`` call, i.e. one transfer-matrix launch, instead of 3 / 7
``evaluate_cost`` round trips.  It is installed with ``compiler.minimizer = B200CostMinimiser(compiler)``
(attribute created at adaptaqc/compilers/approximate_compiler.py:151-156).

``adapt_aqc_b200.minimiser.B200CostMinimiser`` resolves lazily against the reference's own ``CostMinimiser`` when
``adaptaqc`` is importable; the qiskit-free test harness binds the same factory to its restated optimiser
(harness/minimiser.py).
"""
import numpy as np


def make_b200_minimiser(base_cls, replace_1q_gate, minimum_of_sinusoidal, supported_1q_gates=("rx", "ry", "rz")):
    """base_cls: CostMinimiser-shaped class; replace_1q_gate(circuit, index, name, angle) and
    minimum_of_sinusoidal(c0, c_plus, c_minus) -> (theta, cost) are the reference's helpers
    (circuit_operations_basic.py:70-99, utilityfunctions.py:34-57)."""

    class B200CostMinimiser(base_cls):
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
            supports = getattr(c.backend, "supports_shift_costs", None) or getattr(c.backend, "_use_incremental", None)
            return True if supports is None else bool(supports(c))

        def _shift(self, gate_index, candidates):
            costs = self.compiler.backend.shift_costs(self.compiler, gate_index, candidates)
            self.compiler.cost_evaluation_counter += len(candidates)
            return costs

        def replace_with_best_1q_gate(self, gate_index):
            if not self._batched_ok():
                return super().replace_with_best_1q_gate(gate_index)
            h = np.pi / 2
            cands = [("rx", 0.0)] + [(g, s) for g in supported_1q_gates for s in (h, -h)]
            costs = self._shift(gate_index, cands)
            cost_identity = costs[0]
            best_gate_name, best_gate_angle, best_gate_cost = None, None, 1
            for j, gate_name in enumerate(supported_1q_gates):
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

    return B200CostMinimiser


def __getattr__(name):
    if name == "B200CostMinimiser":      # pragma: no cover - needs the reference package (qiskit)
        from adaptaqc.utils import circuit_operations as co
        from adaptaqc.utils.cost_minimiser import CostMinimiser
        from adaptaqc.utils.utilityfunctions import minimum_of_sinusoidal
        cls = make_b200_minimiser(CostMinimiser, co.replace_1q_gate, minimum_of_sinusoidal)
        globals()[name] = cls
        return cls
    raise AttributeError(name)
