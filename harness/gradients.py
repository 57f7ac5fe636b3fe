"""Harness mirror of the general-gradient pair heuristic (adaptaqc/utils/gradients.py:23-224).

Like ``compiler.py`` this exists only so that tests / bench can drive the backends without qiskit.
It talks to the backend through ``backend.mps_ops`` (the aqc_research-shaped functions) and
``backend.simulator`` exactly where the reference does: one small MPS per (pair, generator) and
one ``mps_dot`` against the circuit's MPS.
"""
import numpy as np

from .circuit import Circuit

ROT = ("rx", "ry", "rz")


def _drop_cancelling_cx(circ):
    """Adjacent identical cx pairs resolve to the identity (circuit_operations_optimisation.py:169-204)."""
    changed = True
    while changed:
        changed = False
        for i in range(len(circ.data) - 1):
            a, b = circ.data[i], circ.data[i + 1]
            if a.operation.name == "cx" and b.operation.name == "cx" and a.qubits == b.qubits:
                del circ.data[i:i + 2]
                changed = True
                break
    return circ


def get_generator(ansatz, index, op):
    """Ansatz at zero angles with the rotation at `index` replaced by the Pauli generating `op`
    (gradients.py:170-224)."""
    if op not in ROT:
        raise ValueError("op must be one of rx, ry or rz")
    gen = Circuit(2)
    for i, inst in enumerate(ansatz.data):
        name = inst.operation.name
        if name not in ROT + ("cx",):
            raise ValueError("Circuit must only contain rx, ry, rz and cx gates")
        if i == index:
            getattr(gen, op[1])(inst.qubits[0])
        if name == "cx":
            gen.cx(*inst.qubits)
    return _drop_cancelling_cx(gen)


def get_generators_and_degeneracies(ansatz, rotoselect=False, inverse=False):
    """gradients.py:127-168 + utilityfunctions.py:401-426."""
    gens = []
    for i, inst in enumerate(ansatz.data):
        if inst.operation.name in ROT:
            for op in (ROT if rotoselect else (inst.operation.name,)):
                g = get_generator(ansatz, i, op)
                gens.append(g.inverse() if inverse else g)
    distinct, degeneracies = [], []
    for g in gens:
        for j, d in enumerate(distinct):
            if g == d:
                degeneracies[j] += 1
                break
        else:
            distinct.append(g)
            degeneracies.append(1)
    return distinct, degeneracies


def general_grad_of_pairs(circuit, inverse_zero_ansatz, generators, degeneracies, coupling_map,
                          starting_circuit=None, backend=None):
    """g_pair = sqrt(sum_k deg_k * Im(<s|G_k|psi><psi|U^+(0)|s>)^2)  (gradients.py:23-124)."""
    ops, sim = backend.mps_ops, backend.simulator
    resolves_to_id = inverse_zero_ansatz == Circuit(2)
    circ_mps = ops.mps_from_circuit(circuit.copy(), return_preprocessed=True, sim=sim)
    start = starting_circuit if starting_circuit is not None else Circuit(circuit.num_qubits)
    if resolves_to_id:
        zero_overlap = ops.mps_dot(circ_mps, ops.mps_from_circuit(start.copy(), return_preprocessed=True, sim=sim),
                                   already_preprocessed=True)
    gradients = []
    for control, target in coupling_map:
        if not resolves_to_id:
            on_start = ops.mps_from_circuit(start.compose(inverse_zero_ansatz, [control, target]),
                                            return_preprocessed=True, sim=sim)
            zero_overlap = ops.mps_dot(circ_mps, on_start, already_preprocessed=True)
        total = 0
        for gen, deg in zip(generators, degeneracies):
            gen_mps = ops.mps_from_circuit(start.compose(gen, [control, target]), return_preprocessed=True, sim=sim)
            overlap = ops.mps_dot(gen_mps, circ_mps, already_preprocessed=True)
            total += (-1 * np.imag(overlap * zero_overlap)) ** 2 * deg
        gradients.append(np.sqrt(total))
    return gradients
