"""CPU ORACLE -- statevector path (test infrastructure, NOT product code).

ctypes front end of ``oracle/sv_oracle.c`` plus Python restatements of the
reference methods that sit directly on top of the simulator.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product package never does.

Parity status: the arithmetic lives in qiskit-aer ~=0.16.0 / qiskit ~=1.3.1
(neither vendored in /root/reference nor installable offline), so the oracle is
pinned on the reference's own known-answer tests (tests/test_oracle_kats.py),
not on outputs of the reference run here.

Reference call sites restated here (paths relative to /root/reference):
  evaluate_circuit                    adaptaqc/backends/aer_sv_backend.py:37-47
  evaluate_global_cost                adaptaqc/backends/aer_sv_backend.py:23-30
  measure_qubit_expectation_values    adaptaqc/backends/aer_sv_backend.py:49-59
  evaluate_local_cost                 adaptaqc/backends/aer_sv_backend.py:32-35
  partial_trace(sv, a, b)             adaptaqc/utils/entanglement_measures.py:325-340
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsv_oracle.so")

# Wire format shared with include/b200aqc.h (b200_gate), 40 bytes.
GATE_DTYPE = np.dtype(
    [("op", "<i4"), ("q0", "<i4"), ("q1", "<i4"), ("aux", "<i4"), ("p", "<f8", (3,))]
)

OPCODES = {
    "id": 0, "x": 1, "y": 2, "z": 3, "h": 4, "rx": 5, "ry": 6, "rz": 7,
    "u1": 8, "p": 8, "u2": 9, "u3": 10, "u": 10, "cx": 11, "cz": 12,
    "mat1": 13, "mat2": 14, "s": 15, "sdg": 16, "t": 17, "tdg": 18, "sx": 19,
    "swap": 20,
}


def build(force=False):
    """Compile libsv_oracle.so with the committed Makefile (gcc + OpenMP)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "sv_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        L.orc_sv_apply.argtypes = [ctypes.c_int, dp, ctypes.c_void_p, ctypes.c_int, dp]
        L.orc_sv_apply.restype = ctypes.c_int
        L.orc_sv_simulate.argtypes = L.orc_sv_apply.argtypes
        L.orc_sv_simulate.restype = ctypes.c_int
        L.orc_probabilities.argtypes = [ctypes.c_int, dp, ctypes.c_int, dp]
        L.orc_probabilities.restype = None
        L.orc_partial_trace_pair.argtypes = [ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, dp]
        L.orc_partial_trace_pair.restype = None
        L.orc_vdot.argtypes = [ctypes.c_int, dp, dp, dp]
        L.orc_vdot.restype = None
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_set_num_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def pack_gates(gates):
    """[(name, qubits, params)] -> (records, mats).  ``mat1``/``mat2`` take the
    matrix (2x2 / 4x4, little-endian: first qubit = least significant) as params."""
    rec = np.zeros(len(gates), dtype=GATE_DTYPE)
    mats = []
    off = 0
    for k, (name, qubits, params) in enumerate(gates):
        rec[k]["op"] = OPCODES[name]
        rec[k]["q0"] = qubits[0]
        rec[k]["q1"] = qubits[1] if len(qubits) > 1 else -1
        if name in ("mat1", "mat2"):
            m = np.asarray(params, dtype=np.complex128).reshape(-1)
            rec[k]["aux"] = off
            mats.append(m.view(np.float64))
            off += 2 * m.size
        else:
            for j, v in enumerate(params):
                rec[k]["p"][j] = float(v)
    mats = np.concatenate(mats) if mats else np.zeros(1)
    return rec, np.ascontiguousarray(mats, dtype=np.float64)


def _as_packed(gates):
    if isinstance(gates, tuple) and len(gates) == 2 and isinstance(gates[0], np.ndarray):
        return gates
    return pack_gates(gates)


def evaluate_circuit(num_qubits, gates):
    """All gates from |0..0>; returns the little-endian complex128 statevector."""
    rec, mats = _as_packed(gates)
    psi = np.empty(1 << num_qubits, dtype=np.complex128)
    rc = lib().orc_sv_simulate(
        num_qubits, _dptr(psi.view(np.float64)), rec.ctypes.data, len(rec), _dptr(mats)
    )
    if rc != 0:
        raise ValueError(f"oracle: bad gate at index {-rc - 1}")
    return psi


def apply_gates(psi, gates):
    rec, mats = _as_packed(gates)
    n = int(np.log2(psi.size))
    psi = np.ascontiguousarray(psi, dtype=np.complex128).copy()
    rc = lib().orc_sv_apply(n, _dptr(psi.view(np.float64)), rec.ctypes.data, len(rec), _dptr(mats))
    if rc != 0:
        raise ValueError(f"oracle: bad gate at index {-rc - 1}")
    return psi


def evaluate_global_cost(sv):
    return 1 - (np.absolute(sv[0])) ** 2


def probabilities(sv, qubit):
    n = int(np.log2(sv.size))
    out = np.zeros(2)
    lib().orc_probabilities(n, _dptr(sv.view(np.float64)), qubit, _dptr(out))
    return out


def measure_qubit_expectation_values(sv):
    n = int(np.log2(sv.size))
    vals = []
    for i in range(n):
        p0, p1 = probabilities(sv, i)
        vals.append(p0 - p1)
    return vals


def evaluate_local_cost(sv):
    return 0.5 * (1 - np.mean(measure_qubit_expectation_values(sv)))


def partial_trace(sv, a, b):
    n = int(np.log2(sv.size))
    out = np.zeros(32)
    lib().orc_partial_trace_pair(n, _dptr(sv.view(np.float64)), a, b, _dptr(out))
    return out.view(np.complex128).reshape(4, 4)


def vdot(l, r):
    n = int(np.log2(l.size))
    out = np.zeros(2)
    lib().orc_vdot(n, _dptr(l.view(np.float64)), _dptr(r.view(np.float64)), _dptr(out))
    return complex(out[0], out[1])


def num_threads():
    return lib().orc_num_threads()
