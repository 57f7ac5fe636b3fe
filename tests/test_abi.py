"""The C-ABI library loads and exports every symbol include/b200aqc.h declares; without a GPU the
product fails loudly instead of falling back to a CPU path."""
import ctypes
import os
import re

import numpy as np
import pytest

from adapt_aqc_b200 import lib as blib
from adapt_aqc_b200.gates import GATE_DTYPE, GateStream
from adapt_aqc_b200.sv_engine import plan_stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200aqc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = blib.load()
    declared = header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/b200aqc.h but not exported"
    assert sorted(blib.SYMBOLS) == declared
    assert L.b200_abi_version() == 1


def test_profile_class_table_matches_header():
    """b200_ctx_profile_read fills B200_PROF_CLASSES entries: the host-side table (shared by both engines) has the same length
    and order as the header's enum."""
    from adapt_aqc_b200.sv_engine import SVEngine
    text = open(os.path.join(ROOT, "include", "b200aqc.h")).read()
    count = int(re.search(r"#define\s+B200_PROF_CLASSES\s+(\d+)", text).group(1))
    enum = sorted((int(v), k.lower()) for k, v in re.findall(r"B200_PROF_([A-Z0-9_]+)\s*=\s*(\d+)", text))
    assert [v for v, _ in enum] == list(range(count))
    assert tuple(k for _, k in enum) == SVEngine.PROF_CLASSES


def test_gate_record_layout_matches_header():
    assert GATE_DTYPE.itemsize == 40
    assert [GATE_DTYPE.fields[k][1] for k in ("op", "q0", "q1", "aux", "p")] == [0, 4, 8, 12, 16]


def test_opcode_tables_agree_with_header_and_oracle():
    from adapt_aqc_b200 import gates as G
    from oracle import sv_oracle as orc
    text = open(os.path.join(ROOT, "include", "b200aqc.h")).read()
    enum = dict((k.lower(), int(v)) for k, v in re.findall(r"B200_OP_([A-Z0-9]+)\s*=\s*(\d+)", text))
    for name, (code, _, _) in G.GATE_TABLE.items():
        hname = {"i": "id", "p": "u1", "u": "u3"}.get(name, name)
        assert enum[hname] == code
        assert orc.OPCODES[hname] == code
    assert enum["mat1"] == G.OP_MAT1 == orc.OPCODES["mat1"]
    assert enum["mat2"] == G.OP_MAT2 == orc.OPCODES["mat2"]


def test_planner_runs_without_gpu_and_rejects_bad_gates():
    gs = GateStream.from_gates([("h", [0], []), ("cx", [0, 13], []), ("rz", [13], [0.3])])
    sweeps, rounds, ops, small = plan_stats(14, gs)
    assert small == 0 and sweeps >= 1 and rounds >= sweeps and ops >= 2
    assert plan_stats(5, gs.__class__.from_gates([("h", [0], [])]))[3] == 1
    bad = GateStream.from_gates([("cx", [0, 0], [])])
    with pytest.raises(blib.B200Error, match="q1"):
        plan_stats(4, bad)
    bad2 = GateStream.from_gates([("h", [7], [])])
    with pytest.raises(blib.B200Error, match="out of range"):
        plan_stats(4, bad2)


def test_no_silent_cpu_fallback():
    """Creating a context needs a B200; on any other box it must raise, never emulate."""
    L = blib.load()
    n = ctypes.c_int(0)
    rc = L.b200_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    from adapt_aqc_b200.sv_engine import SVEngine
    with pytest.raises(blib.B200Error):
        SVEngine(3)
    from adapt_aqc_b200.backends import B200SVBackend
    from harness.circuit import Circuit
    from harness.compiler import AdaptCompiler
    qc = Circuit(2); qc.h(0)
    with pytest.raises(blib.B200Error):
        AdaptCompiler(qc, backend=B200SVBackend()).evaluate_cost()


def test_product_never_imports_the_harness_or_the_oracle():
    """The package holds product code only: the oracle (CPU restatement of the reference's arithmetic) and the harness
    (qiskit-free mirror of the reference's compile loop) are test infrastructure and must stay outside it."""
    pkg = os.path.join(ROOT, "adapt-aqc_b200")
    assert not any(os.path.exists(os.path.join(pkg, f)) for f in ("compiler.py", "circuit.py", "measures.py", "gradients.py"))
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "sv_oracle" not in text and "libsv_oracle" not in text, f
                assert "import harness" not in text and "from harness" not in text, f
