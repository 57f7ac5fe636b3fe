"""adapt_aqc_b200.registration against stand-ins of ``adaptaqc`` and ``aqc_research`` (neither is installable here): the
modules below reproduce HOW the reference binds the third-party functions it calls around the backend -- a name
imported into the module (``from aqc_research.mps_operations import mps_from_circuit``:
approximate_compiler.py:20, gradients.py:14, aer_mps_backend.py:14-19) or an attribute looked up on the module
(``mpsops.partial_trace``: entanglement_measures.py:16,77; adapt_compiler.py:19,1129) -- and the tests check that after
``install()`` B200 arguments are served by the device ops at every one of those sites while everything else still reaches
the original function.  No GPU needed (the device ops are recorders)."""
import sys
import types

import pytest

from adapt_aqc_b200 import registration
from adapt_aqc_b200.backends import DeviceStatevector
from adapt_aqc_b200.mps_backend import B200MPSSimulator, DeviceMPSView


class Recorder:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def fn(*a, **k):
            self.calls.append(name)
            return ("device", name)
        return fn


@pytest.fixture
def fakes(monkeypatch):
    mods = {}

    def module(name):
        m = types.ModuleType(name)
        mods[name] = m
        monkeypatch.setitem(sys.modules, name, m)
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(mods[parent], child, m)
        return m

    aqc = module("aqc_research")
    mpsops = module("aqc_research.mps_operations")
    for fname in ("mps_from_circuit", "mps_dot", "mps_expectation", "extract_amplitude", "partial_trace", "check_mps"):
        def make(fname):
            def original(*a, **k):
                return ("original", fname)
            original.__name__ = fname
            return original
        setattr(mpsops, fname, make(fname))

    module("adaptaqc"); module("adaptaqc.utils"); module("adaptaqc.compilers"); module("adaptaqc.compilers.adapt")
    module("adaptaqc.backends")
    em = module("adaptaqc.utils.entanglement_measures")
    em.mpsops = mpsops                                             # import aqc_research.mps_operations as mpsops
    em.partial_trace = lambda statevector, a, b: ("original", "sv_partial_trace")
    em.rho_of = lambda sv, a, b: em.partial_trace(sv, a, b)        # module-internal call by global name (:75)
    em.mps_rho_of = lambda mps, a, b: em.mpsops.partial_trace(mps, [a, b], already_preprocessed=True)   # (:77)
    gr = module("adaptaqc.utils.gradients")
    gr.mps_from_circuit, gr.mps_dot = mpsops.mps_from_circuit, mpsops.mps_dot       # from ... import names
    gr.general_grad_of_pairs = lambda *a, **k: ("original", "general_grad_of_pairs")
    ac = module("adaptaqc.compilers.approximate_compiler")
    ac.mps_from_circuit, ac.check_mps = mpsops.mps_from_circuit, mpsops.check_mps
    adc = module("adaptaqc.compilers.adapt.adapt_compiler")
    adc.mpsops, adc.gr = mpsops, gr
    amb = module("adaptaqc.backends.aer_mps_backend")
    for fname in ("mps_from_circuit", "mps_dot", "mps_expectation", "extract_amplitude"):
        setattr(amb, fname, getattr(mpsops, fname))
    co = module("adaptaqc.utils.circuit_operations")
    co.calculate_overlap_between_circuits = lambda c1, c2, initial_state=None, qubit_subset=None: ("original", "overlap")
    registration.uninstall()
    yield types.SimpleNamespace(mpsops=mpsops, em=em, gr=gr, ac=ac, adc=adc, amb=amb, co=co)
    registration.uninstall()


def _b200_sim():
    sim = B200MPSSimulator.__new__(B200MPSSimulator)
    sim._ops = Recorder()
    return sim


def _device_view(sim):
    v = DeviceMPSView.__new__(DeviceMPSView)
    v._sim, v.handle = sim, None
    return v


def test_install_patches_every_bypass_site(fakes):
    patched = registration.install()
    for name in ("aqc_research.mps_operations.mps_from_circuit", "aqc_research.mps_operations.partial_trace",
                 "adaptaqc.utils.gradients.mps_from_circuit", "adaptaqc.utils.gradients.mps_dot",
                 "adaptaqc.compilers.approximate_compiler.mps_from_circuit",
                 "adaptaqc.backends.aer_mps_backend.extract_amplitude",
                 "adaptaqc.utils.entanglement_measures.partial_trace",
                 "adaptaqc.utils.gradients.general_grad_of_pairs",
                 "adaptaqc.utils.circuit_operations.calculate_overlap_between_circuits"):
        assert name in patched, (name, patched)
    assert registration.install() == []           # idempotent
    assert fakes.ac.check_mps("x") == ("original", "check_mps")      # pure host predicate: left alone


def test_b200_arguments_are_served_on_the_device_everywhere_else_falls_through(fakes):
    registration.install()
    sim = _b200_sim()
    view = _device_view(sim)
    # name-import sites
    assert fakes.ac.mps_from_circuit("qc", sim=sim) == ("device", "mps_from_circuit")
    assert fakes.gr.mps_from_circuit("qc", return_preprocessed=True, sim=sim) == ("device", "mps_from_circuit")
    assert fakes.gr.mps_dot(view, [1, 2], already_preprocessed=True) == ("device", "mps_dot")
    assert fakes.amb.extract_amplitude(view, 4, already_preprocessed=True) == ("device", "extract_amplitude")
    # attribute-lookup sites
    assert fakes.adc.mpsops.mps_from_circuit("qc", sim=sim) == ("device", "mps_from_circuit")
    assert fakes.em.mps_rho_of(view, 0, 1) == ("device", "partial_trace")
    assert fakes.mpsops.mps_expectation(view, "Z", 3, already_preprocessed=True) == ("device", "mps_expectation")
    assert sim._ops.calls.count("mps_from_circuit") == 3
    # anything that is not ours reaches the original
    assert fakes.ac.mps_from_circuit("qc", sim="an AerSimulator") == ("original", "mps_from_circuit")
    assert fakes.ac.mps_from_circuit("qc") == ("original", "mps_from_circuit")
    assert fakes.gr.mps_dot([1], [2]) == ("original", "mps_dot")
    assert fakes.em.mps_rho_of([1, 2, 3], 0, 1) == ("original", "partial_trace")


def test_statevector_partial_trace_and_gradients_dispatch(fakes):
    registration.install()
    sv = DeviceStatevector.__new__(DeviceStatevector)
    sv.partial_trace = lambda a, b: ("device", "sv_partial_trace", a, b)
    assert fakes.em.rho_of(sv, 2, 5) == ("device", "sv_partial_trace", 2, 5)
    assert fakes.em.rho_of("an Aer Statevector", 2, 5) == ("original", "sv_partial_trace")

    class Backend:
        def general_grad_of_pairs(self, *a):
            return ("device", "general_grad_of_pairs", len(a))
    assert fakes.adc.gr.general_grad_of_pairs("c", "u0", [], [], [(0, 1)], None, Backend()) == ("device", "general_grad_of_pairs", 6)
    assert fakes.adc.gr.general_grad_of_pairs("c", "u0", [], [], [(0, 1)], None, "AerMPSBackend") == ("original", "general_grad_of_pairs")
    assert fakes.adc.gr.general_grad_of_pairs("c", "u0", [], [], [(0, 1)]) == ("original", "general_grad_of_pairs")


def test_overlap_uses_the_last_active_backend_and_falls_back(fakes, monkeypatch):
    from adapt_aqc_b200 import backends
    registration.install()
    monkeypatch.setattr(backends, "_LAST_ACTIVE", None)
    assert fakes.co.calculate_overlap_between_circuits("c1", "c2") == ("original", "overlap")

    class B:
        def overlap_between_circuits(self, c1, c2):
            if c1 == "untranslatable":
                raise ValueError("instruction not supported")
            return 0.25
    b = B()
    monkeypatch.setattr(backends, "_LAST_ACTIVE", lambda: b)
    assert fakes.co.calculate_overlap_between_circuits("c1", "c2") == 0.25
    assert fakes.co.calculate_overlap_between_circuits("c1", "c2", qubit_subset=[0]) == ("original", "overlap")
    assert fakes.co.calculate_overlap_between_circuits("untranslatable", "c2") == ("original", "overlap")


def test_uninstall_restores_the_originals(fakes):
    registration.install()
    registration.uninstall()
    sim = _b200_sim()
    assert fakes.ac.mps_from_circuit("qc", sim=sim) == ("original", "mps_from_circuit")
    assert fakes.em.rho_of("x", 0, 1) == ("original", "sv_partial_trace")
    assert fakes.gr.general_grad_of_pairs() == ("original", "general_grad_of_pairs")
