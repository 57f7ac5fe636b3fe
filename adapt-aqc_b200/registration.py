"""Registration hook for a real ``adaptaqc`` installation.

Subclassing gets a B200 backend through every ``isinstance(backend, AerSVBackend / AerMPSBackend)`` gate of the
reference, and the four ``AQCBackend`` methods are overridden.  But the reference also reaches AROUND the backend into
third-party code, through module-level names (SURVEY 8b "de-facto interface"):

  statevector  adaptaqc/utils/entanglement_measures.py:71-75,325-340   partial_trace(statevector, a, b) -> qiskit.quantum_info
               adaptaqc/compilers/adapt/adapt_compiler.py:447-452      co.calculate_overlap_between_circuits(...) -> qiskit Statevector
  MPS          adaptaqc/compilers/approximate_compiler.py:20,133,198,230   mps_from_circuit      (name imported into the module)
               adaptaqc/compilers/adapt/adapt_compiler.py:19,1129          mpsops.mps_from_circuit (attribute of aqc_research.mps_operations)
               adaptaqc/utils/entanglement_measures.py:16,77               mpsops.partial_trace
               adaptaqc/utils/gradients.py:14,60-111                       mps_from_circuit, mps_dot (names imported into the module)
               adaptaqc/utils/utilityfunctions.py:18,197,202               mpsop.mps_from_circuit / mps_expectation
               adaptaqc/backends/aer_mps_backend.py:14-19                  names imported into the module (methods overridden anyway)
               adaptaqc/utils/gradients.py:23-124                          general_grad_of_pairs (called as gr.general_grad_of_pairs)

``install()`` wraps each of those names with a dispatcher: arguments that belong to a B200 backend (a
``B200MPSSimulator`` passed as ``sim=``, a ``DeviceMPSView`` / ``DeviceStatevector`` state, a ``B200MPSBackend``) are
served on the device, everything else falls through to the original function untouched -- so Aer backends keep
working in the same process.  Without this hook a B200 MPS backend still gives correct results (the simulator facade
answers ``sim.run(...)``), but every ``mps_from_circuit`` round-trips the whole MPS through Python lists.

A no-op returning False when ``adaptaqc`` / ``aqc_research`` are not importable (the build image); the wiring itself is
exercised against stand-in modules in tests/test_registration.py.
"""
import importlib
import sys

# modules that did ``from aqc_research.mps_operations import <name>`` (the name is a global of THAT module)
_NAME_IMPORT_SITES = {
    "adaptaqc.compilers.approximate_compiler": ("mps_from_circuit",),
    "adaptaqc.backends.aer_mps_backend": ("mps_from_circuit", "mps_dot", "mps_expectation", "extract_amplitude"),
    "adaptaqc.utils.gradients": ("mps_from_circuit", "mps_dot"),
}
_MPS_FUNCTIONS = ("mps_from_circuit", "mps_dot", "mps_expectation", "extract_amplitude", "partial_trace")
_installed = {}


def _try_import(name):
    try:
        return importlib.import_module(name)
    except Exception:  # noqa: BLE001 - absent or broken third-party package
        return None


def _mps_dispatchers(original):
    """{name: dispatcher} for the aqc_research.mps_operations functions (original: the module being wrapped)."""
    from .mps_backend import B200MPSSimulator, DeviceMPSView
    from .mps_engine import DeviceMPS

    def on_device(x):
        return isinstance(x, (DeviceMPSView, DeviceMPS))

    def ops_of(*states):
        for s in states:
            if isinstance(s, DeviceMPSView):
                return s._sim.ops
        raise TypeError("no device-resident MPS among the arguments")

    def mps_from_circuit(qc, *args, **kwargs):
        sim = kwargs.get("sim")
        if isinstance(sim, B200MPSSimulator):
            return sim.ops.mps_from_circuit(qc, *args, **kwargs)
        return original["mps_from_circuit"](qc, *args, **kwargs)

    def mps_dot(mps1, mps2, *args, **kwargs):
        if on_device(mps1) or on_device(mps2):
            return ops_of(mps1, mps2).mps_dot(mps1, mps2, *args, **kwargs)
        return original["mps_dot"](mps1, mps2, *args, **kwargs)

    def mps_expectation(mps, *args, **kwargs):
        if on_device(mps):
            return ops_of(mps).mps_expectation(mps, *args, **kwargs)
        return original["mps_expectation"](mps, *args, **kwargs)

    def extract_amplitude(mps, *args, **kwargs):
        if on_device(mps):
            return ops_of(mps).extract_amplitude(mps, *args, **kwargs)
        return original["extract_amplitude"](mps, *args, **kwargs)

    def partial_trace(mps, *args, **kwargs):
        if on_device(mps):
            return ops_of(mps).partial_trace(mps, *args, **kwargs)
        return original["partial_trace"](mps, *args, **kwargs)

    return {"mps_from_circuit": mps_from_circuit, "mps_dot": mps_dot, "mps_expectation": mps_expectation,
            "extract_amplitude": extract_amplitude, "partial_trace": partial_trace}


def install_mps():
    """Wrap the aqc_research.mps_operations names the reference calls directly.  Returns the list of patched sites."""
    mpsops = _try_import("aqc_research.mps_operations")
    if mpsops is None:
        return []
    patched = []
    if "mps" not in _installed:
        original = {n: getattr(mpsops, n) for n in _MPS_FUNCTIONS if hasattr(mpsops, n)}
        _installed["mps"] = (original, _mps_dispatchers(original))
    original, disp = _installed["mps"]
    for name, fn in disp.items():
        if name in original and getattr(mpsops, name) is not fn:
            setattr(mpsops, name, fn)              # attribute-lookup sites: mpsops.f(...) / mpsop.f(...)
            patched.append(f"aqc_research.mps_operations.{name}")
    for modname, names in _NAME_IMPORT_SITES.items():
        mod = sys.modules.get(modname) or _try_import(modname)
        if mod is None:
            continue
        for name in names:
            if name in original and getattr(mod, name, None) is original[name]:
                setattr(mod, name, disp[name])     # name-import sites: the module's own global
                patched.append(f"{modname}.{name}")
    return patched


def install_gradients():
    """gr.general_grad_of_pairs (adapt_compiler.py:839-856): a B200MPSBackend answers all (pair, generator) overlaps from
    one batched launch (B200MPSBackend.general_grad_of_pairs) instead of P x (#generators + 1) simulator runs."""
    gr = _try_import("adaptaqc.utils.gradients")
    if gr is None or not hasattr(gr, "general_grad_of_pairs"):
        return []
    if "gradients" not in _installed:
        original = gr.general_grad_of_pairs

        def general_grad_of_pairs(circuit, inverse_zero_ansatz, generators, degeneracies, coupling_map,
                                  starting_circuit=None, backend=None, **kwargs):
            if backend is not None and hasattr(backend, "general_grad_of_pairs") and not kwargs:
                return backend.general_grad_of_pairs(circuit, inverse_zero_ansatz, generators, degeneracies, coupling_map,
                                                     starting_circuit)
            if backend is None:
                return original(circuit, inverse_zero_ansatz, generators, degeneracies, coupling_map, starting_circuit, **kwargs)
            return original(circuit, inverse_zero_ansatz, generators, degeneracies, coupling_map, starting_circuit, backend,
                            **kwargs)

        _installed["gradients"] = (original, general_grad_of_pairs)
    if gr.general_grad_of_pairs is not _installed["gradients"][1]:
        gr.general_grad_of_pairs = _installed["gradients"][1]
        return ["adaptaqc.utils.gradients.general_grad_of_pairs"]
    return []


def install_sv():
    """partial_trace(statevector, a, b) and calculate_overlap_between_circuits for device-resident statevectors."""
    rem = _try_import("adaptaqc.utils.entanglement_measures")
    patched = []
    if rem is not None and hasattr(rem, "partial_trace"):
        from .backends import DeviceStatevector
        if "sv_partial_trace" not in _installed:
            original_pt = rem.partial_trace

            def partial_trace(statevector, qubit_1, qubit_2):
                if isinstance(statevector, DeviceStatevector):
                    return statevector.partial_trace(qubit_1, qubit_2)      # all candidate pairs from one batch of RDM passes
                return original_pt(statevector, qubit_1, qubit_2)

            _installed["sv_partial_trace"] = (original_pt, partial_trace)
        if rem.partial_trace is not _installed["sv_partial_trace"][1]:
            rem.partial_trace = _installed["sv_partial_trace"][1]
            patched.append("adaptaqc.utils.entanglement_measures.partial_trace")
    co = _try_import("adaptaqc.utils.circuit_operations")
    if co is not None and hasattr(co, "calculate_overlap_between_circuits"):
        from . import backends as _b
        if "overlap" not in _installed:
            original_ov = co.calculate_overlap_between_circuits

            def calculate_overlap_between_circuits(circuit1, circuit2, initial_state=None, qubit_subset=None):
                backend = _b.last_active_backend()
                if backend is None or initial_state is not None or qubit_subset:
                    return original_ov(circuit1, circuit2, initial_state, qubit_subset)
                try:
                    return backend.overlap_between_circuits(circuit1, circuit2)   # two device simulations + one inner product
                except Exception:  # noqa: BLE001 - e.g. an instruction the gate translator does not know
                    return original_ov(circuit1, circuit2, initial_state, qubit_subset)

            _installed["overlap"] = (original_ov, calculate_overlap_between_circuits)
        fn = _installed["overlap"][1]
        for mod in (co, sys.modules.get("adaptaqc.utils.circuit_operations.circuit_operations_full_circuit")):
            if mod is not None and getattr(mod, "calculate_overlap_between_circuits", None) is _installed["overlap"][0]:
                mod.calculate_overlap_between_circuits = fn
                patched.append(f"{mod.__name__}.calculate_overlap_between_circuits")
    return patched


def install():
    """Patch every bypass site listed in the module docstring.  Idempotent.  Returns the list of patched names
    (empty -- falsy -- when the reference package is not importable)."""
    if _try_import("adaptaqc") is None:
        return []
    return install_sv() + install_mps() + install_gradients()


def uninstall():
    """Restore the original functions (tests)."""
    for key, modname, attr in (("sv_partial_trace", "adaptaqc.utils.entanglement_measures", "partial_trace"),
                               ("gradients", "adaptaqc.utils.gradients", "general_grad_of_pairs")):
        mod = sys.modules.get(modname)
        if key in _installed and mod is not None and getattr(mod, attr, None) is _installed[key][1]:
            setattr(mod, attr, _installed[key][0])
    if "overlap" in _installed:
        for modname in ("adaptaqc.utils.circuit_operations", "adaptaqc.utils.circuit_operations.circuit_operations_full_circuit"):
            mod = sys.modules.get(modname)
            if mod is not None and getattr(mod, "calculate_overlap_between_circuits", None) is _installed["overlap"][1]:
                mod.calculate_overlap_between_circuits = _installed["overlap"][0]
    if "mps" in _installed:
        original, disp = _installed["mps"]
        for modname in ["aqc_research.mps_operations"] + list(_NAME_IMPORT_SITES):
            mod = sys.modules.get(modname)
            for name, fn in disp.items():
                if mod is not None and getattr(mod, name, None) is fn:
                    setattr(mod, name, original[name])
    _installed.clear()
