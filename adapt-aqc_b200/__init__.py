"""adapt_aqc_b200 -- B200-native (sm_100a) simulation backends for ADAPT-AQC.

Drop-in replacements for ``AerSVBackend`` / ``AerMPSBackend`` of qiskit-community/adapt-aqc:
every cost evaluation and pair heuristic runs on the GPU through the C-ABI of libb200aqc.so
(include/b200aqc.h).  Importing the package does not need a GPU; creating a backend engine does,
and fails loudly if the CUDA library is missing -- there is no CPU path.
"""
from . import gates, measures  # noqa: F401
from .circuit import Circuit, CircuitInstruction, Gate  # noqa: F401
from .lib import B200Error, LIB_PATH  # noqa: F401

__all__ = ["Circuit", "Gate", "CircuitInstruction", "B200Error", "LIB_PATH", "gates", "measures"]


def __getattr__(name):
    # backends / compiler are imported lazily so that `import adapt_aqc_b200` stays cheap
    if name in ("B200SVBackend", "DeviceStatevector"):
        from . import backends
        return getattr(backends, name)
    if name in ("B200MPSBackend",):
        from . import mps_backend
        return getattr(mps_backend, name)
    if name in ("AdaptCompiler", "AdaptConfig", "AdaptResult"):
        from . import compiler
        return getattr(compiler, name)
    if name in ("CostMinimiser", "B200CostMinimiser"):
        from . import minimiser
        return getattr(minimiser, name)
    raise AttributeError(name)
