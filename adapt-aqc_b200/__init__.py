"""adapt_aqc_b200 -- B200-native (sm_100a) simulation backends for ADAPT-AQC.

Drop-in replacements for ``AerSVBackend`` / ``AerMPSBackend`` of qiskit-community/adapt-aqc:
every cost evaluation and pair heuristic runs on the GPU through the C-ABI of libb200aqc.so
(include/b200aqc.h).  Importing the package does not need a GPU; creating a backend engine does,
and fails loudly if the CUDA library is missing -- there is no CPU path.

The package holds only what a maintainer of the reference would ship: the backends, their engines, the gate-stream
translator, the batched optimiser front end and the multi-GPU sharding.  The qiskit-free mirror of the compile loop
that tests / bench use to DRIVE the backends lives outside it (harness/).
"""
from . import gates  # noqa: F401
from .lib import B200Error, LIB_PATH  # noqa: F401

__all__ = ["B200Error", "LIB_PATH", "gates", "B200SVBackend", "B200MPSBackend", "DeviceStatevector", "install"]


def __getattr__(name):
    # backends are imported lazily so that `import adapt_aqc_b200` stays cheap
    if name in ("B200SVBackend", "DeviceStatevector"):
        from . import backends
        return getattr(backends, name)
    if name in ("B200MPSBackend",):
        from . import mps_backend
        return getattr(mps_backend, name)
    if name in ("B200CostMinimiser", "make_b200_minimiser"):
        from . import minimiser
        return getattr(minimiser, name)
    if name == "install":
        from .registration import install
        return install
    raise AttributeError(name)
