import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _cuda_device_count():
    try:
        import ctypes
        from adapt_aqc_b200.lib import load
        n = ctypes.c_int(0)
        if load().b200_device_count(ctypes.byref(n)) != 0:
            return 0
        return n.value
    except Exception:  # noqa: BLE001
        return 0


def pytest_collection_modifyitems(config, items):
    # On a box without a GPU the gpu-marked tests are skipped (they are selected with -m gpu on
    # the B200 box, where a missing device or library must FAIL, not skip).
    if _cuda_device_count() > 0:
        return
    if os.environ.get("B200AQC_REQUIRE_GPU") == "1":
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def emu():
    """CPU emulation of the sweep kernels' thread bodies (tests/emu)."""
    import ctypes
    emu_dir = os.path.join(ROOT, "tests", "emu")
    so = os.path.join(emu_dir, "libb200aqc_emu.so")
    srcs = [os.path.join(emu_dir, "emu.cu")] + [
        os.path.join(ROOT, "adapt-aqc_b200", "csrc", f) for f in ("sv_plan.cpp", "sv_plan.h", "sv_kernels.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", emu_dir, "-s"])
    lib = ctypes.CDLL(so)
    lib.emu_last_error.restype = ctypes.c_char_p
    dp = ctypes.POINTER(ctypes.c_double)
    lib.emu_sv_run.argtypes = [ctypes.c_int, dp, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                               ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]
    lib.emu_sv_run_inner2.argtypes = [ctypes.c_int, dp, dp, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, ctypes.POINTER(ctypes.c_int32)]
    lib.emu_sv_run_project.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]
    lib.emu_sv_run_embedded.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp,
                                        ctypes.POINTER(ctypes.c_int32)]
    lib.emu_set_variant.argtypes = [ctypes.c_int]
    return lib
