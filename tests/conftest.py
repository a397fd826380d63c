import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


@pytest.fixture(scope="session")
def emu_lib():
    """TEST INFRASTRUCTURE: the whole library (kernels + C-ABI layer) built with g++ on the emulation shim
    (tests/emu/emu_capi.cpp), loaded with the product's ctypes signatures.  The product never loads it."""
    import ctypes
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("needs g++")
    from datasketch_b200 import _native as nv
    emu = os.path.join(ROOT, "tests", "emu")
    out = os.path.join(emu, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libdsk_emu.so")
    csrc = os.path.join(ROOT, "datasketch_b200", "csrc")
    srcs = [os.path.join(emu, "emu_capi.cpp"), os.path.join(emu, "cuda_emu.h"), os.path.join(ROOT, "include", "dsk.h")] + \
           [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-std=c++17", "-O1", "-pthread", "-ffp-contract=off", "-DDSK_EMU", "-I" + emu, "-shared",
                        "-fPIC", "-o", so, srcs[0]], check=True)
    lib = ctypes.CDLL(so)
    for name, (res, args) in nv.SIGNATURES.items():
        fn = getattr(lib, name)               # the emulated build exports every C-ABI symbol too
        fn.restype, fn.argtypes = res, args
    return lib
