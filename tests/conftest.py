import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the compiled reference oracle/_ref/libamvref.so")


# Developer switch, never set by the driver: AMV_EMUL=1 points the GPU parity tests at tests/host_emul/simt's
# libamvcuda_emul.so (the kernel sources compiled as C++ on the CPU SIMT emulator), so kernel changes can be checked
# in a container without a GPU before GPU time is spent.  The product library and package are untouched by this.
EMUL = os.environ.get("AMV_EMUL") == "1"
if EMUL:
    import subprocess
    # AMV_EMUL_ASAN_SO: the AddressSanitizer build of the same library (test_simt_emul.py runs a child process that way)
    _so = os.environ.get("AMV_EMUL_ASAN_SO") or subprocess.check_output(
        [os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul", "simt", "build.sh")], text=True).strip().splitlines()[-1]
    import amv_codec_tools_b200 as _amv
    _amv.AmvCuda.__init__.__defaults__ = (-1, None, _so)
    _amv.load_library.__defaults__ = (_so,)


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass on a box without a device
    if EMUL:
        return
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
