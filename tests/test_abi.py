"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol
include/amvcuda.h declares, and it refuses to run without a device (no CPU fallback)."""
import os
import re

import pytest

import amv_codec_tools_b200 as amv


@pytest.fixture(scope="module")
def lib():
    amv.build()
    return amv.load_library()


def test_header_symbols_exported(lib):
    hdr = open(amv.HEADER_PATH).read()
    declared = set(re.findall(r"AMV_API\s+[\w\s\*]+?\b(amv_\w+)\s*\(", hdr))
    assert declared == set(amv.EXPORTS), declared ^ set(amv.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_version_and_strings(lib):
    assert lib.amv_version() == 0x000100
    assert b"no CPU path" in lib.amv_strerror(-2)
    # update_qscale (mpegvideo_enc.c:143-148): quality 0 -> qmin; -qscale N (N*118) -> N
    assert lib.amv_qscale_from_quality(0, 2, 31) == 2
    for q in range(2, 32):
        assert lib.amv_qscale_from_quality(q * 118, 2, 31) == q
    assert lib.amv_qscale_from_quality(10 ** 6, 2, 31) == 31


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    with pytest.raises(amv.AmvError, match="no CPU path"):
        amv.AmvCuda()


def test_product_does_not_touch_oracle():
    """The product tree must not reference oracle/ or the host-emulation harness."""
    root = os.path.dirname(amv.PKG_DIR)
    for d, _, files in os.walk(amv.PKG_DIR):
        if "build" in d.split(os.sep):
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".c", ".py", "Makefile")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "oracle/" not in txt and "amvo_" not in txt and "host_emul" not in txt, os.path.join(d, f)
    assert os.path.isdir(os.path.join(root, "oracle"))
