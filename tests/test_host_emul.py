"""CPU checks of the kernels' __host__ __device__ lane logic (tests/host_emul/emul.cu: the same
headers the CUDA kernels compile, run one emulated lane at a time) against the oracle.  This is
NOT a product path; it lets the Huffman LUTs, BitReader, decode_block, the self-synchronisation
hand-over arithmetic and the register IDCT/FDCT/quantiser be checked without a GPU."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle_lib import Oracle, synth_frames, chroma_dims

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "emul.cu")
SO = os.path.join(HERE, "host_emul", "libemul.so")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

pytestmark = pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def emul():
    csrc = os.path.join(os.path.dirname(HERE), "amv-codec-tools_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc))
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(newest, os.path.getmtime(SRC)):
        subprocess.check_call([NVCC, "-O2", "-std=c++17", "-arch=sm_100a", "--expt-relaxed-constexpr", "-w",
                               "-Xcompiler", "-fPIC", "-shared", "-o", SO, SRC])
    return C.CDLL(SO)


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def test_vlc_table_size(emul):
    n = emul.emul_vlc_entries()
    assert 4 * 1024 < n <= 4 * 1024 + 20 * 64


def test_enc_huff_entries(emul, oracle):
    base = {0: 0, 1: 16, 2: 32, 3: 288}
    for t in range(4):
        ln, cd = oracle.huff(t)
        for s in range(16 if t < 2 else 256):
            e = emul.emul_enc_huff(base[t] + s)
            assert (e & 31, e >> 5) == (int(ln[s]), int(cd[s]))


def test_idct_registers(emul, oracle):
    rng = np.random.default_rng(31)
    sparse = np.zeros((4000, 64), np.int16)
    for i in range(sparse.shape[0]):
        k = rng.integers(0, 8)
        sparse[i, rng.integers(0, 64, k)] = rng.integers(-2000, 2000, k)
        sparse[i, 0] = rng.integers(-3000, 6000)
    extreme = rng.choice(np.array([-32768, -1, 0, 1, 32767], np.int16), (2000, 64))
    dense = rng.integers(-32768, 32768, (2000, 64)).astype(np.int16)
    tworows = dense.copy()                       # only coefficient rows 0 and 1: the short transform of k_idct
    tworows[:, 16:] = 0
    tworows[::3, 8:] = 0
    blocks = np.ascontiguousarray(np.concatenate([sparse, extreme, dense, tworows]))
    out = np.zeros((blocks.shape[0], 64), np.uint8)
    emul.emul_idct(_p(blocks), blocks.shape[0], _p(out))
    assert np.array_equal(out, oracle.idct_put(blocks))


@pytest.mark.parametrize("qscale", [2, 3, 7, 31])
def test_fdct_quant_registers(emul, oracle, qscale):
    rng = np.random.default_rng(32)
    # random pixels plus the extremal patterns of every 2-D basis function (max |coefficient|)
    n = np.arange(8)
    basis = np.cos((2 * n[None, :] + 1) * n[:, None] * np.pi / 16)
    ext = []
    for u in range(8):
        for v in range(8):
            pat = np.outer(basis[u], basis[v])
            ext.append((pat > 0) * 255)
            ext.append((pat < 0) * 255)
    blocks = np.concatenate([rng.integers(0, 256, (6000, 64)), np.array(ext).reshape(-1, 64),
                             np.full((1, 64), 255), np.zeros((1, 64), int)]).astype(np.int16)
    blocks = np.ascontiguousarray(blocks)
    q = np.zeros_like(blocks)
    fd = np.zeros_like(blocks)
    emul.emul_fdct_quant(_p(blocks), blocks.shape[0], qscale, _p(q), _p(fd))
    want_fd = oracle.fdct(blocks)
    assert np.array_equal(fd, want_fd)            # 32-bit wrap arithmetic == the reference's 64-bit for pixel input
    qm = oracle.enc_qmat(qscale)
    want = want_fd.copy()
    for i in range(want.shape[0]):
        oracle.lib.amvo_quantize(_p(want[i]), _p(qm))
    assert np.array_equal(q, want)


def test_division_by_multiply(emul):
    """div_by_magic (block index -> frame / block row in k_idct16): exact for every 32-bit dividend; checked densely at both
    ends of the range and with a coarse stride in between, for the divisors the geometries produce and the edge cases."""
    for d in (1, 2, 3, 5, 6, 7, 10, 20, 26, 40, 45, 80, 160, 429, 1800, 3300, 21600, 65535, 65536, 65537, (1 << 31) - 1, 1 << 31,
              (1 << 32) - 1):
        assert emul.emul_div_magic_check(C.c_uint32(d), C.c_uint32(0), C.c_uint32(200000), C.c_uint32(1)) == 0
        assert emul.emul_div_magic_check(C.c_uint32(d), C.c_uint32((1 << 32) - 200000), C.c_uint32(200000), C.c_uint32(1)) == 0
        assert emul.emul_div_magic_check(C.c_uint32(d), C.c_uint32(12345), C.c_uint32(400000), C.c_uint32(10007)) == 0


@pytest.mark.parametrize("form", [0, 1, 2, 3, 6, 7])
def test_fdct_regrouped_forms(emul, oracle, form):
    """fdct_block_px: the row pass as dot products on the packed pixel bytes and the column pass with its odd half (and
    outputs 2 / 6) written out are the same integer linear maps as the factorised pass of jfdctint.c:184-341."""
    rng = np.random.default_rng(34)
    n = np.arange(8)
    basis = np.cos((2 * n[None, :] + 1) * n[:, None] * np.pi / 16)
    ext = []
    for u in range(8):
        for v in range(8):
            pat = np.outer(basis[u], basis[v])
            ext.append((pat > 0) * 255)
            ext.append((pat < 0) * 255)
    blocks = np.concatenate([rng.integers(0, 256, (6000, 64)), rng.choice([0, 255], (2000, 64)), np.array(ext).reshape(-1, 64),
                             np.full((1, 64), 255), np.zeros((1, 64), int)]).astype(np.int16)
    blocks = np.ascontiguousarray(blocks)
    fd = np.zeros_like(blocks)
    assert emul.emul_fdct_form(_p(blocks), blocks.shape[0], form, _p(fd)) == 0
    assert np.array_equal(fd, oracle.fdct(blocks))


@pytest.mark.parametrize("w,h,kind", [(160, 120, "sinus"), (64, 48, "noise"), (48, 40, "edges"), (32, 32, "flat"),
                                      (208, 176, "sinus"), (72, 24, "sinus")])
@pytest.mark.parametrize("log2p", [0, 1, 3, 5])
def test_lane_decode_matches_oracle(emul, oracle, w, h, kind, log2p):
    y, u, v = synth_frames(2, w, h, seed=33, kind=kind)
    pk, off, sz = oracle.encode_frames(y, u, v, w, h, 2)
    oy, ou, ov, st = oracle.decode_frames(pk, off, sz, w, h)
    cw, ch = chroma_dims(w, h)
    for i in range(2):
        p = np.ascontiguousarray(pk[int(off[i]): int(off[i]) + int(sz[i])])
        ey, eu, ev = np.zeros((h, w), np.uint8), np.zeros((ch, cw), np.uint8), np.zeros((ch, cw), np.uint8)
        rounds = C.c_int(0)
        s = emul.emul_decode_frame(_p(p), len(p), w, h, _p(ey), _p(eu), _p(ev), log2p, C.byref(rounds))
        assert s == 0 and st[i] == 0
        assert np.array_equal(ey, oy[i]) and np.array_equal(eu, ou[i]) and np.array_equal(ev, ov[i])
        if log2p:
            assert 1 <= rounds.value <= (1 << log2p) + 1


def test_adpcm_quotient_by_reciprocal_is_exact():
    """k_adpcm_encode_async finds min(7, |delta| * 4 / step) (adpcm_ima_compress_sample, adpcm.c:219-227) as
    min(7, umulhi(|delta|, ceil(2^34 / step))): every step of the table (read from the product's amv_tables.cuh) x every
    difference two int16 samples can have."""
    import re
    src = open(os.path.join(os.path.dirname(HERE), "amv-codec-tools_b200", "csrc", "amv_tables.cuh")).read()
    body = re.search(r"kImaStep\[89\]\s*=\s*\{([^}]*)\}", src).group(1)
    steps = [int(t) for t in re.findall(r"\d+", body)]
    assert len(steps) == 89 and steps[0] == 7 and steps[-1] == 32767
    x = np.arange(65536, dtype=np.uint64)
    for s in steps:
        m = ((1 << 34) + s - 1) // s
        assert m < (1 << 32)
        assert np.array_equal(np.minimum(7, (4 * x) // s), np.minimum(7, (x * np.uint64(m)) >> np.uint64(32))), s
