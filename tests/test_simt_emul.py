"""The WHOLE kernels on the CPU: tests/host_emul/simt compiles the kernel sources of libamvcuda (csrc/*.cu) as plain C++
against a stand-in CUDA runtime and runs them one CTA at a time, every CUDA thread a fiber that yields at warp / CTA
collectives (shuffles, ballots, __syncwarp, __syncthreads; bulk copies and cp.async complete at once).  What runs here
is the same source nvcc builds -- the launch geometry, the shared-memory bit packing, the token passes, the host-side
copy paths of amv_api.cu -- checked against the oracle without a GPU.  TEST-ONLY: the emulated library is built under
tests/, never shipped, never looked for by the package (whose AmvCuda raises without the real library and a device).

A curated subset for the CPU suite; `AMV_EMUL=1 pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu`
runs the GPU parity tests themselves against the emulated library (developer switch, see conftest.py)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import amv_codec_tools_b200 as amv
from oracle_lib import Oracle, offsets_of, pack, synth_frames, synth_pcm

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "host_emul", "simt", "build.sh")

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")


@pytest.fixture(scope="module")
def emu():
    # AMV_EMUL_ASAN_SO: the sanitizer build, when this file runs as the child of the AddressSanitizer test below
    so = os.environ.get("AMV_EMUL_ASAN_SO") or subprocess.check_output([BUILD], text=True).strip().splitlines()[-1]
    c = amv.AmvCuda(device=0, lib_path=so)
    yield c
    c.close()


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.mark.parametrize("w,h,kind", [(160, 120, "sinus"), (72, 24, "noise"), (48, 40, "flat"), (320, 240, "sinus"), (208, 176, "edges")])
def test_encode_kernels_byte_identical(emu, oracle, w, h, kind):
    n = 2 if w * h > 40000 else 5
    y, u, v = synth_frames(n, w, h, seed=11, kind=kind)
    wpk, woff, wsz = oracle.encode_frames(y, u, v, w, h, 2)
    # k_encode16v2 with the regrouped transform (4, 8) and the factorised one (2, 3), k_encode16 (each + k_encode for
    # handed-back frames), k_encode alone
    for form in (4, 8, 3, 2, 1, 0):
        emu.set_option("encode_rounds", form)
        pk, off, sz, st = emu.encode_frames(y, u, v)
        assert (st == 0).all() and np.array_equal(sz, wsz) and np.array_equal(pk, wpk), "encode_rounds=%d" % form
    emu.set_option("encode_rounds", 4)


def test_encode_qscales_and_packed_vs_slots(emu, oracle):
    w, h, n = 160, 120, 6
    y, u, v = synth_frames(n, w, h, seed=12)
    q = np.array([2, 3, 5, 10, 20, 31], np.int32)
    want = [oracle.encode_frames(y[i:i + 1], u[i:i + 1], v[i:i + 1], w, h, int(q[i]))[0] for i in range(n)]
    pk, off, sz, st = emu.encode_frames(y, u, v, qscale=q)
    assert (st == 0).all()
    for i in range(n):
        assert np.array_equal(pk[int(off[i]): int(off[i]) + int(sz[i])], want[i])
    pk2, off2, sz2, st2 = emu.encode_frames(y, u, v, qscale=q, layout=amv.LAYOUT_SLOTS, pkt_cap=20000)
    assert (st2 == 0).all() and np.array_equal(sz2, sz)
    for i in range(n):
        assert np.array_equal(pk2[int(off2[i]): int(off2[i]) + int(sz2[i])], want[i])


@pytest.mark.parametrize("log2p", [0, 2, 5])
@pytest.mark.parametrize("token_pass", [2, 1, 0])
def test_decode_kernels_identical(emu, oracle, token_pass, log2p):
    cases = [(160, 120, "sinus", 3), (72, 24, "noise", 4), (48, 40, "flat", 5), (16, 16, "noise", 3), (32, 16, "flat", 2),
             (64, 48, "edges", 7), (208, 176, "edges", 2)]      # tiny scans on many lanes: empty subsequences, blocks longer than one
    emu.set_option("decode_token_pass", token_pass)
    emu.set_option("decode_log2_lanes", log2p)
    try:
        for w, h, kind, n in cases:
            y, u, v = synth_frames(n, w, h, seed=13, kind=kind)
            pk, off, sz = oracle.encode_frames(y, u, v, w, h, 2)
            dy, du, dv, st = emu.decode_frames(pk, off, sz, w, h)
            wy, wu, wv, wst = oracle.decode_frames(pk, off, sz, w, h)
            assert (st == 0).all() and np.array_equal(dy, wy) and np.array_equal(du, wu) and np.array_equal(dv, wv), (w, h, kind)
    finally:
        emu.set_option("decode_token_pass", 2)
        emu.set_option("decode_log2_lanes", -1)


def test_decode_corrupt_streams_are_flagged(emu, oracle):
    w, h = 160, 120
    y, u, v = synth_frames(3, w, h, seed=14)
    pk, off, sz = oracle.encode_frames(y, u, v, w, h, 2)
    good = [pk[int(off[i]): int(off[i]) + int(sz[i])].tobytes() for i in range(3)]
    bad = [good[0], good[1][:300] + b"\xff\xd9", b"\xff", good[2][:100] + b"\xff\xc4" + good[2][100:], good[2]]
    bk, boff, bsz = pack(bad)
    dy, du, dv, st = emu.decode_frames(bk, boff, bsz, w, h)
    wy = oracle.decode_frames(pk, off, sz, w, h)[0]
    assert st[0] == 0 and st[4] == 0 and st[1] != 0 and st[2] & amv.ST_SHORT and st[3] & amv.ST_MARKER
    assert np.array_equal(dy[0], wy[0]) and np.array_equal(dy[4], wy[2])


@pytest.mark.parametrize("form", [0, 1, 2])
def test_adpcm_kernels_identical(emu, oracle, form):
    emu.set_option("adpcm_form", form)
    try:
        rng = np.random.default_rng(15)
        nsamp = (rng.integers(0, 900, 70) * 2).astype(np.uint32)
        nsamp[:3] = [0, 2, 1378]
        pcm = synth_pcm(int(nsamp.sum()) + 2, seed=16, kind="noise")
        poff = offsets_of(nsamp)
        step_in = (np.arange(70) * 7 % 89).astype(np.int16)
        out, ooff, osz, so, st = emu.adpcm_encode(pcm, poff, nsamp, step_in)
        wout, _, wsz, wso = oracle.adpcm_encode(pcm, poff, nsamp, step_in)
        assert (st == 0).all() and np.array_equal(osz, wsz) and np.array_equal(out, wout) and np.array_equal(so, wso)
        dec, _, dst = emu.adpcm_decode(out, ooff, osz)
        wdec, _, _ = oracle.adpcm_decode(out, ooff, osz)
        assert (dst == 0).all() and np.array_equal(dec, wdec)
    finally:
        emu.set_option("adpcm_form", 1)


def test_sp5x_and_amvlib_flavours(emu, oracle):
    from oracle_lib import sp5x_from_amv
    w, h, n = 160, 120, 3
    y, u, v = synth_frames(n, w, h, seed=17)
    pk, off, sz = oracle.encode_frames(y, u, v, w, h, 2)
    sp, soff, ssz = sp5x_from_amv(oracle, pk, off, sz)
    sy, su, sv, sst = emu.decode_frames(sp, soff, ssz, w, h, sp5x=True)
    wy, wu, wv, _ = oracle.sp5x_decode_frames(sp, soff, ssz, w, h)
    assert (sst == 0).all() and np.array_equal(sy, wy) and np.array_equal(su, wu) and np.array_equal(sv, wv)
    bgr, bst = emu.decode_frames_bgr24(pk, off, sz, w, h)
    obgr, _ = oracle.amvlib_decode_frames(pk, off, sz, w, h)
    assert (bst == 0).all() and np.array_equal(bgr, obgr)


def test_random_geometries_round_trip_every_lane_count(emu, oracle):
    """Seeded random pictures (size, content, qscale, lanes per frame): the emulated encoder's packets are the oracle's,
    and the emulated decoder -- lean synchronisation pass with its checkpoints, lean token pass, k_idct16 -- returns the
    oracle's planes from them at whatever lane count the draw picked."""
    rng = np.random.default_rng(20261019)
    kinds = ("sinus", "noise", "flat", "edges")
    try:
        for _ in range(14):
            w, h = 16 * int(rng.integers(1, 14)), 8 * int(rng.integers(1, 16))
            if (h // 2) % 8 not in (0, 4):          # the heights the reference's flip addresses correctly
                h += 8
            kind, q, n = kinds[int(rng.integers(0, 4))], int(rng.integers(2, 32)), int(rng.integers(1, 5))
            log2p = int(rng.integers(0, 6))
            y, u, v = synth_frames(n, w, h, seed=int(rng.integers(1, 1 << 30)), kind=kind)
            wpk, woff, wsz = oracle.encode_frames(y, u, v, w, h, q)
            pk, off, sz, st = emu.encode_frames(y, u, v, qscale=q)
            assert (st == 0).all() and np.array_equal(sz, wsz) and np.array_equal(pk, wpk), (w, h, kind, q)
            wy, wu, wv, wst = oracle.decode_frames(wpk, woff, wsz, w, h)
            emu.set_option("decode_log2_lanes", log2p)
            dy, du, dv, dst = emu.decode_frames(wpk, woff, wsz, w, h)
            assert (dst == 0).all() and np.array_equal(dy, wy) and np.array_equal(du, wu) and np.array_equal(dv, wv), (w, h, kind, q, log2p)
    finally:
        emu.set_option("decode_log2_lanes", -1)


@pytest.mark.parametrize("log2p", [0, 3, 5])
def test_bit_flipped_packets_end_and_leave_sound_neighbours_alone(emu, oracle, log2p):
    """Packets with random bit flips in the scan (wrong codes, wrong lengths, runs past 63): every lane count must come
    back -- the walks are bounded by the scan, the checkpoints of the synchronisation pass only ever join what a walk
    really did -- and the sound frames between them decode to the oracle's planes.  The emulator runs the kernels in
    host memory, so a walk that left its buffers would fault here."""
    w, h, n = 208, 176, 6
    rng = np.random.default_rng(31 + log2p)
    y, u, v = synth_frames(n, w, h, seed=32, kind="sinus")
    pk, off, sz = oracle.encode_frames(y, u, v, w, h, 3)
    wy, wu, wv, _ = oracle.decode_frames(pk, off, sz, w, h)
    units = [bytearray(pk[int(off[i]): int(off[i]) + int(sz[i])].tobytes()) for i in range(n)]
    for i in (1, 3, 4):
        for _ in range(1 + 3 * i):
            p = int(rng.integers(2, len(units[i]) - 2))
            units[i][p] ^= 1 << int(rng.integers(0, 8))
            if units[i][p] == 0xff:                 # keep the framing: a new marker would just cut the scan short
                units[i][p] = 0xfe
    bk, boff, bsz = pack([bytes(b) for b in units])
    emu.set_option("decode_log2_lanes", log2p)
    try:
        dy, du, dv, st = emu.decode_frames(bk, boff, bsz, w, h)
    finally:
        emu.set_option("decode_log2_lanes", -1)
    for i in (0, 2, 5):
        assert st[i] == 0 and np.array_equal(dy[i], wy[i]) and np.array_equal(du[i], wu[i]) and np.array_equal(dv[i], wv[i])


def test_emulated_kernels_under_address_sanitizer():
    """The same kernels built with -fsanitize=address (build.sh, ASAN=1) and run in a child process with libasan
    preloaded: every "device", pinned and shared-memory access of the encoder forms, the token passes at 1 / 4 / 32 lanes
    per frame (sound and bit-flipped packets) and the ADPCM forms stays inside its buffer (compute-sanitizer is not
    available on the GPU pool; this is the memcheck the kernels get)."""
    import sys
    built = subprocess.run([BUILD], capture_output=True, text=True, env=dict(os.environ, ASAN="1"))
    if built.returncode == 3:
        pytest.skip("no compiler with libasan.so")
    assert built.returncode == 0, built.stderr[-2000:]
    libasan, so = built.stdout.strip().splitlines()[-2:]
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:abort_on_error=0:exitcode=23")
    out = subprocess.run([sys.executable, os.path.join(HERE, "host_emul", "simt", "asan_target.py"), so], capture_output=True,
                         text=True, env=env, timeout=1500)
    assert out.returncode == 0 and "asan target ok" in out.stdout and "ERROR: AddressSanitizer" not in out.stderr, \
        out.stdout[-2000:] + out.stderr[-4000:]
    # and this file's other tests (flavours, scaler, resampler, container, trellis, host copy paths) against the same build
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-p", "no:cacheprovider",
                          "-k", "not address_sanitizer"], capture_output=True, text=True, env=dict(env, AMV_EMUL_ASAN_SO=so),
                         timeout=2400, cwd=os.path.dirname(HERE))
    assert out.returncode == 0 and "ERROR: AddressSanitizer" not in out.stderr + out.stdout, out.stdout[-4000:] + out.stderr[-4000:]
    # and the GPU parity tests of the kernels this file does not reach, pointed at the sanitizer build (conftest.py, AMV_EMUL=1):
    # plain JPEG, the scaler, the audio resampler, the trellis encoder -- their golden-vector cases (tests that hold torch
    # tensors are left out: torch's own exceptions do not get along with a preloaded libasan)
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_gpu_parity.py"), "-q", "-x", "-m", "gpu",
                          "-p", "no:cacheprovider", "-k", "mjpeg_decode_golden or scaler_matches_golden or audio_resampler_matches_golden "
                          "or adpcm_trellis_golden"],
                         capture_output=True, text=True, env=dict(env, AMV_EMUL="1", AMV_EMUL_ASAN_SO=so), timeout=2400,
                         cwd=os.path.dirname(HERE))
    assert out.returncode == 0 and "ERROR: AddressSanitizer" not in out.stderr + out.stdout, out.stdout[-4000:] + out.stderr[-4000:]
