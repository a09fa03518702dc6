"""The N > 1 host logic on CPU: world_size-2 gloo processes exercising the frame-range sharding,
the host-side gather of the packet (offset, size) tables and the ADPCM stream-state hand-over.
The codec work itself is stood in for by the oracle here (no GPU in this container); the GPU box
runs the same sharding through bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import amv_codec_tools_b200 as amv
from oracle_lib import Oracle, offsets_of, synth_frames, synth_pcm

shard_range = amv.sharding.shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 100000, 1000003):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            lens = [hi - lo for lo, hi in ranges]
            assert max(lens) - min(lens) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = Oracle()
        # ---- video: every rank encodes its own contiguous frame range, tables are gathered
        w, h, n = 64, 48, 13
        y, u, v = synth_frames(n, w, h, seed=81)
        lo, hi = shard_range(n, rank, world)
        pk, off, sz = oracle.encode_frames(y[lo:hi], u[lo:hi], v[lo:hi], w, h, 2)
        gsz, goff, base = amv.sharding.gather_packet_table(sz, dist)
        # ---- audio: one continuous stream, chunks sharded, state chained rank to rank
        nchunks, ns = 11, 200
        pcm = synth_pcm(nchunks * ns, seed=82, kind="noise")
        clo, chi = shard_range(nchunks, rank, world)
        outs = []

        def encode_range(step_in):
            st = step_in
            for c in range(clo, chi):
                o, _, _, so = oracle.adpcm_encode(pcm[c * ns:(c + 1) * ns], np.zeros(1, np.uint64), np.array([ns], np.uint32),
                                                  np.array([st], np.int16))
                outs.append(o)
                st = int(so[0])
            return st
        final = amv.sharding.chain_stream_state(encode_range, dist)
        q.put((rank, pk.tobytes(), gsz, goff, base, b"".join(o.tobytes() for o in outs), final))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gather_and_chain():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    oracle = Oracle()
    w, h, n = 64, 48, 13
    y, u, v = synth_frames(n, w, h, seed=81)
    wpk, woff, wsz = oracle.encode_frames(y, u, v, w, h, 2)
    # the per-rank packets, laid at the gathered base offsets, ARE the single-process stream
    stream = bytearray(int(wsz.sum()))
    for rank, pk, gsz, goff, base, audio, final in res:
        assert np.array_equal(gsz, wsz) and np.array_equal(goff, woff)
        stream[base: base + len(pk)] = pk
    assert bytes(stream) == wpk.tobytes()
    nchunks, ns = 11, 200
    pcm = synth_pcm(nchunks * ns, seed=82, kind="noise")
    st, chunks = 0, []
    for c in range(nchunks):
        o, _, _, so = oracle.adpcm_encode(pcm[c * ns:(c + 1) * ns], np.zeros(1, np.uint64), np.array([ns], np.uint32),
                                          np.array([st], np.int16))
        chunks.append(o.tobytes())
        st = int(so[0])
    assert b"".join(r[5] for r in res) == b"".join(chunks)
    assert res[-1][6] == st


# ------------------------------------------------------------------ one audio stream resampled by several ranks
def _fir_window(win, in_base, k_start, k_count, rate_in, rate_out, bank):
    """stand-in for amv_audio_resample_from on the CPU (the oracle's bank, the stream positions of resample2.c:290-296):
    outputs k_start .. k_start + k_count - 1 from the mono window win = stream[in_base : in_base + len(win)]"""
    flen = bank.shape[1]
    index0 = -1024 * ((flen - 1) // 2)
    out = np.zeros(k_count, np.int16)
    for j in range(k_count):
        index = index0 + ((k_start + j) * rate_in * 1024) // rate_out
        first, f = index >> 10, bank[index & 1023].astype(np.int64)
        pos = np.arange(first, first + flen)
        if first < 0:
            assert in_base == 0
            pos = np.abs(pos) % len(win)
        else:
            pos = pos - in_base
        acc = int((win[pos].astype(np.int64) * f).sum())
        acc = ((acc + (1 << 31)) % (1 << 32)) - (1 << 31)              # the reference's 32-bit accumulator
        out[j] = max(-32768, min(32767, (acc + (1 << 14)) >> 15))
    return out


def _resample_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle, lib = Oracle(), amv.load_library()
        res = []
        for rate, n in ((48000, 6000), (8000, 900), (44100, 40)):
            pcm = synth_pcm(n, seed=rate, kind="noise")
            k0, kc, base, nwin = amv.sharding.resample_shard(n, rate, 22050, rank, world, lib)
            out = _fir_window(pcm[base:base + nwin], base, k0, kc, rate, 22050, oracle.resample_bank(rate, 22050))
            # the window is exactly what the rank's outputs need: the ABI's own count over it is at least the share
            assert kc == 0 or lib.amv_audio_resample_count(base + nwin, rate, 22050) - k0 >= kc
            res.append((k0, out.tobytes()))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_resample_shards_concatenate_to_the_stream(world):
    """every rank resamples its output range from its own window of the stream; the pieces, in rank order, are the
    oracle's (= the reference's) output over the whole stream"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_resample_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    oracle = Oracle()
    for i, (rate, n) in enumerate(((48000, 6000), (8000, 900), (44100, 40))):
        want = oracle.audio_resample(synth_pcm(n, seed=rate, kind="noise"), 1, rate, 22050)
        pos = 0
        for rank, r in res:
            k0, piece = r[i]
            assert k0 == pos or len(piece) == 0
            pos += len(piece) // 2
        assert b"".join(r[i][1] for _, r in res) == want.tobytes()
