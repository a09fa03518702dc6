#!/usr/bin/env python
"""The other BASELINE.json configurations, device resident, one JSON line each (they are parity-test
cases, not bench.py lines; this script records their throughput next to the headline):

  config 4  decode -> encode round trip of 208x176 AMV frames with audio (frames per GPU scaled by --frames)
  config 5  1280x720 AMV-style frames, encode + decode
  amvlib    320x240 packets through the amvlib-flavoured decoder (amv_decode_frames_bgr24)

Usage: python profiles/bench_configs.py [--frames N208] [--frames720 N] [--steps K]
Under torchrun every rank runs its own contiguous frame range (weak scaling) and rank 0 prints the
aggregate, timed on the device as the max over ranks.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import amv_codec_tools_b200 as amv  # noqa: E402
import bench  # noqa: E402


def synth(n, w, h, t0, dev, seed):
    return bench.synth_frames_torch(n, t0, dev, seed, w, h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=200000, help="208x176 frames per GPU (config 4 names 1M in total)")
    ap.add_argument("--frames720", type=int, default=8192)
    ap.add_argument("--frames-amvlib", type=int, default=50000)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    lrank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(lrank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    dev = torch.device("cuda", lrank)
    torch.cuda.set_device(dev)
    ctx = amv.AmvCuda(device=dev.index)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    peak, peak_src = bench.load_peaks()

    def timed(fn, steps):
        for _ in range(3):
            fn()
        ctx.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def video_case(name, w, h, n, with_audio, order):
        cw, ch = (w + 1) // 2, (h + 1) // 2
        Y, U, V = synth(n, w, h, rank * n, dev, 11 + rank)
        cap = n * (w * h // 4 + 4096)
        pk = torch.empty(cap, dtype=torch.uint8, device=dev)
        off = torch.zeros(n, dtype=torch.int64, device=dev); size = torch.zeros(n, dtype=torch.int32, device=dev)
        st = torch.zeros(n, dtype=torch.int32, device=dev); st2 = torch.zeros(n, dtype=torch.int32, device=dev)
        DY, DU, DV = torch.empty_like(Y), torch.empty_like(U), torch.empty_like(V)
        pkt_cap = min(w * h * 3 + 4096, 1 << 20)

        def enc(y, u, v, pk=pk, cap=cap, off=off, size=size):
            ctx.encode_frames_raw(y, u, v, w, cw, w * h, cw * ch, n, w, h, None, pk, cap, pkt_cap, amv.LAYOUT_PACKED, off, size,
                                  st, amv.MEM_DEVICE)

        def dec():
            ctx.decode_frames_raw(pk, cap, off, size, n, w, h, DY, DU, DV, w, cw, w * h, cw * ch, st2, amv.MEM_DEVICE)

        aud = None
        if with_audio:      # one 1378-sample chunk per frame (22050 Hz at 16 fps)
            ns, csz = 1378, 8 + 689
            pcm = (8000 * torch.sin(torch.arange(n * ns, device=dev, dtype=torch.float32) * 0.1254)).to(torch.int16)
            poff = torch.arange(n, device=dev, dtype=torch.int64) * ns
            nsam = torch.full((n,), ns, dtype=torch.int32, device=dev)
            ooff = torch.arange(n, device=dev, dtype=torch.int64) * csz
            osz = torch.full((n,), csz, dtype=torch.int32, device=dev)
            ck = torch.zeros(n * csz, dtype=torch.uint8, device=dev)
            so = torch.zeros(n, dtype=torch.int16, device=dev); ast = torch.zeros(n, dtype=torch.int32, device=dev)
            dpcm = torch.zeros(n * ns, dtype=torch.int16, device=dev)
            ctx.adpcm_enc_chunks_raw(pcm, n * ns, poff, nsam, None, so, n, ck, ck.numel(), ooff, ast, amv.MEM_DEVICE)

            def aud():
                ctx.adpcm_dec_chunks_raw(ck, ck.numel(), ooff, osz, n, dpcm, n * ns, poff, ast, amv.MEM_DEVICE)
                ctx.adpcm_enc_chunks_raw(dpcm, n * ns, poff, nsam, None, so, n, ck, ck.numel(), ooff, ast, amv.MEM_DEVICE)

        torch.cuda.synchronize(dev)        # torch fills its tensors on ITS stream; the context runs on another one
        enc(Y, U, V)
        ctx.sync()
        pkt_bytes = int(size.to(torch.int64).sum().item())
        if order == "dec_enc":      # config 4: decode the packets, re-encode the decoded planes into a second packet buffer
            # (the reference's decoder and encoder use different quantisers, SURVEY 9.1, so the re-encoded packets differ in size)
            cap2 = n * (w * h + 4096)
            pk2 = torch.empty(cap2, dtype=torch.uint8, device=dev)
            off2 = torch.zeros(n, dtype=torch.int64, device=dev); size2 = torch.zeros(n, dtype=torch.int32, device=dev)
            torch.cuda.synchronize(dev)

            def step():
                dec()
                enc(DY, DU, DV, pk2, cap2, off2, size2)
                if aud:
                    aud()
        else:
            def step():
                enc(Y, U, V)
                dec()
        ms = timed(step, args.steps)
        assert int(st.abs().sum().item()) == 0 and int(st2.abs().sum().item()) == 0
        if order == "dec_enc":
            pkt_bytes = (pkt_bytes + int(size2.to(torch.int64).sum().item())) / 2
        bpf = w * h * 3 // 2 + pkt_bytes / n + (3453 if aud else 0)
        fps = world * n / (ms / 1e3)
        gbs = 2 * bpf * n / (ms / 1e3) / 1e9          # both directions move planes + packet (+ chunk)
        return {"config": name, "width": w, "height": h, "frames_per_gpu": n, "n_gpus": world, "ms_per_pass": ms,
                "frames_per_s": fps, "avg_packet_bytes": pkt_bytes / n, "audio_chunks": bool(aud),
                "hbm_roofline": {"achieved_GBs_per_gpu": gbs, "peak": peak, "frac": gbs / peak, "peak_source": peak_src,
                                 "bytes_per_frame_round_trip": 2 * bpf}}

    out = []
    out.append(video_case("config 4: 208x176 decode->encode round trip + ADPCM chunk per frame", 208, 176, args.frames, True, "dec_enc"))
    torch.cuda.empty_cache()
    out.append(video_case("config 5: 1280x720 encode+decode", 1280, 720, args.frames720, False, "enc_dec"))
    torch.cuda.empty_cache()

    # amvlib flavour at 320x240
    w, h, n = 320, 240, args.frames_amvlib
    cw, ch = w // 2, h // 2
    Y, U, V = synth(n, w, h, rank * n, dev, 31 + rank)
    cap = n * 24 * 1024
    pk = torch.empty(cap, dtype=torch.uint8, device=dev)
    off = torch.zeros(n, dtype=torch.int64, device=dev); size = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.zeros(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    ctx.encode_frames_raw(Y, U, V, w, cw, w * h, cw * ch, n, w, h, None, pk, cap, 65536, amv.LAYOUT_PACKED, off, size, st, amv.MEM_DEVICE)
    ctx.sync()
    del Y, U, V
    pkt_bytes = int(size.to(torch.int64).sum().item())
    lb = amv.amvlib_line_bytes(w)
    bgr = torch.empty(n * lb * h, dtype=torch.uint8, device=dev)
    assert int(st.abs().sum().item()) == 0, "encode status %s" % torch.unique(st).tolist()
    ms = timed(lambda: ctx.decode_frames_bgr24_raw(pk, cap, off, size, n, w, h, bgr, lb, lb * h, st, amv.MEM_DEVICE), args.steps)
    assert int(st.abs().sum().item()) == 0, "decode status %s at %s" % (torch.unique(st).tolist(), (st != 0).nonzero()[:8].flatten().tolist())
    bpf = pkt_bytes / n + lb * h
    out.append({"config": "amvlib flavour: 320x240 packets -> BGR24 bitmaps (AmvVideoDecode mirror)", "frames_per_gpu": n,
                "n_gpus": world, "ms_per_pass": ms, "frames_per_s": world * n / (ms / 1e3),
                "hbm_roofline": {"achieved_GBs_per_gpu": bpf * n / (ms / 1e3) / 1e9, "peak": peak,
                                 "frac": bpf * n / (ms / 1e3) / 1e9 / peak, "bytes_per_frame": bpf}})
    if rank == 0:
        for o in out:
            print(json.dumps(o))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
