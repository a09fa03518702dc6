#!/usr/bin/env python
"""Throughput of the stages ffmpeg.c runs in front of the AMV encoders, device resident, one JSON
line each, next to the measured HBM peak: range conversion, the `-s WxH` picture scaler, the audio resampler.
Algorithmic bytes: input planes / samples read once + output written once.

Usage: python profiles/bench_prestages.py [--steps K]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ctx = amv.AmvCuda(device=0, lib_path=os.environ.get("AMV_LIB") or amv.LIB_PATH)      # AMV_LIB: a build variant to compare
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    peak, peak_src = bench.load_peaks()

    def timed(fn):
        for _ in range(3):
            fn()
        ctx.sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            a.record(stream)
            for _ in range(args.steps):
                fn()
            b.record(stream)
        b.synchronize()
        return a.elapsed_time(b) / args.steps

    def line(name, units, unit, ms, nbytes, extra):
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(json.dumps(dict(stage=name, value=units / (ms * 1e-3), unit=unit, ms=ms,
                              roofline=dict(bound="hbm", achieved=gbs, peak=peak, unit="GB/s", frac=gbs / peak, peak_source=peak_src),
                              **extra)), flush=True)

    g = torch.Generator(device=dev); g.manual_seed(1)
    # ---- scaler: 640x480 -> 320x240 (the north star's frame size), 352x288 -> 208x176, 320x240 -> 1280x720
    for (iw, ih, ow, oh, n) in ((640, 480, 320, 240, 8192), (352, 288, 208, 176, 16384), (320, 240, 1280, 720, 2048)):
        icw, ich, ocw, och = iw // 2, ih // 2, ow // 2, oh // 2
        y = torch.randint(0, 256, (n, ih, iw), dtype=torch.uint8, device=dev, generator=g)
        u = torch.randint(0, 256, (n, ich, icw), dtype=torch.uint8, device=dev, generator=g)
        v = torch.randint(0, 256, (n, ich, icw), dtype=torch.uint8, device=dev, generator=g)
        oy = torch.empty((n, oh, ow), dtype=torch.uint8, device=dev)
        ou = torch.empty((n, och, ocw), dtype=torch.uint8, device=dev)
        ov = torch.empty((n, och, ocw), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        nbytes = n * (iw * ih + 2 * icw * ich + ow * oh + 2 * ocw * och)
        for form in (1, 2):
            ctx.set_option("scale_form", form)
            ms = timed(lambda: ctx.scale_frames_raw(y, u, v, iw, icw, iw * ih, icw * ich, n, iw, ih, oy, ou, ov, ow, ocw, ow * oh,
                                                    ocw * och, ow, oh, amv.MEM_DEVICE))
            line("scale", n, "frames/s", ms, nbytes, dict(config="%dx%d->%dx%d x %d" % (iw, ih, ow, oh, n),
                                                          form={1: "tiles", 2: "staged tiles"}[form]))
        ctx.set_option("scale_form", 1)
        del y, u, v, oy, ou, ov
    # ---- range conversion at 320x240
    n, w, h = 32768, 320, 240
    y = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device=dev, generator=g)
    u = torch.randint(0, 256, (n, h // 2, w // 2), dtype=torch.uint8, device=dev, generator=g)
    v = torch.randint(0, 256, (n, h // 2, w // 2), dtype=torch.uint8, device=dev, generator=g)
    torch.cuda.synchronize()
    ms = timed(lambda: ctx.convert_range_raw(y, u, v, w, w // 2, w * h, w * h // 4, n, w, h, 0, y, u, v, w, w // 2, w * h, w * h // 4,
                                             amv.MEM_DEVICE))
    line("range", n, "frames/s", ms, 2 * n * w * h * 3 // 2, dict(config="320x240 x %d in place" % n))
    del y, u, v
    # ---- audio resampler: 44.1 kHz stereo / 48 kHz mono -> 22050 Hz mono
    for (rate, ch, nin) in ((44100, 2, 1 << 28), (48000, 1, 1 << 28), (8000, 1, 1 << 26)):
        pcm = torch.randint(-32768, 32768, (nin * ch,), dtype=torch.int16, device=dev, generator=g)
        k = ctx.audio_resample_count(nin, rate, 22050)
        out = torch.empty((k,), dtype=torch.int16, device=dev)
        torch.cuda.synchronize()
        for form in (1, 2):
            ctx.set_option("resample_form", form)
            ms = timed(lambda: ctx.audio_resample_raw(pcm, nin, ch, rate, 22050, out, k, amv.MEM_DEVICE))
            line("audio_resample", k, "output samples/s", ms, 2 * (nin * ch + k),
                 dict(config="%d Hz x %d ch -> 22050 Hz mono, %d input samples per channel" % (rate, ch, nin),
                      form={1: "tiles", 2: "phase rows"}[form]))
        ctx.set_option("resample_form", 2)
        del pcm, out


if __name__ == "__main__":
    main()
