#!/usr/bin/env python
"""bench.py -- the AMV codec path on B200(s): frames/sec, HBM roofline and the CPU reference beside it.

    python bench.py --gpus N --steps K --warmup W                  # our arm (libamvcuda), BASELINE config 2
    python bench.py --config {2,3,4,5} ...                         # the other BASELINE.json configurations
    python bench.py --impl reference --gpus N --steps K ...        # the reference's CPU codecs on the host cores

A step = one pass of the hot path over one batch of synthetic input, through the C ABI:
  config 2 (default)  320x240: amv_encode_frames over the batch (planes -> packed packets), amv_decode_frames back
  config 3            IMA-ADPCM-AMV: amv_adpcm_enc_chunks over 1 M chunks of 1378 samples, amv_adpcm_dec_chunks back
  config 4            208x176, ONE job of 1 M frames split over the ranks: amv_decode_frames, amv_encode_frames of the
                      decoded planes, one ADPCM chunk per frame decoded and re-encoded
  config 5            1280x720: encode + decode like config 2
`value` is units / device time with every buffer resident in HBM; `e2e` is the same work with pinned HOST buffers
(AMV_MEM_HOST: H2D of the inputs and D2H of the results inside the timed region).  Frames are intra-only, so with N
GPUs the job's units are cut into N contiguous ranges, one per rank, with no collective on the data path; what a split
job still needs is the packet table of the concatenated stream (the muxer consumes packets strictly in order,
libavformat/amvenc.c:378-406): every step ends its encode with an all-gather of the per-rank packet sizes (metadata, 4
bytes per frame over NCCL) and a prefix sum, and rank 0 checks the table after the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "AMV 320x240 frames/sec enc+dec"
W, H, CW, CH, PKT_CAP = 320, 240, 160, 120, 65536       # config 2's shape (profiles/prof_target.py uses these)
PEAK_NOMINAL = 8000.0               # GB/s, NVIDIA's figure for B200 HBM3e (state both denominators)
KERNELS = ("encode", "compact", "unstuff", "sync", "tokens", "idct", "adpcm_dec", "adpcm_enc")

CONFIGS = {
    2: dict(kind="video", w=320, h=240, frames=100000, e2e_frames=16384, order="enc_dec", audio=False, scaling="weak",
            metric=METRIC, unit="frames/s", pkt_cap=65536, cap_per_frame=24 * 1024, ref_frames=256,
            name="BASELINE config 2: %(n)d synthetic 320x240 YUVJ420P frames per GPU, qscale 2: amv_encode_frames (packed packets) "
                 "then amv_decode_frames of those packets"),
    4: dict(kind="video", w=208, h=176, frames=1000000, sub=250000, e2e_frames=32768, order="dec_enc", audio=True, scaling="strong",
            metric="AMV 208x176 frames/sec dec+enc round trip with audio", unit="frames/s", pkt_cap=32768, cap_per_frame=20 * 1024,
            ref_frames=512,
            name="BASELINE config 4: ONE job of %(total)d 208x176 AMV frames (+ one 1378-sample ADPCM chunk each), contiguous frame range "
                 "per GPU (%(n)d frames here, in sub-batches of <= %(sub)d): amv_decode_frames, amv_encode_frames of the decoded planes, "
                 "amv_adpcm_dec_chunks + amv_adpcm_enc_chunks"),
    5: dict(kind="video", w=1280, h=720, frames=8192, e2e_frames=1024, order="enc_dec", audio=False, scaling="weak",
            metric="AMV 1280x720 frames/sec enc+dec", unit="frames/s", pkt_cap=1 << 20, cap_per_frame=320 * 1024, ref_frames=24,
            name="BASELINE config 5: %(n)d synthetic 1280x720 YUVJ420P frames per GPU, qscale 2: amv_encode_frames then amv_decode_frames"),
    3: dict(kind="adpcm", chunks=1000000, ns=1378, e2e_chunks=262144, scaling="weak", metric="IMA-ADPCM-AMV chunks/sec enc+dec",
            unit="chunks/s", ref_chunks=4096,
            name="BASELINE config 3: %(n)d independent 22050 Hz mono chunks of 1378 samples per GPU: amv_adpcm_enc_chunks then "
                 "amv_adpcm_dec_chunks"),
}


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of each hot kernel from the committed ncu --set full capture
    (profiles/traffic.json, bytes per frame of that capture at 320x240); scaled to the frames one launch processes here."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # samples taken under load are the upper half of the clock readings
        sm_load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(sm_load) if sm_load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- synthetic input
def synth_frames_torch(n, t0, device, seed, w=320, h=240):
    """SURVEY 8d generator on the device: noisy sinusoids, YUVJ420P full range."""
    import torch
    cw, ch = (w + 1) // 2, (h + 1) // 2
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    Y = torch.empty((n, h, w), dtype=torch.uint8, device=device)
    U = torch.empty((n, ch, cw), dtype=torch.uint8, device=device)
    V = torch.empty((n, ch, cw), dtype=torch.uint8, device=device)
    xx = torch.arange(w, device=device, dtype=torch.float32)[None, None, :]
    yy = torch.arange(h, device=device, dtype=torch.float32)[None, :, None]
    cx = torch.arange(cw, device=device, dtype=torch.float32)[None, None, :]
    cy = torch.arange(ch, device=device, dtype=torch.float32)[None, :, None]
    step = max(1, (2048 * 320 * 240) // (w * h))
    for a in range(0, n, step):
        b = min(n, a + step)
        t = (torch.arange(a, b, device=device, dtype=torch.float32) + t0)[:, None, None]
        y = 128 + 60 * torch.sin((xx + 3 * t) / 17.0) + 50 * torch.cos((yy - 2 * t) / 11.0)
        y = y + 6.0 * torch.randn((b - a, h, w), device=device, generator=g)
        Y[a:b] = y.round().clamp(0, 255).to(torch.uint8)
        U[a:b] = (128 + 40 * torch.sin((cx + t) / 23.0) + 0 * cy).round().clamp(0, 255).to(torch.uint8)
        V[a:b] = (128 + 40 * torch.cos((cy + t) / 19.0) + 0 * cx).round().clamp(0, 255).to(torch.uint8)
    return Y, U, V


def synth_pcm_torch(n, device, seed):
    """SURVEY 8d audio: two tones + a little noise, mono int16"""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(n, dtype=torch.int16, device=device)
    step = 1 << 26
    for a in range(0, n, step):
        b = min(n, a + step)
        t = torch.arange(a, b, device=device, dtype=torch.float32)
        x = 8000 * torch.sin(t * (2 * np.pi * 440 / 22050)) + 2000 * torch.sin(t * (2 * np.pi * 1234 / 22050))
        x = x + 300 * torch.randn(b - a, device=device, generator=g)
        out[a:b] = x.round().clamp(-32768, 32767).to(torch.int16)
    return out


# --------------------------------------------------------------------------- reference arm
def _pin(core):
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass


def _ref_worker(args):
    """One process = one single-threaded reference codec context pair pinned to its own core (the reference rejects
    thread_count > 1 for AMV, mpegvideo_enc.c:451-456).  Video: (frames, encode s, decode s, packet bytes, kind, amvlib s);
    audio: (chunks, encode s, decode s, bytes, kind, None)."""
    wid, core, cfgno, units, rounds = args
    _pin(core)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, Ref, offsets_of, synth_frames, synth_pcm
    cfg = CONFIGS[cfgno]
    kind = "reference" if Ref.available() else "port"
    codec = Ref() if kind == "reference" else Oracle()
    te = td = 0.0
    nbytes = 0
    lib_s = None
    if cfg["kind"] == "adpcm":
        ns = cfg["ns"]
        pcm = synth_pcm(ns * units + 4096, seed=100 + wid, kind="tones")
        for _ in range(rounds):
            t = time.perf_counter()
            if kind == "reference":
                out, off, sz, cons = codec.adpcm_encode_stream(pcm, ns, max_chunks=units)
            else:
                nsm = np.full(units, ns, np.uint32)
                out, off, sz, _ = codec.adpcm_encode(pcm[: ns * units], offsets_of(nsm), nsm, np.zeros(units, np.int16))
            te += time.perf_counter() - t
            t = time.perf_counter()
            codec.adpcm_decode(out, off, sz)
            td += time.perf_counter() - t
            nbytes += int(np.sum(sz))
        return units * rounds, te, td, nbytes, kind, None
    w, h = cfg["w"], cfg["h"]
    y, u, v = synth_frames(units, w, h, seed=100 + wid, t0=1000 * wid)
    audio = cfg.get("audio")
    if audio:
        ns = 1378
        pcm = synth_pcm(ns * units + 4096, seed=200 + wid, kind="tones")
    for _ in range(rounds):
        t = time.perf_counter()
        if kind == "reference":
            pk, off, sz = codec.encode_frames(y, u, v, w, h, quality=0)
        else:
            pk, off, sz = codec.encode_frames(y, u, v, w, h, 2)
        te += time.perf_counter() - t
        t = time.perf_counter()
        codec.decode_frames(pk, off, sz, w, h)
        td += time.perf_counter() - t
        if audio:       # the chunk that travels with every frame: encode + decode
            t = time.perf_counter()
            if kind == "reference":
                ao, aoff, asz, _ = codec.adpcm_encode_stream(pcm, ns, max_chunks=units)
            else:
                nsm = np.full(units, ns, np.uint32)
                ao, aoff, asz, _ = codec.adpcm_encode(pcm[: ns * units], offsets_of(nsm), nsm, np.zeros(units, np.int16))
            te += time.perf_counter() - t
            t = time.perf_counter()
            codec.adpcm_decode(ao, aoff, asz)
            td += time.perf_counter() - t
        nbytes += int(sz.sum())
    # amvlib (C-AMVDecoder) timed the same way, decode only, packets in memory (it has no encoder)
    if kind == "reference" and cfgno == 2:
        from oracle_lib import AmvlibRef
        if AmvlibRef.available():
            al = AmvlibRef()
            t = time.perf_counter()
            for _ in range(rounds):
                al.video_decode(pk, off, sz, w, h)
            lib_s = time.perf_counter() - t
    return units * rounds, te, td, nbytes, kind, lib_s


def run_cpu_reference(cfgno, units_per_worker, rounds=1, workers=None):
    import multiprocessing as mp
    try:
        cores = sorted(os.sched_getaffinity(0))
    except Exception:
        cores = list(range(os.cpu_count() or 1))
    workers = workers or len(cores)
    ctx = mp.get_context("spawn")
    t = time.perf_counter()
    with ctx.Pool(workers) as pool:
        res = pool.map(_ref_worker, [(i, cores[i % len(cores)], cfgno, units_per_worker, rounds) for i in range(workers)])
    wall = time.perf_counter() - t
    units = sum(r[0] for r in res)
    busy = max(r[1] + r[2] for r in res)                # codec time of the slowest worker (excludes spawn / synthesis)
    enc_s, dec_s = max(r[1] for r in res), max(r[2] for r in res)
    out = {"units": units, "seconds": busy, "wall": wall, "ups": units / busy, "workers": workers, "kind": res[0][4],
           "pkt_bytes": sum(r[3] for r in res), "per_core": units / busy / workers,
           "encode_ups": units / enc_s if enc_s > 0 else None, "decode_ups": units / dec_s if dec_s > 0 else None,
           "cpu_model": cpu_model(), "pinned": True}
    if all(r[5] for r in res):
        out["amvlib_seconds"] = max(r[5] for r in res)
    return out


def cpu_baseline_block(cfgno, cb, sample):
    cfg = CONFIGS[cfgno]
    return {"value": cb["ups"], "unit": cfg["unit"], "cores": cb["workers"], "kind": cb["kind"], "sample": sample,
            "per_core": cb["per_core"], "encode_direction": cb["encode_ups"], "decode_direction": cb["decode_ups"],
            "cpu_model": cb["cpu_model"], "workers_pinned_to_distinct_cores": cb["pinned"],
            "build": "generic C, -O3, MMX disabled (the bit-identical path, SURVEY 8c) -- a reported baseline, not the target"}


# stdout carries exactly ONE line, the JSON result: everything else any library prints there (NCCL's version
# banner, torchrun notices) is diverted to stderr at the file-descriptor level for the whole run
_RESULT_FD = None


def protect_stdout():
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def ref_units(args, cfg):
    if args.ref_frames_per_worker > 0:
        return max(8, args.ref_frames_per_worker)
    return cfg.get("ref_frames") or cfg.get("ref_chunks")


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = CONFIGS[args.config]
    per = ref_units(args, cfg)
    vals = []
    if args.warmup >= 1:
        run_cpu_reference(args.config, max(8, per // 4), 1)
    for _ in range(args.steps):
        vals.append(run_cpu_reference(args.config, per, 1))
    total = sum(v["units"] for v in vals)
    total_s = sum(v["seconds"] for v in vals)
    ups = total / total_s
    what = {2: "AMV 320x240 encode+decode round trip", 3: "ADPCM-IMA-AMV encode+decode of 1378-sample chunks",
            4: "AMV 208x176 decode+encode round trip with one ADPCM chunk per frame", 5: "AMV 1280x720 encode+decode round trip"}[args.config]
    cb = dict(vals[0]); cb["ups"] = ups
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": ups, "unit": cfg["unit"], "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": "u8" if cfg["kind"] == "video" else "s16", "data": "synthetic",
        "config": {"workload": what + ", reference CPU codecs (AMVmuxer libavcodec 51.47.1, generic C), one single-threaded codec "
                               "context per host core, pinned", "units_per_step": vals[0]["units"], "qscale": 2},
        "cpu_baseline": cpu_baseline_block(args.config, cb, "%d units per core per step x %d steps, in memory" % (per, args.steps)),
        "e2e": {"value": ups, "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if cfg["kind"] == "video":
        line["config"].update(width=cfg["w"], height=cfg["h"])
    if all("amvlib_seconds" in v for v in vals):      # the reference's second decoder (C-AMVDecoder/amvlib), decode only
        lib_s = sum(v["amvlib_seconds"] for v in vals)
        line["amvlib"] = {"value": total / lib_s, "unit": "frames/s", "cores": vals[0]["workers"], "kind": "reference",
                          "what": "AmvVideoDecode of the same packets to BGR24, packets in memory, one process per core",
                          "per_core": total / lib_s / vals[0]["workers"]}
    emit(line)
    return 0


# --------------------------------------------------------------------------- our arm: common plumbing
class Job:
    """rank / device / context plumbing shared by the video and the audio benches"""

    def __init__(self, args):
        import torch
        import amv_codec_tools_b200 as amv
        self.torch, self.amv, self.args = torch, amv, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        try:
            self.affinity0 = os.sched_getaffinity(0)
        except Exception:
            self.affinity0 = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # every rank's host threads on their own slice of the cores (pinned-buffer copies and the copy pipeline's
            # host side of N ranks otherwise pile up on the same cores)
            try:
                cores = sorted(os.sched_getaffinity(0))
                per = len(cores) // self.world
                if per >= 1:
                    os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]))
            except Exception:
                pass
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist
        else:
            torch.cuda.set_device(0)
        self.dev = torch.device("cuda", local_rank if self.world > 1 else 0)
        self.warm = max(3, args.warmup)
        self.opts = [(kv.split("=")[0], int(kv.split("=")[1])) for kv in args.opt]
        self.ctx = self.new_ctx()
        self.stream = torch.cuda.Stream(device=self.dev)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.ctx.set_option("profile_events", 1)

    def new_ctx(self):
        c = self.amv.AmvCuda(device=self.dev.index)
        for k, v in self.opts:
            c.set_option(k, v)
        return c

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(self, ok):
        if self.dist is None:
            return bool(ok)
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def timed(self, step_fn, steps):
        """W warm-up steps, then K steps between CUDA events on the launching stream, barrier + synchronize on both sides"""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            for _ in range(self.warm):
                step_fn()
        self.ctx.sync()
        for k in KERNELS:      # drop warm-up samples
            self.ctx.get_stat(k + "_kernel_ns")
        sampler = ClockSampler(self.dev.index)
        launches0 = self.ctx.launch_count()
        self.barrier()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            ev0.record(self.stream)
            for _ in range(steps):
                step_fn()
            ev1.record(self.stream)
        self.barrier()
        clocks = sampler.stop()
        ms_total = ev0.elapsed_time(ev1)
        kern = {}
        for k in KERNELS:
            cnt = self.ctx.get_stat(k + "_kernel_launches")
            ns = self.ctx.get_stat(k + "_kernel_ns")
            kern[k] = {"launches": cnt, "ms": ns / 1e6}
        return ms_total, self.max_over_ranks(ms_total), clocks, self.ctx.launch_count() - launches0, kern

    def finish(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
            self.dist = None
        if self.affinity0 is not None:       # the CPU reference that follows (rank 0) gets all the host cores back
            try:
                os.sched_setaffinity(0, self.affinity0)
            except Exception:
                pass


def shard_range(n, rank, world):
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


# --------------------------------------------------------------------------- our arm: video (configs 2, 4, 5)
def bench_video(args):
    J = Job(args)
    torch, amv, dev, ctx, dist, world, rank = J.torch, J.amv, J.dev, J.ctx, J.dist, J.world, J.rank
    cfg = CONFIGS[args.config]
    W, H = cfg["w"], cfg["h"]
    CW, CH = (W + 1) // 2, (H + 1) // 2
    FRAME_BYTES = W * H + 2 * CW * CH
    PKT_CAP = cfg["pkt_cap"]
    order = cfg["order"]
    total = args.frames if args.frames > 0 else cfg["frames"]
    if cfg["scaling"] == "strong":      # ONE job of `total` frames: rank r owns the contiguous range shard_range(total, r, world)
        lo, hi = shard_range(total, rank, world)
        n = hi - lo
        first_frame = lo
    else:                               # the job grows with the ranks: world * n frames, rank r owns [r * n, (r + 1) * n)
        n = total
        first_frame = rank * n
    sub = min(n, cfg.get("sub", n))     # frames per pass of the kernels (config 4: bounds the workspaces at one GPU)
    nsub = (n + sub - 1) // sub

    # ---- device-resident workload
    cap = sub * cfg["cap_per_frame"]
    size_all = torch.zeros(n, dtype=torch.int32, device=dev)           # packet sizes of the encode direction, all sub-batches
    st_e = torch.zeros(sub, dtype=torch.int32, device=dev)
    st_d = torch.zeros(sub, dtype=torch.int32, device=dev)
    off = torch.zeros(sub, dtype=torch.int64, device=dev)
    DY = torch.empty((sub, H, W), dtype=torch.uint8, device=dev)
    DU = torch.empty((sub, CH, CW), dtype=torch.uint8, device=dev)
    DV = torch.empty((sub, CH, CW), dtype=torch.uint8, device=dev)
    in_pkt_bytes = 0
    if order == "enc_dec":
        assert nsub == 1
        Y, U, V = synth_frames_torch(n, first_frame, dev, 1 + rank, W, H)
        pk = torch.empty(cap, dtype=torch.uint8, device=dev)
        inputs = None
    else:
        # the clip to transcode: packets of our own encoder over the synthetic frames, made sub-batch by sub-batch
        inputs = []
        pk = torch.empty(cap, dtype=torch.uint8, device=dev)
        for j in range(nsub):
            m = min(sub, n - j * sub)
            Y, U, V = synth_frames_torch(m, first_frame + j * sub, dev, 1 + rank + 1000 * j, W, H)
            ioff = torch.zeros(m, dtype=torch.int64, device=dev)
            isz = torch.zeros(m, dtype=torch.int32, device=dev)
            torch.cuda.synchronize(dev)
            ctx.encode_frames_raw(Y, U, V, W, CW, W * H, CW * CH, m, W, H, None, pk, cap, PKT_CAP, amv.LAYOUT_PACKED, ioff, isz, st_e,
                                  amv.MEM_DEVICE)
            ctx.sync()
            assert int(st_e[:m].abs().sum().item()) == 0
            used = int(isz.to(torch.int64).sum().item())
            inputs.append((pk[:used].clone(), ioff, isz, m))
            in_pkt_bytes += used
            del Y, U, V
        Y = U = V = None
        torch.cuda.empty_cache()
    aud = None
    if cfg["audio"]:        # one 1378-sample chunk per frame (22050 Hz at 16 fps), decoded and re-encoded with the frame
        ns, csz = 1378, 8 + 689
        pcm = synth_pcm_torch(sub * ns, dev, 7 + rank)
        poff = torch.arange(sub, device=dev, dtype=torch.int64) * ns
        nsam = torch.full((sub,), ns, dtype=torch.int32, device=dev)
        ooff = torch.arange(sub, device=dev, dtype=torch.int64) * csz
        osz = torch.full((sub,), csz, dtype=torch.int32, device=dev)
        ck = torch.zeros(sub * csz, dtype=torch.uint8, device=dev)
        ck2 = torch.zeros(sub * csz, dtype=torch.uint8, device=dev)
        so = torch.zeros(sub, dtype=torch.int16, device=dev)
        ast = torch.zeros(sub, dtype=torch.int32, device=dev)
        dpcm = torch.zeros(sub * ns, dtype=torch.int16, device=dev)
        torch.cuda.synchronize(dev)
        ctx.adpcm_enc_chunks_raw(pcm, sub * ns, poff, nsam, None, so, sub, ck, ck.numel(), ooff, ast, amv.MEM_DEVICE)
        ctx.sync()
        del pcm

        def aud(m):
            ctx.adpcm_dec_chunks_raw(ck, ck.numel(), ooff, osz, m, dpcm, sub * ns, poff, ast, amv.MEM_DEVICE)
            ctx.adpcm_enc_chunks_raw(dpcm, sub * ns, poff, nsam, None, so, m, ck2, ck2.numel(), ooff, ast, amv.MEM_DEVICE)
    torch.cuda.synchronize(dev)

    # the packet table of the whole (split) job: all-gather of the ranks' sizes + exclusive prefix sum, on the device
    gathered = {}

    def gather_table():
        if dist is None:
            return
        pad = (total + world - 1) // world if cfg["scaling"] == "strong" else n
        loc = size_all if pad == n else torch.nn.functional.pad(size_all, (0, pad - n))
        allsz = torch.empty(world * pad, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(allsz, loc)
        gathered["sizes"] = allsz
        gathered["offsets"] = torch.cumsum(allsz.to(torch.int64), 0) - allsz
        gathered["pad"] = pad

    def step_device():
        if order == "enc_dec":
            ctx.encode_frames_raw(Y, U, V, W, CW, W * H, CW * CH, n, W, H, None, pk, cap, PKT_CAP, amv.LAYOUT_PACKED, off, size_all,
                                  st_e, amv.MEM_DEVICE)
            gather_table()
            ctx.decode_frames_raw(pk, cap, off, size_all, n, W, H, DY, DU, DV, W, CW, W * H, CW * CH, st_d, amv.MEM_DEVICE)
        else:
            for j, (ipk, ioff, isz, m) in enumerate(inputs):
                ctx.decode_frames_raw(ipk, ipk.numel(), ioff, isz, m, W, H, DY, DU, DV, W, CW, W * H, CW * CH, st_d, amv.MEM_DEVICE)
                ctx.encode_frames_raw(DY, DU, DV, W, CW, W * H, CW * CH, m, W, H, None, pk, cap, PKT_CAP, amv.LAYOUT_PACKED, off,
                                      size_all[j * sub: j * sub + m], st_e, amv.MEM_DEVICE)
                if aud:
                    aud(m)
            gather_table()

    ms_total, t_max, clocks, launches, kern = J.timed(step_device, args.steps)
    assert int(st_e.abs().sum().item()) == 0 and int(st_d.abs().sum().item()) == 0, "codec reported errors"
    out_pkt_bytes = int(size_all.to(torch.int64).sum().item())
    job_frames = total if cfg["scaling"] == "strong" else world * n
    value = job_frames * args.steps / (t_max / 1e3)
    table_ok = None
    if dist is not None and rank == 0:     # the concatenated table: rank r's packets start where ranks < r end
        sz_h = gathered["sizes"].cpu().numpy().astype(np.int64)
        of_h = gathered["offsets"].cpu().numpy()
        table_ok = bool(np.array_equal(of_h, np.concatenate([[0], np.cumsum(sz_h)[:-1]]))) and \
            bool(np.array_equal(sz_h[:n], size_all.cpu().numpy().astype(np.int64))) and int(sz_h.sum()) > 0

    # ---- end-to-end leg: pinned host buffers through AMV_MEM_HOST, two contexts (one per host thread) as a two-stage pipeline
    # over the step's sub-batches: while sub-batch j runs its second stage, sub-batch j+1 already runs its first, so both PCIe
    # directions stay busy.  Every step still copies all of its inputs in and all of its results out inside the timed region.
    import threading
    ne = min(args.e2e_frames if args.e2e_frames > 0 else cfg["e2e_frames"], n)
    nsub_e = max(1, min(args.e2e_sub, ne))
    while ne % nsub_e:
        nsub_e -= 1
    ns_ = ne // nsub_e
    hcap = ns_ * cfg["cap_per_frame"]
    del DY, DU, DV
    torch.cuda.empty_cache()
    ctx2 = J.new_ctx()

    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype).pin_memory()

    if order == "enc_dec":
        hY, hU, hV = pinned((ne, H, W), torch.uint8), pinned((ne, CH, CW), torch.uint8), pinned((ne, CH, CW), torch.uint8)
        hY.copy_(Y[:ne]); hU.copy_(U[:ne]); hV.copy_(V[:ne])
    else:
        ipk, ioff, isz, _ = inputs[0]
        used = int(isz[:ne].to(torch.int64).sum().item())
        h_ipk = pinned(used, torch.uint8); h_ipk.copy_(ipk[:used])
        h_ioff = pinned(ne, torch.int64); h_ioff.copy_(ioff[:ne])
        h_isz = pinned(ne, torch.int32); h_isz.copy_(isz[:ne])
        h_ist = pinned(ne, torch.int32)
        h_opk = pinned(ne * cfg["cap_per_frame"], torch.uint8)
        h_ooff, h_osz, h_ost = pinned(ne, torch.int64), pinned(ne, torch.int32), pinned(ne, torch.int32)
    hDY, hDU, hDV = pinned((ne, H, W), torch.uint8), pinned((ne, CH, CW), torch.uint8), pinned((ne, CH, CW), torch.uint8)
    hst2 = pinned(ne, torch.int32)
    hst2.zero_()
    if order == "enc_dec":
        bufs = [dict(pk=pinned(hcap, torch.uint8), off=pinned(ns_, torch.int64), sz=pinned(ns_, torch.int32), st=pinned(ns_, torch.int32),
                     full=threading.Semaphore(0), free=threading.Semaphore(1)) for _ in range(3)]
    else:       # decoded planes travel between the stages
        bufs = [dict(y=pinned((ns_, H, W), torch.uint8), u=pinned((ns_, CH, CW), torch.uint8), v=pinned((ns_, CH, CW), torch.uint8),
                     full=threading.Semaphore(0), free=threading.Semaphore(1)) for _ in range(3)]
    step_pkt_bytes = [0]

    def stage1(b, j):
        sl = slice(j * ns_, (j + 1) * ns_)
        if order == "enc_dec":
            ctx.encode_frames_raw(hY[sl], hU[sl], hV[sl], W, CW, W * H, CW * CH, ns_, W, H, None, b["pk"], hcap, PKT_CAP,
                                  amv.LAYOUT_PACKED, b["off"], b["sz"], b["st"], amv.MEM_HOST)
        else:
            ctx.decode_frames_raw(h_ipk, h_ipk.numel(), h_ioff[sl], h_isz[sl], ns_, W, H, b["y"], b["u"], b["v"], W, CW, W * H, CW * CH,
                                  h_ist[sl], amv.MEM_HOST)

    def stage2(b, j):
        sl = slice(j * ns_, (j + 1) * ns_)
        if order == "enc_dec":
            ctx2.decode_frames_raw(b["pk"], hcap, b["off"], b["sz"], ns_, W, H, hDY[sl], hDU[sl], hDV[sl], W, CW, W * H, CW * CH, hst2[sl],
                                   amv.MEM_HOST)
        else:
            o = j * ns_ * cfg["cap_per_frame"]
            ctx2.encode_frames_raw(b["y"], b["u"], b["v"], W, CW, W * H, CW * CH, ns_, W, H, None, h_opk[o: o + hcap], hcap, PKT_CAP,
                                   amv.LAYOUT_PACKED, h_ooff[sl], h_osz[sl], h_ost[sl], amv.MEM_HOST)

    def run_steps(k, count_bytes=False):
        err = []

        def first():
            try:
                for s_ in range(k * nsub_e):
                    b = bufs[s_ % 3]
                    b["free"].acquire()
                    stage1(b, s_ % nsub_e)
                    b["full"].release()
            except Exception as e:          # noqa: BLE001
                err.append(e)
                for b in bufs:
                    b["full"].release()

        th = threading.Thread(target=first)
        th.start()
        for s_ in range(k * nsub_e):
            b = bufs[s_ % 3]
            b["full"].acquire()
            if err:
                break
            if count_bytes and order == "enc_dec":
                assert int(b["st"].abs().sum().item()) == 0, "codec reported errors (host path, encode)"
                if s_ < nsub_e:
                    step_pkt_bytes[0] += int(b["sz"].to(torch.int64).sum().item())
            stage2(b, s_ % nsub_e)
            b["free"].release()
        th.join()
        if err:
            raise err[0]

    run_steps(2, count_bytes=True)
    if order == "enc_dec":
        assert int(hst2.abs().sum().item()) == 0, "codec reported errors (host path)"
        e_in_pkt = e_out_pkt = step_pkt_bytes[0]
    else:
        assert int(h_ist.abs().sum().item()) == 0 and int(h_ost.abs().sum().item()) == 0, "codec reported errors (host path)"
        e_in_pkt, e_out_pkt = int(h_isz.to(torch.int64).sum().item()), int(h_osz.to(torch.int64).sum().item())
    J.barrier()
    t0 = time.perf_counter()
    run_steps(args.steps)
    torch.cuda.synchronize(dev)
    e2e_s = J.max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * ne * args.steps / e2e_s
    if order == "enc_dec":   # frames in + packets, offsets, sizes back in for the decode; packets + metadata + planes + status out
        h2d = ne * FRAME_BYTES + e_in_pkt + ne * 12
        d2h = e_out_pkt + ne * 16 + ne * FRAME_BYTES + ne * 4
    else:                    # packets + table in, planes out; planes in, packets + table out
        h2d = e_in_pkt + ne * 12 + ne * FRAME_BYTES
        d2h = ne * FRAME_BYTES + ne * 4 + e_out_pkt + ne * 16

    # ---- audit (outside every timed region, on EVERY rank): the library that was just timed must reproduce the committed
    # golden vectors of the reference (tests/golden/amv_golden.npz: reference-made packets and planes) bit for bit, and the
    # timed host path must round-trip its own packets deterministically
    audit = {"frames": 0, "ok": None}
    if args.audit > 0:
        try:
            G = np.load(os.path.join(ROOT, "tests", "golden", "amv_golden.npz"))
            case = "sinus_160x120_q0"
            gy, gu, gv = G[case + "/y"], G[case + "/u"], G[case + "/v"]
            apk, aoff, asz, ast_ = ctx.encode_frames(gy, gu, gv)
            ok = bool((ast_ == 0).all()) and np.array_equal(asz, G[case + "/sz"]) and np.array_equal(apk, G[case + "/pk"])
            ay, au, av, dst = ctx2.decode_frames(G[case + "/pk"], G[case + "/off"], G[case + "/sz"], 160, 120)
            ok = ok and bool((dst == 0).all()) and np.array_equal(ay, G[case + "/dy"]) and np.array_equal(au, G[case + "/du"]) \
                and np.array_equal(av, G[case + "/dv"])
            na = 0
            if order == "enc_dec":
                # the bench frames themselves: their packets (encoded once more, outside the timed region) decode to the planes
                # the timed run wrote
                na = min(args.audit, ns_)
                b = bufs[0]
                ctx.encode_frames_raw(hY[:ns_], hU[:ns_], hV[:ns_], W, CW, W * H, CW * CH, ns_, W, H, None, b["pk"], hcap, PKT_CAP,
                                      amv.LAYOUT_PACKED, b["off"], b["sz"], b["st"], amv.MEM_HOST)
                sub_sz = b["sz"][:na].numpy().astype(np.uint32)
                sub_off = b["off"][:na].numpy().astype(np.uint64)
                sub_pk = b["pk"][: int(sub_off[-1]) + int(sub_sz[-1])].numpy()
                ry, ru, rv, rst = ctx2.decode_frames(sub_pk, sub_off, sub_sz, W, H)
                ok = ok and bool((rst == 0).all()) and np.array_equal(ry, hDY[:na].numpy()) and np.array_equal(ru, hDU[:na].numpy()) \
                    and np.array_equal(rv, hDV[:na].numpy())
            audit = {"frames": int(len(asz) + na), "ok": bool(ok), "against": "tests/golden/amv_golden.npz + self round trip"}
        except Exception as e:      # the audit never decides the timing; report and go on
            audit = {"frames": 0, "ok": None, "error": str(e)[:200]}
        if dist is not None and audit["ok"] is not None:
            audit["ok"] = J.all_ok(audit["ok"])
            audit["ranks"] = world
    J.finish()
    if rank != 0:
        return 0

    peak, peak_src = load_peaks()
    dec_pkt = (out_pkt_bytes if order == "enc_dec" else in_pkt_bytes) / n       # packets the decode direction reads, per frame
    enc_pkt = out_pkt_bytes / n
    audio_bytes = 3453 if cfg["audio"] else 0                                   # (8 + n) + 4 n per direction
    bpf = {"encode": FRAME_BYTES + enc_pkt, "decode": FRAME_BYTES + dec_pkt}
    traffic = load_traffic() if args.config == 2 else {}
    frames_per_step = n
    names = {"encode": "k_encode16v2 (+ k_encode for the frames it hands back; one pair per launch)", "tokens": "k_vlc_tokens_lean",
             "sync": "k_vlc_sync", "idct": "k_idct16"}
    if dict(J.opts).get("encode_rounds", 4) == 1:
        names["encode"] = "k_encode16 (+ k_encode for the frames it hands back)"
    tp = dict(J.opts).get("decode_token_pass", 2)
    if tp != 2:
        names["tokens"], names["idct"] = ("k_vlc_tokens_lean<32-bit>", "k_idct") if tp == 1 else ("k_vlc_tokens", "k_idct")

    def roof_of(k):
        """algorithmic bytes of the frames kernel k processes per step / its device time per step"""
        ms = kern[k]["ms"] / args.steps
        b = bpf["encode" if k in ("encode", "compact") else "decode"]
        gbs = b * frames_per_step / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        tr = traffic.get(k, {}).get("dram_bytes_per_frame")
        return {"bound": "hbm", "kernel": names.get(k, "k_" + k), "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                "peak_source": peak_src, "peak_nominal": PEAK_NOMINAL, "frac_of_nominal": gbs / PEAK_NOMINAL,
                "traffic": tr * frames_per_step if tr else None,
                "traffic_note": "ncu dram bytes per frame (profiles/traffic.json) x frames per step" if tr else None,
                "algorithmic_bytes_per_step": b * frames_per_step, "frames_per_step": frames_per_step,
                "launches_per_step": kern[k]["launches"] / args.steps, "ms_per_step": ms, "bytes_per_frame": b}

    dom = max(("encode", "tokens", "idct"), key=lambda k: kern[k]["ms"])
    enc_dir_ms = (kern["encode"]["ms"] + kern["compact"]["ms"]) / args.steps
    dec_dir_ms = (kern["idct"]["ms"] + kern["tokens"]["ms"] + kern["unstuff"]["ms"] + kern["sync"]["ms"]) / args.steps
    step_ms = ms_total / args.steps
    step_bytes = (bpf["encode"] + bpf["decode"] + 2 * audio_bytes) * frames_per_step

    def direction(ms, b):
        gbs = b * frames_per_step / (ms / 1e3) / 1e9
        return {"achieved": gbs, "unit": "GB/s", "frac": gbs / peak, "frac_of_nominal": gbs / PEAK_NOMINAL}

    line = {
        "metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": world, "steps": args.steps, "warmup": J.warm,
        "ms_per_step": t_max / args.steps, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": cfg["name"] % dict(n=n, total=job_frames, sub=sub), "frames_per_gpu": n, "job_frames": job_frames,
                   "width": W, "height": H, "qscale": 2, "avg_packet_bytes": enc_pkt,
                   "l2": "inputs per step (%.1f GB) far exceed the 126 MB L2; no flush needed" % (n * FRAME_BYTES / 1e9),
                   "sharding": "contiguous frame range per GPU, no collective on the data path; the step ends its encode with the "
                               "all-gather of the ranks' packet sizes (NCCL, 4 B per frame) + prefix sum = packet table of the split job"},
        "encode_fps_per_gpu": n / (enc_dir_ms / 1e3),
        "decode_fps_per_gpu": n / (dec_dir_ms / 1e3),
        "direction_roofline": {"encode": direction(enc_dir_ms, bpf["encode"]), "decode": direction(dec_dir_ms, bpf["decode"])},
        "step_roofline": {"achieved": step_bytes / (step_ms / 1e3) / 1e9, "unit": "GB/s", "frac": step_bytes / (step_ms / 1e3) / 1e9 / peak,
                          "frac_of_nominal": step_bytes / (step_ms / 1e3) / 1e9 / PEAK_NOMINAL,
                          "what": "algorithmic bytes of the whole step (both directions) / step time, per GPU"},
        "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in kern.items() if v["launches"]},
        "kernel_share_of_step": {k: (v["ms"] / args.steps) / step_ms for k, v in kern.items() if v["launches"]},
        "roofline": roof_of(dom),
        "roofline_by_kernel": {k: roof_of(k) for k in ("encode", "tokens", "idct") if k != dom},
        "e2e": {"value": e2e_value, "unit": cfg["unit"], "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "frames_per_step_per_gpu": ne,
                "api": "%s, AMV_MEM_HOST, pinned buffers; two contexts pipeline each step's %d sub-batches" %
                       ("amv_encode_frames + amv_decode_frames" if order == "enc_dec" else "amv_decode_frames + amv_encode_frames", nsub_e)},
        "gpu_launches": int(launches),
        "options": dict(J.opts),
        "clocks": clocks,
        "audit": audit,
        "packet_table_ok": table_ok,
    }
    line["cpu_baseline"] = None
    if not args.no_cpu_baseline:
        per = ref_units(args, cfg)
        cb = run_cpu_reference(args.config, per, 1)
        line["cpu_baseline"] = cpu_baseline_block(args.config, cb, "%d frames per core, one pass, in memory, one single-threaded "
                                                  "reference context per core" % per)
    emit(line)
    return 0


# --------------------------------------------------------------------------- our arm: ADPCM (config 3)
def bench_adpcm(args):
    J = Job(args)
    torch, amv, dev, ctx, dist, world, rank = J.torch, J.amv, J.dev, J.ctx, J.dist, J.world, J.rank
    cfg = CONFIGS[3]
    nc = args.frames if args.frames > 0 else cfg["chunks"]
    ns = cfg["ns"]
    csz = 8 + ns // 2
    pcm = synth_pcm_torch(nc * ns, dev, 3 + rank)
    poff = torch.arange(nc, device=dev, dtype=torch.int64) * ns
    nsam = torch.full((nc,), ns, dtype=torch.int32, device=dev)
    ooff = torch.arange(nc, device=dev, dtype=torch.int64) * csz
    osz = torch.full((nc,), csz, dtype=torch.int32, device=dev)
    out = torch.zeros(nc * csz, dtype=torch.uint8, device=dev)
    so = torch.zeros(nc, dtype=torch.int16, device=dev)
    st = torch.zeros(nc, dtype=torch.int32, device=dev)
    dec = torch.zeros(nc * ns, dtype=torch.int16, device=dev)
    torch.cuda.synchronize(dev)

    def step():
        ctx.adpcm_enc_chunks_raw(pcm, nc * ns, poff, nsam, None, so, nc, out, out.numel(), ooff, st, amv.MEM_DEVICE)
        ctx.adpcm_dec_chunks_raw(out, out.numel(), ooff, osz, nc, dec, nc * ns, poff, st, amv.MEM_DEVICE)

    ms_total, t_max, clocks, launches, kern = J.timed(step, args.steps)
    assert int(st.abs().sum().item()) == 0
    value = world * nc * args.steps / (t_max / 1e3)

    # ---- e2e: host chunks through AMV_MEM_HOST
    ne = min(args.e2e_frames if args.e2e_frames > 0 else cfg["e2e_chunks"], nc)
    h_pcm = torch.empty(ne * ns, dtype=torch.int16).pin_memory(); h_pcm.copy_(pcm[: ne * ns])
    h_poff = (torch.arange(ne, dtype=torch.int64) * ns).pin_memory()
    h_nsam = torch.full((ne,), ns, dtype=torch.int32).pin_memory()
    h_ooff = (torch.arange(ne, dtype=torch.int64) * csz).pin_memory()
    h_osz = torch.full((ne,), csz, dtype=torch.int32).pin_memory()
    h_out = torch.zeros(ne * csz, dtype=torch.uint8).pin_memory()
    h_so = torch.zeros(ne, dtype=torch.int16).pin_memory()
    h_st = torch.zeros(ne, dtype=torch.int32).pin_memory()
    h_dec = torch.zeros(ne * ns, dtype=torch.int16).pin_memory()

    def step_host():
        ctx.adpcm_enc_chunks_raw(h_pcm, ne * ns, h_poff, h_nsam, None, h_so, ne, h_out, h_out.numel(), h_ooff, h_st, amv.MEM_HOST)
        ctx.adpcm_dec_chunks_raw(h_out, h_out.numel(), h_ooff, h_osz, ne, h_dec, ne * ns, h_poff, h_st, amv.MEM_HOST)

    step_host(); step_host()
    assert int(h_st.abs().sum().item()) == 0
    J.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize(dev)
    e2e_s = J.max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * ne * args.steps / e2e_s
    ok_host = bool(torch.equal(h_dec, dec[: ne * ns].cpu()))

    # audit without the oracle: the committed golden chunks of the reference (tests/golden/amv_golden.npz)
    audit = {"ok": None}
    try:
        G = np.load(os.path.join(ROOT, "tests", "golden", "amv_golden.npz"))
        ok = ok_host
        for kind in ("tones", "noise", "square"):
            k = "adpcm_%s/" % kind
            gout, goff, gsz, gcons = G[k + "out"], G[k + "off"], G[k + "sz"], G[k + "cons"]
            firstc = np.array([0, len(gcons)], np.uint32)
            gpoff = np.concatenate([[0], np.cumsum(gcons)[:-1]]).astype(np.uint64)
            step0 = np.array([int(gout[2]) | (int(gout[3]) << 8)], np.int16)
            eo, _, esz, _, est = ctx.adpcm_encode_streams(G[k + "src"], gpoff, gcons, firstc, step0)
            dp, _, dst = ctx.adpcm_decode(gout, goff, gsz)
            ok = ok and bool((est == 0).all()) and np.array_equal(eo, gout) and bool((dst == 0).all()) and np.array_equal(dp, G[k + "dec"])
        audit = {"ok": J.all_ok(ok), "against": "tests/golden/amv_golden.npz (reference-made chunks) + host path equals device path"}
    except Exception as e:      # noqa: BLE001
        audit = {"ok": None, "error": str(e)[:200]}
    J.finish()
    if rank != 0:
        return 0
    peak, peak_src = load_peaks()
    bpc = 5 * (ns // 2) + 8                             # 4 n + (8 + n) per direction

    def roof(k, name):
        ms = kern[k]["ms"] / args.steps
        gbs = bpc * nc / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        return {"bound": "hbm", "kernel": name, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": peak_src,
                "peak_nominal": PEAK_NOMINAL, "frac_of_nominal": gbs / PEAK_NOMINAL, "traffic": None,
                "algorithmic_bytes_per_step": bpc * nc, "chunks_per_step": nc, "ms_per_step": ms, "bytes_per_chunk": bpc}

    dom = "adpcm_enc" if kern["adpcm_enc"]["ms"] >= kern["adpcm_dec"]["ms"] else "adpcm_dec"
    other = "adpcm_dec" if dom == "adpcm_enc" else "adpcm_enc"
    nm = {"adpcm_enc": "k_adpcm_encode", "adpcm_dec": "k_adpcm_decode"}
    line = {
        "metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": world, "steps": args.steps, "warmup": J.warm,
        "ms_per_step": t_max / args.steps, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
        "dtype": "s16", "data": "synthetic",
        "config": {"workload": cfg["name"] % dict(n=nc), "chunks_per_gpu": nc, "samples_per_chunk": ns,
                   "l2": "inputs per step (%.1f GB) far exceed the 126 MB L2; no flush needed" % (nc * ns * 2 / 1e9),
                   "sharding": "contiguous chunk range per GPU, independent chunks, no collective on the data path"},
        "encode_chunks_per_s_per_gpu": nc / (kern["adpcm_enc"]["ms"] / args.steps / 1e3),
        "decode_chunks_per_s_per_gpu": nc / (kern["adpcm_dec"]["ms"] / args.steps / 1e3),
        "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in kern.items() if v["launches"]},
        "roofline": roof(dom, nm[dom]), "roofline_by_kernel": {other: roof(other, nm[other])},
        "e2e": {"value": e2e_value, "unit": cfg["unit"], "h2d_bytes_per_step": int(ne * (ns * 2 + 12 + 8) + ne * (csz + 12 + 8)),
                "d2h_bytes_per_step": int(ne * (csz + 2 + 4) + ne * (ns * 2 + 4)), "chunks_per_step_per_gpu": ne,
                "api": "amv_adpcm_enc_chunks + amv_adpcm_dec_chunks, AMV_MEM_HOST, pinned buffers"},
        "gpu_launches": int(launches), "options": dict(J.opts), "clocks": clocks, "audit": audit,
    }
    line["cpu_baseline"] = None
    if not args.no_cpu_baseline:
        per = ref_units(args, cfg)
        cb = run_cpu_reference(3, per, 1)
        line["cpu_baseline"] = cpu_baseline_block(3, cb, "%d chunks per core, encode + decode, in memory" % per)
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)      # the e2e leg is a two-stage pipeline over the steps: K + 1 stage times for K steps
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="amvcuda", choices=["amvcuda", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configuration (default 2: the headline)")
    ap.add_argument("--frames", type=int, default=0, help="units per GPU per step (config 4: of the whole job); 0 = the configuration's size")
    ap.add_argument("--e2e-frames", type=int, default=0, help="units per GPU per step of the host-buffer (e2e) leg; 0 = the configuration's")
    ap.add_argument("--e2e-sub", type=int, default=1, help="sub-batches a step of the e2e leg travels in (measured: 1 is best, every host call has a fixed cost)")
    ap.add_argument("--ref-frames-per-worker", type=int, default=0, help="units per host core of the CPU reference sample; 0 = the configuration's")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--audit", type=int, default=64, help="frames re-checked after the run (golden vectors + self round trip)")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="amv_set_option on every context (A/B runs of kernel variants, e.g. encode_rounds=1, decode_token_pass=0)")
    args = ap.parse_args()
    protect_stdout()
    if args.impl == "reference":
        return reference_arm(args)
    if CONFIGS[args.config]["kind"] == "adpcm":
        return bench_adpcm(args)
    return bench_video(args)


if __name__ == "__main__":
    sys.exit(main())
