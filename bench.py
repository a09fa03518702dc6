#!/usr/bin/env python
"""bench.py -- AMV 320x240 frames/sec enc+dec on B200(s), with HBM roofline and CPU reference.

    python bench.py --gpus N --steps K --warmup W            # our arm (libamvcuda)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU codecs on the host cores

A step = one pass of the hot path over one batch of synthetic frames: amv_encode_frames over the
batch (YUVJ420P planes -> packed AMV packets) followed by amv_decode_frames over those packets
(-> planes), both through the C ABI.  `value` is frames / (encode time + decode time) with every
buffer resident in HBM; `e2e` is the same round trip with pinned HOST buffers (AMV_MEM_HOST:
H2D of the inputs and D2H of the results inside the timed region).  Frames are intra-only, so
with N GPUs every rank owns its own contiguous frame range (no collective on the data path; the
only torch.distributed traffic is the barrier and the max-reduce of the timings): weak scaling.
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

W, H = 320, 240
CW, CH = W // 2, H // 2
FRAME_BYTES = W * H * 3 // 2
METRIC = "AMV 320x240 frames/sec enc+dec"
PKT_CAP = 65536                     # per-frame packet capacity handed to the encoder
KERNELS = ("encode", "compact", "unstuff", "sync", "tokens", "idct")


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of each hot kernel from the committed ncu --set full capture
    (profiles/traffic.json, bytes per frame of that capture); scaled to the frames one launch processes here."""
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


# --------------------------------------------------------------------------- synthetic frames
def synth_frames_torch(n, t0, device, seed):
    """SURVEY 8d generator on the device: noisy sinusoids, YUVJ420P full range."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    Y = torch.empty((n, H, W), dtype=torch.uint8, device=device)
    U = torch.empty((n, CH, CW), dtype=torch.uint8, device=device)
    V = torch.empty((n, CH, CW), dtype=torch.uint8, device=device)
    xx = torch.arange(W, device=device, dtype=torch.float32)[None, None, :]
    yy = torch.arange(H, device=device, dtype=torch.float32)[None, :, None]
    cx = torch.arange(CW, device=device, dtype=torch.float32)[None, None, :]
    cy = torch.arange(CH, device=device, dtype=torch.float32)[None, :, None]
    step = 2048
    for a in range(0, n, step):
        b = min(n, a + step)
        t = (torch.arange(a, b, device=device, dtype=torch.float32) + t0)[:, None, None]
        y = 128 + 60 * torch.sin((xx + 3 * t) / 17.0) + 50 * torch.cos((yy - 2 * t) / 11.0)
        y = y + 6.0 * torch.randn((b - a, H, W), device=device, generator=g)
        Y[a:b] = y.round().clamp(0, 255).to(torch.uint8)
        U[a:b] = (128 + 40 * torch.sin((cx + t) / 23.0) + 0 * cy).round().clamp(0, 255).to(torch.uint8)
        V[a:b] = (128 + 40 * torch.cos((cy + t) / 19.0) + 0 * cx).round().clamp(0, 255).to(torch.uint8)
    return Y, U, V


def synth_frames_numpy(n, t0, seed):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import synth_frames
    return synth_frames(n, W, H, seed=seed, t0=t0)


# --------------------------------------------------------------------------- reference arm
def _ref_worker(args):
    """One process = one single-threaded reference codec context pair (the reference rejects
    thread_count > 1 for AMV, mpegvideo_enc.c:451-456).  Returns (frames, seconds, packet_bytes)."""
    wid, nframes, rounds = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, Ref
    kind = "reference" if Ref.available() else "port"
    y, u, v = synth_frames_numpy(nframes, 1000 * wid, 100 + wid)
    codec = Ref() if kind == "reference" else Oracle()
    t = time.perf_counter()
    nbytes = 0
    for _ in range(rounds):
        if kind == "reference":
            pk, off, sz = codec.encode_frames(y, u, v, W, H, quality=0)
            codec.decode_frames(pk, off, sz, W, H)
        else:
            pk, off, sz = codec.encode_frames(y, u, v, W, H, 2)
            codec.decode_frames(pk, off, sz, W, H)
        nbytes += int(sz.sum())
    dt = time.perf_counter() - t
    # amvlib (C-AMVDecoder) timed the same way, decode only, packets in memory (it has no encoder)
    lib_s = None
    if kind == "reference":
        from oracle_lib import AmvlibRef
        if AmvlibRef.available():
            al = AmvlibRef()
            t = time.perf_counter()
            for _ in range(rounds):
                al.video_decode(pk, off, sz, W, H)
            lib_s = time.perf_counter() - t
    return nframes * rounds, dt, nbytes, kind, lib_s


def run_cpu_reference(frames_per_worker, rounds=1, workers=None):
    import multiprocessing as mp
    workers = workers or (os.cpu_count() or 1)
    ctx = mp.get_context("spawn")
    t = time.perf_counter()
    with ctx.Pool(workers) as pool:
        res = pool.map(_ref_worker, [(i, frames_per_worker, rounds) for i in range(workers)])
    wall = time.perf_counter() - t
    frames = sum(r[0] for r in res)
    busy = max(r[1] for r in res)                      # codec time of the slowest worker (excludes spawn/synthesis)
    out = {"frames": frames, "seconds": busy, "wall": wall, "fps": frames / busy, "workers": workers,
           "kind": res[0][3], "pkt_bytes": sum(r[2] for r in res), "per_core_fps": frames / busy / workers}
    if all(r[4] for r in res):
        out["amvlib_seconds"] = max(r[4] for r in res)
    return out


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


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per = max(8, args.ref_frames_per_worker)
    vals = []
    for _ in range(args.warmup if args.warmup < 1 else 1):
        run_cpu_reference(max(8, per // 4), 1, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(run_cpu_reference(per, 1, cores))
    total_frames = sum(v["frames"] for v in vals)
    total_s = sum(v["seconds"] for v in vals)
    fps = total_frames / total_s
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "AMV 320x240 encode+decode round trip, reference CPU codecs (AMVmuxer libavcodec 51.47.1, "
                               "generic C), one single-threaded codec context per host core",
                   "frames_per_step": vals[0]["frames"], "width": W, "height": H, "qscale": 2},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": vals[0]["workers"], "kind": vals[0]["kind"],
                         "sample": "%d frames per core per step x %d steps, enc+dec, in memory" % (per, args.steps),
                         "per_core": vals[0]["per_core_fps"]},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if all("amvlib_seconds" in v for v in vals):      # the reference's second decoder (C-AMVDecoder/amvlib), decode only
        lib_s = sum(v["amvlib_seconds"] for v in vals)
        line["amvlib"] = {"value": total_frames / lib_s, "unit": "frames/s", "cores": vals[0]["workers"], "kind": "reference",
                          "what": "AmvVideoDecode of the same packets to BGR24, packets in memory, one process per core",
                          "per_core": total_frames / lib_s / vals[0]["workers"]}
    emit(line)
    return 0


# --------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)      # the e2e leg is a two-stage pipeline over the steps: K + 1 stage times for K steps
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="amvcuda", choices=["amvcuda", "reference"])
    ap.add_argument("--frames", type=int, default=100000, help="frames per GPU per step (BASELINE config 2: 100k)")
    ap.add_argument("--e2e-frames", type=int, default=16384, help="frames per GPU per step of the host-buffer (e2e) leg")
    ap.add_argument("--e2e-sub", type=int, default=1, help="sub-batches a step of the e2e leg travels in (measured: 1 is best, every host call has a fixed ~2 ms)")
    ap.add_argument("--ref-frames-per-worker", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--audit", type=int, default=64, help="frames re-checked after the run (golden vectors + self round trip)")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="amv_set_option on every context (A/B runs of kernel variants, e.g. encode_rounds=1, decode_tokens16=0)")
    args = ap.parse_args()
    protect_stdout()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import amv_codec_tools_b200 as amv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    warm = max(3, args.warmup)
    n = args.frames

    ctx = amv.AmvCuda(device=dev.index)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_option("profile_events", 1)
    opts = [(kv.split("=")[0], int(kv.split("=")[1])) for kv in args.opt]
    for k_, v_ in opts:
        ctx.set_option(k_, v_)

    # ---- device-resident workload: every rank its own frame range [rank*n, (rank+1)*n)
    Y, U, V = synth_frames_torch(n, rank * n, dev, seed=1 + rank)
    out_cap = n * 24 * 1024
    pk = torch.empty(out_cap, dtype=torch.uint8, device=dev)
    off = torch.zeros(n, dtype=torch.int64, device=dev)
    size = torch.zeros(n, dtype=torch.int32, device=dev)
    st_e = torch.zeros(n, dtype=torch.int32, device=dev)
    st_d = torch.zeros(n, dtype=torch.int32, device=dev)
    DY = torch.empty_like(Y); DU = torch.empty_like(U); DV = torch.empty_like(V)
    torch.cuda.synchronize(dev)

    def step_device():
        ctx.encode_frames_raw(Y, U, V, W, CW, W * H, CW * CH, n, W, H, None, pk, out_cap, PKT_CAP, amv.LAYOUT_PACKED,
                              off, size, st_e, amv.MEM_DEVICE)
        ctx.decode_frames_raw(pk, out_cap, off, size, n, W, H, DY, DU, DV, W, CW, W * H, CW * CH, st_d, amv.MEM_DEVICE)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(warm):
        step_device()
    ctx.sync()
    assert int(st_e.abs().sum().item()) == 0 and int(st_d.abs().sum().item()) == 0, "codec reported errors"
    pkt_bytes = int(size.to(torch.int64).sum().item())
    for k in KERNELS:      # drop warm-up samples
        ctx.get_stat(k + "_kernel_ns")

    sampler = ClockSampler(dev.index)
    launches0 = ctx.launch_count()
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            step_device()
        ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    kern = {}
    for k in KERNELS:
        cnt = ctx.get_stat(k + "_kernel_launches")
        ns = ctx.get_stat(k + "_kernel_ns")
        kern[k] = {"launches": cnt, "ms": ns / 1e6}
    t_max = ms_total
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_max = float(t.item())
    value = world * n * args.steps / (t_max / 1e3)

    # ---- end-to-end leg: pinned host buffers through AMV_MEM_HOST
    ne = min(args.e2e_frames, n)
    hY = torch.empty((ne, H, W), dtype=torch.uint8).pin_memory(); hY.copy_(Y[:ne])
    hU = torch.empty((ne, CH, CW), dtype=torch.uint8).pin_memory(); hU.copy_(U[:ne])
    hV = torch.empty((ne, CH, CW), dtype=torch.uint8).pin_memory(); hV.copy_(V[:ne])
    hoff = torch.zeros(ne, dtype=torch.int64).pin_memory()
    hsz = torch.zeros(ne, dtype=torch.int32).pin_memory()
    hst = torch.zeros(ne, dtype=torch.int32).pin_memory()
    hDY = torch.empty((ne, H, W), dtype=torch.uint8).pin_memory()
    hDU = torch.empty((ne, CH, CW), dtype=torch.uint8).pin_memory()
    hDV = torch.empty((ne, CH, CW), dtype=torch.uint8).pin_memory()
    del DY, DU, DV
    torch.cuda.empty_cache()

    # Two contexts, one per host thread, form a two-stage pipeline: the step's frames travel in `nsub` sub-batches,
    # and while sub-batch j's packets are decoded (D2H-heavy) sub-batch j+1 is already being encoded (H2D-heavy),
    # so both PCIe directions stay busy.  Every step still copies all of its own inputs in and all of its own
    # results out inside the timed region; packets / offsets / sizes travel between the stages through a ring of
    # three buffers.  (With whole steps as the pipeline's unit the fill and drain cost one stage in K + 1.)
    import threading
    ctx2 = amv.AmvCuda(device=dev.index)
    for k_, v_ in opts:
        ctx2.set_option(k_, v_)
    nsub = max(1, min(args.e2e_sub, ne))
    while ne % nsub:
        nsub -= 1
    ns_ = ne // nsub
    hcap = ns_ * 24 * 1024
    bufs = [dict(pk=torch.empty(hcap, dtype=torch.uint8).pin_memory(), off=torch.zeros(ns_, dtype=torch.int64).pin_memory(),
                 sz=torch.zeros(ns_, dtype=torch.int32).pin_memory(), st=torch.zeros(ns_, dtype=torch.int32).pin_memory(),
                 full=threading.Semaphore(0), free=threading.Semaphore(1)) for _ in range(3)]
    hst2 = torch.zeros(ne, dtype=torch.int32).pin_memory()
    hpk, hsz = bufs[0]["pk"], bufs[0]["sz"]
    step_pkt_bytes = [0]

    def run_steps(k, count_bytes=False):
        err = []

        def enc():
            try:
                for s_ in range(k * nsub):
                    b = bufs[s_ % 3]
                    j = s_ % nsub
                    b["free"].acquire()
                    ctx.encode_frames_raw(hY[j * ns_:(j + 1) * ns_], hU[j * ns_:(j + 1) * ns_], hV[j * ns_:(j + 1) * ns_], W, CW, W * H,
                                          CW * CH, ns_, W, H, None, b["pk"], hcap, PKT_CAP, amv.LAYOUT_PACKED, b["off"], b["sz"],
                                          b["st"], amv.MEM_HOST)
                    b["full"].release()
            except Exception as e:          # noqa: BLE001
                err.append(e)
                for b in bufs:
                    b["full"].release()

        th = threading.Thread(target=enc)
        th.start()
        for s_ in range(k * nsub):
            b = bufs[s_ % 3]
            j = s_ % nsub
            b["full"].acquire()
            if err:
                break
            if count_bytes:
                assert int(b["st"].abs().sum().item()) == 0, "codec reported errors (host path, encode)"
                if s_ < nsub:
                    step_pkt_bytes[0] += int(b["sz"].to(torch.int64).sum().item())
            ctx2.decode_frames_raw(b["pk"], hcap, b["off"], b["sz"], ns_, W, H, hDY[j * ns_:(j + 1) * ns_], hDU[j * ns_:(j + 1) * ns_],
                                   hDV[j * ns_:(j + 1) * ns_], W, CW, W * H, CW * CH, hst2[j * ns_:(j + 1) * ns_], amv.MEM_HOST)
            b["free"].release()
        th.join()
        if err:
            raise err[0]

    run_steps(2, count_bytes=True)
    assert int(hst2.abs().sum().item()) == 0, "codec reported errors (host path)"
    e_pkt = step_pkt_bytes[0]
    barrier()
    t0 = time.perf_counter()
    run_steps(args.steps)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert all(int(b["st"].abs().sum().item()) == 0 for b in bufs) and int(hst2.abs().sum().item()) == 0, \
        "codec reported errors (host path, timed run)"
    e2e_value = world * ne * args.steps / e2e_s
    h2d = ne * FRAME_BYTES + e_pkt + ne * 12          # frames (encode in) + packets, offsets, sizes (decode in)
    d2h = e_pkt + ne * (8 + 4 + 4) + ne * FRAME_BYTES + ne * 4

    # ---- audit (outside every timed region): the library that was just timed must reproduce the committed golden
    # vectors of the reference (tests/golden/amv_golden.npz: reference-made packets and planes) bit for bit, and the
    # timed host path must round-trip its own packets deterministically
    audit = {"frames": 0, "ok": None}
    if rank == 0 and args.audit > 0:
        try:
            G = np.load(os.path.join(ROOT, "tests", "golden", "amv_golden.npz"))
            case = "sinus_160x120_q0"
            gy, gu, gv = G[case + "/y"], G[case + "/u"], G[case + "/v"]
            apk, aoff, asz, ast = ctx.encode_frames(gy, gu, gv)
            ok = bool((ast == 0).all()) and np.array_equal(asz, G[case + "/sz"]) and np.array_equal(apk, G[case + "/pk"])
            ay, au, av, dst = ctx2.decode_frames(G[case + "/pk"], G[case + "/off"], G[case + "/sz"], 160, 120)
            ok = ok and bool((dst == 0).all()) and np.array_equal(ay, G[case + "/dy"]) and np.array_equal(au, G[case + "/du"]) \
                and np.array_equal(av, G[case + "/dv"])
            # the bench frames themselves: their packets (encoded once more, outside the timed region) decode to the planes the
            # timed run wrote
            na = min(args.audit, ns_)
            ctx.encode_frames_raw(hY[:ns_], hU[:ns_], hV[:ns_], W, CW, W * H, CW * CH, ns_, W, H, None, bufs[0]["pk"], hcap, PKT_CAP,
                                  amv.LAYOUT_PACKED, bufs[0]["off"], bufs[0]["sz"], bufs[0]["st"], amv.MEM_HOST)
            sub_sz = hsz[:na].numpy().astype(np.uint32)
            sub_off = bufs[0]["off"][:na].numpy().astype(np.uint64)
            sub_pk = hpk[: int(sub_off[-1]) + int(sub_sz[-1])].numpy()
            ry, ru, rv, rst = ctx2.decode_frames(sub_pk, sub_off, sub_sz, W, H)
            ok = ok and bool((rst == 0).all()) and np.array_equal(ry, hDY[:na].numpy()) and np.array_equal(ru, hDU[:na].numpy()) \
                and np.array_equal(rv, hDV[:na].numpy())
            audit = {"frames": int(len(asz) + na), "ok": bool(ok), "against": "tests/golden/amv_golden.npz + self round trip"}
        except Exception as e:      # the audit never decides the timing; report and go on
            audit = {"frames": 0, "ok": None, "error": str(e)[:200]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = load_peaks()
    bytes_per_frame = FRAME_BYTES + pkt_bytes / n          # raw planes + packet, per direction

    traffic = load_traffic()

    def roof_of(k):
        """algorithmic bytes of the frames one launch of kernel k processes / its mean launch duration"""
        ms = kern[k]["ms"] / max(1, kern[k]["launches"])
        frames = n * args.steps / max(1, kern[k]["launches"])
        gbs = bytes_per_frame * frames / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        tr = traffic.get(k, {}).get("dram_bytes_per_frame")
        names = {"encode": "k_encode16 (+ k_encode for the frames it hands back; one pair per launch)", "tokens": "k_vlc_tokens",
                 "sync": "k_vlc_sync", "idct": "k_idct"}
        return {"bound": "hbm", "kernel": names.get(k, "k_" + k), "achieved": gbs,
                "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": peak_src,
                "traffic": tr * frames if tr else None,
                "traffic_note": "ncu dram bytes per frame (profiles/traffic.json) x frames per launch" if tr else None,
                "algorithmic_bytes_per_launch": bytes_per_frame * frames, "frames_per_launch": frames,
                "ms_per_launch": ms, "bytes_per_frame": bytes_per_frame}

    dom = max(("encode", "tokens", "idct"), key=lambda k: kern[k]["ms"])
    roof = roof_of(dom)
    enc_dir_ms = (kern["encode"]["ms"] + kern["compact"]["ms"]) / args.steps
    dec_dir_ms = (kern["idct"]["ms"] + kern["tokens"]["ms"] + kern["unstuff"]["ms"] + kern["sync"]["ms"]) / args.steps
    step_ms = ms_total / args.steps
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": t_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "BASELINE config 2: %d synthetic 320x240 YUVJ420P frames per GPU, qscale 2: "
                               "amv_encode_frames (packed packets) then amv_decode_frames of those packets" % n,
                   "frames_per_gpu": n, "width": W, "height": H, "qscale": 2, "avg_packet_bytes": pkt_bytes / n,
                   "l2": "inputs per step (%.1f GB) far exceed the 126 MB L2; no flush needed" % (n * FRAME_BYTES / 1e9),
                   "sharding": "contiguous frame range per GPU, no collective on the data path"},
        "encode_fps_per_gpu": n / (enc_dir_ms / 1e3),
        "decode_fps_per_gpu": n / (dec_dir_ms / 1e3),
        "direction_roofline": {
            "encode": {"achieved": bytes_per_frame * n / (enc_dir_ms / 1e3) / 1e9, "unit": "GB/s",
                       "frac": bytes_per_frame * n / (enc_dir_ms / 1e3) / 1e9 / peak},
            "decode": {"achieved": bytes_per_frame * n / (dec_dir_ms / 1e3) / 1e9, "unit": "GB/s",
                       "frac": bytes_per_frame * n / (dec_dir_ms / 1e3) / 1e9 / peak}},
        "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in kern.items()},
        "kernel_share_of_step": {k: (v["ms"] / args.steps) / step_ms for k, v in kern.items()},
        "roofline": roof,
        "roofline_by_kernel": {k: roof_of(k) for k in ("encode", "tokens", "idct") if k != dom},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "frames_per_step_per_gpu": ne, "api": "amv_encode_frames + amv_decode_frames, AMV_MEM_HOST, pinned buffers; two contexts pipeline each step's %d sub-batches (encode of sub-batch j+1 overlaps decode of sub-batch j)" % nsub},
        "gpu_launches": int(launches),
        "options": dict(opts),
        "clocks": clocks,
        "audit": audit,
    }
    if not args.no_cpu_baseline and world == 1:
        cb = run_cpu_reference(args.ref_frames_per_worker, 1)
        line["cpu_baseline"] = {"value": cb["fps"], "unit": "frames/s", "cores": cb["workers"], "kind": cb["kind"],
                                "sample": "%d frames per core, enc+dec round trip, in memory, one single-threaded reference "
                                          "context per core" % args.ref_frames_per_worker,
                                "per_core": cb["per_core_fps"]}
    else:
        line["cpu_baseline"] = None
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
