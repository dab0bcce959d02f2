#!/usr/bin/env python
"""bench.py — events/s of the X-maps per-event depth path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): synthetic Poisson event stream, 5 M events per projector
frame, camera 640x480, projector 720x1280, ESL_calib_hhi calibration, projector-view depth frames.
One "step" = one pass of the hot path over a batch of `--frames` (default 64) distinct frames that
are resident in HBM (64 x 80 MB = 5.1 GB, far larger than the 126 MB L2, so every step streams its
events from HBM).  `value` = events of all ranks / device time of the K timed steps (CUDA events,
max over ranks).  With N > 1 every rank renders its own frames (frame f -> rank f % N, weak
scaling) and the step includes the NCCL gather of all depth frames on rank 0.

Extra JSON keys: `roofline` (the per-event kernel K1 against the measured HBM peak), `cpu_baseline`
(the NumPy oracle port timed on this box's host cores), `e2e` (host buffers in pinned memory ->
depth frames in pinned memory through `HostFrameStream`), `gpu_launches`, `clocks`.

`--impl reference` times the CPU restatement of the reference path (oracle/xmaps_oracle.py — the
reference itself is pure Python/NumPy and cannot travel to the GPU box) on all host cores it can
use, one process per core, each step = one 5 M-event frame per worker.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

CAM_W, CAM_H, PROJ_W, PROJ_H = 640, 480, 720, 1280
EVENTS_PER_FRAME = 5_000_000
FRAME_US = 16_666
METRIC = "events/sec"
WORKLOAD = "synthetic Poisson stream, 5M events/frame @ 60 fps, cam 640x480, proj 720x1280, ESL_calib_hhi, projector-view depth"


def load_tables():
    from xm_helpers import load_golden_tables

    return load_golden_tables("default")[0]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# synthetic frames
# ------------------------------------------------------------------------------------------------
def synth_frame_cuda(seed, n, device):
    """Uniform-pixel, time-sorted frame of 16-byte EventCD records generated on the GPU
    (homogeneous Poisson process conditioned on N, 90 % positive polarity)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.randint(0, CAM_W, (n,), generator=g, device=device, dtype=torch.int32)
    y = torch.randint(0, CAM_H, (n,), generator=g, device=device, dtype=torch.int32)
    p = (torch.rand(n, generator=g, device=device) < 0.9).to(torch.int32)
    t = torch.sort(torch.randint(0, FRAME_US, (n,), generator=g, device=device, dtype=torch.int64)).values
    raw = torch.empty((n, 4), dtype=torch.int32, device=device)
    raw[:, 0] = x | (y << 16)
    raw[:, 1] = p
    raw.view(torch.int64)[:, 1] = t + seed * FRAME_US
    return raw


def host_frame(raw):
    from xmaps_b200.events import EVENT_DTYPE

    return raw.cpu().numpy().view(EVENT_DTYPE).reshape(-1)


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline (oracle port of the reference's NumPy path)
# ------------------------------------------------------------------------------------------------
_W = {}


def _worker_init(n_events):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    try:
        import cv2

        cv2.setNumThreads(1)
    except Exception:
        pass
    from oracle import xmaps_oracle as orc

    _W["orc"] = orc
    _W["tables"] = load_tables()
    _W["n"] = n_events
    _W["ev"] = None


def _worker_step(seed):
    orc = _W["orc"]
    if _W["ev"] is None:
        _W["ev"] = orc.synth_events(1000 + seed, _W["n"], CAM_W, CAM_H, frame_us=FRAME_US)
    t0 = time.perf_counter()
    depth = orc.frame_depth(_W["tables"], _W["ev"], orc.VIEW_PROJECTOR)
    return time.perf_counter() - t0, float(depth[::97, ::89].sum())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    workers = max(1, min(cores, args.ref_workers))
    n = args.events
    ctx = mp.get_context("fork")
    with ctx.Pool(workers, initializer=_worker_init, initargs=(n,)) as pool:
        for _ in range(max(1, args.warmup)):
            pool.map(_worker_step, range(workers))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_worker_step, range(workers))
        dt = time.perf_counter() - t0
    value = workers * n * args.steps / dt
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": "events/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int16/f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "events_per_frame": n, "frames_per_step": workers},
        "cpu_baseline": {
            "value": value,
            "unit": "events/s",
            "cores": workers,
            "kind": "port",
            "sample": f"{workers} worker processes x 1 frame of {n} events per step (NumPy/OpenCV oracle port, one thread each)",
        },
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(ev_host, tables, runs=5):
    from oracle import xmaps_oracle as orc

    try:
        import cv2

        cv2.setNumThreads(1)
    except Exception:
        pass
    times = []
    depth = None
    for _ in range(runs):
        t0 = time.perf_counter()
        depth = orc.frame_depth(tables, ev_host, orc.VIEW_PROJECTOR)
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return depth, {
        "value": len(ev_host) / med,
        "unit": "events/s",
        "cores": 1,
        "kind": "port",
        "sample": f"1 frame of {len(ev_host)} events, median of {runs} runs, single thread (the reference's per-event stages are single-threaded NumPy)",
        "ms_per_frame": med * 1e3,
    }


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import xmaps_b200  # noqa: F401
    from xmaps_b200 import _native as N
    from xmaps_b200.engine import OUT_DEPTH, VIEW_PROJECTOR, DepthEngine, TableSet
    from xmaps_b200.host_stream import HostFrameStream
    from xmaps_b200.sharding import FrameSharder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    tables = load_tables()
    eng = DepthEngine(
        TableSet(
            lut_x=tables.lut_x, lut_y=tables.lut_y, x_map=tables.x_map, remap_xy=tables.remap_xy,
            rect_w=tables.rect_w, rect_h=tables.rect_h, t_px_scale=tables.t_px_scale, x_offset=tables.x_offset,
            depth_scale=tables.depth_scale,
        ),
        device=device,
    )
    if world > 1 and not args.no_gather:
        # the persistent batch kernel would otherwise own every SM until it ends and NCCL's copy kernels
        # (the gather of the previous chunk) could not run beside it
        eng.set_option("reserve_sms", args.reserve_sms)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))

    n, F = args.events, args.frames
    # global frame index of local frame j on this rank = j * world + rank (round-robin sharding)
    frames = [synth_frame_cuda(j * world + rank, n, device) for j in range(F)]
    out = torch.empty((F, PROJ_H, PROJ_W), dtype=torch.float32, device=device)
    gathered = None
    if world > 1 and rank == 0 and not args.no_gather:
        gathered = [out] + [torch.empty_like(out) for _ in range(world - 1)]

    def render(fr, dst):
        eng.frame_batch(fr, view=VIEW_PROJECTOR, output=OUT_DEPTH, out=dst)

    # frames per render call = per batch-kernel launch.  With the NCCL gather: 16 (measured on 2 GPUs: 219-226 G
    # events/s; a tapering schedule with 4 SMs left to NCCL measured 217 G, so both stay options)
    if args.gather_chunk > 0:
        chunk, schedule = args.gather_chunk, [args.gather_chunk]
    elif world == 1 or args.no_gather:
        chunk, schedule = 32, [32]
    else:
        chunk, schedule = 16, [16]
        if args.taper:  # 16, 16, 16, 8, 4, 4 for 64 frames
            schedule, left = [], F
            while left > 32:
                schedule.append(16)
                left -= 16
            schedule += [s for s in (16, 8, 4, 4) if s <= left] if left == 32 else [left]
            if sum(schedule) != F:
                schedule = [16]
    sharder = FrameSharder(render, rank, world, dst=0, chunk=schedule)
    gather_mode = "none" if (world == 1 or args.no_gather) else "nccl"
    if gather_mode == "nccl" and args.gather_mode == "p2p":
        gather_mode = "p2p" if sharder.enable_peer_copies(gathered) else "nccl (p2p mapping failed)"

    def step():
        sharder.run(frames, out, gathered, gather=(world > 1 and not args.no_gather))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(max(3, args.warmup)):
        step()
    barrier()

    # ---- timed region A: the headline number --------------------------------------------------
    sampler = ClockSampler(local_rank)
    launches0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = N.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = world * F * n * args.steps / (ms * 1e-3)

    # ---- timed region B: same steps with CUDA events around K1 / K2 of every frame --------------
    eng.set_option("profile", -1)
    eng.set_option("profile", 1)
    torch.cuda.synchronize(device)
    for _ in range(args.steps):
        for lo, hi in sharder._spans(F):
            render(frames[lo:hi], out[lo:hi])
    torch.cuda.synchronize(device)
    k1_ns, k2_ns, pf = eng.get_option("profile_k1_ns"), eng.get_option("profile_k2_ns"), eng.get_option("profile_frames")
    eng.set_option("profile", 0)
    k1_us = k1_ns / max(1, pf) / 1e3
    k2_us = k2_ns / max(1, pf) / 1e3

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    # algorithmic bytes of K1 per launch: 16 B per input event + the two tables it
    # gathers from (packed rectify LUT, transposed X-map), each read once per frame
    lut_bytes = CAM_W * CAM_H * 4
    xmap_bytes = tables.x_map.size * 2
    k1_bytes = 16 * n + lut_bytes + xmap_bytes
    kernel_name = "xm::events_lean_kernel (K1: polarity + rectify LUT + X-map lookup + disparity + scatter)"
    prof_launches = max(1, eng.get_option("profile_launches"))
    frames_per_launch = max(1, round(pf / prof_launches))
    # one kernel per frame (fused frame kernel) or per chunk of frames (batch kernel)?
    fused = launches <= 1.5 * F * args.steps
    batch = fused and frames_per_launch > 1
    traffic_file = "k1_dram_bytes.json"
    if fused:
        # whole-frame kernels: K1's bytes + the remap table read + the depth frame written (B(N_in))
        k1_bytes += PROJ_W * PROJ_H * 4 * 2
        traffic_file = "frame_dram_bytes.json"
        if batch:
            kernel_name = "xm::batch_kernel (persistent: event warps + epilogue warp groups, %d frames per launch)" % frames_per_launch
        else:
            kernel_name = "xm::frame_kernel (whole frame: per-event phase + grid barrier + dilate/remap/depth epilogue)"
    # duration of one launch of the dominant kernel: CUDA events around every launch (region B).  Consecutive
    # frame_kernel launches overlap through programmatic dependent launch, which the event records break,
    # so for that kernel the timed region divided by its launches is the in-step figure.
    bytes_per_launch = k1_bytes * frames_per_launch
    us_isolated = k1_us * frames_per_launch
    us_per_launch = ms * 1e3 / (F * args.steps) if (fused and not batch) else us_isolated
    achieved = bytes_per_launch / (us_per_launch * 1e-6) / 1e9 if us_per_launch > 0 else 0.0
    roofline = {
        "bound": "hbm",
        "kernel": kernel_name,
        "achieved": achieved,
        "peak": peak,
        "peak_source": peak_src,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": None,
        "bytes_per_launch": bytes_per_launch,
        "us_per_launch": us_per_launch,
        "us_per_launch_isolated": us_isolated,
        "frames_per_launch": frames_per_launch,
        "algorithmic_bytes_per_frame": k1_bytes,
        "k2_us_per_launch": k2_us,
        "frame_us": ms * 1e3 / (F * args.steps),
    }
    tf = os.path.join(ROOT, "profiles", traffic_file)
    if os.path.exists(tf):
        try:
            with open(tf) as fh:
                tj = json.load(fh)
            per_frame = tj.get("dram_bytes_per_frame")
            roofline["traffic"] = per_frame * frames_per_launch if (batch and per_frame) else tj.get("dram_bytes_per_launch")
            roofline["traffic_source"] = tj.get("source")
        except Exception:
            pass

    if args.quick:  # parameter sweeps: kernel numbers only
        if world > 1:
            dist.barrier()  # (the other ranks wait here before they leave)
            dist.destroy_process_group()
        print(json.dumps({"quick": True, "gather": gather_mode, "options": args.opt, "value": value, "frame_us": roofline["frame_us"],
                          "k1_us": k1_us, "k2_us": k2_us, "k1_frac": roofline["frac"], "frames_per_launch": frames_per_launch, "events": n, "frames": F}), flush=True)
        return

    # ---- parity spot check + CPU baseline on one frame of the same workload -------------------
    ev_host = host_frame(frames[0])
    want, cpu = cpu_baseline_sample(ev_host, tables, runs=args.cpu_runs)
    got = eng.frame(frames[0], view=VIEW_PROJECTOR, output=OUT_DEPTH).cpu().numpy()
    mismatches = int(np.count_nonzero(got != want))

    # ---- e2e: pinned host events -> pinned host depth frames -----------------------------------
    e2e_frames = min(F, args.e2e_frames)
    host_frames = [frames[j].cpu().pin_memory() for j in range(e2e_frames)]
    hs = HostFrameStream(eng, n, view=VIEW_PROJECTOR, output=OUT_DEPTH, depth=3)
    host_out = hs.alloc_outputs(e2e_frames)
    hs.run(host_frames, host_out)  # warm-up
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    reps = max(1, args.e2e_reps)
    for _ in range(reps):
        hs.run(host_frames, host_out)
    torch.cuda.synchronize(device)
    e2e_dt = time.perf_counter() - t0
    e2e_ok = bool(np.array_equal(host_out[0].numpy(), want))
    e2e = {
        "value": e2e_frames * reps * n / e2e_dt,
        "unit": "events/s",
        "h2d_bytes_per_step": e2e_frames * n * 16,
        "d2h_bytes_per_step": e2e_frames * PROJ_H * PROJ_W * 4,
        "frames_per_step": e2e_frames,
        "api": "xmaps_b200.host_stream.HostFrameStream.run (pinned host EventCD buffers -> pinned host depth frames)",
        "matches_oracle": e2e_ok,
    }

    line = {
        "metric": METRIC,
        "value": value,
        "unit": "events/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int16 (+ f64 time normalisation, f64->f32 depth)",
        "data": "synthetic",
        "config": {
            "workload": WORKLOAD,
            "events_per_frame": n,
            "frames_per_step_per_gpu": F,
            "l2": "inputs larger than L2 (%.1f GB of events per step per GPU)" % (F * n * 16 / 1e9),
            "parallelism": "frames round-robin over %d GPU(s)%s" % (world, "" if world == 1 or args.no_gather else ", gather of depth frames to rank 0 inside the step (%s)" % gather_mode),
            "time_bounds": "sorted + device-side verification and fix-up",
            "options": args.opt,
        },
        "frames_per_sec": value / n,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "parity": {"mismatching_pixels": mismatches, "checked": "frame 0 of the bench batch vs the oracle, bit-exact"},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--events", type=int, default=EVENTS_PER_FRAME)
    ap.add_argument("--frames", type=int, default=64, help="distinct frames per step per GPU")
    ap.add_argument("--gather-chunk", type=int, default=0, help="frames per render call (one persistent kernel each); 0 = 32 on one GPU, 16 with the NCCL gather")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--reserve-sms", type=int, default=0, help="SMs left to NCCL while the batch kernel runs (N > 1 with the gather)")
    ap.add_argument("--gather-mode", default="nccl", choices=["nccl", "p2p"], help="p2p: copy-engine peer copies into rank 0's buffer (CUDA IPC)")
    ap.add_argument("--taper", action="store_true", help="tapering render schedule (16, 16, 16, 8, 4, 4) with the NCCL gather")
    ap.add_argument("--cpu-runs", type=int, default=5)
    ap.add_argument("--e2e-frames", type=int, default=16)
    ap.add_argument("--e2e-reps", type=int, default=3)
    ap.add_argument("--ref-workers", type=int, default=32)
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (repeatable)")
    ap.add_argument("--quick", action="store_true", help="print kernel timings only (no CPU baseline / e2e); single GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
