#!/usr/bin/env python
"""bench.py — events/s of the X-maps per-event depth path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload 5m|100k|hd20m|plane|sweep]

Workloads (BASELINE.json `configs`; `5m` = configs[1] is the default and the headline):
  5m     synthetic Poisson stream, 5 M events per projector frame, camera 640x480, projector 720x1280,
         ESL_calib_hhi calibration, projector-view depth frames; 64 distinct frames per step per GPU
  100k   configs[0]: the same geometry at 100 k events per frame (the reference's own CPU-runnable case)
  hd20m  configs[2]: camera 1280x720, projector 1080x1920, 20 M events per frame (nearest X-map lookup)
  plane  SURVEY §8c/§8d: the ~100 %-inlier workload -- a fronto-parallel plane, every lit camera pixel fires a
         burst (16 events by default) when its projector column passes; time-neighbours are space-neighbours
  sweep  configs[4]: 1 M ... 50 M events per frame, one entry per size in `sweep` (headline value = 5 M)

One "step" = one pass of the hot path over a batch of distinct frames that are resident in HBM and far larger
than the 126 MB L2, so every step streams its events from HBM.  `value` = events of all ranks / device time of
the K timed steps (CUDA events, max over ranks).  With N > 1 every rank renders its own frames (frame f -> rank
f % N, weak scaling) and the step includes the gather of all depth frames on rank 0 -- by default written
straight into rank 0's memory by the render kernel's epilogue warps (`--gather-mode direct`, see
x-maps_b200/sharding.py), `nccl` = dist.gather, `copy` = copy-engine peer copies.

Extra JSON keys: `roofline` (the persistent batch kernel against the measured HBM peak), `cpu_baseline` (the
reference's own modules from oracle/_ref timed on one host core; the NumPy restatement where _ref is absent),
`e2e` (pinned host buffers -> pinned host depth frames through `HostFrameStream`, on every rank, aggregate),
`parity` (frames of the TIMED batch output and, for N > 1, gathered frames of every rank against the oracle),
`gpu_launches`, `clocks`.

`--impl reference` times the reference's own implementation of the path (oracle/_ref: the unmodified
x_maps_disparity / cam_proj_calibration / disp_to_depth modules; the NumPy restatement only if _ref is absent)
on all host cores, one process per core, each step = one frame per worker.
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

FRAME_US = 16_666
METRIC = "events/sec"

# name -> (config.workload string, camera w, h, projector w, h, events per frame, frames per step per GPU)
WORKLOADS = {
    "5m": ("synthetic Poisson stream, 5M events/frame @ 60 fps, cam 640x480, proj 720x1280, ESL_calib_hhi, projector-view depth", 640, 480, 720, 1280, 5_000_000, 64),
    "100k": ("synthetic 100k-event frames, cam 640x480, proj 720x1280, ESL_calib_hhi, projector-view depth", 640, 480, 720, 1280, 100_000, 64),
    "hd20m": ("synthetic Poisson stream, 20M events/frame, cam 1280x720, proj 1080x1920, ESL_calib_hhi (camera K x2), nearest X-map lookup, projector-view depth", 1280, 720, 1080, 1920, 20_000_000, 16),
    "plane": ("fronto-parallel plane (~100% inliers), one burst of events per lit camera pixel, cam 640x480, proj 720x1280, ESL_calib_hhi, projector-view depth", 640, 480, 720, 1280, 0, 64),
    "sweep": ("event-rate sweep 1M-50M events/frame, cam 640x480, proj 720x1280, ESL_calib_hhi, projector-view depth", 640, 480, 720, 1280, 5_000_000, 64),
}
SWEEP_SIZES = [1_000_000, 2_000_000, 5_000_000, 10_000_000, 20_000_000, 50_000_000]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_config(wl_name, args):
    """`config` of the JSON line: identical keys and values in both arms (ours / reference).  `events_per_frame` is
    the nominal size (plane: burst length x ~32 k lit pixels; the frames' actual sizes are in `run`)."""
    label, cw, ch, pw, ph, n_default, _ = WORKLOADS[wl_name]
    n = plane_repeat(args) * 32_000 if wl_name == "plane" else (args.events or n_default)
    return {"workload": label, "name": wl_name, "events_per_frame": int(n), "camera": [cw, ch], "projector": [pw, ph], "view": "projector"}


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
# tables of a workload's geometry (host side; oracle form + engine form)
# ------------------------------------------------------------------------------------------------
def load_geometry(wl_name, need_time_map=False, device=None):
    """(OracleTables, time_map_rect or None).  Default geometry: the golden fixture (tables of the real reference);
    HD: the product's host table builder (same OpenCV calls as the reference, hashes pinned in tests) and, for the
    X-map, the GPU builder when a device is given (bit-exact against the reference's hash in the GPU tests) or the
    oracle's NumPy builder on the CPU arm."""
    from oracle import xmaps_oracle as orc
    from xm_helpers import host_time_map_rect, load_golden_tables

    _, cw, ch, pw, ph, _, _ = WORKLOADS[wl_name]
    if (cw, ch, pw, ph) == (640, 480, 720, 1280):
        tables = load_golden_tables("default")[0]
        return tables, (host_time_map_rect() if need_time_map else None)
    from xmaps_b200.calibration import CamProjCalibrationParams, CamProjMaps
    from xmaps_b200.time_map import ProjectorTimeMap

    p = CamProjCalibrationParams.from_yaml(os.path.join(ROOT, "data", "esl_calib_hhi.json"), cw, ch, pw, ph)
    k = p.camera_K.copy()
    k[:2, :] *= 2.0
    k[1, 2] += -120.0
    p.camera_K = k
    maps = CamProjMaps(p)
    tm = ProjectorTimeMap.from_calib(p, maps).projector_time_map_rectified
    x_offset, xmap_w = 4242, pw
    if device is not None:
        from xmaps_b200.engine import build_x_map

        x_map = build_x_map(tm, xmap_w, xmap_w - 1, x_offset, pw, device=device)[0].cpu().numpy()
    else:
        x_map = _cpu_xmap_cached(wl_name, tm, xmap_w, x_offset, pw)
    tables = orc.OracleTables(
        lut_x=maps.disp_cam_mapx_i16, lut_y=maps.disp_cam_mapy_i16, x_map=x_map, remap_xy=maps.disp_proj_mapxy_i16,
        rect_w=p.rect_image_width, rect_h=p.rect_image_height, t_px_scale=xmap_w - 1, x_offset=x_offset, depth_scale=float(maps.P2[0, 3]),
    )
    return tables, tm


def _cpu_xmap_cached(wl_name, tm, xmap_w, x_offset, num_scanlines):
    """X-map of a non-default geometry on a box without a usable GPU context (the reference arm): built ONCE in a
    child process -- by the reference's own Numba builder (oracle/_ref/x_map.py, all threads) or the oracle's NumPy
    one -- and cached under /tmp, so that the forked workers neither repeat it nor inherit Numba's thread pool."""
    import hashlib
    import subprocess
    import tempfile

    key = hashlib.sha256(np.ascontiguousarray(tm).tobytes()).hexdigest()[:16]
    path = os.path.join(tempfile.gettempdir(), f"xmaps_b200_xmap_{wl_name}_{key}.npy")
    if not os.path.exists(path):
        tm_path = path + ".tm.npy"
        np.save(tm_path, tm)
        code = (
            "import sys, numpy as np; sys.path.insert(0, %r)\n"
            "from oracle import ref_chain, xmaps_oracle as orc\n"
            "tm = np.load(%r)\n"
            "if ref_chain.available():\n"
            "    xm = ref_chain.load().x_map.compute_x_map_from_time_map(time_map=tm, x_map_width=%d, t_px_scale=%d, X_OFFSET=%d, num_scanlines=%d)[0]\n"
            "else:\n"
            "    xm = orc.build_x_map(tm, %d, %d, %d, %d)[0]\n"
            "np.save(%r + '.tmp.npy', xm)\n"
        ) % (ROOT, tm_path, xmap_w, xmap_w - 1, x_offset, num_scanlines, xmap_w, xmap_w - 1, x_offset, num_scanlines, path)
        subprocess.run([sys.executable, "-c", code], check=True)
        os.replace(path + ".tmp.npy", path)
        os.remove(tm_path)
    return np.load(path)


def algorithmic_bytes(tables, n):
    """B(N_in) = 16 N_in + B_lut + B_xmap + B_remap + B_out (projector view)."""
    lut = tables.cam_w * tables.cam_h * 4
    xmap = tables.x_map.size * 2
    out = tables.proj_w * tables.proj_h * 4
    return 16 * int(n) + lut + xmap + 2 * out


# ------------------------------------------------------------------------------------------------
# synthetic frames
# ------------------------------------------------------------------------------------------------
def synth_frame_cuda(seed, n, device, cam_w, cam_h):
    """Uniform-pixel, time-sorted frame of 16-byte EventCD records generated on the GPU (homogeneous Poisson
    process conditioned on N, 90 % positive polarity).  Philox: the same seed gives the same frame on any device."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.randint(0, cam_w, (n,), generator=g, device=device, dtype=torch.int32)
    y = torch.randint(0, cam_h, (n,), generator=g, device=device, dtype=torch.int32)
    p = (torch.rand(n, generator=g, device=device) < 0.9).to(torch.int32)
    t = torch.sort(torch.randint(0, FRAME_US, (n,), generator=g, device=device, dtype=torch.int64)).values
    raw = torch.empty((n, 4), dtype=torch.int32, device=device)
    raw[:, 0] = x | (y << 16)
    raw[:, 1] = p
    raw.view(torch.int64)[:, 1] = t + seed * FRAME_US
    return raw


def plane_frame_host(tables, time_map, index, repeat):
    """Frame `index` of the plane workload: depth cycles through 0.30 ... 0.85 m, seeded burst jitter."""
    from oracle import xmaps_oracle as orc

    z = 0.30 + 0.05 * (index % 12)
    return orc.synth_plane_events(tables, time_map, z, frame_us=FRAME_US, repeat=repeat, jitter_us=8 if repeat > 1 else 0, seed=100 + index,
                                  t0=index * FRAME_US)


def host_frame(raw):
    from xmaps_b200.events import EVENT_DTYPE

    return raw.cpu().numpy().view(EVENT_DTYPE).reshape(-1)


# ------------------------------------------------------------------------------------------------
# CPU side: the reference's own modules (oracle/_ref) or, where absent, the NumPy restatement
# ------------------------------------------------------------------------------------------------
class CpuPath:
    def __init__(self, wl_name, tables):
        from oracle import ref_chain
        from oracle import xmaps_oracle as orc

        self.orc, self.tables = orc, tables
        self.ref = None
        if ref_chain.available():
            _, cw, ch, pw, ph, _, _ = WORKLOADS[wl_name]
            hd = (cw, ch) != (640, 480)
            self.ref = ref_chain.RefPath(cw, ch, pw, ph, x_map=tables.x_map, camera_K_scale=2.0 if hd else None, cy_shift=-120.0 if hd else 0.0)
            if not self.ref.tables_match(tables):
                raise RuntimeError("oracle/_ref built different tables than the fixture")
        self.kind = "reference" if self.ref is not None else "port"
        self.what = ("the reference's own modules (oracle/_ref: rectify_cam_coords_i16 -> compute_event_disparity -> compute_disp_map_projector_view -> "
                     "remap_rectified_disp_map_to_proj -> disparity_to_depth_rectified)") if self.ref is not None else \
            "NumPy/OpenCV restatement of the reference (oracle/xmaps_oracle.py; oracle/_ref absent)"

    def frame_depth(self, ev):
        if self.ref is not None:
            return self.ref.frame_depth(ev, 0)
        return self.orc.frame_depth(self.tables, ev, 0)


_W = {}


def _worker_init(wl_name, n_events, repeat):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["NUMBA_NUM_THREADS"] = "1"
    try:
        import cv2

        cv2.setNumThreads(1)
    except Exception:
        pass
    tables, tm = load_geometry(wl_name, need_time_map=(wl_name == "plane"))
    _W.update(cpu=CpuPath(wl_name, tables), tables=tables, tm=tm, n=n_events, wl=wl_name, repeat=repeat, ev=None)


def _worker_step(seed):
    from oracle import xmaps_oracle as orc

    if _W["ev"] is None:
        t = _W["tables"]
        if _W["wl"] == "plane":
            _W["ev"] = plane_frame_host(t, _W["tm"], seed, _W["repeat"])
        else:
            _W["ev"] = orc.synth_events(1000 + seed, _W["n"], t.cam_w, t.cam_h, frame_us=FRAME_US)
    t0 = time.perf_counter()
    depth = _W["cpu"].frame_depth(_W["ev"])
    return time.perf_counter() - t0, len(_W["ev"]), float(depth[::97, ::89].sum())


def plane_repeat(args):
    """Events per lit pixel of the plane workload (~32 k lit pixels at the default geometry)."""
    return max(1, int(round(args.events / 32_000))) if args.events else 16


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    from oracle import ref_chain

    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    workers = max(1, min(cores, args.ref_workers if args.ref_workers > 0 else 64))
    wl = "5m" if args.workload == "sweep" else args.workload
    n = args.events or WORKLOADS[wl][5]
    repeat = plane_repeat(args)
    kind = "reference" if ref_chain.available() else "port"
    ctx = mp.get_context("fork")
    with ctx.Pool(workers, initializer=_worker_init, initargs=(wl, n, repeat)) as pool:
        for _ in range(max(1, args.warmup)):
            res = pool.map(_worker_step, range(workers))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = pool.map(_worker_step, range(workers))
        dt = time.perf_counter() - t0
    events_per_step = sum(r[1] for r in res)
    per_frame = float(np.median([r[0] for r in res]))
    value = events_per_step * args.steps / dt
    if wl == "plane":
        n = events_per_step // workers
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
        "dtype": "int16 (+ f64 time normalisation, f64->f32 depth)",
        "data": "synthetic",
        "config": make_config(args.workload, args),
        "cpu_baseline": {
            "value": value,
            "unit": "events/s",
            "cores": workers,
            "kind": kind,
            "sample": f"{workers} worker processes x 1 frame of ~{n} events per step, one thread each (the per-event stages are single-threaded NumPy); "
                      + ("the reference's own modules from oracle/_ref" if kind == "reference" else "NumPy restatement, oracle/_ref absent"),
            "per_core_events_per_s": n / per_frame if per_frame > 0 else None,
            "ms_per_frame_per_worker": per_frame * 1e3,
        },
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "run": {"frames_per_step": workers, "host_cores": cores, "events_per_frame_actual": int(n)},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(cpu, ev_host, runs=5):
    try:
        import cv2

        cv2.setNumThreads(1)
    except Exception:
        pass
    times = []
    depth = cpu.frame_depth(ev_host)  # warm (Numba JIT of the reference's depth loop)
    for _ in range(runs):
        t0 = time.perf_counter()
        depth = cpu.frame_depth(ev_host)
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return depth, {
        "value": len(ev_host) / med,
        "unit": "events/s",
        "cores": 1,
        "kind": cpu.kind,
        "sample": f"1 frame of {len(ev_host)} events, median of {runs} runs, single thread; {cpu.what}",
        "ms_per_frame": med * 1e3,
    }


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import xmaps_b200  # noqa: F401
    from oracle import xmaps_oracle as orc
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

    wl_name = args.workload
    geo = "5m" if wl_name == "sweep" else wl_name
    label, cam_w, cam_h, proj_w, proj_h, n_default, f_default = WORKLOADS[geo]
    tables, time_map = load_geometry(geo, need_time_map=(geo == "plane"), device=device)
    eng = DepthEngine(
        TableSet(
            lut_x=tables.lut_x, lut_y=tables.lut_y, x_map=tables.x_map, remap_xy=tables.remap_xy,
            rect_w=tables.rect_w, rect_h=tables.rect_h, t_px_scale=tables.t_px_scale, x_offset=tables.x_offset,
            depth_scale=tables.depth_scale,
        ),
        device=device,
    )
    gather = world > 1 and not args.no_gather
    if gather and args.gather_mode == "nccl":
        # the persistent batch kernel would otherwise own every SM until it ends and NCCL's copy kernels
        # (the gather of the previous chunk) could not run beside it
        eng.set_option("reserve_sms", args.reserve_sms)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def render(fr, dst):
        eng.frame_batch(fr, view=VIEW_PROJECTOR, output=OUT_DEPTH, out=dst)

    def render_ptrs(fr, ptrs):
        eng.frame_batch(fr, view=VIEW_PROJECTOR, output=OUT_DEPTH, out_ptrs=ptrs)

    peak, peak_src = peaks()
    cpu = CpuPath(geo, tables) if rank == 0 else None
    sizes = SWEEP_SIZES if wl_name == "sweep" else [args.events or n_default]
    results = []
    headline = None

    for n_req in sizes:
        # ---- frames of this size, resident in HBM --------------------------------------------------
        if args.frames:
            F = args.frames
        elif wl_name == "sweep":
            F = int(max(4, min(64, 6.4e9 // (16 * n_req))))
        else:
            F = f_default
        # global frame index of local frame j on this rank = j * world + rank (round-robin sharding)
        if geo == "plane":
            rep = plane_repeat(args)
            host_frames_np = [plane_frame_host(tables, time_map, j * world + rank, rep) for j in range(F)]
            frames = [eng.events(h).raw for h in host_frames_np]
            n_events_local = sum(len(h) for h in host_frames_np)
            n = n_events_local // F
        else:
            n = n_req
            frames = [synth_frame_cuda(j * world + rank, n, device, cam_w, cam_h) for j in range(F)]
            n_events_local = n * F
        out = torch.empty((F, proj_h, proj_w), dtype=torch.float32, device=device)

        # frames per render call = per batch-kernel launch (<= 32): with the NCCL gather 16, so that the gather of
        # one chunk overlaps the render of the next; the peer modes need no such compromise
        if args.gather_chunk > 0:
            chunk = args.gather_chunk
        elif gather and args.gather_mode == "nccl":
            chunk = 16
        else:
            chunk = 32
        sharder = FrameSharder(render, rank, world, dst=0, chunk=[chunk])
        gather_mode, gathered = "none", None
        if gather:
            gather_mode = args.gather_mode
            if gather_mode in ("direct", "copy"):
                slabs = sharder.enable_peer(out, mode=gather_mode, render_ptrs=render_ptrs)
                if slabs is None:
                    gather_mode = "nccl (peer mapping failed: %s)" % sharder.peer_error
                elif rank == 0:
                    gathered = slabs
            if gather_mode.startswith("nccl") and rank == 0:
                gathered = [out] + [torch.empty_like(out) for _ in range(world - 1)]

        def step():
            sharder.run(frames, out, gathered, gather=gather)

        for _ in range(max(3, args.warmup)):
            step()
        barrier()

        # ---- timed region A: the headline number ----------------------------------------------------
        steps = args.steps if wl_name != "sweep" else max(3, min(args.steps, 5))
        sampler = ClockSampler(local_rank)
        launches0 = N.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.start()
        barrier()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        clocks = sampler.stop()
        launches = N.launch_count() - launches0
        ms = e0.elapsed_time(e1)
        tot = torch.tensor([ms, float(n_events_local)], dtype=torch.float64, device=device)
        if world > 1:
            mx = tot.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            ms, events_all = float(mx[0].item()), float(tot[1].item())
        else:
            events_all = float(n_events_local)
        value = events_all * steps / (ms * 1e-3)

        # ---- parity of what was just timed: frames of the batch output (every rank's own, on rank 0 also one
        #      gathered frame per other rank) against the oracle ------------------------------------------
        parity = None
        if not args.quick or args.check:
            check_idx = sorted({0, min(31, F - 1), F - 1})

            def frame_host_np(j, r):
                if geo == "plane":
                    return plane_frame_host(tables, time_map, j * world + r, plane_repeat(args))
                return host_frame(synth_frame_cuda(j * world + r, n, device, cam_w, cam_h))

            bad, checked = 0, []
            if rank == 0:
                for j in check_idx:
                    want = orc.frame_depth(tables, frame_host_np(j, 0), orc.VIEW_PROJECTOR)
                    bad += int(np.count_nonzero(out[j].cpu().numpy() != want))
                    checked.append("rank0.out[%d]" % j)
                if gathered is not None:
                    for r in range(1, world):
                        j = (7 * r) % F
                        want = orc.frame_depth(tables, frame_host_np(j, r), orc.VIEW_PROJECTOR)
                        bad += int(np.count_nonzero(gathered[r][j].cpu().numpy() != want))
                        checked.append("gathered[%d][%d]" % (r, j))
                parity = {"mismatching_pixels": bad, "frames_checked": len(checked), "checked": checked,
                          "what": "depth frames written by the TIMED batch_kernel launches (and the gathered slabs on rank 0) vs the oracle, bit-exact"}

        # ---- timed region B: same steps with CUDA events around every batch-kernel launch --------------
        eng.set_option("profile", -1)
        eng.set_option("profile", 1)
        torch.cuda.synchronize(device)
        for _ in range(steps):
            for lo, hi in sharder._spans(F):
                render(frames[lo:hi], out[lo:hi])
        torch.cuda.synchronize(device)
        k1_ns, pf, pl = eng.get_option("profile_k1_ns"), eng.get_option("profile_frames"), eng.get_option("profile_launches")
        eng.set_option("profile", 0)
        sharder.close_peer()
        frames_per_launch = max(1, round(pf / max(1, pl)))
        us_per_launch = k1_ns / max(1, pl) / 1e3
        bytes_per_frame = algorithmic_bytes(tables, n)
        bytes_per_launch = bytes_per_frame * frames_per_launch
        achieved = bytes_per_launch / (us_per_launch * 1e-6) / 1e9 if us_per_launch > 0 else 0.0
        roofline = {
            "bound": "hbm",
            "kernel": "xm::batch_kernel (persistent: event warps + epilogue warp groups, %d frames per launch)" % frames_per_launch,
            "achieved": achieved,
            "peak": peak,
            "peak_source": peak_src,
            "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": None,
            "bytes_per_launch": bytes_per_launch,
            "us_per_launch": us_per_launch,
            "how": "CUDA events recorded on the launch stream around every batch_kernel launch in a second pass over the same steps",
            "frames_per_launch": frames_per_launch,
            "algorithmic_bytes_per_frame": bytes_per_frame,
            "frame_us": ms * 1e3 / (F * steps),
        }
        tf = os.path.join(ROOT, "profiles", "frame_dram_bytes.json")
        if wl_name == "5m" and os.path.exists(tf):
            try:
                with open(tf) as fh:
                    tj = json.load(fh)
                per_frame = tj.get("dram_bytes_per_frame")
                roofline["traffic"] = per_frame * frames_per_launch if per_frame else None
                roofline["traffic_source"] = "static file profiles/frame_dram_bytes.json <- " + str(tj.get("source"))
            except Exception:
                pass
        res = {"events_per_frame": n, "frames_per_step_per_gpu": F, "value": value, "frame_us": roofline["frame_us"], "ms_per_step": ms / steps,
               "steps": steps, "roofline_frac": roofline["frac"], "us_per_launch": us_per_launch, "frames_per_launch": frames_per_launch,
               "gpu_launches": int(launches), "gather": gather_mode}
        if parity is not None:
            res["mismatching_pixels"] = parity["mismatching_pixels"]
        results.append(res)
        if headline is None or n_req == 5_000_000:
            headline = dict(n=n, F=F, value=value, ms=ms, steps=steps, roofline=roofline, parity=parity, clocks=clocks, launches=launches,
                            gather_mode=gather_mode, frames=frames if len(sizes) == 1 else None, out=out if len(sizes) == 1 else None)
        if len(sizes) > 1:
            del frames, out, gathered
            torch.cuda.empty_cache()

    if args.quick:  # parameter sweeps: kernel numbers only
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        if rank == 0:
            for r in results:
                print(json.dumps({"quick": True, "options": args.opt, **r}), flush=True)
        return

    # ---- e2e on EVERY rank: pinned host events -> pinned host depth frames, own PCIe link each ---------
    h = headline
    n, F = h["n"], h["F"]
    if h["frames"] is None:  # sweep: rebuild the 5 M frames for the e2e / CPU legs
        n, F = 5_000_000, 16
        h_frames = [synth_frame_cuda(j * world + rank, n, device, cam_w, cam_h) for j in range(F)]
    else:
        h_frames = h["frames"]
    e2e_frames = min(F, args.e2e_frames)
    host_frames = [h_frames[j].cpu().pin_memory() for j in range(e2e_frames)]
    max_ev = max(f.shape[0] for f in host_frames)
    hs = HostFrameStream(eng, max_ev, view=VIEW_PROJECTOR, output=OUT_DEPTH, depth=3)
    host_out = hs.alloc_outputs(e2e_frames)
    hs.run(host_frames, host_out)  # warm-up
    reps = max(1, args.e2e_reps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        hs.run(host_frames, host_out)
    torch.cuda.synchronize(device)
    e2e_dt = time.perf_counter() - t0
    ev_local = float(sum(f.shape[0] for f in host_frames))
    agg = torch.tensor([e2e_dt, ev_local], dtype=torch.float64, device=device)
    if world > 1:
        mx = agg.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        e2e_dt_max, ev_all = float(mx[0].item()), float(agg[1].item())
    else:
        e2e_dt_max, ev_all = e2e_dt, ev_local

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- CPU baseline (one frame of the same workload, one core) + e2e parity -----------------------
    ev_host = host_frame(h_frames[0])
    want, cpu_line = cpu_baseline_sample(cpu, ev_host, runs=args.cpu_runs)
    e2e_ok = bool(np.array_equal(host_out[0].numpy(), want))
    if h["parity"] is not None:
        h["parity"]["reference_modules_agree"] = bool(np.array_equal(want, orc.frame_depth(tables, ev_host, orc.VIEW_PROJECTOR))) if cpu.kind == "reference" else None
    h2d = int(sum(f.shape[0] for f in host_frames) * 16)
    d2h = int(e2e_frames * proj_h * proj_w * 4)
    e2e = {
        "value": ev_all * reps / e2e_dt_max,
        "unit": "events/s",
        "h2d_bytes_per_step": h2d * world,
        "d2h_bytes_per_step": d2h * world,
        "h2d_bytes_per_step_per_rank": h2d,
        "d2h_bytes_per_step_per_rank": d2h,
        "frames_per_step_per_rank": e2e_frames,
        "ranks": world,
        "api": "xmaps_b200.host_stream.HostFrameStream.run on every rank (pinned host EventCD buffers -> pinned host depth frames; own PCIe link each), "
               "aggregate events / slowest rank's wall time between barriers",
        "matches_oracle": e2e_ok,
    }

    line = {
        "metric": METRIC,
        "value": h["value"],
        "unit": "events/s",
        "n_gpus": world,
        "steps": h["steps"],
        "warmup": max(3, args.warmup),
        "ms_per_step": h["ms"] / h["steps"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int16 (+ f64 time normalisation, f64->f32 depth)",
        "data": "synthetic",
        "config": make_config(wl_name, args),
        "run": {
            "frames_per_step_per_gpu": h["F"],
            "events_per_frame_actual": int(h["n"]),
            "l2": "inputs larger than L2 (%.1f GB of events per step per GPU)" % (h["F"] * h["n"] * 16 / 1e9),
            "parallelism": "frames round-robin over %d GPU(s)%s" % (world, "" if not gather else ", gather of depth frames to rank 0 inside the step (%s)" % h["gather_mode"]),
            "time_bounds": "sorted + device-side verification and fix-up",
            "options": args.opt,
        },
        "frames_per_sec": h["value"] / h["n"],
        "roofline": h["roofline"],
        "cpu_baseline": cpu_line,
        "e2e": e2e,
        "gpu_launches": int(h["launches"]),
        "clocks": h["clocks"],
        "parity": h["parity"],
    }
    if wl_name == "sweep":
        line["sweep"] = results
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
    ap.add_argument("--workload", default="5m", choices=sorted(WORKLOADS))
    ap.add_argument("--events", type=int, default=0, help="events per frame (0 = the workload's own; plane: sets the burst length)")
    ap.add_argument("--frames", type=int, default=0, help="distinct frames per step per GPU (0 = the workload's own)")
    ap.add_argument("--gather-chunk", type=int, default=0, help="frames per render call (one persistent kernel each); 0 = 32, or 16 with the NCCL gather")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--reserve-sms", type=int, default=0, help="SMs left to NCCL while the batch kernel runs (N > 1, --gather-mode nccl)")
    ap.add_argument("--gather-mode", default="direct", choices=["direct", "copy", "nccl"],
                    help="direct: the render kernel writes finished frames into rank 0's memory (peer mapping); copy: copy-engine peer copies; nccl: dist.gather")
    ap.add_argument("--cpu-runs", type=int, default=5)
    ap.add_argument("--e2e-frames", type=int, default=16)
    ap.add_argument("--e2e-reps", type=int, default=3)
    ap.add_argument("--ref-workers", type=int, default=0, help="worker processes of the reference arm (0 = one per host core, at most 64)")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (repeatable)")
    ap.add_argument("--quick", action="store_true", help="print kernel timings only (no CPU baseline / e2e)")
    ap.add_argument("--check", action="store_true", help="with --quick: still compare the timed output with the oracle")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
