"""Multi-GPU gather through peer memory (xm_peer_alloc / xm_peer_open + XmFrameArgs.d_out = mapped address):
the CUDA-IPC hand-over is exercised across two PROCESSES -- on two GPUs when the box has them, on one GPU
otherwise (the mapping then needs no peer access, everything else is the same code)."""
import ctypes as C

import numpy as np
import pytest

from oracle import xmaps_oracle as orc
from test_gpu_parity import make_engine
from xm_helpers import load_golden_tables

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _child(handle_bytes, owner_dev, my_dev, frame_bytes, seeds, q):
    try:
        import os
        import sys

        here = os.path.dirname(os.path.abspath(__file__))
        for p in (os.path.dirname(here), here):
            if p not in sys.path:
                sys.path.insert(0, p)
        import torch as th

        import xmaps_b200  # noqa: F401
        from oracle import xmaps_oracle as o
        from test_gpu_parity import make_engine as mk
        from xm_helpers import load_golden_tables as lg
        from xmaps_b200 import _native as N

        th.cuda.set_device(my_dev)
        tables, z = lg("small")
        eng = mk(tables, z) if my_dev == 0 else None
        if eng is None:
            from xmaps_b200.engine import DepthEngine, TableSet

            eng = DepthEngine(TableSet(lut_x=tables.lut_x, lut_y=tables.lut_y, x_map=tables.x_map, remap_xy=tables.remap_xy, rect_w=tables.rect_w,
                                       rect_h=tables.rect_h, t_px_scale=tables.t_px_scale, x_offset=tables.x_offset, depth_scale=tables.depth_scale),
                              device=f"cuda:{my_dev}")
        h = N.XmIpcHandle()
        C.memmove(h.bytes, handle_bytes, 64)
        ptr = C.c_void_p()
        N.check(N.lib.xm_peer_open(my_dev, owner_dev, C.byref(h), C.byref(ptr)))
        frames = [o.synth_events(s, 15_000, 160, 120) for s in seeds]
        eng.frame_batch(frames, view=0, out_ptrs=[ptr.value + i * frame_bytes for i in range(len(frames))])
        th.cuda.synchronize(my_dev)
        N.check(N.lib.xm_peer_close(my_dev, ptr))
        q.put("ok")
    except Exception as exc:  # noqa: BLE001
        q.put("child failed: %r" % (exc,))


def test_frames_rendered_into_another_process_slab():
    import torch.multiprocessing as mp

    from xmaps_b200 import _native as N
    from xmaps_b200.sharding import _wrap_device_memory

    tables, z = load_golden_tables("small")
    h, w = tables.proj_h, tables.proj_w
    frame_bytes, seeds = h * w * 4, [901, 902, 903]
    torch.cuda.init()
    ptr, handle = C.c_void_p(), N.XmIpcHandle()
    N.check(N.lib.xm_peer_alloc(0, frame_bytes * len(seeds), C.byref(ptr), C.byref(handle)))
    try:
        slab = _wrap_device_memory(ptr.value, (len(seeds), h, w), torch.float32, torch.device("cuda", 0))
        slab.zero_()
        torch.cuda.synchronize()
        peer_dev = 1 if torch.cuda.device_count() > 1 else 0
        if peer_dev:
            can, rank = C.c_int32(), C.c_int32()
            N.check(N.lib.xm_peer_info(peer_dev, 0, C.byref(can), C.byref(rank)))
            assert can.value == 1
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        p = ctx.Process(target=_child, args=(bytes(handle.bytes), 0, peer_dev, frame_bytes, seeds, q))
        p.start()
        msg = q.get(timeout=300)
        p.join(timeout=60)
        assert msg == "ok", msg
        got = slab.cpu().numpy()
        for i, s in enumerate(seeds):
            assert np.array_equal(got[i], orc.frame_depth(tables, orc.synth_events(s, 15_000, 160, 120), 0)), f"frame {i}"
    finally:
        N.check(N.lib.xm_peer_free(0, ptr))
