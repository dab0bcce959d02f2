"""Round-robin frame sharding + gather, world_size 2 over gloo on CPU.  The renderer is the oracle
here (test infrastructure); on GPUs it is DepthEngine.frame_batch."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xm_helpers import load_golden_tables


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, q, chunk=2):
    from oracle import xmaps_oracle as orc
    from xmaps_b200.sharding import FrameSharder, global_order, local_frame_indices

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tables, _ = load_golden_tables("small")
    mine = local_frame_indices(n_frames, rank, world)
    frames = [orc.synth_events(500 + g, 3000, 160, 120) for g in mine]

    def render(fr, dst):
        for i, f in enumerate(fr):
            dst[i].copy_(torch.from_numpy(orc.frame_depth(tables, f, orc.VIEW_CAMERA)))

    out = torch.zeros((len(mine), 120, 160), dtype=torch.float32)
    gathered = [torch.zeros_like(out) for _ in range(world)] if rank == 0 else None
    FrameSharder(render, rank, world, dst=0, chunk=chunk).run(frames, out, gathered)
    if rank == 0:
        q.put(global_order(gathered, n_frames).numpy())
    dist.barrier()
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("n_frames,chunk", [(6, 2), (14, (3, 2, 1))])  # fixed chunks / tapering schedule (last size repeats)
def test_round_robin_gather_world2(n_frames, chunk):
    from oracle import xmaps_oracle as orc

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q, chunk)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tables, _ = load_golden_tables("small")
    for g in range(n_frames):
        want = orc.frame_depth(tables, orc.synth_events(500 + g, 3000, 160, 120), orc.VIEW_CAMERA)
        assert np.array_equal(got[g], want), f"frame {g} landed in the wrong slot"


def test_index_helpers():
    from xmaps_b200.sharding import global_order, local_frame_indices, owner_of

    assert [owner_of(f, 4) for f in range(6)] == [0, 1, 2, 3, 0, 1]
    assert local_frame_indices(10, 1, 4) == [1, 5, 9]
    stacks = [torch.arange(3).reshape(3, 1) * 4 + r for r in range(4)]
    assert global_order(stacks, 10).flatten().tolist() == list(range(10))
    from xmaps_b200.sharding import FrameSharder

    sh = FrameSharder(lambda fr, dst: None, 0, 1, chunk=(16, 16, 16, 8, 4, 4))
    assert list(sh._spans(64)) == [(0, 16), (16, 32), (32, 48), (48, 56), (56, 60), (60, 64)]
    assert list(sh._spans(70)) == [(0, 16), (16, 32), (32, 48), (48, 56), (56, 60), (60, 64), (64, 68), (68, 70)]
    assert list(FrameSharder(lambda fr, dst: None, 0, 1, chunk=32)._spans(64)) == [(0, 32), (32, 64)]
