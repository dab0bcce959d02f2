"""Multi-GPU: independent projector frames sharded round-robin, one process per GPU.

A frame depends only on its own events and on read-only tables
(/root/reference/python/depth_reprojection_pipe.py:121-167 keeps no cross-frame state), so frame
``f`` goes to rank ``f % world`` and the path itself needs no exchange.  The only collective is the
final gather of finished depth frames to one rank (NCCL over NVLink / NVSwitch), issued chunk by
chunk on a side stream so that it overlaps the kernels of the following chunk.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def owner_of(frame_index: int, world: int) -> int:
    return frame_index % world


def local_frame_indices(n_frames: int, rank: int, world: int) -> List[int]:
    """Global indices of the frames rank ``rank`` renders, in local order."""
    return list(range(rank, n_frames, world))


def global_order(gathered: Sequence[torch.Tensor], n_frames: int) -> torch.Tensor:
    """Interleave per-rank stacks ``gathered[r][j]`` (= global frame ``j * world + r``) back into
    global frame order."""
    world = len(gathered)
    out = torch.empty((n_frames,) + tuple(gathered[0].shape[1:]), dtype=gathered[0].dtype, device=gathered[0].device)
    for r in range(world):
        k = len(range(r, n_frames, world))
        if k:
            out[r::world] = gathered[r][:k]
    return out


class FrameSharder:
    """Renders this rank's frames with ``render(frames, out)`` and gathers all ranks' frames on
    rank ``dst``.

    ``render`` is the only device-specific piece (``DepthEngine.frame_batch`` in production); the
    sharding / gather logic is backend-agnostic, which is how the gloo tests exercise it on CPU.
    """

    def __init__(self, render: Callable, rank: int, world: int, dst: int = 0, group=None, chunk=8):
        """``chunk``: frames per render call, or a sequence of chunk sizes (the last one repeats).  A tapering
        sequence such as ``(16, 16, 16, 8, 4, 4)`` keeps the render calls large (one persistent kernel each)
        while the only gather that cannot overlap a later render -- the last one -- stays small."""
        self.render = render
        self.rank, self.world, self.dst = rank, world, dst
        self.group = group
        sizes = [chunk] if isinstance(chunk, int) else list(chunk)
        self.chunks = [max(1, int(c)) for c in sizes] or [8]
        self.chunk = self.chunks[0]
        self._comm_stream = None
        self._remote = None  # peer-copy mode: this rank's slab of the destination's gather buffer (CUDA IPC mapping)

    def enable_peer_copies(self, gathered: Optional[List[torch.Tensor]]) -> bool:
        """Gather with copy-engine peer copies instead of NCCL kernels: rank ``dst`` exports its per-rank gather
        slabs through CUDA IPC, every other rank maps its own slab and writes finished frames straight into it
        (``cudaMemcpyPeerAsync`` over NVLink).  No SM is involved, so the copies overlap the persistent render
        kernel, which NCCL's copy kernels cannot (they find no free SM resources next to it).  All ranks of one
        node only.  Returns False (and stays on ``dist.gather``) if any rank could not map its slab."""
        from torch.multiprocessing.reductions import reduce_tensor

        ok = 1
        payload = [None]
        try:
            if self.rank == self.dst:
                payload = [[reduce_tensor(g) for g in gathered]]
            dist.broadcast_object_list(payload, src=self.dst, group=self.group)
            if self.rank != self.dst:
                fn, args = payload[0][self.rank]
                self._remote = fn(*args)
        except Exception:  # noqa: BLE001 - any failure means: fall back to NCCL on every rank
            ok = 0
        dev = torch.device("cuda", torch.cuda.current_device())
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            self._remote = None
            return False
        self._peer = True
        return True

    def _spans(self, n):
        lo, k = 0, 0
        while lo < n:
            size = self.chunks[min(k, len(self.chunks) - 1)]
            hi = min(n, lo + size)
            yield lo, hi
            lo, k = hi, k + 1

    _peer = False

    def _gather_chunk(self, send: torch.Tensor, recv: Optional[List[torch.Tensor]], async_op: bool):
        if self.world == 1:
            if recv is not None and recv[0].data_ptr() != send.data_ptr():
                recv[0].copy_(send)
            return None
        return dist.gather(send, recv if self.rank == self.dst else None, dst=self.dst, group=self.group, async_op=async_op)

    def run(self, local_frames: Sequence, out_local: torch.Tensor, gathered: Optional[List[torch.Tensor]] = None, gather: bool = True):
        """``out_local``: ``[n_local, ...]`` on this rank.  ``gathered`` (rank ``dst`` only): one
        ``[n_local_max, ...]`` tensor per rank.  Every rank must hold the same number of local frames
        (pad the stream to a multiple of ``world``).  Returns after all work has been enqueued; the
        caller synchronises."""
        n = len(local_frames)
        on_cuda = out_local.is_cuda
        if on_cuda and gather and self.world > 1 and self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(out_local.device)
        works = []
        for lo, hi in self._spans(n):
            self.render(local_frames[lo:hi], out_local[lo:hi])
            if not gather:
                continue
            recv = [g[lo:hi] for g in gathered] if (gathered is not None and self.rank == self.dst) else None
            if on_cuda and self.world > 1:
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(out_local.device))
                with torch.cuda.stream(self._comm_stream):
                    self._comm_stream.wait_event(done)
                    if self._peer:
                        if self._remote is not None:  # (the destination's own frames are already in place)
                            self._remote[lo:hi].copy_(out_local[lo:hi], non_blocking=True)
                    else:
                        works.append(self._gather_chunk(out_local[lo:hi], recv, async_op=True))
            else:
                self._gather_chunk(out_local[lo:hi], recv, async_op=False)
        if on_cuda and self._comm_stream is not None:
            for w in works:
                if w is not None:
                    w.wait()  # orders the NCCL stream before the comm stream
            torch.cuda.current_stream(out_local.device).wait_stream(self._comm_stream)
        return gathered if self.rank == self.dst else None
