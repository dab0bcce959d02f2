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


def _wrap_device_memory(ptr: int, shape, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    """A torch view of device memory this process owns but torch did not allocate (xm_peer_alloc)."""
    typestr = {torch.float32: "<f4", torch.uint8: "|u1", torch.int32: "<i4"}[dtype]

    class _Mem:
        pass

    m = _Mem()
    m.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 2, "strides": None}
    return torch.as_tensor(m, device=device)


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
        # peer modes (enable_peer): this rank's slab of the destination's gather buffer, mapped through CUDA IPC
        self._peer_mode = None
        self._remote_ptr, self._remote_dev = None, -1
        self._mapped, self._owned = None, []
        self._render_ptrs, self._frame_bytes = None, 0
        self._done_flag = None
        self.peer_error = None

    # ------------------------------------------------------------------ gather without a collective kernel
    def enable_peer(self, out_local: torch.Tensor, mode: str = "direct", render_ptrs: Optional[Callable] = None):
        """Gather through peer memory instead of ``dist.gather`` (all ranks on one node, NVLink / NVSwitch).

        Rank ``dst`` allocates one slab per other rank with ``xm_peer_alloc`` (C ABI: cudaMalloc + CUDA IPC handle)
        and broadcasts the handles; rank r maps its slab with ``xm_peer_open`` (peer access enabled explicitly and
        checked).  Then, per chunk of frames,

        * ``mode="direct"``: ``render_ptrs(frames, ptrs)`` renders with the MAPPED addresses as output, i.e. the
          epilogue warps of the persistent kernel store finished tiles straight into rank ``dst``'s HBM -- compute
          and gather fused, no extra kernel, no SM taken from the render kernel;
        * ``mode="copy"``: frames are rendered locally and moved with ``xm_peer_copy`` (cudaMemcpyPeerAsync, copy
          engines) on a side stream.

        NCCL's copy kernels, by contrast, cannot be scheduled while the persistent kernel owns the SMs.  Returns the
        list of per-rank slabs on rank ``dst`` (``[out_local, slab_1, ...]``), ``[]`` on other ranks, or ``None`` if
        any rank failed to map (every rank then stays on ``dist.gather``)."""
        import ctypes as C

        from . import _native as N

        if mode not in ("direct", "copy"):
            raise ValueError("mode must be 'direct' or 'copy'")
        if mode == "direct" and render_ptrs is None:
            raise ValueError("mode='direct' needs render_ptrs(frames, ptrs)")
        dev = out_local.device.index
        frame_bytes = out_local[0].numel() * out_local.element_size()
        n_local = out_local.shape[0]
        ok = 1
        slabs, handles = [], [None] * self.world
        try:
            if self.rank == self.dst:
                for r in range(self.world):
                    if r == self.dst:
                        slabs.append(out_local)
                        continue
                    ptr, h = C.c_void_p(), N.XmIpcHandle()
                    N.check(N.lib.xm_peer_alloc(dev, frame_bytes * n_local, C.byref(ptr), C.byref(h)))
                    self._owned.append((dev, ptr.value))
                    slabs.append(_wrap_device_memory(ptr.value, tuple(out_local.shape), out_local.dtype, out_local.device))
                    handles[r] = bytes(h.bytes)
        except Exception as exc:  # noqa: BLE001 - any failure means: every rank falls back to NCCL
            self.peer_error = repr(exc)
            ok = 0
        payload = [handles, dev]
        dist.broadcast_object_list(payload, src=self.dst, group=self.group)
        handles, dst_dev = payload
        if self.rank != self.dst and ok:
            try:
                h = N.XmIpcHandle()
                C.memmove(h.bytes, handles[self.rank], 64)
                ptr = C.c_void_p()
                N.check(N.lib.xm_peer_open(dev, int(dst_dev), C.byref(h), C.byref(ptr)))
                self._mapped = (dev, ptr.value)
                self._remote_ptr, self._remote_dev = ptr.value, int(dst_dev)
            except Exception as exc:  # noqa: BLE001
                self.peer_error = repr(exc)
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=out_local.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            self.close_peer()
            return None
        self._peer_mode, self._render_ptrs, self._frame_bytes = mode, render_ptrs, frame_bytes
        self._done_flag = torch.zeros(1, dtype=torch.int32, device=out_local.device)
        return slabs if self.rank == self.dst else []

    def close_peer(self):
        from . import _native as N

        if self._mapped is not None:
            N.lib.xm_peer_close(self._mapped[0], self._mapped[1])
            self._mapped = None
        for dev, ptr in self._owned:
            N.lib.xm_peer_free(dev, ptr)
        self._owned = []
        self._peer_mode = None
        self._remote_ptr = None

    def _spans(self, n):
        lo, k = 0, 0
        while lo < n:
            size = self.chunks[min(k, len(self.chunks) - 1)]
            hi = min(n, lo + size)
            yield lo, hi
            lo, k = hi, k + 1

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
        peer = self._peer_mode if (gather and self.world > 1) else None
        if on_cuda and gather and self.world > 1 and self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(out_local.device)
        works = []
        for lo, hi in self._spans(n):
            if peer == "direct" and self.rank != self.dst:
                # the render kernel's epilogue writes straight into the destination rank's slab (NVLink stores)
                self._render_ptrs(local_frames[lo:hi], [self._remote_ptr + j * self._frame_bytes for j in range(lo, hi)])
                continue
            self.render(local_frames[lo:hi], out_local[lo:hi])
            if not gather or peer == "direct":
                continue
            recv = [g[lo:hi] for g in gathered] if (gathered is not None and self.rank == self.dst) else None
            if on_cuda and self.world > 1:
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(out_local.device))
                with torch.cuda.stream(self._comm_stream):
                    self._comm_stream.wait_event(done)
                    if peer == "copy":
                        if self.rank != self.dst:  # (the destination's own frames are already in place)
                            from . import _native as N

                            N.check(N.lib.xm_peer_copy(
                                self._remote_ptr + lo * self._frame_bytes, self._remote_dev, out_local[lo].data_ptr(), out_local.device.index,
                                (hi - lo) * self._frame_bytes, self._comm_stream.cuda_stream))
                    else:
                        works.append(self._gather_chunk(out_local[lo:hi], recv, async_op=True))
            else:
                self._gather_chunk(out_local[lo:hi], recv, async_op=False)
        if on_cuda and self._comm_stream is not None:
            for w in works:
                if w is not None:
                    w.wait()  # orders the NCCL stream before the comm stream
            torch.cuda.current_stream(out_local.device).wait_stream(self._comm_stream)
        if peer is not None:
            # stream-ordered arrival: once this tiny all-reduce has completed on rank dst's stream, every peer's
            # render kernels / peer copies of this call have completed, i.e. their frames are in dst's slabs
            dist.all_reduce(self._done_flag, group=self.group)
        return gathered if self.rank == self.dst else None
