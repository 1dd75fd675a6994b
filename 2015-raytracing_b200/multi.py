"""Multi-GPU split of an Assignment-10 render: one process per GPU, each renders a contiguous
range of every pixel's ray slots (split by samples per pixel), then ONE sum-reduce of the
per-pixel accumulation image to rank 0: on GPUs ``Comm.reduce`` = ``rt_render_reduce``, an ncclReduce the C library
issues on the context's stream (NVLink / NVSwitch); ``reduce_accum`` is the same step on host tensors through
``torch.distributed`` (gloo) for the CPU tests of the split / merge logic.

Why slots and not passes: a slot's RNG state lives in ``seeds[id]`` and carries over from pass to
pass with a data-dependent number of draws (A10/code.cl:420-434), so only a split
by slot (or by pixel) reproduces the reference stream.  No kernel reads another slot, so there is
no collective on the data path -- just the final image reduce.
"""
from __future__ import annotations

import numpy as np


def slot_range(rank: int, world: int, rays_per_pixel: int):
    """(slot_begin, slot_count) of ``rank``: contiguous k-ranges, the first ``rays_per_pixel % world``
    ranks take one extra slot.  A rank may get zero slots when world > rays_per_pixel."""
    if not (0 <= rank < world) or rays_per_pixel < 1:
        raise ValueError("slot_range: bad rank/world/rays_per_pixel")
    base, extra = divmod(rays_per_pixel, world)
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def local_seeds(global_seeds, pixels: int, rays_per_pixel: int, slot_begin: int, slot_count: int) -> np.ndarray:
    """This rank's part of the global seed array ``[pixel][k]`` as ``[pixel][k_local]``."""
    s = np.asarray(global_seeds, dtype=np.int32).reshape(pixels, rays_per_pixel)
    return np.ascontiguousarray(s[:, slot_begin:slot_begin + slot_count]).reshape(-1)


def merge_seeds(parts, pixels: int, rays_per_pixel: int) -> np.ndarray:
    """Inverse of :func:`local_seeds` over all ranks' ``(slot_begin, slot_count, seeds)``."""
    out = np.zeros((pixels, rays_per_pixel), dtype=np.int32)
    for begin, count, seeds in parts:
        out[:, begin:begin + count] = np.asarray(seeds, dtype=np.int32).reshape(pixels, count)
    return out.reshape(-1)


def reduce_accum(accum, dst: int = 0, group=None):
    """Sum the ranks' per-pixel accumulation images (float4 per pixel) into rank ``dst`` in place.
    ``accum`` is a torch tensor on the rank's device (or on the CPU with gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


class Comm:
    """rt_comm (include/rt2015.h): the NCCL communicator the C library itself drives -- the reduce is issued from C++
    on the context's stream (``rt_render_reduce``), no torch on the data path.  ``exchange(id_or_None) -> id`` ships rank
    0's 128-byte id to every rank; by default ``torch.distributed`` does it when a process group exists (any backend),
    but a file or a socket serves as well (that is all a Node.js host needs)."""

    def __init__(self, ctx, rank: int, world: int, exchange=None):
        import ctypes as C

        from . import lib as L
        try:   # PyTorch's bundled libnccl.so.2 must be the one the process binds: once the system's older one is loaded under
            import torch  # noqa: F401  -- the same soname, a later `import torch` fails on the symbols it lacks
        except ImportError:
            pass
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            ctx.check(L.dll.rt_comm_unique_id(C.cast(ident, C.c_void_p)))
        raw = (exchange or _exchange_via_torch)(bytes(ident) if rank == 0 else None)
        buf = (C.c_ubyte * 128).from_buffer_copy(raw)
        h = C.c_void_p()
        ctx.check(L.dll.rt_comm_create(ctx.h, self.world, self.rank, C.cast(buf, C.c_void_p), C.byref(h)))
        self.h = h

    def reduce(self, renderer, root: int = 0):
        """Sum every rank's per-pixel accumulation image into ``root``'s (asynchronous on the context's stream)."""
        from . import lib as L
        self.ctx.check(L.dll.rt_render_reduce(renderer.h_render, self.h, int(root)))

    def close(self):
        from . import lib as L
        if self.h:
            L.dll.rt_comm_destroy(self.h)
            self.h = None


def _exchange_via_torch(ident):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("multi.Comm: no torch.distributed process group -- pass exchange=")
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if ident is not None:
        t.copy_(torch.frombuffer(bytearray(ident), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def accum_to_pixel(accum, rays_per_pixel: int, passes: int) -> np.ndarray:
    """Host restatement of copyToPixel's tail (A10/code.cl:1381-1384) for the CPU tests; on the GPU
    ``rt_accum_to_pixel`` does this."""
    a = np.asarray(accum, dtype=np.float32).reshape(-1, 4)
    m = np.float32(1.0 / (rays_per_pixel * passes))
    c = a[:, :3] * (np.float32(255.0) * m)
    c = c * np.float32(1.8)
    c = np.minimum(np.maximum(c, np.float32(0.0)), np.float32(255.0))
    out = np.empty((len(a), 4), dtype=np.uint8)
    out[:, :3] = c.astype(np.uint8)
    out[:, 3] = 255
    return out
