"""Multi-GPU plumbing (one process per GPU, torch.distributed): view/image sharding and the all-reduce callback the
C ABI calls (ezc_set_allreduce) to merge the packed partial reduced system of the LM solve.

The data path has exactly one exchange step per LM iteration pair: a <= 190-double sum of the per-rank
Schur-reduced systems, plus 5 doubles for the candidate evaluation.  Detection shards by image with no collective.
"""
from __future__ import annotations

import ctypes as C


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition [begin, end) of n items (views or images) for `rank`; sizes differ by at most 1.
    Contiguity keeps the sorted file order inside a rank, so the reference's skip rule (ezcalibrator.cpp:50-58) can be
    replayed on the gathered found-flags."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class _CudaArray:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def make_allreduce(device: str = "cuda", group=None):
    """Returns fn(ptr, count, op, stream) that all-reduces `count` float64 values in place at address `ptr`
    (device memory for 'cuda...', host memory for 'cpu' -- the latter is what the gloo tests use).
    op: 0 = sum, 1 = max.  The NCCL call is enqueued relative to torch's current stream, which must be the stream the
    context runs on (Context(stream=torch.cuda.current_stream().cuda_stream))."""
    import torch
    import torch.distributed as dist
    cache = {}

    def fn(ptr: int, count: int, op: int, stream: int):
        key = (ptr, count)
        t = cache.get(key)
        if t is None:
            if device.startswith("cuda"):
                t = torch.as_tensor(_CudaArray(ptr, count), device=device)
            else:
                buf = (C.c_double * count).from_address(ptr)
                t = torch.frombuffer(buf, dtype=torch.float64)
            cache[key] = t
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM, group=group)

    return fn


class TorchComm:
    """The communicator ezcalibrator.run_rig_extrinsics takes, over torch.distributed (host objects: NCCL or gloo group)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def bcast(self, obj, src: int):
        box = [obj]
        self._dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    def allgather(self, obj):
        out = [None] * self.world
        self._dist.all_gather_object(out, obj, group=self.group)
        return out


def init_peer_exchange(ctx, rank: int, world: int, group=None):
    """Install the library-owned peer-memory all-reduce (csrc/comm.cu) on `ctx`: every rank allocates its mailbox, the
    64-byte CUDA IPC handles are all-gathered through torch.distributed (any host transport would do), every rank maps its
    peers' mailboxes.  One process per GPU on one NVLink / NVSwitch node."""
    import torch.distributed as dist
    _, handle = ctx.comm_mailbox(world)
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    ctx.comm_init(rank, world, handles)
    dist.barrier(group=group)
