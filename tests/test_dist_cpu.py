"""CPU (gloo, world_size 2): the host-side multi-GPU plumbing -- shard partition and the all-reduce adapter that
ezc_set_allreduce calls back into."""
import os
import socket

import numpy as np
import pytest

from ezcalib_b200 import dist as ezdist


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 50, 69445):
        for world in (1, 2, 3, 8):
            spans = [ezdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fn = ezdist.make_allreduce("cpu")
    # packed partial reduced systems: rank r holds r+1 everywhere, plus a max slot
    buf = np.full(190, float(rank + 1))
    mx = np.array([10.0 * (rank + 1)])
    fn(buf.ctypes.data, buf.size, 0, 0)
    fn(mx.ctypes.data, 1, 1, 0)
    # a second call on the same address must reuse the cached view and see fresh data
    buf[:] = rank
    fn(buf.ctypes.data, buf.size, 0, 0)
    q.put((rank, float(buf[0]), float(mx[0])))
    dist.destroy_process_group()


def test_allreduce_adapter_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps: p.join(timeout=60)
    assert res == [(0, 1.0, 20.0), (1, 1.0, 20.0)]
