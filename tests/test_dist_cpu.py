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


def _rig_worker(rank, world, port, q):
    """run_rig_extrinsics' host bookkeeping over gloo: cam0's frames broadcast from its owner, every rank builds the blocks of its
    own cameras, the extrinsics of all cameras are gathered on every rank.  The GPU solve is stubbed (returns the initial
    extrinsics and records what it was given)."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ezcalib_b200 import ezcalibrator as E, lm, synth
    cams, tgt = _make_rig()
    seen = {}

    def fake_multicam(ctx, model_id, mono, obs, pts, W, H, focal, pp, dist_, init, n_blocks_global=None, collective=0):
        seen.update(n_blocks=obs.shape[0], n_global=n_blocks_global, collective=collective, obs_sum=float(np.where(obs < 0, 0, obs).sum()),
                    pts_sum=float(pts.sum()))
        return np.array(init, copy=True).reshape(-1, 7), dict(iterations=1, termination_name="stub")

    lm.lm_multicam = fake_multicam
    local = {c: cams[c] for c in range(len(cams)) if c % world == rank}
    out, summ = E.run_rig_extrinsics(None, local, len(cams), tgt, ezdist.TorchComm())
    q.put((rank, sorted(out.keys()), {c: out[c].tolist() for c in out}, seen, summ["n_good_pairs"]))
    dist.destroy_process_group()


def _make_rig():
    from ezcalib_b200 import ezcalibrator as E, synth
    rng = np.random.default_rng(7)
    tgt = synth.chessboard_target(7, 10, 0.05)
    n_cams, n_frames = 5, 9
    poses0 = [np.concatenate([synth.quat_from_rotvec(0.2 * rng.standard_normal(3)), [0.0, 0.0, 0.8] + 0.05 * rng.standard_normal(3)])
              for _ in range(n_frames)]
    cams = []
    for c in range(n_cams):
        ext = np.concatenate([synth.quat_from_rotvec(0.02 * c * np.ones(3)), [0.05 * c, 0.0, 0.0]])
        cam = E.Camera("radtan5", True, 1280, 1024, np.array([900.0 + c]), np.array([640.0, 512.0]), np.zeros(5))
        for i in range(n_frames):
            if c == 2 and i == 4:
                continue   # a frame camera 2 does not have
            T = synth.se3_mul(ext, poses0[i])
            px = synth.project("radtan5", 900.0 + c, 900.0 + c, 640.0, 512.0, np.zeros(5), T, tgt).astype(np.float32)
            fr = E.CalibFrame("%06d.png" % i, px, T, rmse_err=0.1 if not (c == 0 and i == 7) else 2.5)
            cam.frames.append(fr)
        cams.append(cam)
    return cams, tgt


def test_rig_extrinsics_bookkeeping_gloo_world2():
    import torch.multiprocessing as mp
    from ezcalib_b200 import ezcalibrator as E
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_rig_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=180) for _ in ps)
    for p in ps: p.join(timeout=60)
    cams, tgt = _make_rig()
    obs, pts, init, n_good = E.build_multicam_problem(cams, tgt)     # the single-process reference of the same bookkeeping
    for rank, keys, out, seen, n_pairs in res:
        assert keys == [1, 2, 3, 4] and n_pairs == n_good
        assert seen["n_global"] == 4 and seen["collective"] == 1 and seen["n_blocks"] == 2
        mine = [c for c in (1, 2, 3, 4) if c % 2 == rank]
        assert abs(seen["obs_sum"] - float(np.where(obs[[c - 1 for c in mine]] < 0, 0, obs[[c - 1 for c in mine]]).sum())) < 1e-3
        assert abs(seen["pts_sum"] - float(pts.sum())) < 1e-9
        for c in (1, 2, 3, 4):
            assert np.allclose(out[c], init[c - 1], atol=1e-15)
