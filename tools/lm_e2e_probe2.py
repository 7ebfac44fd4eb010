"""The bench's per-model sequence (device-resident solve, then host-buffer solve) with a time breakdown of the e2e leg."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from ezcalib_b200 import synth, lm

ctx = lm.Context(0)
fixed = dict(function_tolerance=-1.0, parameter_tolerance=-1.0, gradient_tolerance=-1.0)
def sync(): torch.cuda.synchronize()
for mi, model in enumerate(bench.BENCH_MODELS):
    P = bench.make_problem(model, bench.N_VIEWS_10M, seed_offset=mi)
    mid = synth.MODEL_ID[model]
    obs_pin = torch.from_numpy(P.obs).pin_memory()
    poses_pin = torch.from_numpy(P.poses_init).pin_memory()
    obs_dev = obs_pin.cuda(non_blocking=True)
    pts_dev = torch.from_numpy(P.target).cuda()
    solver = lm.LmSolver(ctx, mid, True, obs_dev.data_ptr(), pts_dev.data_ptr(), P.width, P.height, pts_period=144,
                         device_ptrs=True, n_blocks=bench.N_VIEWS_10M, obs_per_block=144, n_blocks_global=bench.N_VIEWS_10M)
    solver.set_params(P.focal_init, P.pp_init, P.dist_init, P.poses_init)
    solver.solve(max_num_iterations=3, **fixed)
    solver.close()
    del obs_dev
    sync(); t0 = time.perf_counter()
    s2 = lm.LmSolver(ctx, mid, True, obs_pin.numpy(), P.target, P.width, P.height)
    sync(); t1 = time.perf_counter()
    s2.set_params(P.focal_init, P.pp_init, P.dist_init, poses_pin.numpy())
    sync(); t2 = time.perf_counter()
    summ2, _ = s2.solve(max_num_iterations=10, **fixed)
    sync(); t3 = time.perf_counter()
    out = s2.get_params()
    sync(); t4 = time.perf_counter()
    s2.close()
    sync(); t5 = time.perf_counter()
    print(f"{model:14s}: create {1e3*(t1-t0):7.2f} ms | set_params {1e3*(t2-t1):6.2f} | solve(10 it) {1e3*(t3-t2):7.2f} | get_params {1e3*(t4-t3):6.2f} | close {1e3*(t5-t4):6.2f} | total {1e3*(t5-t0):7.2f}", flush=True)
