"""Where does the host-buffer (e2e) LM solve spend its time?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from ezcalib_b200 import synth, lm

model = "radtan5"
ctx = lm.Context(0)
P = bench.make_problem(model, bench.N_VIEWS_10M, seed_offset=4)
mid = synth.MODEL_ID[model]
obs_pin = torch.from_numpy(P.obs).pin_memory()
poses_pin = torch.from_numpy(P.poses_init).pin_memory()
fixed = dict(function_tolerance=-1.0, parameter_tolerance=-1.0, gradient_tolerance=-1.0)
def sync(): torch.cuda.synchronize()
for rep in range(4):
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
    print(f"rep {rep}: create {1e3*(t1-t0):.2f} ms | set_params {1e3*(t2-t1):.2f} | solve(10 it) {1e3*(t3-t2):.2f} | get_params {1e3*(t4-t3):.2f} | close {1e3*(t5-t4):.2f} | total {1e3*(t5-t0):.2f}")
# pageable for comparison
s2 = lm.LmSolver(ctx, mid, True, P.obs, P.target, P.width, P.height); sync(); t0 = time.perf_counter()
s3 = lm.LmSolver(ctx, mid, True, P.obs, P.target, P.width, P.height); sync(); print("create from pageable obs: %.2f ms" % (1e3 * (time.perf_counter() - t0)))
