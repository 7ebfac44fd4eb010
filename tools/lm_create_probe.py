import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, bench
from ezcalib_b200 import synth, lm
ctx = lm.Context(0)
fixed = dict(function_tolerance=-1.0, parameter_tolerance=-1.0, gradient_tolerance=-1.0)
model='radtan5'
P = bench.make_problem(model, bench.N_VIEWS_10M, seed_offset=4)
mid = synth.MODEL_ID[model]
obs_pin = torch.from_numpy(P.obs).pin_memory(); poses_pin = torch.from_numpy(P.poses_init).pin_memory()
dst = torch.empty_like(obs_pin, device='cuda')
for _ in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter(); dst.copy_(obs_pin, non_blocking=True); torch.cuda.synchronize(); print('torch H2D 80MB %.3f ms'%(1e3*(time.perf_counter()-t0)))
for rep in range(4):
    torch.cuda.synchronize(); t0=time.perf_counter()
    s2 = lm.LmSolver(ctx, mid, True, obs_pin.numpy(), P.target, P.width, P.height)
    torch.cuda.synchronize(); t1=time.perf_counter()
    ta=time.perf_counter(); s2.set_params(P.focal_init, P.pp_init, P.dist_init, poses_pin.numpy()); tb=time.perf_counter(); print('set_params wall %.3f ms'%(1e3*(tb-ta)), poses_pin.dtype, poses_pin.is_pinned())
    summ2,_ = s2.solve(max_num_iterations=10, **fixed)
    tc=time.perf_counter(); out = s2.get_params(); td=time.perf_counter(); print('get_params wall %.3f ms'%(1e3*(td-tc))); s2.close()
    torch.cuda.synchronize(); print('rep',rep,'create %.3f ms total %.3f ms'%(1e3*(t1-t0),1e3*(time.perf_counter()-t0)), flush=True)
