"""One LM solve (radtan5 mono, 69445 views x 144) for an ncu launch list: warm-up solve, then ONE profiled 5-iteration
solve bracketed by cudaProfilerStart/Stop (run under `ncu --profile-from-start off`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from ezcalib_b200 import synth, lm

model = sys.argv[1] if len(sys.argv) > 1 else "radtan5"
n_views = int(sys.argv[2]) if len(sys.argv) > 2 else bench.N_VIEWS_10M
ctx = lm.Context(0)
P = bench.make_problem(model, n_views, seed_offset=4)
mid = synth.MODEL_ID[model]
solver = lm.LmSolver(ctx, mid, True, P.obs, P.target, P.width, P.height)
fixed = dict(function_tolerance=-1.0, parameter_tolerance=-1.0, gradient_tolerance=-1.0)
solver.set_params(P.focal_init, P.pp_init, P.dist_init, P.poses_init)
solver.solve(max_num_iterations=3, **fixed)
solver.set_params(P.focal_init, P.pp_init, P.dist_init, P.poses_init)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
summ, _ = solver.solve(max_num_iterations=5, **fixed)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(model, summ["iterations"], summ["final_cost"])
