"""How many views are still searching at every attempt of findChessboardCorners (EZC_CB_TIMING=1 prints it): the bench's
cfg1 / cfg2 chessboard batches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["EZC_CB_TIMING"] = "1"
import argparse
import numpy as np
import torch
import bench
from ezcalib_b200 import lm
ctx = lm.Context(0, stream=torch.cuda.current_stream().cuda_stream)
args = argparse.Namespace(cb_images=400, no_cpu=True, warmup=1)
for cfg in ("cfg1", "cfg2"):
    print("=====", cfg, file=sys.stderr)
    out = bench.run_detect_chessboard(ctx, args, 0, 1, torch.cuda.current_stream(), cfg)
    print(cfg, out["value"], out["config"]["found"])
