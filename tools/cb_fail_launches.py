"""Launch list of the chessboard detector on ONE image that fails (board partly occluded => all 8 + 48 attempts run)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cv2
from ezcalib_b200 import detect, lm

G = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "chessboard_golden.npz"))
img = G["render0_img"].copy()
ok, c = cv2.findChessboardCorners(img, (9, 6))
x, y = c.reshape(-1, 2)[0]
cv2.circle(img, (int(x), int(y)), 40, 128, -1)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 2
img = cv2.resize(img, None, fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR)
ctx = lm.Context(0)
imgs = detect.Images(ctx, img)
found, corners = detect.find_chessboard_corners(ctx, imgs, (9, 6), True)
torch.cuda.synchronize()
t = time.perf_counter()
torch.cuda.cudart().cudaProfilerStart()
found, corners = detect.find_chessboard_corners(ctx, imgs, (9, 6), True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
dt = time.perf_counter() - t
t = time.perf_counter()
ok2, _ = cv2.findChessboardCorners(img, (9, 6), flags=cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK)
print("image", img.shape, "gpu found", bool(found[0]), f"{dt*1e3:.1f} ms | cv2 found", ok2, f"{(time.perf_counter()-t)*1e3:.1f} ms")
