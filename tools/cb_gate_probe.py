import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "cb_harness"))
import numpy as np, cv2
from ezcalib_b200 import detect, lm
from erased import FLAGS, gate_corpus
import cb_pipeline as P, fast_check as FC
G = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "chessboard_golden.npz"))
ctx = lm.Context(0)
batch = gate_corpus(G)
imgs = detect.Images(ctx, np.stack(batch))
f_gate, c_gate = detect.find_chessboard_corners(ctx, imgs, (9, 6), False, fast_check=True)
f_full, c_full = detect.find_chessboard_corners(ctx, imgs, (9, 6), False, fast_check=False)
for k, im in enumerate(batch):
    ok_gate, _ = cv2.findChessboardCorners(im, (9, 6), flags=FLAGS + cv2.CALIB_CB_FAST_CHECK)
    ok_full, _ = cv2.findChessboardCorners(im, (9, 6), flags=FLAGS)
    thr = P.binarize_hist_threshold(im); B = np.where(im >= thr, 255, 0).astype(np.uint8) if thr > 0 else im.copy()
    pyb = FC.check_chessboard_binary(B, (9, 6)); pyg = pyb or FC.check_chessboard(im, (9, 6))
    mark = "" if (bool(f_gate[k]) == bool(ok_gate) and bool(f_full[k]) == bool(ok_full)) else "  <== MISMATCH"
    print(k, "cv2 gate/full", ok_gate, ok_full, "| gpu gate/full", bool(f_gate[k]), bool(f_full[k]), "| py gate binary/any", pyb, pyg, mark)
for k, im in enumerate(batch):
    ok_gate, c_ref = cv2.findChessboardCorners(im, (9, 6), flags=FLAGS + cv2.CALIB_CB_FAST_CHECK)
    if ok_gate and f_gate[k]:
        d = np.abs(c_gate[k] - c_ref.reshape(-1, 2)).max()
        if d > 0.001 or not np.array_equal(c_gate[k], c_full[k]): print("corner diff", k, d, np.array_equal(c_gate[k], c_full[k]))
