import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "cb_harness"))
import numpy as np, cv2
from ezcalib_b200 import detect, lm
from erased import FLAGS, gate_corpus
import cb_pipeline as P
G = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "chessboard_golden.npz"))
ctx = lm.Context(0)
batch = gate_corpus(G)
for k in [int(a) for a in sys.argv[1:]]:
    im = batch[k]
    imgs = detect.Images(ctx, im)
    f, c = detect.find_chessboard_corners(ctx, imgs, (9, 6), False)
    imgs.close()
    info = {}
    okh, ch = P.find_chessboard_corners(im, (9, 6), info=info)
    okr, chr_ = P.find_chessboard_corners(im, (9, 6), refine=False)
    d = np.abs(c[0] - ch).max(axis=1)
    print(k, "gpu found", f[0], "harness", okh, info, "max diff", d.max(), "n>0.01", int((d > 0.01).sum()))
    for i in np.argsort(-d)[:4]:
        print("   corner", i, "gpu", c[0][i], "harness refined", ch[i], "harness raw", chr_[i])
