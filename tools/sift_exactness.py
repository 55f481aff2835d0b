"""SIFT Gaussian pyramid / DoG vs the reference: mismatching elements per octave (vector body = columns below floor8(width))"""
import sys

import numpy as np
import torch

sys.path.insert(0, "."); sys.path.insert(0, "tests")
import opencv_b200 as cvb
from oracle.api import Oracle, unpack_pyramid

cvb.init(0)
ref = Oracle("ref")
rng = np.random.default_rng(9)
for shape in ((540, 960), (135, 240), (100, 75)):
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    img = (np.asarray(img, np.float32) * 0.5 + np.roll(img, 3, 1) * 0.5).astype(np.uint8)
    wg, wd = ref.sift_pyramid(img, 3, 1.6, True)
    G, D, dims = cvb.sift_pyramid(torch.from_numpy(img).cuda(), 3, 1.6, True)
    torch.cuda.synchronize()
    gg, gd = unpack_pyramid(G.cpu().numpy()[0], D.cpu().numpy()[0], dims.reshape(-1), len(dims), 3)
    for o in range(len(dims)):
        w = wg[o][0].shape[1]; body = (w // 8) * 8
        nb = sum(int((gg[o][i][:, :body] != wg[o][i][:, :body]).sum()) for i in range(6))
        nt = sum(int((gg[o][i][:, body:] != wg[o][i][:, body:]).sum()) for i in range(6))
        nd = sum(int((gd[o][i][:, :body] != wd[o][i][:, :body]).sum()) for i in range(5))
        mx = max(float(np.abs(gg[o][i] - wg[o][i]).max()) for i in range(6))
        print("%s octave %d (%dx%d): gauss body %d, remainder cols %d, dog body %d, max |d| %.3g" % (shape, o, wg[o][0].shape[0], w, nb, nt, nd, mx))
