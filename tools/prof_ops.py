"""run a fixed list of ops once each (after one warm-up) for ncu captures: python tools/prof_ops.py [names...]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import opencv_b200 as cvb

cvb.init(0)
which = set(sys.argv[1:])
dev = "cuda"
u8 = torch.randint(0, 256, (16, 2160, 3840, 1), dtype=torch.uint8, device=dev)
o8 = torch.empty_like(u8)
f32 = u8[:8].float()
o32 = torch.empty_like(f32)
oh = torch.empty((16, 2160, 3840, 1), dtype=torch.float32, device=dev)
bgr = torch.randint(0, 256, (2, 4320, 7680, 3), dtype=torch.uint8, device=dev)
obgr = torch.empty_like(bgr)
half = torch.empty((2, 2160, 3840, 3), dtype=torch.uint8, device=dev)
k5 = torch.empty((2, 2880, 5120, 3), dtype=torch.uint8, device=dev)
M = np.array([[0.8911, 0.1094, 182.0], [-0.1094, 0.8911, 655.0]])
templ = u8[0, 700:764, 1000:1064, 0].contiguous()
g3 = np.array([0.2, 0.55, 0.25], np.float32)
g5 = np.array([0.1, 0.2, 0.35, 0.25, 0.1], np.float32)
g7 = cvb.getGaussianKernel(7, 1.3).astype(np.float32).ravel()
g11 = cvb.getGaussianKernel(11, 0).astype(np.float32).ravel()
g31 = cvb.getGaussianKernel(31, 0).astype(np.float32).ravel()
ops = {
    "gauss_u8_k3": lambda: cvb.GaussianBlur(u8, (3, 3), 0, dst=o8),
    "gauss_u8_k5": lambda: cvb.GaussianBlur(u8, (5, 5), 0, dst=o8),
    "gauss_u8_k15": lambda: cvb.GaussianBlur(u8, (15, 15), 0, dst=o8),
    "gauss_f32_k5": lambda: cvb.GaussianBlur(f32, (5, 5), 0, dst=o32),
    "gauss_f32_k15": lambda: cvb.GaussianBlur(f32, (15, 15), 0, dst=o32),
    "harris": lambda: cvb.cornerHarris(u8, 2, 3, 0.04, dst=oh),
    "resize_nn": lambda: cvb.resize(bgr, (3840, 2160), interpolation=0, dst=half),
    "resize_lin5k": lambda: cvb.resize(bgr, (5120, 2880), interpolation=1, dst=k5),
    "warp_lin": lambda: cvb.warpAffine(bgr, M, (7680, 4320), 1, dst=obgr),
    "warp_nn": lambda: cvb.warpAffine(bgr, M, (7680, 4320), 0, dst=obgr),
    "hsv2bgr": lambda: cvb.cvtColor(bgr, 54, dst=obgr),
    "bgr2yuv": lambda: cvb.cvtColor(bgr, 82, dst=obgr),
    "match": lambda: cvb.matchTemplate(u8[:2], templ, 2),
    "gauss_f32_k11": lambda: cvb.GaussianBlur(f32, (11, 11), 0, dst=o32),
    "gauss_f32_k3": lambda: cvb.GaussianBlur(f32, (3, 3), 0, dst=o32),
    "sep_u8_k11": lambda: cvb.sepFilter2D(u8, -1, g11, g11, dst=o8),
    "sep_u8_k31": lambda: cvb.sepFilter2D(u8, -1, g31, g31, dst=o8),
    "filter2d_u8_k31": lambda: cvb.filter2D(u8, -1, np.outer(g31, g31), dst=o8),
    "filter2d_u8_k11": lambda: cvb.filter2D(u8, -1, np.outer(g11, g11), dst=o8),
    "filter2d_f32_k31": lambda: cvb.filter2D(f32, -1, np.outer(g31, g31), dst=o32),
    "filter2d_u8_k7": lambda: cvb.filter2D(u8, -1, np.outer(g7, g7), dst=o8),
    "filter2d_u8_k11d": lambda: cvb.filter2D(u8, -1, np.outer(g11, g11) + 1e-4, dst=o8),
    "filter2d_u8_k3": lambda: cvb.filter2D(u8, -1, np.outer(g3, g3), dst=o8),
    "filter2d_f32_k5": lambda: cvb.filter2D(f32, -1, np.outer(g5, g5), dst=o32),
    "warp_cub": lambda: cvb.warpAffine(bgr, M, (7680, 4320), 2, dst=obgr),
    "resize_cub5k": lambda: cvb.resize(bgr, (5120, 2880), interpolation=2, dst=k5),
    "resize_lin_up": lambda: cvb.resize(half, (7680, 4320), interpolation=1, dst=obgr),
}
for name, fn in ops.items():
    if which and name not in which:
        continue
    fn(); torch.cuda.synchronize()
    fn(); torch.cuda.synchronize()
print("done")
