"""quick device-side timing probe (development aid; bench.py is the contract)"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import opencv_b200 as cvb

cvb.init(0)
PEAK = 6568.0


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = {}
N = 32
u8 = torch.randint(0, 256, (N, 2160, 3840, 1), dtype=torch.uint8, device="cuda")
f32 = u8[:8].float()
out8 = torch.empty_like(u8)
out32 = torch.empty_like(f32)
# copy baseline
ms = timeit(lambda: out8.copy_(u8))
res["copy_u8"] = dict(ms=ms, gbs=2 * u8.numel() / ms / 1e6)
for k in (3, 5, 7, 9, 11, 13, 15, 21, 31):
    ms = timeit(lambda: cvb.GaussianBlur(u8, (k, k), 0, dst=out8))
    res["gauss_u8_k%d" % k] = dict(ms=ms, mpix=u8.numel() / ms / 1e3, gbs=2 * u8.numel() / ms / 1e6, frac=2 * u8.numel() / ms / 1e6 / PEAK)
    ms = timeit(lambda: cvb.GaussianBlur(f32, (k, k), 0, dst=out32))
    res["gauss_f32_k%d" % k] = dict(ms=ms, mpix=f32.numel() / ms / 1e3, gbs=8 * f32.numel() / ms / 1e6, frac=8 * f32.numel() / ms / 1e6 / PEAK)
import numpy as np
for k in (3, 5, 9, 15, 31):
    ker = np.random.rand(k, k).astype(np.float32); ker /= ker.sum()
    ms = timeit(lambda: cvb.filter2D(u8[:8], -1, ker, dst=out8[:8]), iters=3, warm=1)
    res["filter2d_u8_k%d" % k] = dict(ms=ms, mpix=u8[:8].numel() / ms / 1e3)
    ms = timeit(lambda: cvb.filter2D(f32, -1, ker, dst=out32), iters=3, warm=1)
    res["filter2d_f32_k%d" % k] = dict(ms=ms, mpix=f32.numel() / ms / 1e3)
del u8, f32, out8, out32
bgr = torch.randint(0, 256, (8, 4320, 7680, 3), dtype=torch.uint8, device="cuda")
for name, code, dcn in (("BGR2GRAY", 6, 1), ("BGR2YUV", 82, 3), ("BGR2HSV", 40, 3), ("YUV2BGR", 84, 3), ("HSV2BGR", 54, 3)):
    dst = torch.empty((8, 4320, 7680, dcn), dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: cvb.cvtColor(bgr, code, dcn, dst=dst))
    px = 8 * 4320 * 7680
    res["cvt_" + name] = dict(ms=ms, mpix=px / ms / 1e3, gbs=px * (3 + dcn) / ms / 1e6, frac=px * (3 + dcn) / ms / 1e6 / PEAK)
gray = torch.randint(0, 256, (8, 4320, 7680, 1), dtype=torch.uint8, device="cuda")
dst = torch.empty((8, 4320, 7680, 3), dtype=torch.uint8, device="cuda")
ms = timeit(lambda: cvb.cvtColor(gray, 8, 3, dst=dst))
px = 8 * 4320 * 7680
res["cvt_GRAY2BGR"] = dict(ms=ms, mpix=px / ms / 1e3, gbs=px * 4 / ms / 1e6, frac=px * 4 / ms / 1e6 / PEAK)
for k, v in res.items():
    print(k, json.dumps({a: round(b, 3) for a, b in v.items()}))
json.dump(res, open("gpurun_out/perf_probe.json", "w"), indent=1)
