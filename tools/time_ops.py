"""CUDA-event timing of the section-8(f) ops that are not in bench.py's workloads: python tools/time_ops.py [names...]
Prints ms per launch and algorithmic GB/s (source bytes read once + destination bytes written once)."""
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
bgr = torch.randint(0, 256, (4, 4320, 7680, 3), dtype=torch.uint8, device=dev)
obgr = torch.empty_like(bgr)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def nbytes(*ts):
    return sum(t.numel() * t.element_size() for t in ts)


ops = {
    "blur_u8_k3": (lambda: cvb.blur(u8, (3, 3), dst=o8), nbytes(u8, o8)),
    "blur_u8_k5": (lambda: cvb.blur(u8, (5, 5), dst=o8), nbytes(u8, o8)),
    "blur_u8_k21": (lambda: cvb.blur(u8, (21, 21), dst=o8), nbytes(u8, o8)),
    "blur_u8c3_8k_k5": (lambda: cvb.blur(bgr, (5, 5), dst=obgr), nbytes(bgr, obgr)),
    "blur_f32_k5": (lambda: cvb.blur(f32, (5, 5), dst=o32), nbytes(f32, o32)),
    "box_u8_f32_k7": (lambda: cvb.boxFilter(u8[:8], 5, (7, 7), dst=o32), nbytes(u8[:8], o32)),
    "pyrdown_u8c3_8k": (lambda: cvb.pyrDown(bgr), nbytes(bgr) * 5 // 4),
    "scharr_u8_s16": (lambda: cvb.Scharr(u8, 3, 1, 0), nbytes(u8) * 3),
}
for name, (fn, nb) in ops.items():
    if which and name not in which:
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()                       # 256 MB > L2: every timed launch starts cold
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    print("%-18s %8.3f ms  %8.1f GB/s algorithmic" % (name, ms, nb / ms / 1e6), flush=True)
