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


nv12 = torch.randint(0, 256, (4, 4320 * 3 // 2, 7680, 1), dtype=torch.uint8, device=dev)      # 4 x 8K NV12 / I420 frames
yuy2 = torch.randint(0, 256, (4, 4320, 7680, 2), dtype=torch.uint8, device=dev)
k5 = torch.empty((4, 2880, 5120, 3), dtype=torch.uint8, device=dev)
k25 = torch.empty((4, 1440, 2560, 3), dtype=torch.uint8, device=dev)
u16 = (torch.randint(0, 65536, (8, 2160, 3840, 1), dtype=torch.int32, device=dev)).to(torch.uint16)
o16 = torch.empty_like(u16)
_sm = torch.nn.functional.interpolate(torch.rand((1, 1, 135, 240), device=dev), size=(1080, 1920), mode="bilinear")
sift_img = (_sm[0, 0] * 255).to(torch.uint8).contiguous()
f32c3 = torch.rand((4, 2160, 3840, 3), device=dev)
of32c3 = torch.empty_like(f32c3)
u16c3 = torch.randint(0, 65536, (8, 2160, 3840, 3), dtype=torch.int32, device=dev).to(torch.uint16)
ou16c3 = torch.empty_like(u16c3)
smooth4k = (torch.nn.functional.interpolate(torch.rand((4, 1, 270, 480), device=dev), size=(2160, 3840), mode="bicubic").clamp(0, 1) * 255).to(torch.uint8).reshape(4, 2160, 3840, 1).contiguous()

ops = {
    "blur_u8_k3": (lambda: cvb.blur(u8, (3, 3), dst=o8), nbytes(u8, o8)),
    "blur_u8_k5": (lambda: cvb.blur(u8, (5, 5), dst=o8), nbytes(u8, o8)),
    "blur_u8_k21": (lambda: cvb.blur(u8, (21, 21), dst=o8), nbytes(u8, o8)),
    "blur_u8c3_8k_k5": (lambda: cvb.blur(bgr, (5, 5), dst=obgr), nbytes(bgr, obgr)),
    "blur_f32_k5": (lambda: cvb.blur(f32, (5, 5), dst=o32), nbytes(f32, o32)),
    "box_u8_f32_k7": (lambda: cvb.boxFilter(u8[:8], 5, (7, 7), dst=o32), nbytes(u8[:8], o32)),
    "pyrdown_u8c3_8k": (lambda: cvb.pyrDown(bgr), nbytes(bgr) * 5 // 4),
    "scharr_u8_s16": (lambda: cvb.Scharr(u8, 3, 1, 0), nbytes(u8) * 3),
    # written after the round-1 GPU budget was spent: first numbers belong to round 2
    "nv12_to_bgr_8k": (lambda: cvb.cvtColor(nv12, cvb.COLOR_YUV2BGR_NV12, dst=obgr), nbytes(nv12, obgr)),
    "i420_to_bgr_8k": (lambda: cvb.cvtColor(nv12, cvb.COLOR_YUV2BGR_I420, dst=obgr), nbytes(nv12, obgr)),
    "yuy2_to_bgr_8k": (lambda: cvb.cvtColor(yuy2, cvb.COLOR_YUV2BGR_YUY2, dst=obgr), nbytes(yuy2, obgr)),
    "bgr_to_i420_8k": (lambda: cvb.cvtColor(bgr, cvb.COLOR_BGR2YUV_I420, dst=nv12), nbytes(bgr, nv12)),
    "bayer_to_bgr_4k": (lambda: cvb.cvtColor(u8, cvb.COLOR_BayerRG2BGR), nbytes(u8) * 4),
    "area_8k_to_5k": (lambda: cvb.resize(bgr, (5120, 2880), interpolation=cvb.INTER_AREA, dst=k5), nbytes(bgr, k5)),
    "area_8k_div3": (lambda: cvb.resize(bgr, (2560, 1440), interpolation=cvb.INTER_AREA, dst=k25), nbytes(bgr, k25)),
    "lanczos4_8k_to_5k": (lambda: cvb.resize(bgr, (5120, 2880), interpolation=cvb.INTER_LANCZOS4, dst=k5), nbytes(bgr, k5)),
    "linear_exact_8k_to_5k": (lambda: cvb.resize(bgr, (5120, 2880), interpolation=cvb.INTER_LINEAR_EXACT, dst=k5), nbytes(bgr, k5)),
    "nearest_exact_8k_to_5k": (lambda: cvb.resize(bgr, (5120, 2880), interpolation=cvb.INTER_NEAREST_EXACT, dst=k5), nbytes(bgr, k5)),
    "bgr_to_lab_8k": (lambda: cvb.cvtColor(bgr, cvb.COLOR_BGR2Lab, dst=obgr), nbytes(bgr, obgr)),
    "lab_to_bgr_8k": (lambda: cvb.cvtColor(bgr, cvb.COLOR_Lab2BGR, dst=obgr), nbytes(bgr, obgr)),
    "match_masked_4k_64": (lambda: cvb.matchTemplate(u8[:1], u8[0, 700:764, 1000:1064, 0].contiguous(), 5, mask=torch.ones((64, 64), dtype=torch.uint8, device=dev)), nbytes(u8[:1]) * 5),
    "integral_4k": (lambda: cvb.integral(u8), nbytes(u8) * 5),
    "integral_sq_4k": (lambda: cvb.integral(u8[:4], with_sqsum=True), nbytes(u8[:4]) * 13),
    "gauss_u16_k5": (lambda: cvb.GaussianBlur(u16, (5, 5), 0, dst=o16), nbytes(u16, o16)),
    "gauss_u16_k15": (lambda: cvb.GaussianBlur(u16, (15, 15), 0, dst=o16), nbytes(u16, o16)),
    "gftt_4k_noise": (lambda: cvb.goodFeaturesToTrack(u8[:4], 1000, 0.01, 10, 3, 3, True, 0.04), nbytes(u8[:4])),
    "gftt_4k_smooth": (lambda: cvb.goodFeaturesToTrack(smooth4k, 1000, 0.01, 10, 3, 3, True, 0.04), nbytes(smooth4k)),
    "cvt_f32_bgr2gray_4k": (lambda: cvb.cvtColor(f32c3, cvb.COLOR_BGR2GRAY, dst=o32[:4]), nbytes(f32c3, o32[:4])),
    "cvt_f32_bgr2yuv_4k": (lambda: cvb.cvtColor(f32c3, cvb.COLOR_BGR2YUV, dst=of32c3), nbytes(f32c3, of32c3)),
    "cvt_f32_yuv2bgr_4k": (lambda: cvb.cvtColor(f32c3, cvb.COLOR_YUV2BGR, dst=of32c3), nbytes(f32c3, of32c3)),
    "cvt_u16_bgr2gray_4k": (lambda: cvb.cvtColor(u16c3, cvb.COLOR_BGR2GRAY, dst=o16), nbytes(u16c3, o16)),
    "cvt_u16_bgr2ycrcb_4k": (lambda: cvb.cvtColor(u16c3, cvb.COLOR_BGR2YCrCb, dst=ou16c3), nbytes(u16c3, ou16c3)),
    "cvt_u16_bgr2rgb_4k": (lambda: cvb.cvtColor(u16c3, cvb.COLOR_BGR2RGB, dst=ou16c3), nbytes(u16c3, ou16c3)),
    "sift_detect_1080p": (lambda: cvb.sift_detectAndCompute(sift_img), nbytes(sift_img) * 5),
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
