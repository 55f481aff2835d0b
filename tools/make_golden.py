"""Regenerate the golden INPUT fixtures under tests/golden/ from the reference itself (oracle/_ref, built from /root/reference).

The reference's known-answer tests (modules/imgproc/test/test_color.cpp:2823-2900, runCvtColorBitExactCheck) hash cv::cvtColor's
output on cv::RNG(0).fill(UNIFORM, 0, 255) images; the hashes are constants of the reference's test file and live in
tests/test_oracle.py, the inputs are its RNG stream and live here as .npy files (the reference does not exist on the GPU box).
    python tools/make_golden.py          # needs oracle/_ref/libocvref.so
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.api import Oracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
FIXTURES = {
    "cvtcolor_kat_input.npy": (255, 263, 3),          # Size(263, 255) CV_8UC3: GRAY / YUV / HSV families
    "cvtcolor_kat_yuv420_input.npy": (510, 262),      # Size(262, 510) CV_8UC1: NV12 / NV21 / YV12 / IYUV -> BGR family, GRAY_420
    "cvtcolor_kat_yuv422_input.npy": (510, 262, 2),   # Size(262, 510) CV_8UC2: UYVY / YUY2 / YVYU
    "cvtcolor_kat_bgr_262x254_input.npy": (254, 262, 3),   # Size(262, 254) CV_8UC3: BGR family -> I420 / YV12
}

if __name__ == "__main__":
    ref = Oracle("ref")
    for name, shape in FIXTURES.items():
        a = ref.rng_fill(shape, np.uint8, 0, 0, 255)
        path = os.path.join(GOLD, name)
        if os.path.exists(path) and np.array_equal(np.load(path), a):
            print("unchanged", name)
            continue
        np.save(path, a)
        print("wrote", name, a.shape)
