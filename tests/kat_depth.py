"""Known-answer hashes (adler32 of the output bytes) for the cvtColor families added late in round 2 -- 16-bit / float images and the edge-aware
Bayer codes -- on inputs derived from the committed fixture tests/golden/cvtcolor_kat_input.npy (the reference's own RNG(0) 263 x 255 8UC3 image).
Generated from the compiled reference (oracle/_ref) in the build container; tests/test_oracle.py checks that the reference still reproduces them
(CPU), tests/test_gpu_color_depths.py that the kernels do (GPU).  The reference does not exist on the GPU box: these constants do."""
import os
import zlib

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")

KAT_DEPTH = {
    (6, "f32"): 0xcc00c0bd,   # BGR2GRAY
    (6, "u16"): 0x70d00809,
    (7, "f32"): 0xca4268c3,   # RGB2GRAY
    (7, "u16"): 0xa9a8fb64,
    (36, "f32"): 0xdef3d5fd,  # BGR2YCrCb
    (36, "u16"): 0x17656878,
    (39, "f32"): 0x0b1e5c7f,  # YCrCb2RGB
    (39, "u16"): 0x22f4883a,
    (82, "f32"): 0x9809d173,  # BGR2YUV
    (82, "u16"): 0x9a2f6e6e,
    (84, "f32"): 0x790ea326,  # YUV2BGR
    (84, "u16"): 0xd29093ac,
    (32, "f32"): 0x856be209,  # BGR2XYZ
    (32, "u16"): 0x07c3a247,
    (35, "f32"): 0xbff5c0e4,  # XYZ2RGB
    (35, "u16"): 0xa22883fd,
    (2, "f32"): 0xe7f59c12,   # BGR2RGBA
    (2, "u16"): 0x80f6e33f,
    (40, "f32"): 0xb66ab386,  # BGR2HSV
    (67, "f32"): 0xbe2f7f64,  # RGB2HSV_FULL
    (54, "hsv"): 0x9e1a8762,  # HSV2BGR       (input: hue = channel 0 * 360)
    (71, "hsv"): 0xddd18762,  # HSV2RGB_FULL
    (135, "mosaic8"): 0x810c0ce8,   # BayerBG2BGR_EA on channel 0 of the fixture
    (135, "mosaic16"): 0x5d4096ba,
    (138, "mosaic8"): 0x9a39b60c,   # BayerGR2BGR_EA
    (138, "mosaic16"): 0xd76ce7bd,
}


def kat_input(kind):
    img = np.load(os.path.join(GOLD, "cvtcolor_kat_input.npy"))
    f = (img.astype(np.float32) * np.float32(1 / 255)).astype(np.float32)
    if kind == "f32":
        return f
    if kind == "u16":
        return img.astype(np.uint16) * np.uint16(257)
    if kind == "hsv":
        return np.stack([f[..., 0] * np.float32(360), f[..., 1], f[..., 2]], -1).astype(np.float32)
    mos = np.ascontiguousarray(img[:, :, 0])
    return mos if kind == "mosaic8" else mos.astype(np.uint16) * np.uint16(257)


def kat_dcn(code):
    return 1 if code in (6, 7) else 4 if code == 2 else 3


def kat_hash(a):
    return zlib.adler32(np.ascontiguousarray(a).tobytes())
