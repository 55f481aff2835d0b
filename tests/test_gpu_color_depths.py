"""GPU parity: cvtColor on CV_16U and CV_32F images (cvtcolor_depth.cu) against the compiled reference -- BIT-EXACT, floats included: the
kernel reproduces the reference's vector-body / scalar-tail split of every row (the last width % 8 float pixels run GCC's contraction of
the scalar statement; XYZ lives in a baseline-ISA unit: 4 lanes, no FMA).  Widths with and without tails, 3 and 4 channels, through the device ABI and through the host (HAL) path."""
import numpy as np
import pytest

import opencv_b200 as C
from util import assert_exact, cpu, gpu

pytestmark = pytest.mark.gpu

FROM3 = [(C.COLOR_BGR2BGRA, 4), (C.COLOR_BGR2RGBA, 4), (C.COLOR_BGR2RGB, 3), (C.COLOR_BGR2GRAY, 1), (C.COLOR_RGB2GRAY, 1),
         (C.COLOR_BGR2YCrCb, 3), (C.COLOR_RGB2YCrCb, 3), (C.COLOR_BGR2YUV, 3), (C.COLOR_RGB2YUV, 3),
         (C.COLOR_YCrCb2BGR, 3), (C.COLOR_YCrCb2RGB, 3), (C.COLOR_YUV2BGR, 3), (C.COLOR_YUV2RGB, 3), (C.COLOR_YCrCb2BGR, 4), (C.COLOR_YUV2RGB, 4),
         (C.COLOR_BGR2XYZ, 3), (C.COLOR_RGB2XYZ, 3), (C.COLOR_XYZ2BGR, 3), (C.COLOR_XYZ2RGB, 3), (C.COLOR_XYZ2BGR, 4)]
FROM4 = [(C.COLOR_BGRA2BGR, 3), (C.COLOR_RGBA2BGR, 3), (C.COLOR_BGRA2RGBA, 4), (C.COLOR_BGRA2GRAY, 1), (C.COLOR_RGBA2GRAY, 1),
         (C.COLOR_BGR2YCrCb, 3), (C.COLOR_RGB2YUV, 3), (C.COLOR_BGR2XYZ, 3)]
FROM1 = [(C.COLOR_GRAY2BGR, 3), (C.COLOR_GRAY2BGRA, 4)]
SHAPES = [(37, 29), (64, 1024), (255, 263), (1, 1), (3, 8), (5, 6)]


def _img(rng, dtype, h, w, cn):
    if dtype == np.uint16:
        a = rng.integers(0, 65536, (h, w, cn), dtype=np.uint16)
    else:
        a = (rng.random((h, w, cn), dtype=np.float32) * 1.5 - 0.25).astype(np.float32)       # beyond [0, 1] on both sides: nothing clamps
    return a[..., 0] if cn == 1 else a


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("scn,code,dcn", [(3,) + c for c in FROM3] + [(4,) + c for c in FROM4] + [(1,) + c for c in FROM1])
def test_cvt_depths(cvb, ref, rng, dtype, shape, scn, code, dcn):
    img = _img(rng, dtype, shape[0], shape[1], scn)
    got = cpu(cvb.cvtColor(gpu(img), code, dcn))
    assert got.dtype == dtype
    assert_exact(got, ref.cvtColor(img, code, dcn), "cvtColor %s code=%d scn=%d dcn=%d %s" % (np.dtype(dtype).name, code, scn, dcn, shape))


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_cvt_depths_host_path_and_4k(cvb, ref, rng, dtype):
    """numpy in, numpy out (b200cv_host_cvt_color = what the HAL entries call) on a 4K frame"""
    img = _img(rng, dtype, 2160, 3840, 3)
    for code, dcn in ((C.COLOR_BGR2GRAY, 1), (C.COLOR_BGR2YUV, 3), (C.COLOR_YCrCb2RGB, 3)):
        got = cvb.cvtColor(img, code, dcn)
        assert isinstance(got, np.ndarray)
        assert_exact(got, ref.cvtColor(img, code, dcn), "host cvtColor %s code=%d" % (np.dtype(dtype).name, code))


HSV_FWD = [(C.COLOR_BGR2HSV, 3), (C.COLOR_RGB2HSV, 3), (C.COLOR_BGR2HSV_FULL, 3), (C.COLOR_RGB2HSV_FULL, 3)]
HSV_INV = [(C.COLOR_HSV2BGR, 3), (C.COLOR_HSV2RGB, 3), (C.COLOR_HSV2BGR_FULL, 3), (C.COLOR_HSV2RGB_FULL, 4), (C.COLOR_HSV2BGR, 4)]


@pytest.mark.parametrize("shape", SHAPES + [(480, 640)])
def test_cvt_float_hsv(cvb, ref, rng, shape):
    """float HSV both ways, bit-exact: 8-lane vector body (fma(hsel, 60 / (diff + eps), res)) and the scalar tail (double division, the
    compiler's contractions, + 360 for negative hues; s == 0 shortcut, floored hue) are separate code paths in the reference and in the kernel"""
    h, w = shape
    bgr = _img(rng, np.float32, h, w, 3)
    if h > 8 and w > 8:
        bgr[2:5, 1:7] = bgr[2:5, 1:7, :1]                        # greys: diff == 0
        bgr[5, :6] = [[.2, .2, .7], [.7, .2, .2], [.2, .7, .7], [.7, .7, .2], [0, 0, 0], [1, 1, 1]]       # ties between the maxima
    for scn in (3, 4):
        src = bgr if scn == 3 else np.concatenate([bgr, bgr[..., :1]], -1)
        for code, dcn in HSV_FWD:
            assert_exact(cpu(cvb.cvtColor(gpu(src), code, dcn)), ref.cvtColor(src, code, dcn), "f32 to HSV code=%d scn=%d %s" % (code, scn, shape))
    hsv = np.stack([rng.random((h, w), dtype=np.float32) * np.float32(420) - np.float32(30),       # hues outside [0, 360) on both sides
                    rng.random((h, w), dtype=np.float32), rng.random((h, w), dtype=np.float32)], -1).astype(np.float32)
    if h > 8 and w > 8:
        hsv[3, :, 1] = 0                                          # s == 0
        hsv[4, :6, 0] = [0, 60, 120, 359.99997, 360, 720]
    for code, dcn in HSV_INV:
        assert_exact(cpu(cvb.cvtColor(gpu(hsv), code, dcn)), ref.cvtColor(hsv, code, dcn), "f32 from HSV code=%d dcn=%d %s" % (code, dcn, shape))


def test_cvt_depths_declined(cvb, rng):
    """families that exist only for 8-bit images say so (a stock OpenCV then runs its own code)"""
    img = gpu(_img(rng, np.uint16, 16, 16, 3))
    for code in (C.COLOR_BGR2HSV, C.COLOR_BGR2Lab, C.COLOR_HSV2BGR):
        with pytest.raises(Exception):
            cvb.cvtColor(img, code, 3)


def test_depth_known_answer_hashes(cvb):
    """the committed constants of tests/kat_depth.py (adler32 of the reference's output on inputs derived from the reference's own RNG(0) fixture):
    parity that does not need the reference library on the box"""
    from kat_depth import KAT_DEPTH, kat_dcn, kat_hash, kat_input
    for (code, kind), want in sorted(KAT_DEPTH.items()):
        got = cpu(cvb.cvtColor(gpu(kat_input(kind)), code, kat_dcn(code)))
        assert kat_hash(got) == want, "cvtColor code %d on %s: hash differs from the reference's" % (code, kind)
