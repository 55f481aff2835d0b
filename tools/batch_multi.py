#!/usr/bin/env python3
"""One PROCESS driving every GPU of the box through the batch driver (include/b200cv_batch.h): worker thread + streams per device inside the
library, NUMA-local page-locked frame blocks, ncclBroadcast of the template.

    gpurun --gpus 8 -- 'python tools/batch_multi.py > gpurun_out/batch_multi.txt'

1. C5 (SIFT pyramid + DoG + Harris, 4K 8UC1) over 32 frames per device, Harris responses downloaded: frames/s at 1, 2, 4, ... all devices.
2. A copy-bound op (GaussianBlur 3x3 8UC1 4K, 2 bytes per pixel over PCIe each way) on device subsets: which GPUs share a PCIe uplink
   (per-device GB/s halves when two GPUs behind one switch stream at once) -- the limiter of end-to-end scaling on this box.
3. matchTemplate with the template broadcast over NCCL, checked against the one-device result."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opencv_b200 as C                      # noqa: E402
from opencv_b200.batch import BatchDriver    # noqa: E402

W, H = 3840, 2160


def timed(fn, reps=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def c5(devs, per_dev=32):
    d = BatchDriver(devs)
    n = d.n_devices * per_dev
    src = d.pinned_frames((n, H, W, 1), np.uint8)
    rng = np.random.default_rng(5)
    base = (np.kron(rng.random((H // 8, W // 8)), np.ones((8, 8))) * 255).astype(np.uint8)
    for f in range(n):
        src[f, :, :, 0] = np.roll(base, (17 * f, 31 * f), (0, 1))
    har = d.pinned_frames((n, H, W, 1), np.float32)
    t = timed(lambda: d.sift_harris(src, har, wave=4))
    t2 = timed(lambda: d.sift_harris(src, None, wave=4))
    print("C5  devices %-18s %4d frames  %.1f frames/s (%.0f Mpix/s, 2 ops) with Harris download;  %.1f frames/s responses kept on device   counts %s"
          % (devs, n, n / t, 2 * n * W * H / t / 1e6, n / t2, d.last_counts()), flush=True)
    d.close()
    return n / t


def copy_bound(devs, per_dev=48):
    d = BatchDriver(devs)
    n = d.n_devices * per_dev
    src = d.pinned_frames((n, H, W, 1), np.uint8)
    src[...] = 7
    dst = d.pinned_frames((n, H, W, 1), np.uint8)
    t = timed(lambda: d.GaussianBlur(src, (3, 3), 0, dst=dst), reps=3)
    gbs = per_dev * W * H / t / 1e9
    print("PCIe devices %-18s GaussianBlur 3x3 8UC1: %.1f GB/s per device each way (%.0f Mpix/s total)" % (devs, gbs, n * W * H / t / 1e6), flush=True)
    d.close()
    return gbs


def main():
    n = C.lib().b200cv_device_count()
    print("devices visible:", n, flush=True)
    os.system("nvidia-smi topo -m | head -12")
    quick = "--quick" in sys.argv
    sets = [[0]]
    k = 2
    while k <= n:
        sets.append(list(range(k)))
        k *= 2
    if quick:
        sets = [[0], list(range(n))] if n > 1 else [[0]]
    base = None
    for s in sets:
        r = c5(s)
        base = base or r
        print("     scaling vs 1 device: %.2fx" % (r / base), flush=True)
    one = copy_bound([0])
    for j in range(1, min(n, 5) if quick else n):
        copy_bound([0, j])
    for s in sets[1:]:
        g = copy_bound(s)
        print("     per-device share vs alone: %.2f" % (g / one), flush=True)
    if n > 1:
        d1, dn = BatchDriver([0]), BatchDriver(None)
        rng = np.random.default_rng(1)
        img = rng.integers(0, 256, (2 * n + 1, 540, 960, 1), dtype=np.uint8)
        t = np.ascontiguousarray(img[1, 100:164, 200:264, 0])
        a, b = d1.matchTemplate(img, t, C.TM_CCORR_NORMED), dn.matchTemplate(img, t, C.TM_CCORR_NORMED)
        print("matchTemplate over %d devices (template by ncclBroadcast: %s) == one device: %s" % (n, dn.uses_nccl, np.array_equal(a, b)), flush=True)
        d1.close(); dn.close()


if __name__ == "__main__":
    main()
