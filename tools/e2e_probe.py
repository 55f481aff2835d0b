"""host-path throughput of one cheap op (GaussianBlur 3x3) on 16 4K u8 / 8 4K f32 frames, for a few pipeline settings
(set through the environment before the library is first used, hence one subprocess per setting)"""
import os
import subprocess
import sys

CHILD = r'''
import time, numpy as np, sys
sys.path.insert(0, ".")
from opencv_b200 import hal
for shape, dt in (((16, 2160, 3840, 1), np.uint8), ((8, 2160, 3840, 1), np.float32)):
    src = hal.pinned_empty(shape, dt); dst = hal.pinned_empty(shape, dt)
    src[...] = 7
    hal.GaussianBlur(src, (3, 3), 0, dst=dst)
    t0 = time.perf_counter()
    for _ in range(5):
        hal.GaussianBlur(src, (3, 3), 0, dst=dst)
    dt_s = (time.perf_counter() - t0) / 5
    print("   %s %-8s %.2f ms  %.1f GB/s each way" % (shape, np.dtype(dt).name, dt_s * 1e3, src.nbytes / dt_s / 1e9))
'''
for pipe, chunk in ((3, 32), (4, 32), (2, 32), (3, 8), (4, 64), (4, 128), (1, 1024)):
    env = dict(os.environ, B200CV_HOST_PIPE=str(pipe), B200CV_HOST_CHUNK_MB=str(chunk))
    print("pipe=%d chunk=%d MB" % (pipe, chunk), flush=True)
    subprocess.run([sys.executable, "-c", CHILD], env=env)
