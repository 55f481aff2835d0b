import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import opencv_b200 as cvb
cvb.init(0)
rng = np.random.default_rng(5)
for (H, W, h, w) in ((300, 500, 64, 64), (2160, 3840, 64, 64), (257, 321, 17, 23), (400, 400, 100, 64)):
    img = rng.integers(0, 256, (H, W), dtype=np.uint8); tp = img[5:5 + h, 9:9 + w].copy()
    ti, tt = torch.from_numpy(img).cuda(), torch.from_numpy(tp).cuda()
    os.environ["B200CV_MATCHTEMPLATE_PATH"] = "dp4a"
    a = cvb.matchTemplate(ti, tt, cvb.TM_CCORR).cpu().numpy()
    os.environ["B200CV_MATCHTEMPLATE_PATH"] = "tc"
    b = cvb.matchTemplate(ti, tt, cvb.TM_CCORR).cpu().numpy()
    d = np.abs(a.astype(np.float64) - b)
    print((H, W, h, w), "max diff", d.max(), "nbad", int((d > 0).sum()), "of", d.size, "first bad", np.argwhere(d > 0)[:3].tolist())
u8 = torch.randint(0, 256, (16, 2160, 3840, 1), dtype=torch.uint8, device="cuda")
tt = u8[0, 700:764, 1000:1064, 0].contiguous()
res = torch.empty((16, 2097, 3777, 1), dtype=torch.float32, device="cuda")
for path in ("dp4a", "tc"):
    os.environ["B200CV_MATCHTEMPLATE_PATH"] = path
    for _ in range(2): cvb.matchTemplate(u8, tt, cvb.TM_CCORR, result=res)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): cvb.matchTemplate(u8, tt, cvb.TM_CCORR, result=res)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3 / 16
    print(path, "ms/frame %.4f" % ms, "TMAC/s useful %.1f" % (3.244e10 / ms / 1e9))
