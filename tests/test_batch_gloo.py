"""CPU: the multi-GPU host logic (frame sharding + operand broadcast) under a world_size-2 gloo group."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from opencv_b200 import batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = batch.shard_range(1024 + 3, rank, world)
    taps = np.arange(31, dtype=np.float32) if rank == 0 else np.zeros(31, np.float32)
    got = batch.broadcast_operand(taps, 0)
    counts = batch.gather_counts(hi - lo)
    q.put((rank, lo, hi, got.tolist(), counts))
    dist.destroy_process_group()


def test_shard_and_broadcast_world2():
    from opencv_b200 import batch
    assert [batch.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res[0][1:3] == (0, 514) and res[1][1:3] == (514, 1027)
    for r in res:
        assert r[3] == list(map(float, range(31)))       # every rank holds rank 0's operand
        assert r[4] == [514, 513]


def test_c_abi_shard_matches_python():
    """b200cv_batch_shard (the batch driver's frame blocks, include/b200cv_batch.h) = shard_range: pure host arithmetic, no GPU"""
    import ctypes
    import opencv_b200 as cvb
    from opencv_b200 import batch
    L = cvb.lib()
    f, c = ctypes.c_int(), ctypes.c_int()
    for frames, n in ((10, 4), (1027, 2), (1024, 8), (3, 8), (0, 2), (1, 1)):
        got = []
        for i in range(n):
            assert L.b200cv_batch_shard(frames, i, n, ctypes.byref(f), ctypes.byref(c)) == 0
            got.append((f.value, f.value + c.value))
        assert got == [batch.shard_range(frames, i, n) for i in range(n)]
    assert L.b200cv_batch_shard(5, 2, 2, ctypes.byref(f), ctypes.byref(c)) == -2


def test_batch_driver_refuses_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import opencv_b200 as cvb
    from opencv_b200.batch import BatchDriver
    with pytest.raises(cvb.B200cvError, match="no CPU fallback"):
        BatchDriver()
