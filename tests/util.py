"""helpers shared by the GPU parity tests"""
import numpy as np


def gpu(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def cpu(t):
    import torch
    torch.cuda.synchronize()
    return t.cpu().numpy()


def assert_exact(got, want, what=""):
    got = np.asarray(got); want = np.asarray(want)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (what, got.shape, want.shape)
    if not np.array_equal(got, want):
        d = np.abs(got.astype(np.float64) - want.astype(np.float64))
        idx = np.argwhere(d > 0)
        raise AssertionError("%s: %d/%d elements differ, max |d| = %g, first at %s (got %s want %s)" % (
            what, len(idx), d.size, d.max(), tuple(idx[0]), got[tuple(idx[0])], want[tuple(idx[0])]))


def assert_close(got, want, atol=0.0, rtol=0.0, what=""):
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (what, got.shape, want.shape)
    d = np.abs(got - want)
    lim = atol + rtol * np.abs(want)
    bad = d > lim
    if bad.any():
        idx = np.argwhere(bad)
        raise AssertionError("%s: %d/%d elements out of tolerance (atol %g rtol %g), max |d| = %g at %s (got %r want %r)" % (
            what, len(idx), d.size, atol, rtol, d.max(), tuple(np.unravel_index(d.argmax(), d.shape)),
            got[tuple(idx[0])], want[tuple(idx[0])]))


def rand_u8(rng, h, w, cn=1):
    shape = (h, w) if cn == 1 else (h, w, cn)
    return rng.integers(0, 256, shape, dtype=np.uint8)
