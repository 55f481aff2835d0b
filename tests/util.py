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


def assert_exact_body(got, want, lanes, atol=0.0, rtol=0.0, what=""):
    """Float separable filters reproduce the operation order of the reference's SIMD loops, so every row element the reference
    computes in a full vector is bit-identical.  The last (W*cn mod lanes) elements of a row come from the reference's scalar
    remainder loops, whose rounding depends on how its compiler contracted them: those are held to the tolerance instead.
    lanes = 8 for float sources, 32 for 8-bit sources (measured against the reference build: filter.simd.hpp row/column loops)."""
    got = np.asarray(got); want = np.asarray(want)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (what, got.shape, want.shape)
    g2 = got.reshape(got.shape[0], -1) if got.ndim <= 3 else got.reshape(got.shape[0] * got.shape[1], -1)
    w2 = want.reshape(g2.shape)
    body = (g2.shape[1] // lanes) * lanes
    assert_exact(g2[:, :body], w2[:, :body], what + " [vector body]")
    if body < g2.shape[1]:
        assert_close(g2[:, body:], w2[:, body:], atol=atol, rtol=rtol, what=what + " [scalar remainder columns]")


def rand_u8(rng, h, w, cn=1):
    shape = (h, w) if cn == 1 else (h, w, cn)
    return rng.integers(0, 256, shape, dtype=np.uint8)
