"""Opt-in overlap / angle filter for matched line pairs (BASELINE config 2): StereoFrame::lineSegmentOverlap
(stvo-pl/src/stereoFrame.cpp:521-627) + the direction test of matchGrid (matching.cpp:221).
CPU: the C restatement against an independent numpy form.  GPU: plm_line_pair_filter against the restatement."""
import numpy as np
import pytest

import oracle

port = oracle.port


def make_pairs(seed, n1=400, n2=380):
    rng = np.random.default_rng(seed)
    s = rng.uniform(0, 752, (n1, 2))
    d = rng.uniform(-120, 120, (n1, 2))
    ln1 = np.concatenate([s, s + d], 1)
    ln1[:40, 2] = ln1[:40, 0] + rng.uniform(-0.99, 0.99, 40)      # vertical (|dx| < 1)
    ln1[40:80, 3] = ln1[40:80, 1] + rng.uniform(-0.99, 0.99, 40)  # horizontal (|dy| < 1)
    ln1[80:85, 2:] = ln1[80:85, :2]                               # zero length -> NaN direction / division by zero
    m12 = rng.integers(-1, n2, n1).astype(np.int32)
    m12[::9] = -1
    m12[5] = n2 + 3                                               # out of range: treated as unmatched
    ln2 = rng.uniform(0, 752, (n2, 4))
    hit = m12[(m12 >= 0) & (m12 < n2)]
    # most matched partners are shifted / shortened / rotated copies so that every overlap branch occurs
    for i1 in range(n1):
        i2 = m12[i1]
        if 0 <= i2 < n2 and i1 % 4:
            t0, t1 = np.sort(rng.uniform(-0.6, 1.6, 2))
            p0 = ln1[i1, :2] + t0 * (ln1[i1, 2:] - ln1[i1, :2]) + rng.normal(0, 2, 2)
            p1 = ln1[i1, :2] + t1 * (ln1[i1, 2:] - ln1[i1, :2]) + rng.normal(0, 2, 2)
            ln2[i2] = np.concatenate([p1, p0] if i1 % 8 == 1 else [p0, p1])
    return ln1.astype(np.float32), ln2.astype(np.float32), m12


def numpy_filter(ln1, ln2, m12, overlap_th, sim_th):
    """Vectorised float64 form with the same operation order (numpy does not contract to FMA)."""
    n1, n2 = len(ln1), len(ln2)
    ok = (m12 >= 0) & (m12 < n2)
    j = np.where(ok, m12, 0)
    o, p = ln1.astype(np.float64), ln2.astype(np.float64)[j]
    sox, soy, eox, eoy = o.T
    spx, spy, epx, epy = p.T
    with np.errstate(all="ignore"):
        l0, l1 = eox - sox, eoy - soy

        def lam(ls, le):
            lo, hi = np.where(le < ls, le, ls), np.where(ls < le, le, ls)   # std::min / std::max
            return np.select([(lo < 0) & (hi > 1), (hi < 0) | (lo > 1), lo < 0, hi > 1], [1.0, 0.0, hi, 1.0 - lo], hi - lo)
        vert = lam((spy - soy) / l1, (epy - soy) / l1)
        horz = lam((spx - sox) / l0, (epx - sox) / l0)
        a, b = soy - eoy, eox - sox
        c = sox * eoy - eox * soy
        lxy = 1.0 / (a * a + b * b)
        sx = (b * (b * spx - a * spy) - a * c) * lxy
        ex = (b * (b * epx - a * epy) - a * c) * lxy
        gen = lam((sx - sox) / l0, (ex - sox) / l0)
        ov = np.where(np.abs(sox - eox) < 1.0, vert, np.where(np.abs(soy - eoy) < 1.0, horz, gen))
        vx, vy, wx, wy = eox - sox, eoy - soy, epx - spx, epy - spy
        mv, mw = np.sqrt(vx * vx + vy * vy), np.sqrt(wx * wx + wy * wy)
        sim = np.abs((vx / mv) * (wx / mw) + (vy / mv) * (wy / mw))
    keep = ok & (ov > overlap_th) & ~(sim < sim_th)
    return keep.astype(np.uint8), np.where(ok, ov, 0.0), np.where(ok, sim, 0.0)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_port_vs_numpy(seed):
    ln1, ln2, m12 = make_pairs(seed)
    n, keep, ov, sim = port.line_pair_filter(ln1, ln2, m12, 0.75, 0.75)
    k2, o2, s2 = numpy_filter(ln1, ln2, m12, 0.75, 0.75)
    assert np.array_equal(keep, k2) and n == int(k2.sum())
    assert np.array_equal(ov.view(np.uint64), o2.view(np.uint64))
    assert np.array_equal(sim.view(np.uint64), s2.view(np.uint64))
    assert 0 < n < (m12 >= 0).sum()                       # both outcomes occur
    assert np.isnan(sim[80:85][m12[80:85] >= 0]).all()    # zero-length segments: NaN similarity


def same_bits_or_nan(a, b):
    """Bit-exact where finite / infinite; NaN must meet NaN (x86 and the GPU differ in the sign / payload of a
    generated NaN, which no comparison of the filter can observe)."""
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a[~na].view(np.uint64), b[~nb].view(np.uint64))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_gpu_line_pair_filter(seed):
    from pl_inertial_slam_b200 import matching as M
    ln1, ln2, m12 = make_pairs(seed)
    for ov_th, sim_th in ((0.75, 0.75), (0.3, 0.95), (0.0, 0.0)):
        n_o, k_o, ov_o, s_o = port.line_pair_filter(ln1, ln2, m12, ov_th, sim_th)
        n_g, k_g, ov_g, s_g = M.line_pair_filter(ln1, ln2, m12, ov_th, sim_th)
        assert n_g == n_o and np.array_equal(k_g, k_o)
        assert same_bits_or_nan(ov_g, ov_o)
        assert same_bits_or_nan(s_g, s_o)
    n_g, k_g, _, _ = M.line_pair_filter(ln1[:0], ln2, m12[:0])
    assert n_g == 0 and len(k_g) == 0
