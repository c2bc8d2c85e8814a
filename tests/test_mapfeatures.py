"""MapPoint / MapLine::updateAverageDescDir (src/mapFeatures.cpp:51-93, :121-163).

CPU: the C restatement against the reference's own mapFeatures.cpp (compiled unmodified, oracle/Makefile),
against an independent numpy form, and against the golden vectors generated from the compiled reference.
GPU (-m gpu): plm_med_desc / plm_dev_med_desc through the C ABI against the restatement, bit for bit."""
import os

import numpy as np
import pytest

import oracle
from conftest import ROOT
from pl_inertial_slam_b200 import synth

port, ref = oracle.port, oracle.ref
needs_ref = pytest.mark.skipif(not (ref.available() and hasattr(ref.lib, "plref_med_desc")),
                               reason="oracle/_ref/libplref.so (with mapFeatures.cpp) not built")
GOLDEN = os.path.join(ROOT, "tests", "golden", "med_desc.npz")

CASES = [  # (seed offset, kwargs)
    (0, dict(n_lm=300, mean_obs=6)),
    (1, dict(n_lm=200, mean_obs=12, tie=True)),
    (2, dict(n_lm=120, mean_obs=20, max_obs=32, empty_frac=0.1)),
    (3, dict(n_lm=60, mean_obs=5, long_lists=6, long_len=70)),
    (4, dict(n_lm=40, mean_obs=3, long_lists=3, long_len=300, tie=True)),
    (5, dict(n_lm=1, mean_obs=1)),
]


def numpy_med_desc(desc, dirs, obs_start):
    """Independent form: sort-based order statistic on the full distance matrix."""
    bits = np.unpackbits(desc, axis=1).astype(np.int32)
    out_i, out_d = [], []
    for l in range(len(obs_start) - 1):
        lo, hi = int(obs_start[l]), int(obs_start[l + 1])
        n = hi - lo
        if n == 0:
            out_i.append(-1); out_d.append(np.zeros(3)); continue
        if n == 1:
            out_i.append(0); out_d.append(dirs[lo].copy()); continue
        b = bits[lo:hi]
        conf = (b[:, None, :] != b[None, :, :]).sum(-1)
        med = np.sort(conf, axis=1)[:, int(1 + 0.5 * (n - 1))]
        out_i.append(int(np.argmin(med)))  # first minimum
        s = np.zeros(3)
        for i in range(lo, hi):
            s = s + dirs[i]
        out_d.append(s / n)
    return np.array(out_i, np.int32), np.array(out_d)


@pytest.mark.parametrize("off,kw", CASES)
def test_port_vs_numpy(off, kw):
    desc, dirs, os_ = synth.make_landmark_observations(synth.SEED0 + 20 + off, **kw)
    i_p, med, d_p = port.med_desc(desc, dirs, os_)
    i_n, d_n = numpy_med_desc(desc, dirs, os_)
    assert np.array_equal(i_p, i_n)
    assert np.array_equal(d_p, d_n)
    ok = i_p >= 0
    assert np.array_equal(med[ok], desc[os_[:-1][ok] + i_p[ok]])
    assert not med[~ok].any()


@needs_ref
@pytest.mark.parametrize("off,kw", CASES)
@pytest.mark.parametrize("is_line", [False, True])
def test_port_vs_reference(off, kw, is_line):
    desc, dirs, os_ = synth.make_landmark_observations(synth.SEED0 + 20 + off, **kw)
    i_p, _, d_p = port.med_desc(desc, dirs, os_)
    i_r, d_r = ref.med_desc(desc, dirs, os_, is_line=is_line)
    assert np.array_equal(i_p, i_r)
    assert np.array_equal(d_p.view(np.uint64), d_r.view(np.uint64))  # bit for bit, signed zeros included


def test_port_vs_golden():
    z = np.load(GOLDEN)
    for k in range(int(z["n_cases"])):
        i_p, med, d_p = port.med_desc(z[f"desc_{k}"], z[f"dirs_{k}"], z[f"obs_start_{k}"])
        assert np.array_equal(i_p, z[f"med_idx_{k}"])
        assert np.array_equal(d_p.view(np.uint64), z[f"med_dir_{k}"].view(np.uint64))


def test_host_validation_needs_no_gpu(plm_lib):
    import ctypes as C
    from pl_inertial_slam_b200 import _lib as L
    d = np.zeros((4, 32), np.uint8)
    idx = np.zeros(2, np.int32)
    bad = np.array([0, 3, 2], np.int32)     # decreasing
    far = np.array([0, 2, 9], np.int32)     # past n_obs
    nul = C.cast(None, L.f64p)
    call = lambda os_: plm_lib.plm_med_desc(None, d.ctypes.data_as(L.u8p), 4, 32, nul, os_.ctypes.data_as(L.i32p), 2,  # noqa: E731
                                            idx.ctypes.data_as(L.i32p), C.cast(None, L.u8p), nul)
    assert call(bad) == L.PLM_E_INVALID
    assert call(far) == L.PLM_E_INVALID
    assert plm_lib.plm_med_desc(None, d.ctypes.data_as(L.u8p), 4, 16, nul, far.ctypes.data_as(L.i32p), 2,
                                idx.ctypes.data_as(L.i32p), C.cast(None, L.u8p), nul) == L.PLM_E_INVALID


# ---- GPU parity ------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("off,kw", CASES + [(6, dict(n_lm=20000, mean_obs=8)),
                                            (7, dict(n_lm=5000, mean_obs=10, long_lists=40, long_len=45, empty_frac=0.02))])
def test_gpu_med_desc_host_call(off, kw):
    from pl_inertial_slam_b200 import mapfeatures as MF
    desc, dirs, os_ = synth.make_landmark_observations(synth.SEED0 + 20 + off, **kw)
    i_p, med_p, d_p = port.med_desc(desc, dirs, os_)
    i_g, med_g, d_g = MF.med_desc_batch(desc, os_, dirs)
    assert np.array_equal(i_g, i_p)
    assert np.array_equal(med_g, med_p)
    assert np.array_equal(d_g.view(np.uint64), d_p.view(np.uint64))
    # without directions / strided descriptor rows
    wide = np.zeros((len(desc), 48), np.uint8)
    wide[:, :32] = desc
    i_g2, med_g2, none = MF.med_desc_batch(wide[:, :32], os_)
    assert none is None and np.array_equal(i_g2, i_p) and np.array_equal(med_g2, med_p)


@pytest.mark.gpu
def test_gpu_med_desc_golden():
    from pl_inertial_slam_b200 import mapfeatures as MF
    z = np.load(GOLDEN)
    for k in range(int(z["n_cases"])):
        i_g, _, d_g = MF.med_desc_batch(z[f"desc_{k}"], z[f"obs_start_{k}"], z[f"dirs_{k}"])
        assert np.array_equal(i_g, z[f"med_idx_{k}"])
        assert np.array_equal(d_g.view(np.uint64), z[f"med_dir_{k}"].view(np.uint64))


@pytest.mark.gpu
def test_gpu_med_desc_device_scatter():
    """plm_dev_med_desc writing med_desc rows straight into a resident map shard (dst_rows scatter)."""
    import torch
    from pl_inertial_slam_b200 import mapfeatures as MF
    from pl_inertial_slam_b200 import matching as M
    desc, dirs, os_ = synth.make_landmark_observations(synth.SEED0 + 29, n_lm=3000, mean_obs=7, long_lists=5, long_len=40)
    n_lm = len(os_) - 1
    i_p, med_p, d_p = port.med_desc(desc, dirs, os_)
    ctx = M.Context(0)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(5)
    rows = rng.permutation(2 * n_lm)[:n_lm].astype(np.int32)
    rows[::17] = -1                                   # landmarks that are not part of the shard
    shard = torch.full((2 * n_lm, 32), 0xAB, dtype=torch.uint8, device=dev)
    t = lambda a: torch.from_numpy(a).to(dev)          # noqa: E731
    med_idx = torch.empty(n_lm, dtype=torch.int32, device=dev)
    med_dir = torch.empty((n_lm, 3), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    MF.dev_med_desc(ctx, t(desc), t(os_), med_idx, med_desc=shard, dir_obs=t(dirs), med_dir=med_dir, dst_rows=t(rows))
    ctx.synchronize()
    assert np.array_equal(med_idx.cpu().numpy(), i_p)
    assert np.array_equal(med_dir.cpu().numpy().view(np.uint64), d_p.view(np.uint64))
    got = shard.cpu().numpy()
    want = np.full((2 * n_lm, 32), 0xAB, np.uint8)
    want[rows[rows >= 0]] = med_p[rows >= 0]
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_map_landmark_classes():
    """The reference's usage: constructor + addMapPointObservation per observation; deferred adds + one batch."""
    from pl_inertial_slam_b200 import mapfeatures as MF
    desc, dirs, os_ = synth.make_landmark_observations(synth.SEED0 + 30, n_lm=25, mean_obs=5)
    i_p, med_p, d_p = port.med_desc(desc, dirs, os_)
    pts, lns = [], []
    for l in range(len(os_) - 1):
        lo, hi = int(os_[l]), int(os_[l + 1])
        p = MF.MapPoint(l, np.zeros(3), desc[lo], 0, np.zeros(2), dirs[lo])
        q = MF.MapLine(l, np.zeros(6), desc[lo], 0, np.zeros(3), dirs[lo], np.zeros(4))
        for i in range(lo + 1, hi):
            if l % 2:
                p.addMapPointObservation(desc[i], i - lo, np.zeros(2), dirs[i])
            else:
                p.addMapPointObservation(desc[i], i - lo, np.zeros(2), dirs[i], defer=True)
            q.addMapLineObservation(desc[i], i - lo, np.zeros(3), dirs[i], np.zeros(4), defer=True)
        pts.append(p); lns.append(q)
    MF.update_average_desc_dir([p for p in pts if len(p.desc_list) > 1])
    got = MF.update_average_desc_dir(lns)
    assert np.array_equal(got, i_p)
    for l, (p, q) in enumerate(zip(pts, lns)):
        for lm in (p, q):
            assert np.array_equal(lm.med_desc, med_p[l])
            assert np.array_equal(np.asarray(lm.med_obs_dir).view(np.uint64), d_p[l].view(np.uint64))


@pytest.mark.gpu
def test_gpu_med_desc_group_boundaries():
    """Lists of exactly 1, 2, 8, 9, 16, 17, 32 and 33 observations next to each other: every packing mode of the warp
    kernel (4 x 8, 2 x 16, 1 x 32 lanes), the hand-over to the CTA kernel at 33, its shared-memory cache limit at
    128 / 129, and identical descriptors (all medians equal -> first row)."""
    from pl_inertial_slam_b200 import mapfeatures as MF
    rng = np.random.default_rng(12)
    counts = np.array([1, 2, 8, 9, 16, 17, 32, 33, 8, 8, 8, 8, 16, 16, 3, 32, 128, 129, 5, 0, 7], np.int64)
    counts = np.concatenate([counts, rng.integers(0, 34, 200)])
    start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    n_obs = int(start[-1])
    owner = np.repeat(np.arange(len(counts)), counts)
    desc = synth.flip_bits(rng, synth.rand_desc(rng, len(counts))[owner], 0.05)
    same = np.nonzero(counts >= 3)[0][::5]
    for l in same:                                   # all observations identical
        desc[start[l]:start[l + 1]] = desc[start[l]]
    dirs = rng.normal(size=(n_obs, 3))
    i_p, med_p, d_p = port.med_desc(desc, dirs, start)
    i_g, med_g, d_g = MF.med_desc_batch(desc, start, dirs)
    assert np.array_equal(i_g, i_p)
    assert np.array_equal(med_g, med_p)
    assert np.array_equal(d_g.view(np.uint64), d_p.view(np.uint64))
    assert (i_p[same] == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("tie", [False, True])
def test_gpu_med_desc_every_length(tie):
    """One landmark of every list length 0 .. 140, shuffled (classes of the warp kernel: <= 8, <= 16, <= 32 lanes;
    CTA kernel: 2 / 4 lanes per row up to 128, one thread per row above), noisy copies and the tie-stress set
    (few distinct rows -> many equal medians, first row wins, mapFeatures.cpp:80-84)."""
    from pl_inertial_slam_b200 import mapfeatures as MF
    rng = np.random.default_rng(77 + tie)
    counts = rng.permutation(np.concatenate([np.arange(0, 141), np.arange(2, 34), np.arange(2, 34)])).astype(np.int64)
    start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    n_obs = int(start[-1])
    owner = np.repeat(np.arange(len(counts)), counts)
    if tie:
        desc = synth.tie_stress_desc(rng, n_obs)
    else:
        desc = synth.flip_bits(rng, synth.rand_desc(rng, len(counts))[owner], 0.07)
    dirs = rng.normal(size=(n_obs, 3))
    i_p, med_p, d_p = port.med_desc(desc, dirs, start)
    i_g, med_g, d_g = MF.med_desc_batch(desc, start, dirs)
    assert np.array_equal(i_g, i_p), np.flatnonzero(i_g != i_p)[:10]
    assert np.array_equal(med_g, med_p)
    assert np.array_equal(d_g.view(np.uint64), d_p.view(np.uint64))
