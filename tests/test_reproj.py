"""Local-map selection and reprojection gates of MapHandler::matchMap2KFPoints / Lines (src/mapHandler.cpp:583-682,
:685-803).  mapHandler.cpp cannot be compiled here (g2o, Eigen, OpenCV) and the summation order of Eigen's 3x3 * 3x1
product is not reproducible without Eigen: the restatement (oracle/plm_oracle.c plo_map_*) is PARITY UNPINNED and is
cross-checked here against an independent numpy form; the CUDA kernels are then compared with it bit for bit
(tests/test_gpu_reproj.py)."""
import numpy as np

import oracle

port = oracle.port
W, H = 752, 480
CAM = (435.2, 435.2, 367.4, 252.2)
INV_W, INV_H = 64.0 / W, 48.0 / H


def make_scene(seed, n, lines=False):
    rng = np.random.default_rng(seed)
    ang = rng.normal(0, 0.05, 3)
    cx, cy, cz = np.cos(ang)
    sx, sy, sz = np.sin(ang)
    R = np.array([[cy * cz, -cy * sz, sy], [sx * sy * cz + cx * sz, -sx * sy * sz + cx * cz, -sx * cy],
                  [-cx * sy * cz + sx * sz, cx * sy * sz + sx * cz, cx * cy]])
    T = np.concatenate([R, rng.normal(0, 0.2, (3, 1))], 1)
    X = np.concatenate([rng.uniform(-6, 6, (n, 2)), rng.uniform(-2, 12, (n, 1))], 1)      # some behind the camera
    if lines:
        X = np.concatenate([X, X + rng.normal(0, 0.4, (n, 3))], 1)
    active = (rng.random(n) < 0.8).astype(np.uint8)
    return T, X, active


def numpy_select(X, active, T, lines):
    per = 2 if lines else 1
    keep = active.astype(bool).copy()
    uvs = []
    for k in range(per):
        P = X[:, 3 * k:3 * k + 3] @ T[:, :3].T + T[:, 3]
        with np.errstate(all="ignore"):
            uv = np.stack([CAM[2] + CAM[0] * P[:, 0] / P[:, 2], CAM[3] + CAM[1] * P[:, 1] / P[:, 2]], 1)
        keep &= (uv[:, 0] > 0) & (uv[:, 0] < W) & (uv[:, 1] > 0) & (uv[:, 1] < H) & (P[:, 2] > 0)
        uvs.append(uv)
    uv = np.concatenate(uvs, 1)
    sel = np.flatnonzero(keep)
    pf = uv[sel]
    coords = np.trunc(pf * np.tile([INV_W, INV_H], per)).astype(np.int32)
    return sel.astype(np.int32), coords, pf


def test_select_port_vs_numpy():
    for lines in (False, True):
        for seed in range(4):
            T, X, active = make_scene(seed, 5000, lines)
            sel, coords, pf = port.map_select(X, active, T, CAM, INV_W, INV_H, W, H)
            sel2, coords2, pf2 = numpy_select(X, active, T, lines)
            assert 500 < len(sel) < 4500
            # numpy's matmul may sum in another order: identical selections away from the image border, fp64 within a few ulp
            assert np.array_equal(sel, sel2) and np.array_equal(coords, coords2)
            assert np.allclose(pf, pf2, rtol=1e-13, atol=1e-10)
    sel, _, _ = port.map_select(X, None, T, CAM, INV_W, INV_H, W, H)
    assert len(sel) > len(sel2)                                                      # no mask = all landmarks


def test_gate_port_vs_numpy():
    rng = np.random.default_rng(3)
    n_sel, n2 = 3000, 500
    pf = rng.uniform(0, 700, (n_sel, 2))
    pl = rng.uniform(0, 700, (n2, 2))
    m12 = rng.integers(-1, n2, n_sel).astype(np.int32)
    close = rng.random(n_sel) < 0.5
    idx = np.flatnonzero(close & (m12 >= 0))
    pf[idx] = pl[m12[idx]] + rng.normal(0, 0.6, (len(idx), 2))
    cnt, ok = port.map_gate(pf, m12, pl, 1.0, int((m12 >= 0).sum()))
    d = np.hypot(pf[:, 0] - pl[m12.clip(0), 0], pf[:, 1] - pl[m12.clip(0), 1])
    want = (m12 >= 0) & (d < 1.0)
    assert np.array_equal(ok.astype(bool), want) and cnt == int(want.sum()) and 300 < cnt < 2500
    # lines: signed point-line errors, both endpoints (mapHandler.cpp:784-786 -- no fabs in the reference)
    pf4 = rng.uniform(0, 700, (n_sel, 4))
    le = rng.normal(0, 1, (n2, 3))
    le /= np.hypot(le[:, 0], le[:, 1])[:, None]
    le[:, 2] = -(le[:, 0] * 350 + le[:, 1] * 240) + rng.normal(0, 30, n2)
    cnt, ok = port.map_gate(pf4, m12, le, 1.0, int((m12 >= 0).sum()))
    l = le[m12.clip(0)]
    e0 = l[:, 0] * pf4[:, 0] + l[:, 1] * pf4[:, 1] + l[:, 2]
    e1 = l[:, 0] * pf4[:, 2] + l[:, 1] * pf4[:, 3] + l[:, 2]
    want = (m12 >= 0) & (e0 < 1.0) & (e1 < 1.0)
    assert np.array_equal(ok.astype(bool), want) and cnt == int(want.sum()) and cnt > 100
