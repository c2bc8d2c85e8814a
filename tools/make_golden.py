"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libplref.so = the reference's
matching.cpp / gridStructure.cpp / lineIterator.cpp / mapFeatures.cpp compiled unmodified, see oracle/Makefile).

Run in the build container (needs /root/reference to build libplref.so):
    python tools/make_golden.py [brute] [grid] [line_coords] [med_desc] [bow] [stereo]
The fixtures are small (a few hundred KB) and committed; tests compare the oracle port and the CUDA
path against them, so parity stays pinned on machines where the reference cannot be built.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from helpers import random_grid_case  # noqa: E402
from pl_inertial_slam_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ref = oracle.ref
assert ref.available(), "build oracle/_ref/libplref.so first (make -C oracle ref)"


def save(name, **arrays):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote", name, {k: getattr(v, "shape", v) for k, v in arrays.items()})


def brute_cases():
    rng = np.random.default_rng(synth.SEED0 + 9)
    shapes = [(2, 2, 1), (50, 2, 1), (70, 3, 1), (100, 33, 1), (64, 1025, 1), (200, 200, 0), (301, 277, 0)]
    for ci, (n1, n2, tie) in enumerate(shapes):
        d2 = synth.tie_stress_desc(rng, n2) if tie else synth.rand_desc(rng, n2)
        if tie:
            d1 = synth.tie_stress_desc(rng, n1)
        else:
            d1 = synth.rand_desc(rng, n1)
            k = min(n1, n2) * 2 // 3
            d1[rng.choice(n1, k, replace=False)] = synth.flip_bits(rng, d2[rng.choice(n2, k, replace=False)], 0.08)
        stale = np.full(n1, -1, np.int32)
        if n1 > 10:
            pos = rng.choice(n1, n1 // 4, replace=False)
            stale[pos] = rng.integers(0, n2, len(pos))
        out = dict(d1=d1, d2=d2, stale=stale)
        for nnr in (0.75, 0.9):
            for blr in (0, 1):
                n, m = ref.match(d1, d2, nnr, best_lr=blr)
                out[f"m_{nnr}_{blr}"] = m
                out[f"n_{nnr}_{blr}"] = np.int32(n)
                n, m = ref.match(d1, d2, nnr, best_lr=blr, m12=stale)
                out[f"ms_{nnr}_{blr}"] = m
                out[f"ns_{nnr}_{blr}"] = np.int32(n)
        save(f"brute_{ci}", **out)


def grid_cases():
    specs = [
        (300, 300, False, False, (10, 0, 0, 0), 0, 0),
        (250, 200, False, True, (3, 3, 3, 3), 6, 0),
        (40, 6, False, True, (70, 70, 50, 50), 0, 0),
        (120, 120, True, False, (10, 0, 0, 0), 0, 3),
        (150, 90, True, True, (3, 3, 3, 3), 0, 5),
        (1500, 200, False, False, (3, 3, 3, 3), 0, 0),
    ]
    for ci, (n1, n2, is_lines, tie, win, bad, zl) in enumerate(specs):
        rng = np.random.default_rng(synth.SEED0 + 100 + ci)
        case = random_grid_case(rng, n1, n2, is_lines=is_lines, tie=tie, win=win, bad_items=bad, zero_len=zl)
        out = {k: v for k, v in case.items() if v is not None}
        out["is_lines"] = np.int32(is_lines)
        for ratio in (0.75, 0.9, 1.0):
            for blr in (0, 1):
                if is_lines:
                    n, m = ref.match_grid_lines(case["coords"], case["d1"], case["cell_start"], case["cell_items"],
                                                case["rows"], case["cols"], case["d2"], case["dirs2"], 0.75,
                                                case["win"], ratio, blr)
                else:
                    n, m = ref.match_grid_points(case["coords"], case["d1"], case["cell_start"], case["cell_items"],
                                                 case["rows"], case["cols"], case["d2"], case["win"], ratio, blr)
                out[f"m_{ratio}_{blr}"] = m
                out[f"n_{ratio}_{blr}"] = np.int32(n)
        save(f"grid_{ci}", **out)


def line_coord_cases():
    rng = np.random.default_rng(synth.SEED0 + 200)
    seg = rng.uniform(-4, 70, (200, 4))
    seg[:10, 2:] = seg[:10, :2]            # zero-length
    seg[10:20, 1] = seg[10:20, 3]          # horizontal
    seg[20:30, 0] = seg[20:30, 2]          # vertical
    flat, offs = [], [0]
    for s in seg:
        c = ref.line_coords(*s)
        flat.append(c)
        offs.append(offs[-1] + len(c))
    save("line_coords", seg=seg, cells=np.concatenate(flat), offs=np.array(offs, np.int64))


def med_desc_cases():
    """MapPoint / MapLine::updateAverageDescDir from the reference's own src/mapFeatures.cpp (points and
    lines share one body; both are run and must agree)."""
    specs = [dict(n_lm=150, mean_obs=6), dict(n_lm=100, mean_obs=10, tie=True),
             dict(n_lm=40, mean_obs=4, long_lists=4, long_len=50, empty_frac=0.1),
             dict(n_lm=12, mean_obs=3, long_lists=2, long_len=90, tie=True)]
    out = {"n_cases": np.int32(len(specs))}
    for k, kw in enumerate(specs):
        desc, dirs, obs_start = synth.make_landmark_observations(synth.SEED0 + 300 + k, **kw)
        i_pt, d_pt = ref.med_desc(desc, dirs, obs_start, is_line=False)
        i_ln, d_ln = ref.med_desc(desc, dirs, obs_start, is_line=True)
        assert np.array_equal(i_pt, i_ln) and np.array_equal(d_pt.view(np.uint64), d_ln.view(np.uint64))
        out.update({f"desc_{k}": desc, f"dirs_{k}": dirs, f"obs_start_{k}": obs_start, f"med_idx_{k}": i_pt,
                    f"med_dir_{k}": d_pt})
    save("med_desc", **out)


def bow_cases():
    """DBoW2 from the reference's vendored sources (oracle/_ref/libplref_dbow.so): vocabularies built by its own
    create() (hierarchical k-means++ on synthetic training descriptors; seeds pinned), then transform() of a
    few descriptor sets and the full score() matrix between them, for every weighting type."""
    rdb = oracle.ref_dbow
    assert rdb.available(), "build oracle/_ref/libplref_dbow.so first (make -C oracle ref)"
    rng = np.random.default_rng(synth.SEED0 + 400)
    centres = synth.rand_desc(rng, 300)
    train = [synth.flip_bits(rng, centres[rng.integers(0, 300, 150)], 0.15) for _ in range(30)]
    tdesc, tstart = np.concatenate(train), np.arange(31, dtype=np.int32) * 150
    sets = [synth.flip_bits(rng, centres[rng.integers(0, 300, n)], 0.15) for n in (200, 180, 260, 90, 33, 1)]
    sets += [train[0], synth.rand_desc(rng, 150), synth.tie_stress_desc(rng, 64), np.zeros((0, 32), np.uint8)]
    out = {"n_sets": np.int32(len(sets)), "n_voc": np.int32(4)}
    for k, d in enumerate(sets):
        out[f"set_{k}"] = d
    for weighting in range(4):
        h = rdb.create(tdesc, tstart, k=8, L=3, weighting=weighting, seed=11 + weighting)
        fv = rdb.export(h, 8, 3, weighting)
        for name, arr in fv.arrays().items():
            out[f"voc{weighting}_{name}"] = arr
        bows = [rdb.transform(h, d) for d in sets]
        for k, (ids, vals) in enumerate(bows):
            out[f"voc{weighting}_ids_{k}"] = ids
            out[f"voc{weighting}_vals_{k}"] = vals
        out[f"voc{weighting}_scores"] = np.array([[rdb.score(h, a, b) for b in bows] for a in bows])
        rdb.destroy(h)
    save("bow", **out)


def stereo_cases():
    """StereoFrame::matchStereoPoints / matchStereoLines + the scalar gates from the reference's own stereoFrame.cpp /
    stereoFeatures.cpp / pinholeStereoCamera.cpp (oracle/_ref/libplref_stereo.so), on frames salted with the
    degenerate inputs the gates special-case (see tests/test_ref_stereo.py::degenerate_pair)."""
    from test_ref_stereo import CAM_REF, H, W, degenerate_pair
    rs = oracle.ref_stereo
    assert rs.available(), "build oracle/_ref/libplref_stereo.so first (make -C oracle ref)"
    cfgs = [dict(ratio=0.9, best_lr=True, matching_s_ws=10), dict(ratio=0.75, best_lr=False, matching_s_ws=10),
            dict(ratio=0.9, best_lr=True, matching_s_ws=4, max_dist_epip=1.5, min_disp=5.0, line_horiz_th=1.0,
                 stereo_overlap_th=0.6, ls_min_disp_ratio=0.8)]
    out = {"n_frames": np.int32(3), "n_cfg": np.int32(len(cfgs)), "cam_ref": CAM_REF, "img_wh": np.array([W, H], np.int32)}
    for f in range(3):
        names = ("kp_l", "pdesc_l", "kp_r", "pdesc_r", "ln_l", "ldesc_l", "ln_r", "ldesc_r")
        data = dict(zip(names, degenerate_pair(40 + f)))
        for k, v in data.items():
            out[f"f{f}_{k}"] = v
        for c, cfg in enumerate(cfgs):
            r = rs.stereo_points(data["kp_l"], data["pdesc_l"], data["kp_r"], data["pdesc_r"], W, H, CAM_REF, **cfg)
            q = rs.stereo_lines(data["ln_l"], data["ldesc_l"], data["ln_r"], data["ldesc_r"], W, H, CAM_REF, **cfg)
            for k in ("kept_i1", "disp", "P"):
                out[f"f{f}_c{c}_pt_{k}"] = r[k]
            for k in ("kept_i1", "disp_se", "sP", "eP", "le"):
                out[f"f{f}_c{c}_ls_{k}"] = q[k]
    for c, cfg in enumerate(cfgs):
        for k, v in cfg.items():
            out[f"cfg{c}_{k}"] = np.float64(v)
    rng = np.random.default_rng(synth.SEED0 + 500)
    v8 = rng.uniform(0, 700, (600, 8))
    v8[::5, 2] = v8[::5, 0] + rng.uniform(-1.5, 1.5, 120)
    v8[1::5, 3] = v8[1::5, 1] + rng.uniform(-1.5, 1.5, 120)
    v8[2::31, 2:4] = v8[2::31, 0:2]
    v4 = rng.uniform(0, 480, (600, 4))
    v4[::7, 1] = v4[::7, 0] + rng.uniform(-0.2, 0.2, len(v4[::7]))
    out.update(ov_in=v8, ov_out=rs.line_overlap(v8), ovs_in=v4, ovs_out=rs.line_overlap_stereo(v4, 0.1))
    save("stereo", **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, fn in (("brute", brute_cases), ("grid", grid_cases), ("line_coords", line_coord_cases),
                     ("med_desc", med_desc_cases), ("bow", bow_cases), ("stereo", stereo_cases)):
        if not only or name in only:
            fn()
