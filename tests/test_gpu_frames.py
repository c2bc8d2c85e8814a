"""Device-resident stereo-frame pipeline (plm_frames_*) against the oracle's restatement of
StereoFrame::matchStereoPoints / matchStereoLines + StVO::match on the compacted descriptors."""
import numpy as np
import pytest

import oracle
from pl_inertial_slam_b200 import _lib as L
from pl_inertial_slam_b200 import synth
from pl_inertial_slam_b200.frames import FrameConfig, FramePipeline, replay_frame_records

pytestmark = pytest.mark.gpu


def arenas_from_pairs(pairs):
    """Arenas + frame records for a list of StereoPair-like objects (left/right counts may differ)."""
    desc, kps, lns = [], [], []
    rec = np.zeros(len(pairs), L.FRAME_REC_DTYPE)
    nd = nk = nl = 0
    for f, sp in enumerate(pairs):
        for key, arr in (("desc_pl", sp.pdesc_l), ("desc_pr", sp.pdesc_r), ("desc_ll", sp.ldesc_l), ("desc_lr", sp.ldesc_r)):
            rec[key][f] = nd
            desc.append(arr.reshape(-1, 32))
            nd += len(arr)
        rec["kp_l"][f] = nk; kps.append(sp.kp_l.reshape(-1, 2)); nk += len(sp.kp_l)
        rec["kp_r"][f] = nk; kps.append(sp.kp_r.reshape(-1, 2)); nk += len(sp.kp_r)
        rec["ln_l"][f] = nl; lns.append(sp.ln_l.reshape(-1, 4)); nl += len(sp.ln_l)
        rec["ln_r"][f] = nl; lns.append(sp.ln_r.reshape(-1, 4)); nl += len(sp.ln_r)
        rec["n_pl"][f], rec["n_pr"][f] = len(sp.kp_l), len(sp.kp_r)
        rec["n_ll"][f], rec["n_lr"][f] = len(sp.ln_l), len(sp.ln_r)
    cat = lambda xs, w, dt: (np.ascontiguousarray(np.concatenate(xs), dt) if xs else np.zeros((0, w), dt))  # noqa: E731
    return cat(desc, 32, np.uint8), cat(kps, 2, np.float32), cat(lns, 4, np.float32), rec


def oracle_pipeline(desc, kp, ln, rec, cfg: FrameConfig):
    """Per frame: the oracle's stereo drivers, then StVO::match on the compacted descriptors."""
    port = oracle.port
    frames = []
    for r in rec:
        sl = lambda off, n, a: a[int(off):int(off) + int(n)]  # noqa: E731
        dpl, dpr = sl(r["desc_pl"], r["n_pl"], desc), sl(r["desc_pr"], r["n_pr"], desc)
        dll, dlr = sl(r["desc_ll"], r["n_ll"], desc), sl(r["desc_lr"], r["n_lr"], desc)
        kl, kr = sl(r["kp_l"], r["n_pl"], kp), sl(r["kp_r"], r["n_pr"], kp)
        ll, lr = sl(r["ln_l"], r["n_ll"], ln), sl(r["ln_r"], r["n_lr"], ln)
        p = port.stereo_points(kl, dpl, kr, dpr, cfg.inv_width, cfg.inv_height, cfg.cam, cfg.grid_rows, cfg.grid_cols,
                               cfg.matchingSWs, cfg.minRatio12P, cfg.bestLRMatches, cfg.maxDistEpip, cfg.minDisp)
        q = port.stereo_lines(ll, dll, lr, dlr, cfg.inv_width, cfg.inv_height, cfg.cam, cfg.grid_rows, cfg.grid_cols,
                              cfg.matchingSWs, cfg.minRatio12P, cfg.lineSimTh, cfg.bestLRMatches, cfg.minDisp,
                              cfg.lineHorizTh, cfg.stereoOverlapTh, cfg.lsMinDispRatio)
        p["cdesc"] = np.ascontiguousarray(dpl[p["kept_i1"]])
        q["cdesc"] = np.ascontiguousarray(dll[q["kept_i1"]])
        p["n_match"] = int((p["m12"] >= 0).sum())
        q["n_match"] = int((q["m12"] >= 0).sum())
        frames.append((p, q))
    f2f = []
    for f in range(1, len(frames)):
        row = []
        for k, nnr in ((0, cfg.minRatio12P), (1, cfg.minRatio12L)):
            a, b = frames[f - 1][k]["cdesc"], frames[f][k]["cdesc"]
            if len(a) == 0 or len(b) == 0:
                row.append((0, np.zeros(0, np.int32)))
            elif len(b) < 2 or (cfg.bestLRMatches and len(a) < 2):
                row.append((np.iinfo(np.int32).min, np.full(len(a), -1, np.int32)))
            else:
                n, m = port.match(a, b, np.float32(nnr), cfg.bestLRMatches)
                row.append((n, m))
        f2f.append(row)
    return frames, f2f


def check(desc, kp, ln, rec, cfg):
    pipe = FramePipeline()
    pipe.upload(desc, kp, ln, rec, cfg)
    pipe.run()
    out = pipe.fetch()
    frames, f2f = oracle_pipeline(desc, kp, ln, rec, cfg)
    for f, (p, q) in enumerate(frames):
        lo, hi = int(pipe.off_p[f]), int(pipe.off_p[f + 1])
        k = len(p["kept_i1"])
        assert out["counts"][f, 0] == p["n_match"] and out["counts"][f, 1] == k, (f, out["counts"][f], p["n_match"], k)
        assert np.array_equal(out["stereo_m12_p"][lo:hi], p["m12"]), f
        assert np.array_equal(out["kept_p"][lo:lo + k], p["kept_i1"]), f
        assert np.array_equal(out["pt_disp"][lo:lo + k], p["disp"]), f
        assert np.array_equal(out["pt_P"][lo:lo + k], p["P"]), f
        lo, hi = int(pipe.off_l[f]), int(pipe.off_l[f + 1])
        k = len(q["kept_i1"])
        assert out["counts"][f, 2] == q["n_match"] and out["counts"][f, 3] == k, (f, out["counts"][f], q["n_match"], k)
        assert np.array_equal(out["stereo_m12_l"][lo:hi], q["m12"]), f
        assert np.array_equal(out["kept_l"][lo:lo + k], q["kept_i1"]), f
        assert np.array_equal(out["ls_disp"][lo:lo + k], q["disp_se"]), f
        assert np.array_equal(out["ls_sP"][lo:lo + k], q["sP"]), f
        assert np.array_equal(out["ls_eP"][lo:lo + k], q["eP"]), f
        assert np.array_equal(out["ls_le"][lo:lo + k], q["le"], equal_nan=True), f
    assert (out["counts"][0, 4:] == 0).all()
    for f in range(1, len(frames)):
        for k, (name, off) in enumerate((("f2f_m12_p", pipe.off_p), ("f2f_m12_l", pipe.off_l))):
            n, m = f2f[f - 1][k]
            assert out["counts"][f, 4 + k] == n, (f, k, out["counts"][f], n)
            lo = int(off[f - 1])
            assert np.array_equal(out[name][lo:lo + len(m)], m), (f, k)
    pipe.close()
    return out


def test_replay_pipeline_matches_oracle():
    rp = synth.make_replay(synth.SEED0 + 3, 10, mean_pts=300, sd_pts=40, mean_lines=90, sd_lines=15)
    kp, ln, rec = replay_frame_records(rp)
    out = check(rp.arena, kp, ln, rec, FrameConfig())
    assert (out["counts"][:, 1] > 50).all() and (out["counts"][1:, 4] > 10).all()  # real stereo + temporal matches


@pytest.mark.parametrize("ratio,best_lr", [(0.75, True), (0.9, False), (1.0, True)])
def test_full_size_frames_and_config_variants(ratio, best_lr):
    rp = synth.make_replay(synth.SEED0 + 30, 4)
    kp, ln, rec = replay_frame_records(rp)
    cfg = FrameConfig(minRatio12P=ratio, minRatio12L=0.75, bestLRMatches=best_lr, matchingSWs=6)
    check(rp.arena, kp, ln, rec, cfg)


def test_ragged_and_empty_frames():
    """Left / right counts differ; frames with no right features, no lines, a single point."""
    rng = np.random.default_rng(synth.SEED0 + 31)
    pairs = []
    shapes = [(120, 40), (3, 2), (0, 30), (90, 0), (1, 1), (64, 33), (64, 33)]
    for i, (n_p, n_l) in enumerate(shapes):
        sp = synth.make_stereo_pair(synth.SEED0 + 100 + i, n_pts=max(n_p, 1), n_lines=max(n_l, 1))
        if n_p == 0:
            sp.kp_l, sp.kp_r, sp.pdesc_l, sp.pdesc_r = sp.kp_l[:0], sp.kp_r[:0], sp.pdesc_l[:0], sp.pdesc_r[:0]
        if n_l == 0:
            sp.ln_l, sp.ln_r, sp.ldesc_l, sp.ldesc_r = sp.ln_l[:0], sp.ln_r[:0], sp.ldesc_l[:0], sp.ldesc_r[:0]
        if i == 5:  # fewer right than left features
            k = 20
            sp.kp_r, sp.pdesc_r = sp.kp_r[:k].copy(), sp.pdesc_r[:k].copy()
            sp.ln_r, sp.ldesc_r = sp.ln_r[:9].copy(), sp.ldesc_r[:9].copy()
        if i == 6:  # no right features at all: the stereo drivers return early
            sp.kp_r, sp.pdesc_r = sp.kp_r[:0], sp.pdesc_r[:0]
        pairs.append(sp)
    # features on / beyond the image border (off-grid cells, truncation toward zero of negatives)
    sp = synth.make_stereo_pair(synth.SEED0 + 140, n_pts=80, n_lines=30)
    sp.kp_r[:10, 0] = rng.uniform(-9.0, 3.0, 10).astype(np.float32)
    sp.kp_l[:10, 1] = rng.uniform(478.0, 486.0, 10).astype(np.float32)
    sp.ln_r[:6, 0] = rng.uniform(-30.0, 5.0, 6).astype(np.float32)
    sp.ln_l[:6, 3] = rng.uniform(470.0, 500.0, 6).astype(np.float32)
    sp.ln_r[6] = sp.ln_r[7]  # duplicate segment
    sp.ln_r[8, 2:] = sp.ln_r[8, :2]  # zero-length right segment -> NaN direction
    pairs.append(sp)
    desc, kp, ln, rec = arenas_from_pairs(pairs)
    check(desc, kp, ln, rec, FrameConfig())


def test_tie_stress_descriptors():
    rng = np.random.default_rng(synth.SEED0 + 9)
    pairs = []
    for i in range(3):
        sp = synth.make_stereo_pair(synth.SEED0 + 150 + i, n_pts=200, n_lines=70)
        sp.pdesc_l, sp.pdesc_r = synth.tie_stress_desc(rng, 200), synth.tie_stress_desc(rng, 200)
        sp.ldesc_l, sp.ldesc_r = synth.tie_stress_desc(rng, 70), synth.tie_stress_desc(rng, 70)
        sp.kp_r[:, 1] = sp.kp_l[:, 1]  # everything passes the epipolar gate -> many kept rows with tied distances
        pairs.append(sp)
    desc, kp, ln, rec = arenas_from_pairs(pairs)
    check(desc, kp, ln, rec, FrameConfig(minRatio12P=1.0, minRatio12L=1.0))


def test_rejects_bad_input():
    pipe = FramePipeline()
    sp = synth.make_stereo_pair(synth.SEED0 + 160, n_pts=20, n_lines=5)
    desc, kp, ln, rec = arenas_from_pairs([sp])
    bad = rec.copy(); bad["n_pl"][0] = 10_000
    with pytest.raises(L.PlmError):
        pipe.upload(desc, kp, ln, bad, FrameConfig())
    far = ln.copy(); far[len(sp.ln_l), 2] = 1e7
    with pytest.raises(L.PlmError):
        pipe.upload(desc, kp, far, rec, FrameConfig())
    with pytest.raises(L.PlmError):
        pipe.upload(desc, kp, ln, rec, FrameConfig(minRatio12P=1.5))
    pipe.close()


def test_repeated_runs_are_identical():
    """The column top-2 merge uses shared-memory atomics under racing thresholds: results must not depend on timing."""
    rp = synth.make_replay(synth.SEED0 + 33, 40)
    kp, ln, rec = replay_frame_records(rp)
    cfg = FrameConfig()
    pipe = FramePipeline()
    pipe.upload(rp.arena, kp, ln, rec, cfg)
    pipe.run()
    first = pipe.fetch()
    frames, f2f = oracle_pipeline(rp.arena, kp, ln, rec, cfg)
    for f in range(1, len(frames)):
        assert first["counts"][f, 4] == f2f[f - 1][0][0] and first["counts"][f, 5] == f2f[f - 1][1][0], f
    for _ in range(25):
        pipe.run()
        again = pipe.fetch()
        for k in ("counts", "f2f_m12_p", "f2f_m12_l", "stereo_m12_p", "stereo_m12_l", "kept_p", "kept_l"):
            assert np.array_equal(first[k], again[k]), k
    pipe.close()


@pytest.mark.parametrize("pairs_p,pairs_l", [(0, 0), (1, 2), (8, 32), (64, 64)])
def test_pair_list_and_chunk_paths_agree(pairs_p, pairs_l):
    """The stereo matcher has two forms (pair list / chunk phases); both must give the oracle's result,
    also when only some frames overflow the pair list."""
    lib = L.load()
    try:
        L.check(lib.plm_set_option(b"frames_pairs_p", pairs_p), "opt")
        L.check(lib.plm_set_option(b"frames_pairs_l", pairs_l), "opt")
        rp = synth.make_replay(synth.SEED0 + 34, 6, mean_pts=400, mean_lines=120)
        kp, ln, rec = replay_frame_records(rp)
        # pile the features of two frames into a corner so that windows hold dozens of candidates
        for f in (2, 4):
            a, n = int(rec["kp_l"][f]), int(rec["n_pl"][f])
            kp[a:a + n] = kp[a:a + n] * np.float32(0.08) + np.float32(20.0)
            b = int(rec["kp_r"][f])
            kp[b:b + n] = kp[b:b + n] * np.float32(0.08) + np.float32(18.0)
            a, n = int(rec["ln_l"][f]), int(rec["n_ll"][f])
            ln[a:a + n] = ln[a:a + n] * np.float32(0.15) + np.float32(25.0)
            b = int(rec["ln_r"][f])
            ln[b:b + n] = ln[b:b + n] * np.float32(0.15) + np.float32(23.0)
        check(rp.arena, kp, ln, rec, FrameConfig())
    finally:
        lib.plm_set_option(b"frames_pairs_p", 8)
        lib.plm_set_option(b"frames_pairs_l", 0)


@pytest.mark.parametrize("chunk", [1, 3, 7, 256])
def test_pipelined_process_equals_staged_calls(chunk):
    """plm_frames_process (chunks over three streams) gives the same outputs as upload / run / fetch."""
    import torch
    rp = synth.make_replay(synth.SEED0 + 35, 20, mean_pts=250, mean_lines=80)
    kp, ln, rec = replay_frame_records(rp)
    cfg = FrameConfig()
    ref_out = check(rp.arena, kp, ln, rec, cfg)
    pipe = FramePipeline()
    pin = lambda a: torch.from_numpy(a).pin_memory()  # noqa: E731
    pipe.set_layout(rec)
    out = pipe.alloc_outputs(pinned=True)
    for _ in range(2):  # second pass reuses the buffers of the first
        for v in out.values():
            v.fill_(-7)
        pipe.process(pin(rp.arena), pin(kp), pin(ln), rec, cfg, out, chunk_frames=chunk)
        for name, want in ref_out.items():
            got = out[name].numpy()
            if name.startswith(("kept", "stereo_m12", "f2f", "counts")):
                assert np.array_equal(got, want), name
            else:  # double outputs: only the kept prefix of each frame is defined
                off = pipe.off_p if name.startswith("pt_") else pipe.off_l
                col = 1 if name.startswith("pt_") else 3
                for f in range(len(rec)):
                    k = int(ref_out["counts"][f, col])
                    lo = int(off[f])
                    assert np.array_equal(got[lo:lo + k], want[lo:lo + k], equal_nan=True), (name, f)
    assert pipe.d2h_bytes > 0 and pipe.h2d_bytes >= rp.arena.nbytes + kp.nbytes + ln.nbytes
    pipe.close()


def _golden_stereo():
    import os
    from conftest import ROOT
    return np.load(os.path.join(ROOT, "tests", "golden", "stereo.npz"))


def _golden_cfg(z, c):
    g = lambda k, d: float(z[f"cfg{c}_{k}"]) if f"cfg{c}_{k}" in z.files else d  # noqa: E731
    fx, _, cx, cy, b = z["cam_ref"]
    return FrameConfig(img_width=int(z["img_wh"][0]), img_height=int(z["img_wh"][1]), matchingSWs=int(g("matching_s_ws", 10)),
                       bestLRMatches=bool(g("best_lr", 1)), minRatio12P=g("ratio", 0.9), maxDistEpip=g("max_dist_epip", 1.0),
                       minDisp=g("min_disp", 1.0), lineHorizTh=g("line_horiz_th", 0.1), stereoOverlapTh=g("stereo_overlap_th", 0.75),
                       lsMinDispRatio=g("ls_min_disp_ratio", 0.7), cam_b=float(b), cam_fx=float(fx), cam_cx=float(cx), cam_cy=float(cy))


def test_frame_pipeline_vs_reference_golden():
    """tests/golden/stereo.npz holds the outputs of the reference's OWN StereoFrame::matchStereoPoints / matchStereoLines
    (stereoFrame.cpp compiled unmodified, tools/make_golden.py stereo): kept features, disparities, back-projected
    3-D points / segments and line equations must be bit-identical."""
    from types import SimpleNamespace
    z = _golden_stereo()
    pairs = [SimpleNamespace(**{k: z[f"f{f}_{k}"] for k in ("kp_l", "kp_r", "pdesc_l", "pdesc_r", "ln_l", "ln_r", "ldesc_l", "ldesc_r")})
             for f in range(int(z["n_frames"]))]
    desc, kp, ln, rec = arenas_from_pairs(pairs)
    bits = lambda a: np.ascontiguousarray(a, np.float64).view(np.uint64)  # noqa: E731
    for c in range(int(z["n_cfg"])):
        pipe = FramePipeline()
        pipe.upload(desc, kp, ln, rec, _golden_cfg(z, c))
        pipe.run()
        out = pipe.fetch()
        for f in range(len(pairs)):
            lo = int(pipe.off_p[f])
            want = z[f"f{f}_c{c}_pt_kept_i1"]
            k = len(want)
            assert out["counts"][f, 1] == k and np.array_equal(out["kept_p"][lo:lo + k], want), (c, f)
            assert np.array_equal(bits(out["pt_disp"][lo:lo + k]), bits(z[f"f{f}_c{c}_pt_disp"])), (c, f)
            assert np.array_equal(bits(out["pt_P"][lo:lo + k]), bits(z[f"f{f}_c{c}_pt_P"])), (c, f)
            lo = int(pipe.off_l[f])
            want = z[f"f{f}_c{c}_ls_kept_i1"]
            k = len(want)
            assert out["counts"][f, 3] == k and np.array_equal(out["kept_l"][lo:lo + k], want), (c, f)
            for name, key in (("ls_disp", "disp_se"), ("ls_sP", "sP"), ("ls_eP", "eP"), ("ls_le", "le")):
                assert np.array_equal(bits(out[name][lo:lo + k]), bits(z[f"f{f}_c{c}_ls_{key}"])), (c, f, name)
        pipe.close()


def test_line_overlap_kernel_vs_reference_golden():
    """StereoFrame::lineSegmentOverlap (stereoFrame.cpp:521-627) outputs of the compiled reference against
    line_pair_filter_kernel: segment i of set 1 is the projection, segment i of set 2 the observation."""
    from pl_inertial_slam_b200 import matching as M
    z = _golden_stereo()
    v = z["ov_in"]
    n = len(v)
    # inputs are float32 in the ABI (KeyLine fields): round them first and take the reference of the rounded values
    # from the restatement, which tests/test_ref_stereo.py pins on the compiled reference for arbitrary doubles
    obs = v[:, 0:4].astype(np.float32)
    proj = v[:, 4:8].astype(np.float32)
    m12 = np.arange(n, dtype=np.int32)
    n_g, keep_g, ov_g, sim_g = M.line_pair_filter(proj, obs, m12, 0.75, 0.75)
    n_o, keep_o, ov_o, sim_o = oracle.port.line_pair_filter(proj, obs, m12, 0.75, 0.75)
    assert n_g == n_o and np.array_equal(keep_g, keep_o)
    assert np.array_equal(ov_g.view(np.uint64), ov_o.view(np.uint64))
