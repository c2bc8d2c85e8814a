"""The C++ drop-in for stvo-pl/src/matching.cpp (pl_inertial_slam_b200/csrc/stvo_matching_gpu.cpp):
StVO::matchNNR / match / distance / matchGrid x2 with the reference's exact signatures, built against
the reference's own headers (oracle/Makefile target stvo_gpu) and driven through std::vector<int>& /
cv::Mat / GridStructure arguments by the extern "C" harness.  Compared with the oracle port."""
import os

import numpy as np
import pytest

import oracle
from helpers import oracle_grid, random_grid_case
from pl_inertial_slam_b200 import synth

pytestmark = pytest.mark.gpu
port = oracle.port
HAVE = os.path.exists(os.path.join(os.path.dirname(oracle.__file__), "_ref", "libstvo_gpu.so"))
needs_lib = pytest.mark.skipif(not HAVE, reason="oracle/_ref/libstvo_gpu.so not built (needs the reference headers)")


@needs_lib
def test_stvo_match_and_distance(plm_lib):
    gpu = oracle.stvo_gpu
    rng = np.random.default_rng(1)
    for n1, n2, tie in [(2, 2, True), (300, 280, False), (64, 1025, True)]:
        d1 = synth.tie_stress_desc(rng, n1) if tie else synth.rand_desc(rng, n1)
        d2 = synth.tie_stress_desc(rng, n2) if tie else synth.rand_desc(rng, n2)
        if not tie:
            d1[:100] = synth.flip_bits(rng, d2[:100], 0.08)
        for blr in (0, 1):
            for nnr in (0.75, 0.9):
                n_o, m_o = port.match(d1, d2, nnr, blr)
                n_g, m_g = gpu.match(d1, d2, nnr, best_lr=blr)
                assert n_g == n_o and (m_g == m_o).all()
        n_o, m_o = port.match_nnr(d1, d2, 0.9)
        n_g, m_g = gpu.match_nnr(d1, d2, 0.9)
        assert n_g == n_o and (m_g == m_o).all()
        assert gpu.distance(d1[0], d2[0]) == port.distance(d1[0], d2[0])


@needs_lib
def test_stvo_match_throws_like_the_reference(plm_lib):
    gpu = oracle.stvo_gpu
    d = synth.rand_desc(np.random.default_rng(4), 4)
    with pytest.raises(RuntimeError):
        gpu.match_nnr(d, d[:0], 0.9)  # empty train set: "[matchNNR] Different size for matches and descriptors!"
    with pytest.raises(RuntimeError):
        gpu.match(d, d[:0], 0.9, best_lr=1)


@needs_lib
def test_stvo_match_sparse_frames_do_not_throw(plm_lib):
    """A frame with a single stereo point / line on one side (low-texture scenes) reaches StVO::match through
    matchF2FPoints/Lines and mapHandler.cpp:328/475/648/763/3332 without any try/catch: the reference carries on
    there (it reads matches_[idx][1] out of bounds, matching.cpp:54, but does not throw), so the drop-in must not
    terminate the process either.  A row without a second neighbour is not accepted."""
    gpu = oracle.stvo_gpu
    d = synth.rand_desc(np.random.default_rng(4), 4)
    n, m = gpu.match_nnr(d, d[:1], 0.9)
    assert n == 0 and (m == -1).all()
    n, m = gpu.match(d[:0], d, 0.9, best_lr=1)
    assert n == 0 and len(m) == 0
    # one query row: accepted in the 12 direction (it IS train row 0), no reverse match exists -> culled
    n, m = gpu.match(d[:1], d, 0.9, best_lr=1)
    assert n == 0 and (m == -1).all()
    n, m = gpu.match(d[:1], d, 0.9, best_lr=0)
    assert n == 1 and m[0] == 0
    # one train row + stale entries from a preceding matchGrid: nothing new is accepted, the mutual check culls
    stale = np.array([2, -1, 0, 1], np.int32)
    n, m = gpu.match(d, d[:1], 0.9, best_lr=1, m12=stale)
    assert n == -3 and (m == -1).all()
    stale = np.array([0, -1, -1, -1], np.int32)       # row 0 <-> train 0 is mutual: survives
    n, m = gpu.match(d, d[:1], 0.9, best_lr=1, m12=stale)
    assert n == 0 and m[0] == 0 and (m[1:] == -1).all()


@needs_lib
@pytest.mark.parametrize("is_lines", [False, True])
def test_stvo_match_grid(plm_lib, is_lines):
    gpu = oracle.stvo_gpu
    rng = np.random.default_rng(2 + int(is_lines))
    for n1, n2, tie, win in [(400, 350, False, (10, 0, 0, 0)), (200, 260, True, (3, 3, 3, 3)), (6000, 300, False, (3, 3, 3, 3))]:
        case = random_grid_case(rng, n1, n2, is_lines=is_lines, tie=tie, win=win, zero_len=3 if is_lines else 0)
        for ratio in (0.75, 0.9):
            for blr in (0, 1):
                n_o, m_o = oracle_grid(port, case, ratio, blr)
                if is_lines:
                    n_g, m_g = gpu.match_grid_lines(case["coords"], case["d1"], case["cell_start"], case["cell_items"],
                                                    case["rows"], case["cols"], case["d2"], case["dirs2"], 0.75,
                                                    case["win"], ratio, blr)
                else:
                    n_g, m_g = gpu.match_grid_points(case["coords"], case["d1"], case["cell_start"], case["cell_items"],
                                                     case["rows"], case["cols"], case["d2"], case["win"], ratio, blr)
                assert n_g == n_o and (m_g == m_o).all()


@needs_lib
def test_stvo_gpu_frame_one_launch(plm_lib):
    """StVO::GpuFrame (csrc/stvo_gpu_frame.h): the four matcher calls of one frame recorded with the reference's own
    argument types and executed as one launch -- results equal the oracle call by call, stale in/out entries included."""
    from pl_inertial_slam_b200 import synth as S
    gpu = oracle.stvo_gpu
    for seed, n_pts, n_lines in ((0, 600, 200), (1, 150, 40), (2, 5, 3)):
        prev, curr = S.make_temporal_pair(S.SEED0 + 60 + seed, n_pts=n_pts, n_lines=n_lines)
        a, b = S.stereo_points_grid_args(curr), S.stereo_lines_grid_args(curr)
        rng = np.random.default_rng(seed)
        stale = [np.full(n, -1, np.int32) for n in (n_pts, n_lines, len(prev.pdesc_l), len(prev.ldesc_l))]
        for k, n2 in enumerate((n_pts, n_lines, len(curr.pdesc_l), len(curr.ldesc_l))):
            stale[k][::6] = rng.integers(0, n2, len(stale[k][::6]))
        for blr, ratio in ((1, 0.9), (0, 0.75)):
            gpu.set_config(bool(blr), True, ratio, 0.75)
            counts, m, _ = gpu.gpu_frame((a["xy"], a["d1"], a["cell_start"], a["cell_items"], a["d2"], a["win"]),
                                         (b["xyxy"], b["d1"], b["cell_start"], b["cell_items"], b["d2"], b["dirs2"], b["win"]),
                                         (prev.pdesc_l, curr.pdesc_l), (prev.ldesc_l, curr.ldesc_l), a["rows"], a["cols"], 0.9, stale=stale)
            want = [port.match_grid_points(a["xy"], a["d1"], a["cell_start"], a["cell_items"], a["rows"], a["cols"], a["d2"], a["win"], ratio, blr, stale[0]),
                    port.match_grid_lines(b["xyxy"], b["d1"], b["cell_start"], b["cell_items"], b["rows"], b["cols"], b["d2"], b["dirs2"], 0.75,
                                          b["win"], ratio, blr, stale[1]),
                    port.match(prev.pdesc_l, curr.pdesc_l, 0.9, blr, m12=stale[2]),
                    port.match(prev.ldesc_l, curr.ldesc_l, 0.9, blr, m12=stale[3])]
            for k, (n_o, m_o) in enumerate(want):
                assert int(counts[k]) == n_o and (m[k] == m_o).all(), (seed, blr, k)
    gpu.set_config(True, True, 0.9, 0.75)


needs_map_lib = pytest.mark.skipif(not oracle.map_gpu.available(),
                                   reason="oracle/_ref/libmapfeatures_gpu.so not built (needs the reference headers)")


@needs_map_lib
@pytest.mark.parametrize("is_line", [False, True])
def test_map_features_dropin(plm_lib, is_line):
    """PLSLAM::MapPoint / MapLine from pl_inertial_slam_b200/csrc/map_features_gpu.cpp, built against the
    reference's include/mapFeatures.h: observation-by-observation use (one plm_med_desc per append, as the
    reference recomputes) and the batch entry point, against the restatement."""
    for seed, kw in [(1, dict(n_lm=40, mean_obs=5)), (2, dict(n_lm=25, mean_obs=9, tie=True, empty_frac=0.1)),
                     (3, dict(n_lm=6, mean_obs=3, long_lists=2, long_len=40))]:
        desc, dirs, start = synth.make_landmark_observations(synth.SEED0 + 500 + seed, **kw)
        i_p, _, d_p = port.med_desc(desc, dirs, start)
        for batch in (False, True):
            i_g, d_g = oracle.map_gpu.med_desc(desc, dirs, start, is_line=is_line, batch=batch)
            assert np.array_equal(i_g, i_p), (seed, batch)
            assert np.array_equal(d_g.view(np.uint64), d_p.view(np.uint64)), (seed, batch)


needs_dbow_lib = pytest.mark.skipif(not oracle.dbow_gpu.available(),
                                    reason="oracle/_ref/libdbow_gpu.so not built (needs the reference's DBoW2 headers)")


@needs_dbow_lib
@pytest.mark.parametrize("w", [0, 1, 2, 3])
def test_dbow_vocabulary_dropin(plm_lib, w):
    """PLM::GpuVocabulary (csrc/dbow_vocabulary_gpu.h), a subclass of the reference's own DBoW2 vocabulary type:
    transform() / score() / scoreAll() through DBoW2's BowVector interface against the restatement."""
    gpu = oracle.dbow_gpu
    fv = synth.make_vocabulary(synth.SEED0 + 600 + w, k=7, L=3, weighting=w)
    h = gpu.from_flat(fv)
    try:
        sets = [synth.vocabulary_features(700 + i, fv, n) for i, n in enumerate((250, 120, 1, 0, 400))]
        bows = []
        for d in sets:
            ids, vals = gpu.transform(h, d)
            wi, wv = port.bow_transform(fv, d)
            assert np.array_equal(ids, wi) and np.array_equal(vals.view(np.uint64), wv.view(np.uint64))
            bows.append((ids, vals))
        for a in bows[:3]:
            for b in bows:
                assert np.float64(gpu.score(h, a, b)).view(np.uint64) == np.float64(port.bow_score(a, b)).view(np.uint64)
        got = gpu.score_all(h, bows[0], bows)
        want = np.array([port.bow_score(bows[0], b) for b in bows])
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    finally:
        gpu.destroy(h)
