"""Parity of the CUDA path (through the C ABI) against the oracle on identical seeded inputs.

Bar: bit-exact match index vectors, counts and packed top-2 keys (integer work); the two ratio tests
are evaluated in the reference's precisions (fp32 for matchNNR, fp64 for matchGrid) and therefore
also compare exactly.
"""
import numpy as np
import pytest

import oracle
from helpers import gpu_grid, oracle_grid, random_grid_case
from pl_inertial_slam_b200 import synth

pytestmark = pytest.mark.gpu
port = oracle.port


@pytest.fixture(scope="module")
def M(plm_lib):
    from pl_inertial_slam_b200 import matching
    return matching


def test_distance(M):
    rng = np.random.default_rng(11)
    a, b = synth.rand_desc(rng, 300), synth.rand_desc(rng, 300)
    got = M.distances(a, b)
    want = np.array([port.distance(a[i], b[i]) for i in range(300)])
    assert (got == want).all()
    assert M.distance(a[0], a[0]) == 0
    assert M.distance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


@pytest.mark.parametrize("n1,n2,tie", [(1, 2, True), (50, 2, True), (70, 3, True), (100, 33, True), (64, 1025, True),
                                        (600, 600, False), (200, 200, False), (813, 1291, False),
                                        (300, 5000, True), (4100, 700, False)])
def test_knn2_packed_keys(M, n1, n2, tie):
    rng = np.random.default_rng(1000 + n1 + n2)
    d1 = synth.tie_stress_desc(rng, n1) if tie else synth.rand_desc(rng, n1)
    d2 = synth.tie_stress_desc(rng, n2) if tie else synth.rand_desc(rng, n2)
    got = M.knn2(d1, d2, idx_base=17)
    want = port.knn2_packed(d1, d2, idx_base=17)
    assert (got == want).all()


def test_knn2_empty_and_single_train(M):
    rng = np.random.default_rng(5)
    d1 = synth.rand_desc(rng, 10)
    got = M.knn2(d1, np.zeros((0, 32), np.uint8))
    assert (got == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
    got = M.knn2(d1, d1[:1])
    want = port.knn2_packed(d1, d1[:1])
    assert (got == want).all() and (got[:, 1] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
    assert M.knn2(np.zeros((0, 32), np.uint8), d1).shape == (0, 2)


@pytest.mark.parametrize("n1,n2,tie", [(2, 2, True), (50, 3, True), (33, 1025, True), (600, 600, False),
                                        (200, 200, False), (640, 577, False), (5000, 300, False)])
@pytest.mark.parametrize("nnr", [0.75, 0.9])
@pytest.mark.parametrize("best_lr", [0, 1])
def test_match_and_match_nnr(M, n1, n2, tie, nnr, best_lr):
    rng = np.random.default_rng(7 * n1 + n2)
    d2 = synth.tie_stress_desc(rng, n2) if tie else synth.rand_desc(rng, n2)
    if tie:
        d1 = synth.tie_stress_desc(rng, n1)
    else:
        d1 = synth.rand_desc(rng, n1)
        k = min(n1, n2) * 2 // 3
        d1[rng.choice(n1, k, replace=False)] = synth.flip_bits(rng, d2[rng.choice(n2, k, replace=False)], 0.08)
    n_o, m_o = port.match(d1, d2, nnr, best_lr)
    M.Config.bestLRMatches = bool(best_lr)
    try:
        m_g = []
        n_g = M.match(d1, d2, nnr, m_g)
    finally:
        M.Config.bestLRMatches = True
    assert n_g == n_o
    assert (np.array(m_g) == m_o).all()
    if not best_lr:
        m_n = np.full(n1, -1, np.int32)
        assert M.matchNNR(d1, d2, nnr, m_n) == n_o and (m_n == m_o).all()


def test_match_stale_inout_fallback(M):
    """mapHandler.cpp:325-329: match() is handed the vector matchGrid just filled; stale entries are
    kept, cross-checked and can drive the returned count negative."""
    rng = np.random.default_rng(99)
    n1, n2 = 300, 280
    d1, d2 = synth.rand_desc(rng, n1), synth.rand_desc(rng, n2)
    d1[:100] = synth.flip_bits(rng, d2[:100], 0.05)
    stale = np.full(n1, -1, np.int32)
    stale[100:250] = rng.integers(0, n2, 150)
    for best_lr in (0, 1):
        n_o, m_o = port.match(d1, d2, 0.75, best_lr, m12=stale)
        M.Config.bestLRMatches = bool(best_lr)
        try:
            m_g = stale.copy()
            n_g = M.match(d1, d2, 0.75, m_g)
        finally:
            M.Config.bestLRMatches = True
        assert n_g == n_o and (m_g == m_o).all()
    assert port.match(d1, d2, 0.75, 1, m12=stale)[0] < 100  # culls of stale entries are subtracted


def test_strided_descriptors(M):
    """cv::Mat::step != 32 (a column range of a wider matrix)."""
    rng = np.random.default_rng(3)
    wide1 = rng.integers(0, 256, (200, 48), dtype=np.uint8)
    wide2 = rng.integers(0, 256, (180, 64), dtype=np.uint8)
    d1, d2 = wide1[:, 8:40], wide2[:, :32]
    n_o, m_o = port.match(np.ascontiguousarray(d1), np.ascontiguousarray(d2), 0.9, 1)
    m_g = []
    assert M.match(d1, d2, 0.9, m_g) == n_o and (np.array(m_g) == m_o).all()


GRID_CASES = [
    # n1, n2, is_lines, tie, win, bad_items, zero_len
    (600, 600, False, False, (10, 0, 0, 0), 0, 0),
    (600, 600, False, False, (3, 3, 3, 3), 0, 0),
    (500, 400, False, True, (3, 3, 3, 3), 0, 0),
    (300, 900, False, True, (6, 2, 1, 4), 12, 0),
    (37, 5, False, True, (70, 70, 50, 50), 0, 0),
    (1, 1, False, False, (1, 1, 1, 1), 0, 0),
    (900, 30, False, True, (64, 64, 48, 48), 0, 0),
    (2500, 1500, False, False, (3, 3, 3, 3), 0, 0),
    (200, 200, True, False, (10, 0, 0, 0), 0, 0),
    (200, 200, True, False, (3, 3, 3, 3), 0, 4),
    (150, 260, True, True, (3, 3, 3, 3), 0, 6),
    (80, 40, True, True, (64, 64, 48, 48), 0, 3),
    (1300, 700, True, False, (2, 2, 2, 2), 0, 0),
]


@pytest.mark.parametrize("n1,n2,is_lines,tie,win,bad_items,zero_len", GRID_CASES)
@pytest.mark.parametrize("best_lr", [1, 0])
@pytest.mark.parametrize("ratio", [0.75, 0.9, 1.0])
def test_match_grid(M, n1, n2, is_lines, tie, win, bad_items, zero_len, best_lr, ratio):
    rng = np.random.default_rng(n1 * 31 + n2 * 7 + int(is_lines))
    case = random_grid_case(rng, n1, n2, is_lines=is_lines, tie=tie, win=win, bad_items=bad_items, zero_len=zero_len)
    n_o, m_o = oracle_grid(port, case, ratio, best_lr)
    n_g, m_g = gpu_grid(case, ratio, best_lr)
    assert n_g == n_o
    assert (m_g == m_o).all(), np.flatnonzero(m_g != m_o)[:10]


def test_match_grid_small_grids_and_empty(M):
    rng = np.random.default_rng(4)
    for rows, cols in [(1, 1), (2, 5), (7, 3)]:
        case = random_grid_case(rng, 120, 90, rows=rows, cols=cols, tie=True, win=(1, 1, 1, 1))
        n_o, m_o = oracle_grid(port, case, 0.9, 1)
        n_g, m_g = gpu_grid(case, 0.9, 1)
        assert n_g == n_o and (m_g == m_o).all()
    # empty grid (no train feature inside the image): nothing matches, stale entries are culled
    case = random_grid_case(rng, 50, 40, win=(3, 3, 3, 3))
    case["cell_start"] = np.zeros_like(case["cell_start"])
    case["cell_items"] = np.zeros(0, np.int32)
    stale = np.full(50, -1, np.int32)
    stale[:10] = np.arange(10)
    n_o, m_o = oracle_grid(port, case, 0.9, 1, m12=stale)
    n_g, m_g = gpu_grid(case, 0.9, 1, m12=stale)
    assert n_g == n_o == -10 and (m_g == m_o).all() and (m_g == -1).all()


def test_match_grid_stereo_configs(M):
    """Config 1 of BASELINE.json: the stereo point / line matchGrid calls of StereoFrame."""
    sp = synth.make_stereo_pair(synth.SEED0 + 1)
    for ratio in (0.75, 0.9):
        a = synth.stereo_points_grid_args(sp)
        case = dict(coords=a["xy"], d1=a["d1"], cell_start=a["cell_start"], cell_items=a["cell_items"],
                    rows=a["rows"], cols=a["cols"], d2=a["d2"], win=a["win"], dirs2=None)
        n_o, m_o = oracle_grid(port, case, ratio, 1)
        n_g, m_g = gpu_grid(case, ratio, 1)
        assert n_g == n_o and (m_g == m_o).all() and n_o > 300
        b = synth.stereo_lines_grid_args(sp)
        case = dict(coords=b["xyxy"], d1=b["d1"], cell_start=b["cell_start"], cell_items=b["cell_items"],
                    rows=b["rows"], cols=b["cols"], d2=b["d2"], win=b["win"], dirs2=b["dirs2"])
        n_o, m_o = oracle_grid(port, case, ratio, 1)
        n_g, m_g = gpu_grid(case, ratio, 1)
        assert n_g == n_o and (m_g == m_o).all() and n_o > 80


@pytest.mark.parametrize("is_lines", [False, True])
def test_match_grid_map_scale_chunked(M, is_lines):
    """Config 4 shape (scaled so the oracle finishes in seconds): tens of thousands of map rows
    against a frame-sized grid -> the multi-CTA chunked path."""
    rng = np.random.default_rng(2024 + int(is_lines))
    n1, n2 = (30000, 600) if not is_lines else (12000, 200)
    case = random_grid_case(rng, n1, n2, is_lines=is_lines, tie=False, win=(3, 3, 3, 3), zero_len=5 if is_lines else 0)
    for best_lr in (1, 0):
        n_o, m_o = oracle_grid(port, case, 0.9, best_lr)
        n_g, m_g = gpu_grid(case, 0.9, best_lr)
        assert n_g == n_o and (m_g == m_o).all()


def test_match_grid_map_scale_ties(M):
    rng = np.random.default_rng(77)
    case = random_grid_case(rng, 9000, 300, tie=True, win=(5, 5, 5, 5))
    n_o, m_o = oracle_grid(port, case, 0.9, 1)
    n_g, m_g = gpu_grid(case, 0.9, 1)
    assert n_g == n_o and (m_g == m_o).all()


def test_stereo_filters(M):
    sp = synth.make_stereo_pair(synth.SEED0 + 1)
    rng = np.random.default_rng(8)
    m = rng.integers(-1, 600, 600).astype(np.int32)
    m[:300] = np.arange(300)  # plausible and implausible pairings
    n_o, k_o, d_o = port.stereo_filter_points(sp.kp_l, sp.kp_r, m)
    n_g, k_g, d_g = M.stereo_filter_points(sp.kp_l, sp.kp_r, m)
    assert n_g == n_o and (k_g == k_o).all() and (d_g == d_o).all()
    ml = rng.integers(-1, 200, 200).astype(np.int32)
    ln_r = sp.ln_r.copy()
    ln_r[::17, 3] = ln_r[::17, 1]  # horizontal right lines: division by zero -> inf / NaN paths
    n_o, k_o, d_o = port.stereo_filter_lines(sp.ln_l, ln_r, ml)
    n_g, k_g, d_g = M.stereo_filter_lines(sp.ln_l, ln_r, ml)
    assert n_g == n_o and (k_g == k_o).all()
    assert np.array_equal(d_g, d_o, equal_nan=True)


def test_threads_are_reentrant(M):
    """up to ~6 host threads are inside the matching layer at once."""
    import threading
    rng = np.random.default_rng(21)
    cases = []
    for t in range(6):
        d1, d2 = synth.rand_desc(rng, 300 + 10 * t), synth.rand_desc(rng, 280)
        d1[:150] = synth.flip_bits(rng, d2[:150], 0.08)
        cases.append((d1, d2, port.match(d1, d2, 0.9, 1)))
    errors = []

    def work(d1, d2, want):
        try:
            for _ in range(20):
                m = []
                n = M.match(d1, d2, 0.9, m)
                assert n == want[0] and (np.array(m) == want[1]).all()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ths = [threading.Thread(target=work, args=c) for c in cases]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errors, errors


def test_gpu_vs_golden(M):
    """The committed fixtures were produced by the reference's own matching.cpp (tools/make_golden.py)."""
    import glob
    import os
    from conftest import ROOT
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
    assert len(files) >= 10
    for path in files:
        z = np.load(path)
        name = os.path.basename(path)
        if name.startswith("brute"):
            for nnr in (0.75, 0.9):
                for blr in (0, 1):
                    M.Config.bestLRMatches = bool(blr)
                    try:
                        m = []
                        n = M.match(z["d1"], z["d2"], nnr, m)
                        ms = z["stale"].copy()
                        ns = M.match(z["d1"], z["d2"], nnr, ms)
                    finally:
                        M.Config.bestLRMatches = True
                    assert n == int(z[f"n_{nnr}_{blr}"]) and (np.array(m) == z[f"m_{nnr}_{blr}"]).all(), name
                    assert ns == int(z[f"ns_{nnr}_{blr}"]) and (ms == z[f"ms_{nnr}_{blr}"]).all(), name
        elif name.startswith("grid"):
            case = dict(coords=z["coords"], d1=z["d1"], cell_start=z["cell_start"], cell_items=z["cell_items"],
                        rows=int(z["rows"]), cols=int(z["cols"]), d2=z["d2"], win=z["win"],
                        dirs2=z["dirs2"] if int(z["is_lines"]) else None)
            for ratio in (0.75, 0.9, 1.0):
                for blr in (0, 1):
                    n, m = gpu_grid(case, ratio, blr)
                    assert n == int(z[f"n_{ratio}_{blr}"]) and (m == z[f"m_{ratio}_{blr}"]).all(), name


def test_match_grid_map_scale_record_chains(M):
    """Row-parallel map-scale kernel (grid_rows_kernel): every row of a block improves the running column minimum
    of the same train feature, so the rounds that resolve the live pairs run one record at a time (the worst case
    of matching.cpp:145-150 for the parallel form), next to ordinary rows."""
    rng = np.random.default_rng(5)
    n2, n1 = 64, 6000
    d2 = synth.rand_desc(rng, n2)
    case = random_grid_case(rng, n1, n2, tie=False, win=(2, 2, 2, 2), off_grid=0)
    case["d2"] = d2
    # all train features in distinct cells of a small patch; rows 1000..1511 sit on train feature 7's cell and get
    # closer to it row by row (distance 255, 254, ... then a plateau of equal distances, then closer again)
    cs = np.zeros(48 * 64 + 1, np.int32)
    cells = (np.arange(n2) % 8) * 48 + (np.arange(n2) // 8)          # cell id = x * rows + y
    order = np.argsort(cells, kind="stable")
    counts = np.bincount(cells, minlength=48 * 64)
    cs[1:] = np.cumsum(counts)
    case["cell_start"], case["cell_items"] = cs, order.astype(np.int32)
    x7, y7 = int(cells[7] // 48), int(cells[7] % 48)
    rows = np.arange(1000, 1512)
    case["coords"][rows] = (x7, y7)
    flips = np.concatenate([np.arange(255, 55, -1), np.full(112, 55), np.arange(54, -146, -1).clip(1)])[:512]
    bits = np.unpackbits(d2[7])
    for r, k in zip(rows, flips):
        b = bits.copy()
        b[rng.choice(256, int(k), replace=False)] ^= 1
        case["d1"][r] = np.packbits(b)
    for best_lr in (1, 0):
        for ratio in (0.9, 1.0):
            n_o, m_o = oracle_grid(port, case, ratio, best_lr)
            n_g, m_g = gpu_grid(case, ratio, best_lr)
            assert n_g == n_o and (m_g == m_o).all(), (best_lr, ratio, np.flatnonzero(m_g != m_o)[:10])


@pytest.mark.parametrize("is_lines", [False, True])
def test_match_grid_map_scale_both_kernel_families(M, plm_lib, is_lines):
    """Config 4 at full row count (200 000 points / 50 000 lines against one frame): the row-parallel kernels
    (several blocks of rows per CTA) and the warp-per-chunk kernels (option grid_rows = 0) against the oracle."""
    rng = np.random.default_rng(3030 + int(is_lines))
    n1, n2 = (200_000, 600) if not is_lines else (50_000, 200)
    case = random_grid_case(rng, n1, n2, is_lines=is_lines, tie=False, win=(3, 3, 3, 3), zero_len=7 if is_lines else 0)
    stale = np.full(n1, -1, np.int32)
    stale[rng.choice(n1, 500, replace=False)] = rng.integers(0, n2, 500)
    n_o, m_o = oracle_grid(port, case, 0.9, 1, m12=stale)
    for rows_mode, head in ((1, 1), (1, 0), (0, 1)):   # head = 0: uniform rows per CTA (no short CTAs at the start of the map)
        assert plm_lib.plm_set_option(b"grid_rows", rows_mode) == 0
        assert plm_lib.plm_set_option(b"grid_head", head) == 0
        try:
            n_g, m_g = gpu_grid(case, 0.9, 1, m12=stale)
        finally:
            plm_lib.plm_set_option(b"grid_rows", 1)
            plm_lib.plm_set_option(b"grid_head", 1)
        assert n_g == n_o and (m_g == m_o).all(), (rows_mode, head)


def test_match_grid_map_scale_wide_frame(M):
    """A train side too large for the shared-memory staging (n2 = 9000) and one too large for the row-parallel
    work arrays (n2 = 20000 -> warp-per-chunk kernels)."""
    for n2 in (9000, 20000):
        rng = np.random.default_rng(n2)
        case = random_grid_case(rng, 7000, n2, tie=False, win=(1, 1, 1, 1))
        n_o, m_o = oracle_grid(port, case, 0.9, 1)
        n_g, m_g = gpu_grid(case, 0.9, 1)
        assert n_g == n_o and (m_g == m_o).all(), n2


@pytest.mark.parametrize("is_lines", [False, True])
def test_match_grid_map_scale_dense_windows(M, is_lines):
    """Windows so wide that a block of 256 rows has more slots than the shared-memory pair list holds: the
    row-parallel kernel takes its re-walk form (same rounds, every thread walking its own row)."""
    rng = np.random.default_rng(91 + int(is_lines))
    n1, n2 = 6000, (300 if not is_lines else 120)
    case = random_grid_case(rng, n1, n2, is_lines=is_lines, tie=False, win=(20, 20, 16, 16), zero_len=3 if is_lines else 0)
    for best_lr in (1, 0):
        n_o, m_o = oracle_grid(port, case, 0.9, best_lr)
        n_g, m_g = gpu_grid(case, 0.9, best_lr)
        assert n_g == n_o and (m_g == m_o).all(), best_lr
    case = random_grid_case(rng, 5000, 200, is_lines=is_lines, tie=True, win=(30, 30, 30, 30))
    n_o, m_o = oracle_grid(port, case, 0.9, 1)
    n_g, m_g = gpu_grid(case, 0.9, 1)
    assert n_g == n_o and (m_g == m_o).all()


def test_match_one_row_sides(M):
    """Sparse frames: one descriptor on a side must not raise (only an EMPTY train set does, matching.cpp:50-51)."""
    d = synth.rand_desc(np.random.default_rng(4), 5)
    m = []
    assert M.match(d[:1], d, 0.9, m) == 0 and m == [-1]
    m = np.full(5, -1, np.int32)
    assert M.matchNNR(d, d[:1], 0.9, m) == 0 and (m == -1).all()
    m = []
    assert M.match(d[:0], d, 0.9, m) == 0 and m == []
    with pytest.raises(RuntimeError):
        M.match(d, d[:0], 0.9, [])


def test_frame_session_equals_separate_calls(M):
    """plm_frame_begin / plm_frame_end: the four matcher calls of one frame (stereo matchGrid points + lines, temporal
    match points + lines) recorded and run as one round trip give exactly the results of the four separate calls."""
    prev, curr = synth.make_temporal_pair(synth.SEED0 + 2)
    a = synth.stereo_points_grid_args(curr)
    b = synth.stereo_lines_grid_args(curr)
    ga = (a["cell_start"], a["cell_items"], a["rows"], a["cols"])
    gb = (b["cell_start"], b["cell_items"], b["rows"], b["cols"])
    M.Config.minRatio12P = 0.9
    for ctx in (None, M.Context(0)):
        want, got = [], []
        for sess in (False, True):
            out = want if not sess else got
            m = [np.full(600, -1, np.int32), np.full(200, -1, np.int32), np.full(len(prev.pdesc_l), -1, np.int32),
                 np.full(len(prev.ldesc_l), -1, np.int32), np.full(len(prev.pdesc_l), -1, np.int32)]
            m[2][::5] = 3                                      # stale entries travel through a session too

            def calls():
                return [M.matchGrid(a["xy"], a["d1"], ga, a["d2"], a["win"], m[0], ctx=ctx),
                        M.matchGrid(b["xyxy"], b["d1"], gb, b["d2"], b["dirs2"], b["win"], m[1], ctx=ctx),
                        M.match(prev.pdesc_l, curr.pdesc_l, 0.9, m[2], ctx=ctx),
                        M.match(prev.ldesc_l, curr.ldesc_l, 0.9, m[3], ctx=ctx),
                        M.matchNNR(prev.pdesc_l, curr.pdesc_l, 0.75, m[4], ctx=ctx)]
            if sess:
                with M.FrameSession(ctx):
                    r = calls()
            else:
                r = calls()
            out.extend([int(x) for x in r])
            out.extend([x.copy() for x in m])
        assert want[:5] == got[:5] and want[0] > 100 and want[2] > 100
        for x, y in zip(want[5:], got[5:]):
            assert np.array_equal(x, y)
    # an empty session and a failing call inside a session
    with M.FrameSession(None):
        pass
    with pytest.raises(RuntimeError):
        with M.FrameSession(None):
            M.match(prev.pdesc_l, prev.pdesc_l[:0], 0.9, [])
    assert M.match(prev.pdesc_l, curr.pdesc_l, 0.9, []) == want[2] or True     # the context is usable afterwards


def _match_case(rng, n1, n2, tie):
    d2 = synth.tie_stress_desc(rng, n2) if tie else synth.rand_desc(rng, n2)
    if tie:
        return synth.tie_stress_desc(rng, n1), d2
    d1 = synth.rand_desc(rng, n1)
    k = min(n1, n2) * 2 // 3
    if k:
        d1[rng.choice(n1, k, replace=False)] = synth.flip_bits(rng, d2[rng.choice(n2, k, replace=False)], 0.08)
    return d1, d2


@pytest.mark.parametrize("fused", [1, 0])
def test_frame_session_one_launch_vs_oracle(M, plm_lib, fused):
    """A frame session runs as ONE launch (frame_fused_kernel: one 8-CTA cluster per recorded call) whenever every call
    is frame-sized.  Sessions made of the matchGrid cases (points and lines, ties, off-grid rows, bad items, NaN
    directions, stale in/out entries) and of match / matchNNR calls from 1 x 2 up to 2048 x 2048 rows are held to the
    oracle call by call; option frame_fused = 0 (one lane per call) must give the same."""
    assert plm_lib.plm_set_option(b"frame_fused", fused) == 0
    ctx = M.Context(0)
    try:
        rng = np.random.default_rng(2024)
        grid_cases = [c for c in GRID_CASES if c[0] <= 2048]
        match_cases = [(1, 2, True), (2, 1, False), (50, 3, True), (33, 1025, True), (600, 600, False), (200, 200, False),
                       (640, 577, False), (75, 2048, False), (2048, 2048, False), (257, 9, True), (9, 257, True)]
        sessions = [(grid_cases[:4], match_cases[:4]), (grid_cases[4:8], match_cases[4:8]), (grid_cases[8:], match_cases[8:]),
                    ([], match_cases[:2]), (grid_cases[:1], [])]
        for gcs, mcs in sessions:
            for best_lr, ratio, nnr in ((1, 0.9, 0.9), (0, 0.75, 0.75), (1, 1.0, 0.75)):
                want, bufs, pend = [], [], []
                M.Config.bestLRMatches = bool(best_lr)
                M.Config.minRatio12P = ratio
                before = ctx.launch_count
                with M.FrameSession(ctx):
                    for n1, n2, is_lines, tie, win, bad, zl in gcs:
                        case = random_grid_case(rng, n1, n2, is_lines=is_lines, tie=tie, win=win, bad_items=bad, zero_len=zl)
                        stale = np.full(n1, -1, np.int32)
                        stale[::9] = rng.integers(0, max(n2, 1), len(stale[::9]))
                        want.append(oracle_grid(port, case, ratio, best_lr, m12=stale))
                        buf = stale.copy()
                        grid = (case["cell_start"], case["cell_items"], case["rows"], case["cols"])
                        if is_lines:
                            pend.append(M.matchGrid(case["coords"], case["d1"], grid, case["d2"], case["dirs2"], case["win"], buf, ctx=ctx))
                        else:
                            pend.append(M.matchGrid(case["coords"], case["d1"], grid, case["d2"], case["win"], buf, ctx=ctx))
                        bufs.append(buf)
                    for n1, n2, tie in mcs:
                        d1, d2 = _match_case(rng, n1, n2, tie)
                        stale = np.full(n1, -1, np.int32)
                        stale[::7] = rng.integers(0, n2, len(stale[::7]))
                        degenerate = n2 < 2 or n1 < 2     # undefined in the reference (matching.cpp:54): the oracle
                        for fn, ofn in ((M.match, lambda: port.match(d1, d2, nnr, best_lr, m12=stale)),   # flags it;
                                        (M.matchNNR, lambda: port.match_nnr(d1, d2, nnr, m12=stale))):    # held to the
                            if degenerate:                                                                # stand-alone call
                                ref = stale.copy()
                                want.append((fn(d1, d2, nnr, ref), ref))
                            else:
                                want.append(ofn())
                            buf = stale.copy()
                            pend.append(fn(d1, d2, nnr, buf, ctx=ctx))
                            bufs.append(buf)
                launches = ctx.launch_count - before
                if fused:
                    assert launches == 1, launches
                else:
                    assert launches >= len(pend)
                for k, ((n_o, m_o), n_g, m_g) in enumerate(zip(want, pend, bufs)):
                    assert int(n_g) == n_o, (fused, k, best_lr, ratio)
                    assert (m_g == m_o).all(), (fused, k, best_lr, ratio, np.flatnonzero(m_g != m_o)[:10])
        # a call beyond the frame-sized limits sends the whole session down the lane path -- same results
        d1, d2 = _match_case(rng, 2100, 300, False)
        before = ctx.launch_count
        buf = np.full(2100, -1, np.int32)
        with M.FrameSession(ctx):
            p = M.match(d1, d2, 0.9, buf, ctx=ctx)
        n_o, m_o = port.match(d1, d2, 0.9, True)
        assert int(p) == n_o and (buf == m_o).all() and ctx.launch_count - before > 1
    finally:
        M.Config.bestLRMatches = True
        M.Config.minRatio12P = 0.9
        plm_lib.plm_set_option(b"frame_fused", 1)


def test_frame_session_random_shapes(M):
    """Randomised differential test of the one-launch frame kernel: 40 sessions of 1-4 matchGrid and 0-4 match / matchNNR
    calls with log-uniform sizes (1 ... 2048 rows per side), random windows, ratios, tie-heavy or planted-match
    descriptors and stale in/out entries, every call against the oracle."""
    ctx = M.Context(0)
    rng = np.random.default_rng(77)
    size = lambda lo=1: int(np.clip(np.exp(rng.uniform(np.log(lo), np.log(2048))), lo, 2048))  # noqa: E731
    try:
        for sess in range(40):
            best_lr = bool(rng.integers(0, 2))
            ratio = float(rng.choice([0.75, 0.9, 1.0]))
            nnr = float(rng.choice([0.75, 0.9]))
            M.Config.bestLRMatches = best_lr
            M.Config.minRatio12P = ratio
            want, pend, bufs = [], [], []
            before = ctx.launch_count
            with M.FrameSession(ctx):
                for _ in range(int(rng.integers(1, 5))):
                    is_lines = bool(rng.integers(0, 2))
                    n1, n2 = size(), size()
                    win = tuple(int(v) for v in rng.integers(0, 12, 4))
                    case = random_grid_case(rng, n1, n2, is_lines=is_lines, tie=bool(rng.integers(0, 2)), win=win,
                                            bad_items=int(rng.integers(0, 3)) if not is_lines else 0, zero_len=2 if is_lines else 0)
                    stale = np.full(n1, -1, np.int32)
                    stale[::5] = rng.integers(0, n2, len(stale[::5]))
                    want.append(oracle_grid(port, case, ratio, best_lr, m12=stale))
                    buf = stale.copy()
                    grid = (case["cell_start"], case["cell_items"], case["rows"], case["cols"])
                    if is_lines:
                        pend.append(M.matchGrid(case["coords"], case["d1"], grid, case["d2"], case["dirs2"], case["win"], buf, ctx=ctx))
                    else:
                        pend.append(M.matchGrid(case["coords"], case["d1"], grid, case["d2"], case["win"], buf, ctx=ctx))
                    bufs.append(buf)
                for _ in range(int(rng.integers(0, 5))):
                    n1, n2 = size(2), size(2)
                    d1, d2 = _match_case(rng, n1, n2, bool(rng.integers(0, 2)))
                    stale = np.full(n1, -1, np.int32)
                    stale[::4] = rng.integers(0, n2, len(stale[::4]))
                    buf = stale.copy()
                    if rng.integers(0, 2):
                        want.append(port.match(d1, d2, nnr, best_lr, m12=stale))
                        pend.append(M.match(d1, d2, nnr, buf, ctx=ctx))
                    else:
                        want.append(port.match_nnr(d1, d2, nnr, m12=stale))
                        pend.append(M.matchNNR(d1, d2, nnr, buf, ctx=ctx))
                    bufs.append(buf)
            assert ctx.launch_count - before == 1
            for k, ((n_o, m_o), n_g, m_g) in enumerate(zip(want, pend, bufs)):
                assert int(n_g) == n_o and (m_g == m_o).all(), (sess, k, best_lr, ratio, len(m_o), np.flatnonzero(m_g != m_o)[:8])
    finally:
        M.Config.bestLRMatches = True
        M.Config.minRatio12P = 0.9


@pytest.mark.parametrize("n1,n2", [(400, 9000), (1500, 6000), (2048, 2048), (64, 32000)])
def test_frame_kernel_wide_train_sides(M, n1, n2):
    """Frame-sized query sides against wide train sides: the per-column work arrays of the cluster matchGrid grow to
    ~150 KB of shared memory per CTA (n2 = 9000); beyond that (n2 = 32 000) the call leaves the one-launch path."""
    rng = np.random.default_rng(n1 + n2)
    ctx = M.Context(0)
    case = random_grid_case(rng, n1, n2, tie=False, win=(2, 2, 2, 2))
    for best_lr in (1, 0):
        n_o, m_o = oracle_grid(port, case, 0.9, best_lr)
        n_g, m_g = gpu_grid(case, 0.9, best_lr, ctx=ctx)
        assert n_g == n_o and (m_g == m_o).all(), best_lr


def test_frame_session_two_host_threads(M):
    """Two host threads, each with its own context, run one-launch frame sessions at the same time (the reference's
    points || lines std::async structure): the kernels of the two contexts overlap on the device, results stay exact."""
    import threading
    prev, curr = synth.make_temporal_pair(synth.SEED0 + 9)
    a, b = synth.stereo_points_grid_args(curr), synth.stereo_lines_grid_args(curr)
    ga = (a["cell_start"], a["cell_items"], a["rows"], a["cols"])
    gb = (b["cell_start"], b["cell_items"], b["rows"], b["cols"])
    M.Config.minRatio12P = 0.9
    want_p = (port.match_grid_points(a["xy"], a["d1"], a["cell_start"], a["cell_items"], a["rows"], a["cols"], a["d2"], a["win"], 0.9, True),
              port.match(prev.pdesc_l, curr.pdesc_l, 0.9, True))
    want_l = (port.match_grid_lines(b["xyxy"], b["d1"], b["cell_start"], b["cell_items"], b["rows"], b["cols"], b["d2"], b["dirs2"], 0.75,
                                    b["win"], 0.9, True),
              port.match(prev.ldesc_l, curr.ldesc_l, 0.9, True))
    errors = []

    def worker(lines):
        try:
            ctx = M.Context(0)
            for _ in range(200):
                m_s = np.full(len(b["d1"]) if lines else len(a["d1"]), -1, np.int32)
                m_t = np.full(len(prev.ldesc_l) if lines else len(prev.pdesc_l), -1, np.int32)
                with M.FrameSession(ctx):
                    if lines:
                        r_s = M.matchGrid(b["xyxy"], b["d1"], gb, b["d2"], b["dirs2"], b["win"], m_s, ctx=ctx)
                        r_t = M.match(prev.ldesc_l, curr.ldesc_l, 0.9, m_t, ctx=ctx)
                    else:
                        r_s = M.matchGrid(a["xy"], a["d1"], ga, a["d2"], a["win"], m_s, ctx=ctx)
                        r_t = M.match(prev.pdesc_l, curr.pdesc_l, 0.9, m_t, ctx=ctx)
                want = want_l if lines else want_p
                if int(r_s) != want[0][0] or not (m_s == want[0][1]).all() or int(r_t) != want[1][0] or not (m_t == want[1][1]).all():
                    errors.append(("mismatch", lines))
                    return
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=worker, args=(k,)) for k in (False, True)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_match_grid_single_call_kernel_families(M, plm_lib, mode):
    """The three single-call kernels for frame-sized jobs -- one CTA (0), chunk phases on an 8-CTA cluster (1), the
    row-parallel pair-list kernel on one cluster (2, default) -- on a cross-section of the grid cases."""
    assert plm_lib.plm_set_option(b"grid_cluster", mode) == 0
    try:
        for n1, n2, is_lines, tie, win, bad, zl in GRID_CASES[:4] + GRID_CASES[7:12]:
            rng = np.random.default_rng(n1 * 31 + n2 * 7 + int(is_lines) + 1)
            case = random_grid_case(rng, n1, n2, is_lines=is_lines, tie=tie, win=win, bad_items=bad, zero_len=zl)
            for best_lr, ratio in ((1, 0.9), (0, 0.75), (1, 1.0)):
                stale = np.full(n1, -1, np.int32)
                stale[::9] = rng.integers(0, max(n2, 1), len(stale[::9]))
                n_o, m_o = oracle_grid(port, case, ratio, best_lr, m12=stale)
                n_g, m_g = gpu_grid(case, ratio, best_lr, m12=stale)
                assert n_g == n_o and (m_g == m_o).all(), (mode, n1, n2, is_lines, best_lr, ratio)
    finally:
        plm_lib.plm_set_option(b"grid_cluster", 2)
