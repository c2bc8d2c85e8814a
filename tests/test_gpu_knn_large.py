"""Parity of the brute-force kernel variants the headline number comes from.

`knn2_slice_kernel<*, 2/3>` (carry-save Hamming + blocked threshold top-2) is only selected for train sets
of >= 65536 rows (plmatch.cu `plan_knn`), so these cases use 65 535 ... 140 000 train rows for both CTA
widths (n1 < 4096 -> 64 threads, n1 >= 4096 -> 128 threads), random descriptors with planted matches and
the tie-stress set, ragged tails (n2 % 8 != 0, n2 % 256 != 0), and every forced variant.  Reference
semantics: stvo-pl/src/matching.cpp:41-61 over cv::BFMatcher::knnMatch(k=2) (lowest train index wins ties).
"""
import numpy as np
import pytest

import oracle
from pl_inertial_slam_b200 import synth

pytestmark = pytest.mark.gpu
port = oracle.port


@pytest.fixture(scope="module")
def M(plm_lib):
    from pl_inertial_slam_b200 import matching
    return matching


def _case(seed, n1, n2, tie):
    rng = np.random.default_rng(seed)
    if tie:
        # many exact duplicates and only 2^8 distinct rows: every query has thousands of equal distances
        return synth.tie_stress_desc(rng, n1), synth.tie_stress_desc(rng, n2)
    d1, d2 = synth.rand_desc(rng, n1), synth.rand_desc(rng, n2)
    k = n1 * 2 // 3
    rows = rng.choice(n1, k, replace=False)
    d1[rows] = synth.flip_bits(rng, d2[rng.integers(0, n2, k)], 0.08)
    # exact duplicates of a train row far apart: the lower index must win both slots' ties
    src = rng.integers(0, n2 // 2, 64)
    d2[n2 - 1 - np.arange(64)] = d2[src]
    return d1, d2


@pytest.mark.parametrize("n1,n2,tie", [
    (300, 70000, False), (300, 70000, True),          # 64-thread CTAs
    (4100, 70000, False), (4100, 70003, True),        # 128-thread CTAs, ragged tail
    (70, 65535, False), (70, 65536, True), (70, 65537, False),
    (4096, 65536, True), (129, 140001, False),
])
def test_knn2_large_train_sets(M, n1, n2, tie):
    d1, d2 = _case(7000 + n1 + n2, n1, n2, tie)
    got = M.knn2(d1, d2, idx_base=5)
    want = port.knn2_packed(d1, d2, idx_base=5)
    assert (got == want).all(), np.flatnonzero((got != want).any(1))[:10]


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("n1,n2", [(300, 66000), (4100, 66000), (200, 3000)])
def test_knn2_every_variant(M, plm_lib, variant, n1, n2):
    d1, d2 = _case(99 + n1, n1, n2, tie=(variant % 2 == 0))
    want = port.knn2_packed(d1, d2)
    assert plm_lib.plm_set_option(b"knn_variant", variant) == 0
    try:
        got = M.knn2(d1, d2)
    finally:
        plm_lib.plm_set_option(b"knn_variant", -1)
    assert (got == want).all()


@pytest.mark.parametrize("n1,n2,tie", [(4100, 70003, False), (4096, 65536, True), (9001, 66001, False), (4097, 131072, True)])
def test_knn2_two_queries_per_thread(M, plm_lib, n1, n2, tie):
    """Option knn_qpt = 2: knn2_slice_kernel<128, 6> (a CTA owns 256 queries, thread t the queries t and t + 128);
    ragged query blocks (n1 % 256 in {4, 0, 41, 1}) and ragged train tails."""
    d1, d2 = _case(555 + n1, n1, n2, tie)
    want = port.knn2_packed(d1, d2, idx_base=3)
    assert plm_lib.plm_set_option(b"knn_qpt", 2) == 0
    try:
        got = M.knn2(d1, d2, idx_base=3)
        m_g = np.full(n1, -1, np.int32)
        n_g = M.matchNNR(d1, d2, 0.8, m_g)
    finally:
        plm_lib.plm_set_option(b"knn_qpt", 1)
    assert (got == want).all(), np.flatnonzero((got != want).any(1))[:10]
    n_o, m_o = port.match_nnr(d1, d2, 0.8)
    assert n_g == n_o and (m_g == m_o).all()


@pytest.mark.parametrize("n1,n2", [(300, 70000), (4100, 70000)])
@pytest.mark.parametrize("nnr", [0.75, 0.9])
def test_match_nnr_large_vs_reference(M, n1, n2, nnr):
    """matchNNR against the compiled reference matching.cpp (or the restatement when it is not built)."""
    d1, d2 = _case(31 + n1, n1, n2, tie=False)
    n_o, m_o = oracle.ref.match_nnr(d1, d2, nnr) if oracle.ref.available() else port.match_nnr(d1, d2, nnr)
    m_g = np.full(n1, -1, np.int32)
    n_g = M.matchNNR(d1, d2, nnr, m_g)
    assert n_g == n_o and (m_g == m_o).all() and n_o > n1 // 2


def test_match_bidirectional_large(M):
    """StVO::match with a map-sized query side (config 4 fallback: 200 000 x 600 scaled to 70 000 x 600):
    the 21 direction scans >= 65536 rows and therefore runs the blocked kernel."""
    rng = np.random.default_rng(8)
    n1, n2 = 70000, 600
    d1, d2 = synth.rand_desc(rng, n1), synth.rand_desc(rng, n2)
    rows = rng.choice(n1, 500, replace=False)
    d1[rows] = synth.flip_bits(rng, d2[rng.integers(0, n2, 500)], 0.06)
    n_o, m_o = port.match(d1, d2, 0.9, 1)
    m_g = []
    n_g = M.match(d1, d2, 0.9, m_g)
    assert n_g == n_o and (np.array(m_g) == m_o).all() and n_o > 200


def test_full_size_property_checks(M):
    """BASELINE config 5 shard size (2 M rows): size-independent properties instead of a full oracle pass --
    (i) the result for a block of queries equals the lexicographic merge of the results over two halves of the
    train set (idx_base carries the global index), (ii) a planted exact duplicate is found at distance 0 with the
    lowest index, (iii) 64 sampled queries agree with the oracle on the whole set."""
    rng = np.random.default_rng(123)
    n2, n1 = 2_000_000, 512
    d2 = synth.rand_desc(rng, n2)
    d1 = synth.rand_desc(rng, n1)
    d1[:100] = d2[rng.integers(1000, n2, 100)]
    d2[7] = d1[3]
    d2[900_001] = d1[3]
    full = M.knn2(d1, d2)
    h = n2 // 2 + 3
    lo, hi = M.knn2(d1, d2[:h]), M.knn2(d1, d2[h:], idx_base=h)
    merged = np.sort(np.concatenate([lo, hi], 1), 1)[:, :2]
    assert (full == merged).all()
    assert (full[:100, 0] >> np.uint64(32) == 0).all()
    assert full[3, 0] == np.uint64(7) and (full[3, 1] >> np.uint64(32)) == 0 and (full[3, 1] & np.uint64(0xFFFFFFFF)) > 7
    sample = rng.choice(n1, 64, replace=False)
    assert (full[sample] == port.knn2_packed(d1[sample], d2)).all()
