"""Bag-of-words loop-candidate scoring (the reference's vendored DBoW2; src/mapHandler.cpp:3116-3237).

CPU: the C restatement against golden vectors produced by the reference's own DBoW2 (vocabularies from its
create(), transform(), score()), and against the compiled reference on synthetic trees when it is built.
GPU (-m gpu): plm_voc_* / plm_bow_* through the C ABI, bit for bit against the restatement and the goldens."""
import os

import numpy as np
import pytest

import oracle
from conftest import ROOT
from pl_inertial_slam_b200 import synth

port, rdb = oracle.port, oracle.ref_dbow
needs_ref = pytest.mark.skipif(not rdb.available(), reason="oracle/_ref/libplref_dbow.so not built")
GOLDEN = os.path.join(ROOT, "tests", "golden", "bow.npz")


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def golden_voc(z, w):
    return oracle.FlatVocabulary.from_arrays(z, f"voc{w}_")


def feature_sets(voc, seed, sizes=(300, 220, 1, 0, 640, 64)):
    sets = [synth.vocabulary_features(seed + i, voc, n) for i, n in enumerate(sizes)]
    rng = np.random.default_rng(seed)
    sets.append(synth.rand_desc(rng, 200))
    sets.append(synth.tie_stress_desc(rng, 120))       # equal distances in the descent -> first-child rule
    sets.append(np.repeat(sets[0][:3], 40, axis=0))    # one word hit many times -> repeated-addition fold
    return sets


@pytest.mark.parametrize("w", [0, 1, 2, 3])
def test_port_vs_golden(w):
    z = np.load(GOLDEN)
    fv = golden_voc(z, w)
    bows = []
    for k in range(int(z["n_sets"])):
        ids, vals = port.bow_transform(fv, z[f"set_{k}"])
        assert np.array_equal(ids, z[f"voc{w}_ids_{k}"])
        assert np.array_equal(bits(vals), bits(z[f"voc{w}_vals_{k}"]))
        bows.append((ids, vals))
    got = np.array([[port.bow_score(a, b) for b in bows] for a in bows])
    assert np.array_equal(bits(got), bits(z[f"voc{w}_scores"]))


@needs_ref
@pytest.mark.parametrize("w,k,L", [(0, 10, 3), (1, 6, 4), (2, 9, 2), (3, 4, 5)])
def test_port_vs_reference_on_synthetic_trees(w, k, L):
    voc = synth.make_vocabulary(synth.SEED0 + 40 + w, k=k, L=L, weighting=w)
    h = rdb.from_flat(voc)
    try:
        bows = []
        for d in feature_sets(voc, synth.SEED0 + 50 + w):
            a, b = rdb.transform(h, d), port.bow_transform(voc, d)
            assert np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1]))
            bows.append(b)
        for a in bows:
            for b in bows:
                assert np.float64(rdb.score(h, a, b)).view(np.uint64) == np.float64(port.bow_score(a, b)).view(np.uint64)
    finally:
        rdb.destroy(h)


def test_vocabulary_validation_needs_no_gpu(plm_lib):
    import ctypes as C
    from pl_inertial_slam_b200 import _lib as L
    voc = synth.make_vocabulary(1, k=3, L=2)
    h = C.c_void_p()

    def create(cs=voc.child_start, ci=voc.child_ids, wd=voc.node_word, weighting=0, scoring=0):
        return plm_lib.plm_voc_create(None, voc.n_nodes, cs.ctypes.data_as(L.i32p), ci.ctypes.data_as(L.i32p),
                                      voc.node_desc.ctypes.data_as(L.u8p), voc.node_weight.ctypes.data_as(L.f64p),
                                      wd.ctypes.data_as(L.i32p), weighting, scoring, C.byref(h))
    assert create(scoring=1) == L.PLM_E_UNSUPPORTED
    assert create(weighting=7) == L.PLM_E_INVALID
    bad = voc.child_ids.copy(); bad[0] = 0                       # a child that is not below its parent
    assert create(ci=bad) == L.PLM_E_INVALID
    dup = voc.child_ids.copy(); dup[1] = dup[0]                  # listed twice
    assert create(ci=dup) == L.PLM_E_INVALID
    now = voc.node_word.copy(); now[now >= 0] = -1               # leaves without a word id
    assert create(wd=now) == L.PLM_E_INVALID


# ---- GPU parity ------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("w", [0, 1, 2, 3])
def test_gpu_bow_golden(w):
    from pl_inertial_slam_b200 import bow as B
    z = np.load(GOLDEN)
    voc = B.Vocabulary.from_flat(golden_voc(z, w))
    sets = [z[f"set_{k}"] for k in range(int(z["n_sets"]))]
    start = np.concatenate([[0], np.cumsum([len(s) for s in sets])]).astype(np.int32)
    bows = voc.transform_batch(np.concatenate(sets), start)
    for k, (ids, vals) in enumerate(bows):
        assert np.array_equal(ids, z[f"voc{w}_ids_{k}"]), k
        assert np.array_equal(bits(vals), bits(z[f"voc{w}_vals_{k}"])), k
    got = B.score_matrix(bows, bows)
    assert np.array_equal(bits(got), bits(z[f"voc{w}_scores"]))
    assert voc.score(bows[0], bows[1]) == z[f"voc{w}_scores"][0, 1]


@pytest.mark.gpu
@pytest.mark.parametrize("w,k,L", [(0, 10, 3), (1, 6, 4), (2, 9, 2), (3, 4, 5), (0, 10, 4)])
def test_gpu_bow_vs_port(w, k, L):
    from pl_inertial_slam_b200 import bow as B
    fv = synth.make_vocabulary(synth.SEED0 + 40 + w + 10 * L, k=k, L=L, weighting=w)
    voc = B.Vocabulary.from_flat(fv)
    assert voc.size() == fv.n_words
    sets = feature_sets(fv, synth.SEED0 + 60 + w, sizes=(300, 220, 1, 0, 640, 64, 2500))
    start = np.concatenate([[0], np.cumsum([len(s) for s in sets])]).astype(np.int32)
    bows = voc.transform_batch(np.concatenate(sets), start)
    want = [port.bow_transform(fv, d) for d in sets]
    for (ids, vals), (wi, wv) in zip(bows, want):
        assert np.array_equal(ids, wi)
        assert np.array_equal(bits(vals), bits(wv))
    single = voc.transform(sets[0])
    assert np.array_equal(single[0], want[0][0]) and np.array_equal(bits(single[1]), bits(want[0][1]))
    got = B.score_matrix(bows, bows)
    ref = np.array([[port.bow_score(a, b) for b in want] for a in want])
    assert np.array_equal(bits(got), bits(ref))


@pytest.mark.gpu
def test_gpu_insert_kf_bow_vectors():
    """insertKFBowVectorP / L / PL over a short keyframe sequence with a culled keyframe (mapHandler.cpp:3116-3237)."""
    from pl_inertial_slam_b200 import bow as B
    fp = synth.make_vocabulary(synth.SEED0 + 70, k=8, L=3)
    fl = synth.make_vocabulary(synth.SEED0 + 71, k=6, L=3)
    vp, vl = B.Vocabulary.from_flat(fp), B.Vocabulary.from_flat(fl)
    rng = np.random.default_rng(9)
    kfs = []
    for i in range(7):
        n_p, n_l = int(rng.integers(100, 300)), int(rng.integers(30, 90))
        kfs.append(dict(p=synth.vocabulary_features(100 + i, fp, n_p), l=synth.vocabulary_features(200 + i, fl, n_l),
                        pt=rng.uniform(0, 700, (n_p, 2)), ls=rng.uniform(0, 450, (n_l, 2))))
    kfs[5]["p"] = kfs[1]["p"].copy()                  # a revisit
    for mode in ("P", "L", "PL"):
        conf = B.BowConfusion(vp, vl)
        want = np.zeros((7, 7))
        bows_p, bows_l = [], []
        for i, k in enumerate(kfs):
            kf = B.KeyFrameBow(i, k["p"], k["l"], k["pt"], k["ls"])
            getattr(conf, "insertKFBowVector" + mode)(kf)
            if i == 2:
                conf.map_keyframes[2] = None          # culled keyframe: skipped from now on
            bows_p.append(port.bow_transform(fp, k["p"]))
            bows_l.append(port.bow_transform(fl, k["l"]))
            for j in range(i + 1):
                if j < i and (j == 2):
                    continue
                sp, sl = port.bow_score(bows_p[i], bows_p[j]), port.bow_score(bows_l[i], bows_l[j])
                if mode == "P":
                    s = sp
                elif mode == "L":
                    s = sl
                else:
                    n_pt, n_ls = len(k["pt"]), len(k["ls"])
                    std_pt = B.vector_stdv(k["pt"][:, 0]) + B.vector_stdv(k["pt"][:, 1])
                    std_ls = B.vector_stdv(k["ls"][:, 0]) + B.vector_stdv(k["ls"][:, 1])
                    s = 0.0
                    s += (sp * n_pt + sl * n_ls) / (n_pt + n_ls)
                    s += (sp * std_pt + sl * std_ls) / (std_ls + std_pt)
                want[i, j] = want[j, i] = s
        assert np.array_equal(bits(conf.conf_matrix), bits(want)), mode
        if mode == "P":
            assert conf.conf_matrix[5, 1] > 0.99      # the revisit scores like the keyframe itself


@pytest.mark.gpu
def test_gpu_bow_limits_and_large_sets():
    """The largest supported set (8192 features: widest shared-memory sort, hash table of 16384 slots), one feature
    more (rejected), a vocabulary whose bitmap does not fit next to a long query (hash table only) and the
    no-common-word / identical-vector scores."""
    from pl_inertial_slam_b200 import _lib as L
    from pl_inertial_slam_b200 import bow as B
    fv = synth.make_vocabulary(synth.SEED0 + 80, k=10, L=4, ragged=0.0, stop_frac=0.0)
    voc = B.Vocabulary.from_flat(fv)
    big = synth.vocabulary_features(synth.SEED0 + 81, fv, 8192, flip_p=0.04)
    small = synth.vocabulary_features(synth.SEED0 + 82, fv, 50, flip_p=0.04)
    bows = voc.transform_batch(np.concatenate([big, small]), [0, 8192, 8242])
    want = [port.bow_transform(fv, big), port.bow_transform(fv, small)]
    for (i, v), (wi, wv) in zip(bows, want):
        assert np.array_equal(i, wi) and np.array_equal(bits(v), bits(wv))
    got = B.score_matrix(bows, bows)
    ref = np.array([[port.bow_score(a, b) for b in want] for a in want])
    assert np.array_equal(bits(got), bits(ref))
    with pytest.raises(L.PlmError) as e:
        voc.transform(np.concatenate([big, small[:1]]))
    assert e.value.status == L.PLM_E_UNSUPPORTED
    # word ids far beyond any bitmap: synthetic vectors with ids up to 2^31 - 2
    rng = np.random.default_rng(3)
    ids_a = np.unique(rng.integers(0, 2**31 - 1, 3000)).astype(np.uint32)
    ids_b = np.unique(np.concatenate([ids_a[::7], rng.integers(0, 2**31 - 1, 2000).astype(np.uint32)]))
    va, vb = rng.random(len(ids_a)), rng.random(len(ids_b))
    va /= va.sum(); vb /= vb.sum()
    a, b, empty = (ids_a, va), (ids_b, vb), (np.zeros(0, np.uint32), np.zeros(0))
    got = B.score_matrix([a, b, empty], [a, b, empty])
    ref = np.array([[port.bow_score(x, y) for y in (a, b, empty)] for x in (a, b, empty)])
    assert np.array_equal(bits(got), bits(ref))
