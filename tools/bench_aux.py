"""Side kernels only (landmark medoids, BoW transform / score): python tools/bench_aux.py [--cpu] [--events]

--events: additionally times the scoring / medoid calls with CUDA events on a torch stream the context is bound to,
next to the host time spent enqueueing (separates kernel time from launch overhead)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench_extras  # noqa: E402
from pl_inertial_slam_b200.matching import Context  # noqa: E402


def events(ctx):
    import ctypes as C
    import numpy as np
    from pl_inertial_slam_b200 import _lib as L
    from pl_inertial_slam_b200 import bow as B
    from pl_inertial_slam_b200 import mapfeatures as MF
    from pl_inertial_slam_b200 import synth
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    out = {}
    with torch.cuda.stream(stream):
        fvoc = synth.make_vocabulary(synth.SEED0 + 42, k=10, L=5, ragged=0.02)
        voc = B.Vocabulary.from_flat(fvoc, ctx=ctx)
        n_kf, per = 2000, 800
        feats = synth.vocabulary_features(synth.SEED0 + 43, fvoc, n_kf * per, flip_p=0.05)
        d_desc = torch.from_numpy(feats).to(dev)
        d_start = torch.arange(0, (n_kf + 1) * per, per, dtype=torch.int32, device=dev)
        ids = torch.zeros(n_kf * per, dtype=torch.int32, device=dev)
        vals = torch.zeros(n_kf * per, dtype=torch.float64, device=dev)
        lens = torch.zeros(n_kf, dtype=torch.int32, device=dev)
        lib = L.load()
        ptr = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
        L.check(lib.plm_dev_bow_transform(voc._h, ptr(d_desc), n_kf * per, ptr(d_start), n_kf, per, ptr(ids), ptr(vals), ptr(lens)), "t")
        n_db, rep = 20_000, 10
        db_ids, db_vals = ids.repeat(rep), vals.repeat(rep)
        db_start = torch.arange(n_db, device=dev, dtype=torch.int64) * per
        db_len = lens.repeat(rep).contiguous()
        q_start = torch.zeros(1, dtype=torch.int64, device=dev)
        q_len = lens[:1].contiguous()
        scores = torch.zeros(n_db, dtype=torch.float64, device=dev)

        def score():
            L.check(lib.plm_dev_bow_score(ctx.handle, ptr(ids), ptr(vals), ptr(q_start), ptr(q_len), 1, per, fvoc.n_words,
                                          ptr(db_ids), ptr(db_vals), ptr(db_start), ptr(db_len), n_db, ptr(scores)), "s")

        def timed(fn, reps=20):
            fn(); stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            t1 = time.perf_counter()
            e1.record(stream); stream.synchronize()
            return {"device_ms": e0.elapsed_time(e1) / reps, "host_enqueue_ms": (t1 - t0) / reps * 1e3}
        out["bow_score"] = timed(score)
        # the same without the 10 copies of the query itself in the database (782 common words each: the long serial sums)
        db_len_full = db_len
        db_len = db_len_full.clone()
        db_len[::n_kf] = 0
        out["bow_score_without_self_copies"] = timed(score)
        db_len = db_len_full
        n_db_full = n_db
        n_db = 64
        out["bow_score_64_vectors"] = timed(score)
        db_len = db_len_full.clone()
        db_len[::n_kf] = 0
        out["bow_score_64_vectors_without_self"] = timed(score)
        n_db = 4736
        out["bow_score_4736_vectors_without_self"] = timed(score)
        db_len = db_len_full
        n_db = n_db_full
        common = [int(np.isin(ids[:int(lens[0])].cpu().numpy(), ids[k * per:k * per + int(lens[k])].cpu().numpy()).sum()) for k in (0, 1, 2, 3)]
        out["bow_score"]["common_words_with_kf_0_1_2_3"] = common
        desc, dirs, start = synth.make_landmark_observations(synth.SEED0 + 21, 250_000, mean_obs=8, long_lists=50, long_len=60)
        t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
        d_d, d_dirs, d_st = t(desc), t(dirs), t(start)
        med_idx = torch.empty(250_000, dtype=torch.int32, device=dev)
        med_rows = torch.empty((250_000, 32), dtype=torch.uint8, device=dev)
        med_dir = torch.empty((250_000, 3), dtype=torch.float64, device=dev)
        out["med_desc"] = timed(lambda: MF.dev_med_desc(ctx, d_d, d_st, med_idx, med_desc=med_rows, dir_obs=d_dirs, med_dir=med_dir))
    ctx.set_stream(None)
    return out


if __name__ == "__main__":
    cpu = "--cpu" in sys.argv
    torch.cuda.set_device(0)
    ctx = Context()
    if "--events" in sys.argv:
        print(json.dumps(events(ctx)))
    else:
        out = {"map_landmarks": bench_extras.map_landmarks(ctx, cpu=cpu), "bow_scoring": bench_extras.bow_scoring(ctx, cpu=cpu)}
        print(json.dumps(out))
