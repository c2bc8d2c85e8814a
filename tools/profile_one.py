"""Small single-purpose drivers for ncu captures of one kernel family:  python tools/profile_one.py med|bow"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_inertial_slam_b200 import matching as M  # noqa: E402
from pl_inertial_slam_b200 import synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "med"
ctx = M.Context(0)
dev = torch.device("cuda", 0)
t_ = lambda x: torch.from_numpy(x).to(dev)  # noqa: E731
if what == "med":
    from pl_inertial_slam_b200 import mapfeatures as MF
    ldesc, ldirs, lstart = synth.make_landmark_observations(synth.SEED0 + 21, 250_000, mean_obs=8, long_lists=50, long_len=60)
    n_lm = len(lstart) - 1
    med_idx = torch.empty(n_lm, dtype=torch.int32, device=dev)
    med_rows = torch.empty((n_lm, 32), dtype=torch.uint8, device=dev)
    med_dir = torch.empty((n_lm, 3), dtype=torch.float64, device=dev)
    ld, ldr, lst = t_(ldesc), t_(ldirs), t_(lstart)
    for _ in range(2):
        MF.dev_med_desc(ctx, ld, lst, med_idx, med_desc=med_rows, dir_obs=ldr, med_dir=med_dir)
    ctx.synchronize()
elif what == "bow":
    from pl_inertial_slam_b200 import bow as B
    fvoc = synth.make_vocabulary(synth.SEED0 + 42, k=10, L=5, ragged=0.02)
    voc = B.Vocabulary.from_flat(fvoc, ctx=ctx)
    feats = synth.vocabulary_features(synth.SEED0 + 43, fvoc, 500 * 800, flip_p=0.05)
    bows = voc.transform_batch(feats, np.arange(501, dtype=np.int32) * 800)
    B.score_matrix(bows[:1], [bows[i % 500] for i in range(20000)], ctx=ctx)
    ctx.synchronize()
print("profile_one ok", what)
