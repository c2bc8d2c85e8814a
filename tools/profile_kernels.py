"""One pass over every kernel of the library on a representative shape, for ncu captures.

    python tools/profile_kernels.py            # plain run (must exit 0 before the same line runs under ncu)
    ncu --set full --clock-control none --import-source on -o gpurun_out/prof_all python tools/profile_kernels.py

Shapes: frame-sized calls of configs 1-2 (600 ORB + 200 LBD), the map-sized matchGrid of config 4
(200 000 x 600), a 300-frame replay batch (config 3, batched stages and the device-resident frame pipeline) and one 6400 x 2 000 000 brute-force launch (config 5,
shortened so the ~40 ncu replays stay short).  Every kernel is launched once after one warm-up call.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_inertial_slam_b200 import grid as G  # noqa: E402
from pl_inertial_slam_b200 import matching as M  # noqa: E402
from pl_inertial_slam_b200 import replay, synth  # noqa: E402
from pl_inertial_slam_b200.database import DeviceOps, GridFrame, ShardedMap  # noqa: E402

quick = "--quick" in sys.argv
ctx = M.Context(0)
M.Config.minRatio12P = 0.9
prev, curr = synth.make_temporal_pair(synth.SEED0 + 2)
a = synth.stereo_points_grid_args(prev)
b = synth.stereo_lines_grid_args(prev)
ga = (a["cell_start"], a["cell_items"], a["rows"], a["cols"])
gb = (b["cell_start"], b["cell_items"], b["rows"], b["cols"])

# frame-sized host-buffer calls: cluster matchGrid (points, lines), match (both directions + merge + cross-check),
# stereo geometry gates, StVO::distance
m_p, m_l = [], []
M.matchGrid(a["xy"], a["d1"], ga, a["d2"], a["win"], m_p, ctx=ctx)
M.matchGrid(b["xyxy"], b["d1"], gb, b["d2"], b["dirs2"], b["win"], m_l, ctx=ctx)
M.match(prev.pdesc_l, curr.pdesc_l, 0.9, [], ctx=ctx)
M.matchNNR(prev.ldesc_l, curr.ldesc_l, 0.9, [], ctx=ctx)
M.stereo_filter_points(prev.kp_l, prev.kp_r, m_p, ctx=ctx)
M.stereo_filter_lines(prev.ln_l, prev.ln_r, m_l, ctx=ctx)
M.distances(prev.pdesc_l, prev.pdesc_r, ctx=ctx)
M.line_pair_filter(prev.ln_l, prev.ln_r, m_l, ctx=ctx)
# the same four matcher calls as ONE frame session: frame_fused_kernel with four jobs (one launch per frame)
with M.FrameSession(ctx):
    M.matchGrid(a["xy"], a["d1"], ga, a["d2"], a["win"], [], ctx=ctx)
    M.matchGrid(b["xyxy"], b["d1"], gb, b["d2"], b["dirs2"], b["win"], [], ctx=ctx)
    M.match(prev.pdesc_l, curr.pdesc_l, 0.9, [], ctx=ctx)
    M.match(prev.ldesc_l, curr.ldesc_l, 0.9, [], ctx=ctx)
# ... and with the one-launch form switched off: the per-call kernels (cluster matchGrid, brute-force slices + merge +
# mutual check) that frames beyond the frame-sized limits still take
from pl_inertial_slam_b200 import _lib as _L  # noqa: E402
_L.load().plm_set_option(b"frame_fused", 0)
M.matchGrid(a["xy"], a["d1"], ga, a["d2"], a["win"], [], ctx=ctx)
M.matchGrid(b["xyxy"], b["d1"], gb, b["d2"], b["dirs2"], b["win"], [], ctx=ctx)
M.match(prev.pdesc_l, curr.pdesc_l, 0.9, [], ctx=ctx)
M.matchNNR(prev.ldesc_l, curr.ldesc_l, 0.9, [], ctx=ctx)
_L.load().plm_set_option(b"frame_fused", 1)
# local-map selection + reprojection gates of matchMap2KF* (map_select_kernel, gather_rows_kernel, map_gate_kernel)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
try:
    from test_reproj import CAM, H, INV_H, INV_W, W, make_scene  # noqa: E402
    from pl_inertial_slam_b200 import drivers as D  # noqa: E402
    rngm = np.random.default_rng(7)
    for lines_ in (False, True):
        T_, X_, act_ = make_scene(5, 20_000 if quick else 70_000, lines_)
        view_ = D.map_view(T_, CAM, INV_W, INV_H, W, H)
        sel_, coords_, pf_ = D.mapSelect(X_, act_, view_)
        m12_ = rngm.integers(-1, 300, len(sel_)).astype(np.int32)
        feat_ = rngm.uniform(0, 700, (300, 2)) if not lines_ else rngm.normal(0, 1, (300, 3))
        D.mapGate(pf_, m12_, feat_, 1.0, int((m12_ >= 0).sum()))
except Exception as e:  # noqa: BLE001 -- the capture of the other kernels does not depend on this block
    print("reprojection kernels skipped:", repr(e)[:200])

# replay batch: fused matchGrid kernel (one CTA per job), list kernels of the brute-force path
rp = synth.make_replay(synth.SEED0 + 3, 100 if quick else 300)
gjobs, tjobs = replay.stereo_grid_jobs(rp), replay.temporal_match_jobs(rp)
m_in = np.full(rp.n_m, -1, np.int32)
gbatch, tbatch = replay.MatchBatch(ctx), replay.MatchBatch(ctx)
gbatch.set_match_grid(rp.arena, rp.coords, rp.cell_start, rp.cell_items, rp.dirs2, 48, 64, gjobs, 0.9, 0.75, True, m_in)
tbatch.set_match(rp.arena, tjobs, 0.9, True, m_in)
gbatch.run(); tbatch.run()
gbatch.fetch(); tbatch.fetch()

# map-sized matchGrid (chunked two-pass kernels + scan + m21 + cross-check) and the brute-force fallback
dev = torch.device("cuda", 0)
ops = DeviceOps(0)
sp = synth.make_stereo_pair(synth.SEED0 + 4)
n_map = 50_000 if quick else 200_000
d1, xy = synth.make_map_points(synth.SEED0 + 4, n_map, sp)
c = sp.kp_l.astype(np.float64)
cs, ci = G.csr_from_points(c[:, 0] * synth.INV_W, c[:, 1] * synth.INV_H)
frame = GridFrame(torch.from_numpy(sp.pdesc_l).to(dev), torch.from_numpy(cs).to(dev), torch.from_numpy(ci).to(dev),
                  G.GRID_ROWS, G.GRID_COLS)
smap = ShardedMap(n_map, torch.from_numpy(d1).to(dev), torch.from_numpy(xy).to(dev), ops=ops)
win = np.array([3, 3, 3, 3], np.int32)
cnt, m12 = smap.match_grid(frame, win, 0.9, 0.75, True)
smap.match(frame.d2, 0.9, True, m12_inout=m12)

# config-5 shaped brute force (variant csa4) + slice merge + multi-GPU style top-2 merge + acceptance
gen = torch.Generator(device="cuda").manual_seed(1)
q = torch.randint(0, 256, (6400, 32), dtype=torch.uint8, device="cuda", generator=gen)
db = torch.randint(0, 256, (200_000 if quick else 2_000_000, 32), dtype=torch.uint8, device="cuda", generator=gen)
t2 = ops.knn2(q, db)
parts = torch.stack([t2, t2 + 1])
merged = ops.top2_merge(parts)
m = torch.full((6400,), -1, dtype=torch.int32, device="cuda")
cntt = torch.zeros(1, dtype=torch.int32, device="cuda")
ops.nnr_accept(merged, 0.9, m, cntt)
# device-resident frame pipeline (plm_frames_*): stereo_frame_kernel x2 (points, lines), f2f_match_kernel x2
from pl_inertial_slam_b200.frames import FrameConfig, FramePipeline, replay_frame_records  # noqa: E402
kp, ln, rec = replay_frame_records(rp)
pipe = FramePipeline(ctx)
pipe.upload(rp.arena, kp, ln, rec, FrameConfig())
pipe.run()
pipe.fetch()
# map landmarks: med_desc of 50 000 landmarks (mean 8 observations, a few long lists) -- med_desc_warp_kernel + med_desc_cta_kernel
from pl_inertial_slam_b200 import mapfeatures as MF  # noqa: E402
ldesc, ldirs, lstart = synth.make_landmark_observations(synth.SEED0 + 20, 20_000 if quick else 50_000, mean_obs=8,
                                                        long_lists=64, long_len=48)
t_ = lambda x: torch.from_numpy(x).to(dev)  # noqa: E731
n_lm = len(lstart) - 1
med_idx = torch.empty(n_lm, dtype=torch.int32, device=dev)
med_rows = torch.empty((n_lm, 32), dtype=torch.uint8, device=dev)
med_dir = torch.empty((n_lm, 3), dtype=torch.float64, device=dev)
ld, ldr, lst = t_(ldesc), t_(ldirs), t_(lstart)
torch.cuda.synchronize()
MF.dev_med_desc(ctx, ld, lst, med_idx, med_desc=med_rows, dir_obs=ldr, med_dir=med_dir)
# bag of words: DBoW2 transform of 300 keyframes (k = 10, L = 4 vocabulary) and one keyframe scored against 20 000 database vectors
from pl_inertial_slam_b200 import bow as B  # noqa: E402
fvoc = synth.make_vocabulary(synth.SEED0 + 40, k=10, L=4)
voc = B.Vocabulary.from_flat(fvoc, ctx=ctx)
n_kf = 100 if quick else 300
feats = synth.vocabulary_features(synth.SEED0 + 41, fvoc, n_kf * 600)
bows = voc.transform_batch(feats, np.arange(n_kf + 1, dtype=np.int32) * 600)
db = [bows[i % n_kf] for i in range(5_000 if quick else 20_000)]
B.score_matrix(bows[:1], db, ctx=ctx)
# peer-memory exchange kernels with a world of ONE rank (under ncu kernels are serialised, so several ranks on one GPU
# would wait for each other until the time-out): push into the own buffer, flag, wait (satisfied at once), merge
import ctypes as C  # noqa: E402
from pl_inertial_slam_b200 import _lib as L  # noqa: E402
lib = L.load()
xb, gb_, hdl = C.c_void_p(), C.c_void_p(), (C.c_uint8 * 64)()
L.check(lib.plm_peer_alloc(ctx.handle, 1, 8192, C.byref(xb), hdl), "plm_peer_alloc")
L.check(lib.plm_peer_alloc_bytes(ctx.handle, lib.plm_peer_gather_bytes(1, 200_000), C.byref(gb_), hdl), "plm_peer_alloc_bytes")
xs, gs = (C.c_void_p * 1)(xb), (C.c_void_p * 1)(gb_)
err = torch.zeros(1, dtype=torch.int32, device=dev)
pm12 = torch.full((6400,), -1, dtype=torch.int32, device=dev)
pcnt = torch.zeros(1, dtype=torch.int32, device=dev)
pout = torch.empty_like(t2)
p_ = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
L.check(lib.plm_dev_top2_exchange(ctx.handle, xs, 0, 1, 8192, 1, p_(t2), 6400, p_(pout), C.c_float(0.9), p_(pm12), p_(pcnt), p_(err)),
        "plm_dev_top2_exchange")
keys = torch.randint(0, 1 << 40, (600,), dtype=torch.int64, device=dev)
kout = torch.empty_like(keys)
L.check(lib.plm_dev_peer_reduce(ctx.handle, xs, 0, 1, 8192, 2, 0, p_(keys), 300, p_(kout), p_(err)), "plm_dev_peer_reduce")
rows = torch.randint(-1, 600, (200_000,), dtype=torch.int32, device=dev)
rout, rtot = torch.empty_like(rows), torch.zeros(1, dtype=torch.int32, device=dev)
L.check(lib.plm_dev_peer_allgather_i32(ctx.handle, gs, 0, 1, 200_000, 1, p_(rows), 0, 200_000, 200_000, p_(pcnt), p_(rout), p_(rtot),
                                       p_(err)), "plm_dev_peer_allgather_i32")
ctx.synchronize()
assert int(err.item()) == 0 and torch.equal(rout, rows)
torch.cuda.synchronize()
ctx.synchronize()
print("profile_kernels ok: launches", ctx.launch_count + ops.ctx.launch_count)
