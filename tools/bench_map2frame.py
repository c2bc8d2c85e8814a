"""Config 4 at N GPUs (torchrun): the 200k-point / 50k-line local map row-sharded over the ranks,
matchGrid (+-3 window) then the forced match() fallback, exchanges included (the peer-memory kernels and, beside
them, the NCCL form).  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_inertial_slam_b200 import grid as G  # noqa: E402
from pl_inertial_slam_b200 import synth  # noqa: E402
from pl_inertial_slam_b200.database import GridFrame, ShardedMap, shard_bounds  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sp = synth.make_stereo_pair(synth.SEED0 + 4)
    n_map = 200_000
    d1, xy = synth.make_map_points(synth.SEED0 + 4, n_map, sp)
    c = sp.kp_l.astype(np.float64)
    cs, ci = G.csr_from_points(c[:, 0] * synth.INV_W, c[:, 1] * synth.INV_H)
    frame = GridFrame(torch.from_numpy(sp.pdesc_l).to(dev), torch.from_numpy(cs).to(dev), torch.from_numpy(ci).to(dev),
                      G.GRID_ROWS, G.GRID_COLS)
    lo, hi = shard_bounds(n_map, world, rank)
    d1_dev, xy_dev = torch.from_numpy(d1[lo:hi].copy()).to(dev), torch.from_numpy(xy[lo:hi].copy()).to(dev)
    win = np.array([3, 3, 3, 3], np.int32)
    out = {"workload": "config 4: 200k map points x 600 frame points", "n_gpus": world}
    forms = ["auto"] if world == 1 else ["auto", "nccl"]
    ops = None
    for form in forms:
        smap = ShardedMap(n_map, d1_dev, xy_dev, device=local, exchange=form, ops=ops)
        ops = smap.ops
        name = "single" if world == 1 else ("peer_memory" if smap.peer is not None else "nccl")
        out[name] = run(smap, frame, win, world, rank, dev, n_map)
        if smap.peer is not None:
            smap.peer.check()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run(smap, frame, win, world, rank, dev, n_map):

    def timed(fn, reps=20):
        for _ in range(3):
            r = fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(3_000_000)   # ~1.5 ms of GPU spin: the host enqueues ahead, the events see device time only
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), r

    g_ms, (cnt, m12) = timed(lambda: smap.match_grid(frame, win, 0.9, 0.75, True))
    f_ms, (cnt2, _) = timed(lambda: smap.match(frame.d2, 0.9, True, m12_inout=m12))
    return {"matchGrid_ms": g_ms, "matchGrid_matches": int(cnt.item()), "match_fallback_ms": f_ms,
            "match_fallback_count": int(cnt2.item()), "match_fallback_unique_pairs_per_s": n_map * 600.0 / (f_ms * 1e-3)}


if __name__ == "__main__":
    main()
