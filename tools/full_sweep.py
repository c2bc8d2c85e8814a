"""Config 5 as stated in BASELINE.json: ALL 20 000 keyframes (x 800 descriptors) matched against the whole 16 M-row
database, sharded over N GPUs (torchrun) -- 2500 steps of 8 query keyframes, 2.56e14 unique descriptor pairs.
Prints one JSON line: whole-sweep device time (max over ranks), pairs/s, and a crc32 over every match vector of the
sweep (identical at every N).  `--kfs K` restricts the sweep to the first K query keyframes."""
import argparse
import json
import os
import sys
import time
import zlib

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synthetic database / query generators)
from pl_inertial_slam_b200.database import ShardedDescriptorDB, shard_bounds  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kfs", type=int, default=bench.N_KF)
    ap.add_argument("--kf-batch", type=int, default=8)
    ap.add_argument("--nnr", type=float, default=0.9)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_rows = bench.N_KF * bench.PER_KF
    lo, hi = shard_bounds(n_rows, world, rank)
    shard = torch.from_numpy(bench.gen_rows(lo // bench.PER_KF, hi // bench.PER_KF)).to(dev)
    db = ShardedDescriptorDB(n_rows=n_rows, shard=shard, device=local)
    # the query keyframes ARE the database keyframes (all-pairs)
    steps = (args.kfs + args.kf_batch - 1) // args.kf_batch
    crc, matches = 0, 0
    pinned = [torch.empty((args.kf_batch * bench.PER_KF, 32), dtype=torch.uint8).pin_memory() for _ in range(2)]
    out_host = torch.empty(args.kf_batch * bench.PER_KF, dtype=torch.int32).pin_memory()
    for s in range(2):
        db.match_nnr(torch.from_numpy(bench.gen_queries(s, args.kf_batch, bench.N_KF)).to(dev), args.nnr)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for s in range(steps):
        k0 = s * args.kf_batch
        k1 = min(args.kfs, k0 + args.kf_batch)
        q = bench.gen_rows(k0, k1)
        buf = pinned[s & 1][: len(q)]
        buf.copy_(torch.from_numpy(q))
        count, m12 = db.match_nnr(buf.to(dev, non_blocking=True), args.nnr)
        out_host[: len(q)].copy_(m12, non_blocking=False)
        crc = zlib.crc32(out_host[: len(q)].numpy().tobytes(), crc)
        matches += int(count.item())
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        pairs = float(args.kfs) * bench.PER_KF * n_rows
        print(json.dumps({"workload": "config 5 full sweep: every keyframe against the whole database (matchNNR, flat database)",
                          "n_gpus": world, "query_keyframes": args.kfs, "steps": steps, "unique_pairs": pairs,
                          "seconds_device": float(t.item()) * 1e-3, "seconds_wall": wall,
                          "pairs_per_s": pairs / (float(t.item()) * 1e-3), "matches": matches, "crc32_all_m12": crc & 0xFFFFFFFF,
                          "note": "queries = the database keyframes themselves (every row finds itself at distance 0, so matches == rows "
                                  "unless a duplicate row exists); host query upload and match-vector read-back inside the timed region"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
