"""torchrun target: latency of the per-query top-2 exchange + merge + ratio test across GPUs, the library's
peer-memory kernel (plm_dev_top2_exchange) next to NCCL all_gather + merge kernel + acceptance kernel.

    python -m torch.distributed.run --nproc-per-node N tools/bench_exchange.py
Prints one JSON line on rank 0 (device time per exchange, CUDA events, max over ranks)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_inertial_slam_b200.database import DeviceOps, PeerExchange, _Group  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g, ops = _Group(), DeviceOps(local)
    peer = PeerExchange.create(g, ops, 8192)
    out = {"world": world, "peer_available": peer is not None, "unit": "us per exchange (device time, max over ranks)"}
    gen = torch.Generator(device=dev).manual_seed(rank)
    for n1 in (200, 600, 6400):
        local_keys = torch.randint(0, 1 << 40, (n1, 2), dtype=torch.int64, device=dev, generator=gen).sort(dim=1).values
        m12 = torch.full((n1,), -1, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)

        def nccl():
            top2 = ops.top2_merge(g.all_gather(local_keys))
            ops.nnr_accept(top2, 0.9, m12, cnt)
            return top2

        def p2p():
            return peer.exchange(local_keys, nnr=0.9, m12=m12, count=cnt, out=buf)

        buf = torch.empty((n1, 2), dtype=torch.int64, device=dev)
        res = {}
        for name, fn in (("nccl_all_gather_merge_accept", nccl),) + ((("peer_memory_kernel", p2p),) if peer else ()):
            for _ in range(20):
                ref = fn()
            torch.cuda.synchronize(); dist.barrier()
            reps = 200
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[name] = float(t.item())
            res[name + "_checksum"] = int(ref.sum().item())
        if peer:
            assert res["nccl_all_gather_merge_accept_checksum"] == res["peer_memory_kernel_checksum"], "exchange forms disagree"
            peer.check()
        out[f"q{n1}"] = {k: v for k, v in res.items() if not k.endswith("_checksum")}
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
