"""torchrun target: the sharded database / map paths over NCCL on real GPUs, checked against the
oracle on every rank.  `python -m torch.distributed.run --nproc-per-node N tests/dist_gpu_check.py`"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
from helpers import oracle_grid, random_grid_case  # noqa: E402
from pl_inertial_slam_b200 import synth  # noqa: E402
from pl_inertial_slam_b200.database import GridFrame, ShardedDescriptorDB, ShardedMap, shard_bounds  # noqa: E402


def run_checks(rank: int, world: int, local: int, dev, compact: bool = False) -> dict:
    """Every sharded path against the oracle on this rank (the process group must be initialised).  Raises
    AssertionError on the first mismatch; returns a summary.  compact=True (bench.py's pre-timing check) runs fewer
    exchange epochs and smaller maps."""
    port = oracle.port
    rng = np.random.default_rng(2026)
    checks = []

    # flat database (config 5 shape, small)
    db = synth.tie_stress_desc(rng, 30011)
    q = synth.tie_stress_desc(rng, 801)
    lo, hi = shard_bounds(len(db), world, rank)
    sdb = ShardedDescriptorDB(n_rows=len(db), shard=torch.from_numpy(db[lo:hi].copy()).to(dev), device=local)
    got = sdb.knn2(torch.from_numpy(q).to(dev)).cpu().numpy().view(np.uint64)
    assert (got == port.knn2_packed(q, db)).all(), "sharded knn2"
    count, m12 = sdb.match_nnr(torch.from_numpy(q).to(dev), 0.9)
    n_o, m_o = port.match_nnr(q, db, 0.9)
    assert int(count.item()) == n_o and (m12.cpu().numpy() == m_o).all(), "sharded matchNNR"
    # the same through both exchange forms, many epochs in a row (double-buffered peer slots, ragged sizes)
    mode = os.environ.get("PLM_EXPECT_EXCHANGE", "")
    if mode == "peer":
        assert sdb.peer is not None, "peer-memory exchange unavailable"
    nccl_db = ShardedDescriptorDB(n_rows=len(db), shard=sdb.shard, device=local, ops=sdb.ops, exchange="nccl")
    assert nccl_db.peer is None
    for n1 in ((1, 129, 801, 640, 5) if compact else (1, 127, 128, 129, 801, 640, 5, 800, 333, 802)):
        qq = synth.tie_stress_desc(rng, n1)
        qd = torch.from_numpy(qq).to(dev)
        want = port.knn2_packed(qq, db)
        for d in (sdb, nccl_db):
            assert (d.knn2(qd).cpu().numpy().view(np.uint64) == want).all(), ("knn2", n1, d.peer is not None)
            c, m = d.match_nnr(qd, 0.9)
            n_o, m_o = port.match_nnr(qq, db, 0.9)
            assert int(c.item()) == n_o and (m.cpu().numpy() == m_o).all(), ("matchNNR", n1, d.peer is not None)
    if sdb.peer is not None:
        sdb.peer.check()
    checks.append("flat database knn2 + matchNNR, peer-memory and NCCL exchange")

    # map -> frame (config 4 shape, scaled)
    for is_lines in (False, True):
        n1, n2 = ((24000, 600) if not is_lines else (9000, 200)) if not compact else ((12000, 600) if not is_lines else (5000, 200))
        case = random_grid_case(rng, n1, n2, is_lines=is_lines, win=(3, 3, 3, 3), zero_len=3 if is_lines else 0)
        lo, hi = shard_bounds(n1, world, rank)
        frame = GridFrame(torch.from_numpy(case["d2"]).to(dev), torch.from_numpy(case["cell_start"]).to(dev),
                          torch.from_numpy(case["cell_items"]).to(dev), case["rows"], case["cols"],
                          torch.from_numpy(case["dirs2"]).to(dev) if is_lines else None)
        smap = ShardedMap(n1, torch.from_numpy(case["d1"][lo:hi].copy()).to(dev),
                          torch.from_numpy(case["coords"][lo:hi].copy()).to(dev), ops=sdb.ops)
        for best_lr in (True, False):
            count, m12 = smap.match_grid(frame, case["win"], 0.9, 0.75, best_lr)
            n_o, m_o = oracle_grid(port, case, 0.9, best_lr)
            assert int(count.item()) == n_o and (m12.cpu().numpy() == m_o).all(), ("sharded matchGrid", is_lines, best_lr)
            count2, m12b = smap.match(frame.d2, 0.9, best_lr, m12_inout=m12)
            n_o2, m_o2 = port.match(case["d1"], case["d2"], 0.9, best_lr, m12=m_o)
            assert int(count2.item()) == n_o2 and (m12b.cpu().numpy() == m_o2).all(), ("sharded match fallback", is_lines)
        checks.append(f"sharded matchGrid + match fallback, {'lines' if is_lines else 'points'} {n1} x {n2}")
    dist.barrier()
    return {"ok": True, "world": world, "exchange": "peer" if sdb.peer is not None else "nccl", "checks": checks}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = run_checks(rank, world, local, dev)
    if rank == 0:
        print("DIST_GPU_CHECK_OK world", world, "exchange", res["exchange"])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
