"""World-size-2/3 gloo runs (CPU) of the multi-GPU orchestration in pl_inertial_slam_b200.database:
row sharding, all_gather of packed top-2 keys, lexicographic merge, seeded column minima for the
sharded matchGrid.  The per-shard arithmetic comes from the oracle (tests/cpu_backend.py); what is
under test is that G shards + exchanges reproduce the unsharded reference result bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _worker(rank, world, port_no, fn_name, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        globals()[fn_name](rank, world)
        open(os.path.join(out_dir, f"ok_{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _run(fn_name, world, tmp_path):
    port_no = 29500 + (os.getpid() + hash(fn_name) + world) % 2000
    mp.start_processes(_worker, args=(world, port_no, fn_name, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok_{r}")) for r in range(world))


def _check_flat_db(rank, world):
    import oracle
    from cpu_backend import CpuOps
    from pl_inertial_slam_b200 import synth
    from pl_inertial_slam_b200.database import ShardedDescriptorDB, shard_bounds
    rng = np.random.default_rng(123)
    n_db, nq = 1001, 77
    db_rows = synth.tie_stress_desc(rng, n_db)          # heavy ties across shard boundaries
    q = synth.tie_stress_desc(rng, nq)
    lo, hi = shard_bounds(n_db, world, rank)
    db = ShardedDescriptorDB(n_rows=n_db, shard=torch.from_numpy(db_rows[lo:hi].copy()), ops=CpuOps())
    got = db.knn2(torch.from_numpy(q)).numpy().view(np.uint64)
    want = oracle.port.knn2_packed(q, db_rows)
    assert (got == want).all()
    count, m12 = db.match_nnr(torch.from_numpy(q), 0.9)
    n_o, m_o = oracle.port.match_nnr(q, db_rows, 0.9)
    assert int(count) == n_o and (m12.numpy() == m_o).all()


def _check_map_match(rank, world):
    import oracle
    from cpu_backend import CpuOps
    from pl_inertial_slam_b200 import synth
    from pl_inertial_slam_b200.database import ShardedMap, shard_bounds
    rng = np.random.default_rng(321)
    n_map, n_f = 503, 61
    d2 = synth.rand_desc(rng, n_f)
    d1 = synth.rand_desc(rng, n_map)
    d1[rng.choice(n_map, 40, replace=False)] = synth.flip_bits(rng, d2[rng.choice(n_f, 40)], 0.06)
    stale = np.full(n_map, -1, np.int32)
    stale[rng.choice(n_map, 60, replace=False)] = rng.integers(0, n_f, 60)
    lo, hi = shard_bounds(n_map, world, rank)
    smap = ShardedMap(n_map, torch.from_numpy(d1[lo:hi].copy()), ops=CpuOps())
    for best_lr in (True, False):
        for m_in in (None, stale):
            count, m12 = smap.match(torch.from_numpy(d2), 0.9, best_lr,
                                    None if m_in is None else torch.from_numpy(m_in.copy()))
            n_o, m_o = oracle.port.match(d1, d2, 0.9, best_lr, m12=m_in)
            assert int(count) == n_o and (m12.numpy() == m_o).all()


def _check_map_grid(rank, world):
    import oracle
    from cpu_backend import CpuOps
    from helpers import oracle_grid, random_grid_case
    from pl_inertial_slam_b200.database import GridFrame, ShardedMap, shard_bounds
    for is_lines in (False, True):
        for tie in (False, True):
            rng = np.random.default_rng(55 + int(is_lines) * 2 + int(tie))
            n1, n2 = 700, 90
            case = random_grid_case(rng, n1, n2, is_lines=is_lines, tie=tie, win=(3, 3, 3, 3), zero_len=3 if is_lines else 0)
            lo, hi = shard_bounds(n1, world, rank)
            frame = GridFrame(torch.from_numpy(case["d2"]), torch.from_numpy(case["cell_start"]),
                              torch.from_numpy(case["cell_items"]), case["rows"], case["cols"],
                              torch.from_numpy(case["dirs2"]) if is_lines else None)
            smap = ShardedMap(n1, torch.from_numpy(case["d1"][lo:hi].copy()),
                              torch.from_numpy(case["coords"][lo:hi].copy()), ops=CpuOps())
            for best_lr in (True, False):
                count, m12 = smap.match_grid(frame, case["win"], 0.9, 0.75, best_lr)
                n_o, m_o = oracle_grid(oracle.port, case, 0.9, best_lr)
                assert int(count) == n_o, (is_lines, tie, best_lr, int(count), n_o)
                assert (m12.numpy() == m_o).all()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_flat_db_gloo(world, tmp_path):
    _run("_check_flat_db", world, tmp_path)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_map_match_gloo(world, tmp_path):
    _run("_check_map_match", world, tmp_path)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_map_grid_gloo(world, tmp_path):
    _run("_check_map_grid", world, tmp_path)


def test_shard_bounds_cover_exactly():
    from pl_inertial_slam_b200.database import shard_bounds
    for n in (0, 1, 7, 800, 16_000_000):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
