"""CUDA kernels of the sharded paths: (a) G shards emulated one after another on ONE GPU through the
same DeviceOps primitives the multi-rank code uses (seeded chunked matchGrid, global-index top-2,
merge), (b) a real NCCL run of tests/dist_gpu_check.py under torchrun when >= 2 GPUs are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle
from helpers import oracle_grid, random_grid_case
from pl_inertial_slam_b200 import synth

pytestmark = pytest.mark.gpu
port = oracle.port
INT64_MIN = -(1 << 63)


@pytest.fixture(scope="module")
def ops(plm_lib):
    from pl_inertial_slam_b200.database import DeviceOps
    return DeviceOps(0)


def test_knn2_shards_merge_to_global(ops):
    from pl_inertial_slam_b200.database import shard_bounds
    rng = np.random.default_rng(9)
    db = synth.tie_stress_desc(rng, 5003)
    q = synth.tie_stress_desc(rng, 333)
    want = port.knn2_packed(q, db)
    qd = torch.from_numpy(q).cuda()
    for world in (1, 2, 5):
        parts = []
        for r in range(world):
            lo, hi = shard_bounds(len(db), world, r)
            parts.append(ops.knn2(qd, torch.from_numpy(db[lo:hi].copy()).cuda(), idx_base=lo))
        merged = ops.top2_merge(torch.stack(parts).contiguous())
        assert (merged.cpu().numpy().view(np.uint64) == want).all()


@pytest.mark.parametrize("is_lines", [False, True])
@pytest.mark.parametrize("world", [2, 3])
def test_grid_shards_emulated(ops, is_lines, world):
    from pl_inertial_slam_b200.database import GridFrame, shard_bounds
    rng = np.random.default_rng(400 + world + int(is_lines))
    n1, n2 = 5000, 300
    for tie in (False, True):
        case = random_grid_case(rng, n1, n2, is_lines=is_lines, tie=tie, win=(3, 3, 3, 3), zero_len=4 if is_lines else 0)
        frame = GridFrame(torch.from_numpy(case["d2"]).cuda(), torch.from_numpy(case["cell_start"]).cuda(),
                          torch.from_numpy(case["cell_items"]).cuda(), case["rows"], case["cols"],
                          torch.from_numpy(case["dirs2"]).cuda() if is_lines else None)
        d1 = torch.from_numpy(case["d1"]).cuda()
        co = torch.from_numpy(case["coords"]).cuda()
        for best_lr in (True, False):
            n_o, m_o = oracle_grid(port, case, 0.9, best_lr)
            spans = [shard_bounds(n1, world, r) for r in range(world)]
            m12 = [torch.full((hi - lo,), -1, dtype=torch.int32, device="cuda") for lo, hi in spans]
            cnt = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in spans]
            cms = [ops.grid_colmin(co[lo:hi].contiguous(), d1[lo:hi].contiguous(), lo, frame, case["win"], 0.9, 0.75, best_lr)
                   for lo, hi in spans]
            allcm = torch.stack(cms).to(torch.int32) & 0xFFFF
            keys = []
            for r, (lo, hi) in enumerate(spans):
                seed = allcm[:r].min(dim=0).values.to(torch.int16).contiguous() if (r > 0 and best_lr) else None
                keys.append(ops.grid_match(co[lo:hi].contiguous(), d1[lo:hi].contiguous(), lo, frame, case["win"], 0.9, 0.75,
                                           best_lr, m12[r], cnt[r], seed))
            if best_lr:
                key = ((torch.stack(keys) ^ INT64_MIN).min(dim=0).values ^ INT64_MIN).contiguous()
                m21 = ops.m21_from_keys(key)
                for r, (lo, hi) in enumerate(spans):
                    ops.cross_check(m12[r], lo, m21, cnt[r])
            got = torch.cat(m12).cpu().numpy()
            total = int(sum(int(c.item()) for c in cnt))
            assert total == n_o and (got == m_o).all(), (is_lines, world, tie, best_lr)


def test_sharded_classes_world1(ops):
    """ShardedDescriptorDB / ShardedMap without a process group (world = 1) on the GPU."""
    from pl_inertial_slam_b200.database import GridFrame, ShardedDescriptorDB, ShardedMap
    rng = np.random.default_rng(31)
    db = synth.rand_desc(rng, 40000)
    q = synth.flip_bits(rng, db[rng.choice(40000, 500)], 0.08)
    sdb = ShardedDescriptorDB(rows=db, ops=ops)
    count, m12 = sdb.match_nnr(torch.from_numpy(q).cuda(), 0.9)
    n_o, m_o = port.match_nnr(q, db, 0.9)
    assert int(count.item()) == n_o and (m12.cpu().numpy() == m_o).all()

    case = random_grid_case(rng, 20000, 600, win=(3, 3, 3, 3))
    frame = GridFrame(torch.from_numpy(case["d2"]).cuda(), torch.from_numpy(case["cell_start"]).cuda(),
                      torch.from_numpy(case["cell_items"]).cuda(), case["rows"], case["cols"])
    smap = ShardedMap(20000, torch.from_numpy(case["d1"]).cuda(), torch.from_numpy(case["coords"]).cuda(), ops=ops)
    count, m12 = smap.match_grid(frame, case["win"], 0.9)
    n_o, m_o = oracle_grid(port, case, 0.9, 1)
    assert int(count.item()) == n_o and (m12.cpu().numpy() == m_o).all()
    # fallback: match() on the vector matchGrid just filled (mapHandler.cpp:645-650)
    count2, m12b = smap.match(frame.d2, 0.9, True, m12_inout=m12)
    n_o2, m_o2 = port.match(case["d1"], case["d2"], 0.9, True, m12=m_o)
    assert int(count2.item()) == n_o2 and (m12b.cpu().numpy() == m_o2).all()


@pytest.mark.parametrize("is_lines", [False, True])
def test_dev_match_grid_one_call(ops, is_lines):
    """plm_dev_match_grid (ShardedMap.match_grid at world 1): the whole device-resident matchGrid as one C call --
    fresh vector initialised inside the first pass, or an in/out vector with stale entries; both ratio modes; map-sized,
    frame-sized and chunk-kernel (wide frame) shapes; an empty frame."""
    from pl_inertial_slam_b200.database import GridFrame, ShardedMap
    shapes = [(30000, 600), (700, 90), (3, 5), (5000, 20000)] if not is_lines else [(12000, 200), (650, 70)]
    for n1, n2 in shapes:
        rng = np.random.default_rng(n1 + n2 + int(is_lines))
        case = random_grid_case(rng, n1, n2, is_lines=is_lines, win=(3, 3, 3, 3) if n2 < 10000 else (1, 1, 1, 1), zero_len=3 if is_lines else 0)
        frame = GridFrame(torch.from_numpy(case["d2"]).cuda(), torch.from_numpy(case["cell_start"]).cuda(),
                          torch.from_numpy(case["cell_items"]).cuda(), case["rows"], case["cols"],
                          torch.from_numpy(case["dirs2"]).cuda() if is_lines else None)
        smap = ShardedMap(n1, torch.from_numpy(case["d1"]).cuda(), torch.from_numpy(case["coords"]).cuda(), ops=ops)
        stale = np.full(n1, -1, np.int32)
        stale[::11] = rng.integers(0, n2, len(stale[::11]))
        for best_lr in (True, False):
            for m_in in (None, stale):
                count, m12 = smap.match_grid(frame, case["win"], 0.9, 0.75, best_lr,
                                             m12_inout=None if m_in is None else torch.from_numpy(m_in).cuda())
                n_o, m_o = oracle_grid(port, case, 0.9, best_lr, m12=m_in)
                assert int(count.item()) == n_o and (m12.cpu().numpy() == m_o).all(), (is_lines, n1, n2, best_lr, m_in is None)
    if not is_lines:  # a frame without features: nothing matches, a fresh vector is all -1
        case = random_grid_case(np.random.default_rng(1), 400, 0, win=(3, 3, 3, 3))
        frame = GridFrame(torch.zeros((0, 32), dtype=torch.uint8).cuda(), torch.from_numpy(case["cell_start"]).cuda(),
                          torch.zeros(1, dtype=torch.int32).cuda(), case["rows"], case["cols"])
        smap = ShardedMap(400, torch.from_numpy(case["d1"]).cuda(), torch.from_numpy(case["coords"]).cuda(), ops=ops)
        count, m12 = smap.match_grid(frame, case["win"], 0.9)
        assert int(count.item()) == 0 and (m12.cpu().numpy() == -1).all()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_nccl_two_ranks():
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dist_gpu_check.py")
    n = min(torch.cuda.device_count(), 8)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", script],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST_GPU_CHECK_OK" in res.stdout


@pytest.mark.parametrize("world", [2, 4])
def test_peer_exchange_kernel_one_gpu(plm_lib, world):
    """The peer-memory kernels (top-2 exchange + merge, the two element-wise reductions, the all-gather) with `world`
    ranks emulated on ONE GPU.  Kernels that wait for each other must not be separate launches on one GPU, so the
    ranks' calls are recorded (plm_peer_emulate_begin) and executed as ONE cooperative launch with blockIdx.y = rank
    (plm_peer_emulate_run): every CTA that waits is co-resident with the CTAs it waits for.  Buffers are plain
    device allocations instead of peer mappings; the device code is the one the multi-GPU launches run."""
    import ctypes as C
    from pl_inertial_slam_b200 import _lib as L
    from pl_inertial_slam_b200.database import DeviceOps, shard_bounds
    rng = np.random.default_rng(77 + world)
    db = synth.tie_stress_desc(rng, 9001)
    q_cap = 1500
    ops = DeviceOps(0)
    ctx = ops.ctx.handle
    ops._bind_stream()
    bufs = (C.c_void_p * world)()
    for r in range(world):
        p, h = C.c_void_p(), (C.c_uint8 * 64)()
        L.check(plm_lib.plm_peer_alloc(ctx, world, q_cap, C.byref(p), h), "plm_peer_alloc")
        bufs[r] = p
    shards = [torch.from_numpy(db[slice(*shard_bounds(len(db), world, r))].copy()).cuda() for r in range(world)]
    err = torch.zeros(world, dtype=torch.int32, device="cuda")

    def run_all(record):
        L.check(plm_lib.plm_peer_emulate_begin(world), "plm_peer_emulate_begin")
        for r in range(world):
            record(r)
        L.check(plm_lib.plm_peer_emulate_run(ctx), "plm_peer_emulate_run")
        torch.cuda.synchronize()
        assert not err.any().item(), "a rank timed out"

    epoch = 0
    for n1 in (1, 128, 129, 1500, 700, 33, 1024):            # several epochs: both parities, ragged last block
        q = synth.tie_stress_desc(rng, n1)
        qd = torch.from_numpy(q).cuda()
        want = port.knn2_packed(q, db)
        n_o, m_o = port.match_nnr(q, db, 0.9)
        epoch += 1
        locals_ = [ops.knn2(qd, shards[r], idx_base=shard_bounds(len(db), world, r)[0]) for r in range(world)]
        outs = [torch.empty((n1, 2), dtype=torch.int64, device="cuda") for _ in range(world)]
        m12s = [torch.full((n1,), -1, dtype=torch.int32, device="cuda") for _ in range(world)]
        cnts = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(world)]
        run_all(lambda r: L.check(plm_lib.plm_dev_top2_exchange(ctx, bufs, r, world, q_cap, epoch, C.c_void_p(locals_[r].data_ptr()), n1,
                                                                C.c_void_p(outs[r].data_ptr()), C.c_float(0.9), C.c_void_p(m12s[r].data_ptr()),
                                                                C.c_void_p(cnts[r].data_ptr()), C.c_void_p(err[r:].data_ptr())),
                                  "plm_dev_top2_exchange"))
        for r in range(world):
            assert (outs[r].cpu().numpy().view(np.uint64) == want).all(), (n1, r)
            assert int(cnts[r].item()) == n_o and (m12s[r].cpu().numpy() == m_o).all(), (n1, r)
    # element-wise reductions through the same buffers (the two exchanges of the row-sharded matchGrid)
    for n in (1, 7, 600, 1333):
        u16 = rng.integers(0, 300, (world, n)).astype(np.uint16)
        u16[rng.random((world, n)) < 0.3] = 0xFFFF
        u64 = rng.integers(0, 1 << 62, (world, n), dtype=np.uint64)
        u64[rng.random((world, n)) < 0.2] = np.uint64(0xFFFFFFFFFFFFFFFF)
        for op, host, pad_to, fill in ((1, u16, 8, 0xFFFF), (0, u64, 2, 0xFFFFFFFFFFFFFFFF)):
            npad = (n + pad_to - 1) // pad_to * pad_to
            padded = np.full((world, npad), fill, host.dtype)
            padded[:, :n] = host
            src = [torch.from_numpy(padded[r].view(np.int16 if op == 1 else np.int64).copy()).cuda() for r in range(world)]
            outs = [torch.empty_like(x) for x in src]
            epoch += 1
            run_all(lambda r: L.check(plm_lib.plm_dev_peer_reduce(ctx, bufs, r, world, q_cap, epoch, op, C.c_void_p(src[r].data_ptr()),
                                                                  npad // pad_to, C.c_void_p(outs[r].data_ptr()), C.c_void_p(err[r:].data_ptr())),
                                      "plm_dev_peer_reduce"))
            for r in range(world):
                got = outs[r].cpu().numpy().view(host.dtype)[:n]
                if op == 0:
                    want = host.min(axis=0)
                else:
                    want = host[:r].min(axis=0) if r > 0 else np.full(n, 0xFFFF, np.uint16)
                assert np.array_equal(got, want), (op, n, r)
    # all-gather of per-rank int32 shards + sum of per-rank counts through a second set of buffers
    n_cap = 50_000
    gbufs = (C.c_void_p * world)()
    for r in range(world):
        p, h = C.c_void_p(), (C.c_uint8 * 64)()
        L.check(plm_lib.plm_peer_alloc_bytes(ctx, plm_lib.plm_peer_gather_bytes(world, n_cap), C.byref(p), h), "plm_peer_alloc_bytes")
        gbufs[r] = p
    g_epoch = 0
    for n_rows in (5000, 1, 4097, 37, 50_000, 5000):
        full = rng.integers(-1, 1000, n_rows).astype(np.int32)
        counts = rng.integers(-50, 500, world).astype(np.int32)
        spans = [shard_bounds(n_rows, world, r) for r in range(world)]
        loc = [torch.from_numpy(full[lo:hi].copy()).cuda() for lo, hi in spans]
        cnt = [torch.from_numpy(counts[r:r + 1].copy()).cuda() for r in range(world)]
        outs = [torch.full((n_rows,), -7, dtype=torch.int32, device="cuda") for _ in range(world)]
        tot = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in range(world)]
        g_epoch += 1
        run_all(lambda r: L.check(plm_lib.plm_dev_peer_allgather_i32(ctx, gbufs, r, world, n_cap, g_epoch, C.c_void_p(loc[r].data_ptr()),
                                                                     spans[r][0], spans[r][1] - spans[r][0], n_rows, C.c_void_p(cnt[r].data_ptr()),
                                                                     C.c_void_p(outs[r].data_ptr()), C.c_void_p(tot[r].data_ptr()),
                                                                     C.c_void_p(err[r:].data_ptr())), "plm_dev_peer_allgather_i32"))
        for r in range(world):
            assert np.array_equal(outs[r].cpu().numpy(), full), (n_rows, r)
            assert int(tot[r].item()) == int(counts.sum()), (n_rows, r)
    for r in range(world):
        plm_lib.plm_peer_free(ctx, gbufs[r])
        plm_lib.plm_peer_free(ctx, bufs[r])


def test_keyframe_db_mode_a(plm_lib):
    """Per-keyframe-pair loop-closure matching against a resident database (plm_batch_set_match_dev): StVO::match of
    one query keyframe against every keyframe, each pair with its own mutual check, against the restatement."""
    from pl_inertial_slam_b200.database import KeyframeDB
    rng = np.random.default_rng(31)
    sizes = rng.integers(30, 90, 40)
    sizes[3], sizes[17] = 1, 0                                  # the reference would be UB / throw: INT32_MIN
    kf_start = np.concatenate([[0], np.cumsum(sizes)])
    rows = synth.rand_desc(rng, int(kf_start[-1]))
    db = KeyframeDB(rows, kf_start, device=0, q_cap=128)
    for trial, nq in enumerate((64, 64, 90, 2)):
        q = synth.rand_desc(rng, nq)
        k = 5 + trial                                           # the query revisits keyframe k
        m = min(nq, sizes[k])
        q[:m] = synth.flip_bits(rng, rows[kf_start[k]:kf_start[k] + m], 0.05)
        for best_lr in (True, False):
            counts, m12 = db.match_all(q, 0.9, best_lr, want_matches=True)
            for j in range(len(sizes)):
                d2 = rows[kf_start[j]:kf_start[j + 1]]
                if len(d2) < 2 or (best_lr and nq < 2):
                    assert counts[j] == np.iinfo(np.int32).min
                    continue
                n_o, m_o = port.match(q, d2, np.float32(0.9), best_lr)
                assert counts[j] == n_o and (m12[j] == m_o).all(), (trial, best_lr, j)
            c2, _ = db.match_all(q, 0.9, best_lr)
            assert (c2 == counts).all()
            if m >= 20:
                assert counts[k] == counts[counts > np.iinfo(np.int32).min].max()   # the revisited keyframe wins
