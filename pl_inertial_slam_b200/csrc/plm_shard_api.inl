// Multi-device keyframe database / local map inside ONE process (plm_shard_*; included by plmatch.cu).
//
// The reference's host (MapHandler: matchMap2KFPoints/Lines src/mapHandler.cpp:583-803, isLoopClosure :3301-3409) is
// one C++ process; this is the entry point it can call to spread the map / keyframe database over the GPUs of a box
// without any launcher: the library owns the per-device contexts, enables peer access between every pair of
// devices (no CUDA IPC needed inside one process), allocates the exchange buffers and runs the same peer-memory
// kernels (csrc/plm_peer.cuh) the process-per-GPU path uses.  Rows are sharded contiguously; row indices in every
// result are GLOBAL, so results are bit-identical to the single-GPU calls.
//
// Ordering rule that keeps the bounded spins of the peer kernels from ever seeing a stalled peer: a call first does
// everything that may allocate or synchronise on every device (phase "prepare": scratch, copies, local kernels),
// and only then enqueues the kernels that wait for each other, device after device, with no host-side
// synchronisation in between.

struct plm_shard {
    int n = 0;
    std::vector<int> devices;
    std::vector<plm_ctx *> ctx;
    std::vector<uint4 *> rows;      // this device's rows [lo, hi)
    std::vector<int32_t *> coords;  // grid-cell coordinates of those rows (config 4), may be null
    std::vector<int64_t> lo, hi;
    std::vector<char *> scratch;    // per device: queries / frame side, local keys, m12 slices, outputs
    std::vector<size_t> scratch_cap;
    std::vector<void *> xchg, gather;
    int coords_per_row = 0;
    int64_t rows_cap = 0, per_cap = 0, n_rows = 0;
    int q_cap = 0;
    uint32_t xchg_epoch = 0, gather_epoch = 0;
    int32_t *error_host = nullptr;  // mapped pinned: written by the peer kernels of any device on a timeout
    bool failed = false;
    char *h_stage = nullptr;        // pinned staging of one call
    size_t h_cap = 0;

    int ensure_scratch(int g, size_t bytes) {
        if (bytes <= scratch_cap[g]) return PLM_OK;
        CU_TRY(cudaSetDevice(devices[g]));
        CU_TRY(cudaStreamSynchronize(ctx[g]->stream));
        if (scratch[g]) CU_TRY(cudaFree(scratch[g]));
        scratch[g] = nullptr;
        scratch_cap[g] = 0;
        const size_t cap = align_up(bytes + bytes / 4, 1 << 20);
        CU_TRY(cudaMalloc(reinterpret_cast<void **>(&scratch[g]), cap));
        scratch_cap[g] = cap;
        return PLM_OK;
    }
    int ensure_stage(size_t bytes) {
        if (bytes <= h_cap) return PLM_OK;
        for (int g = 0; g < n; ++g) {
            CU_TRY(cudaSetDevice(devices[g]));
            CU_TRY(cudaStreamSynchronize(ctx[g]->stream));
        }
        if (h_stage) CU_TRY(cudaFreeHost(h_stage));
        h_stage = nullptr;
        h_cap = 0;
        const size_t cap = align_up(bytes + bytes / 4, 1 << 16);
        CU_TRY(cudaHostAlloc(reinterpret_cast<void **>(&h_stage), cap, cudaHostAllocPortable));
        h_cap = cap;
        return PLM_OK;
    }
};

namespace {

int shard_check_alive(plm_shard *s) {
    if (!s) return fail(PLM_E_INVALID, "null shard set");
    if (s->failed || (s->error_host && *s->error_host != 0)) {
        s->failed = true;
        return fail(PLM_E_PEER, "a device did not arrive at a peer-memory exchange within the spin limit (option "
                                "peer_spin_ms); the exchange buffers are in an undefined state -- destroy and recreate the shard set");
    }
    return PLM_OK;
}

int shard_sync_all(plm_shard *s) {
    for (int g = 0; g < s->n; ++g) {
        CU_TRY(cudaSetDevice(s->devices[g]));
        CU_TRY(cudaStreamSynchronize(s->ctx[g]->stream));
    }
    return PLM_OK;
}

} // namespace

PLM_API int plm_shard_create(const int *devices, int n_devices, int q_cap, int64_t rows_cap, plm_shard **out) {
    if (!out) return fail(PLM_E_INVALID, "null out");
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > PLM_PEER_MAX_RANKS) return fail(PLM_E_INVALID, "1 .. 16 devices");
    if (q_cap < 1 || rows_cap < 1 || rows_cap > (1ll << 32)) return fail(PLM_E_INVALID, "q_cap >= 1, 1 <= rows_cap <= 2^32");
    for (int a = 0; a < n_devices; ++a)
        for (int b = a + 1; b < n_devices; ++b)
            if (devices[a] == devices[b])
                return fail(PLM_E_INVALID, "every shard needs its own device (kernels that wait for each other must not share a GPU)");
    int prev = 0;
    cudaGetDevice(&prev);
    plm_shard *s = new (std::nothrow) plm_shard();
    if (!s) return fail(PLM_E_NOMEM, "host allocation failed");
    s->n = n_devices;
    s->devices.assign(devices, devices + n_devices);
    s->ctx.assign(n_devices, nullptr);
    s->rows.assign(n_devices, nullptr);
    s->coords.assign(n_devices, nullptr);
    s->lo.assign(n_devices, 0);
    s->hi.assign(n_devices, 0);
    s->scratch.assign(n_devices, nullptr);
    s->scratch_cap.assign(n_devices, 0);
    s->xchg.assign(n_devices, nullptr);
    s->gather.assign(n_devices, nullptr);
    s->q_cap = q_cap;
    s->rows_cap = rows_cap;
    s->per_cap = (rows_cap + n_devices - 1) / n_devices;
    auto bail = [&](int st) {
        plm_shard_destroy(s);
        cudaSetDevice(prev);
        return st;
    };
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(&s->error_host), 64, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) return bail(fail(PLM_E_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e)));
    *s->error_host = 0;
    for (int g = 0; g < n_devices; ++g) {
        int st = plm_ctx_create(devices[g], &s->ctx[g]);
        if (st != PLM_OK) return bail(st);
        for (int p = 0; p < n_devices; ++p) {
            if (p == g) continue;
            int can = 0;
            e = cudaDeviceCanAccessPeer(&can, devices[g], devices[p]);
            if (e != cudaSuccess || !can) return bail(fail(PLM_E_UNSUPPORTED, "devices cannot access each other's memory (no NVLink / PCIe peer path)"));
            e = cudaDeviceEnablePeerAccess(devices[p], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return bail(fail(PLM_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)));
        }
        e = cudaMalloc(reinterpret_cast<void **>(&s->rows[g]), static_cast<size_t>(s->per_cap) * 32);
        const size_t xb = plm::peer_buffer_bytes(n_devices, q_cap, (q_cap + plm::PEER_THREADS - 1) / plm::PEER_THREADS);
        const size_t gb = plm::peer_gather_bytes(n_devices, rows_cap);
        if (e == cudaSuccess && n_devices > 1) e = cudaMalloc(&s->xchg[g], xb);
        if (e == cudaSuccess && n_devices > 1) e = cudaMalloc(&s->gather[g], gb);
        if (e == cudaSuccess && n_devices > 1) e = cudaMemset(s->xchg[g], 0, xb); // flags start at epoch 0 = "nothing arrived"
        if (e == cudaSuccess && n_devices > 1) e = cudaMemset(s->gather[g], 0, gb);
        if (e != cudaSuccess)
            return bail(fail(e == cudaErrorMemoryAllocation ? PLM_E_NOMEM : PLM_E_CUDA, std::string("plm_shard_create: ") + cudaGetErrorString(e)));
    }
    for (int g = 0; g < n_devices; ++g) {
        cudaSetDevice(devices[g]);
        cudaDeviceSynchronize();
    }
    cudaSetDevice(prev);
    *out = s;
    return PLM_OK;
}

PLM_API int plm_shard_destroy(plm_shard *s) {
    if (!s) return PLM_OK;
    for (int g = 0; g < s->n; ++g) {
        if (!s->ctx[g]) continue;
        cudaSetDevice(s->devices[g]);
        cudaStreamSynchronize(s->ctx[g]->stream);
    }
    for (int g = 0; g < s->n; ++g) {
        if (!s->ctx[g]) continue;
        cudaSetDevice(s->devices[g]);
        if (s->rows[g]) cudaFree(s->rows[g]);
        if (s->coords[g]) cudaFree(s->coords[g]);
        if (s->scratch[g]) cudaFree(s->scratch[g]);
        if (s->xchg[g]) cudaFree(s->xchg[g]);
        if (s->gather[g]) cudaFree(s->gather[g]);
        plm_ctx_destroy(s->ctx[g]);
    }
    if (s->h_stage) cudaFreeHost(s->h_stage);
    if (s->error_host) cudaFreeHost(s->error_host);
    delete s;
    return PLM_OK;
}

PLM_API int plm_shard_n_devices(const plm_shard *s) { return s ? s->n : 0; }
PLM_API int64_t plm_shard_n_rows(const plm_shard *s) { return s ? s->n_rows : 0; }
PLM_API int plm_shard_range(const plm_shard *s, int i, int64_t *row_lo, int64_t *row_hi) {
    if (!s || i < 0 || i >= s->n || !row_lo || !row_hi) return fail(PLM_E_INVALID, "bad shard index");
    *row_lo = s->lo[i];
    *row_hi = s->hi[i];
    return PLM_OK;
}
PLM_API uint64_t plm_shard_launch_count(const plm_shard *s) {
    uint64_t t = 0;
    if (s)
        for (int g = 0; g < s->n; ++g) t += s->ctx[g]->launches;
    return t;
}

// rows [0, n_rows) of the database / map, `step` bytes apart; shard g keeps rows [g * per, min(n_rows, (g + 1) * per)),
// per = ceil(n_rows / n_devices).  coords (may be NULL): n_rows x coords_per_row grid-cell coordinates (2 = points,
// 4 = lines) of the same rows, for plm_shard_match_grid.
PLM_API int plm_shard_upload(plm_shard *s, const uint8_t *rows, int64_t n_rows, size_t step, const int32_t *coords, int coords_per_row) {
    int st = shard_check_alive(s);
    if (st != PLM_OK) return st;
    if (n_rows < 0 || n_rows > s->rows_cap) return fail(PLM_E_INVALID, "n_rows outside [0, rows_cap]");
    if (n_rows > 0 && (!rows || step < 32)) return fail(PLM_E_INVALID, "null rows / step < 32");
    if (coords && coords_per_row != 2 && coords_per_row != 4) return fail(PLM_E_INVALID, "coords_per_row must be 2 or 4");
    int prev = 0;
    cudaGetDevice(&prev);
    const int64_t per = (n_rows + s->n - 1) / s->n;
    s->n_rows = n_rows;
    s->coords_per_row = coords ? coords_per_row : 0;
    for (int g = 0; g < s->n; ++g) {
        const int64_t lo = std::min<int64_t>(n_rows, static_cast<int64_t>(g) * per), hi = std::min<int64_t>(n_rows, lo + per);
        s->lo[g] = lo;
        s->hi[g] = hi;
        CU_TRY(cudaSetDevice(s->devices[g]));
        if (hi > lo) {
            if (step == 32) CU_TRY(cudaMemcpy(s->rows[g], rows + static_cast<size_t>(lo) * 32, static_cast<size_t>(hi - lo) * 32, cudaMemcpyHostToDevice));
            else CU_TRY(cudaMemcpy2D(s->rows[g], 32, rows + static_cast<size_t>(lo) * step, step, 32, static_cast<size_t>(hi - lo), cudaMemcpyHostToDevice));
        }
        if (coords) {
            if (!s->coords[g]) CU_TRY(cudaMalloc(reinterpret_cast<void **>(&s->coords[g]), static_cast<size_t>(s->per_cap) * 4 * sizeof(int32_t)));
            if (hi > lo)
                CU_TRY(cudaMemcpy(s->coords[g], coords + static_cast<size_t>(lo) * coords_per_row,
                                  static_cast<size_t>(hi - lo) * coords_per_row * sizeof(int32_t), cudaMemcpyHostToDevice));
        }
    }
    cudaSetDevice(prev);
    return PLM_OK;
}

// StVO::matchNNR / cv::BFMatcher::knnMatch(k = 2) of host queries against the whole sharded database (config 5, flat
// database): H2D of the queries to every device, local top-2 with global indices, ONE peer-memory kernel per device
// (push, wait, merge, fp32 ratio test), D2H from the first device.  top2 (n1 x 2 packed keys) and m12_inout /
// n_matches may each be NULL.  Batches larger than q_cap are processed in slices of q_cap queries.
PLM_API int plm_shard_match_nnr(plm_shard *s, const uint8_t *q, int n1, size_t step, float nnr, uint64_t *top2, int32_t *m12_inout,
                                int *n_matches) {
    int st = shard_check_alive(s);
    if (st != PLM_OK) return st;
    if ((st = check_desc(q, n1, step)) != PLM_OK) return st;
    if (n_matches) *n_matches = 0;
    if (m12_inout && !n_matches) return fail(PLM_E_INVALID, "null n_matches");
    if (n1 == 0) return PLM_OK;
    if (s->n_rows == 0) return fail(PLM_E_TRAIN, "matchNNR: empty train set");
    int prev = 0;
    cudaGetDevice(&prev);
    int total = 0;
    for (int q0 = 0; q0 < n1; q0 += s->q_cap) {
        const int nq = std::min(s->q_cap, n1 - q0);
        Layout H;
        const size_t h_q = H.add(size_t(nq) * 32), h_m = H.add(size_t(nq) * 4 + 4), h_t = H.add(size_t(nq) * 16);
        if ((st = s->ensure_stage(H.total)) != PLM_OK) return st;
        pack_rows(s->h_stage + h_q, q + static_cast<size_t>(q0) * step, nq, step);
        if (m12_inout) std::memcpy(s->h_stage + h_m, m12_inout + q0, size_t(nq) * 4);
        else std::memset(s->h_stage + h_m, 0xFF, size_t(nq) * 4);
        std::memset(s->h_stage + h_m + size_t(nq) * 4, 0, 4);
        Layout D;
        const size_t d_q = D.add(size_t(nq) * 32), d_m = D.add(size_t(nq) * 4 + 4), d_loc = D.add(size_t(nq) * 16), d_out = D.add(size_t(nq) * 16);
        // prepare: scratch, copies, local kernels (may allocate / synchronise per device)
        for (int g = 0; g < s->n; ++g) {
            if ((st = s->ensure_scratch(g, D.total)) != PLM_OK) return st;
            CU_TRY(cudaSetDevice(s->devices[g]));
            cudaStream_t sg = s->ctx[g]->stream;
            char *X = s->scratch[g];
            CU_TRY(cudaMemcpyAsync(X + d_q, s->h_stage + h_q, size_t(nq) * 32, cudaMemcpyHostToDevice, sg));
            CU_TRY(cudaMemcpyAsync(X + d_m, s->h_stage + h_m, size_t(nq) * 4 + 4, cudaMemcpyHostToDevice, sg));
            if (s->hi[g] > s->lo[g]) {
                if ((st = plm_dev_knn2(s->ctx[g], X + d_q, nq, s->rows[g], s->hi[g] - s->lo[g], static_cast<uint64_t>(s->lo[g]),
                                       reinterpret_cast<uint64_t *>(X + d_loc))) != PLM_OK)
                    return st;
            } else {
                CU_TRY(cudaMemsetAsync(X + d_loc, 0xFF, size_t(nq) * 16, sg)); // an empty shard contributes absent keys
            }
        }
        // exchange: nothing below allocates or synchronises until every device has its kernel enqueued
        if (s->n > 1) {
            const uint32_t epoch = ++s->xchg_epoch;
            for (int g = 0; g < s->n; ++g) {
                char *X = s->scratch[g];
                if ((st = plm_dev_top2_exchange(s->ctx[g], s->xchg.data(), g, s->n, s->q_cap, epoch, reinterpret_cast<const uint64_t *>(X + d_loc),
                                                nq, reinterpret_cast<uint64_t *>(X + d_out), nnr, m12_inout ? reinterpret_cast<int32_t *>(X + d_m) : nullptr,
                                                m12_inout ? reinterpret_cast<int32_t *>(X + d_m) + nq : nullptr, s->error_host)) != PLM_OK)
                    return st;
            }
        } else {
            char *X = s->scratch[0];
            CU_TRY(cudaMemcpyAsync(X + d_out, X + d_loc, size_t(nq) * 16, cudaMemcpyDeviceToDevice, s->ctx[0]->stream));
            if (m12_inout && (st = plm_dev_nnr_accept(s->ctx[0], reinterpret_cast<const uint64_t *>(X + d_loc), nq, nnr,
                                                      reinterpret_cast<int32_t *>(X + d_m), reinterpret_cast<int32_t *>(X + d_m) + nq)) != PLM_OK)
                return st;
        }
        CU_TRY(cudaSetDevice(s->devices[0]));
        char *X0 = s->scratch[0];
        if (m12_inout) CU_TRY(cudaMemcpyAsync(s->h_stage + h_m, X0 + d_m, size_t(nq) * 4 + 4, cudaMemcpyDeviceToHost, s->ctx[0]->stream));
        if (top2) CU_TRY(cudaMemcpyAsync(s->h_stage + h_t, X0 + d_out, size_t(nq) * 16, cudaMemcpyDeviceToHost, s->ctx[0]->stream));
        CU_TRY(cudaStreamSynchronize(s->ctx[0]->stream));
        if ((st = shard_check_alive(s)) != PLM_OK) return st;
        if (m12_inout) {
            std::memcpy(m12_inout + q0, s->h_stage + h_m, size_t(nq) * 4);
            int32_t c;
            std::memcpy(&c, s->h_stage + h_m + size_t(nq) * 4, 4);
            total += c;
        }
        if (top2) std::memcpy(top2 + 2 * static_cast<size_t>(q0), s->h_stage + h_t, size_t(nq) * 16);
    }
    if (n_matches) *n_matches = total;
    cudaSetDevice(prev);
    return PLM_OK;
}

namespace {

// Frame side of a config-4 call staged on every device.
struct ShardFrame {
    size_t d_d2 = 0, d_cs = 0, d_ci = 0, d_dir = 0, d_m12 = 0, d_cnt = 0, d_out = 0, d_tot = 0, total = 0;
};

} // namespace

// StVO::matchGrid with the sharded map as desc1 / the query side (matchMap2KFPoints / Lines, mapHandler.cpp:637-642,
// :752-757): frame side (CSR grid over the frame features, descriptors, line directions) from the host, map rows and
// their cell coordinates resident (plm_shard_upload).  m12_inout = the global in/out vector (n_rows).
PLM_API int plm_shard_match_grid(plm_shard *s, const int32_t *cell_start, const int32_t *cell_items, int grid_rows, int grid_cols,
                                 const uint8_t *d2, int n2, size_t step2, const double *dirs2, double line_sim_th, const int32_t win[4],
                                 double ratio, int best_lr, int32_t *m12_inout, int *n_matches) {
    int st = shard_check_alive(s);
    if (st != PLM_OK) return st;
    if ((st = check_desc(d2, n2, step2)) != PLM_OK) return st;
    if (!n_matches || !win) return fail(PLM_E_INVALID, "null n_matches / win");
    *n_matches = 0;
    if (s->coords_per_row == 0) return fail(PLM_E_INVALID, "no coordinates uploaded (plm_shard_upload)");
    const int is_lines = s->coords_per_row == 4;
    if (is_lines && n2 > 0 && !dirs2) return fail(PLM_E_INVALID, "null dirs2");
    if (ratio > 1.0) return fail(PLM_E_RATIO, plm_status_string(PLM_E_RATIO));
    int n_items = 0;
    if ((st = validate_grid(cell_start, cell_items, grid_rows, grid_cols, &n_items)) != PLM_OK) return st;
    if (n2 > GRID_N2_MAX) return fail(PLM_E_UNSUPPORTED, "matchGrid supports at most 32768 train features");
    const int64_t n_rows = s->n_rows;
    if (n_rows == 0) return PLM_OK;
    if (!m12_inout) return fail(PLM_E_INVALID, "null m12");
    if (n_rows > INT_MAX) return fail(PLM_E_UNSUPPORTED, "more than 2^31 map rows");
    if (s->n > 1 && ((size_t(n2) + 7) / 8 > size_t(s->q_cap))) return fail(PLM_E_UNSUPPORTED, "frame too large for the exchange buffers (q_cap)");
    int prev = 0;
    cudaGetDevice(&prev);
    const int n_cells = grid_rows * grid_cols;
    Layout H;
    const size_t h_d2 = H.add(size_t(std::max(n2, 1)) * 32), h_cs = H.add(size_t(n_cells + 1) * 4), h_ci = H.add(size_t(std::max(n_items, 1)) * 4),
                 h_dir = H.add(is_lines ? size_t(std::max(n2, 1)) * 16 : 0), h_m12 = H.add(size_t(n_rows) * 4 + 4);
    const size_t frame_bytes = h_m12; // everything before the match vector goes to every device
    if ((st = s->ensure_stage(H.total)) != PLM_OK) return st;
    pack_rows(s->h_stage + h_d2, d2, n2, step2);
    std::memcpy(s->h_stage + h_cs, cell_start, size_t(n_cells + 1) * 4);
    if (n_items > 0) std::memcpy(s->h_stage + h_ci, cell_items, size_t(n_items) * 4);
    if (is_lines && n2 > 0) std::memcpy(s->h_stage + h_dir, dirs2, size_t(n2) * 16);
    std::memcpy(s->h_stage + h_m12, m12_inout, size_t(n_rows) * 4);
    Layout D;
    const size_t d_frame = D.add(frame_bytes), d_m12 = D.add(size_t(s->per_cap) * 4), d_cnt = D.add(16), d_out = D.add(size_t(n_rows) * 4),
                 d_tot = D.add(16), d_key = D.add(size_t(std::max(n2, 1)) * 8);
    std::vector<plm_dev_grid_args> args(s->n);
    for (int g = 0; g < s->n; ++g) {
        if ((st = s->ensure_scratch(g, D.total)) != PLM_OK) return st;
        CU_TRY(cudaSetDevice(s->devices[g]));
        cudaStream_t sg = s->ctx[g]->stream;
        char *X = s->scratch[g];
        const int64_t nl = s->hi[g] - s->lo[g];
        CU_TRY(cudaMemcpyAsync(X + d_frame, s->h_stage, frame_bytes, cudaMemcpyHostToDevice, sg));
        if (nl > 0) CU_TRY(cudaMemcpyAsync(X + d_m12, s->h_stage + h_m12 + size_t(s->lo[g]) * 4, size_t(nl) * 4, cudaMemcpyHostToDevice, sg));
        CU_TRY(cudaMemsetAsync(X + d_cnt, 0, 16, sg));
        plm_dev_grid_args &a = args[g];
        std::memset(&a, 0, sizeof(a));
        a.coords = s->coords[g];
        a.d1 = s->rows[g];
        a.cell_start = reinterpret_cast<const int32_t *>(X + d_frame + h_cs);
        a.cell_items = reinterpret_cast<const int32_t *>(X + d_frame + h_ci);
        a.d2 = X + d_frame + h_d2;
        a.dirs2 = is_lines ? reinterpret_cast<const double *>(X + d_frame + h_dir) : nullptr;
        a.m12_inout = reinterpret_cast<int32_t *>(X + d_m12);
        a.count = reinterpret_cast<int32_t *>(X + d_cnt);
        a.i1_base = s->lo[g];
        a.ratio = ratio;
        a.line_sim_th = line_sim_th;
        a.n1 = static_cast<int32_t>(nl);
        a.n2 = n2;
        a.grid_rows = grid_rows;
        a.grid_cols = grid_cols;
        a.is_lines = is_lines;
        a.best_lr = best_lr ? 1 : 0;
        for (int i = 0; i < 4; ++i) a.win[i] = win[i];
        // size the context's own scratch now (the fused call below must not allocate while peers spin)
        if (s->n > 1 && (st = plm_dev_sharded_match_grid_prepare(s->ctx[g], &a)) != PLM_OK) return st;
    }
    if (s->n > 1) {
        plm_peer_group grp;
        grp.xchg = s->xchg.data();
        grp.gather = s->gather.data();
        grp.world = s->n;
        grp.q_cap = s->q_cap;
        grp.pad_ = 0;
        grp.n_rows_cap = s->rows_cap;
        grp.xchg_epoch = s->xchg_epoch + 1;
        grp.gather_epoch = s->gather_epoch + 1;
        s->xchg_epoch += (best_lr && n2 > 0) ? 2 : 0;
        s->gather_epoch += 1;
        for (int g = 0; g < s->n; ++g) {
            grp.rank = g;
            char *X = s->scratch[g];
            if ((st = plm_dev_sharded_match_grid(s->ctx[g], &args[g], &grp, n_rows, reinterpret_cast<int32_t *>(X + d_out),
                                                 reinterpret_cast<int32_t *>(X + d_tot), s->error_host)) != PLM_OK)
                return st;
        }
    } else {
        char *X = s->scratch[0];
        plm_dev_grid_args &a = args[0];
        uint64_t *key = reinterpret_cast<uint64_t *>(X + d_key);
        if ((st = plm_dev_grid_match(s->ctx[0], &a, nullptr, key)) != PLM_OK) return st;
        if (best_lr && n2 > 0 && a.n1 > 0) {
            plm::cross_check_keys_kernel<<<(a.n1 + 255) / 256, 256, 0, s->ctx[0]->stream>>>(a.m12_inout, a.n1, 0, reinterpret_cast<const unsigned long long *>(key),
                                                                                           n2, a.count);
            s->ctx[0]->launches++;
            CU_TRY(cudaGetLastError());
        }
        CU_TRY(cudaMemcpyAsync(X + d_out, X + d_m12, size_t(n_rows) * 4, cudaMemcpyDeviceToDevice, s->ctx[0]->stream));
        CU_TRY(cudaMemcpyAsync(X + d_tot, X + d_cnt, 4, cudaMemcpyDeviceToDevice, s->ctx[0]->stream));
    }
    CU_TRY(cudaSetDevice(s->devices[0]));
    char *X0 = s->scratch[0];
    CU_TRY(cudaMemcpyAsync(s->h_stage + h_m12, X0 + d_out, size_t(n_rows) * 4, cudaMemcpyDeviceToHost, s->ctx[0]->stream));
    CU_TRY(cudaMemcpyAsync(s->h_stage + h_m12 + size_t(n_rows) * 4, X0 + d_tot, 4, cudaMemcpyDeviceToHost, s->ctx[0]->stream));
    CU_TRY(cudaStreamSynchronize(s->ctx[0]->stream));
    if ((st = shard_check_alive(s)) != PLM_OK) return st;
    std::memcpy(m12_inout, s->h_stage + h_m12, size_t(n_rows) * 4);
    int32_t c;
    std::memcpy(&c, s->h_stage + h_m12 + size_t(n_rows) * 4, 4);
    *n_matches = c;
    cudaSetDevice(prev);
    return PLM_OK;
}

// StVO::match with the sharded map as desc1 (the brute-force fallback of matchMap2KF*, mapHandler.cpp:645-650, on the
// vector matchGrid just filled): direction 12 local per shard, direction 21 through the peer-memory top-2 exchange,
// mutual check, all-gather.  Same degenerate-size rules as plm_match.
PLM_API int plm_shard_match(plm_shard *s, const uint8_t *d2, int n2, size_t step2, float nnr, int best_lr, int32_t *m12_inout,
                            int *n_matches) {
    int st = shard_check_alive(s);
    if (st != PLM_OK) return st;
    if ((st = check_desc(d2, n2, step2)) != PLM_OK) return st;
    if (!n_matches) return fail(PLM_E_INVALID, "null n_matches");
    *n_matches = 0;
    const int64_t n_rows = s->n_rows;
    if (n_rows == 0) return PLM_OK;
    if (!m12_inout) return fail(PLM_E_INVALID, "null m12");
    if (n2 == 0) return fail(PLM_E_TRAIN, "matchNNR: empty train set");
    if (n_rows > INT_MAX) return fail(PLM_E_UNSUPPORTED, "more than 2^31 map rows");
    if (s->n > 1 && n2 > s->q_cap) return fail(PLM_E_UNSUPPORTED, "frame too large for the exchange buffers (q_cap)");
    int prev = 0;
    cudaGetDevice(&prev);
    Layout H;
    const size_t h_d2 = H.add(size_t(n2) * 32), h_m12 = H.add(size_t(n_rows) * 4 + 4);
    if ((st = s->ensure_stage(H.total)) != PLM_OK) return st;
    pack_rows(s->h_stage + h_d2, d2, n2, step2);
    std::memcpy(s->h_stage + h_m12, m12_inout, size_t(n_rows) * 4);
    Layout D;
    const size_t d_d2 = D.add(size_t(n2) * 32), d_m12 = D.add(size_t(s->per_cap) * 4), d_out = D.add(size_t(n_rows) * 4), d_tot = D.add(16),
                 d_top = D.add(size_t(s->per_cap) * 16), d_part = D.add(size_t(n2) * 16), d_m21 = D.add(size_t(n2) * 4);
    for (int g = 0; g < s->n; ++g) {
        if ((st = s->ensure_scratch(g, D.total)) != PLM_OK) return st;
        CU_TRY(cudaSetDevice(s->devices[g]));
        cudaStream_t sg = s->ctx[g]->stream;
        char *X = s->scratch[g];
        const int64_t nl = s->hi[g] - s->lo[g];
        CU_TRY(cudaMemcpyAsync(X + d_d2, s->h_stage + h_d2, size_t(n2) * 32, cudaMemcpyHostToDevice, sg));
        if (nl > 0) CU_TRY(cudaMemcpyAsync(X + d_m12, s->h_stage + h_m12 + size_t(s->lo[g]) * 4, size_t(nl) * 4, cudaMemcpyHostToDevice, sg));
        if (s->n > 1 && (st = plm_dev_sharded_match_prepare(s->ctx[g], static_cast<int>(nl), n2)) != PLM_OK) return st;
    }
    if (s->n > 1) {
        plm_peer_group grp;
        grp.xchg = s->xchg.data();
        grp.gather = s->gather.data();
        grp.world = s->n;
        grp.q_cap = s->q_cap;
        grp.pad_ = 0;
        grp.n_rows_cap = s->rows_cap;
        grp.xchg_epoch = s->xchg_epoch + 1;
        grp.gather_epoch = s->gather_epoch + 1;
        s->xchg_epoch += best_lr ? 1 : 0;
        s->gather_epoch += 1;
        for (int g = 0; g < s->n; ++g) {
            grp.rank = g;
            char *X = s->scratch[g];
            if ((st = plm_dev_sharded_match(s->ctx[g], s->rows[g], static_cast<int>(s->hi[g] - s->lo[g]), s->lo[g], X + d_d2, n2, nnr, best_lr ? 1 : 0,
                                            reinterpret_cast<int32_t *>(X + d_m12), &grp, n_rows, reinterpret_cast<int32_t *>(X + d_out),
                                            reinterpret_cast<int32_t *>(X + d_tot), s->error_host)) != PLM_OK)
                return st;
        }
    } else {
        char *X = s->scratch[0];
        plm_ctx *c0 = s->ctx[0];
        const int n1 = static_cast<int>(n_rows);
        int32_t *m12 = reinterpret_cast<int32_t *>(X + d_m12), *cnt = reinterpret_cast<int32_t *>(X + d_tot), *m21 = reinterpret_cast<int32_t *>(X + d_m21);
        uint64_t *top = reinterpret_cast<uint64_t *>(X + d_top), *part = reinterpret_cast<uint64_t *>(X + d_part);
        CU_TRY(cudaMemsetAsync(cnt, 0, 16, c0->stream));
        if ((st = plm_dev_knn2(c0, s->rows[0], n1, X + d_d2, n2, 0, top)) != PLM_OK) return st;
        if ((st = plm_dev_nnr_accept(c0, top, n1, nnr, m12, cnt)) != PLM_OK) return st;
        if (best_lr) {
            CU_TRY(cudaMemsetAsync(m21, 0xFF, size_t(n2) * 4, c0->stream));
            if ((st = plm_dev_knn2(c0, X + d_d2, n2, s->rows[0], n1, 0, part)) != PLM_OK) return st;
            if ((st = plm_dev_nnr_accept(c0, part, n2, nnr, m21, nullptr)) != PLM_OK) return st;
            if ((st = plm_dev_cross_check(c0, m12, n1, 0, m21, n2, cnt)) != PLM_OK) return st;
        }
        CU_TRY(cudaMemcpyAsync(X + d_out, m12, size_t(n_rows) * 4, cudaMemcpyDeviceToDevice, c0->stream));
    }
    CU_TRY(cudaSetDevice(s->devices[0]));
    char *X0 = s->scratch[0];
    CU_TRY(cudaMemcpyAsync(s->h_stage + h_m12, X0 + d_out, size_t(n_rows) * 4, cudaMemcpyDeviceToHost, s->ctx[0]->stream));
    CU_TRY(cudaMemcpyAsync(s->h_stage + h_m12 + size_t(n_rows) * 4, X0 + d_tot, 4, cudaMemcpyDeviceToHost, s->ctx[0]->stream));
    CU_TRY(cudaStreamSynchronize(s->ctx[0]->stream));
    if ((st = shard_check_alive(s)) != PLM_OK) return st;
    std::memcpy(m12_inout, s->h_stage + h_m12, size_t(n_rows) * 4);
    int32_t c;
    std::memcpy(&c, s->h_stage + h_m12 + size_t(n_rows) * 4, 4);
    *n_matches = c;
    cudaSetDevice(prev);
    return PLM_OK;
}

PLM_API int plm_shard_synchronize(plm_shard *s) {
    if (!s) return fail(PLM_E_INVALID, "null shard set");
    int prev = 0;
    cudaGetDevice(&prev);
    int st = shard_sync_all(s);
    cudaSetDevice(prev);
    if (st != PLM_OK) return st;
    return shard_check_alive(s);
}
