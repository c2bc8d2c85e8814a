// Host side of the device-resident stereo-frame pipeline (plm_frames_*; included by plmatch.cu).
//
// Frames are handled in CHUNKS of consecutive frames.  A chunk is the unit of host preparation (validation,
// shared-memory capacities, job tables), of the host -> device copies (the chunk's slice of every arena), of the
// four launches and of the device -> host copies.  plm_frames_upload / run / fetch use one chunk for the
// whole upload; plm_frames_process pipelines many chunks over three streams so that the copy engines and the
// SMs work on different chunks at the same time.

struct plm_frames_chunk {
    int f0 = 0, f1 = 0; // frames [f0, f1)
    plm::StereoCaps caps_p{1, 1, 1, 8, 8, 0}, caps_l{1, 1, 1, 4, 8, 0};
    size_t smem_p = 0, smem_l = 0, smem_fp = 0, smem_fl = 0;
};

struct plm_frames {
    plm_ctx *ctx = nullptr;
    char *d_buf = nullptr;
    size_t d_cap = 0;
    char *h_tab = nullptr; // pinned staging of the job tables (all chunks of one call)
    size_t h_cap = 0;
    int n_frames = 0;
    int64_t NP = 0, NL = 0; // left point / line slots
    int64_t n_rows = 0, n_kp = 0, n_ln = 0;
    plm::FrameCfg cfg;
    float nnr_p = 0.f, nnr_l = 0.f;
    std::vector<int64_t> lp_off, ll_off;
    std::vector<plm_frames_chunk> chunks;
    size_t o_desc = 0, o_kp = 0, o_ln = 0, o_cdesc_p = 0, o_cdesc_l = 0;
    size_t o_jobs = 0; // [n_frames] FrameJobs
    size_t o_m12_p = 0, o_m12_l = 0, o_kept_p = 0, o_kept_l = 0, o_pt_disp = 0, o_pt_P = 0, o_ls_disp = 0, o_ls_sP = 0,
           o_ls_eP = 0, o_ls_le = 0, o_f2f_p = 0, o_f2f_l = 0, o_counts = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr, s_lines = nullptr;
    std::vector<cudaEvent_t> events;
    int64_t h2d = 0, d2h = 0;
    bool ready = false;

    int ensure(size_t bytes) {
        if (bytes <= d_cap) return PLM_OK;
        CU_TRY(cudaStreamSynchronize(ctx->stream));
        if (d_buf) CU_TRY(cudaFree(d_buf));
        d_buf = nullptr;
        d_cap = 0;
        CU_TRY(cudaMalloc(reinterpret_cast<void **>(&d_buf), align_up(bytes, 1 << 20)));
        d_cap = align_up(bytes, 1 << 20);
        return PLM_OK;
    }
    int ensure_tab(size_t bytes) {
        if (bytes <= h_cap) return PLM_OK;
        CU_TRY(cudaStreamSynchronize(ctx->stream));
        if (h_tab) CU_TRY(cudaFreeHost(h_tab));
        h_tab = nullptr;
        h_cap = 0;
        CU_TRY(cudaHostAlloc(reinterpret_cast<void **>(&h_tab), align_up(bytes, 1 << 16), cudaHostAllocDefault));
        h_cap = align_up(bytes, 1 << 16);
        return PLM_OK;
    }
    bool timing_events = false; // PLM_FRAMES_TRACE: per-chunk device timeline
    int event(size_t i, cudaEvent_t *out) {
        while (events.size() <= i) {
            cudaEvent_t e;
            CU_TRY(cudaEventCreateWithFlags(&e, timing_events ? cudaEventDefault : cudaEventDisableTiming));
            events.push_back(e);
        }
        *out = events[i];
        return PLM_OK;
    }
};

namespace {

constexpr int FRAMES_MAX_FEATURES = 4096;
constexpr int FRAMES_THREADS_P = 256, FRAMES_THREADS_L = 128;

// Number of cells the Bresenham walk of lineIterator.cpp:34-77 visits for a segment with already
// scaled endpoints (same arithmetic as plm::line_walk: the walk runs x = int(x1') .. int(x2')).
inline long long line_walk_cells(double x1, double y1, double x2, double y2) {
    const bool steep = std::fabs(y2 - y1) > std::fabs(x2 - x1);
    if (steep) {
        std::swap(x1, y1);
        std::swap(x2, y2);
    }
    if (x1 > x2) std::swap(x1, x2);
    if (!(std::fabs(x1) < 1e9) || !(std::fabs(x2) < 1e9)) return LLONG_MAX; // NaN / absurd coordinates
    const long long n = static_cast<long long>(static_cast<int>(x2)) - static_cast<int>(x1) + 1;
    return n > 0 ? n : 0;
}

// The four jobs of one frame as one record: a chunk of consecutive frames is ONE host -> device copy of the job tables.
struct FrameJobs {
    plm::StereoJob sp, sl; // stereo stage: points, lines
    plm::F2FJob fp, fl;    // temporal stage against the previous frame: points, lines
};

template <int THREADS>
int launch_stereo(plm_ctx *ctx, cudaStream_t stream, const plm::StereoJob *jobs, int n_jobs, const plm::FrameCfg &cfg,
                  const plm::StereoCaps &caps, size_t smem) {
    if (n_jobs == 0) return PLM_OK;
    CU_TRY(cudaFuncSetAttribute(plm::stereo_frame_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
    plm::stereo_frame_kernel<THREADS><<<n_jobs, THREADS, smem, stream>>>(jobs, sizeof(FrameJobs), cfg, caps);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
    return PLM_OK;
}

template <int THREADS>
int launch_f2f(plm_ctx *ctx, cudaStream_t stream, const plm::F2FJob *jobs, int n_jobs, int best_lr, size_t smem) {
    if (n_jobs == 0) return PLM_OK;
    CU_TRY(cudaFuncSetAttribute(plm::f2f_match_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
    plm::f2f_match_kernel<THREADS><<<n_jobs, THREADS, smem, stream>>>(jobs, sizeof(FrameJobs), best_lr);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
    return PLM_OK;
}

// Argument checks, slot offsets, device layout and allocation for a whole call.
int frames_layout(plm_frames *fr, const uint8_t *desc_arena, int64_t n_rows, const float *kp_arena, int64_t n_kp,
                  const float *ln_arena, int64_t n_ln, const plm_frame_rec *frames, int n_frames,
                  const plm_frame_config *c) {
    if (!fr || !c) return fail(PLM_E_INVALID, "null frames / config");
    if (n_rows < 0 || n_kp < 0 || n_ln < 0 || n_frames < 0) return fail(PLM_E_INVALID, "negative size");
    if ((n_rows > 0 && !desc_arena) || (n_kp > 0 && !kp_arena) || (n_ln > 0 && !ln_arena) || (n_frames > 0 && !frames))
        return fail(PLM_E_INVALID, "null pointer");
    if (c->grid_rows <= 0 || c->grid_cols <= 0) return fail(PLM_E_GRID, "[GridStructure] invalid dimension");
    if (static_cast<long long>(c->grid_rows) * c->grid_cols > (1 << 14)) return fail(PLM_E_UNSUPPORTED, "grid too large for the frame pipeline");
    if (c->min_ratio_12p > 1.0) return fail(PLM_E_RATIO, plm_status_string(PLM_E_RATIO));
    fr->ready = false;
    fr->lp_off.assign(static_cast<size_t>(n_frames) + 1, 0);
    fr->ll_off.assign(static_cast<size_t>(n_frames) + 1, 0);
    for (int f = 0; f < n_frames; ++f) {
        const plm_frame_rec &r = frames[f];
        if (r.n_pl < 0 || r.n_pr < 0 || r.n_ll < 0 || r.n_lr < 0 || r.desc_pl < 0 || r.desc_pr < 0 || r.desc_ll < 0 ||
            r.desc_lr < 0 || r.kp_l < 0 || r.kp_r < 0 || r.ln_l < 0 || r.ln_r < 0 || r.desc_pl + r.n_pl > n_rows ||
            r.desc_pr + r.n_pr > n_rows || r.desc_ll + r.n_ll > n_rows || r.desc_lr + r.n_lr > n_rows ||
            r.kp_l + r.n_pl > n_kp || r.kp_r + r.n_pr > n_kp || r.ln_l + r.n_ll > n_ln || r.ln_r + r.n_lr > n_ln)
            return fail(PLM_E_INVALID, "frame record outside its arena");
        if (r.n_pl > FRAMES_MAX_FEATURES || r.n_pr > FRAMES_MAX_FEATURES || r.n_ll > FRAMES_MAX_FEATURES ||
            r.n_lr > FRAMES_MAX_FEATURES)
            return fail(PLM_E_UNSUPPORTED, "more than 4096 features of one kind in a frame");
        fr->lp_off[f + 1] = fr->lp_off[f] + r.n_pl;
        fr->ll_off[f + 1] = fr->ll_off[f] + r.n_ll;
    }
    const int64_t NP = fr->lp_off[n_frames], NL = fr->ll_off[n_frames];
    const size_t F = size_t(std::max(n_frames, 1));
    Layout L;
    fr->o_desc = L.add(size_t(n_rows) * 32);
    fr->o_kp = L.add(size_t(n_kp) * 8);
    fr->o_ln = L.add(size_t(n_ln) * 16);
    fr->o_cdesc_p = L.add(size_t(NP) * 32);
    fr->o_cdesc_l = L.add(size_t(NL) * 32);
    fr->o_m12_p = L.add(size_t(NP) * 4);
    fr->o_m12_l = L.add(size_t(NL) * 4);
    fr->o_kept_p = L.add(size_t(NP) * 4);
    fr->o_kept_l = L.add(size_t(NL) * 4);
    fr->o_pt_disp = L.add(size_t(NP) * 8);
    fr->o_pt_P = L.add(size_t(NP) * 24);
    fr->o_ls_disp = L.add(size_t(NL) * 16);
    fr->o_ls_sP = L.add(size_t(NL) * 24);
    fr->o_ls_eP = L.add(size_t(NL) * 24);
    fr->o_ls_le = L.add(size_t(NL) * 24);
    fr->o_f2f_p = L.add(size_t(NP) * 4);
    fr->o_f2f_l = L.add(size_t(NL) * 4);
    fr->o_counts = L.add(F * 6 * 4);
    fr->o_jobs = L.add(F * sizeof(FrameJobs));
    int st = fr->ensure(L.total);
    if (st != PLM_OK) return st;
    if ((st = fr->ensure_tab(F * sizeof(FrameJobs))) != PLM_OK) return st;
    fr->n_frames = n_frames;
    fr->NP = NP;
    fr->NL = NL;
    fr->n_rows = n_rows;
    fr->n_kp = n_kp;
    fr->n_ln = n_ln;
    fr->nnr_p = static_cast<float>(c->min_ratio_12p);
    fr->nnr_l = static_cast<float>(c->min_ratio_12l);
    plm::FrameCfg &g = fr->cfg;
    g.inv_w = c->inv_width;
    g.inv_h = c->inv_height;
    g.ratio = c->min_ratio_12p;
    g.line_sim_th = c->line_sim_th;
    g.max_dist_epip = c->max_dist_epip;
    g.min_disp = c->min_disp;
    g.line_horiz_th = c->line_horiz_th;
    g.stereo_overlap_th = c->stereo_overlap_th;
    g.ls_min_disp_ratio = c->ls_min_disp_ratio;
    g.cam_b = c->cam_b;
    g.cam_fx = c->cam_fx;
    g.cam_cx = c->cam_cx;
    g.cam_cy = c->cam_cy;
    g.grid_rows = c->grid_rows;
    g.grid_cols = c->grid_cols;
    g.matching_s_ws = c->matching_s_ws;
    g.best_lr = c->best_lr ? 1 : 0;
    fr->chunks.clear();
    fr->h2d = fr->d2h = 0;
    return PLM_OK;
}

struct FrameSpan {
    int64_t lo = INT64_MAX, hi = 0;
    void add(int64_t off, int64_t n) {
        if (n > 0) {
            lo = std::min(lo, off);
            hi = std::max(hi, off + n);
        }
    }
};

// Result of the host-only preparation of a chunk (may run on a helper thread: no CUDA calls, no
// thread-local error state).
struct FrameChunkPrep {
    plm_frames_chunk ch;
    FrameSpan sd[4], sk[2], sn[2]; // arena slices: descriptors (4 sets), keypoints (l/r), segments (l/r)
    int status = PLM_OK;
    const char *error = "";
};

// Host preparation of frames [f0, f1): shared-memory capacities, arena slices and the job tables, written
// into the pinned staging block at this chunk's rows.
FrameChunkPrep frames_chunk_prep(const plm_frames *fr, const float *ln_arena, const plm_frame_rec *frames, int f0, int f1) {
    FrameChunkPrep out;
    const plm_ctx *ctx = fr->ctx;
    const plm::FrameCfg &g = fr->cfg;
    plm_frames_chunk &ch = out.ch;
    FrameSpan *sd = out.sd, *sk = out.sk, *sn = out.sn;
    auto bad = [&](int status, const char *msg) {
        out.status = status;
        out.error = msg;
        return out;
    };
    ch.f0 = f0;
    ch.f1 = f1;
    int cap_pl = 1, cap_pr = 1, cap_ll = 1, cap_lr = 1;
    long long cap_items_l = 1;
    const long long walk_max = static_cast<long long>(g.grid_rows) + g.grid_cols + 2;
    for (int f = f0; f < f1; ++f) {
        const plm_frame_rec &r = frames[f];
        cap_pl = std::max(cap_pl, r.n_pl);
        cap_pr = std::max(cap_pr, r.n_pr);
        cap_ll = std::max(cap_ll, r.n_ll);
        cap_lr = std::max(cap_lr, r.n_lr);
        long long items = 0, longest = 0;
        const float *l = ln_arena + 4 * r.ln_r;
        for (int j = 0; j < r.n_lr; ++j, l += 4) {
            const long long n = line_walk_cells(double(l[0]) * g.inv_w, double(l[1]) * g.inv_h, double(l[2]) * g.inv_w,
                                                double(l[3]) * g.inv_h);
            longest = std::max(longest, n);
            items += std::min(n, walk_max + 1);
        }
        if (longest > walk_max) return bad(PLM_E_UNSUPPORTED, "line segment far outside the image (Bresenham walk too long)");
        cap_items_l = std::max(cap_items_l, items);
        sd[0].add(r.desc_pl, r.n_pl);
        sd[1].add(r.desc_pr, r.n_pr);
        sd[2].add(r.desc_ll, r.n_ll);
        sd[3].add(r.desc_lr, r.n_lr);
        sk[0].add(r.kp_l, r.n_pl);
        sk[1].add(r.kp_r, r.n_pr);
        sn[0].add(r.ln_l, r.n_ll);
        sn[1].add(r.ln_r, r.n_lr);
    }
    // the temporal jobs of frame f0 read frame f0 - 1: their capacities count too
    int cap_fp = cap_pl, cap_fl = cap_ll;
    if (f0 > 0) {
        cap_fp = std::max(cap_fp, frames[f0 - 1].n_pl);
        cap_fl = std::max(cap_fl, frames[f0 - 1].n_ll);
    }
    const int n_cells = g.grid_rows * g.grid_cols;
    // pair-list capacity: a stereo window holds 2-3 point candidates per query on EuRoC-like frames; denser
    // frames take the chunk phases inside the same kernel
    const int pairs_p = std::min(g_frames_pairs_per_row[0] * cap_pl, 16384);
    const int pairs_l = std::min(g_frames_pairs_per_row[1] * cap_ll, 16384);
    ch.caps_p = plm::StereoCaps{cap_pl, cap_pr, cap_pr, g_frames_threads_p / 32, pairs_p, 0};
    ch.caps_l = plm::StereoCaps{cap_ll, cap_lr, static_cast<int>(cap_items_l), g_frames_threads_l / 32, pairs_l, 0};
    ch.smem_p = plm::stereo_frame_smem(ch.caps_p, n_cells, false);
    ch.smem_l = plm::stereo_frame_smem(ch.caps_l, n_cells, true);
    ch.smem_fp = plm::f2f_smem(cap_fp, cap_fp);
    ch.smem_fl = plm::f2f_smem(cap_fl, cap_fl);
    const size_t budget = ctx->smem_optin - 1024;
    if (ch.smem_p > budget || ch.smem_l > budget || ch.smem_fp > budget || ch.smem_fl > budget)
        return bad(PLM_E_UNSUPPORTED, "frame too large for shared memory");

    // ---- job tables of the chunk ---------------------------------------------------------------------------
    char *D = fr->d_buf;
    const size_t F = size_t(std::max(fr->n_frames, 1));
    FrameJobs *tab = reinterpret_cast<FrameJobs *>(fr->h_tab);
    const uint4 *d_desc = reinterpret_cast<const uint4 *>(D + fr->o_desc);
    const float *d_kp = reinterpret_cast<const float *>(D + fr->o_kp);
    const float *d_ln = reinterpret_cast<const float *>(D + fr->o_ln);
    int32_t *d_counts = reinterpret_cast<int32_t *>(D + fr->o_counts);
    uint4 *cdesc_p = reinterpret_cast<uint4 *>(D + fr->o_cdesc_p), *cdesc_l = reinterpret_cast<uint4 *>(D + fr->o_cdesc_l);
    const std::vector<int64_t> &lp_off = fr->lp_off, &ll_off = fr->ll_off;
    for (int f = f0; f < f1; ++f) {
        const plm_frame_rec &r = frames[f];
        plm::StereoJob &a = tab[f].sp;
        std::memset(&a, 0, sizeof(a));
        a.geo_l = d_kp + 2 * r.kp_l;
        a.geo_r = d_kp + 2 * r.kp_r;
        a.d_l = d_desc + 2 * r.desc_pl;
        a.d_r = d_desc + 2 * r.desc_pr;
        a.m12 = reinterpret_cast<int32_t *>(D + fr->o_m12_p) + lp_off[f];
        a.cdesc = cdesc_p + 2 * lp_off[f];
        a.kept_i1 = reinterpret_cast<int32_t *>(D + fr->o_kept_p) + lp_off[f];
        a.o0 = reinterpret_cast<double *>(D + fr->o_pt_disp) + lp_off[f];
        a.o1 = reinterpret_cast<double *>(D + fr->o_pt_P) + 3 * lp_off[f];
        a.counts = d_counts + 6 * f;
        a.n_l = r.n_pl;
        a.n_r = r.n_pr;
        a.is_lines = 0;
        plm::StereoJob &b = tab[f].sl;
        std::memset(&b, 0, sizeof(b));
        b.geo_l = d_ln + 4 * r.ln_l;
        b.geo_r = d_ln + 4 * r.ln_r;
        b.d_l = d_desc + 2 * r.desc_ll;
        b.d_r = d_desc + 2 * r.desc_lr;
        b.m12 = reinterpret_cast<int32_t *>(D + fr->o_m12_l) + ll_off[f];
        b.cdesc = cdesc_l + 2 * ll_off[f];
        b.kept_i1 = reinterpret_cast<int32_t *>(D + fr->o_kept_l) + ll_off[f];
        b.o0 = reinterpret_cast<double *>(D + fr->o_ls_disp) + 2 * ll_off[f];
        b.o1 = reinterpret_cast<double *>(D + fr->o_ls_sP) + 3 * ll_off[f];
        b.o2 = reinterpret_cast<double *>(D + fr->o_ls_eP) + 3 * ll_off[f];
        b.o3 = reinterpret_cast<double *>(D + fr->o_ls_le) + 3 * ll_off[f];
        b.counts = d_counts + 6 * f + 2;
        b.n_l = r.n_ll;
        b.n_r = r.n_lr;
        b.is_lines = 1;
        // temporal job of (frame f - 1, frame f); row 0 of the tables stays unused
        plm::F2FJob &p = tab[f].fp;
        plm::F2FJob &q = tab[f].fl;
        std::memset(&p, 0, sizeof(p));
        std::memset(&q, 0, sizeof(q));
        if (f >= 1) {
            p.d1 = cdesc_p + 2 * lp_off[f - 1];
            p.d2 = a.cdesc;
            p.n1_ptr = d_counts + 6 * (f - 1) + 1;
            p.n2_ptr = d_counts + 6 * f + 1;
            p.m12 = reinterpret_cast<int32_t *>(D + fr->o_f2f_p) + lp_off[f - 1];
            p.count = d_counts + 6 * f + 4;
            p.cap1 = frames[f - 1].n_pl;
            p.cap2 = r.n_pl;
            p.nnr = fr->nnr_p;
            q.d1 = cdesc_l + 2 * ll_off[f - 1];
            q.d2 = b.cdesc;
            q.n1_ptr = d_counts + 6 * (f - 1) + 3;
            q.n2_ptr = d_counts + 6 * f + 3;
            q.m12 = reinterpret_cast<int32_t *>(D + fr->o_f2f_l) + ll_off[f - 1];
            q.count = d_counts + 6 * f + 5;
            q.cap1 = frames[f - 1].n_ll;
            q.cap2 = r.n_ll;
            q.nnr = fr->nnr_l;
        }
    }
    return out;
}

// Host -> device copies of a prepared chunk on `s`: its rows of the job tables and its arena slices (a
// scattered arena layout only makes the slices larger).  Appends the chunk to fr->chunks.
int frames_chunk_copy(plm_frames *fr, FrameChunkPrep &pr, const uint8_t *desc_arena, const float *kp_arena,
                      const float *ln_arena, cudaStream_t s) {
    if (pr.status != PLM_OK) return fail(pr.status, pr.error);
    char *D = fr->d_buf;
    const int f0 = pr.ch.f0, f1 = pr.ch.f1;
    const size_t F = size_t(std::max(fr->n_frames, 1));
    const FrameJobs *tab = reinterpret_cast<const FrameJobs *>(fr->h_tab);
    const size_t nf = size_t(f1 - f0);
    if (nf) {
        CU_TRY(cudaMemcpyAsync(D + fr->o_jobs + f0 * sizeof(FrameJobs), tab + f0, nf * sizeof(FrameJobs), cudaMemcpyHostToDevice, s));
        fr->h2d += static_cast<int64_t>(nf * sizeof(FrameJobs));
    }
    auto copy_spans = [&](FrameSpan *sp_, int n_sp, size_t dev_off, const void *host, size_t elem) -> cudaError_t {
        std::sort(sp_, sp_ + n_sp, [](const FrameSpan &x, const FrameSpan &y) { return x.lo < y.lo; });
        for (int i = 0; i < n_sp; ++i) {
            if (sp_[i].hi <= sp_[i].lo) continue;
            int64_t lo = sp_[i].lo, hi = sp_[i].hi;
            while (i + 1 < n_sp && sp_[i + 1].hi > sp_[i + 1].lo && sp_[i + 1].lo <= hi) hi = std::max(hi, sp_[++i].hi);
            const cudaError_t e = cudaMemcpyAsync(D + dev_off + size_t(lo) * elem, static_cast<const char *>(host) + size_t(lo) * elem,
                                                  size_t(hi - lo) * elem, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) return e;
            fr->h2d += (hi - lo) * static_cast<int64_t>(elem);
        }
        return cudaSuccess;
    };
    CU_TRY(copy_spans(pr.sd, 4, fr->o_desc, desc_arena, 32));
    CU_TRY(copy_spans(pr.sk, 2, fr->o_kp, kp_arena, 8));
    CU_TRY(copy_spans(pr.sn, 2, fr->o_ln, ln_arena, 16));
    fr->chunks.push_back(pr.ch);
    return PLM_OK;
}

// The four launches of chunk number k.  Points and lines are independent chains (stereo stage -> temporal
// stage), so the line chain runs on a second stream forked from / joined back into the context's stream:
// a chunk of a few hundred frames does not fill the chip with one kernel at a time.
int frames_chunk_run(plm_frames *fr, const plm_frames_chunk &ch, size_t k) {
    plm_ctx *ctx = fr->ctx;
    char *D = fr->d_buf;
    const int n = ch.f1 - ch.f0;
    if (n <= 0) return PLM_OK;
    cudaStream_t s = ctx->stream;
    if (!fr->s_lines) CU_TRY(cudaStreamCreateWithFlags(&fr->s_lines, cudaStreamNonBlocking));
    cudaEvent_t e_fork, e_join;
    int st;
    if ((st = fr->event(4 * k + 3, &e_fork)) != PLM_OK) return st;
    if ((st = fr->event(4 * k + 4, &e_join)) != PLM_OK) return st;
    const int64_t p0 = fr->lp_off[ch.f0], p1 = fr->lp_off[ch.f1], l0 = fr->ll_off[ch.f0], l1 = fr->ll_off[ch.f1];
    CU_TRY(cudaMemsetAsync(D + fr->o_counts + size_t(ch.f0) * 24, 0, size_t(n) * 24, s));
    // slots past a frame's kept count (and the last frame's f2f range) read as -1
    if (p1 > p0) {
        CU_TRY(cudaMemsetAsync(D + fr->o_kept_p + size_t(p0) * 4, 0xFF, size_t(p1 - p0) * 4, s));
        CU_TRY(cudaMemsetAsync(D + fr->o_f2f_p + size_t(p0) * 4, 0xFF, size_t(p1 - p0) * 4, s));
    }
    if (l1 > l0) {
        CU_TRY(cudaMemsetAsync(D + fr->o_kept_l + size_t(l0) * 4, 0xFF, size_t(l1 - l0) * 4, s));
        CU_TRY(cudaMemsetAsync(D + fr->o_f2f_l + size_t(l0) * 4, 0xFF, size_t(l1 - l0) * 4, s));
    }
    CU_TRY(cudaEventRecord(e_fork, s));
    CU_TRY(cudaStreamWaitEvent(fr->s_lines, e_fork, 0));
    const int t0 = std::max(ch.f0, 1); // frame 0 has no predecessor
    const bool wide_p = ch.caps_p.warps * 32 == 512;
    if ((st = wide_p ? launch_stereo<512>(ctx, s, &reinterpret_cast<const FrameJobs *>(D + fr->o_jobs)[ch.f0].sp, n, fr->cfg,
                                          ch.caps_p, ch.smem_p)
                     : launch_stereo<FRAMES_THREADS_P>(ctx, s, &reinterpret_cast<const FrameJobs *>(D + fr->o_jobs)[ch.f0].sp, n,
                                                       fr->cfg, ch.caps_p, ch.smem_p)) != PLM_OK) return st;
    const bool wide_l = ch.caps_l.warps * 32 == FRAMES_THREADS_P;
    if ((st = wide_l ? launch_stereo<FRAMES_THREADS_P>(ctx, fr->s_lines, &reinterpret_cast<const FrameJobs *>(D + fr->o_jobs)[ch.f0].sl,
                                                       n, fr->cfg, ch.caps_l, ch.smem_l)
                     : launch_stereo<FRAMES_THREADS_L>(ctx, fr->s_lines, &reinterpret_cast<const FrameJobs *>(D + fr->o_jobs)[ch.f0].sl,
                                                       n, fr->cfg, ch.caps_l, ch.smem_l)) != PLM_OK) return st;
    if (ch.f1 > t0) {
        if ((st = wide_p ? launch_f2f<512>(ctx, s, &reinterpret_cast<const FrameJobs *>(D + fr->o_jobs)[t0].fp, ch.f1 - t0,
                                           fr->cfg.best_lr, ch.smem_fp)
                         : launch_f2f<FRAMES_THREADS_P>(ctx, s, &reinterpret_cast<const FrameJobs *>(D + fr->o_jobs)[t0].fp, ch.f1 - t0,
                                                        fr->cfg.best_lr, ch.smem_fp)) != PLM_OK) return st;
        if ((st = wide_l ? launch_f2f<FRAMES_THREADS_P>(ctx, fr->s_lines, &reinterpret_cast<const FrameJobs *>(D + fr->o_jobs)[t0].fl,
                                                        ch.f1 - t0, fr->cfg.best_lr, ch.smem_fl)
                         : launch_f2f<FRAMES_THREADS_L>(ctx, fr->s_lines, &reinterpret_cast<const FrameJobs *>(D + fr->o_jobs)[t0].fl,
                                                        ch.f1 - t0, fr->cfg.best_lr, ch.smem_fl)) != PLM_OK) return st;
    }
    CU_TRY(cudaEventRecord(e_join, fr->s_lines));
    CU_TRY(cudaStreamWaitEvent(s, e_join, 0));
    return PLM_OK;
}

// Device -> host copies of the outputs of frames [f0, f1) on `s`.  The f2f vector of frame f is written by
// the job of frame f + 1, so its slot range lags by one frame: this copies frames [f0 - 1, f1 - 1), plus
// frame f1 - 1 itself (never written: -1) when it is the last frame of the upload.
int frames_copy_out(plm_frames *fr, const plm_frames_out *out, int f0, int f1, cudaStream_t s) {
    char *D = fr->d_buf;
    auto pull = [&](void *dst, size_t off, size_t elem, int64_t lo, int64_t hi) -> cudaError_t {
        if (!dst || hi <= lo) return cudaSuccess;
        fr->d2h += (hi - lo) * static_cast<int64_t>(elem);
        return cudaMemcpyAsync(static_cast<char *>(dst) + size_t(lo) * elem, D + off + size_t(lo) * elem, size_t(hi - lo) * elem,
                               cudaMemcpyDeviceToHost, s);
    };
    const std::vector<int64_t> &P = fr->lp_off, &Lo = fr->ll_off;
    const int64_t p0 = P[f0], p1 = P[f1], l0 = Lo[f0], l1 = Lo[f1];
    CU_TRY(pull(out->stereo_m12_p, fr->o_m12_p, 4, p0, p1));
    CU_TRY(pull(out->stereo_m12_l, fr->o_m12_l, 4, l0, l1));
    CU_TRY(pull(out->kept_p, fr->o_kept_p, 4, p0, p1));
    CU_TRY(pull(out->kept_l, fr->o_kept_l, 4, l0, l1));
    CU_TRY(pull(out->pt_disp, fr->o_pt_disp, 8, p0, p1));
    CU_TRY(pull(out->pt_P, fr->o_pt_P, 24, p0, p1));
    CU_TRY(pull(out->ls_disp, fr->o_ls_disp, 16, l0, l1));
    CU_TRY(pull(out->ls_sP, fr->o_ls_sP, 24, l0, l1));
    CU_TRY(pull(out->ls_eP, fr->o_ls_eP, 24, l0, l1));
    CU_TRY(pull(out->ls_le, fr->o_ls_le, 24, l0, l1));
    const int g0 = std::max(f0 - 1, 0), g1 = (f1 == fr->n_frames) ? f1 : f1 - 1;
    if (g1 > g0) {
        CU_TRY(pull(out->f2f_m12_p, fr->o_f2f_p, 4, P[g0], P[g1]));
        CU_TRY(pull(out->f2f_m12_l, fr->o_f2f_l, 4, Lo[g0], Lo[g1]));
    }
    CU_TRY(pull(out->counts, fr->o_counts, 24, f0, f1));
    return PLM_OK;
}

} // namespace

PLM_API int plm_frames_create(plm_ctx *ctx, plm_frames **out) {
    if (!out) return fail(PLM_E_INVALID, "null out");
    *out = nullptr;
    int st = resolve_ctx(ctx);
    if (st != PLM_OK) return st;
    plm_frames *f = new (std::nothrow) plm_frames();
    if (!f) return fail(PLM_E_NOMEM, "host allocation failed");
    f->ctx = ctx;
    *out = f;
    return PLM_OK;
}

PLM_API int plm_frames_destroy(plm_frames *fr) {
    if (!fr) return PLM_OK;
    if (fr->ctx) {
        cudaSetDevice(fr->ctx->device);
        cudaStreamSynchronize(fr->ctx->stream);
    }
    if (fr->s_in) {
        cudaStreamSynchronize(fr->s_in);
        cudaStreamDestroy(fr->s_in);
    }
    if (fr->s_out) {
        cudaStreamSynchronize(fr->s_out);
        cudaStreamDestroy(fr->s_out);
    }
    if (fr->s_lines) {
        cudaStreamSynchronize(fr->s_lines);
        cudaStreamDestroy(fr->s_lines);
    }
    for (cudaEvent_t e : fr->events) cudaEventDestroy(e);
    if (fr->d_buf) cudaFree(fr->d_buf);
    if (fr->h_tab) cudaFreeHost(fr->h_tab);
    delete fr;
    return PLM_OK;
}

PLM_API int64_t plm_frames_h2d_bytes(const plm_frames *fr) { return fr ? fr->h2d : 0; }
PLM_API int64_t plm_frames_d2h_bytes(const plm_frames *fr) { return fr ? fr->d2h : 0; }

PLM_API int plm_frames_upload(plm_frames *fr, const uint8_t *desc_arena, int64_t n_rows, const float *kp_arena,
                              int64_t n_kp, const float *ln_arena, int64_t n_ln, const plm_frame_rec *frames,
                              int n_frames, const plm_frame_config *c) {
    if (!fr) return fail(PLM_E_INVALID, "null frames");
    plm_ctx *ctx = fr->ctx;
    int st = resolve_ctx(ctx);
    if (st != PLM_OK) return st;
    if ((st = frames_layout(fr, desc_arena, n_rows, kp_arena, n_kp, ln_arena, n_ln, frames, n_frames, c)) != PLM_OK) return st;
    FrameChunkPrep pr = frames_chunk_prep(fr, ln_arena, frames, 0, n_frames);
    if ((st = frames_chunk_copy(fr, pr, desc_arena, kp_arena, ln_arena, ctx->stream)) != PLM_OK) return st;
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    fr->ready = true;
    return PLM_OK;
}

PLM_API int plm_frames_run(plm_frames *fr) {
    if (!fr || !fr->ready) return fail(PLM_E_INVALID, "frames not uploaded");
    plm_ctx *ctx = fr->ctx;
    int st = resolve_ctx(ctx);
    if (st != PLM_OK) return st;
    for (size_t k = 0; k < fr->chunks.size(); ++k)
        if ((st = frames_chunk_run(fr, fr->chunks[k], k)) != PLM_OK) return st;
    return PLM_OK;
}

PLM_API int plm_frames_fetch(plm_frames *fr, const plm_frames_out *out) {
    if (!fr || !fr->ready || !out) return fail(PLM_E_INVALID, "frames not uploaded / null out");
    plm_ctx *ctx = fr->ctx;
    int st = resolve_ctx(ctx);
    if (st != PLM_OK) return st;
    fr->d2h = 0;
    if (fr->n_frames > 0 && (st = frames_copy_out(fr, out, 0, fr->n_frames, ctx->stream)) != PLM_OK) return st;
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return PLM_OK;
}

PLM_API int plm_frames_process(plm_frames *fr, const uint8_t *desc_arena, int64_t n_rows, const float *kp_arena,
                               int64_t n_kp, const float *ln_arena, int64_t n_ln, const plm_frame_rec *frames,
                               int n_frames, const plm_frame_config *c, const plm_frames_out *out, int chunk_frames) {
    if (!fr || !out) return fail(PLM_E_INVALID, "null frames / out");
    plm_ctx *ctx = fr->ctx;
    int st = resolve_ctx(ctx);
    if (st != PLM_OK) return st;
    const bool trace = std::getenv("PLM_FRAMES_TRACE") != nullptr;
    if (trace && !fr->timing_events) { // events created so far carry no time stamps: start over with timing ones
        for (cudaEvent_t e : fr->events) cudaEventDestroy(e);
        fr->events.clear();
        fr->timing_events = true;
    }
    std::vector<cudaEvent_t> trace_out;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    if ((st = frames_layout(fr, desc_arena, n_rows, kp_arena, n_kp, ln_arena, n_ln, frames, n_frames, c)) != PLM_OK) return st;
    if (trace) std::fprintf(stderr, "[plm_frames_process] layout done %.3f ms\n", since());
    if (!fr->s_in) CU_TRY(cudaStreamCreateWithFlags(&fr->s_in, cudaStreamNonBlocking));
    if (!fr->s_out) CU_TRY(cudaStreamCreateWithFlags(&fr->s_out, cudaStreamNonBlocking));
    if (chunk_frames <= 0) chunk_frames = 320;
    // everything enqueued earlier on the compute stream (a previous run on the same buffers) comes first
    cudaEvent_t e_start;
    if ((st = fr->event(0, &e_start)) != PLM_OK) return st;
    CU_TRY(cudaEventRecord(e_start, ctx->stream));
    CU_TRY(cudaStreamWaitEvent(fr->s_in, e_start, 0));
    // Host preparation of the chunks (capacities, job tables) runs on a few helper threads, in chunk order,
    // ahead of this thread, which only issues the CUDA calls.
    // chunk boundaries: the first chunks are short (32, 64, 128, ... frames) so that the copy engine starts
    // after a few microseconds of host preparation instead of a full chunk's worth
    std::vector<int> bounds{0};
    for (int len = std::min(32, chunk_frames); bounds.back() < n_frames; len = std::min(2 * len, chunk_frames))
        bounds.push_back(std::min(n_frames, bounds.back() + len));
    const int n_chunks = static_cast<int>(bounds.size()) - 1;
    std::vector<FrameChunkPrep> preps(static_cast<size_t>(n_chunks));
    std::vector<std::atomic<int>> prepared(static_cast<size_t>(n_chunks));
    for (auto &p : prepared) p.store(0, std::memory_order_relaxed);
    std::atomic<int> next_chunk{0};
    auto worker = [&]() {
        for (int k = next_chunk.fetch_add(1); k < n_chunks; k = next_chunk.fetch_add(1)) {
            preps[k] = frames_chunk_prep(fr, ln_arena, frames, bounds[k], bounds[k + 1]);
            prepared[k].store(1, std::memory_order_release);
        }
    };
    const int n_helpers = std::max(0, std::min({n_chunks - 1, 6, static_cast<int>(std::thread::hardware_concurrency()) - 1}));
    std::vector<std::thread> helpers;
    for (int i = 0; i < n_helpers; ++i) helpers.emplace_back(worker);
    if (n_helpers == 0) worker();
    const int out_group = std::max(1, g_frames_out_group);
    int out_first = 0;
    for (int k = 0; k < n_chunks; ++k) {
        const int f0 = bounds[k], f1 = bounds[k + 1];
        cudaEvent_t e_in, e_run;
        if ((st = fr->event(4 * size_t(k) + 1, &e_in)) != PLM_OK) break;
        if ((st = fr->event(4 * size_t(k) + 2, &e_run)) != PLM_OK) break;
        while (!prepared[k].load(std::memory_order_acquire)) std::this_thread::yield();
        if ((st = frames_chunk_copy(fr, preps[k], desc_arena, kp_arena, ln_arena, fr->s_in)) != PLM_OK) break;
        if (cudaEventRecord(e_in, fr->s_in) != cudaSuccess || cudaStreamWaitEvent(ctx->stream, e_in, 0) != cudaSuccess) {
            st = fail(PLM_E_CUDA, "event record / wait failed");
            break;
        }
        if ((st = frames_chunk_run(fr, fr->chunks.back(), size_t(k))) != PLM_OK) break;
        if (cudaEventRecord(e_run, ctx->stream) != cudaSuccess || cudaStreamWaitEvent(fr->s_out, e_run, 0) != cudaSuccess) {
            st = fail(PLM_E_CUDA, "event record / wait failed");
            break;
        }
        // copy-out in groups of chunks: every output array is one copy per group (13 fixed copy costs per group); the
        // last chunks go out one by one so that the tail after the last kernel stays short
        (void)f0;
        const int group = (k + 3 >= n_chunks) ? 1 : out_group;
        if (k + 1 - out_first >= group || k + 1 == n_chunks) {
            if ((st = frames_copy_out(fr, out, bounds[out_first], f1, fr->s_out)) != PLM_OK) break;
            out_first = k + 1;
        }
        if (trace) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, fr->s_out);
            trace_out.push_back(e);
            std::fprintf(stderr, "[plm_frames_process] chunk %d issued %.3f ms\n", k, since());
        }
    }
    next_chunk.store(n_chunks); // after a failure: helpers stop picking up chunks
    for (std::thread &t : helpers) t.join();
    // drain all three streams even after a failure: the caller's buffers must not be in flight on return
    cudaStreamSynchronize(fr->s_in);
    if (trace) std::fprintf(stderr, "[plm_frames_process] copy-in stream drained %.3f ms\n", since());
    cudaStreamSynchronize(ctx->stream);
    if (fr->s_lines) cudaStreamSynchronize(fr->s_lines);
    if (trace) std::fprintf(stderr, "[plm_frames_process] compute streams drained %.3f ms\n", since());
    cudaStreamSynchronize(fr->s_out);
    if (trace) {
        std::fprintf(stderr, "[plm_frames_process] copy-out stream drained %.3f ms\n", since());
        for (size_t k = 0; k < trace_out.size(); ++k) { // device timeline of every chunk, relative to the start event
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, fr->events[0], fr->events[4 * k + 1]);
            cudaEventElapsedTime(&b, fr->events[0], fr->events[4 * k + 2]);
            cudaEventElapsedTime(&c, fr->events[0], trace_out[k]);
            std::fprintf(stderr, "[plm_frames_process] chunk %zu frames %d..%d: copied in %.3f ms, computed %.3f ms, copied out %.3f ms\n", k,
                         bounds[k], bounds[k + 1], a, b, c);
            cudaEventDestroy(trace_out[k]);
        }
    }
    if (st != PLM_OK) return st;
    CU_TRY(cudaGetLastError());
    fr->ready = true;
    return PLM_OK;
}
