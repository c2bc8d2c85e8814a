// Host side of the device-resident stereo-frame pipeline (plm_frames_*; included by plmatch.cu).

struct plm_frames {
    plm_ctx *ctx = nullptr;
    char *d_buf = nullptr;
    size_t d_cap = 0;
    int n_frames = 0;
    int64_t NP = 0, NL = 0; // left point / line slots
    plm::FrameCfg cfg;
    float nnr_p = 0.f, nnr_l = 0.f;
    plm::StereoCaps caps_p{1, 1, 1, 8}, caps_l{1, 1, 1, 4};
    size_t smem_p = 0, smem_l = 0, smem_fp = 0, smem_fl = 0;
    size_t o_sjobs_p = 0, o_sjobs_l = 0, o_fjobs_p = 0, o_fjobs_l = 0;
    size_t o_m12_p = 0, o_m12_l = 0, o_kept_p = 0, o_kept_l = 0, o_pt_disp = 0, o_pt_P = 0, o_ls_disp = 0, o_ls_sP = 0,
           o_ls_eP = 0, o_ls_le = 0, o_f2f_p = 0, o_f2f_l = 0, o_counts = 0;
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

template <int THREADS>
int launch_stereo(plm_ctx *ctx, const plm::StereoJob *jobs, int n_jobs, const plm::FrameCfg &cfg,
                  const plm::StereoCaps &caps, size_t smem) {
    if (n_jobs == 0) return PLM_OK;
    CU_TRY(cudaFuncSetAttribute(plm::stereo_frame_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
    plm::stereo_frame_kernel<THREADS><<<n_jobs, THREADS, smem, ctx->stream>>>(jobs, cfg, caps);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
    return PLM_OK;
}

template <int THREADS>
int launch_f2f(plm_ctx *ctx, const plm::F2FJob *jobs, int n_jobs, int best_lr, size_t smem) {
    if (n_jobs == 0) return PLM_OK;
    CU_TRY(cudaFuncSetAttribute(plm::f2f_match_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
    plm::f2f_match_kernel<THREADS><<<n_jobs, THREADS, smem, ctx->stream>>>(jobs, best_lr);
    ctx->launches++;
    CU_TRY(cudaGetLastError());
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
    if (fr->d_buf) cudaFree(fr->d_buf);
    delete fr;
    return PLM_OK;
}

PLM_API int64_t plm_frames_h2d_bytes(const plm_frames *fr) { return fr ? fr->h2d : 0; }
PLM_API int64_t plm_frames_d2h_bytes(const plm_frames *fr) { return fr ? fr->d2h : 0; }

PLM_API int plm_frames_upload(plm_frames *fr, const uint8_t *desc_arena, int64_t n_rows, const float *kp_arena,
                              int64_t n_kp, const float *ln_arena, int64_t n_ln, const plm_frame_rec *frames,
                              int n_frames, const plm_frame_config *c) {
    if (!fr || !c) return fail(PLM_E_INVALID, "null frames / config");
    if (n_rows < 0 || n_kp < 0 || n_ln < 0 || n_frames < 0) return fail(PLM_E_INVALID, "negative size");
    if ((n_rows > 0 && !desc_arena) || (n_kp > 0 && !kp_arena) || (n_ln > 0 && !ln_arena) || (n_frames > 0 && !frames))
        return fail(PLM_E_INVALID, "null pointer");
    if (c->grid_rows <= 0 || c->grid_cols <= 0) return fail(PLM_E_GRID, "[GridStructure] invalid dimension");
    if (static_cast<long long>(c->grid_rows) * c->grid_cols > (1 << 14)) return fail(PLM_E_UNSUPPORTED, "grid too large for the frame pipeline");
    if (c->min_ratio_12p > 1.0) return fail(PLM_E_RATIO, plm_status_string(PLM_E_RATIO));
    plm_ctx *ctx = fr->ctx;
    int st = resolve_ctx(ctx);
    if (st != PLM_OK) return st;
    fr->ready = false;

    // ---- validation, slot offsets, shared-memory capacities --------------------------------------------
    std::vector<int64_t> lp_off(static_cast<size_t>(n_frames) + 1, 0), ll_off(static_cast<size_t>(n_frames) + 1, 0);
    int cap_pl = 1, cap_pr = 1, cap_ll = 1, cap_lr = 1;
    long long cap_items_l = 1;
    const long long walk_max = static_cast<long long>(c->grid_rows) + c->grid_cols + 2;
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
        lp_off[f + 1] = lp_off[f] + r.n_pl;
        ll_off[f + 1] = ll_off[f] + r.n_ll;
        cap_pl = std::max(cap_pl, r.n_pl);
        cap_pr = std::max(cap_pr, r.n_pr);
        cap_ll = std::max(cap_ll, r.n_ll);
        cap_lr = std::max(cap_lr, r.n_lr);
        long long items = 0;
        for (int j = 0; j < r.n_lr; ++j) {
            const float *l = ln_arena + 4 * (r.ln_r + j);
            const long long n = line_walk_cells(double(l[0]) * c->inv_width, double(l[1]) * c->inv_height,
                                                double(l[2]) * c->inv_width, double(l[3]) * c->inv_height);
            if (n > walk_max) return fail(PLM_E_UNSUPPORTED, "line segment far outside the image (Bresenham walk too long)");
            items += n;
        }
        cap_items_l = std::max(cap_items_l, items);
    }
    const int n_cells = c->grid_rows * c->grid_cols;
    fr->caps_p = plm::StereoCaps{cap_pl, cap_pr, cap_pr, FRAMES_THREADS_P / 32};
    fr->caps_l = plm::StereoCaps{cap_ll, cap_lr, static_cast<int>(cap_items_l), FRAMES_THREADS_L / 32};
    fr->smem_p = plm::stereo_frame_smem(fr->caps_p, n_cells, false);
    fr->smem_l = plm::stereo_frame_smem(fr->caps_l, n_cells, true);
    fr->smem_fp = plm::f2f_smem(cap_pl, cap_pl);
    fr->smem_fl = plm::f2f_smem(cap_ll, cap_ll);
    const size_t budget = ctx->smem_optin - 1024;
    if (fr->smem_p > budget || fr->smem_l > budget || fr->smem_fp > budget || fr->smem_fl > budget)
        return fail(PLM_E_UNSUPPORTED, "frame too large for shared memory");

    const int64_t NP = lp_off[n_frames], NL = ll_off[n_frames];
    Layout L;
    const size_t o_desc = L.add(size_t(n_rows) * 32);
    const size_t o_kp = L.add(size_t(n_kp) * 8);
    const size_t o_ln = L.add(size_t(n_ln) * 16);
    const size_t o_cdesc_p = L.add(size_t(NP) * 32), o_cdesc_l = L.add(size_t(NL) * 32);
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
    fr->o_counts = L.add(size_t(std::max(n_frames, 1)) * 6 * 4);
    fr->o_sjobs_p = L.add(size_t(std::max(n_frames, 1)) * sizeof(plm::StereoJob));
    fr->o_sjobs_l = L.add(size_t(std::max(n_frames, 1)) * sizeof(plm::StereoJob));
    fr->o_fjobs_p = L.add(size_t(std::max(n_frames, 1)) * sizeof(plm::F2FJob));
    fr->o_fjobs_l = L.add(size_t(std::max(n_frames, 1)) * sizeof(plm::F2FJob));
    if ((st = fr->ensure(L.total)) != PLM_OK) return st;
    char *D = fr->d_buf;

    // ---- job tables ------------------------------------------------------------------------------------
    const float nnr_p = static_cast<float>(c->min_ratio_12p), nnr_l = static_cast<float>(c->min_ratio_12l);
    std::vector<plm::StereoJob> sp(static_cast<size_t>(n_frames)), sl(static_cast<size_t>(n_frames));
    std::vector<plm::F2FJob> fp(static_cast<size_t>(std::max(n_frames - 1, 0))), fl(fp.size());
    const uint4 *d_desc = reinterpret_cast<const uint4 *>(D + o_desc);
    const float *d_kp = reinterpret_cast<const float *>(D + o_kp);
    const float *d_ln = reinterpret_cast<const float *>(D + o_ln);
    int32_t *d_counts = reinterpret_cast<int32_t *>(D + fr->o_counts);
    for (int f = 0; f < n_frames; ++f) {
        const plm_frame_rec &r = frames[f];
        plm::StereoJob &a = sp[f];
        std::memset(&a, 0, sizeof(a));
        a.geo_l = d_kp + 2 * r.kp_l;
        a.geo_r = d_kp + 2 * r.kp_r;
        a.d_l = d_desc + 2 * r.desc_pl;
        a.d_r = d_desc + 2 * r.desc_pr;
        a.m12 = reinterpret_cast<int32_t *>(D + fr->o_m12_p) + lp_off[f];
        a.cdesc = reinterpret_cast<uint4 *>(D + o_cdesc_p) + 2 * lp_off[f];
        a.kept_i1 = reinterpret_cast<int32_t *>(D + fr->o_kept_p) + lp_off[f];
        a.o0 = reinterpret_cast<double *>(D + fr->o_pt_disp) + lp_off[f];
        a.o1 = reinterpret_cast<double *>(D + fr->o_pt_P) + 3 * lp_off[f];
        a.counts = d_counts + 6 * f;
        a.n_l = r.n_pl;
        a.n_r = r.n_pr;
        a.is_lines = 0;
        plm::StereoJob &b = sl[f];
        std::memset(&b, 0, sizeof(b));
        b.geo_l = d_ln + 4 * r.ln_l;
        b.geo_r = d_ln + 4 * r.ln_r;
        b.d_l = d_desc + 2 * r.desc_ll;
        b.d_r = d_desc + 2 * r.desc_lr;
        b.m12 = reinterpret_cast<int32_t *>(D + fr->o_m12_l) + ll_off[f];
        b.cdesc = reinterpret_cast<uint4 *>(D + o_cdesc_l) + 2 * ll_off[f];
        b.kept_i1 = reinterpret_cast<int32_t *>(D + fr->o_kept_l) + ll_off[f];
        b.o0 = reinterpret_cast<double *>(D + fr->o_ls_disp) + 2 * ll_off[f];
        b.o1 = reinterpret_cast<double *>(D + fr->o_ls_sP) + 3 * ll_off[f];
        b.o2 = reinterpret_cast<double *>(D + fr->o_ls_eP) + 3 * ll_off[f];
        b.o3 = reinterpret_cast<double *>(D + fr->o_ls_le) + 3 * ll_off[f];
        b.counts = d_counts + 6 * f + 2;
        b.n_l = r.n_ll;
        b.n_r = r.n_lr;
        b.is_lines = 1;
        if (f >= 1) {
            plm::F2FJob &p = fp[f - 1];
            std::memset(&p, 0, sizeof(p));
            p.d1 = sp[f - 1].cdesc;
            p.d2 = a.cdesc;
            p.n1_ptr = d_counts + 6 * (f - 1) + 1;
            p.n2_ptr = d_counts + 6 * f + 1;
            p.m12 = reinterpret_cast<int32_t *>(D + fr->o_f2f_p) + lp_off[f - 1];
            p.count = d_counts + 6 * f + 4;
            p.cap1 = frames[f - 1].n_pl;
            p.cap2 = r.n_pl;
            p.nnr = nnr_p;
            plm::F2FJob &q = fl[f - 1];
            std::memset(&q, 0, sizeof(q));
            q.d1 = sl[f - 1].cdesc;
            q.d2 = b.cdesc;
            q.n1_ptr = d_counts + 6 * (f - 1) + 3;
            q.n2_ptr = d_counts + 6 * f + 3;
            q.m12 = reinterpret_cast<int32_t *>(D + fr->o_f2f_l) + ll_off[f - 1];
            q.count = d_counts + 6 * f + 5;
            q.cap1 = frames[f - 1].n_ll;
            q.cap2 = r.n_ll;
            q.nnr = nnr_l;
        }
    }

    Layout S;
    const size_t s_sp = S.add(sp.size() * sizeof(plm::StereoJob)), s_sl = S.add(sl.size() * sizeof(plm::StereoJob));
    const size_t s_fp = S.add(fp.size() * sizeof(plm::F2FJob)), s_fl = S.add(fl.size() * sizeof(plm::F2FJob));
    if ((st = ctx->ensure_pinned(std::max<size_t>(S.total, 256))) != PLM_OK) return st;
    char *H = ctx->h_buf;
    if (!sp.empty()) std::memcpy(H + s_sp, sp.data(), sp.size() * sizeof(plm::StereoJob));
    if (!sl.empty()) std::memcpy(H + s_sl, sl.data(), sl.size() * sizeof(plm::StereoJob));
    if (!fp.empty()) std::memcpy(H + s_fp, fp.data(), fp.size() * sizeof(plm::F2FJob));
    if (!fl.empty()) std::memcpy(H + s_fl, fl.data(), fl.size() * sizeof(plm::F2FJob));
    cudaStream_t s = ctx->stream;
    if (n_rows > 0) CU_TRY(cudaMemcpyAsync(D + o_desc, desc_arena, size_t(n_rows) * 32, cudaMemcpyHostToDevice, s));
    if (n_kp > 0) CU_TRY(cudaMemcpyAsync(D + o_kp, kp_arena, size_t(n_kp) * 8, cudaMemcpyHostToDevice, s));
    if (n_ln > 0) CU_TRY(cudaMemcpyAsync(D + o_ln, ln_arena, size_t(n_ln) * 16, cudaMemcpyHostToDevice, s));
    if (!sp.empty()) CU_TRY(cudaMemcpyAsync(D + fr->o_sjobs_p, H + s_sp, sp.size() * sizeof(plm::StereoJob), cudaMemcpyHostToDevice, s));
    if (!sl.empty()) CU_TRY(cudaMemcpyAsync(D + fr->o_sjobs_l, H + s_sl, sl.size() * sizeof(plm::StereoJob), cudaMemcpyHostToDevice, s));
    if (!fp.empty()) CU_TRY(cudaMemcpyAsync(D + fr->o_fjobs_p, H + s_fp, fp.size() * sizeof(plm::F2FJob), cudaMemcpyHostToDevice, s));
    if (!fl.empty()) CU_TRY(cudaMemcpyAsync(D + fr->o_fjobs_l, H + s_fl, fl.size() * sizeof(plm::F2FJob), cudaMemcpyHostToDevice, s));
    CU_TRY(cudaStreamSynchronize(s)); // the pinned staging block is reused by the next call

    fr->n_frames = n_frames;
    fr->NP = NP;
    fr->NL = NL;
    fr->nnr_p = nnr_p;
    fr->nnr_l = nnr_l;
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
    fr->h2d = static_cast<int64_t>(size_t(n_rows) * 32 + size_t(n_kp) * 8 + size_t(n_ln) * 16 + S.total);
    fr->d2h = 0;
    fr->ready = true;
    return PLM_OK;
}

PLM_API int plm_frames_run(plm_frames *fr) {
    if (!fr || !fr->ready) return fail(PLM_E_INVALID, "frames not uploaded");
    plm_ctx *ctx = fr->ctx;
    int st = resolve_ctx(ctx);
    if (st != PLM_OK) return st;
    if (fr->n_frames == 0) return PLM_OK;
    char *D = fr->d_buf;
    const int F = fr->n_frames;
    CU_TRY(cudaMemsetAsync(D + fr->o_counts, 0, size_t(F) * 6 * 4, ctx->stream));
    // slots past a frame's kept count (and the last frame's f2f range) read as -1
    if (fr->NP) CU_TRY(cudaMemsetAsync(D + fr->o_kept_p, 0xFF, size_t(fr->NP) * 4, ctx->stream));
    if (fr->NL) CU_TRY(cudaMemsetAsync(D + fr->o_kept_l, 0xFF, size_t(fr->NL) * 4, ctx->stream));
    if (fr->NP) CU_TRY(cudaMemsetAsync(D + fr->o_f2f_p, 0xFF, size_t(fr->NP) * 4, ctx->stream));
    if (fr->NL) CU_TRY(cudaMemsetAsync(D + fr->o_f2f_l, 0xFF, size_t(fr->NL) * 4, ctx->stream));
    if ((st = launch_stereo<FRAMES_THREADS_P>(ctx, reinterpret_cast<const plm::StereoJob *>(D + fr->o_sjobs_p), F, fr->cfg,
                                              fr->caps_p, fr->smem_p)) != PLM_OK) return st;
    if ((st = launch_stereo<FRAMES_THREADS_L>(ctx, reinterpret_cast<const plm::StereoJob *>(D + fr->o_sjobs_l), F, fr->cfg,
                                              fr->caps_l, fr->smem_l)) != PLM_OK) return st;
    if ((st = launch_f2f<FRAMES_THREADS_P>(ctx, reinterpret_cast<const plm::F2FJob *>(D + fr->o_fjobs_p), F - 1,
                                           fr->cfg.best_lr, fr->smem_fp)) != PLM_OK) return st;
    return launch_f2f<FRAMES_THREADS_L>(ctx, reinterpret_cast<const plm::F2FJob *>(D + fr->o_fjobs_l), F - 1,
                                        fr->cfg.best_lr, fr->smem_fl);
}

PLM_API int plm_frames_fetch(plm_frames *fr, const plm_frames_out *out) {
    if (!fr || !fr->ready || !out) return fail(PLM_E_INVALID, "frames not uploaded / null out");
    plm_ctx *ctx = fr->ctx;
    int st = resolve_ctx(ctx);
    if (st != PLM_OK) return st;
    cudaStream_t s = ctx->stream;
    char *D = fr->d_buf;
    int64_t bytes = 0;
    auto pull = [&](void *dst, size_t off, size_t n) -> cudaError_t {
        if (!dst || n == 0) return cudaSuccess;
        bytes += static_cast<int64_t>(n);
        return cudaMemcpyAsync(dst, D + off, n, cudaMemcpyDeviceToHost, s);
    };
    const size_t NP = size_t(fr->NP), NL = size_t(fr->NL);
    CU_TRY(pull(out->stereo_m12_p, fr->o_m12_p, NP * 4));
    CU_TRY(pull(out->stereo_m12_l, fr->o_m12_l, NL * 4));
    CU_TRY(pull(out->kept_p, fr->o_kept_p, NP * 4));
    CU_TRY(pull(out->kept_l, fr->o_kept_l, NL * 4));
    CU_TRY(pull(out->pt_disp, fr->o_pt_disp, NP * 8));
    CU_TRY(pull(out->pt_P, fr->o_pt_P, NP * 24));
    CU_TRY(pull(out->ls_disp, fr->o_ls_disp, NL * 16));
    CU_TRY(pull(out->ls_sP, fr->o_ls_sP, NL * 24));
    CU_TRY(pull(out->ls_eP, fr->o_ls_eP, NL * 24));
    CU_TRY(pull(out->ls_le, fr->o_ls_le, NL * 24));
    CU_TRY(pull(out->f2f_m12_p, fr->o_f2f_p, NP * 4));
    CU_TRY(pull(out->f2f_m12_l, fr->o_f2f_l, NL * 4));
    CU_TRY(pull(out->counts, fr->o_counts, size_t(fr->n_frames) * 6 * 4));
    CU_TRY(cudaStreamSynchronize(s));
    fr->d2h = bytes;
    return PLM_OK;
}
