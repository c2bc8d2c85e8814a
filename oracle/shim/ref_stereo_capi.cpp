// extern "C" entry points over the UNMODIFIED reference sources
//   /root/reference/stvo-pl/src/{stereoFrame,stereoFeatures,pinholeStereoCamera}.cpp
// (plus matching / gridStructure / lineIterator, compiled against the same owning cv::Mat stand-in) so that tests can
// drive StereoFrame::matchStereoPoints / matchStereoLines / filterLineSegmentDisparity / lineSegmentOverlapStereo /
// lineSegmentOverlap and PinholeStereoCamera::backProjection through ctypes.  Test infrastructure only.
#include <stereoFrame.h>

#include <cstdint>
#include <cstring>

#define PLREF_API extern "C" __attribute__((visibility("default")))

namespace {

cv::Mat owned_desc(const uint8_t *d, int n) {
    cv::Mat m;
    if (n > 0) {
        m.create(n, 32, CV_8U);
        std::memcpy(m.data, d, static_cast<size_t>(n) * 32);
    }
    return m;
}

PinholeStereoCamera *make_cam(int w, int h, const double *cam) {
    return new PinholeStereoCamera(w, h, cam[0], cam[1], cam[2], cam[3], cam[4]);
}

} // namespace

PLREF_API void plref_set_config(int best_lr, int lr_parallel, double min_ratio_12p, double line_sim_th) {
    Config::bestLRMatches() = best_lr != 0;
    Config::lrInParallel() = lr_parallel != 0;
    Config::minRatio12P() = min_ratio_12p;
    Config::lineSimTh() = line_sim_th;
}

PLREF_API void plref_set_stereo_config(int matching_s_ws, double max_dist_epip, double min_disp, double line_horiz_th,
                                       double stereo_overlap_th, double ls_min_disp_ratio, double orb_scale_factor,
                                       double lsd_scale) {
    Config::matchingSWs() = matching_s_ws;
    Config::maxDistEpip() = max_dist_epip;
    Config::minDisp() = min_disp;
    Config::lineHorizTh() = line_horiz_th;
    Config::stereoOverlapTh() = stereo_overlap_th;
    Config::lsMinDispRatio() = ls_min_disp_ratio;
    Config::orbScaleFactor() = orb_scale_factor;
    Config::lsdScale() = lsd_scale;
}

// StereoFrame::matchStereoPoints (stereoFrame.cpp:131-184) called the way detectStereoPoints does (:110): the member
// descriptors are both the matchGrid input and the in/out argument.  Returns the number of stereo points; every
// output array has room for n_l entries.
PLREF_API int plref_stereo_points(const float *kp_l, const int32_t *oct_l, const uint8_t *d_l, int n_l, const float *kp_r,
                                  const uint8_t *d_r, int n_r, int img_w, int img_h, const double *cam, int initial,
                                  double *pl, double *disp, double *P, int32_t *idx, int32_t *level, double *sigma2,
                                  uint8_t *desc_out, int32_t *n_desc_out) {
    PinholeStereoCamera *c = make_cam(img_w, img_h, cam);
    int n = -1;
    {
        cv::Mat img(img_h, img_w, CV_8UC1, nullptr);
        StVO::StereoFrame f(img, img, initial ? 0 : 1, c);
        f.points_l.resize(n_l);
        f.points_r.resize(n_r);
        for (int i = 0; i < n_l; i++) {
            f.points_l[i].pt = cv::Point2f(kp_l[2 * i], kp_l[2 * i + 1]);
            f.points_l[i].octave = oct_l ? oct_l[i] : 0;
        }
        for (int i = 0; i < n_r; i++) f.points_r[i].pt = cv::Point2f(kp_r[2 * i], kp_r[2 * i + 1]);
        f.pdesc_l = owned_desc(d_l, n_l);
        f.pdesc_r = owned_desc(d_r, n_r);
        f.matchStereoPoints(f.points_l, f.points_r, f.pdesc_l, f.pdesc_r, initial != 0);
        n = static_cast<int>(f.stereo_pt.size());
        for (int k = 0; k < n; k++) {
            const StVO::PointFeature *p = f.stereo_pt[k];
            pl[2 * k] = p->pl(0);
            pl[2 * k + 1] = p->pl(1);
            disp[k] = p->disp;
            for (int a = 0; a < 3; a++) P[3 * k + a] = p->P(a);
            idx[k] = p->idx;
            level[k] = p->level;
            sigma2[k] = p->sigma2;
        }
        *n_desc_out = f.pdesc_l.rows;
        for (int r = 0; r < f.pdesc_l.rows; r++) std::memcpy(desc_out + static_cast<size_t>(r) * 32, f.pdesc_l.ptr<uint8_t>(r), 32);
    }
    delete c;
    return n;
}

// StereoFrame::matchStereoLines (stereoFrame.cpp:320-409).  ln_* are n x 4 floats (startPointX, startPointY, endPointX,
// endPointY); angle_l / oct_l are KeyLine::angle / ::octave of the left lines.
PLREF_API int plref_stereo_lines(const float *ln_l, const float *angle_l, const int32_t *oct_l, const uint8_t *d_l, int n_l,
                                 const float *ln_r, const uint8_t *d_r, int n_r, int img_w, int img_h, const double *cam,
                                 int initial, double *spl, double *epl, double *disp_se, double *sP, double *eP, double *le,
                                 double *angle, int32_t *idx, int32_t *level, double *sigma2, uint8_t *desc_out,
                                 int32_t *n_desc_out) {
    PinholeStereoCamera *c = make_cam(img_w, img_h, cam);
    int n = -1;
    {
        cv::Mat img(img_h, img_w, CV_8UC1, nullptr);
        StVO::StereoFrame f(img, img, initial ? 0 : 1, c);
        f.lines_l.resize(n_l);
        f.lines_r.resize(n_r);
        for (int i = 0; i < n_l; i++) {
            KeyLine &k = f.lines_l[i];
            k.startPointX = ln_l[4 * i];
            k.startPointY = ln_l[4 * i + 1];
            k.endPointX = ln_l[4 * i + 2];
            k.endPointY = ln_l[4 * i + 3];
            k.angle = angle_l ? angle_l[i] : 0.f;
            k.octave = oct_l ? oct_l[i] : 0;
        }
        for (int i = 0; i < n_r; i++) {
            KeyLine &k = f.lines_r[i];
            k.startPointX = ln_r[4 * i];
            k.startPointY = ln_r[4 * i + 1];
            k.endPointX = ln_r[4 * i + 2];
            k.endPointY = ln_r[4 * i + 3];
        }
        f.ldesc_l = owned_desc(d_l, n_l);
        f.ldesc_r = owned_desc(d_r, n_r);
        f.matchStereoLines(f.lines_l, f.lines_r, f.ldesc_l, f.ldesc_r, initial != 0);
        n = static_cast<int>(f.stereo_ls.size());
        for (int k = 0; k < n; k++) {
            const StVO::LineFeature *p = f.stereo_ls[k];
            for (int a = 0; a < 2; a++) {
                spl[2 * k + a] = p->spl(a);
                epl[2 * k + a] = p->epl(a);
            }
            disp_se[2 * k] = p->sdisp;
            disp_se[2 * k + 1] = p->edisp;
            for (int a = 0; a < 3; a++) {
                sP[3 * k + a] = p->sP(a);
                eP[3 * k + a] = p->eP(a);
                le[3 * k + a] = p->le(a);
            }
            angle[k] = p->angle;
            idx[k] = p->idx;
            level[k] = p->level;
            sigma2[k] = p->sigma2;
        }
        *n_desc_out = f.ldesc_l.rows;
        for (int r = 0; r < f.ldesc_l.rows; r++) std::memcpy(desc_out + static_cast<size_t>(r) * 32, f.ldesc_l.ptr<uint8_t>(r), 32);
    }
    delete c;
    return n;
}

// scalar helpers, one call per element (stereoFrame.cpp:416-426, :484-519, :521-627, :829-840; pinholeStereoCamera.cpp:229-237)
PLREF_API void plref_filter_line_disparity(int n, const double *spl, const double *epl, const double *spr, const double *epr,
                                           double *disp_se) {
    StVO::StereoFrame f;
    for (int i = 0; i < n; i++) {
        double ds = 0, de = 0;
        f.filterLineSegmentDisparity(Vector2d(spl[2 * i], spl[2 * i + 1]), Vector2d(epl[2 * i], epl[2 * i + 1]),
                                     Vector2d(spr[2 * i], spr[2 * i + 1]), Vector2d(epr[2 * i], epr[2 * i + 1]), ds, de);
        disp_se[2 * i] = ds;
        disp_se[2 * i + 1] = de;
    }
}

PLREF_API void plref_filter_disparity_pair(int n, double *disp_se) {
    StVO::StereoFrame f;
    for (int i = 0; i < n; i++) f.filterLineSegmentDisparity(disp_se[2 * i], disp_se[2 * i + 1]);
}

PLREF_API void plref_line_overlap_stereo(int n, const double *v4, double *out) {
    StVO::StereoFrame f;
    for (int i = 0; i < n; i++) out[i] = f.lineSegmentOverlapStereo(v4[4 * i], v4[4 * i + 1], v4[4 * i + 2], v4[4 * i + 3]);
}

// v8 = (spl_obs.x, spl_obs.y, epl_obs.x, epl_obs.y, spl_proj.x, spl_proj.y, epl_proj.x, epl_proj.y)
PLREF_API void plref_line_overlap(int n, const double *v8, double *out) {
    StVO::StereoFrame f;
    for (int i = 0; i < n; i++) {
        const double *v = v8 + 8 * i;
        out[i] = f.lineSegmentOverlap(Vector2d(v[0], v[1]), Vector2d(v[2], v[3]), Vector2d(v[4], v[5]), Vector2d(v[6], v[7]));
    }
}

PLREF_API void plref_back_projection(int img_w, int img_h, const double *cam, int n, const double *uvd, double *P) {
    PinholeStereoCamera *c = make_cam(img_w, img_h, cam);
    for (int i = 0; i < n; i++) {
        const Vector3d p = c->backProjection(uvd[3 * i], uvd[3 * i + 1], uvd[3 * i + 2]);
        for (int a = 0; a < 3; a++) P[3 * i + a] = p(a);
    }
    delete c;
}
