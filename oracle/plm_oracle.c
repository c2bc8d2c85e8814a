/*
 * plm_oracle.c -- CPU restatement of the PL-inertial-slam descriptor-matching path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and there only as the checker (or the CPU arm), never as the thing shipped.
 *
 * Parity status: the reference holds no golden vectors / known-answer tests for this path
 * (SURVEY.md section 4).  This restatement is therefore pinned against the reference ITSELF where its
 * sources compile here (oracle/Makefile, outputs in oracle/_ref/):
 *   - libplref.so = /root/reference/stvo-pl/src/{matching,gridStructure,lineIterator}.cpp and
 *     /root/reference/src/mapFeatures.cpp compiled unmodified  -> plo_hamming256, plo_match*, plo_match_grid*,
 *     plo_line_coords, plo_med_desc (tests/test_oracle.py, tests/test_mapfeatures.py);
 *   - libplref_dbow.so = the reference's vendored DBoW2 (3rdparty/DBoW2) compiled unmodified
 *     -> plo_bow_word, plo_bow_transform, plo_bow_score (tests/test_bow.py);
 *   - libplref_stereo.so = /root/reference/stvo-pl/src/{stereoFrame,stereoFeatures,pinholeStereoCamera}.cpp (+ matching /
 *     gridStructure / lineIterator) compiled unmodified against the stand-in headers under oracle/shim_stereo/
 *     -> plo_stereo_points, plo_stereo_lines (grid fill, matchGrid, gates, compaction, back-projection),
 *     plo_stereo_filter_*, plo_csr_from_*, plo_line_overlap_stereo, plo_line_segment_overlap / plo_line_pair_filter
 *     (tests/test_ref_stereo.py; tests/golden/stereo.npz);
 *   tests/golden/ holds vectors generated from those builds (tools/make_golden.py) for machines without
 *   /root/reference.
 * PARITY UNPINNED (restated from the source only, no reference build possible here):
 *   - cv::BFMatcher::knnMatch (OpenCV 3.3, features2d; call site matching.cpp:47-48) is not in
 *     /root/reference: its K=2 batchDistance insertion rule is restated in plo_knn2() and cross-checked
 *     against the in-container cv2 4.13 wheel (tests/test_oracle.py::test_knn2_vs_cv2).
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PLO_API __attribute__((visibility("default")))

/* Built with -ffp-contract=off: the reference is compiled for baseline x86-64 (no FMA,
 * CMakeLists.txt:41), so every product that feeds a comparison is rounded separately. */

/* std::min / std::max semantics (NaN behaviour differs from fmin/fmax). */
static inline double plo_min(double a, double b) { return (b < a) ? b : a; }
static inline double plo_max(double a, double b) { return (a < b) ? b : a; }

/* stvo-pl/src/matching.cpp:93-109  StVO::distance -- 8 x int32 xor + SWAR popcount. */
PLO_API int plo_hamming256(const uint8_t *a, const uint8_t *b)
{
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t wa, wb;
        memcpy(&wa, a + 4 * i, 4);
        memcpy(&wb, b + 4 * i, 4);
        uint32_t v = wa ^ wb;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

/*
 * cv::BFMatcher::knnMatch(k=2) as called at stvo-pl/src/matching.cpp:47-48.  OpenCV's
 * batchDistance (K>0 branch) keeps, per query row, K (dist, idx) slots initialised to
 * (INT_MAX, -1); train row j with distance d enters iff d < dist[K-1] and is shifted up while
 * dist[k] > d  ==> lexicographic (distance, train index) ordering, lowest index first on ties.
 * idx/dist are n1 x 2; absent slots stay (-1, INT_MAX).
 */
PLO_API void plo_knn2(const uint8_t *d1, int n1, size_t step1, const uint8_t *d2, int n2,
                      size_t step2, int32_t *idx, int32_t *dist)
{
    for (int i = 0; i < n1; i++) {
        int32_t bd[2] = {INT_MAX, INT_MAX};
        int32_t bi[2] = {-1, -1};
        const uint8_t *q = d1 + (size_t)i * step1;
        for (int j = 0; j < n2; j++) {
            int d = plo_hamming256(q, d2 + (size_t)j * step2);
            if (d < bd[1]) {
                int k = 0;
                if (bd[0] > d) { bd[1] = bd[0]; bi[1] = bi[0]; k = -1; }
                bd[k + 1] = d;
                bi[k + 1] = j;
            }
        }
        idx[2 * i] = bi[0]; idx[2 * i + 1] = bi[1];
        dist[2 * i] = bd[0]; dist[2 * i + 1] = bd[1];
    }
}

/*
 * stvo-pl/src/matching.cpp:41-61  StVO::matchNNR.
 * m12 is IN/OUT: the reference does matches_12.resize(n1,-1), which only initialises slots the
 * caller's vector did not already have; here the caller passes the already-resized buffer.
 * Acceptance test is evaluated in float (DMatch::distance is float, nnr is float) -- :54.
 * Returns the number of accepted rows, or -1 for the n2 < 2 case (UB in the reference).
 */
PLO_API int plo_match_nnr(const uint8_t *d1, int n1, size_t step1, const uint8_t *d2, int n2,
                          size_t step2, float nnr, int32_t *m12)
{
    if (n2 < 2) return -1;
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(n1 > 0 ? n1 : 1));
    int32_t *dist = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(n1 > 0 ? n1 : 1));
    plo_knn2(d1, n1, step1, d2, n2, step2, idx, dist);
    int matches = 0;
    for (int i = 0; i < n1; i++) {
        const float lhs = (float)dist[2 * i];
        const float rhs = (float)dist[2 * i + 1] * nnr;
        if (lhs < rhs) {
            m12[i] = idx[2 * i];
            matches++;
        }
    }
    free(idx);
    free(dist);
    return matches;
}

/*
 * stvo-pl/src/matching.cpp:63-91  StVO::match.  best_lr = Config::bestLRMatches().
 * The mutual check walks EVERY entry >= 0 of m12, stale ones included (:80-86), so the
 * returned count can undershoot (even go negative) on the stale-fallback call sites
 * (src/mapHandler.cpp:325-329).  Stale entries must be < n2 (the reference indexes
 * matches_21[i2] unchecked).
 */
PLO_API int plo_match(const uint8_t *d1, int n1, size_t step1, const uint8_t *d2, int n2,
                      size_t step2, float nnr, int best_lr, int32_t *m12)
{
    if (!best_lr) return plo_match_nnr(d1, n1, step1, d2, n2, step2, nnr, m12);
    if (n2 < 2 || n1 < 2) return INT_MIN;
    int matches = plo_match_nnr(d1, n1, step1, d2, n2, step2, nnr, m12);
    int32_t *m21 = (int32_t *)malloc(sizeof(int32_t) * (size_t)n2);
    for (int j = 0; j < n2; j++) m21[j] = -1;
    plo_match_nnr(d2, n2, step2, d1, n1, step1, nnr, m21);
    for (int i1 = 0; i1 < n1; i1++) {
        int i2 = m12[i1];
        if (i2 >= 0 && m21[i2] != i1) {
            m12[i1] = -1;
            matches--;
        }
    }
    free(m21);
    return matches;
}

/* ---------------------------------------------------------------------------------------
 * Grid (CSR form of StVO::GridStructure).  Cell (x, y), 0 <= x < cols, 0 <= y < rows, has
 * linear id  x*rows + y  -- the reference stores grid[x][y] with x outermost
 * (stvo-pl/src/gridStructure.cpp:49).  cell_start has cols*rows+1 entries.
 * ------------------------------------------------------------------------------------- */

/*
 * stvo-pl/src/lineIterator.cpp:34-77 + stvo-pl/src/gridStructure.cpp:33-41 getLineCoords:
 * Bresenham from double endpoints.  NOTE the reference never initialises y/x/maxX from the
 * *rounded* endpoints: y = static_cast<int>(y1), x = static_cast<int>(x1),
 * maxX = static_cast<int>(x2) after the steep / order swaps (:50-55).
 * Writes up to max_cells (x, y) pairs, returns the number of cells of the full walk.
 */
PLO_API int plo_line_coords(double x1, double y1, double x2, double y2, int32_t *cells,
                            int max_cells)
{
    const int steep = fabs(y2 - y1) > fabs(x2 - x1);
    double t;
    if (steep) { t = x1; x1 = y1; y1 = t; t = x2; x2 = y2; y2 = t; }
    if (x1 > x2) { t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
    const double dx = x2 - x1;
    const double dy = fabs(y2 - y1);
    double error = dx / 2.0;
    const int ystep = (y1 < y2) ? 1 : -1;
    int y = (int)y1;
    int x = (int)x1;
    const int maxX = (int)x2;
    int n = 0;
    while (x <= maxX) {
        if (n < max_cells) {
            cells[2 * n] = steep ? y : x;
            cells[2 * n + 1] = steep ? x : y;
        }
        n++;
        error -= dy;
        if (error < 0) { y += ystep; error += dx; }
        x++;
    }
    return n;
}

/* GridStructure::get (stvo-pl/src/gridStructure.cpp:65-76): clamp the window to the grid. */
static void plo_window(int x, int y, const int32_t win[4], int rows, int cols, int *min_x,
                       int *max_x, int *min_y, int *max_y)
{
    int a = x - win[0]; *min_x = a > 0 ? a : 0;
    int b = x + win[1] + 1; *max_x = b < cols ? b : cols;
    int c = y - win[2]; *min_y = c > 0 ? c : 0;
    int d = y + win[3] + 1; *max_y = d < rows ? d : rows;
}

typedef struct {
    int32_t *list;  /* de-duplicated candidate ids (the unordered_set)            */
    int32_t *stamp; /* stamp[id - lo] == tag  <=> id already in the set           */
    int32_t lo, hi; /* id range covered by stamp                                   */
    int n;
} plo_set;

static void plo_set_add_window(plo_set *s, int tag, int x, int y, const int32_t win[4],
                               const int32_t *cell_start, const int32_t *cell_items, int rows,
                               int cols)
{
    int min_x, max_x, min_y, max_y;
    plo_window(x, y, win, rows, cols, &min_x, &max_x, &min_y, &max_y);
    for (int x_ = min_x; x_ < max_x; ++x_)
        for (int y_ = min_y; y_ < max_y; ++y_) {
            int c = x_ * rows + y_;
            for (int k = cell_start[c]; k < cell_start[c + 1]; k++) {
                int id = cell_items[k];
                if (id < s->lo || id >= s->hi) { /* out-of-range ids: keep once each is moot,
                                                    the matcher skips them (:141) */
                    continue;
                }
                if (s->stamp[id - s->lo] != tag) {
                    s->stamp[id - s->lo] = tag;
                    s->list[s->n++] = id;
                }
            }
        }
}

/*
 * Shared body of both StVO::matchGrid overloads (stvo-pl/src/matching.cpp:111-177 points,
 * :179-258 lines).  is_lines selects the two-window candidate union (:213-215) and the
 * direction filter (:221-222, NaN passes because !(NaN < th)).
 * ratio = Config::minRatio12P() for BOTH overloads (:160, :241), evaluated in double.
 * The unordered_set iteration order of the reference is unspecified; for ratio <= 1 an accepted
 * row has a strict unique minimum so the order cannot change any output (SURVEY 8a note 1).
 */
static int plo_match_grid_ex(int is_lines, const int32_t *coords, const uint8_t *d1, int n1,
                             size_t step1, const int32_t *cell_start, const int32_t *cell_items,
                             int rows, int cols, const uint8_t *d2, int n2, size_t step2,
                             const double *dirs2, double line_sim_th, const int32_t win[4],
                             double ratio, int best_lr, int32_t *m12, const uint16_t *seed,
                             int do_cross, int64_t i1_base, uint16_t *colmin_out,
                             uint64_t *m21key_out);

static int plo_match_grid(int is_lines, const int32_t *coords, const uint8_t *d1, int n1,
                          size_t step1, const int32_t *cell_start, const int32_t *cell_items,
                          int rows, int cols, const uint8_t *d2, int n2, size_t step2,
                          const double *dirs2, double line_sim_th, const int32_t win[4],
                          double ratio, int best_lr, int32_t *m12)
{
    return plo_match_grid_ex(is_lines, coords, d1, n1, step1, cell_start, cell_items, rows, cols,
                             d2, n2, step2, dirs2, line_sim_th, win, ratio, best_lr, m12, NULL, 1,
                             0, NULL, NULL);
}

/*
 * The loop above with three hooks for ROW-SHARDED execution: `seed` initialises
 * distances[] (the running column minima left behind by the rows of lower-ranked shards; 0xFFFF =
 * INT_MAX), do_cross = 0 defers the final mutual check (it needs every shard's m21), and the
 * per-column state is exported: colmin_out = final distances[] (0xFFFF = INT_MAX), m21key_out =
 * (distance << 32 | i1_base + i1) of the last live pair of this shard (UINT64_MAX = none).
 * With seed == NULL, do_cross == 1 it is exactly the reference function.
 */
static int plo_match_grid_ex(int is_lines, const int32_t *coords, const uint8_t *d1, int n1,
                             size_t step1, const int32_t *cell_start, const int32_t *cell_items,
                             int rows, int cols, const uint8_t *d2, int n2, size_t step2,
                             const double *dirs2, double line_sim_th, const int32_t win[4],
                             double ratio, int best_lr, int32_t *m12, const uint16_t *seed,
                             int do_cross, int64_t i1_base, uint16_t *colmin_out,
                             uint64_t *m21key_out)
{
    int matches = 0;
    int32_t *m21 = NULL, *distances = NULL;
    if (best_lr) {
        m21 = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n2 > 0 ? n2 : 1));
        distances = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n2 > 0 ? n2 : 1));
        for (int j = 0; j < n2; j++) {
            m21[j] = -1;
            distances[j] = (seed && seed[j] != 0xFFFFu) ? (int)seed[j] : INT_MAX;
            if (m21key_out) m21key_out[j] = UINT64_MAX;
        }
    }
    const int n_items = cell_start[rows * cols];
    plo_set set;
    set.lo = 0;
    set.hi = n2;
    set.list = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n2 > 0 ? n2 : 1));
    set.stamp = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n2 > 0 ? n2 : 1));
    for (int j = 0; j < n2; j++) set.stamp[j] = -1;
    (void)n_items;

    for (int i1 = 0; i1 < n1; ++i1) {
        int best_d = INT_MAX, best_d2 = INT_MAX, best_idx = -1;
        const uint8_t *desc = d1 + (size_t)i1 * step1;
        double vx = 0.0, vy = 0.0;
        set.n = 0;
        if (!is_lines) {
            plo_set_add_window(&set, i1, coords[2 * i1], coords[2 * i1 + 1], win, cell_start,
                               cell_items, rows, cols);
        } else {
            const int32_t *c = coords + 4 * i1;
            vx = (double)(c[2] - c[0]);
            vy = (double)(c[3] - c[1]);
            /* matching.h:43-48 normalize: magnitude = sqrt(x*x + y*y), no FMA contraction */
            const double xx = vx * vx, yy = vy * vy;
            double mag = sqrt(xx + yy);
            vx /= mag;
            vy /= mag;
            plo_set_add_window(&set, i1, c[0], c[1], win, cell_start, cell_items, rows, cols);
            plo_set_add_window(&set, i1, c[2], c[3], win, cell_start, cell_items, rows, cols);
        }
        /* `if (candidates.empty()) continue;` (:139/:217): ids outside [0,n2) would make the
         * set non-empty but are skipped below, and an all-skipped row is never accepted for
         * ratio <= 1, so dropping them from the set is equivalent. */
        if (set.n == 0) continue;
        for (int k = 0; k < set.n; k++) {
            const int i2 = set.list[k];
            if (is_lines) {
                const double p0 = vx * dirs2[2 * i2], p1 = vy * dirs2[2 * i2 + 1];
                double dp = p0 + p1;
                if (fabs(dp) < line_sim_th) continue;
            }
            const int d = plo_hamming256(desc, d2 + (size_t)i2 * step2);
            if (best_lr) {
                if (d < distances[i2]) {
                    distances[i2] = d;
                    m21[i2] = i1;
                    if (m21key_out)
                        m21key_out[i2] = ((uint64_t)d << 32) | (uint64_t)(i1_base + i1);
                } else
                    continue;
            }
            if (d < best_d) {
                best_d2 = best_d;
                best_d = d;
                best_idx = i2;
            } else if (d < best_d2)
                best_d2 = d;
        }
        if ((double)best_d < (double)best_d2 * ratio) {
            m12[i1] = best_idx;
            matches++;
        }
    }
    if (best_lr && colmin_out)
        for (int j = 0; j < n2; j++)
            colmin_out[j] = distances[j] == INT_MAX ? 0xFFFFu : (uint16_t)distances[j];
    if (best_lr && do_cross) {
        for (int i1 = 0; i1 < n1; ++i1) {
            int i2 = m12[i1];
            if (i2 >= 0 && m21[i2] != i1) {
                m12[i1] = -1;
                matches--;
            }
        }
    }
    free(m21);
    free(distances);
    free(set.list);
    free(set.stamp);
    return matches;
}

/* stvo-pl/src/matching.cpp:111-177.  xy = n1 x (x, y) grid-cell coords; win = {width.first,
 * width.second, height.first, height.second} of GridWindow (gridStructure.h:35-37). */
PLO_API int plo_match_grid_points(const int32_t *xy, const uint8_t *d1, int n1, size_t step1,
                                  const int32_t *cell_start, const int32_t *cell_items,
                                  int rows, int cols, const uint8_t *d2, int n2, size_t step2,
                                  const int32_t win[4], double ratio, int best_lr,
                                  int32_t *m12)
{
    return plo_match_grid(0, xy, d1, n1, step1, cell_start, cell_items, rows, cols, d2, n2,
                          step2, NULL, 0.0, win, ratio, best_lr, m12);
}

/* stvo-pl/src/matching.cpp:179-258.  xyxy = n1 x (sx, sy, ex, ey); dirs2 = n2 x (dx, dy). */
PLO_API int plo_match_grid_lines(const int32_t *xyxy, const uint8_t *d1, int n1, size_t step1,
                                 const int32_t *cell_start, const int32_t *cell_items,
                                 int rows, int cols, const uint8_t *d2, int n2, size_t step2,
                                 const double *dirs2, double line_sim_th,
                                 const int32_t win[4], double ratio, int best_lr,
                                 int32_t *m12)
{
    return plo_match_grid(1, xyxy, d1, n1, step1, cell_start, cell_items, rows, cols, d2, n2,
                          step2, dirs2, line_sim_th, win, ratio, best_lr, m12);
}

/* Row-shard form of both matchGrid overloads (tests of the multi-GPU orchestration). */
PLO_API int plo_match_grid_shard(int is_lines, const int32_t *coords, const uint8_t *d1, int n1,
                                 size_t step1, int64_t i1_base, const int32_t *cell_start,
                                 const int32_t *cell_items, int rows, int cols, const uint8_t *d2,
                                 int n2, size_t step2, const double *dirs2, double line_sim_th,
                                 const int32_t win[4], double ratio, int best_lr, int32_t *m12,
                                 const uint16_t *seed, uint16_t *colmin_out, uint64_t *m21key_out)
{
    return plo_match_grid_ex(is_lines, coords, d1, n1, step1, cell_start, cell_items, rows, cols,
                             d2, n2, step2, dirs2, line_sim_th, win, ratio, best_lr, m12, seed, 0,
                             i1_base, colmin_out, m21key_out);
}

/* ---------------------------------------------------------------------------------------
 * Stereo post-filters (the drivers' per-match geometry gates).
 * ------------------------------------------------------------------------------------- */

/*
 * stvo-pl/src/stereoFrame.cpp:162-171 (matchStereoPoints): a match (i1 -> i2) survives iff
 * |y_l - y_r| <= maxDistEpip (float subtraction, std::abs(float) promoted to double for the
 * compare) and disp = x_l - x_r (float subtraction widened to double) >= minDisp.
 * kp_* are n x (x, y) float32 pixel coordinates (cv::KeyPoint::pt).
 * keep[i1] = 1/0, disp[i1] = disparity of kept rows (else 0).  Returns the number kept.
 */
PLO_API int plo_stereo_filter_points(const float *kp_l, const float *kp_r, const int32_t *m12,
                                     int n1, double max_dist_epip, double min_disp,
                                     uint8_t *keep, double *disp)
{
    int kept = 0;
    for (int i1 = 0; i1 < n1; i1++) {
        keep[i1] = 0;
        disp[i1] = 0.0;
        const int i2 = m12[i1];
        if (i2 < 0) continue;
        const float dyf = kp_l[2 * i1 + 1] - kp_r[2 * i2 + 1];
        if ((double)fabsf(dyf) <= max_dist_epip) {
            const float dxf = kp_l[2 * i1] - kp_r[2 * i2];
            double disp_ = (double)dxf;
            if (disp_ >= min_disp) {
                keep[i1] = 1;
                disp[i1] = disp_;
                kept++;
            }
        }
    }
    return kept;
}

/* stvo-pl/src/stereoFrame.cpp:484-519  lineSegmentOverlapStereo (all double). */
PLO_API double plo_line_overlap_stereo(double spl_obs, double epl_obs, double spl_proj,
                                       double epl_proj, double line_horiz_th)
{
    double overlap = 1.f;
    if (fabs(epl_obs - spl_obs) > line_horiz_th) {
        double sln = plo_min(spl_obs, epl_obs);
        double eln = plo_max(spl_obs, epl_obs);
        double spn = plo_min(spl_proj, epl_proj);
        double epn = plo_max(spl_proj, epl_proj);
        double length = eln - spn;
        if ((epn < sln) || (spn > eln))
            overlap = 0.f;
        else {
            if ((epn > eln) && (spn < sln))
                overlap = eln - sln;
            else
                overlap = plo_min(eln, epn) - plo_max(sln, spn);
        }
        if (length > 0.01f)
            overlap = overlap / length;
        else
            overlap = 0.f;
        if (overlap > 1.f) overlap = 1.f;
    }
    return overlap;
}

/*
 * stvo-pl/src/stereoFrame.cpp:359-385 (matchStereoLines) + :416-426 filterLineSegmentDisparity.
 * ln_* are n x (sx, sy, ex, ey) float32 (cv::line_descriptor::KeyLine start/endPoint), widened
 * to double on entry exactly like the Vector3d initialisation at :366-370.
 * Quirk kept: sp_r is overwritten (x interpolated at y = sp_l.y, :377) BEFORE ep_r is
 * interpolated (:378), so the second interpolation and the |sp_r.y - ep_r.y| gate (:384) see
 * the updated values.
 * keep[i1] = 1/0; disp_se[2*i1 .. +1] = (disp_s, disp_e) as computed (including the -1/-1
 * rejection marker) for rows with a match, else (0,0).  Returns the number kept.
 */
PLO_API int plo_stereo_filter_lines(const float *ln_l, const float *ln_r, const int32_t *m12,
                                    int n1, double min_disp, double line_horiz_th,
                                    double stereo_overlap_th, double ls_min_disp_ratio,
                                    uint8_t *keep, double *disp_se)
{
    int kept = 0;
    for (int i1 = 0; i1 < n1; i1++) {
        keep[i1] = 0;
        disp_se[2 * i1] = disp_se[2 * i1 + 1] = 0.0;
        const int i2 = m12[i1];
        if (i2 < 0) continue;
        double sp_l[2] = {ln_l[4 * i1 + 0], ln_l[4 * i1 + 1]};
        double ep_l[2] = {ln_l[4 * i1 + 2], ln_l[4 * i1 + 3]};
        double sp_r[2] = {ln_r[4 * i2 + 0], ln_r[4 * i2 + 1]};
        double ep_r[2] = {ln_r[4 * i2 + 2], ln_r[4 * i2 + 3]};

        double overlap =
            plo_line_overlap_stereo(sp_l[1], ep_l[1], sp_r[1], ep_r[1], line_horiz_th);

        {
            const double a = sp_r[0] * (sp_l[1] - ep_r[1]);
            const double b = ep_r[0] * (sp_r[1] - sp_l[1]);
            double nx = (a + b) / (sp_r[1] - ep_r[1]);
            sp_r[0] = nx;
            sp_r[1] = sp_l[1];
        }
        {
            const double a = sp_r[0] * (ep_l[1] - ep_r[1]);
            const double b = ep_r[0] * (sp_r[1] - ep_l[1]);
            double nx = (a + b) / (sp_r[1] - ep_r[1]);
            ep_r[0] = nx;
            ep_r[1] = ep_l[1];
        }
        double disp_s = sp_l[0] - sp_r[0];
        double disp_e = ep_l[0] - ep_r[0];
        if (plo_min(disp_s, disp_e) / plo_max(disp_s, disp_e) < ls_min_disp_ratio) {
            disp_s = -1.0;
            disp_e = -1.0;
        }
        disp_se[2 * i1] = disp_s;
        disp_se[2 * i1 + 1] = disp_e;
        if (disp_s >= min_disp && disp_e >= min_disp &&
            fabs(sp_l[1] - ep_l[1]) > line_horiz_th &&
            fabs(sp_r[1] - ep_r[1]) > line_horiz_th && overlap > stereo_overlap_th) {
            keep[i1] = 1;
            kept++;
        }
    }
    return kept;
}

/* ---------------------------------------------------------------------------------------
 * Whole stereo drivers (StereoFrame::matchStereoPoints / matchStereoLines): grid fill,
 * matchGrid, geometry gates, compaction of the kept rows and back-projection.  These follow
 * stvo-pl/src/stereoFrame.cpp, which cannot be compiled here (OpenCV + Eigen + line_descriptor),
 * so unlike the matching functions above they are pinned by restatement only.
 * ------------------------------------------------------------------------------------- */

/* C++ double -> int conversion (truncation toward zero) of an already scaled coordinate. */
static inline int plo_trunc(double v) { return (int)v; }

/* grid.at(kp.x * inv_width, kp.y * inv_height).push_back(idx), stereoFrame.cpp:146-150 with
 * GridStructure::at (gridStructure.cpp:56-63: off-grid pushes go to a sink list).  kp = n x (x, y)
 * float pixel coordinates; float * double is evaluated in double.  CSR out. */
PLO_API int plo_csr_from_points(const float *kp, int n, double inv_w, double inv_h, int rows,
                                int cols, int32_t *cell_start, int32_t *cell_items)
{
    const int n_cells = rows * cols;
    int32_t *cnt = (int32_t *)calloc((size_t)n_cells + 1, sizeof(int32_t));
    int32_t *cid = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        const int x = plo_trunc((double)kp[2 * i] * inv_w), y = plo_trunc((double)kp[2 * i + 1] * inv_h);
        cid[i] = (x >= 0 && x < cols && y >= 0 && y < rows) ? x * rows + y : -1;
        if (cid[i] >= 0) cnt[cid[i]]++;
    }
    cell_start[0] = 0;
    for (int c = 0; c < n_cells; c++) cell_start[c + 1] = cell_start[c] + cnt[c];
    memset(cnt, 0, sizeof(int32_t) * (size_t)n_cells);
    for (int i = 0; i < n; i++)
        if (cid[i] >= 0) cell_items[cell_start[cid[i]] + cnt[cid[i]]++] = i;
    free(cnt);
    free(cid);
    return cell_start[n_cells];
}

/* stereoFrame.cpp:336-349: directions (:342-344, float subtraction, * double, matching.h:43-48
 * normalize) and the Bresenham grid fill of every right line (getLineCoords, :346-348).
 * ln = n x (sx, sy, ex, ey) float.  Returns the item count (cell_items needs that many slots;
 * call with cell_items == NULL to size it). */
PLO_API int plo_csr_from_lines(const float *ln, int n, double inv_w, double inv_h, int rows, int cols,
                               int32_t *cell_start, int32_t *cell_items, double *dirs)
{
    const int n_cells = rows * cols;
    int32_t *cnt = (int32_t *)calloc((size_t)n_cells + 1, sizeof(int32_t));
    const int cap = rows + cols + 4;
    int32_t *cells = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)cap * 4);
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            cell_start[0] = 0;
            for (int c = 0; c < n_cells; c++) cell_start[c + 1] = cell_start[c] + cnt[c];
            memset(cnt, 0, sizeof(int32_t) * (size_t)n_cells);
            if (!cell_items) break;
        }
        for (int i = 0; i < n; i++) {
            const float *l = ln + 4 * i;
            if (pass == 0 && dirs) {
                double vx = (double)(l[2] - l[0]) * inv_w, vy = (double)(l[3] - l[1]) * inv_h;
                const double xx = vx * vx, yy = vy * vy;
                const double mag = sqrt(xx + yy);
                dirs[2 * i] = vx / mag;
                dirs[2 * i + 1] = vy / mag;
            }
            int m = plo_line_coords((double)l[0] * inv_w, (double)l[1] * inv_h, (double)l[2] * inv_w,
                                    (double)l[3] * inv_h, cells, cap * 4);
            if (m > cap * 4) m = cap * 4; /* cannot happen for on-image lines */
            for (int k = 0; k < m; k++) {
                const int x = cells[2 * k], y = cells[2 * k + 1];
                if (x < 0 || x >= cols || y < 0 || y >= rows) continue;
                const int c = x * rows + y;
                if (pass == 0) cnt[c]++;
                else cell_items[cell_start[c] + cnt[c]++] = i;
            }
        }
    }
    const int total = cell_start[n_cells];
    free(cnt);
    free(cells);
    return total;
}

/* PinholeStereoCamera::backProjection (stvo-pl/src/pinholeStereoCamera.cpp:229-237).
 * cam = {b, fx, cx, cy}. */
static void plo_back_projection(const double cam[4], double u, double v, double disp, double P[3])
{
    const double bd = cam[0] / disp;
    P[0] = bd * (u - cam[2]);
    P[1] = bd * (v - cam[3]);
    P[2] = bd * cam[1];
}

/*
 * StereoFrame::matchStereoPoints (stvo-pl/src/stereoFrame.cpp:131-184).  Outputs: m12 (n_l, the
 * matchGrid vector, starts at -1), and for the kept rows in i1 order (= the order of stereo_pt and
 * of the compacted pdesc_l, :172-178): kept_i1, disp, P (3 per row).  Returns the number kept, or
 * 0 with m12 untouched when either side is empty (:137-138).
 */
PLO_API int plo_stereo_points(const float *kp_l, const uint8_t *d_l, int n_l, const float *kp_r,
                              const uint8_t *d_r, int n_r, double inv_w, double inv_h, int rows, int cols,
                              int matching_s_ws, double ratio, int best_lr, double max_dist_epip,
                              double min_disp, const double cam[4], int32_t *m12, int32_t *kept_i1,
                              double *disp, double *P)
{
    if (n_l <= 0 || n_r <= 0) return 0;
    const int n_cells = rows * cols;
    int32_t *xy = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)n_l);
    int32_t *cs = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n_cells + 1));
    int32_t *ci = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_r);
    for (int i = 0; i < n_l; i++) { /* :142-143, pair<double,double> -> pair<int,int> */
        xy[2 * i] = plo_trunc((double)kp_l[2 * i] * inv_w);
        xy[2 * i + 1] = plo_trunc((double)kp_l[2 * i + 1] * inv_h);
        m12[i] = -1;
    }
    plo_csr_from_points(kp_r, n_r, inv_w, inv_h, rows, cols, cs, ci);
    const int32_t win[4] = {matching_s_ws, 0, 0, 0}; /* :152-154 */
    plo_match_grid(0, xy, d_l, n_l, 32, cs, ci, rows, cols, d_r, n_r, 32, NULL, 0.0, win, ratio, best_lr, m12);
    int kept = 0;
    for (int i1 = 0; i1 < n_l; i1++) {
        const int i2 = m12[i1];
        if (i2 < 0) continue;
        const float dyf = kp_l[2 * i1 + 1] - kp_r[2 * i2 + 1];
        if (!((double)fabsf(dyf) <= max_dist_epip)) continue;
        const double disp_ = (double)(kp_l[2 * i1] - kp_r[2 * i2]);
        if (!(disp_ >= min_disp)) continue;
        kept_i1[kept] = i1;
        disp[kept] = disp_;
        plo_back_projection(cam, (double)kp_l[2 * i1], (double)kp_l[2 * i1 + 1], disp_, P + 3 * kept);
        kept++;
    }
    free(xy);
    free(cs);
    free(ci);
    return kept;
}

/*
 * StereoFrame::matchStereoLines (stvo-pl/src/stereoFrame.cpp:320-409).  Per kept row: kept_i1,
 * disp_se (2), sP / eP (3 each), le (3) = normalised left line equation (:368: cross product of the
 * homogeneous endpoints divided by sqrt(le0^2 + le1^2)).
 */
PLO_API int plo_stereo_lines(const float *ln_l, const uint8_t *d_l, int n_l, const float *ln_r,
                             const uint8_t *d_r, int n_r, double inv_w, double inv_h, int rows, int cols,
                             int matching_s_ws, double ratio, double line_sim_th, int best_lr,
                             double min_disp, double line_horiz_th, double stereo_overlap_th,
                             double ls_min_disp_ratio, const double cam[4], int32_t *m12,
                             int32_t *kept_i1, double *disp_se, double *sP, double *eP, double *le)
{
    if (n_l <= 0 || n_r <= 0) return 0;
    const int n_cells = rows * cols;
    int32_t *xyxy = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)n_l);
    int32_t *cs = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n_cells + 1));
    double *dirs = (double *)malloc(sizeof(double) * 2 * (size_t)n_r);
    for (int i = 0; i < n_l; i++) { /* :330-333 */
        xyxy[4 * i + 0] = plo_trunc((double)ln_l[4 * i + 0] * inv_w);
        xyxy[4 * i + 1] = plo_trunc((double)ln_l[4 * i + 1] * inv_h);
        xyxy[4 * i + 2] = plo_trunc((double)ln_l[4 * i + 2] * inv_w);
        xyxy[4 * i + 3] = plo_trunc((double)ln_l[4 * i + 3] * inv_h);
        m12[i] = -1;
    }
    const int n_items = plo_csr_from_lines(ln_r, n_r, inv_w, inv_h, rows, cols, cs, NULL, NULL);
    int32_t *ci = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_items > 0 ? n_items : 1));
    plo_csr_from_lines(ln_r, n_r, inv_w, inv_h, rows, cols, cs, ci, dirs);
    const int32_t win[4] = {matching_s_ws, 0, 0, 0}; /* :351-353 */
    plo_match_grid(1, xyxy, d_l, n_l, 32, cs, ci, rows, cols, d_r, n_r, 32, dirs, line_sim_th, win, ratio,
                   best_lr, m12);
    uint8_t *keep = (uint8_t *)malloc((size_t)n_l);
    double *dse = (double *)malloc(sizeof(double) * 2 * (size_t)n_l);
    plo_stereo_filter_lines(ln_l, ln_r, m12, n_l, min_disp, line_horiz_th, stereo_overlap_th, ls_min_disp_ratio,
                            keep, dse);
    int kept = 0;
    for (int i1 = 0; i1 < n_l; i1++) {
        if (!keep[i1]) continue;
        const double x1 = ln_l[4 * i1], y1 = ln_l[4 * i1 + 1], x2 = ln_l[4 * i1 + 2], y2 = ln_l[4 * i1 + 3];
        /* Eigen cross of (x1, y1, 1) and (x2, y2, 1), each product rounded separately */
        const double c0 = y1 * 1.0 - 1.0 * y2;
        const double c1 = 1.0 * x2 - x1 * 1.0;
        const double p0 = x1 * y2, p1 = y1 * x2;
        const double c2 = p0 - p1;
        const double q0 = c0 * c0, q1 = c1 * c1;
        const double nrm = sqrt(q0 + q1);
        kept_i1[kept] = i1;
        disp_se[2 * kept] = dse[2 * i1];
        disp_se[2 * kept + 1] = dse[2 * i1 + 1];
        plo_back_projection(cam, x1, y1, dse[2 * i1], sP + 3 * kept);
        plo_back_projection(cam, x2, y2, dse[2 * i1 + 1], eP + 3 * kept);
        le[3 * kept] = c0 / nrm;
        le[3 * kept + 1] = c1 / nrm;
        le[3 * kept + 2] = c2 / nrm;
        kept++;
    }
    free(xyxy); free(cs); free(ci); free(dirs); free(keep); free(dse);
    return kept;
}

/* ------------------------------------------------------------------------------------------
 * Map landmarks: representative descriptor and mean observation direction.
 *
 * src/mapFeatures.cpp:51-93 (PLSLAM::MapPoint::updateAverageDescDir) and :121-163
 * (PLSLAM::MapLine::updateAverageDescDir) -- the two bodies are identical.  For a landmark with
 * n observations:
 *   - conf_desc(i,j) = cv::norm(desc_i, desc_j, NORM_HAMMING) (:59-68; 256-bit popcount of the
 *     xor, symmetric, zero diagonal, stored as float and read back as int -- exact for 0..256);
 *   - per row i the distances are sorted and dist[int(1 + 0.5 * (n - 1))] is its "median"
 *     (:75-79; the self distance 0 takes part);
 *   - the first row with the strictly smallest median wins (:80-84, start value 99999);
 *   - med_obs_dir = (sum of dir_list in list order) / n (:88-91).  The reference adds into an
 *     uninitialised Eigen vector (:88); this restatement starts the sum at zero, which is what
 *     the code means and what the compiled reference (zero-initialising Eigen stand-in,
 *     oracle/shim_map/) does.
 * A landmark is created with one observation (ctor, :29-40: med_desc = that descriptor,
 * med_obs_dir = its direction) and updateAverageDescDir only runs from n = 2 on (:42-50); for
 * n = 1 the index int(1 + 0) would be out of range, so n = 1 follows the constructor.
 *
 * Batched over landmarks: observations of landmark l are rows obs_start[l] .. obs_start[l+1]-1
 * of desc (n_obs x 32, `step` bytes apart) and of dirs (n_obs x 3, may be NULL).  Outputs:
 * med_idx[l] = winning position inside the landmark's list (-1 for an empty list), med_desc
 * (n_lm x 32, may be NULL), med_dir (n_lm x 3, may be NULL).
 */
static int plo_cmp_int(const void *a, const void *b)
{
    const int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

PLO_API void plo_med_desc(const uint8_t *desc, size_t step, const double *dirs, const int32_t *obs_start,
                          int n_lm, int32_t *med_idx, uint8_t *med_desc, double *med_dir)
{
    for (int l = 0; l < n_lm; l++) {
        const int lo = obs_start[l], n = obs_start[l + 1] - lo;
        int best = n > 0 ? 0 : -1;
        if (n >= 2) {
            int *conf = (int *)malloc(sizeof(int) * (size_t)n * (size_t)n);
            int *row = (int *)malloc(sizeof(int) * (size_t)n);
            for (int i = 0; i < n; i++) {
                conf[(size_t)i * n + i] = 0;
                for (int j = i + 1; j < n; j++) {
                    const int d = plo_hamming256(desc + (size_t)(lo + i) * step, desc + (size_t)(lo + j) * step);
                    conf[(size_t)i * n + j] = d;
                    conf[(size_t)j * n + i] = d;
                }
            }
            int max_dist = 99999;
            for (int i = 0; i < n; i++) {
                memcpy(row, conf + (size_t)i * n, sizeof(int) * (size_t)n);
                qsort(row, (size_t)n, sizeof(int), plo_cmp_int);
                const int med = row[(int)(1 + 0.5 * (n - 1))];
                if (med < max_dist) {
                    max_dist = med;
                    best = i;
                }
            }
            free(conf);
            free(row);
        }
        med_idx[l] = best;
        if (med_desc) {
            if (best >= 0) memcpy(med_desc + (size_t)l * 32, desc + (size_t)(lo + best) * step, 32);
            else memset(med_desc + (size_t)l * 32, 0, 32);
        }
        if (med_dir && dirs) {
            double s[3] = {0.0, 0.0, 0.0};
            for (int i = 0; i < n; i++)
                for (int c = 0; c < 3; c++) s[c] += dirs[3 * (size_t)(lo + i) + c];
            for (int c = 0; c < 3; c++) /* n == 1: the constructor's plain copy (:38) */
                med_dir[3 * (size_t)l + c] = (n >= 2) ? s[c] / n : (n == 1 ? dirs[3 * (size_t)lo + c] : 0.0);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Bag-of-words loop-candidate scoring (the reference's vendored DBoW2, 3rdparty/DBoW2).
 *
 * The vocabulary is a flat copy of TemplatedVocabulary::m_nodes (include/DBoW2/TemplatedVocabulary.h:
 * 275-307): node 0 = root, children of node i = child_ids[child_start[i] .. child_start[i+1]-1] in the
 * reference's vector order, node_desc 32 bytes per node, node_weight, node_word >= 0 on leaves.
 *
 * plo_bow_word     TemplatedVocabulary::transform(feature, word_id, weight)  (:1196-1238): from the root,
 *                  move to the child with the smallest FORB::distance (src/DBoW2/FORB.cpp:78-100, 256-bit
 *                  Hamming), FIRST child on ties (strict `<`, :1225), until a leaf.
 * plo_bow_transform  TemplatedVocabulary::transform(features, BowVector)  (:1045-1101) for weighting
 *                  TF_IDF / TF (0 / 1): v[word] += weight per feature in feature order (BowVector::addWeight,
 *                  src/DBoW2/BowVector.cpp:31-43; the same weight is added once per occurrence, so the value
 *                  is the weight summed `count` times), IDF / BINARY (2 / 3): v[word] = weight
 *                  (addIfNotExist, :47-55); words with weight <= 0 are dropped (:1074, :1094); then
 *                  L1 normalisation (BowVector::normalize, :59-81: sum of fabs in word order, each /= norm)
 *                  -- L1_NORM scoring, the DBoW2 default, asks for it (ScoringObject.h:73).
 * plo_bow_score    L1Scoring::score (src/DBoW2/ScoringObject.cpp:25-69): over the common words in word
 *                  order, score += fabs(vi - wi) - fabs(vi) - fabs(wi); result -score / 2.
 * Call sites: src/mapHandler.cpp:3125-3137, :3150-3162, :3176-3235.
 */
PLO_API int plo_bow_word(const int32_t *child_start, const int32_t *child_ids, const uint8_t *node_desc,
                         const uint8_t *feature)
{
    int node = 0;
    if (child_start[1] == child_start[0]) return -1; /* empty vocabulary */
    do {
        const int c0 = child_start[node], c1 = child_start[node + 1];
        int best = child_ids[c0];
        int best_d = plo_hamming256(feature, node_desc + 32 * (size_t)best);
        for (int c = c0 + 1; c < c1; c++) {
            const int id = child_ids[c];
            const int d = plo_hamming256(feature, node_desc + 32 * (size_t)id);
            if (d < best_d) {
                best_d = d;
                best = id;
            }
        }
        node = best;
    } while (child_start[node + 1] > child_start[node]);
    return node; /* leaf node id */
}

typedef struct { uint32_t word; int32_t order; double w; } plo_bow_item;

static int plo_cmp_bow(const void *a, const void *b)
{
    const plo_bow_item *x = (const plo_bow_item *)a, *y = (const plo_bow_item *)b;
    if (x->word != y->word) return (x->word > y->word) - (x->word < y->word);
    return (x->order > y->order) - (x->order < y->order);
}

/* -> number of entries written to ids / vals (capacity n) */
PLO_API int plo_bow_transform(const int32_t *child_start, const int32_t *child_ids, const uint8_t *node_desc,
                              const double *node_weight, const int32_t *node_word, int weighting,
                              const uint8_t *desc, int n, size_t step, uint32_t *ids, double *vals)
{
    if (n <= 0 || child_start[1] == child_start[0]) return 0;
    plo_bow_item *it = (plo_bow_item *)malloc(sizeof(plo_bow_item) * (size_t)n);
    int m = 0;
    for (int f = 0; f < n; f++) {
        const int leaf = plo_bow_word(child_start, child_ids, node_desc, desc + (size_t)f * step);
        const double w = node_weight[leaf];
        if (w > 0) {
            it[m].word = (uint32_t)node_word[leaf];
            it[m].order = f;
            it[m].w = w;
            m++;
        }
    }
    qsort(it, (size_t)m, sizeof(plo_bow_item), plo_cmp_bow); /* std::map order; feature order inside a word */
    int len = 0;
    for (int i = 0; i < m;) {
        int j = i;
        double v = it[i].w; /* insert(id, w) */
        for (j = i + 1; j < m && it[j].word == it[i].word; j++)
            if (weighting == 0 || weighting == 1) v += it[j].w; /* addWeight; addIfNotExist keeps the first */
        ids[len] = it[i].word;
        vals[len] = v;
        len++;
        i = j;
    }
    free(it);
    double norm = 0.0;
    for (int i = 0; i < len; i++) norm += fabs(vals[i]);
    if (norm > 0.0)
        for (int i = 0; i < len; i++) vals[i] /= norm;
    return len;
}

PLO_API double plo_bow_score(const uint32_t *ids1, const double *vals1, int n1, const uint32_t *ids2,
                             const double *vals2, int n2)
{
    double score = 0;
    int i = 0, j = 0;
    while (i < n1 && j < n2) {
        if (ids1[i] == ids2[j]) {
            const double vi = vals1[i], wi = vals2[j];
            score += fabs(vi - wi) - fabs(vi) - fabs(wi);
            i++;
            j++;
        } else if (ids1[i] < ids2[j]) {
            i++; /* lower_bound jump == linear advance on sorted ids */
        } else {
            j++;
        }
    }
    return -score / 2.0;
}

/* ------------------------------------------------------------------------------------------
 * Opt-in geometric filter for matched line pairs (BASELINE config 2 "NNR line matching with overlap/angle
 * filter"; SURVEY 8 note 6).  In this fork the temporal line matcher applies no geometric filter
 * (stereoFrameHandler.cpp:182-207); the two tests below are the reference's own functions, evaluated per
 * matched pair:
 *   plo_line_segment_overlap  StereoFrame::lineSegmentOverlap (stvo-pl/src/stereoFrame.cpp:521-627): the
 *       fraction of the observed segment covered by the other segment projected onto its line, with the
 *       vertical (|dx| < 1) and horizontal (|dy| < 1) special cases; all double, float literals promoted.
 *   direction similarity      |dot(normalize(e1 - s1), normalize(e2 - s2))| with dot / normalize of
 *       stvo-pl/include/matching.h:39-48 and the test of matching.cpp:221 (`abs(dot) < lineSimTh` rejects,
 *       so a NaN similarity passes).
 * Pinned on the reference's own stereoFrame.cpp compiled unmodified (oracle/_ref/libplref_stereo.so,
 * tests/test_ref_stereo.py::test_scalar_gates_vs_reference; golden outputs in tests/golden/stereo.npz).
 */
static double plo_overlap_from_lambdas(double lambda_s, double lambda_e)
{
    const double lambda_min = plo_min(lambda_s, lambda_e);
    const double lambda_max = plo_max(lambda_s, lambda_e);
    if (lambda_min < 0.f && lambda_max > 1.f) return 1.f;
    if (lambda_max < 0.f || lambda_min > 1.f) return 0.f;
    if (lambda_min < 0.f) return lambda_max;
    if (lambda_max > 1.f) return 1.f - lambda_min;
    return lambda_max - lambda_min;
}

PLO_API double plo_line_segment_overlap(const double spl_obs[2], const double epl_obs[2],
                                        const double spl_proj[2], const double epl_proj[2])
{
    const double l0 = epl_obs[0] - spl_obs[0], l1 = epl_obs[1] - spl_obs[1];
    if (fabs(spl_obs[0] - epl_obs[0]) < 1.0) { /* vertical lines, :526-554 */
        const double lambda_s = (spl_proj[1] - spl_obs[1]) / l1;
        const double lambda_e = (epl_proj[1] - spl_obs[1]) / l1;
        return plo_overlap_from_lambdas(lambda_s, lambda_e);
    }
    if (fabs(spl_obs[1] - epl_obs[1]) < 1.0) { /* horizontal lines, :555-584 */
        const double lambda_s = (spl_proj[0] - spl_obs[0]) / l0;
        const double lambda_e = (epl_proj[0] - spl_obs[0]) / l0;
        return plo_overlap_from_lambdas(lambda_s, lambda_e);
    }
    /* non-degenerate, :585-622: foot points of the other segment's endpoints on the observed line */
    const double a = spl_obs[1] - epl_obs[1];
    const double b = epl_obs[0] - spl_obs[0];
    const double c = spl_obs[0] * epl_obs[1] - epl_obs[0] * spl_obs[1];
    const double lxy = 1.f / (a * a + b * b);
    const double sx = (b * (b * spl_proj[0] - a * spl_proj[1]) - a * c) * lxy;
    const double ex = (b * (b * epl_proj[0] - a * epl_proj[1]) - a * c) * lxy;
    const double lambda_s = (sx - spl_obs[0]) / l0;
    const double lambda_e = (ex - spl_obs[0]) / l0;
    return plo_overlap_from_lambdas(lambda_s, lambda_e);
}

/* ln1 / ln2: n x (sx, sy, ex, ey) float32.  keep[i1] = 1 iff m12[i1] in [0, n2), overlap > overlap_th and
 * !(sim < line_sim_th); overlap / sim are written for every matched row (0 elsewhere).  Returns #kept. */
PLO_API int plo_line_pair_filter(const float *ln1, int n1, const float *ln2, int n2, const int32_t *m12,
                                 double overlap_th, double line_sim_th, uint8_t *keep, double *overlap,
                                 double *sim)
{
    int kept = 0;
    for (int i1 = 0; i1 < n1; i1++) {
        keep[i1] = 0;
        overlap[i1] = 0.0;
        sim[i1] = 0.0;
        const int i2 = m12[i1];
        if (i2 < 0 || i2 >= n2) continue;
        const double so[2] = {ln1[4 * i1], ln1[4 * i1 + 1]}, eo[2] = {ln1[4 * i1 + 2], ln1[4 * i1 + 3]};
        const double sp[2] = {ln2[4 * i2], ln2[4 * i2 + 1]}, ep[2] = {ln2[4 * i2 + 2], ln2[4 * i2 + 3]};
        const double ov = plo_line_segment_overlap(so, eo, sp, ep);
        double v[2] = {eo[0] - so[0], eo[1] - so[1]}, w[2] = {ep[0] - sp[0], ep[1] - sp[1]};
        const double mv = sqrt(v[0] * v[0] + v[1] * v[1]), mw = sqrt(w[0] * w[0] + w[1] * w[1]);
        v[0] /= mv; v[1] /= mv;
        w[0] /= mw; w[1] /= mw;
        const double s = fabs(v[0] * w[0] + v[1] * w[1]);
        overlap[i1] = ov;
        sim[i1] = s;
        if (ov > overlap_th && !(s < line_sim_th)) {
            keep[i1] = 1;
            kept++;
        }
    }
    return kept;
}


/* ---------------------------------------------------------------------------------------------------------------
 * Local-map selection and reprojection gates of MapHandler::matchMap2KFPoints / matchMap2KFLines
 * (src/mapHandler.cpp:583-682, :685-803; PinholeStereoCamera::projection stvo-pl/src/pinholeStereoCamera.cpp:239-245).
 * PARITY UNPINNED: src/mapHandler.cpp needs g2o / Eigen / OpenCV to compile, and the summation order of Eigen's fixed
 * 3x3 * 3x1 product (Twf.block(0,0,3,3) * X) is not reproducible without Eigen.  Restated from the source with the
 * product summed left to right; tests/test_reproj.py cross-checks selections, cell coordinates and gate decisions
 * against an independent numpy form and bounds the fp64 difference to it.
 */
typedef struct plo_map_view {
    double T[12];
    double fx, fy, cx, cy;
    double inv_width, inv_height;
    int32_t width, height;
} plo_map_view;

static void plo_view_point(const plo_map_view *v, const double *X, double P[3], double uv[2])
{
    for (int r = 0; r < 3; r++)
        P[r] = ((v->T[4 * r] * X[0] + v->T[4 * r + 1] * X[1]) + v->T[4 * r + 2] * X[2]) + v->T[4 * r + 3];
    uv[0] = v->cx + v->fx * P[0] / P[2];
    uv[1] = v->cy + v->fy * P[1] / P[2];
}

/* mapHandler.cpp:596-609 / :698-714 -> number of selected landmarks */
PLO_API int plo_map_select(int is_lines, const double *X, const uint8_t *active, int n, const plo_map_view *v,
                           int32_t *sel, int32_t *coords, double *pf)
{
    const int per = is_lines ? 2 : 1;
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (active && !active[i]) continue;
        double uv[2][2];
        int keep = 1;
        for (int k = 0; k < per; k++) {
            double P[3];
            plo_view_point(v, X + (size_t)i * 3 * per + 3 * k, P, uv[k]);
            keep = keep && uv[k][0] > 0 && uv[k][0] < v->width && uv[k][1] > 0 && uv[k][1] < v->height && P[2] > 0.0;
        }
        if (!keep) continue;
        sel[m] = i;
        for (int k = 0; k < per; k++) {
            coords[(size_t)m * 2 * per + 2 * k] = (int32_t)(uv[k][0] * v->inv_width);      /* pair<int,int>(double, double) */
            coords[(size_t)m * 2 * per + 2 * k + 1] = (int32_t)(uv[k][1] * v->inv_height);
            pf[(size_t)m * 2 * per + 2 * k] = uv[k][0];
            pf[(size_t)m * 2 * per + 2 * k + 1] = uv[k][1];
        }
        m++;
    }
    return m;
}

/* mapHandler.cpp:652-680 / :767-800 -> count after the rejections */
PLO_API int plo_map_gate(int is_lines, const double *pf, const int32_t *m12, int n_sel, const double *feat, int n2,
                         double max_epip, uint8_t *ok, int count)
{
    for (int i1 = 0; i1 < n_sel; i1++) {
        const int i2 = m12[i1];
        ok[i1] = 0;
        if (i2 < 0 || i2 >= n2) continue;
        int pass;
        if (!is_lines) {
            const double dx = pf[2 * (size_t)i1] - feat[2 * (size_t)i2], dy = pf[2 * (size_t)i1 + 1] - feat[2 * (size_t)i2 + 1];
            pass = sqrt(dx * dx + dy * dy) < max_epip;
        } else {
            const double *l = feat + 3 * (size_t)i2, *p = pf + 4 * (size_t)i1;
            const double e0 = l[0] * p[0] + l[1] * p[1] + l[2];
            const double e1 = l[0] * p[2] + l[1] * p[3] + l[2];
            pass = e0 < max_epip && e1 < max_epip;
        }
        if (pass) ok[i1] = 1;
        else --count;
    }
    return count;
}
