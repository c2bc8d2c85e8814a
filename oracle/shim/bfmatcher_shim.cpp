// cv::BFMatcher::knnMatch stand-in (NORM_HAMMING, 32-byte rows, no mask, no cross-check).
//
// Follows OpenCV 3.3 modules/core/src/stat.cpp  batchDistance / BatchDistInvoker, K > 0 branch
// (the library is NOT vendored under /root/reference; the reference links libopencv_features2d
// 3.3): per query row, K slots start at (INT_MAX, -1); train row j with distance d is inserted
// iff d < dist[K-1], shifting slots up while dist[k] > d.  BFMatcher::knnMatchImpl then emits,
// per row, one DMatch for every slot whose index is >= 0, distance converted to float.
// Query rows are split over worker threads the way OpenCV's parallel_for_ does.
// Test infrastructure only (it is also the hot loop of the CPU "reference" arm in bench.py).
#include <opencv2/features2d.hpp>

#include <climits>
#include <cstring>
#include <thread>

static int g_threads = 1;
extern "C" __attribute__((visibility("default"))) void plref_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
extern "C" __attribute__((visibility("default"))) int plref_get_threads() { return g_threads; }

namespace {

inline int hamming32(const unsigned char *a, const unsigned char *b) {
    uint64_t x[4], y[4];
    std::memcpy(x, a, 32);
    std::memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
           __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
}

inline int hamming_any(const unsigned char *a, const unsigned char *b, int len) {
    int d = 0;
    for (int i = 0; i < len; i++) d += __builtin_popcount(static_cast<unsigned>(a[i] ^ b[i]));
    return d;
}

void knn_rows(const cv::Mat &q, const cv::Mat &t, std::vector<std::vector<cv::DMatch>> &out, int K,
              int r0, int r1) {
    std::vector<int> dist(K), nidx(K);
    for (int i = r0; i < r1; i++) {
        for (int k = 0; k < K; k++) { dist[k] = INT_MAX; nidx[k] = -1; }
        const unsigned char *qp = q.data + static_cast<size_t>(i) * q.step;
        for (int j = 0; j < t.rows; j++) {
            const unsigned char *tp = t.data + static_cast<size_t>(j) * t.step;
            const int d = (q.cols == 32) ? hamming32(qp, tp) : hamming_any(qp, tp, q.cols);
            if (d < dist[K - 1]) {
                int k;
                for (k = K - 2; k >= 0 && dist[k] > d; k--) {
                    nidx[k + 1] = nidx[k];
                    dist[k + 1] = dist[k];
                }
                nidx[k + 1] = j;
                dist[k + 1] = d;
            }
        }
        std::vector<cv::DMatch> &row = out[i];
        row.clear();
        row.reserve(K);
        for (int k = 0; k < K; k++)
            if (nidx[k] >= 0) row.push_back(cv::DMatch(i, nidx[k], static_cast<float>(dist[k])));
    }
}

} // namespace

namespace cv {

Ptr<BFMatcher> BFMatcher::create(int normType, bool crossCheck) {
    Ptr<BFMatcher> p = std::make_shared<BFMatcher>();
    p->normType_ = normType;
    p->crossCheck_ = crossCheck;
    return p;
}

void BFMatcher::knnMatch(const Mat &q, const Mat &t, std::vector<std::vector<DMatch>> &matches,
                         int K) const {
    matches.clear();
    // DescriptorMatcher::knnMatch: empty query or empty train collection -> no rows at all.
    if (q.empty() || t.empty()) return;
    matches.resize(q.rows);
    const int nt = (g_threads > 1 && static_cast<long>(q.rows) * t.rows > 20000) ? g_threads : 1;
    if (nt == 1) {
        knn_rows(q, t, matches, K, 0, q.rows);
        return;
    }
    std::vector<std::thread> pool;
    const int chunk = (q.rows + nt - 1) / nt;
    for (int w = 0; w < nt; w++) {
        const int r0 = w * chunk, r1 = std::min(q.rows, r0 + chunk);
        if (r0 >= r1) break;
        pool.emplace_back(knn_rows, std::cref(q), std::cref(t), std::ref(matches), K, r0, r1);
    }
    for (auto &th : pool) th.join();
}

} // namespace cv
