// C-level latency of the host-buffer entry points (no Python in the loop).
//   g++ -O2 -std=c++17 tools/latency_bench.cpp -Iinclude -Lpl_inertial_slam_b200/lib -lplmatch -Wl,-rpath,$PWD/pl_inertial_slam_b200/lib -o /tmp/latency_bench
#include <algorithm>
#include <dlfcn.h>
#include <atomic>
#include <chrono>
#include <cmath>
#include <thread>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

#include "plmatch.h"

using clk = std::chrono::steady_clock;

template <class F> double median_us(F &&f, int reps) {
    for (int i = 0; i < 10; ++i) f();
    std::vector<double> t;
    for (int i = 0; i < reps; ++i) {
        auto a = clk::now();
        f();
        t.push_back(std::chrono::duration<double, std::micro>(clk::now() - a).count());
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

// One stereo frame the way the reference drives the matcher: StereoFrame::extractStereoFeatures runs points and
// lines on two threads (stereoFrame.cpp:75-76), then StereoFrameHandler::f2fTracking does the same for the
// temporal matches (stereoFrameHandler.cpp:142-143).  Two persistent host threads (each with its own per-thread
// plm context = its own CUDA stream), two stages per frame with a join after each: wall time per frame.
struct Feat {
    int n;
    std::vector<uint8_t> d1, d2;
    std::vector<int32_t> coords, cell_start, cell_items, m12;
    std::vector<double> dirs2;
};

static Feat make_feat(std::mt19937_64 &rng, int n, bool lines) {
    const int rows = 48, cols = 64;
    Feat f;
    f.n = n;
    f.d1.resize(size_t(n) * 32);
    f.d2.resize(size_t(n) * 32);
    for (auto &b : f.d1) b = uint8_t(rng());
    for (size_t i = 0; i < f.d2.size(); ++i) f.d2[i] = f.d1[i] ^ ((rng() % 12 == 0) ? uint8_t(1u << (rng() % 8)) : 0);
    f.coords.resize(size_t(n) * (lines ? 4 : 2));
    f.cell_start.assign(rows * cols + 1, 0);
    f.cell_items.resize(n);
    f.m12.resize(n);
    std::vector<int> cx(n), cy(n);
    for (int i = 0; i < n; ++i) {
        cx[i] = 10 + int(rng() % (cols - 10));
        cy[i] = int(rng() % rows);
        const int qx = std::min(cols - 1, cx[i] + int(rng() % 8)), qy = cy[i]; // left feature: right one shifted by a disparity
        if (lines) {
            const int dx = int(rng() % 7) - 3, dy = int(rng() % 7) - 3;
            f.coords[4 * i] = qx; f.coords[4 * i + 1] = qy; f.coords[4 * i + 2] = qx + dx; f.coords[4 * i + 3] = qy + dy;
            const double nrm = std::sqrt(double(dx * dx + dy * dy));
            f.dirs2.push_back(nrm > 0 ? dx / nrm : 1.0);
            f.dirs2.push_back(nrm > 0 ? dy / nrm : 0.0);
        } else {
            f.coords[2 * i] = qx; f.coords[2 * i + 1] = qy;
        }
        f.cell_start[cx[i] * rows + cy[i] + 1]++;
    }
    for (int c = 0; c < rows * cols; ++c) f.cell_start[c + 1] += f.cell_start[c];
    std::vector<int32_t> cur(f.cell_start.begin(), f.cell_start.end() - 1);
    for (int i = 0; i < n; ++i) f.cell_items[cur[cx[i] * rows + cy[i]]++] = i;
    return f;
}

static void frame_mode(int reps) {
    std::mt19937_64 rng(7);
    Feat P = make_feat(rng, 600, false), Ln = make_feat(rng, 200, true);
    const int32_t win_st[4] = {10, 0, 0, 0};
    auto stereo = [&](Feat &f, bool lines) {
        int cnt = 0;
        std::fill(f.m12.begin(), f.m12.end(), -1);
        if (lines)
            plm_match_grid_lines(nullptr, f.coords.data(), f.d1.data(), f.n, 32, f.cell_start.data(), f.cell_items.data(), 48, 64,
                                 f.d2.data(), f.n, 32, f.dirs2.data(), 0.75, win_st, 0.9, 1, f.m12.data(), &cnt);
        else
            plm_match_grid_points(nullptr, f.coords.data(), f.d1.data(), f.n, 32, f.cell_start.data(), f.cell_items.data(), 48, 64,
                                  f.d2.data(), f.n, 32, win_st, 0.9, 1, f.m12.data(), &cnt);
    };
    auto temporal = [&](Feat &f) {
        int cnt = 0;
        std::fill(f.m12.begin(), f.m12.end(), -1);
        plm_match(nullptr, f.d1.data(), f.n, 32, f.d2.data(), f.n, 32, 0.9f, 1, f.m12.data(), &cnt);
    };
    // serial: the four calls one after the other on one thread
    const double serial = median_us([&] { stereo(P, false); stereo(Ln, true); temporal(P); temporal(Ln); }, reps);
    // frame session: the same four calls recorded and executed as ONE round trip (plm_frame_begin / plm_frame_end)
    int c4[4];
    auto session_frame = [&] {
        std::fill(P.m12.begin(), P.m12.end(), -1);
        std::fill(Ln.m12.begin(), Ln.m12.end(), -1);
        std::vector<int32_t> &tp12 = P.m12, &tl12 = Ln.m12;
        static std::vector<int32_t> t_p, t_l;
        t_p.assign(P.n, -1);
        t_l.assign(Ln.n, -1);
        plm_frame_begin(nullptr);
        plm_match_grid_points(nullptr, P.coords.data(), P.d1.data(), P.n, 32, P.cell_start.data(), P.cell_items.data(), 48, 64, P.d2.data(), P.n, 32,
                              win_st, 0.9, 1, tp12.data(), &c4[0]);
        plm_match_grid_lines(nullptr, Ln.coords.data(), Ln.d1.data(), Ln.n, 32, Ln.cell_start.data(), Ln.cell_items.data(), 48, 64, Ln.d2.data(), Ln.n,
                             32, Ln.dirs2.data(), 0.75, win_st, 0.9, 1, tl12.data(), &c4[1]);
        plm_match(nullptr, P.d1.data(), P.n, 32, P.d2.data(), P.n, 32, 0.9f, 1, t_p.data(), &c4[2]);
        plm_match(nullptr, Ln.d1.data(), Ln.n, 32, Ln.d2.data(), Ln.n, 32, 0.9f, 1, t_l.data(), &c4[3]);
        plm_frame_end(nullptr);
    };
    plm_set_option("frame_fused", 0);
    const double session_lanes = median_us(session_frame, reps);
    plm_set_option("frame_fused", 1);
    const double session = median_us(session_frame, reps);
    // the same session through a window of +-3 cells (the keyframe matchers' window): ~6x the candidates per row
    const int32_t win_kf[4] = {3, 3, 3, 3};
    auto session_kf = [&] {
        static std::vector<int32_t> a, b, t_p, t_l;
        a.assign(P.n, -1); b.assign(Ln.n, -1); t_p.assign(P.n, -1); t_l.assign(Ln.n, -1);
        int c[4];
        plm_frame_begin(nullptr);
        plm_match_grid_points(nullptr, P.coords.data(), P.d1.data(), P.n, 32, P.cell_start.data(), P.cell_items.data(), 48, 64, P.d2.data(), P.n, 32,
                              win_kf, 0.9, 1, a.data(), &c[0]);
        plm_match_grid_lines(nullptr, Ln.coords.data(), Ln.d1.data(), Ln.n, 32, Ln.cell_start.data(), Ln.cell_items.data(), 48, 64, Ln.d2.data(), Ln.n,
                             32, Ln.dirs2.data(), 0.75, win_kf, 0.9, 1, b.data(), &c[1]);
        plm_match(nullptr, P.d1.data(), P.n, 32, P.d2.data(), P.n, 32, 0.9f, 1, t_p.data(), &c[2]);
        plm_match(nullptr, Ln.d1.data(), Ln.n, 32, Ln.d2.data(), Ln.n, 32, 0.9f, 1, t_l.data(), &c[3]);
        plm_frame_end(nullptr);
    };
    const double session_w3 = median_us(session_kf, reps);
    if (std::getenv("PLM_LATENCY_TIMELINE")) { // debug builds (PLM_BUILD_DEFINES=-DPLM_TIMELINE): phase stamps of the frame kernel
        typedef int (*tl_fn)(long long *);
        tl_fn tl = reinterpret_cast<tl_fn>(dlsym(RTLD_DEFAULT, "plm_debug_timeline"));
        if (!tl) {
            printf("  no plm_debug_timeline in this build\n");
        } else {
            std::vector<long long> buf(128 * 24);
            std::vector<double> acc(128 * 24, 0.0);
            const int runs = 200;
            for (int r = 0; r < runs + 20; ++r) {
                session_frame();
                if (r < 20) continue;
                tl(buf.data());
                for (int c = 0; c < 128; ++c)
                    for (int k = 0; k < 24; ++k) acc[c * 24 + k] += double(buf[c * 24 + k] - buf[c * 24]) / runs;
            }
            printf("  phase stamps, SM cycles since kernel entry (mean of %d frames); columns = stamps 1..15\n", runs);
            for (int c = 0; c < 96; ++c) {
                if (c >= 16 && c % 8) continue;
                printf("  cta %3d:", c);
                for (int k = 1; k < 16; ++k) printf(" %7.0f", acc[c * 24 + k]);
                printf("\n");
            }
        }
    }
    if (std::getenv("PLM_LATENCY_PARTS")) { // each call of the frame as a session of its own (kernel time per job under ncu)
        static std::vector<int32_t> a, b, t_p, t_l;
        int c;
        auto one = [&](int which) {
            a.assign(P.n, -1); b.assign(Ln.n, -1); t_p.assign(P.n, -1); t_l.assign(Ln.n, -1);
            plm_frame_begin(nullptr);
            if (which == 0)
                plm_match_grid_points(nullptr, P.coords.data(), P.d1.data(), P.n, 32, P.cell_start.data(), P.cell_items.data(), 48, 64, P.d2.data(), P.n, 32,
                                      win_st, 0.9, 1, a.data(), &c);
            if (which == 1)
                plm_match_grid_lines(nullptr, Ln.coords.data(), Ln.d1.data(), Ln.n, 32, Ln.cell_start.data(), Ln.cell_items.data(), 48, 64, Ln.d2.data(),
                                     Ln.n, 32, Ln.dirs2.data(), 0.75, win_st, 0.9, 1, b.data(), &c);
            if (which == 2) plm_match(nullptr, P.d1.data(), P.n, 32, P.d2.data(), P.n, 32, 0.9f, 1, t_p.data(), &c);
            if (which == 3) plm_match(nullptr, Ln.d1.data(), Ln.n, 32, Ln.d2.data(), Ln.n, 32, 0.9f, 1, t_l.data(), &c);
            if (which == 4)
                plm_match_grid_points(nullptr, P.coords.data(), P.d1.data(), P.n, 32, P.cell_start.data(), P.cell_items.data(), 48, 64, P.d2.data(), P.n, 32,
                                      win_kf, 0.9, 1, a.data(), &c);
            plm_frame_end(nullptr);
        };
        const char *names[5] = {"grid points (10,0,0,0)", "grid lines", "match points", "match lines", "grid points +-3"};
        for (int w = 0; w < 5; ++w) printf("  session of one call, %-24s %7.1f us\n", names[w], median_us([&] { one(w); }, reps));
    }
    // two threads, as the reference: a generation counter starts a stage, a done counter joins it
    std::atomic<int> go{0}, done{0};
    std::atomic<bool> quit{false};
    auto worker = [&](Feat *f, bool lines) {
        int seen = 0;
        while (true) {
            while (go.load(std::memory_order_acquire) == seen)
                if (quit.load()) return;
            ++seen;
            if (seen & 1) stereo(*f, lines);
            else temporal(*f);
            done.fetch_add(1, std::memory_order_release);
        }
    };
    std::thread tp(worker, &P, false), tl(worker, &Ln, true);
    auto frame = [&] {
        for (int stage = 0; stage < 2; ++stage) {
            const int target = done.load() + 2;
            go.fetch_add(1, std::memory_order_release);
            while (done.load(std::memory_order_acquire) < target) {
            }
        }
    };
    const double threaded = median_us(frame, reps);
    quit.store(true);
    tp.join();
    tl.join();
    printf("frame (600 pts + 200 lines): stereo matchGrid + temporal match, 4 calls serial %7.1f us; points || lines on two host threads "
           "(the reference's std::async structure) %7.1f us; one frame session (plm_frame_begin/end): one lane per call %7.1f us, ONE launch "
           "(frame_fused_kernel) %7.1f us, one launch with +-3 windows %7.1f us [matches %d %d %d %d]\n", serial, threaded, session_lanes, session,
           session_w3, c4[0], c4[1], c4[2], c4[3]);
    printf("JSON {\"frame_us\": {\"four_calls_serial\": %.1f, \"two_host_threads\": %.1f, \"session_one_lane_per_call\": %.1f, "
           "\"session_one_launch\": %.1f, \"session_one_launch_pm3_windows\": %.1f}, \"features\": \"600 points + 200 lines per side\", "
           "\"reps\": %d}\n", serial, threaded, session_lanes, session, session_w3, reps);
}

int main(int argc, char **argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 200;
    std::mt19937_64 rng(20261018);
    const int rows = 48, cols = 64;
    for (int n : {200, 600, 1200}) {
        std::vector<uint8_t> d1(size_t(n) * 32), d2(size_t(n) * 32);
        for (auto &b : d1) b = uint8_t(rng());
        for (size_t i = 0; i < d2.size(); ++i) d2[i] = d1[i] ^ ((rng() % 12 == 0) ? uint8_t(1u << (rng() % 8)) : 0);
        // train features scattered uniformly over the grid, queries displaced by <= 3 cells
        std::vector<int32_t> cx(n), cy(n), xy(size_t(n) * 2), cell_start(rows * cols + 1, 0), cell_items(n);
        for (int i = 0; i < n; ++i) {
            cx[i] = int(rng() % cols);
            cy[i] = int(rng() % rows);
            xy[2 * i] = cx[i] + int(rng() % 5) - 2;
            xy[2 * i + 1] = cy[i] + int(rng() % 5) - 2;
            cell_start[cx[i] * rows + cy[i] + 1]++;
        }
        for (int c = 0; c < rows * cols; ++c) cell_start[c + 1] += cell_start[c];
        std::vector<int32_t> cur(cell_start.begin(), cell_start.end() - 1);
        for (int i = 0; i < n; ++i) cell_items[cur[cx[i] * rows + cy[i]]++] = i;
        std::vector<int32_t> m12(n);
        int cnt = 0;
        const int32_t win_kf[4] = {3, 3, 3, 3}, win_st[4] = {10, 0, 0, 0};
        auto grid_kf = [&] {
            std::fill(m12.begin(), m12.end(), -1);
            plm_match_grid_points(nullptr, xy.data(), d1.data(), n, 32, cell_start.data(), cell_items.data(), rows, cols, d2.data(), n, 32, win_kf, 0.9, 1, m12.data(), &cnt);
        };
        auto grid_st = [&] {
            std::fill(m12.begin(), m12.end(), -1);
            plm_match_grid_points(nullptr, xy.data(), d1.data(), n, 32, cell_start.data(), cell_items.data(), rows, cols, d2.data(), n, 32, win_st, 0.9, 1, m12.data(), &cnt);
        };
        auto match = [&] {
            std::fill(m12.begin(), m12.end(), -1);
            plm_match(nullptr, d1.data(), n, 32, d2.data(), n, 32, 0.9f, 1, m12.data(), &cnt);
        };
        auto nnr = [&] {
            std::fill(m12.begin(), m12.end(), -1);
            plm_match_nnr(nullptr, d1.data(), n, 32, d2.data(), n, 32, 0.9f, m12.data(), &cnt);
        };
        const double a = median_us(grid_kf, reps), a_cnt = cnt;
        const double b = median_us(grid_st, reps);
        const double c = median_us(match, reps), c_cnt = cnt;
        const double d = median_us(nnr, reps);
        printf("n=%4d  matchGrid(+-3) %7.1f us (%d matches)  matchGrid(10,0,0,0) %7.1f us  match %7.1f us (%d)  matchNNR %7.1f us\n", n, a,
               int(a_cnt), b, c, int(c_cnt), d);
    }
    frame_mode(reps);
    return 0;
}
