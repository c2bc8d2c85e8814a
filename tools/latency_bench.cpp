// C-level latency of the host-buffer entry points (no Python in the loop).
//   g++ -O2 -std=c++17 tools/latency_bench.cpp -Iinclude -Lpl_inertial_slam_b200/lib -lplmatch -Wl,-rpath,$PWD/pl_inertial_slam_b200/lib -o /tmp/latency_bench
#include <algorithm>
#include <chrono>
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
    return 0;
}
