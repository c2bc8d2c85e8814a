// Stand-in for stvo-pl/src/config.cpp (which needs yaml-cpp + boost, absent here): the Config
// singleton with the constructor defaults the matching path reads (config.cpp:49-51,60,63-69,
// 90-92).  loadFromFile is not provided; tests set fields through the static reference getters.
// Test infrastructure only.
#include "config.h"

#include <cstring>

Config::Config() {
    std::memset(static_cast<void *>(this), 0, sizeof(Config));
    has_points = true;
    has_lines = true;
    lr_in_parallel = true;
    pl_in_parallel = true;
    best_lr_matches = true;
    max_dist_epip = 1.0;
    min_disp = 1.0;
    min_ratio_12_p = 0.9;
    line_sim_th = 0.75;
    stereo_overlap_th = 0.75;
    f2f_overlap_th = 0.75;
    min_line_length = 0.025;
    line_horiz_th = 0.1;
    min_ratio_12_l = 0.9;
    ls_min_disp_ratio = 0.7;
    matching_strategy = 0;
    matching_s_ws = 10;
    matching_f2f_ws = 3;
    orb_scale_factor = 1.2;
    lsd_scale = 1.2;
}

Config::~Config() {}

Config &Config::getInstance() {
    static Config instance;
    return instance;
}
