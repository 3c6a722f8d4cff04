// host/replay_main.cpp -- offline replay driver: the ROS-free stand-in for the node's timer
// loop.  Loads maps/<name>.yaml and config/mcl_config.yaml, drives the C++ ParticleFilter along
// a synthetic straight-line trajectory with scans cast by the filter's own calc_range_many, and
// prints the pose estimate and the update time.
//
//   mcl_replay <map.yaml> [mcl_config.yaml] [--particles N] [--steps K] [--x X --y Y --theta T]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "particle_filter.hpp"

using namespace particle_filter_cpp;

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <map.yaml> [mcl_config.yaml] [--particles N] [--steps K] [--x X --y Y --theta T]\n", argv[0]);
        return 2;
    }
    Parameters prm;
    std::string map_yaml = argv[1];
    int steps = 20;
    double x0 = NAN, y0 = NAN, th0 = 0.0;
    for (int i = 2; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--particles") && i + 1 < argc) prm.max_particles = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--steps") && i + 1 < argc) steps = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--x") && i + 1 < argc) x0 = std::atof(argv[++i]);
        else if (!std::strcmp(argv[i], "--y") && i + 1 < argc) y0 = std::atof(argv[++i]);
        else if (!std::strcmp(argv[i], "--theta") && i + 1 < argc) th0 = std::atof(argv[++i]);
        else if (argv[i][0] != '-') {
            std::string err;
            const int keep = prm.max_particles;
            if (!prm.load_yaml(argv[i], &err)) {
                std::fprintf(stderr, "config: %s\n", err.c_str());
                return 2;
            }
            (void)keep;
        }
    }
    OccupancyGrid grid;
    std::string err;
    if (!load_map(map_yaml, grid, &err)) {
        std::fprintf(stderr, "map: %s\n", err.c_str());
        return 2;
    }
    ParticleFilter pf(prm);
    pf.get_omap(grid);
    if (std::isnan(x0)) {   // default start: the free cell closest to the middle of the grid
        double best = 1e300;
        for (uint32_t r = 0; r < grid.height; ++r)
            for (uint32_t c = 0; c < grid.width; ++c)
                if (grid.data[static_cast<size_t>(r) * grid.width + c] == 0) {
                    const double d = std::hypot(double(r) - grid.height / 2.0, double(c) - grid.width / 2.0);
                    if (d < best) {
                        best = d;
                        x0 = (c + 0.5) * grid.resolution + grid.origin_x;
                        y0 = (r + 0.5) * grid.resolution + grid.origin_y;
                    }
                }
    }
    const int nb = 1080;
    const float amin = -2.35f, ainc = 4.7f / 1079.0f;
    Vector3d gt{{x0, y0, th0}};
    pf.initialize_particles_pose(gt);
    std::printf("map %s %ux%u res %.6f MAX_RANGE_PX %d particles %d\n", map_yaml.c_str(), grid.width, grid.height,
                grid.resolution, pf.max_range_px(), prm.max_particles);
    const double v = 1.0, dt = 0.025;
    for (int t = 0; t < steps; ++t) {
        // creep forward only while there is room ahead
        const float ahead = pf.cast_ray(gt[0], gt[1], gt[2]);
        const double vel = ahead > 1.0f ? v : 0.0;
        gt[0] += vel * dt * std::cos(gt[2]);
        gt[1] += vel * dt * std::sin(gt[2]);
        std::vector<double> q(static_cast<size_t>(3) * nb);
        for (int i = 0; i < nb; ++i) {
            q[i] = gt[0];
            q[nb + i] = gt[1];
            q[2 * nb + i] = gt[2] + static_cast<double>(amin + i * ainc);
        }
        const std::vector<float> scan = pf.calc_range_many(q);
        // the node's loop: scan callback, odometry callback, timer tick
        pf.lidarCB(amin, ainc, scan);
        pf.odomCB(Vector3d{{100.0 + gt[0], -50.0 + gt[1], gt[2]}}, vel, 0.0);   // odom frame != map frame
        if (!pf.timer_update(dt)) {
            std::fprintf(stderr, "update skipped\n");
            return 1;
        }
        const Vector3d cur = pf.get_current_pose();   // what publish_tf would send (:839-845)
        if (!pf.is_pose_valid(cur) || std::hypot(cur[0] - gt[0], cur[1] - gt[1]) > 1.0) {
            std::fprintf(stderr, "get_current_pose off: [%.3f %.3f %.3f]\n", cur[0], cur[1], cur[2]);
            return 1;
        }
        const Vector3d p = pf.inferred_pose();
        std::printf("iter %3d  gt [%.3f %.3f %.3f]  est [%.3f %.3f %.3f]  err %.3f m  %.3f ms\n", pf.iterations(), gt[0], gt[1],
                    gt[2], p[0], p[1], p[2], std::hypot(p[0] - gt[0], p[1] - gt[1]), pf.last_update_ms());
        if (!pf.is_pose_valid(p)) return 1;
    }
    const Vector3d p = pf.inferred_pose();
    return std::hypot(p[0] - gt[0], p[1] - gt[1]) < 0.5 ? 0 : 1;
}
