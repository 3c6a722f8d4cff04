// host/host_capi.cpp -- C hooks over the C++ host pieces that need no GPU (map loading,
// config parsing) so the CPU test-suite can cross-check them against the Python twins.
#include <cstring>
#include <string>

#include "map_loader.hpp"
#include "particle_filter.hpp"

using namespace particle_filter_cpp;

extern "C" {

// Loads yaml + image.  First call with data == nullptr to get the size.
int pfhost_load_map(const char* yaml_path, int8_t* data, int capacity, int* width, int* height, float* resolution,
                    double origin[3], char* err, int err_cap) {
    OccupancyGrid g;
    std::string e;
    if (!load_map(yaml_path, g, &e)) {
        if (err && err_cap > 0) {
            std::strncpy(err, e.c_str(), err_cap - 1);
            err[err_cap - 1] = 0;
        }
        return -1;
    }
    *width = static_cast<int>(g.width);
    *height = static_cast<int>(g.height);
    *resolution = g.resolution;
    origin[0] = g.origin_x;
    origin[1] = g.origin_y;
    origin[2] = g.origin_yaw;
    if (data) {
        if (capacity < static_cast<int>(g.data.size())) return -2;
        std::memcpy(data, g.data.data(), g.data.size());
    }
    return 0;
}

// Parses config/mcl_config.yaml into the reference's parameter set; out[] order:
// max_particles, angle_step, max_viz_particles, squash_factor, max_range, z_short, z_max, z_rand,
// z_hit, sigma_hit, disp_x, disp_y, disp_theta, lidar_offset_x, timer_frequency, num_threads,
// delay_compensation_factor
int pfhost_load_config(const char* yaml_path, double out[17]) {
    Parameters p;
    if (!p.load_yaml(yaml_path)) return -1;
    const double v[17] = {double(p.max_particles), double(p.angle_step), double(p.max_viz_particles), p.squash_factor,
                          p.max_range, p.z_short, p.z_max, p.z_rand, p.z_hit, p.sigma_hit, p.motion_dispersion_x,
                          p.motion_dispersion_y, p.motion_dispersion_theta, p.lidar_offset_x, p.timer_frequency,
                          double(p.num_threads), p.delay_compensation_factor};
    std::memcpy(out, v, sizeof v);
    return 0;
}

}  // extern "C"
