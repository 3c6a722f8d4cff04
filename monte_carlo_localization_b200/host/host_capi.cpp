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

// ---- UpdateShell driven from outside (CPU parity test against the reference's own timer_update) ----
// The MCL result of a tick (pose, elapsed ms) and the start-up jitter's normal draws are supplied by the
// caller; the shell's decisions (action, tracking state, compensated pose, current pose) are read back.
void* pfhost_shell_create(double delay_compensation_factor, double max_pose_range) {
    auto* s = new UpdateShell();
    s->delay_compensation_factor = delay_compensation_factor;
    s->max_pose_range = max_pose_range;
    return s;
}
void pfhost_shell_destroy(void* h) { delete static_cast<UpdateShell*>(h); }
void pfhost_shell_odom(void* h, const double pose[3], double v, double w, int map_initialized) {
    static_cast<UpdateShell*>(h)->odomCB({{pose[0], pose[1], pose[2]}}, v, w, map_initialized != 0);
}
void pfhost_shell_clicked_pose(void* h, const double pose[3]) {
    static_cast<UpdateShell*>(h)->clicked_pose({{pose[0], pose[1], pose[2]}}, [](const Vector3d&) {});
}
// mcl_pose / mcl_ms: what MCL + expected_pose return for the action the shell synthesises (the caller runs the
// update itself, after pfhost_shell_peek_action, or supplies the reference's result); normals: 3 draws
int pfhost_shell_timer_update(void* h, double dt, int map_initialized, int lidar_initialized, int num_ranges,
                              const double mcl_pose[3], double mcl_ms, int mcl_ok, const double normals[3],
                              double action_out[3]) {
    auto* s = static_cast<UpdateShell*>(h);
    std::vector<float> ranges(static_cast<size_t>(num_ranges > 0 ? num_ranges : 0), 1.0f);
    int ndraw = 0;
    const bool ran = s->timer_update(
        dt, map_initialized != 0, lidar_initialized != 0, ranges,
        [&](const Vector3d& action, const std::vector<float>&) {
            UpdateShell::MclResult r;
            for (int k = 0; k < 3; ++k) action_out[k] = action[k];
            r.pose = {{mcl_pose[0], mcl_pose[1], mcl_pose[2]}};
            r.elapsed_ms = mcl_ms;
            r.ok = mcl_ok != 0;
            return r;
        },
        [&]() { return normals[ndraw++ % 3]; });
    return ran ? 1 : 0;
}
// out[0..2] inferred, [3..5] odom_pose, [6..8] odom_reference_pose, [9..11] odom_reference_odom, [12..14] last_pose,
// [15] iters, [16] odom_initialized, [17] pose_initialized_from_rviz, [18] odom_tracking_active,
// [19] window total ms, [20] window count, [21] velocity, [22] angular velocity
void pfhost_shell_state(void* h, double out[23]) {
    auto* s = static_cast<UpdateShell*>(h);
    for (int k = 0; k < 3; ++k) {
        out[k] = s->inferred_pose_[k];
        out[3 + k] = s->odom_pose_[k];
        out[6 + k] = s->odom_reference_pose_[k];
        out[9 + k] = s->odom_reference_odom_[k];
        out[12 + k] = s->last_pose_[k];
    }
    out[15] = s->iters_;
    out[16] = s->odom_initialized_;
    out[17] = s->pose_initialized_from_rviz_;
    out[18] = s->odom_tracking_active_;
    out[19] = s->window_total_ms_;
    out[20] = s->window_count_;
    out[21] = s->current_velocity_;
    out[22] = s->current_angular_vel_;
}
void pfhost_shell_set_inferred(void* h, const double pose[3]) {
    static_cast<UpdateShell*>(h)->inferred_pose_ = {{pose[0], pose[1], pose[2]}};
}
void pfhost_shell_current_pose(void* h, int map_initialized, const double particle_mean[3], int have_mean, double out[3]) {
    const Vector3d c = static_cast<UpdateShell*>(h)->get_current_pose(map_initialized != 0, [&](Vector3d* m) {
        if (!have_mean) return false;
        *m = {{particle_mean[0], particle_mean[1], particle_mean[2]}};
        return true;
    });
    for (int k = 0; k < 3; ++k) out[k] = c[k];
}

}  // extern "C"
