// oracle/ref_harness.cpp -- C ABI around the UNMODIFIED reference class (TEST INFRASTRUCTURE).
//
// oracle/Makefile compiles /root/reference/src/particle_filter.cpp and src/utils.cpp where
// they lie, against the interface shims in oracle/shim/, together with this file, into
// oracle/_ref/libref_pf.so.  This file is compiled with -fno-access-control so it can call
// the reference's private members (MCL, expected_pose, cast_ray, ...) and seed its RNG; it
// contains no MCL arithmetic of its own.  Used to (1) pin oracle/mcl_oracle.cpp bit-for-bit
// and (2) time the reference's own CPU implementation (bench.py --impl reference).
#include <cstring>
#include <memory>
#include <vector>

#include "particle_filter_cpp/particle_filter.hpp"

using particle_filter_cpp::ParticleFilter;

struct ref_pf {
    std::unique_ptr<ParticleFilter> pf;
};

ref_clock::State& ref_clock::state() {
    static State s;
    return s;
}

extern "C" {

void ref_clear_params() { rclcpp::ShimGlobals::get().overrides.clear(); }
void ref_set_param_int(const char* name, long long v) { rclcpp::ShimGlobals::get().overrides[name] = static_cast<int64_t>(v); }
void ref_set_param_double(const char* name, double v) { rclcpp::ShimGlobals::get().overrides[name] = v; }
void ref_set_param_bool(const char* name, int v) { rclcpp::ShimGlobals::get().overrides[name] = (v != 0); }
void ref_set_verbose(int v) { rclcpp::ShimGlobals::get().verbose = (v != 0); }

// what nav2 map_server would answer on /map_server/map
void ref_install_map(const int8_t* data, int width, int height, float resolution, double ox, double oy,
                     double oyaw) {
    auto& g = rclcpp::ShimGlobals::get();
    g.map.info.resolution = resolution;
    g.map.info.width = static_cast<uint32_t>(width);
    g.map.info.height = static_cast<uint32_t>(height);
    g.map.info.origin.position.x = ox;
    g.map.info.origin.position.y = oy;
    g.map.info.origin.orientation = particle_filter_cpp::utils::geometry::yaw_to_quaternion(oyaw);
    g.map.data.assign(data, data + static_cast<size_t>(width) * height);
    g.have_map = true;
}

// Constructs the node: declares parameters, get_omap(), initialize_global() -- the
// reference's own ctor (src/particle_filter.cpp:19-170).
ref_pf* ref_create(unsigned seed) {
    auto* h = new ref_pf();
    h->pf = std::make_unique<ParticleFilter>();
    h->pf->rng_.seed(seed);
    h->pf->normal_dist_.reset();
    return h;
}
void ref_destroy(ref_pf* h) { delete h; }

void ref_seed(ref_pf* h, unsigned seed) {
    h->pf->rng_.seed(seed);
    h->pf->normal_dist_.reset();
}

int ref_num_particles(ref_pf* h) { return h->pf->MAX_PARTICLES; }
int ref_max_range_px(ref_pf* h) { return h->pf->MAX_RANGE_PX; }
double ref_map_resolution(ref_pf* h) { return h->pf->map_resolution_; }
int ref_num_threads(ref_pf* h) { return h->pf->NUM_THREADS; }

// lidarCB (src/particle_filter.cpp:295-323) with a full scan; returns #downsampled beams
int ref_lidar(ref_pf* h, float angle_min, float angle_increment, const float* ranges, int n) {
    auto msg = std::make_shared<sensor_msgs::msg::LaserScan>();
    msg->angle_min = angle_min;
    msg->angle_increment = angle_increment;
    msg->ranges.assign(ranges, ranges + n);
    h->pf->lidarCB(msg);
    return static_cast<int>(h->pf->downsampled_angles_.size());
}
int ref_get_beam_angles(ref_pf* h, float* out) {
    const auto& a = h->pf->downsampled_angles_;
    std::memcpy(out, a.data(), a.size() * sizeof(float));
    return static_cast<int>(a.size());
}
int ref_get_downsampled_ranges(ref_pf* h, float* out) {
    const auto& a = h->pf->downsampled_ranges_;
    std::memcpy(out, a.data(), a.size() * sizeof(float));
    return static_cast<int>(a.size());
}

void ref_get_sensor_table(ref_pf* h, double* out) {
    const auto& t = h->pf->sensor_model_table_;
    std::memcpy(out, t.data(), sizeof(double) * t.rows() * t.cols());
}

void ref_set_state(ref_pf* h, const double* particles_colmajor, const double* weights) {
    const int N = h->pf->MAX_PARTICLES;
    if (particles_colmajor) std::memcpy(h->pf->particles_.data(), particles_colmajor, sizeof(double) * 3 * N);
    if (weights) std::memcpy(h->pf->weights_.data(), weights, sizeof(double) * N);
}
void ref_get_state(ref_pf* h, double* particles_colmajor, double* weights) {
    const int N = h->pf->MAX_PARTICLES;
    if (particles_colmajor) std::memcpy(particles_colmajor, h->pf->particles_.data(), sizeof(double) * 3 * N);
    if (weights) std::memcpy(weights, h->pf->weights_.data(), sizeof(double) * N);
}

void ref_init_pose(ref_pf* h, const double pose[3]) {
    h->pf->initialize_particles_pose(Eigen::Vector3d(pose[0], pose[1], pose[2]));
}
void ref_init_global(ref_pf* h) { h->pf->initialize_global(); }

float ref_cast_ray(ref_pf* h, double x, double y, double a) { return h->pf->cast_ray(x, y, a); }

// MCL(action, observation) then expected_pose(), exactly as timer_update does (:777-778)
void ref_mcl(ref_pf* h, const double action[3], const float* obs, int n_obs, double pose_out[3]) {
    std::vector<float> o(obs, obs + n_obs);
    h->pf->MCL(Eigen::Vector3d(action[0], action[1], action[2]), o);
    Eigen::Vector3d p = h->pf->expected_pose();
    h->pf->inferred_pose_ = p;
    pose_out[0] = p[0];
    pose_out[1] = p[1];
    pose_out[2] = p[2];
}
void ref_get_ranges(ref_pf* h, float* out) {
    std::memcpy(out, h->pf->ranges_.data(), sizeof(float) * h->pf->ranges_.size());
}
long long ref_num_ranges(ref_pf* h) { return static_cast<long long>(h->pf->ranges_.size()); }

// the reference's own TimingStats buckets (utils.hpp:49-61), ms accumulated
void ref_get_timing(ref_pf* h, double out[6], int* count) {
    const auto& t = h->pf->timing_stats_;
    out[0] = t.total_mcl_time;
    out[1] = t.resampling_time;
    out[2] = t.motion_model_time;
    out[3] = t.query_prep_time;
    out[4] = t.ray_casting_time;
    out[5] = t.sensor_model_time;
    *count = t.measurement_count;
}
void ref_reset_timing(ref_pf* h) { h->pf->timing_stats_.reset(); }

// ---- the node's update shell: the reference's own callbacks, with scripted clocks -------------------
// (timer_update keeps its timer in function-local statics, :735-746: one timeline per process.  Scripted
// steady time must therefore only move forward, in steps of at most 1 s -- a larger step makes the
// reference return before it stores the new time (:750-752), and every later tick would do the same.)
void ref_clock_fake(int on) { ref_clock::state().fake = on != 0; }
void ref_clock_set_steady(double seconds) { ref_clock::state().steady_ns = static_cast<long long>(seconds * 1e9 + 0.5); }
void ref_clock_set_hr_quantum(double ms) { ref_clock::state().hr_quantum_ns = static_cast<long long>(ms * 1e6 + 0.5); }

void ref_odom(ref_pf* h, double x, double y, double yaw, double v, double w) {
    auto msg = std::make_shared<nav_msgs::msg::Odometry>();
    msg->pose.pose.position.x = x;
    msg->pose.pose.position.y = y;
    msg->pose.pose.orientation = particle_filter_cpp::utils::geometry::yaw_to_quaternion(yaw);
    msg->twist.twist.linear.x = v;
    msg->twist.twist.angular.z = w;
    msg->header.stamp.sec = 1;
    h->pf->odomCB(msg);
}
void ref_clicked_pose(ref_pf* h, double x, double y, double yaw) {
    auto msg = std::make_shared<geometry_msgs::msg::PoseWithCovarianceStamped>();
    msg->pose.pose.position.x = x;
    msg->pose.pose.position.y = y;
    msg->pose.pose.orientation = particle_filter_cpp::utils::geometry::yaw_to_quaternion(yaw);
    h->pf->clicked_pose(msg);
}
void ref_timer_update(ref_pf* h) { h->pf->timer_update(); }
void ref_current_pose(ref_pf* h, double out[3]) {
    const Eigen::Vector3d p = h->pf->get_current_pose();
    for (int k = 0; k < 3; ++k) out[k] = p[k];
}
void ref_set_inferred(ref_pf* h, const double pose[3]) { h->pf->inferred_pose_ = Eigen::Vector3d(pose[0], pose[1], pose[2]); }
// same layout as pfhost_shell_state (monte_carlo_localization_b200/host/host_capi.cpp)
void ref_shell_state(ref_pf* h, double out[23]) {
    const auto& p = *h->pf;
    for (int k = 0; k < 3; ++k) {
        out[k] = p.inferred_pose_[k];
        out[3 + k] = p.odom_pose_[k];
        out[6 + k] = p.odom_reference_pose_[k];
        out[9 + k] = p.odom_reference_odom_[k];
        out[12 + k] = p.last_pose_[k];
    }
    out[15] = p.iters_;
    out[16] = p.odom_initialized_;
    out[17] = p.pose_initialized_from_rviz_;
    out[18] = p.odom_tracking_active_;
    out[19] = p.timing_stats_.total_mcl_time;
    out[20] = p.timing_stats_.measurement_count;
    out[21] = p.current_velocity_;
    out[22] = p.current_angular_vel_;
}
// visualize()'s weighted sub-sample (:946-958) with the reference's own generator: k draws of
// discrete_distribution(weights_) -> indices
void ref_viz_sample(ref_pf* h, int k, int* idx_out) {
    std::discrete_distribution<int> dist(h->pf->weights_.begin(), h->pf->weights_.end());
    for (int i = 0; i < k; ++i) idx_out[i] = dist(h->pf->rng_);
}

}  // extern "C"
