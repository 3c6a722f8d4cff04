// host/particle_filter.cpp -- see particle_filter.hpp.  Thin: every numerical step of the
// path is one C-ABI call into libmcl_b200.so.
#include "particle_filter.hpp"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <stdexcept>

#include "../../include/mcl_b200.h"

namespace particle_filter_cpp {

namespace {
void log_error(const char* what, int rc) {
    std::fprintf(stderr, "[particle_filter] %s: %s (%s)\n", what, mcl_status_str(rc), mcl_last_error());
}
std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
    while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
    return s.substr(a, b - a);
}
bool as_bool(const std::string& v) { return v == "true" || v == "True" || v == "1"; }
}  // namespace

bool Parameters::load_yaml(const std::string& path, std::string* err) {
    std::ifstream f(path);
    if (!f) {
        if (err) *err = "cannot open " + path;
        return false;
    }
    std::string line;
    bool in_pf = false, in_params = false;
    while (std::getline(f, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        if (trim(line).empty()) continue;
        const size_t indent = line.find_first_not_of(' ');
        const std::string s = trim(line);
        const size_t colon = s.find(':');
        if (colon == std::string::npos) continue;
        const std::string key = trim(s.substr(0, colon));
        std::string val = trim(s.substr(colon + 1));
        if (val.size() >= 2 && (val.front() == '"' || val.front() == '\'')) val = val.substr(1, val.size() - 2);
        if (indent == 0) {
            in_pf = (key == "particle_filter");
            in_params = false;
            continue;
        }
        if (!in_pf) continue;
        if (key == "ros__parameters") {
            in_params = true;
            continue;
        }
        if (!in_params || val.empty()) continue;
        if (key == "angle_step") angle_step = std::atoi(val.c_str());
        else if (key == "max_particles") max_particles = std::atoi(val.c_str());
        else if (key == "max_viz_particles") max_viz_particles = std::atoi(val.c_str());
        else if (key == "squash_factor") squash_factor = std::atof(val.c_str());
        else if (key == "max_range") max_range = std::atof(val.c_str());
        else if (key == "publish_odom") publish_odom = as_bool(val);
        else if (key == "viz") viz = as_bool(val);
        else if (key == "z_short") z_short = std::atof(val.c_str());
        else if (key == "z_max") z_max = std::atof(val.c_str());
        else if (key == "z_rand") z_rand = std::atof(val.c_str());
        else if (key == "z_hit") z_hit = std::atof(val.c_str());
        else if (key == "sigma_hit") sigma_hit = std::atof(val.c_str());
        else if (key == "motion_dispersion_x") motion_dispersion_x = std::atof(val.c_str());
        else if (key == "motion_dispersion_y") motion_dispersion_y = std::atof(val.c_str());
        else if (key == "motion_dispersion_theta") motion_dispersion_theta = std::atof(val.c_str());
        else if (key == "lidar_offset_x") lidar_offset_x = std::atof(val.c_str());
        else if (key == "lidar_offset_y") lidar_offset_y = std::atof(val.c_str());
        else if (key == "wheelbase") wheelbase = std::atof(val.c_str());
        else if (key == "scan_topic") scan_topic = val;
        else if (key == "odom_topic") odom_topic = val;
        else if (key == "timer_frequency") timer_frequency = std::atof(val.c_str());
        else if (key == "use_parallel_raycasting") use_parallel_raycasting = as_bool(val);
        else if (key == "num_threads") num_threads = std::atoi(val.c_str());
        else if (key == "max_pose_range") max_pose_range = std::atof(val.c_str());
        else if (key == "delay_compensation_factor") delay_compensation_factor = std::atof(val.c_str());
        // sim_mode, range_method, theta_discretization, rangelib_variant, fine_timing, *_frame:
        // present in the YAML, never declared or read by the reference -> ignored
    }
    return true;
}

ParticleFilter::ParticleFilter(const Parameters& params) : p_(params) {
    shell_.delay_compensation_factor = p_.delay_compensation_factor;
    shell_.max_pose_range = p_.max_pose_range;
    mcl_params mp;
    mcl_default_params(&mp);
    mp.max_particles = p_.max_particles;
    mp.max_viz_particles = p_.max_viz_particles;
    mp.angle_step = p_.angle_step;
    mp.squash_factor = p_.squash_factor;
    mp.max_range = p_.max_range;
    mp.z_short = p_.z_short;
    mp.z_max = p_.z_max;
    mp.z_rand = p_.z_rand;
    mp.z_hit = p_.z_hit;
    mp.sigma_hit = p_.sigma_hit;
    mp.motion_dispersion_x = p_.motion_dispersion_x;
    mp.motion_dispersion_y = p_.motion_dispersion_y;
    mp.motion_dispersion_theta = p_.motion_dispersion_theta;
    mp.seed = p_.seed;
    mp.num_filters = 1;
    const int rc = mcl_create(&mp, p_.device, &ctx_);
    if (rc != MCL_OK) {
        // no CPU fallback: a filter without its device context cannot run at all
        throw std::runtime_error(std::string("ParticleFilter: mcl_create failed: ") + mcl_status_str(rc) + " (" +
                                 mcl_last_error() + ")");
    }
}

ParticleFilter::~ParticleFilter() { mcl_destroy(ctx_); }

void ParticleFilter::get_omap(const OccupancyGrid& map) {
    const int rc = mcl_set_map(ctx_, map.data.data(), static_cast<int>(map.width), static_cast<int>(map.height),
                               map.resolution, map.origin_x, map.origin_y, map.origin_yaw);
    if (rc != MCL_OK) {
        log_error("Failed to set map", rc);   // reference: "Failed to get map from map server" (:228)
        return;
    }
    MAX_RANGE_PX = mcl_max_range_px(ctx_);
    map_initialized_ = true;
}

bool ParticleFilter::get_omap(const std::string& map_yaml_path) {
    OccupancyGrid g;
    std::string err;
    if (!load_map(map_yaml_path, g, &err)) {
        std::fprintf(stderr, "[particle_filter] Failed to load map %s: %s\n", map_yaml_path.c_str(), err.c_str());
        return false;
    }
    get_omap(g);
    return map_initialized_;
}

void ParticleFilter::precompute_sensor_model() {
    // :233-292 -- built inside mcl_set_map from the same parameters; nothing to do here
}

void ParticleFilter::lidarCB(float angle_min, float angle_increment, const std::vector<float>& ranges) {
    if (laser_angles_.empty()) {
        laser_angles_.resize(ranges.size());
        for (size_t i = 0; i < ranges.size(); ++i) laser_angles_[i] = angle_min + i * angle_increment;   // float32 (:303)
        for (size_t i = 0; i < laser_angles_.size(); i += static_cast<size_t>(p_.angle_step))
            downsampled_angles_.push_back(laser_angles_[i]);
        const int rc = mcl_set_beam_angles(ctx_, downsampled_angles_.data(), static_cast<int>(downsampled_angles_.size()));
        if (rc != MCL_OK) {
            log_error("LiDAR initialisation failed", rc);
            laser_angles_.clear();
            downsampled_angles_.clear();
            return;
        }
    }
    downsampled_ranges_.clear();
    for (size_t i = 0; i < ranges.size(); i += static_cast<size_t>(p_.angle_step)) downsampled_ranges_.push_back(ranges[i]);
    lidar_initialized_ = true;
}

void ParticleFilter::initialize_particles_pose(const Vector3d& pose) {
    const int rc = mcl_init_pose(ctx_, 0, pose.data(), nullptr);
    if (rc != MCL_OK) log_error("initialize_particles_pose", rc);
}

void ParticleFilter::initialize_global() {
    if (!map_initialized_) return;   // :403-404
    const int rc = mcl_init_global(ctx_, 0, nullptr, nullptr);
    if (rc != MCL_OK) log_error("initialize_global", rc);   // "No free space found in map!" (:425)
}

// MCL(action, observation) followed by expected_pose(), as timer_update calls them (:777-778): ONE C-ABI call
UpdateShell::MclResult ParticleFilter::run_mcl(const Vector3d& action, const std::vector<float>& observation) {
    UpdateShell::MclResult r;
    const auto t0 = std::chrono::steady_clock::now();
    double pose[3];
    const int rc = mcl_update(ctx_, action.data(), observation.data(), static_cast<int>(observation.size()), nullptr, pose);
    if (rc != MCL_OK) {
        log_error("MCL", rc);   // e.g. a scan whose downsampled length differs from the first scan's
        return r;               // ok == false: the caller leaves every piece of tracking state alone
    }
    r.pose = {pose[0], pose[1], pose[2]};
    r.elapsed_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    r.ok = true;
    last_update_ms_ = r.elapsed_ms;
    return r;
}

void ParticleFilter::MCL(const Vector3d& action, const std::vector<float>& observation) {
    const UpdateShell::MclResult r = run_mcl(action, observation);
    if (!r.ok) return;
    shell_.inferred_pose_ = r.pose;           // expected_pose() of the same update (:778)
    shell_.window_total_ms_ += r.elapsed_ms;  // timing_stats_.total_mcl_time / measurement_count (:692-693)
    ++shell_.window_count_;
}

Vector3d ParticleFilter::expected_pose() {
    double pose[3] = {0, 0, 0};
    const int rc = mcl_expected_pose(ctx_, 0, pose);
    if (rc != MCL_OK) log_error("expected_pose", rc);
    return {pose[0], pose[1], pose[2]};
}

std::vector<float> ParticleFilter::calc_range_many(const std::vector<double>& q) {
    const int64_t n = static_cast<int64_t>(q.size() / 3);
    std::vector<float> out(static_cast<size_t>(n), static_cast<float>(p_.max_range));
    if (!map_initialized_ || n == 0) return out;
    const int rc = mcl_calc_range_many(ctx_, q.data(), n, out.data());
    if (rc != MCL_OK) log_error("calc_range_many", rc);
    return out;
}

float ParticleFilter::cast_ray(double x, double y, double angle) {
    if (!map_initialized_) return static_cast<float>(p_.max_range);   // :613-614
    float r = static_cast<float>(p_.max_range);
    const int rc = mcl_cast_ray(ctx_, x, y, angle, &r);
    if (rc != MCL_OK) log_error("cast_ray", rc);
    return r;
}

bool ParticleFilter::update(double dt, double current_velocity, double current_angular_vel) {
    if (!map_initialized_) return false;                              // :722-724
    if (dt > 1.0) return false;                                        // :750-752
    if (!lidar_initialized_ || downsampled_ranges_.empty()) return false;   // :758
    ++shell_.iters_;
    Vector3d action{{0.0, 0.0, 0.0}};
    const bool apply_motion = dt >= 0.0001;                            // :754
    if (apply_motion && (std::abs(current_velocity) > 0.0001 || std::abs(current_angular_vel) > 0.0001)) {
        action[0] = current_velocity * dt;                             // :764-766
        action[1] = 0.0;
        action[2] = current_angular_vel * dt;
    }
    const std::vector<float> observation = downsampled_ranges_;        // :774
    MCL(action, observation);                                          // :777-778
    return true;
}

// ---- the node's update shell: logic in host/update_shell.hpp, hooks bound here ----

namespace {
// standard normal for the start-up jitter only (three draws per tick for the first 15 ticks);
// the reference draws these from its mt19937 (:769-771), which is not reproducible either
double jitter_normal(uint64_t& s) {
    auto next = [&]() {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        return (static_cast<double>(s >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    };
    const double u1 = next(), u2 = next();
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * M_PI * u2);
}
}  // namespace

void ParticleFilter::odomCB(const Vector3d& odom_pose, double linear_velocity, double angular_velocity) {
    shell_.odomCB(odom_pose, linear_velocity, angular_velocity, map_initialized_);
}

void ParticleFilter::clicked_pose(const Vector3d& pose) {
    shell_.clicked_pose(pose, [this](const Vector3d& p) { initialize_particles_pose(p); });
}

bool ParticleFilter::timer_update(double dt) {
    return shell_.timer_update(
        dt, map_initialized_, lidar_initialized_, downsampled_ranges_,
        [this](const Vector3d& action, const std::vector<float>& obs) { return run_mcl(action, obs); },
        [this]() { return jitter_normal(jitter_state_); });
}

Vector3d ParticleFilter::get_current_pose() {
    return shell_.get_current_pose(map_initialized_, [this](Vector3d* c) {   // particles_.colwise().mean() (:904)
        const std::vector<double> p = particles();
        const size_t n = static_cast<size_t>(p_.max_particles);
        if (n == 0) return false;
        for (int k = 0; k < 3; ++k) {
            double s = 0.0;
            for (size_t i = 0; i < n; ++i) s += p[k * n + i];
            (*c)[k] = s / static_cast<double>(n);
        }
        return true;
    });
}

std::vector<double> ParticleFilter::particles() const {
    std::vector<double> out(static_cast<size_t>(3) * p_.max_particles);
    const int rc = mcl_get_particles(ctx_, 0, out.data());
    if (rc != MCL_OK) log_error("particles", rc);
    return out;
}

std::vector<double> ParticleFilter::weights() const {
    std::vector<double> out(static_cast<size_t>(p_.max_particles));
    const int rc = mcl_get_weights(ctx_, 0, out.data());
    if (rc != MCL_OK) log_error("weights", rc);
    return out;
}

std::vector<double> ParticleFilter::sample_particles(int k) const {
    std::vector<double> out(static_cast<size_t>(3) * k);
    const int rc = mcl_sample_particles(ctx_, 0, k, out.data());
    if (rc != MCL_OK) log_error("sample_particles", rc);
    return out;
}

}  // namespace particle_filter_cpp
