// host/particle_filter.hpp -- ROS-free C++ mirror of particle_filter_cpp::ParticleFilter whose
// hot path runs on the B200 through the C ABI (include/mcl_b200.h).
//
// Same class name, namespace, method names and parameter set as the reference
// (include/particle_filter_cpp/particle_filter.hpp:32-189).  Where the reference uses Eigen
// and ROS message types in signatures, plain std types of the same shape are used
// (Vector3d -> std::array<double,3>, MatrixXd N x 3 column-major -> std::vector<double> of
// x[N] y[N] theta[N]).  The rclcpp node shell (timers, pubs/subs, TF) is out of scope; a ROS 2
// node would own one of these and call lidarCB / MCL / expected_pose from its callbacks exactly
// as timer_update does (src/particle_filter.cpp:756-833).  See INTEGRATION.md.
//
// Error behaviour follows the reference: no exceptions on the path; failures are logged to
// stderr ("[particle_filter] ...") and the call returns early (cf. :228, :238, :425);
// cast_ray without a map returns MAX_RANGE_METERS (:613-614).  The one hard failure is
// construction without a CUDA device: there is no CPU fallback, so the constructor throws.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "map_loader.hpp"
#include "update_shell.hpp"

struct mcl_ctx;

namespace particle_filter_cpp {

// The declared ROS parameters (src/particle_filter.cpp:23-47), same names and defaults.
struct Parameters {
    int angle_step = 18;
    int max_particles = 2000;
    int max_viz_particles = 60;
    double squash_factor = 2.2;
    double max_range = 12.0;
    bool publish_odom = true;
    bool viz = true;
    double z_short = 0.01, z_max = 0.07, z_rand = 0.12, z_hit = 0.80, sigma_hit = 8.0;
    double motion_dispersion_x = 0.05, motion_dispersion_y = 0.025, motion_dispersion_theta = 0.25;
    double lidar_offset_x = 0.0, lidar_offset_y = 0.0, wheelbase = 0.325;
    std::string scan_topic = "/scan", odom_topic = "/odom";
    double timer_frequency = 100.0;
    bool use_parallel_raycasting = true;   // accepted, meaningless on the GPU
    int num_threads = 0;                   // accepted, meaningless on the GPU
    double max_pose_range = 10000.0;
    double delay_compensation_factor = 1.5;
    // not a reference parameter: device + RNG seed (reference seeds from std::random_device, :20)
    int device = 0;
    uint64_t seed = 0x9E3779B97F4A7C15ull;

    // Overlay values from config/mcl_config.yaml (particle_filter: ros__parameters: ...).
    // Unknown keys (range_method, theta_discretization, ... -- never read by the reference
    // either) are ignored.  Returns false if the file cannot be read.
    bool load_yaml(const std::string& path, std::string* err = nullptr);
};

class ParticleFilter {
  public:
    explicit ParticleFilter(const Parameters& params = Parameters());
    ~ParticleFilter();
    ParticleFilter(const ParticleFilter&) = delete;
    ParticleFilter& operator=(const ParticleFilter&) = delete;

    // --- core MCL algorithm (reference: private members, :39-43) ---
    void MCL(const Vector3d& action, const std::vector<float>& observation);
    Vector3d expected_pose();

    // --- initialisation (:46-48) ---
    void initialize_global();
    void initialize_particles_pose(const Vector3d& pose);
    void precompute_sensor_model();   // table is rebuilt on the device side at get_omap

    // --- sensor / map ingest ---
    // lidarCB (:295-323): first call fixes the beam angles, every call refreshes the ranges
    void lidarCB(float angle_min, float angle_increment, const std::vector<float>& ranges);
    // get_omap (:173-230): from an already decoded grid, or from maps/<name>.yaml
    void get_omap(const OccupancyGrid& map);
    bool get_omap(const std::string& map_yaml_path);

    // --- ray casting (:74-75) ---
    std::vector<float> calc_range_many(const std::vector<double>& queries_colmajor);   // n x 3
    float cast_ray(double x, double y, double angle);

    // --- timer_update's call sequence (:761-778): action synthesis + MCL + expected_pose ---
    // velocity/angular velocity as odomCB stores them (:328-329); returns false if skipped
    bool update(double dt, double current_velocity, double current_angular_vel);

    // --- the node's update shell without ROS ---
    // odomCB (:325-352): twist -> velocities, pose -> last_pose_, odometry tracking update
    void odomCB(const Vector3d& odom_pose, double linear_velocity, double angular_velocity);
    // clicked_pose (:355-374): re-initialise around a pose and start odometry tracking from it
    void clicked_pose(const Vector3d& pose);
    // timer_update (:720-846) with the wall-clock dt passed in: action synthesis (odometry, or the
    // decaying start-up jitter :767-772), MCL + expected_pose, odometry-tracking re-anchor with
    // delay compensation (:781-807).  Returns false when the tick is skipped (:722, :750, :758).
    bool timer_update(double dt);
    // get_current_pose (:892-916): odometry tracking > filter estimate > particle mean > last odom
    Vector3d get_current_pose();
    // mean MCL time the delay compensation uses (:792-795), milliseconds: the mean over the CURRENT
    // 200-iteration TimingStats window (:814-827)
    double mean_mcl_ms() const { return shell_.window_count_ ? shell_.window_total_ms_ / shell_.window_count_ : 0.0; }
    bool odom_tracking_active() const { return shell_.odom_tracking_active_; }

    // --- state access (visualize() :944-963, get_current_pose :892-916) ---
    std::vector<double> particles() const;          // column-major N x 3
    std::vector<double> weights() const;
    std::vector<double> sample_particles(int k) const;   // weighted subsample, k x 3 column-major
    Vector3d inferred_pose() const { return shell_.inferred_pose_; }
    bool is_pose_valid(const Vector3d& pose) const { return shell_.is_pose_valid(pose); }
    int iterations() const { return shell_.iters_; }
    int max_range_px() const { return MAX_RANGE_PX; }
    const std::vector<float>& downsampled_angles() const { return downsampled_angles_; }
    const std::vector<float>& downsampled_ranges() const { return downsampled_ranges_; }
    const Parameters& parameters() const { return p_; }
    double last_update_ms() const { return last_update_ms_; }

  private:
    Parameters p_;
    mcl_ctx* ctx_ = nullptr;
    int MAX_RANGE_PX = 0;
    bool map_initialized_ = false, lidar_initialized_ = false;
    std::vector<float> laser_angles_, downsampled_angles_, downsampled_ranges_;
    double last_update_ms_ = 0.0;
    // the node's update shell (timer_update / odomCB / clicked_pose / get_current_pose): pure host state,
    // host/update_shell.hpp; this class binds its MCL / mean / init hooks to the C ABI
    UpdateShell shell_;
    UpdateShell::MclResult run_mcl(const Vector3d& action, const std::vector<float>& observation);
    uint64_t jitter_state_ = 0x243F6A8885A308D3ull;   // start-up jitter noise (:769-771)
};

}  // namespace particle_filter_cpp
