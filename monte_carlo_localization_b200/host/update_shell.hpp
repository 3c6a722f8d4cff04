// host/update_shell.hpp -- the node's update shell around MCL(), without ROS and without CUDA.
//
// A restatement of the host-side logic either side of the hot path:
//   timer_update        src/particle_filter.cpp:720-846   action synthesis (odometry, or the decaying
//                       start-up jitter :767-772), MCL + expected_pose, odometry-tracking re-anchor
//                       with delay compensation (:781-807), TimingStats window reset every 200
//                       iterations (:814-827)
//   odomCB              :325-352
//   clicked_pose        :355-374
//   get_current_pose    :892-916 (priority chain)
//   initialize_odom_tracking / update_odom_pose   :988-1013
// Pure state machine: the MCL update, the particle mean, the initialiser and the jitter's normal
// draws are INJECTED, so the same code runs in the product (host/particle_filter.cpp binds them to
// the C ABI) and in the CPU parity test, which drives it beside the reference's own timer_update
// (tests/test_update_shell.py).  The wall-clock dt is an argument: the reference reads
// std::chrono::steady_clock inside timer_update (:735-741).
#pragma once
#include <array>
#include <cmath>
#include <functional>
#include <vector>

namespace particle_filter_cpp {

using Vector3d = std::array<double, 3>;

class UpdateShell {
  public:
    struct MclResult {
        Vector3d pose{{0, 0, 0}};   // expected_pose() of the update (:778)
        double elapsed_ms = 0.0;    // what TimingStats::total_mcl_time receives (:691-693)
        bool ok = false;            // false: the update could not run; the tick changes nothing else
    };
    using MclFn = std::function<MclResult(const Vector3d& action, const std::vector<float>& observation)>;
    using NormalFn = std::function<double()>;                 // normal_dist_(rng_) of the start-up jitter (:769-771)
    using MeanFn = std::function<bool(Vector3d* mean)>;       // particles_.colwise().mean() (:904)
    using InitFn = std::function<void(const Vector3d& pose)>; // initialize_particles_pose (:361)

    double delay_compensation_factor = 1.5;   // :47
    double max_pose_range = 10000.0;          // :46

    // odomCB (:325-352)
    void odomCB(const Vector3d& odom_pose, double linear_velocity, double angular_velocity, bool map_initialized) {
        current_velocity_ = linear_velocity;       // :328-329
        current_angular_vel_ = angular_velocity;
        const bool can_track = pose_initialized_from_rviz_ || (map_initialized && iters_ > 0 && is_pose_valid(inferred_pose_));
        if (can_track && odom_tracking_active_) update_odom_pose(odom_pose);   // :332-337
        last_pose_ = odom_pose;                    // :343-350
        odom_initialized_ = true;
    }

    // clicked_pose (:355-374)
    void clicked_pose(const Vector3d& pose, const InitFn& init) {
        init(pose);                                // :361
        initialize_odom_tracking(pose, true);      // :364
        inferred_pose_ = pose;                     // :367
    }

    // timer_update (:720-846) after its first (timer-initialising) call, with the steady-clock dt passed in.
    // Returns false when the tick does nothing (:722, :750, :758) or the update failed.
    bool timer_update(double dt, bool map_initialized, bool lidar_initialized, const std::vector<float>& downsampled_ranges,
                      const MclFn& mcl, const NormalFn& normal) {
        if (!map_initialized) return false;                                    // :722-724
        const bool has_odom = odom_initialized_;                               // :726
        if (dt > 1.0) return false;                                            // :750-752
        const bool apply_motion = dt >= 0.0001;                                // :754
        if (!lidar_initialized || downsampled_ranges.empty()) return false;    // :758
        ++iters_;                                                              // :759
        Vector3d action{{0.0, 0.0, 0.0}};
        if (has_odom && apply_motion && (std::abs(current_velocity_) > 0.0001 || std::abs(current_angular_vel_) > 0.0001)) {
            action[0] = current_velocity_ * dt;                                // :764-766
            action[1] = 0.0;
            action[2] = current_angular_vel_ * dt;
        } else if (!has_odom && !pose_initialized_from_rviz_ && iters_ < 15) {
            const double noise_factor = std::max(0.1, 1.0 - (static_cast<double>(iters_) / 15.0));   // :768
            action[0] = normal() * 0.02 * noise_factor;
            action[1] = normal() * 0.01 * noise_factor;
            action[2] = normal() * 0.05 * noise_factor;
        }
        last_action_ = action;
        const MclResult r = mcl(action, downsampled_ranges);                   // :774-778
        if (!r.ok) return false;   // (the reference has no failure path here; nothing is re-anchored from a stale pose)
        inferred_pose_ = r.pose;
        window_total_ms_ += r.elapsed_ms;                                      // :692-693
        ++window_count_;
        const bool can_track = has_odom && (pose_initialized_from_rviz_ ||
                                            (map_initialized && iters_ > 0 && is_pose_valid(inferred_pose_)));   // :781-782
        if (can_track) {
            if (!odom_tracking_active_ && is_pose_valid(inferred_pose_)) initialize_odom_tracking(inferred_pose_, false);   // :785-788
            Vector3d compensated = inferred_pose_;                             // :791-802
            if (window_count_ > 0) {
                const double delay = window_total_ms_ / window_count_ / 1000.0;
                const double lon = current_velocity_ * delay * delay_compensation_factor;
                const double ang = current_angular_vel_ * delay * delay_compensation_factor;
                compensated[0] += lon * std::cos(inferred_pose_[2]);
                compensated[1] += lon * std::sin(inferred_pose_[2]);
                compensated[2] += ang;
            }
            odom_reference_pose_ = compensated;                                // :804-806
            odom_reference_odom_ = last_pose_;
            odom_pose_ = compensated;
        }
        if (iters_ % 200 == 0) {   // timing_stats_.reset() (:814-827): the delay is the mean of the CURRENT window
            window_total_ms_ = 0.0;
            window_count_ = 0;
        }
        return true;
    }

    // get_current_pose (:892-916): odometry tracking > filter estimate > particle mean > last odom > origin
    Vector3d get_current_pose(bool map_initialized, const MeanFn& particle_mean) const {
        if (odom_tracking_active_ && is_pose_valid(odom_pose_)) return odom_pose_;   // :895-896
        if (is_pose_valid(inferred_pose_)) return inferred_pose_;                    // :899-900
        if (map_initialized) {                                                       // :903-908
            Vector3d c{{0, 0, 0}};
            if (particle_mean(&c) && is_pose_valid(c)) return c;
        }
        if (is_pose_valid(last_pose_)) return last_pose_;                            // :911-912
        return Vector3d{{0, 0, 0}};
    }

    bool is_pose_valid(const Vector3d& pose) const {   // src/utils.cpp:80-84
        return std::isfinite(pose[0]) && std::isfinite(pose[1]) && std::isfinite(pose[2]) &&
               std::abs(pose[0]) < max_pose_range && std::abs(pose[1]) < max_pose_range;
    }

    // state (particle_filter.hpp:104-113, 170-178)
    int iters_ = 0;
    Vector3d inferred_pose_{{0, 0, 0}};
    Vector3d last_pose_{{0, 0, 0}}, odom_pose_{{0, 0, 0}}, odom_reference_pose_{{0, 0, 0}}, odom_reference_odom_{{0, 0, 0}};
    bool odom_initialized_ = false, pose_initialized_from_rviz_ = false, odom_tracking_active_ = false;
    double current_velocity_ = 0.0, current_angular_vel_ = 0.0;
    double window_total_ms_ = 0.0;   // timing_stats_.total_mcl_time of the current 200-iteration window
    int window_count_ = 0;           // timing_stats_.measurement_count
    Vector3d last_action_{{0, 0, 0}};

  private:
    static double norm3(const Vector3d& v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    void initialize_odom_tracking(const Vector3d& initial_pose, bool from_rviz) {   // :988-1002
        odom_pose_ = initial_pose;
        odom_reference_pose_ = initial_pose;
        if (norm3(last_pose_) > 0) odom_reference_odom_ = last_pose_;
        pose_initialized_from_rviz_ = from_rviz;
        odom_tracking_active_ = true;
    }
    void update_odom_pose(const Vector3d& current_odom) {                            // :1004-1013
        if (!odom_tracking_active_) return;
        for (int k = 0; k < 3; ++k) odom_pose_[k] = odom_reference_pose_[k] + (current_odom[k] - odom_reference_odom_[k]);
    }
};

}  // namespace particle_filter_cpp
