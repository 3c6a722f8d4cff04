// shim: yaw<->quaternion helpers with tf2's names; only used off the hot path
// (origin yaw, visualisation).  TEST INFRASTRUCTURE.
#pragma once
#include <cmath>
namespace tf2 {
class Quaternion {
  public:
    Quaternion() : x_(0), y_(0), z_(0), w_(1) {}
    Quaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {}
    void setRPY(double roll, double pitch, double yaw) {
        const double cr = std::cos(roll * 0.5), sr = std::sin(roll * 0.5);
        const double cp = std::cos(pitch * 0.5), sp = std::sin(pitch * 0.5);
        const double cy = std::cos(yaw * 0.5), sy = std::sin(yaw * 0.5);
        x_ = sr * cp * cy - cr * sp * sy;
        y_ = cr * sp * cy + sr * cp * sy;
        z_ = cr * cp * sy - sr * sp * cy;
        w_ = cr * cp * cy + sr * sp * sy;
    }
    double x() const { return x_; }
    double y() const { return y_; }
    double z() const { return z_; }
    double w() const { return w_; }

  private:
    double x_, y_, z_, w_;
};
class Matrix3x3 {
  public:
    explicit Matrix3x3(const Quaternion& q) : q_(q) {}
    void getRPY(double& roll, double& pitch, double& yaw) const {
        const double x = q_.x(), y = q_.y(), z = q_.z(), w = q_.w();
        roll = std::atan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y));
        double s = 2.0 * (w * y - z * x);
        s = s > 1.0 ? 1.0 : (s < -1.0 ? -1.0 : s);
        pitch = std::asin(s);
        yaw = std::atan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z));
    }

  private:
    Quaternion q_;
};
}  // namespace tf2
