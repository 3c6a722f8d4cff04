// oracle/shim/msgs_common.hpp -- plain-struct stand-ins for the ROS 2 messages the
// reference node names (TEST INFRASTRUCTURE).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace builtin_interfaces::msg {
struct Time {
    int32_t sec = 0;
    uint32_t nanosec = 0;
};
}  // namespace builtin_interfaces::msg

namespace std_msgs::msg {
struct Header {
    builtin_interfaces::msg::Time stamp;
    std::string frame_id;
};
}  // namespace std_msgs::msg

namespace geometry_msgs::msg {
struct Point {
    double x = 0, y = 0, z = 0;
};
struct Vector3 {
    double x = 0, y = 0, z = 0;
};
struct Quaternion {
    double x = 0, y = 0, z = 0, w = 1;
};
struct Pose {
    Point position;
    Quaternion orientation;
};
struct PoseWithCovariance {
    Pose pose;
    double covariance[36] = {};
};
struct Twist {
    Vector3 linear, angular;
};
struct TwistWithCovariance {
    Twist twist;
    double covariance[36] = {};
};
struct PoseStamped {
    using SharedPtr = std::shared_ptr<PoseStamped>;
    std_msgs::msg::Header header;
    Pose pose;
};
struct PoseArray {
    using SharedPtr = std::shared_ptr<PoseArray>;
    std_msgs::msg::Header header;
    std::vector<Pose> poses;
};
struct PoseWithCovarianceStamped {
    using SharedPtr = std::shared_ptr<PoseWithCovarianceStamped>;
    std_msgs::msg::Header header;
    PoseWithCovariance pose;
};
struct PointStamped {
    using SharedPtr = std::shared_ptr<PointStamped>;
    std_msgs::msg::Header header;
    Point point;
};
struct Transform {
    Vector3 translation;
    Quaternion rotation;
};
struct TransformStamped {
    std_msgs::msg::Header header;
    std::string child_frame_id;
    Transform transform;
};
}  // namespace geometry_msgs::msg

namespace nav_msgs::msg {
struct MapMetaData {
    builtin_interfaces::msg::Time map_load_time;
    float resolution = 0.f;
    uint32_t width = 0, height = 0;
    geometry_msgs::msg::Pose origin;
};
struct OccupancyGrid {
    using SharedPtr = std::shared_ptr<OccupancyGrid>;
    std_msgs::msg::Header header;
    MapMetaData info;
    std::vector<int8_t> data;
};
struct Odometry {
    using SharedPtr = std::shared_ptr<Odometry>;
    std_msgs::msg::Header header;
    std::string child_frame_id;
    geometry_msgs::msg::PoseWithCovariance pose;
    geometry_msgs::msg::TwistWithCovariance twist;
};
}  // namespace nav_msgs::msg

namespace nav_msgs::srv {
struct GetMap {
    struct Request {};
    struct Response {
        nav_msgs::msg::OccupancyGrid map;
    };
};
}  // namespace nav_msgs::srv

namespace sensor_msgs::msg {
struct LaserScan {
    using SharedPtr = std::shared_ptr<LaserScan>;
    std_msgs::msg::Header header;
    float angle_min = 0, angle_max = 0, angle_increment = 0, time_increment = 0, scan_time = 0;
    float range_min = 0, range_max = 0;
    std::vector<float> ranges, intensities;
};
}  // namespace sensor_msgs::msg
