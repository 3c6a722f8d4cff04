// shim: TF broadcasting is a no-op outside ROS (TEST INFRASTRUCTURE)
#pragma once
#include "msgs_common.hpp"
#include "rclcpp/rclcpp.hpp"
namespace tf2_ros {
class TransformBroadcaster {
  public:
    template <typename NodeT>
    explicit TransformBroadcaster(NodeT&) {}
    void sendTransform(const geometry_msgs::msg::TransformStamped&) {}
};
}  // namespace tf2_ros
