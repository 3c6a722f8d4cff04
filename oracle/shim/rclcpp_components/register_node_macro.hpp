// shim: component registration is a no-op outside ROS (TEST INFRASTRUCTURE)
#pragma once
#define RCLCPP_COMPONENTS_REGISTER_NODE(cls) static_assert(sizeof(cls) > 0, "shim");
