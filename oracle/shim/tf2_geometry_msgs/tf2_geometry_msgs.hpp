// shim (TEST INFRASTRUCTURE)
#pragma once
#include "msgs_common.hpp"
#include "tf2/LinearMath/Quaternion.h"
namespace tf2 {
inline geometry_msgs::msg::Quaternion toMsg(const Quaternion& q) {
    geometry_msgs::msg::Quaternion m;
    m.x = q.x();
    m.y = q.y();
    m.z = q.z();
    m.w = q.w();
    return m;
}
}  // namespace tf2
