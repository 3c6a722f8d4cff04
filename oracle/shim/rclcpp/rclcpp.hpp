// oracle/shim/rclcpp/rclcpp.hpp -- no-op stand-in for the slice of rclcpp the reference node
// uses (TEST INFRASTRUCTURE).  Parameters come from a process-global override table set by
// oracle/ref_harness.cpp; the GetMap client hands back a grid the harness installed; pubs,
// subs, timers, TF and logging do nothing.  No arithmetic of the MCL path lives here.
#pragma once
#include <chrono>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <variant>

#include "msgs_common.hpp"

namespace rclcpp {

using ParamValue = std::variant<bool, int64_t, double, std::string>;

struct ShimGlobals {
    std::map<std::string, ParamValue> overrides;
    nav_msgs::msg::OccupancyGrid map;
    bool have_map = false;
    bool verbose = false;
    static ShimGlobals& get() {
        static ShimGlobals g;
        return g;
    }
};

class Parameter {
  public:
    Parameter() = default;
    explicit Parameter(ParamValue v) : v_(std::move(v)) {}
    int64_t as_int() const { return std::get<int64_t>(v_); }
    double as_double() const {
        if (std::holds_alternative<int64_t>(v_)) return static_cast<double>(std::get<int64_t>(v_));
        return std::get<double>(v_);
    }
    bool as_bool() const { return std::get<bool>(v_); }
    std::string as_string() const { return std::get<std::string>(v_); }

  private:
    ParamValue v_;
};

class Time {
  public:
    Time() = default;
    explicit Time(int64_t ns) : ns_(ns) {}
    Time(const builtin_interfaces::msg::Time& t) : ns_(int64_t(t.sec) * 1000000000LL + t.nanosec) {}
    int64_t nanoseconds() const { return ns_; }
    operator builtin_interfaces::msg::Time() const {
        builtin_interfaces::msg::Time t;
        t.sec = static_cast<int32_t>(ns_ / 1000000000LL);
        t.nanosec = static_cast<uint32_t>(ns_ % 1000000000LL);
        return t;
    }

  private:
    int64_t ns_ = 0;
};

class Clock {
  public:
    using SharedPtr = std::shared_ptr<Clock>;
    Time now() const {
        auto t = std::chrono::steady_clock::now().time_since_epoch();
        return Time(std::chrono::duration_cast<std::chrono::nanoseconds>(t).count());
    }
};

class Logger {};

class QoS {
  public:
    QoS(int depth) : depth_(depth) {}
    QoS& transient_local() { return *this; }

  private:
    int depth_;
};

class NodeOptions {};

template <typename MsgT>
class Publisher {
  public:
    using SharedPtr = std::shared_ptr<Publisher<MsgT>>;
    void publish(const MsgT&) {}
    size_t get_subscription_count() const { return 0; }
};

template <typename MsgT>
class Subscription {
  public:
    using SharedPtr = std::shared_ptr<Subscription<MsgT>>;
};

class TimerBase {
  public:
    using SharedPtr = std::shared_ptr<TimerBase>;
};

enum class FutureReturnCode { SUCCESS, INTERRUPTED, TIMEOUT };

template <typename T>
class ShimFuture {
  public:
    explicit ShimFuture(std::shared_ptr<T> v) : v_(std::move(v)) {}
    std::shared_ptr<T> get() { return v_; }
    bool valid() const { return static_cast<bool>(v_); }

  private:
    std::shared_ptr<T> v_;
};

template <typename SrvT>
class Client {
  public:
    using SharedPtr = std::shared_ptr<Client<SrvT>>;
    template <typename D>
    bool wait_for_service(D) {
        return true;
    }
    ShimFuture<typename SrvT::Response> async_send_request(std::shared_ptr<typename SrvT::Request>) {
        auto& g = ShimGlobals::get();
        if (!g.have_map) return ShimFuture<typename SrvT::Response>(nullptr);
        auto r = std::make_shared<typename SrvT::Response>();
        r->map = g.map;
        return ShimFuture<typename SrvT::Response>(r);
    }
};

struct NodeBaseInterface {
    using SharedPtr = std::shared_ptr<NodeBaseInterface>;
};

class Node {
  public:
    Node(const std::string& name, const NodeOptions& = NodeOptions()) : name_(name), clock_(std::make_shared<Clock>()) {}
    virtual ~Node() = default;

    template <typename T>
    void declare_parameter(const std::string& name, const T& dflt) {
        ParamValue v;
        if constexpr (std::is_same_v<T, bool>)
            v = dflt;
        else if constexpr (std::is_integral_v<T>)
            v = static_cast<int64_t>(dflt);
        else if constexpr (std::is_floating_point_v<T>)
            v = static_cast<double>(dflt);
        else
            v = std::string(dflt);
        auto& ov = ShimGlobals::get().overrides;
        auto it = ov.find(name);
        if (it != ov.end()) {
            // keep the declared type (ROS would reject a type mismatch)
            if (std::holds_alternative<double>(v) && std::holds_alternative<int64_t>(it->second))
                v = static_cast<double>(std::get<int64_t>(it->second));
            else
                v = it->second;
        }
        params_[name] = v;
    }
    Parameter get_parameter(const std::string& name) const { return Parameter(params_.at(name)); }
    Logger get_logger() const { return Logger{}; }
    Clock::SharedPtr get_clock() const { return clock_; }
    NodeBaseInterface::SharedPtr get_node_base_interface() { return std::make_shared<NodeBaseInterface>(); }

    template <typename MsgT>
    typename Publisher<MsgT>::SharedPtr create_publisher(const std::string&, const QoS&) {
        return std::make_shared<Publisher<MsgT>>();
    }
    template <typename MsgT, typename CB>
    typename Subscription<MsgT>::SharedPtr create_subscription(const std::string&, const QoS&, CB&&) {
        return std::make_shared<Subscription<MsgT>>();
    }
    template <typename SrvT>
    typename Client<SrvT>::SharedPtr create_client(const std::string&) {
        return std::make_shared<Client<SrvT>>();
    }
    template <typename D, typename CB>
    TimerBase::SharedPtr create_wall_timer(D, CB&&) {
        return std::make_shared<TimerBase>();
    }

  private:
    std::string name_;
    Clock::SharedPtr clock_;
    std::map<std::string, ParamValue> params_;
};

template <typename F>
FutureReturnCode spin_until_future_complete(NodeBaseInterface::SharedPtr, F& fut) {
    return fut.valid() ? FutureReturnCode::SUCCESS : FutureReturnCode::INTERRUPTED;
}

inline bool ok() { return true; }
inline void init(int, char**) {}
inline void shutdown() {}
inline void spin(std::shared_ptr<Node>) {}

}  // namespace rclcpp

#define RCLCPP_SHIM_LOG(...)                                    \
    do {                                                        \
        if (rclcpp::ShimGlobals::get().verbose) {               \
            std::fprintf(stderr, "[ref] ");                     \
            std::fprintf(stderr, __VA_ARGS__);                  \
            std::fprintf(stderr, "\n");                         \
        }                                                       \
    } while (0)
#define RCLCPP_INFO(logger, ...) do { (void)(logger); RCLCPP_SHIM_LOG(__VA_ARGS__); } while (0)
#define RCLCPP_WARN(logger, ...) do { (void)(logger); RCLCPP_SHIM_LOG(__VA_ARGS__); } while (0)
#define RCLCPP_ERROR(logger, ...) do { (void)(logger); RCLCPP_SHIM_LOG(__VA_ARGS__); } while (0)
#define RCLCPP_INFO_THROTTLE(logger, clock, period, ...) do { (void)(logger); (void)(clock); } while (0)
