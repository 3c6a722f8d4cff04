// oracle/shim/ref_clock_prelude.hpp -- controllable clocks for the UNMODIFIED reference (TEST INFRASTRUCTURE).
//
// Forced into every translation unit of oracle/_ref/libref_pf.so (oracle/Makefile: -include).  The
// reference's timer_update reads std::chrono::steady_clock for its dt (src/particle_filter.cpp:735-741)
// and MCL() reads std::chrono::high_resolution_clock for TimingStats (:654-693), which feeds the delay
// compensation (:791-802).  To compare the node shell of monte_carlo_localization_b200/host/ with the
// reference's own timer_update tick by tick, both clocks must be reproducible: after every standard
// header that names them has been included, the two identifiers are redirected to clocks that pass
// through to the real ones unless the harness switches them to scripted time (ref_clock_*).
// No arithmetic of the MCL path lives here.
#pragma once
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <random>
#include <shared_mutex>
#include <string>
#include <thread>
#include <variant>
#include <vector>

namespace ref_clock {
struct State {
    bool fake = false;
    long long steady_ns = 0;       // scripted steady time
    long long hr_ns = 0;           // scripted high-resolution time: advances by the quantum at every now()
    long long hr_quantum_ns = 0;
};
State& state();   // oracle/ref_harness.cpp
}  // namespace ref_clock

namespace std {
namespace chrono {
struct ref_fake_steady_clock {
    using duration = nanoseconds;
    using rep = duration::rep;
    using period = duration::period;
    using time_point = chrono::time_point<ref_fake_steady_clock, duration>;
    static constexpr bool is_steady = true;
    static time_point now() noexcept {
        const auto& s = ref_clock::state();
        if (!s.fake) return time_point(duration_cast<duration>(steady_clock::now().time_since_epoch()));
        return time_point(duration(s.steady_ns));
    }
};
struct ref_fake_hr_clock {
    using duration = nanoseconds;
    using rep = duration::rep;
    using period = duration::period;
    using time_point = chrono::time_point<ref_fake_hr_clock, duration>;
    static constexpr bool is_steady = false;
    static time_point now() noexcept {
        auto& s = ref_clock::state();
        if (!s.fake) return time_point(duration_cast<duration>(high_resolution_clock::now().time_since_epoch()));
        s.hr_ns += s.hr_quantum_ns;
        return time_point(duration(s.hr_ns));
    }
};
}  // namespace chrono
}  // namespace std

#define steady_clock ref_fake_steady_clock
#define high_resolution_clock ref_fake_hr_clock
