// oracle/mcl_oracle.cpp -- CPU oracle for the MCL update path (TEST INFRASTRUCTURE ONLY).
//
// Restates, operation for operation, the arithmetic of the reference's hot path so the
// CUDA implementation can be diffed against it on identical inputs and identical noise.
// Nothing under monte_carlo_localization_b200/ may include, link or call this file.
//
// Reference lines followed (all in /root/reference):
//   sensor table   src/particle_filter.cpp:233-292
//   motion model   src/particle_filter.cpp:449-503, src/utils.cpp:43-48
//   sensor model   src/particle_filter.cpp:506-583
//   ray batch/march src/particle_filter.cpp:586-650
//   MCL driver     src/particle_filter.cpp:652-694
//   expected pose  src/particle_filter.cpp:696-716
//   initialisers   src/particle_filter.cpp:382-446
//   libstdc++ 13   bits/random.tcc:2657-2714 (discrete_distribution), :1812-1844 (normal),
//                  :3349-3381 (generate_canonical)
//
// Build: g++ -std=c++17 -O3 -fopenmp, no -march=native, no -ffast-math
// (reference CMakeLists.txt:5-11, 73-76) so no FMA contraction and no reassociation.
//
// Parity status: PINNED against the unmodified reference sources (oracle/_ref).

#include "mcl_oracle.h"

#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

namespace {

using clk = std::chrono::high_resolution_clock;
inline double ms_since(clk::time_point t0) {
    return std::chrono::duration<double, std::milli>(clk::now() - t0).count();
}

}  // namespace

struct orc_filter {
    orc_params prm{};
    int N = 0;
    // particles_: N x 3 column-major == three contiguous columns
    std::vector<double> px, py, pt;
    std::vector<double> w;      // weights_
    std::vector<double> w_raw;  // weights before normalisation (diagnostic)
    // map
    std::vector<int8_t> grid;
    int W = 0, H = 0;
    double res = 0.0, ox = 0.0, oy = 0.0, oyaw = 0.0;
    int M = 0;  // MAX_RANGE_PX
    bool map_ok = false;
    std::vector<int32_t> free_cells;  // row*W+col of cells == 0, row-major scan order
    // sensor table (M+1)^2 column-major
    std::vector<double> table;
    // beams
    std::vector<float> angles;
    // scratch kept between updates like the reference's queries_/ranges_
    std::vector<double> qx, qy, qa;
    std::vector<float> ranges;
    double inv_squash = 1.0 / 2.2;
    orc_timing tm{};
};

struct orc_rng {
    std::mt19937 eng;
    std::normal_distribution<double> nd{0.0, 1.0};
    explicit orc_rng(uint32_t s) : eng(s) {}
};

static double wrap_angle(double a) {  // src/utils.cpp:43-48
    while (a > M_PI) a -= 2.0 * M_PI;
    while (a < -M_PI) a += 2.0 * M_PI;
    return a;
}

// src/particle_filter.cpp:233-292
static void build_table(orc_filter* f) {
    const int tw = f->M + 1;
    f->table.assign(static_cast<size_t>(tw) * tw, 0.0);
    const double zs = f->prm.z_short, zm = f->prm.z_max, zr = f->prm.z_rand, zh = f->prm.z_hit;
    const double sg = f->prm.sigma_hit;
    for (int d = 0; d < tw; ++d) {
        double* col = &f->table[static_cast<size_t>(d) * tw];
        double norm = 0.0;
        for (int r = 0; r < tw; ++r) {
            double prob = 0.0;
            const double z = static_cast<double>(r - d);
            prob += zh * std::exp(-(z * z) / (2.0 * sg * sg)) / (sg * std::sqrt(2.0 * M_PI));
            if (r < d) prob += 2.0 * zs * (d - r) / static_cast<double>(d);
            if (r == f->M) prob += zm;
            if (r < f->M) prob += zr * 1.0 / static_cast<double>(f->M);
            norm += prob;
            col[r] = prob;
        }
        if (norm > 0) {
            for (int r = 0; r < tw; ++r) col[r] /= norm;  // Eigen >= 3.3: true division
        }
    }
}

// src/particle_filter.cpp:611-650
static inline float march(const orc_filter* f, double x, double y, double angle) {
    if (!f->map_ok) return static_cast<float>(f->prm.max_range);
    const double res = f->res;
    const double dx = std::cos(angle) * res;
    const double dy = std::sin(angle) * res;
    double cx = x, cy = y;
    const int W = f->W, H = f->H;
    const int8_t* g = f->grid.data();
    const int cells = W * H;
    for (int step = 0; step < f->M; ++step) {
        cx += dx;
        cy += dy;
        const int gx = static_cast<int>((cx - f->ox) / res);
        const int gy = static_cast<int>((cy - f->oy) / res);
        if (gx < 0 || gx >= W || gy < 0 || gy >= H) return step * res;
        const int k = gy * W + gx;
        if (k >= 0 && k < cells && g[k] > 50) return step * res;
    }
    return static_cast<float>(f->prm.max_range);
}

// src/particle_filter.cpp:586-609
static void range_batch(orc_filter* f, const double* qx, const double* qy, const double* qa,
                        int64_t n, float* out) {
    auto t0 = clk::now();
    if (f->prm.use_parallel_raycasting) {
#pragma omp parallel for schedule(dynamic)
        for (int64_t i = 0; i < n; ++i) out[i] = march(f, qx[i], qy[i], qa[i]);
    } else {
        for (int64_t i = 0; i < n; ++i) out[i] = march(f, qx[i], qy[i], qa[i]);
    }
    f->tm.raycast_ms += ms_since(t0);
}

// src/particle_filter.cpp:449-503
static void motion(orc_filter* f, double* X, double* Y, double* T, const double action[3],
                   const double* z) {
    double dt = 0.01, vel = 0.0, omega = 0.0;
    const double fwd = action[0];
    const double ang = action[2];
    if (std::abs(fwd) > 0.001) {
        if (std::abs(fwd) < 0.1)
            dt = std::abs(fwd) / 1.0;
        else
            dt = std::abs(fwd) / 5.0;
        dt = std::max(0.001, std::min(dt, 0.1));
        vel = fwd / dt;
    }
    if (std::abs(ang) > 0.001) omega = ang / dt;

    const double sx = f->prm.motion_dispersion_x, sy = f->prm.motion_dispersion_y,
                 st = f->prm.motion_dispersion_theta;
    const int N = f->N;
    for (int i = 0; i < N; ++i) {
        const double x = X[i], y = Y[i], th = T[i];
        double nx, ny, nt;
        if (std::abs(omega) < 1e-6) {
            nx = x + vel * dt * std::cos(th);
            ny = y + vel * dt * std::sin(th);
            nt = th;
        } else {
            const double radius = vel / omega;
            const double dth = omega * dt;
            nx = x + radius * (std::sin(th + dth) - std::sin(th));
            ny = y - radius * (std::cos(th + dth) - std::cos(th));
            nt = th + dth;
        }
        nx += z[3 * i + 0] * sx;
        ny += z[3 * i + 1] * sy;
        nt += z[3 * i + 2] * st;
        X[i] = nx;
        Y[i] = ny;
        T[i] = wrap_angle(nt);
    }
}

// src/particle_filter.cpp:506-583
static void sensor(orc_filter* f, const double* X, const double* Y, const double* T,
                   const float* obs, int n_obs, double* weights) {
    const int R = static_cast<int>(f->angles.size());
    const int N = f->N;
    const size_t NR = static_cast<size_t>(N) * R;
    if (f->qx.size() != NR) {
        f->qx.assign(NR, 0.0);
        f->qy.assign(NR, 0.0);
        f->qa.assign(NR, 0.0);
        f->ranges.assign(NR, 0.f);
    }
    auto t0 = clk::now();
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j < R; ++j) {
            const size_t k = static_cast<size_t>(i) * R + j;
            f->qx[k] = X[i];
            f->qy[k] = Y[i];
            f->qa[k] = T[i] + f->angles[j];
        }
    }
    f->tm.query_ms += ms_since(t0);

    range_batch(f, f->qx.data(), f->qy.data(), f->qa.data(), static_cast<int64_t>(NR),
                f->ranges.data());

    auto t1 = clk::now();
    const int M = f->M;
    const double res = f->res;
    std::vector<float> obs_px(n_obs);
    std::vector<float> rng_px(NR);
    for (int i = 0; i < n_obs; ++i) {
        obs_px[i] = obs[i] / res;
        if (obs_px[i] > M) obs_px[i] = M;
    }
    for (size_t i = 0; i < NR; ++i) {
        rng_px[i] = f->ranges[i] / res;
        if (rng_px[i] > M) rng_px[i] = M;
    }
    const int tw = M + 1;
    const double* tab = f->table.data();
    for (int i = 0; i < N; ++i) {
        double acc = 1.0;
        for (int j = 0; j < R; ++j) {
            int oi = static_cast<int>(std::round(obs_px[j]));
            int ri = static_cast<int>(std::round(rng_px[static_cast<size_t>(i) * R + j]));
            oi = std::max(0, std::min(oi, M));
            ri = std::max(0, std::min(ri, M));
            acc *= tab[static_cast<size_t>(ri) * tw + oi];  // table(obs, range), column-major
        }
        weights[i] = std::pow(acc, f->inv_squash);
    }
    f->tm.sensor_ms += ms_since(t1);
}

// discrete_distribution::param_type::_M_initialize, random.tcc:2657-2678
static void build_cdf(const double* wts, int n, std::vector<double>& cp) {
    std::vector<double> p(wts, wts + n);
    cp.clear();
    if (n < 2) return;  // libstdc++ clears the table; every draw returns 0
    const double s = std::accumulate(p.begin(), p.end(), 0.0);
    for (double& v : p) v /= s;
    cp.resize(n);
    std::partial_sum(p.begin(), p.end(), cp.begin());
    cp[n - 1] = 1.0;
}

extern "C" {

void orc_default_params(orc_params* p) {
    p->max_particles = 2000;
    p->num_threads = 0;
    p->use_parallel_raycasting = 1;
    p->squash_factor = 2.2;
    p->max_range = 12.0;
    p->z_short = 0.01;
    p->z_max = 0.07;
    p->z_rand = 0.12;
    p->z_hit = 0.80;
    p->sigma_hit = 8.0;
    p->motion_dispersion_x = 0.05;
    p->motion_dispersion_y = 0.025;
    p->motion_dispersion_theta = 0.25;
}

orc_filter* orc_create(const orc_params* p) {
    auto* f = new orc_filter();
    f->prm = *p;
    f->N = p->max_particles;
    f->inv_squash = 1.0 / p->squash_factor;
    f->px.assign(f->N, 0.0);
    f->py.assign(f->N, 0.0);
    f->pt.assign(f->N, 0.0);
    f->w.assign(f->N, 1.0 / f->N);
    f->w_raw.assign(f->N, 0.0);
    if (p->use_parallel_raycasting) {
        int nt = p->num_threads == 0 ? omp_get_max_threads() : p->num_threads;
        omp_set_num_threads(nt);
    }
    return f;
}

void orc_destroy(orc_filter* f) { delete f; }

int orc_set_map(orc_filter* f, const int8_t* data, int width, int height, float resolution,
                double ox, double oy, double oyaw) {
    if (!f || !data || width <= 0 || height <= 0) return -1;
    f->grid.assign(data, data + static_cast<size_t>(width) * height);
    f->W = width;
    f->H = height;
    f->res = resolution;  // double <- float32, as map_resolution_ = info.resolution (:191)
    f->ox = ox;
    f->oy = oy;
    f->oyaw = oyaw;
    f->M = static_cast<int>(f->prm.max_range / f->res);  // :195
    f->free_cells.clear();
    for (int r = 0; r < height; ++r)
        for (int c = 0; c < width; ++c)
            if (f->grid[static_cast<size_t>(r) * width + c] == 0) f->free_cells.push_back(r * width + c);
    f->map_ok = true;
    if (f->res <= 0.0) return -2;  // :236-240
    build_table(f);
    return 0;
}

int orc_max_range_px(const orc_filter* f) { return f->M; }

int orc_get_sensor_table(const orc_filter* f, double* out) {
    std::memcpy(out, f->table.data(), f->table.size() * sizeof(double));
    return 0;
}

int orc_set_beam_angles(orc_filter* f, const float* a, int n) {
    f->angles.assign(a, a + n);
    return 0;
}

int orc_set_state(orc_filter* f, const double* P, const double* wts) {
    const int N = f->N;
    if (P) {
        std::copy(P, P + N, f->px.begin());
        std::copy(P + N, P + 2 * N, f->py.begin());
        std::copy(P + 2 * N, P + 3 * N, f->pt.begin());
    }
    if (wts) std::copy(wts, wts + N, f->w.begin());
    return 0;
}

int orc_get_state(const orc_filter* f, double* P, double* wts) {
    const int N = f->N;
    if (P) {
        std::copy(f->px.begin(), f->px.end(), P);
        std::copy(f->py.begin(), f->py.end(), P + N);
        std::copy(f->pt.begin(), f->pt.end(), P + 2 * N);
    }
    if (wts) std::copy(f->w.begin(), f->w.end(), wts);
    return 0;
}

// :382-399
int orc_init_pose(orc_filter* f, const double pose[3], const double* z) {
    const int N = f->N;
    std::fill(f->w.begin(), f->w.end(), 1.0 / N);
    for (int i = 0; i < N; ++i) {
        f->px[i] = pose[0] + z[3 * i + 0] * 0.5;
        f->py[i] = pose[1] + z[3 * i + 1] * 0.5;
        f->pt[i] = wrap_angle(pose[2] + z[3 * i + 2] * 0.4);
    }
    return 0;
}

// :401-446
int orc_init_global(orc_filter* f, const int32_t* cell, const double* theta) {
    if (!f->map_ok) return -1;
    if (f->free_cells.empty()) return -2;
    const int N = f->N;
    for (int i = 0; i < N; ++i) {
        const int32_t lin = f->free_cells[cell[i]];
        const int row = lin / f->W, col = lin % f->W;
        f->px[i] = col * f->res + f->ox;
        f->py[i] = row * f->res + f->oy;
        f->pt[i] = theta[i];
    }
    std::fill(f->w.begin(), f->w.end(), 1.0 / N);
    return 0;
}

int orc_num_free_cells(const orc_filter* f) { return static_cast<int>(f->free_cells.size()); }

float orc_cast_ray(const orc_filter* f, double x, double y, double a) { return march(f, x, y, a); }

int orc_calc_range_many(orc_filter* f, const double* Q, int64_t n, float* out) {
    range_batch(f, Q, Q + n, Q + 2 * n, n, out);
    return 0;
}

// :652-694
int orc_update(orc_filter* f, const double action[3], const float* obs, int n_obs,
               const double* u, const double* z, int32_t* idx_out) {
    const int N = f->N;
    auto t_all = clk::now();

    auto t0 = clk::now();
    std::vector<double> cp;
    build_cdf(f->w.data(), N, cp);
    std::vector<double> nx(N), ny(N), nt(N);
    for (int i = 0; i < N; ++i) {
        int k = 0;
        if (!cp.empty()) k = static_cast<int>(std::lower_bound(cp.begin(), cp.end(), u[i]) - cp.begin());
        if (idx_out) idx_out[i] = k;
        nx[i] = f->px[k];
        ny[i] = f->py[k];
        nt[i] = f->pt[k];
    }
    f->tm.resample_ms += ms_since(t0);

    auto t1 = clk::now();
    motion(f, nx.data(), ny.data(), nt.data(), action, z);
    f->tm.motion_ms += ms_since(t1);

    sensor(f, nx.data(), ny.data(), nt.data(), obs, n_obs, f->w.data());
    f->w_raw = f->w;

    const double s = std::accumulate(f->w.begin(), f->w.end(), 0.0);
    if (s > 0)
        for (double& v : f->w) v /= s;

    f->px = nx;
    f->py = ny;
    f->pt = nt;
    f->tm.total_ms += ms_since(t_all);
    f->tm.count++;
    return 0;
}

// :696-716
int orc_expected_pose(orc_filter* f, double pose[3]) {
    auto t0 = clk::now();
    double ax = 0.0, ay = 0.0, ss = 0.0, sc = 0.0;
    for (int i = 0; i < f->N; ++i) {
        ax += f->w[i] * f->px[i];
        ay += f->w[i] * f->py[i];
        ss += f->w[i] * std::sin(f->pt[i]);
        sc += f->w[i] * std::cos(f->pt[i]);
    }
    pose[0] = ax;
    pose[1] = ay;
    pose[2] = std::atan2(ss, sc);
    f->tm.pose_ms += ms_since(t0);
    return 0;
}

int orc_get_ranges(const orc_filter* f, float* out) {
    std::memcpy(out, f->ranges.data(), f->ranges.size() * sizeof(float));
    return 0;
}

int orc_get_raw_weights(const orc_filter* f, double* out) {
    std::memcpy(out, f->w_raw.data(), f->w_raw.size() * sizeof(double));
    return 0;
}

double orc_mean_cells_per_ray(const orc_filter* f) {
    if (f->ranges.empty()) return 0.0;
    double tot = 0.0;
    const float maxr = static_cast<float>(f->prm.max_range);
    for (float r : f->ranges) {
        if (r == maxr) {
            tot += f->M;
        } else {
            float px = r / f->res;
            tot += static_cast<int>(std::round(px)) + 1;
        }
    }
    return tot / static_cast<double>(f->ranges.size());
}

void orc_get_timing(const orc_filter* f, orc_timing* t) { *t = f->tm; }
void orc_reset_timing(orc_filter* f) { f->tm = orc_timing{}; }

int orc_motion_model(orc_filter* f, double* P, const double action[3], const double* z) {
    motion(f, P, P + f->N, P + 2 * f->N, action, z);
    return 0;
}

int orc_sensor_weights(orc_filter* f, const double* P, const float* obs, int n_obs, double* out) {
    sensor(f, P, P + f->N, P + 2 * f->N, obs, n_obs, out);
    return 0;
}

int orc_resample_indices(const double* wts, int n, const double* u, int nd, int32_t* idx,
                         double* cdf_out) {
    std::vector<double> cp;
    build_cdf(wts, n, cp);
    for (int i = 0; i < nd; ++i)
        idx[i] = cp.empty() ? 0 : static_cast<int32_t>(std::lower_bound(cp.begin(), cp.end(), u[i]) - cp.begin());
    if (cdf_out && !cp.empty()) std::memcpy(cdf_out, cp.data(), sizeof(double) * n);
    return 0;
}

double orc_normalize_angle(double a) { return wrap_angle(a); }

orc_rng* orc_rng_create(uint32_t seed) { return new orc_rng(seed); }
void orc_rng_destroy(orc_rng* r) { delete r; }

void orc_rng_canonical(orc_rng* r, int64_t n, double* out) {
    // what discrete_distribution::operator() pulls through _Adaptor<_, double>
    for (int64_t i = 0; i < n; ++i) out[i] = std::generate_canonical<double, 53>(r->eng);
}

void orc_rng_normal(orc_rng* r, int64_t n, double* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = r->nd(r->eng);
}

void orc_rng_uniform_int(orc_rng* r, int64_t n, int32_t lo, int32_t hi, int32_t* out) {
    std::uniform_int_distribution<int> d(lo, hi);
    for (int64_t i = 0; i < n; ++i) out[i] = d(r->eng);
}

void orc_rng_uniform_real(orc_rng* r, int64_t n, double lo, double hi, double* out) {
    std::uniform_real_distribution<double> d(lo, hi);
    for (int64_t i = 0; i < n; ++i) out[i] = d(r->eng);
}

void orc_rng_global_init(orc_rng* r, int64_t n, int32_t n_free, int32_t* cell, double* theta) {
    std::uniform_int_distribution<int> pos(0, n_free - 1);
    std::uniform_real_distribution<double> ang(0.0, 2.0 * M_PI);
    for (int64_t i = 0; i < n; ++i) {
        cell[i] = pos(r->eng);
        theta[i] = ang(r->eng);
    }
}

uint32_t orc_rng_raw(orc_rng* r) { return static_cast<uint32_t>(r->eng()); }

}  // extern "C"
