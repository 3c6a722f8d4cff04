// host/sharded_main.cpp -- C++ caller of the particle-sharded filter (mcl_create_sharded): one process per
// GPU, ONE global filter.  The binary forks `world` ranks before any CUDA call; rank 0 obtains the NCCL id
// (mcl_nccl_unique_id) and the parent hands it to the other ranks through pipes -- the "caller-supplied
// bootstrap" of include/mcl_b200.h.  Every rank then drives the same replay (the scan is cast from the
// ground truth by the filter's own calc_range_many) and reports its view of the whole filter's pose; the
// parent checks that the ranks agree bit for bit and track the ground truth, and prints one JSON line.
//
//   mcl_sharded <map.yaml> [--world W] [--particles N_PER_GPU] [--steps K] [--x X --y Y --theta T] [--nccl-barrier]
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mcl_b200.h"
#include "map_loader.hpp"

using particle_filter_cpp::OccupancyGrid;

namespace {

struct Options {
    std::string map_yaml;
    int world = 2, steps = 20;
    long particles = 262144;
    double x = NAN, y = NAN, theta = 0.0;
    bool nccl_barrier = false;
};

bool read_all(int fd, void* buf, size_t n) {
    size_t got = 0;
    while (got < n) {
        const ssize_t r = read(fd, static_cast<char*>(buf) + got, n - got);
        if (r <= 0) return false;
        got += static_cast<size_t>(r);
    }
    return true;
}

struct Report {
    double pose[3];
    double err;
    double ms_per_update;
    int ok;
};

int run_rank(const Options& o, int rank, int id_in, int id_out, int report_fd) {
    Report rep{};
    auto fail = [&](const char* what, int rc) {
        std::fprintf(stderr, "[rank %d] %s: %s (%s)\n", rank, what, mcl_status_str(rc), mcl_last_error());
        rep.ok = 0;
        if (write(report_fd, &rep, sizeof rep) != static_cast<ssize_t>(sizeof rep)) return 3;
        return 1;
    };
    char id[128];
    if (rank == 0) {
        const int rc = mcl_nccl_unique_id(id, sizeof id);
        if (rc != MCL_OK) return fail("mcl_nccl_unique_id", rc);
        if (write(id_out, id, sizeof id) != static_cast<ssize_t>(sizeof id)) return 3;
    } else if (!read_all(id_in, id, sizeof id)) {
        return 3;
    }
    OccupancyGrid grid;
    std::string err;
    if (!particle_filter_cpp::load_map(o.map_yaml, grid, &err)) {
        std::fprintf(stderr, "[rank %d] map: %s\n", rank, err.c_str());
        return 2;
    }
    mcl_params p;
    mcl_default_params(&p);
    p.max_particles = static_cast<int32_t>(o.particles * o.world);   // particles of the WHOLE filter
    p.seed = 20250;
    mcl_ctx* ctx = nullptr;
    int rc = mcl_create_sharded(&p, rank, o.world, rank, id, &ctx);
    if (rc != MCL_OK) return fail("mcl_create_sharded", rc);
    rc = mcl_set_map(ctx, grid.data.data(), static_cast<int>(grid.width), static_cast<int>(grid.height), grid.resolution,
                     grid.origin_x, grid.origin_y, grid.origin_yaw);
    if (rc != MCL_OK) return fail("mcl_set_map", rc);
    const int nb = 1080, step = 18;
    const float amin = -2.35f, ainc = 4.7f / 1079.0f;
    std::vector<float> angles;
    for (int i = 0; i < nb; i += step) angles.push_back(amin + i * ainc);
    rc = mcl_set_beam_angles(ctx, angles.data(), static_cast<int>(angles.size()));
    if (rc != MCL_OK) return fail("mcl_set_beam_angles", rc);
    if (o.nccl_barrier) {
        rc = mcl_shard_set_exchange(ctx, 0, nullptr, nullptr);
        if (rc != MCL_OK) return fail("mcl_shard_set_exchange", rc);
    }
    double gt[3] = {o.x, o.y, o.theta};
    if (std::isnan(gt[0])) {   // default start: the free cell closest to the middle of the grid
        double best = 1e300;
        for (uint32_t r = 0; r < grid.height; ++r)
            for (uint32_t c = 0; c < grid.width; ++c)
                if (grid.data[static_cast<size_t>(r) * grid.width + c] == 0) {
                    const double d = std::hypot(double(r) - grid.height / 2.0, double(c) - grid.width / 2.0);
                    if (d < best) {
                        best = d;
                        gt[0] = (c + 0.5) * grid.resolution + grid.origin_x;
                        gt[1] = (r + 0.5) * grid.resolution + grid.origin_y;
                    }
                }
    }
    rc = mcl_init_pose(ctx, 0, gt, nullptr);
    if (rc != MCL_OK) return fail("mcl_init_pose", rc);
    const double v = 1.0, dt = 0.025;
    double pose[3] = {0, 0, 0}, total_ms = 0.0;
    std::vector<double> q(static_cast<size_t>(3) * angles.size());
    std::vector<float> scan(angles.size());
    for (int t = 0; t < o.steps; ++t) {
        float ahead = 0.f;
        rc = mcl_cast_ray(ctx, gt[0], gt[1], gt[2], &ahead);
        if (rc != MCL_OK) return fail("mcl_cast_ray", rc);
        const double vel = ahead > 1.0f ? v : 0.0;
        gt[0] += vel * dt * std::cos(gt[2]);
        gt[1] += vel * dt * std::sin(gt[2]);
        const size_t R = angles.size();
        for (size_t i = 0; i < R; ++i) {
            q[i] = gt[0];
            q[R + i] = gt[1];
            q[2 * R + i] = gt[2] + static_cast<double>(angles[i]);
        }
        rc = mcl_calc_range_many(ctx, q.data(), static_cast<int64_t>(R), scan.data());
        if (rc != MCL_OK) return fail("mcl_calc_range_many", rc);
        const double action[3] = {vel * dt, 0.0, 0.0};
        const auto t0 = std::chrono::steady_clock::now();
        rc = mcl_update(ctx, action, scan.data(), static_cast<int>(R), nullptr, pose);   // every rank gets the whole filter's pose
        if (rc != MCL_OK) return fail("mcl_update", rc);
        if (t >= o.steps / 2) total_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    // whole-filter read-back through the library's communicator: the normalised weights sum to one
    std::vector<double> w(static_cast<size_t>(o.particles) * o.world);
    rc = mcl_sharded_gather(ctx, nullptr, w.data());
    if (rc != MCL_OK) return fail("mcl_sharded_gather", rc);
    double sw = 0.0;
    for (double x : w) sw += x;
    rep.ok = std::abs(sw - 1.0) < 1e-9 ? 1 : 0;
    for (int k = 0; k < 3; ++k) rep.pose[k] = pose[k];
    rep.err = std::hypot(pose[0] - gt[0], pose[1] - gt[1]);
    rep.ms_per_update = total_ms / std::max(1, o.steps - o.steps / 2);
    mcl_destroy(ctx);
    return write(report_fd, &rep, sizeof rep) == static_cast<ssize_t>(sizeof rep) ? 0 : 3;
}

}  // namespace

int main(int argc, char** argv) {
    Options o;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--world") && i + 1 < argc) o.world = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--particles") && i + 1 < argc) o.particles = std::atol(argv[++i]);
        else if (!std::strcmp(argv[i], "--steps") && i + 1 < argc) o.steps = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--x") && i + 1 < argc) o.x = std::atof(argv[++i]);
        else if (!std::strcmp(argv[i], "--y") && i + 1 < argc) o.y = std::atof(argv[++i]);
        else if (!std::strcmp(argv[i], "--theta") && i + 1 < argc) o.theta = std::atof(argv[++i]);
        else if (!std::strcmp(argv[i], "--nccl-barrier")) o.nccl_barrier = true;
        else if (argv[i][0] != '-') o.map_yaml = argv[i];
    }
    if (o.map_yaml.empty() || o.world < 2 || o.world > 16) {
        std::fprintf(stderr, "usage: %s <map.yaml> [--world W>=2] [--particles N_PER_GPU] [--steps K] [--x X --y Y --theta T] [--nccl-barrier]\n", argv[0]);
        return 2;
    }
    // pipes: rank 0 -> parent (the id), parent -> rank r (the id), rank r -> parent (the report)
    std::vector<int> to_rank(o.world * 2), from_rank(o.world * 2);
    int id_pipe[2];
    if (pipe(id_pipe)) return 3;
    for (int r = 0; r < o.world; ++r)
        if (pipe(&to_rank[2 * r]) || pipe(&from_rank[2 * r])) return 3;
    std::vector<pid_t> pids(o.world);
    std::fflush(stdout);
    for (int r = 0; r < o.world; ++r) {
        pids[r] = fork();   // BEFORE any CUDA call: every rank creates its own context on its own GPU
        if (pids[r] == 0) {
            const int rc = run_rank(o, r, to_rank[2 * r], id_pipe[1], from_rank[2 * r + 1]);
            _exit(rc);
        }
    }
    char id[128];
    bool ok = read_all(id_pipe[0], id, sizeof id);
    for (int r = 1; r < o.world && ok; ++r) ok = write(to_rank[2 * r + 1], id, sizeof id) == static_cast<ssize_t>(sizeof id);
    std::vector<Report> reps(o.world);
    for (int r = 0; r < o.world; ++r)
        if (!read_all(from_rank[2 * r], &reps[r], sizeof(Report))) ok = false;
    int bad = 0;
    for (int r = 0; r < o.world; ++r) {
        int st = 0;
        waitpid(pids[r], &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) ++bad;
    }
    bool agree = ok && bad == 0;
    for (int r = 0; r < o.world && agree; ++r)
        agree = reps[r].ok && std::memcmp(reps[r].pose, reps[0].pose, sizeof reps[0].pose) == 0;
    std::printf("{\"world\": %d, \"particles_per_gpu\": %ld, \"steps\": %d, \"exchange\": \"%s\", \"ranks_agree\": %s, "
                "\"pose\": [%.6f, %.6f, %.6f], \"pose_error_m\": %.4f, \"host_ms_per_update\": %.4f, \"failed_ranks\": %d}\n",
                o.world, o.particles, o.steps, o.nccl_barrier ? "nccl barrier" : "fused", agree ? "true" : "false", reps[0].pose[0],
                reps[0].pose[1], reps[0].pose[2], reps[0].err, reps[0].ms_per_update, bad);
    return (agree && reps[0].err < 0.5) ? 0 : 1;
}
