// csrc/kernels.cuh -- the sm_100a kernels of the MCL update.
//
// One update of the reference (src/particle_filter.cpp:652-716) becomes, per filter:
//   exact sums / CDF   k_tile_sums, k_exact_pass (x3), k_exact_emit (exact_kernels.cuh) (:658, :679)
//                      (k_exact_single: all of a pass in one CTA for a filter of one tile)
//   resample + motion  k_resample_motion                                         (:661-665, :449-503)
//   heading sort       histogram inside k_resample_motion, k_sort_scatter (processing order only)
//   ray cast + weight  k_prepare_obs, then either k_raycast_weight (isotropic skip map, weights in
//                      its epilogue) or the directional stage of dir_kernels.cuh: k_dir_gather,
//                      k_dir_plan, k_raycast_dir, k_weight_steps                   (:506-650)
//   normalise + pose   inside the exact pass that sums w_raw / S1; k_normalize_pose for one-tile filters
//                      and mcl_expected_pose                                     (:679-686, :696-716)
//   sharded filter     k_route_request + k_route_serve (or k_route): the draws reach the rank that owns
//                      their source particle, the poses come back over NVLink       (:658-665)
// blockIdx.y is the filter of a batch; every per-filter array is [F][...] contiguous.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include <type_traits>

#include "device_utils.cuh"
#include "exact_sum.cuh"
#include "map_prep.h"
#include "march.cuh"
#include "shard.cuh"

namespace mclb200 {

constexpr int kMaxBeams = 128;
constexpr int kMaxBuckets = 4096;     // heading buckets of the coherence sort

// Everything the ray/weight kernels need to know about the map and the beams.  Passed by
// value: kernel parameters live in the constant bank, so the beam tables below are read
// through the constant cache with warp-uniform indices.
struct MapDev {
    const int8_t* grid;    // reference occupancy, W*H
    const uint8_t* v8;     // skip map, PH*PW
    const uint8_t* v4;     // nibble skip map, PH*PW/2
    int W, H, PW, PH;
    double res, ox, oy;
    int M;                 // MAX_RANGE_PX
    // shared-memory window geometry (cells); ww == 0 disables the window path.
    // wbits = 8: one byte per cell (full skip code); 4: two cells per byte (code clamped to 15)
    int ww, wh, wbits;
};

// Shared-memory form of the 4-bit window: a 32-bit shared-space address, so the march issues
// plain LDS.U8 without generic-address conversion.
struct WindowV4S {
    uint32_t saddr;          // shared address of the window's first byte
    int offx, offy, pitch;
    __device__ __forceinline__ int get_p(uint32_t px, uint32_t py) const {
        return get(static_cast<int>(px >> kFrac), static_cast<int>(py >> kFrac));
    }
    __device__ __forceinline__ int get(int lx, int ly) const {
        const int x = lx + offx, y = ly + offy;
        uint32_t b;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(saddr + static_cast<uint32_t>(y * pitch + (x >> 1))));
        return (b >> ((x & 1) << 2)) & 15;
    }
};

// One byte per cell: the march reads the full 8-bit skip code with a single LDS.U8.
struct WindowV8S {
    uint32_t base;           // shared address of window cell (offx, offy): saddr + offy * pitch + offx
    int pitch;
    // main lookup straight from the fixed-point position: shift+add, shift, multiply-add, LDS
    __device__ __forceinline__ int get_p(uint32_t px, uint32_t py) const {
        uint32_t b;
        const uint32_t t = (px >> kFrac) + base;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"((py >> kFrac) * static_cast<uint32_t>(pitch) + t));
        return static_cast<int>(b);
    }
    __device__ __forceinline__ int get(int lx, int ly) const {
        uint32_t b;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(base + static_cast<uint32_t>(ly * pitch + lx)));
        return static_cast<int>(b);
    }
};

// keeps a value in a register: the compiler may not rematerialise it inside the march loop
__device__ __forceinline__ int pin_reg(int v) {
    int r;
    asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

struct BeamDev {
    int R;
    float angle[kMaxBeams];     // downsampled_angles_ (float32)
    double cosa[kMaxBeams];     // cos((double)angle)
    double sina[kMaxBeams];
};

}  // namespace mclb200

#include "dir_kernels.cuh"
#include "exact_kernels.cuh"

namespace mclb200 {

// ------------------------------------------------------------------------------------------
// resample (:658-665) + motion model (:449-503)
// ------------------------------------------------------------------------------------------
struct MotionArgs {
    int64_t N;                // particles of this rank's slice (CDF length, array length)
    int64_t glo;              // global index of local slot 0: noise and RNG counters are keyed by the global slot
    const double* cdf;        // [F][N]
    const double* sx;         // source state [F][N] each
    const double* sy;
    const double* st;
    double* dx;               // destination state
    double* dy;
    double* dt;
    int32_t* idx_out;         // [F][N]
    const double* u;          // [F][NG] injected or nullptr (indexed by the global slot)
    const double* z;          // [F][3 NG] injected or nullptr
    // packed copy of the state, one 32-byte (x, y, theta, 0) entry per particle: the source-pose
    // gather of the resampling touches one memory sector instead of three.  spose4 = source buffer
    // (nullptr: not valid, use the SoA arrays), dpose4 = destination buffer, always written.
    const double4* spose4;
    double4* dpose4;
    // sharded filter: the source pose (and index) of every own slot, pushed here by the rank that
    // owns the source (k_route); nullptr = search and gather locally (k_resample_motion)
    const double4* routed;
    const uint32_t* where;       // [N] two-hop routing: where in `routed` the answer of each slot lies (nullptr: at the slot)
    // coarse level of the CDF search, staged in shared memory: coarse[f][k] = cdf[(k+1) << cshift) - 1]
    const double* coarse;     // [F][nc] or nullptr
    int nc, cshift;
    const double* mid;        // [F][C] chunk-end values of the CDF (middle level of the search) or nullptr
    int64_t C;                // chunks per filter (stride of mid)
    const double* action;     // [F][3] device
    double disp_x, disp_y, disp_t;
    uint64_t seed;
    const unsigned long long* update_no;   // device counter of completed updates (keys the Philox stream)
    double* centre;           // [F][2] accumulators (sum x, sum y)
    // heading histogram of the counting sort (k_sort_scatter), accumulated here instead of in a pass of its own:
    // block-private counts in shared memory, one global atomic per (block, non-empty bucket); nullptr = no sort
    int* hist;                // [F][hist_B], zeroed by k_prepare_obs
    int hist_B;
    // directional ray stage (dir_kernels.cuh): ray-start records in slot order, or nullptr;
    // k_dir_gather moves them to their heading-sorted slots
    DirRec* rec;
    MapDev map;
    int B;                    // heading buckets of the directional stage's sector arithmetic
};

struct MotionScalars {
    double dt, vel, omega, radius, dtheta, vdt;
    int straight;
};

// the scalar preamble of motion_model (:452-471, :487-488), same operations in the same order
__device__ __forceinline__ MotionScalars motion_scalars(double fwd, double ang) {
    MotionScalars m;
    m.dt = 0.01;
    m.vel = 0.0;
    m.omega = 0.0;
    if (fabs(fwd) > 0.001) {
        if (fabs(fwd) < 0.1)
            m.dt = __ddiv_rn(fabs(fwd), 1.0);
        else
            m.dt = __ddiv_rn(fabs(fwd), 5.0);
        m.dt = fmax(0.001, fmin(m.dt, 0.1));
        m.vel = __ddiv_rn(fwd, m.dt);
    }
    if (fabs(ang) > 0.001) m.omega = __ddiv_rn(ang, m.dt);
    m.straight = fabs(m.omega) < 1e-6;
    m.vdt = __dmul_rn(m.vel, m.dt);
    m.radius = m.straight ? 0.0 : __ddiv_rn(m.vel, m.omega);
    m.dtheta = __dmul_rn(m.omega, m.dt);
    return m;
}

__device__ __forceinline__ double wrap_angle_dev(double a) {  // src/utils.cpp:43-48
    const double pi = 3.14159265358979323846, two_pi = 2.0 * 3.14159265358979323846;
    while (a > pi) a = __dsub_rn(a, two_pi);
    while (a < -pi) a = __dadd_rn(a, two_pi);
    return a;
}

// Device noise streams (production mode; the parity tests inject the reference's own draws).
// Philox4x32-10 keyed by (seed, update) and counted by the GLOBAL slot, so a sharded filter draws
// what the same filter draws on one GPU.  One call yields the resampling uniforms of TWO slots
// (stream 0, counter = slot / 2) or the three motion normals of one slot (stream 1).
__device__ __forceinline__ Philox4 noise_words(uint64_t counter, int f, uint32_t stream, uint64_t seed, uint64_t update_no) {
    return philox4x32_10(static_cast<uint32_t>(counter), static_cast<uint32_t>(counter >> 32), static_cast<uint32_t>(f), stream,
                         static_cast<uint32_t>(seed) ^ static_cast<uint32_t>(update_no),
                         static_cast<uint32_t>(seed >> 32) ^ static_cast<uint32_t>(update_no >> 32) ^ 0x5bd1e995u);
}
__device__ __forceinline__ double resample_uniform(int64_t i, int f, uint64_t seed, uint64_t update_no) {
    const Philox4 r = noise_words(static_cast<uint64_t>(i) >> 1, f, 0u, seed, update_no);
    return (i & 1) ? canonical_from_words(r.v[2], r.v[3]) : canonical_from_words(r.v[0], r.v[1]);
}

// lower_bound(cp.begin(), cp.end(), u)  (random.tcc:2709-2713) in three levels.  coarse (shared memory): the
// exact CDF values at the ends of 2^cshift-element segments -- the first segment whose end value is >= u contains
// the answer.  mid (global): the value at the end of every 8-element chunk -- the segment's chunk ends are ONE
// 64-byte line when cshift == 6, so the first chunk whose end is >= u costs one memory round trip; the chunk's 8
// entries are another 64-byte line and the answer is the chunk's base + the number of entries below u.  Two
// dependent L2 accesses instead of the six of a binary search over the segment (the search is latency bound).
// 32 bytes per lane in ONE request (LDG.256, sm_100): a divergent gather costs the L1TEX one slot per lane and
// request whatever its width, so the search is counted in requests, not bytes
__device__ __forceinline__ void ldg256(const double* p, double& a, double& b, double& c, double& d) {
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ int count_below(const double* __restrict__ p, double u) {   // p: 8 doubles
    if (reinterpret_cast<uintptr_t>(p) & 31u) {   // (a batch of filters whose particle count is not a multiple of 4)
        int n = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) n += __ldg(p + i) < u;
        return n;
    }
    double v0, v1, v2, v3, v4, v5, v6, v7;
    ldg256(p, v0, v1, v2, v3);
    ldg256(p + 4, v4, v5, v6, v7);
    return (v0 < u) + (v1 < u) + (v2 < u) + (v3 < u) + (v4 < u) + (v5 < u) + (v6 < u) + (v7 < u);
}
// The coarse level is searched on 32-bit keys: for non-negative doubles the order of the values is the order of
// their bit patterns, so the high word decides unless it ties (then the exact double from global memory does).
// Random 4-byte shared-memory reads cost a third of the bank-conflict replays of 8-byte ones, and the table is
// half the size.
__device__ __forceinline__ int64_t cdf_lower_bound(const double* __restrict__ cp, const double* __restrict__ mid, int64_t N,
                                                   const uint32_t* ts, const double* __restrict__ coarse, int nc, int cshift, double u) {
    int64_t lo = 0, hi = N;
    if (nc > 0) {
        const uint32_t uh = static_cast<uint32_t>(__double2hiint(u));
        int kl = 0, kh = nc;
        while (kl < kh) {
            const int km = (kl + kh) >> 1;
            const uint32_t t = ts[coarse_slot(km)];
            bool below = t < uh;
            if (t == uh) below = __ldg(coarse + km) < u;   // rare: same 32 leading bits
            if (below)
                kl = km + 1;
            else
                kh = km;
        }
        lo = static_cast<int64_t>(kl) << cshift;
        hi = min(N, lo + (int64_t{1} << cshift));
    }
    if (mid != nullptr && hi > lo) {
        // first chunk of [lo, hi) whose end value is >= u (the last chunk's end is cp[hi - 1] or the filter's last entry)
        int64_t cl = lo >> 3, ch = (hi + 7) >> 3;
        if (ch - cl == 8 && (cl & 7) == 0) {
            cl += count_below(mid + cl, u);
            if (cl >= ch) cl = ch - 1;   // (cannot happen for u <= the segment's end value)
        } else {
            while (cl < ch - 1) {        // generic: bisect the chunk ends
                const int64_t cm = (cl + ch - 1) >> 1;
                if (__ldg(mid + cm) < u)
                    cl = cm + 1;
                else
                    ch = cm + 1;
            }
        }
        const int64_t base = cl << 3;
        if (base + 8 <= N) {
            lo = base + count_below(cp + base, u);
        } else {
            lo = base;
            while (lo < N && __ldg(cp + lo) < u) ++lo;
        }
        if (lo >= N) lo = N - 1;
        return lo;
    }
    while (lo < hi) {
        const int64_t mid_i = (lo + hi) >> 1;
        if (__ldg(cp + mid_i) < u)
            lo = mid_i + 1;
        else
            hi = mid_i;
    }
    if (lo >= N) lo = N - 1;  // unreachable for u < 1 == cp[N-1]; keeps reads in range
    return lo;
}

// motion_model for one particle (:474-502): kinematics + noise + wrap.  zp: the slot's three injected normals or
// nullptr (device Philox keyed by the GLOBAL slot i, so the result does not depend on which GPU evaluates it)
struct MotionNoise {
    double disp_x, disp_y, disp_t;
    uint64_t seed;
};
__device__ __forceinline__ void motion_apply(const MotionScalars& m, const MotionNoise& q, const double* zp, uint64_t update_no, int f,
                                             int64_t i, double x, double y, double th, double* ox, double* oy, double* ot) {
    double z0, z1, z2;
    if (zp) {
        z0 = zp[0];
        z1 = zp[1];
        z2 = zp[2];
    } else {
        const Philox4 r = noise_words(static_cast<uint64_t>(i), f, 1u, q.seed, update_no);
        double t;
        normal_pair(r.v[0], r.v[1], &z0, &z1);
        normal_pair(r.v[2], r.v[3], &z2, &t);
    }
    double nx, ny, nt;
    if (m.straight) {
        double s, c;
        sincos(th, &s, &c);
        nx = __dadd_rn(x, __dmul_rn(m.vdt, c));
        ny = __dadd_rn(y, __dmul_rn(m.vdt, s));
        nt = th;
    } else {
        double s0, c0, s1, c1;
        sincos(th, &s0, &c0);
        sincos(__dadd_rn(th, m.dtheta), &s1, &c1);
        nx = __dadd_rn(x, __dmul_rn(m.radius, __dsub_rn(s1, s0)));
        ny = __dsub_rn(y, __dmul_rn(m.radius, __dsub_rn(c1, c0)));
        nt = __dadd_rn(th, m.dtheta);
    }
    nx = __dadd_rn(nx, __dmul_rn(z0, q.disp_x));
    ny = __dadd_rn(ny, __dmul_rn(z1, q.disp_y));
    nt = __dadd_rn(nt, __dmul_rn(z2, q.disp_t));
    *ox = nx;
    *oy = ny;
    *ot = wrap_angle_dev(nt);
}

// the stores of the proposal: SoA state, packed copy, ray-start record, heading count; adds the (finite) position to
// the cloud-centre sums.  moved: (x, y, th) is already the moved pose (two-hop routing: the serving rank applied
// the motion model), otherwise the source pose.
__device__ __forceinline__ void motion_store(const MotionArgs& a, const MotionScalars& m, uint64_t update_no, int f, int64_t li,
                                             double x, double y, double th, bool moved, double* sum_x, double* sum_y, int* sort_cnt) {
    const int64_t fo = static_cast<int64_t>(f) * a.N;
    const int64_t i = a.glo + li;
    double nx = x, ny = y, nt = th;
    if (!moved)
        motion_apply(m, MotionNoise{a.disp_x, a.disp_y, a.disp_t, a.seed}, a.z ? a.z + 3 * (fo + i) : nullptr, update_no, f, i, x, y, th,
                     &nx, &ny, &nt);
    a.dx[fo + li] = nx;
    a.dy[fo + li] = ny;
    a.dt[fo + li] = nt;
    {
        double2* d4 = reinterpret_cast<double2*>(a.dpose4 + fo + li);
        d4[0] = make_double2(nx, ny);
        d4[1] = make_double2(nt, 0.0);
    }
    int bucket = -1;
    if (a.rec) {
        bucket = theta_bucket(nt, a.B);
        dir_write_record(a.map, a.rec, fo + li, nx, ny, nt, bucket);
    }
    if (sort_cnt) atomicAdd(&sort_cnt[(bucket >= 0 && a.hist_B == a.B) ? bucket : theta_bucket(nt, a.hist_B)], 1);
    if (!(fabs(nx) < 1e12) || !(fabs(ny) < 1e12)) nx = ny = 0.0;  // keep the window centre finite
    *sum_x += nx;
    *sum_y += ny;
}

constexpr int kMotionThreads = 1024;
constexpr int kWhereShift = 28;     // two-hop routing: slots per rank < 2^28, ranks <= 16 (MotionArgs::where)

// One thread per output slot, persistent blocks when the coarse CDF level is large (one staging of
// the table per SM).  a.routed == nullptr: search the CDF and gather the source pose here (whole
// filter on this GPU).  a.routed != nullptr (sharded filter): the source pose of every slot was
// pushed into `routed` by k_route on the rank that owns the source; only the motion runs here.
__global__ void __launch_bounds__(kMotionThreads) k_resample_motion(MotionArgs a) {
    pdl_enter();
    __shared__ double sm[kMotionThreads / 32];
    __shared__ int sort_cnt[kMaxBuckets];
    extern __shared__ uint32_t ts[];   // a.nc keys (high words of the coarse level) when it is used
    const int f = blockIdx.y;
    if (a.hist) {
        for (int b = threadIdx.x; b < a.hist_B; b += kMotionThreads) sort_cnt[b] = 0;
        __syncthreads();
    }
    const bool search = a.routed == nullptr;
    const int nc = (search && a.coarse != nullptr) ? a.nc : 0;
    const double* coarse = nc > 0 ? a.coarse + static_cast<int64_t>(f) * a.nc : nullptr;
    if (nc > 0) {
        for (int t = threadIdx.x; t < nc; t += kMotionThreads) ts[coarse_slot(t)] = static_cast<uint32_t>(__double2hiint(coarse[t]));
        __syncthreads();
    }
    const int64_t N = a.N;
    const int64_t fo = static_cast<int64_t>(f) * N;
    const MotionScalars m = motion_scalars(a.action[3 * f + 0], a.action[3 * f + 2]);
    const uint64_t update_no = *a.update_no;
    double sum_x = 0.0, sum_y = 0.0;
    for (int64_t li = static_cast<int64_t>(blockIdx.x) * kMotionThreads + threadIdx.x; li < N;
         li += static_cast<int64_t>(gridDim.x) * kMotionThreads) {
        double x, y, th;
        if (search) {
            const int64_t i = a.glo + li;
            const double u = a.u ? a.u[fo + i] : resample_uniform(i, f, a.seed, update_no);
            int64_t lo = cdf_lower_bound(a.cdf + fo, a.mid ? a.mid + static_cast<int64_t>(f) * a.C : nullptr, N, ts, coarse, nc, a.cshift, u);
            if (N < 2) lo = 0;        // libstdc++ clears the table for fewer than 2 weights
            a.idx_out[fo + li] = static_cast<int32_t>(lo);
            if (a.spose4) {
                double pad;
                ldg256(reinterpret_cast<const double*>(a.spose4 + fo + lo), x, y, th, pad);
            } else {
                x = a.sx[fo + lo];
                y = a.sy[fo + lo];
                th = a.st[fo + lo];
            }
        } else {
            // written by other GPUs over NVLink during k_route: system-scope loads (never a line of this SM's L1)
            int64_t at = li;
            if (a.where) {   // two-hop routing: the answer lies at [server][position of the request]
                const uint32_t w = a.where[li];
                at = static_cast<int64_t>(w >> kWhereShift) * N + (w & ((1u << kWhereShift) - 1u));
            }
            // one 32-byte request per slot (the answers are gathered in request order, i.e. at random)
            double w3;
            asm volatile("ld.relaxed.sys.global.v4.f64 {%0, %1, %2, %3}, [%4];"
                         : "=d"(x), "=d"(y), "=d"(th), "=d"(w3)
                         : "l"(a.routed + at)
                         : "memory");
            a.idx_out[li] = static_cast<int32_t>(__double_as_longlong(w3));
        }
        motion_store(a, m, update_no, f, li, x, y, th, !search && a.where != nullptr, &sum_x, &sum_y, a.hist ? sort_cnt : nullptr);
    }
    // cloud centre for the shared-memory window of the ray kernel
    const double bx = block_sum<kMotionThreads>(sum_x, sm);
    const double by = block_sum<kMotionThreads>(sum_y, sm);
    if (threadIdx.x == 0) {
        atomicAdd(a.centre + 2 * f + 0, bx);
        atomicAdd(a.centre + 2 * f + 1, by);
    }
    if (a.hist) {   // (block_sum ended with a barrier: every thread's count is in)
        int* hist = a.hist + static_cast<int64_t>(f) * a.hist_B;
        for (int b = threadIdx.x; b < a.hist_B; b += kMotionThreads) {
            const int c = sort_cnt[b];
            if (c) atomicAdd(hist + b, c);
        }
    }
}

// ------------------------------------------------------------------------------------------
// sharded filter: sender-driven resampling
// ------------------------------------------------------------------------------------------
// Every rank owns a slice of the global CDF.  The draw u_i of EVERY global slot i is a pure function
// of (seed, update, i), so each rank evaluates all of them, keeps the draws that fall into its own
// CDF range (rank_end[rank-1], rank_end[rank]] -- exactly one rank claims each slot --, finds the
// source particle in its local CDF slice and PUSHES the source pose (x, y, theta, global index) over
// NVLink into slot i of the owner's `routed` array.  Nothing is ever read from a peer.  A warp tests
// 64 slots per round (one Philox call per lane) and queues the claimed ones in shared memory, so the
// searches run with full warps although only 1 / world of the slots are claimed.
struct RouteArgs {
    int64_t NG;               // global slots
    int64_t N;                // particles of this rank (== slots per rank)
    const double* cdf;        // [N] this rank's slice of the global CDF
    const double* coarse;     // [nc] coarse level of the local slice
    int nc, cshift;
    const double* mid;        // [C] chunk-end values of the local slice
    const double* rank_end;   // [world] exact CDF value at the end of every rank's slice
    const double4* spose4;    // [N] local source state, packed (nullptr: use the SoA arrays)
    const double* sx;
    const double* sy;
    const double* st;
    double4* routed[kMaxWorld];   // every rank's `routed` array (own included)
    const double* u;          // [NG] injected uniforms or nullptr
    uint64_t seed;
    const unsigned long long* update_no;
    unsigned int* done;
    unsigned long long* dbg;    // diagnostics (nullable): [8] slowest CTA's scan+serve cycles, [9] incl. the system fence, [10] last CTA's wait for the peers
    ShardDev sh;
};
constexpr int kRouteThreads = 1024;
constexpr int kRouteQueue = 128;   // claimed slots a warp can hold (<= 31 left over + 64 new)

__device__ __forceinline__ void route_finish(const ShardDev& sh, unsigned long long epoch) {
    if (threadIdx.x == 0) *sh.xseq = epoch;
}

__global__ void __launch_bounds__(kRouteThreads) k_route(RouteArgs a) {
    pdl_enter();
    extern __shared__ __align__(8) uint32_t ts[];   // nc keys of the coarse level (padded to even), then the warps' queues
    __shared__ bool is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long c_begin = clock64();
    for (int t = tid; t < a.nc; t += kRouteThreads) ts[coarse_slot(t)] = static_cast<uint32_t>(__double2hiint(a.coarse[t]));
    double* qbase = reinterpret_cast<double*>(ts + ((coarse_slots(a.nc) + 1) & ~1));
    double* qu = qbase + warp * kRouteQueue;                                               // queued draws
    int* qi = reinterpret_cast<int*>(qbase + (kRouteThreads / 32) * kRouteQueue) + warp * kRouteQueue;   // queued slots
    __syncthreads();
    const int me = a.sh.rank, world = a.sh.world;
    const double lo_u = me > 0 ? a.rank_end[me - 1] : -1.0;          // claim u in (lo_u, hi_u]
    const double hi_u = me + 1 < world ? a.rank_end[me] : 2.0;
    const uint64_t update_no = *a.update_no;
    const int64_t glo = static_cast<int64_t>(me) * a.N;
    int qn = 0;   // entries in this warp's queue (warp-uniform)

    auto serve = [&](int k) {   // lane k < count serves queue entry k
        const double u = qu[k];
        const int64_t i = qi[k];
        const int64_t j = cdf_lower_bound(a.cdf, a.mid, a.N, ts, a.coarse, a.nc, a.cshift, u);
        double x, y, th;
        if (a.spose4) {
            double pad;
            ldg256(reinterpret_cast<const double*>(a.spose4 + j), x, y, th, pad);
        } else {
            x = a.sx[j];
            y = a.sy[j];
            th = a.st[j];
        }
        const uint32_t owner = static_cast<uint32_t>(i) / static_cast<uint32_t>(a.N);   // (slots and slice sizes are below 2^31)
        double* dst = reinterpret_cast<double*>(a.routed[owner] + (static_cast<uint32_t>(i) - owner * static_cast<uint32_t>(a.N)));
        // one 32-byte store: one NVLink write transaction per routed particle
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(x), "d"(y), "d"(th),
                     "d"(__longlong_as_double(glo + j))
                     : "memory");
    };

    const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (kRouteThreads / 32);
    const int64_t gw = static_cast<int64_t>(blockIdx.x) * (kRouteThreads / 32) + warp;
    for (int64_t base = gw * 64; base < a.NG; base += warps_total * 64) {
        const int64_t i0 = base + 2 * lane;   // this lane's two slots: one Philox call
        double u0 = 2.0, u1 = 2.0;
        if (i0 < a.NG) {
            if (a.u) {
                u0 = a.u[i0];
                if (i0 + 1 < a.NG) u1 = a.u[i0 + 1];
            } else {
                const Philox4 r = noise_words(static_cast<uint64_t>(i0) >> 1, 0, 0u, a.seed, update_no);
                u0 = canonical_from_words(r.v[0], r.v[1]);
                if (i0 + 1 < a.NG) u1 = canonical_from_words(r.v[2], r.v[3]);
            }
        }
        const bool c0 = u0 > lo_u && u0 <= hi_u && u0 < 1.5, c1 = u1 > lo_u && u1 <= hi_u && u1 < 1.5;
        const unsigned b0 = __ballot_sync(kFullMask, c0), b1 = __ballot_sync(kFullMask, c1);
        const unsigned below = (1u << lane) - 1u;
        if (c0) {
            const int p = qn + __popc(b0 & below);
            qu[p] = u0;
            qi[p] = static_cast<int>(i0);
        }
        if (c1) {
            const int p = qn + __popc(b0) + __popc(b1 & below);
            qu[p] = u1;
            qi[p] = static_cast<int>(i0 + 1);
        }
        qn += __popc(b0) + __popc(b1);
        __syncwarp();
        while (qn >= 32) {   // full warps of claimed slots, taken from the tail of the queue
            qn -= 32;
            serve(qn + lane);
            __syncwarp();
        }
    }
    if (lane < qn) serve(lane);
    const long long c_loop = clock64();
    // every store of this CTA is performed at system scope before the CTA counts itself done
    // ONE system-scope fence per CTA: the barrier orders every thread's stores before thread 0's fence, which is
    // cumulative over them (a fence by each of the 1024 threads cost 8-19 us per kernel)
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        if (a.dbg) {
            atomicMax(a.dbg + 8, static_cast<unsigned long long>(c_loop - c_begin));
            atomicMax(a.dbg + 9, static_cast<unsigned long long>(clock64() - c_begin));
        }
        is_last = atomicAdd(a.done, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (tid == 0) *a.done = 0;
    __syncthreads();
    const unsigned long long epoch = *a.sh.xseq + 1ull;
    const long long c_pub = clock64();
    shard_publish(a.sh, epoch, nullptr, 0);
    if (!a.sh.fused) return;
    if (!shard_wait(a.sh, epoch)) return;
    if (a.dbg && tid == 0) a.dbg[10] = static_cast<unsigned long long>(clock64() - c_pub);
    route_finish(a.sh, epoch);
}

// host-ordered ranks: checks that every rank's k_route has published
__global__ void __launch_bounds__(32) k_route_check(ShardDev sh) {
    pdl_enter();
    const unsigned long long epoch = *sh.xseq + 1ull;
    if (!shard_wait(sh, epoch)) return;
    route_finish(sh, epoch);
}

// ------------------------------------------------------------------------------------------
// sharded filter: sender-driven resampling in two hops (the default)
// ------------------------------------------------------------------------------------------
// k_route makes every rank evaluate ALL world * N draws to find the 1 / world it has to serve: work per rank
// grows with the number of ranks.  Here the owner of a SLOT evaluates its own N draws, finds the rank whose
// CDF range holds each draw (world - 1 comparisons against the replicated rank_end table) and appends the
// slot's local index to that rank's inbox (k_route_request; 4 bytes per slot over NVLink, written as
// contiguous runs).  After one exchange of the per-destination counts the source ranks serve their inboxes
// with full warps (k_route_serve): draw recomputed from the global slot, search in the local CDF slice, 32-byte
// push of the source pose into the slot owner's `routed` array -- at position [server][k], k being the request's
// position in the inbox, so that the pushes of a warp are CONSECUTIVE 32-byte records (whole NVLink packets
// instead of one packet per particle); the owner noted (server, k) of each of its slots in `where` when it
// made the request and gathers the record from its own memory.  Per rank: N classifications + (on balanced
// weights) N searches, independent of the world size.  Results are identical to k_route's by construction:
// the same draw, the same claim rule (rank_end[q-1] < u <= rank_end[q]), the same local lower_bound.
struct RouteReqArgs {
    int64_t N;                      // slots of this rank
    const double* rank_end;         // [world]
    uint32_t* inbox[kMaxWorld];     // every rank's inbox [world senders][N] (own included)
    unsigned int* req_count;        // [kMaxWorld] own: requests appended per destination (zeroed by the serve kernel)
    uint32_t* where;                // [N] own: server << kWhereShift | position of the slot's request (and of its answer)
    const double* u;                // [NG] injected uniforms or nullptr
    uint64_t seed;
    const unsigned long long* update_no;
    unsigned int* done;
    unsigned long long* dbg;        // diagnostics (nullable): [8] slowest CTA until its stores are issued, [9] incl. the system fence, [10] last CTA's wait
    ShardDev sh;
};
constexpr int kReqThreads = 1024;
constexpr int kReqPer = 8;          // slots per thread (four Philox calls)
constexpr int kReqBlock = kReqThreads * kReqPer;

__global__ void __launch_bounds__(kReqThreads) k_route_request(RouteReqArgs a) {
    pdl_enter();
    __shared__ uint32_t stage[kReqBlock];           // the block's slots grouped by destination
    __shared__ unsigned int s_cnt[kMaxWorld], s_off[kMaxWorld + 1], s_base[kMaxWorld];
    __shared__ double s_end[kMaxWorld];
    __shared__ unsigned long long s_pay[kMaxWorld];
    __shared__ bool is_last;
    const int tid = threadIdx.x, lane = tid & 31;
    const int me = a.sh.rank, world = a.sh.world;
    const long long c_begin = clock64();
    if (tid < kMaxWorld) {
        s_cnt[tid] = 0;
        s_end[tid] = tid < world - 1 ? a.rank_end[tid] : 2.0;
    }
    __syncthreads();
    const uint64_t update_no = *a.update_no;
    const int64_t glo = static_cast<int64_t>(me) * a.N;
    const int64_t b0 = static_cast<int64_t>(blockIdx.x) * kReqBlock;
    // pass 1: destination of every slot and its position among the block's slots for that destination
    int dest[kReqPer];
    unsigned int posn[kReqPer];
#pragma unroll
    for (int e = 0; e < kReqPer; e += 2) {
        // a thread's pair of slots shares one Philox call; consecutive threads take consecutive pairs
        const int64_t li = b0 + static_cast<int64_t>(e / 2) * (2 * kReqThreads) + 2 * tid;
        double u0 = 3.0, u1 = 3.0;
        if (li < a.N) {
            const int64_t i = glo + li;          // glo and li are even or N is odd: pair by GLOBAL slot parity below
            if (a.u) {
                u0 = a.u[i];
                if (li + 1 < a.N) u1 = a.u[i + 1];
            } else if ((i & 1) == 0) {
                const Philox4 r = noise_words(static_cast<uint64_t>(i) >> 1, 0, 0u, a.seed, update_no);
                u0 = canonical_from_words(r.v[0], r.v[1]);
                if (li + 1 < a.N) u1 = canonical_from_words(r.v[2], r.v[3]);
            } else {
                u0 = resample_uniform(i, 0, a.seed, update_no);
                if (li + 1 < a.N) u1 = resample_uniform(i + 1, 0, a.seed, update_no);
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double u = h ? u1 : u0;
            int q = -1;
            if (u < 2.5) {
                q = 0;
                for (int k = 0; k < world - 1; ++k) q += s_end[k] < u;   // the rank with rank_end[q-1] < u <= rank_end[q]
            }
            dest[e + h] = q;
            // warp-aggregated count: one shared-memory atomic per (warp, destination)
            const unsigned same = __match_any_sync(kFullMask, q);
            const int leader = __ffs(same) - 1;
            unsigned int base = 0;
            if (lane == leader && q >= 0) base = atomicAdd(&s_cnt[q], static_cast<unsigned int>(__popc(same)));
            base = __shfl_sync(kFullMask, base, leader);
            posn[e + h] = base + static_cast<unsigned int>(__popc(same & ((1u << lane) - 1u)));
        }
    }
    __syncthreads();
    if (tid == 0) {
        unsigned int o = 0;
        for (int q = 0; q < world; ++q) {
            s_off[q] = o;
            o += s_cnt[q];
        }
        s_off[world] = o;
    }
    if (tid < world) s_base[tid] = s_cnt[tid] ? atomicAdd(a.req_count + tid, s_cnt[tid]) : 0u;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < kReqPer; ++e) {
        if (dest[e] >= 0) {
            const int64_t li = b0 + static_cast<int64_t>(e / 2) * (2 * kReqThreads) + 2 * tid + (e & 1);
            stage[s_off[dest[e]] + posn[e]] = static_cast<uint32_t>(li);
            a.where[li] = (static_cast<uint32_t>(dest[e]) << kWhereShift) | (s_base[dest[e]] + posn[e]);
        }
    }
    __syncthreads();
    // pass 2: every destination's run goes out as consecutive 4-byte stores (whole sectors over NVLink)
    for (int q = 0; q < world; ++q) {
        uint32_t* dst = a.inbox[q] + static_cast<size_t>(me) * static_cast<size_t>(a.N) + s_base[q];
        const unsigned int n = s_cnt[q], o = s_off[q];
        for (unsigned int t = tid; t < n; t += kReqThreads) dst[t] = stage[o + t];
    }
    const long long c_loop = clock64();
    // ONE system-scope fence per CTA: the barrier orders every thread's stores before thread 0's fence, which is
    // cumulative over them (a fence by each of the 1024 threads cost 8-19 us per kernel)
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        if (a.dbg) {
            atomicMax(a.dbg + 8, static_cast<unsigned long long>(c_loop - c_begin));
            atomicMax(a.dbg + 9, static_cast<unsigned long long>(clock64() - c_begin));
        }
        is_last = atomicAdd(a.done, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (tid == 0) *a.done = 0;
    // payload: this rank's request count for every destination
    if (tid < kMaxWorld) s_pay[tid] = tid < world ? static_cast<unsigned long long>(ld_sys_u32(a.req_count + tid)) : 0ull;
    __syncthreads();
    const unsigned long long epoch = *a.sh.xseq + 1ull;
    const long long c_pub = clock64();
    shard_publish(a.sh, epoch, s_pay, world);
    if (!a.sh.fused) return;
    if (!shard_wait(a.sh, epoch)) return;
    if (a.dbg && tid == 0) a.dbg[10] = static_cast<unsigned long long>(clock64() - c_pub);
    route_finish(a.sh, epoch);
}

struct RouteServeArgs {
    int64_t N;
    const double* cdf;
    const double* coarse;
    int nc, cshift;
    const double* mid;
    const double4* spose4;
    const double* sx;
    const double* sy;
    const double* st;
    const uint32_t* inbox;          // own inbox [world][N]
    double4* routed[kMaxWorld];     // every rank's answer array [world servers][N]
    // the serving rank also applies the motion model (the noise is keyed by the global slot, so any GPU computes the same
    // bits): the search is L1TEX-bound with idle issue slots, and the owner's kernel becomes a pure scatter of the answers
    const double* action;           // [3] device
    const double* z;                // [3 NG] injected normals or nullptr
    MotionNoise noise;
    unsigned int* req_count;        // own request counters: cleared here for the next update
    const double* u;
    uint64_t seed;
    const unsigned long long* update_no;
    unsigned int* done;
    unsigned long long* dbg;        // diagnostics (nullable): as RouteReqArgs
    ShardDev sh;
};

__global__ void __launch_bounds__(kRouteThreads, 1) k_route_serve(RouteServeArgs a) {
    pdl_enter();
    extern __shared__ __align__(8) uint32_t ts[];   // nc keys of the coarse level
    __shared__ unsigned int s_pre[kMaxWorld + 1];
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    const int me = a.sh.rank, world = a.sh.world;
    const long long c_begin = clock64();
    for (int t = tid; t < a.nc; t += kRouteThreads) ts[coarse_slot(t)] = static_cast<uint32_t>(__double2hiint(a.coarse[t]));
    // the counts travelled in the exchange the request kernel (or its check kernel) completed: epoch == *xseq
    const unsigned long long req_epoch = *a.sh.xseq;
    if (tid == 0) {
        unsigned int o = 0;
        for (int r = 0; r < world; ++r) {
            s_pre[r] = o;
            o += static_cast<unsigned int>(ld_sys_u64(mbox_slot(a.sh, me, req_epoch, r) + me));
        }
        s_pre[world] = o;
    }
    if (blockIdx.x == 0 && tid < kMaxWorld) a.req_count[tid] = 0;   // (own counters: only this rank's request kernel adds to them)
    __syncthreads();
    const unsigned int total = s_pre[world];
    const uint64_t update_no = *a.update_no;
    const int64_t glo = static_cast<int64_t>(me) * a.N;
    const MotionScalars m = motion_scalars(a.action[0], a.action[2]);
    for (unsigned int t = blockIdx.x * kRouteThreads + tid; t < total; t += gridDim.x * kRouteThreads) {
        int r = 0;
        while (r + 1 < world && s_pre[r + 1] <= t) ++r;
        const uint32_t li = ld_sys_u32(a.inbox + static_cast<size_t>(r) * static_cast<size_t>(a.N) + (t - s_pre[r]));
        const int64_t i = static_cast<int64_t>(r) * a.N + li;
        const double u = a.u ? a.u[i] : resample_uniform(i, 0, a.seed, update_no);
        const int64_t j = cdf_lower_bound(a.cdf, a.mid, a.N, ts, a.coarse, a.nc, a.cshift, u);
        double x, y, th;
        if (a.spose4) {
            double pad;
            ldg256(reinterpret_cast<const double*>(a.spose4 + j), x, y, th, pad);
        } else {
            x = a.sx[j];
            y = a.sy[j];
            th = a.st[j];
        }
        motion_apply(m, a.noise, a.z ? a.z + 3 * i : nullptr, update_no, 0, i, x, y, th, &x, &y, &th);
        // answer k of this rank for rank r: consecutive lanes write consecutive records
        double* dst = reinterpret_cast<double*>(a.routed[r] + static_cast<size_t>(me) * static_cast<size_t>(a.N) + (t - s_pre[r]));
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(x), "d"(y), "d"(th),
                     "d"(__longlong_as_double(glo + j))
                     : "memory");
    }
    const long long c_loop = clock64();
    // ONE system-scope fence per CTA: the barrier orders every thread's stores before thread 0's fence, which is
    // cumulative over them (a fence by each of the 1024 threads cost 8-19 us per kernel)
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        if (a.dbg) {
            atomicMax(a.dbg + 8, static_cast<unsigned long long>(c_loop - c_begin));
            atomicMax(a.dbg + 9, static_cast<unsigned long long>(clock64() - c_begin));
        }
        is_last = atomicAdd(a.done, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (tid == 0) *a.done = 0;
    __syncthreads();
    const unsigned long long epoch = req_epoch + 1ull;
    const long long c_pub = clock64();
    shard_publish(a.sh, epoch, nullptr, 0);
    if (!a.sh.fused) return;
    if (!shard_wait(a.sh, epoch)) return;
    if (a.dbg && tid == 0) a.dbg[10] = static_cast<unsigned long long>(clock64() - c_pub);
    route_finish(a.sh, epoch);
}

// Counting sort of the particles by heading bucket, in two kernels of fat blocks so that the
// only global atomics are one per (block, non-empty bucket):
//   histogram       block-private shared-memory counts inside k_resample_motion, flushed once per block
//   k_sort_scatter  bucket bases (scan of the global histogram), a block-private count to
//                   reserve the block's range in each bucket, then shared-memory ranks
// perm[f][pos] = particle index.  The order inside a bucket is arbitrary; it only decides
// which warp marches which particle, never a result.
struct SortArgs {
    int64_t N;            // particles per filter (array stride)
    int64_t lo, cnt;      // the slots [lo, lo+cnt) being sorted; perm holds indices relative to lo
    const double* pt;     // [F][N]
    int* hist;            // [F][B] filled by k_resample_motion
    int* cursor;          // [F][B] zeroed
    int32_t* perm;        // [F][N]
    int B;
    int64_t chunk;        // particles per block (multiple of kSortThreads)
};
constexpr int kSortThreads = 1024;

__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(SortArgs a) {
    pdl_enter();
    __shared__ int base[kMaxBuckets];
    __shared__ int cnt[kMaxBuckets];
    __shared__ int wsum[kSortThreads / 32];
    const int f = blockIdx.y, tid = threadIdx.x;
    const int* hist = a.hist + static_cast<int64_t>(f) * a.B;
    // exclusive scan of the global histogram, kMaxBuckets / kSortThreads entries per thread
    constexpr int kPer = kMaxBuckets / kSortThreads;
    int h[kPer];
    int loc = 0;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int b = tid * kPer + q;
        h[q] = b < a.B ? hist[b] : 0;
        loc += h[q];
        cnt[b] = 0;
    }
    int v = loc;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(kFullMask, v, d);
        if (lane >= d) v += o;
    }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int w = wsum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(kFullMask, w, d);
            if (lane >= d) w += o;
        }
        wsum[lane] = w;
    }
    __syncthreads();
    int excl = v - loc + (warp ? wsum[warp - 1] : 0);
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        base[tid * kPer + q] = excl;
        excl += h[q];
    }
    __syncthreads();
    const int64_t fo = static_cast<int64_t>(f) * a.N + a.lo;
    const int64_t lo = static_cast<int64_t>(blockIdx.x) * a.chunk;
    const int64_t hi = min(a.cnt, lo + a.chunk);
    // this block's population of every bucket
    for (int64_t i = lo + tid; i < hi; i += kSortThreads) atomicAdd(&cnt[theta_bucket(a.pt[fo + i], a.B)], 1);
    __syncthreads();
    // reserve the block's range inside each bucket; cnt[] becomes the running cursor
    int* cursor = a.cursor + static_cast<int64_t>(f) * a.B;
    for (int b = tid; b < a.B; b += kSortThreads) {
        const int c = cnt[b];
        if (c) base[b] += atomicAdd(cursor + b, c);
        cnt[b] = 0;
    }
    __syncthreads();
    for (int64_t i = lo + tid; i < hi; i += kSortThreads) {
        const int b = theta_bucket(a.pt[fo + i], a.B);
        const int pos = base[b] + atomicAdd(&cnt[b], 1);
        a.perm[fo + pos] = static_cast<int32_t>(i);
    }
}

// ------------------------------------------------------------------------------------------
// sensor model (:506-583): observation -> table rows, ray march, weight product
// ------------------------------------------------------------------------------------------
struct ObsArgs {
    const float* obs;          // [F][R]
    const double* tabT;        // [(M+1)][(M+1)] transposed table: tabT[obs_idx*(M+1) + range_idx]
    const int32_t* step2idx;   // [M+1] range_idx the reference derives from step r (M = no hit)
    double* slice;             // [F][R][M+1]  slice[j][r] = table(obs_idx_j, range_idx(r))
    int R, M;
    double res;
    double* centre;            // [F][2] cleared here for k_resample_motion's sums
    int* hist;                 // [nhist] heading histogram + scatter cursors, cleared here (nullable)
    int nhist;
};

__global__ void k_prepare_obs(ObsArgs a) {
    pdl_enter();
    const int j = blockIdx.x, f = blockIdx.y;
    {   // per-update accumulators: cleared by the whole grid, ahead of the kernels that add to them
        const int nthreads = gridDim.x * gridDim.y * blockDim.x;
        const int gtid = (blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
        if (a.hist)
            for (int i = gtid; i < a.nhist; i += nthreads) a.hist[i] = 0;
        if (gtid < 2 * static_cast<int>(gridDim.y)) a.centre[gtid] = 0.0;
    }
    // obs_px = obs / res, clamp, round (:549-554, :570, :573)
    float opx = static_cast<float>(static_cast<double>(a.obs[f * a.R + j]) / a.res);
    if (opx > static_cast<float>(a.M)) opx = static_cast<float>(a.M);
    int oi = isnan(opx) ? 0 : static_cast<int>(roundf(opx));  // x86 cvttss2si(NaN) = INT_MIN -> clamps to 0
    oi = max(0, min(oi, a.M));
    const double* row = a.tabT + static_cast<int64_t>(oi) * (a.M + 1);
    double* dst = a.slice + (static_cast<int64_t>(f) * a.R + j) * (a.M + 1);
    for (int r = threadIdx.x; r <= a.M; r += blockDim.x) dst[r] = row[a.step2idx[r]];
}

struct RayArgs {
    MapDev map;
    BeamDev beams;
    int64_t N;
    const double* px;          // [F][N] proposal particles
    const double* py;
    const double* pt;
    int64_t lo, cnt;           // slots [lo, lo+cnt) computed by this launch (shard)
    const int32_t* perm;       // [F][N] processing order (heading-sorted, shard-local) or nullptr
    const double* slice;       // [F][R][M+1]
    double* w_raw;             // [F][N]
    uint8_t* steps;            // [F][N*R] or nullptr
    const double* centre;      // [F][2] sums of x and y over the filter's particles
    double inv_squash;
    int64_t* replay_count;     // diagnostics (nullable)
    const int* plan;           // directional stage's plan (nullable): mode 1 = that stage does this update
};

constexpr int kRayThreads = 1024;

// One lane = one particle; the lane walks its R beams in order and folds the table
// entries into the weight product in the reference's multiplication order (:564-579).
// Lanes of a warp hold heading-neighbours (perm), so their rays and trip counts agree.
// MC > 0: MAX_RANGE_PX known at compile time (the loop bound becomes an immediate); 0: run time
template <int WBITS, int MC>
__global__ void __launch_bounds__(kRayThreads, 1) k_raycast_weight(RayArgs a) {
    pdl_enter();
    extern __shared__ __align__(16) uint8_t smem_win[];
    if (a.plan && a.plan[0] == 1) return;   // k_raycast_dir + k_weight_steps take this update (dir_kernels.cuh)
    const int f = blockIdx.y;
    const MapDev& mp = a.map;
    const int64_t N = a.N;
    const int64_t fo = static_cast<int64_t>(f) * N;
    const int M = MC > 0 ? MC : pin_reg(mp.M);
    const int R = a.beams.R;

    // ---- stage the window of the skip map around the cloud centre --------------------------
    int wx0 = 0, wy0 = 0, vx0 = 0, vx1 = 0, vy0 = 0, vy1 = 0;
    const int pitch = WBITS == 8 ? mp.ww : (mp.ww >> 1);
    if (mp.ww > 0) {
        const double mx = a.centre[2 * f + 0] / static_cast<double>(a.cnt);
        const double my = a.centre[2 * f + 1] / static_cast<double>(a.cnt);
        const double qx = (mx - mp.ox) / mp.res + kPadL, qy = (my - mp.oy) / mp.res + kPadL;
        int cx = (qx > -1e9 && qx < 1e9) ? static_cast<int>(floor(qx)) : 0;
        int cy = (qy > -1e9 && qy < 1e9) ? static_cast<int>(floor(qy)) : 0;
        wx0 = cx - mp.ww / 2;
        wy0 = cy - mp.wh / 2;
        wx0 = max(0, min(wx0, mp.PW - mp.ww)) & ~31;
        wy0 = max(0, min(wy0, mp.PH - mp.wh));
        // particles whose rays provably stay inside the window
        vx0 = (wx0 == 0) ? 2 : wx0 + M + 2;
        vx1 = (wx0 + mp.ww >= mp.PW) ? mp.PW - 2 : wx0 + mp.ww - M - 2;
        vy0 = (wy0 == 0) ? 2 : wy0 + M + 2;
        vy1 = (wy0 + mp.wh >= mp.PH) ? mp.PH - 2 : wy0 + mp.wh - M - 2;
        // Stage the window with the bulk-copy engine (TMA, cp.async.bulk): one global->shared
        // copy per window row, completion counted in bytes on an mbarrier.  Rows are 16-byte
        // aligned on both sides (PW, ww, wx0 are multiples of 32 cells).
        const int gpitch = WBITS == 8 ? mp.PW : (mp.PW >> 1);
        const uint8_t* gsrc = WBITS == 8 ? mp.v8 + wx0 : mp.v4 + (wx0 >> 1);
        __shared__ __align__(8) unsigned long long win_bar;
        const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&win_bar));
        const uint32_t dst0 = static_cast<uint32_t>(__cvta_generic_to_shared(smem_win));
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t total = static_cast<uint32_t>(pitch) * static_cast<uint32_t>(mp.wh);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
        }
        __syncthreads();
        for (int row = threadIdx.x; row < mp.wh; row += kRayThreads) {
            const uint8_t* src = gsrc + static_cast<int64_t>(wy0 + row) * gpitch;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             dst0 + static_cast<uint32_t>(row * pitch)),
                         "l"(src), "r"(static_cast<uint32_t>(pitch)), "r"(bar)
                         : "memory");
        }
        {   // every thread waits for phase 0 of the barrier: all bytes have landed
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(bar)
                    : "memory");
            }
        }
    }
    // shared-space address of the window; the volatile asm keeps it in a register instead of
    // being rematerialised (S2R + LEA) inside the march loop
    uint32_t win_saddr;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(win_saddr) : "l"(smem_win));
    const RefGrid rg{mp.grid, mp.W, mp.H, mp.res, mp.ox, mp.oy};
    const int tw = pin_reg(M + 1);
    const double* slice = a.slice + static_cast<int64_t>(f) * R * tw;
    asm volatile("mov.b64 %0, %0;" : "+l"(slice));   // keep the table base in registers across rays
    const int32_t* perm = a.perm ? a.perm + fo + a.lo : nullptr;
    int replays = 0;

    for (int64_t s = static_cast<int64_t>(blockIdx.x) * kRayThreads + threadIdx.x; s < a.cnt;
         s += static_cast<int64_t>(gridDim.x) * kRayThreads) {
        const int64_t i = a.lo + (perm ? perm[s] : s);
        const double x = a.px[fo + i], y = a.py[fo + i], th = a.pt[fo + i];
        double sth, cth;
        sincos(th, &sth, &cth);
        const double qx = p_coord(x, mp.ox, mp.res, kPadL);
        const double qy = p_coord(y, mp.oy, mp.res, kPadL);
        const bool inside = p_inside(qx, qy, mp.PW, mp.PH);
        double acc = 1.0;
        uint8_t* steps = a.steps ? a.steps + (fo + i) * R : nullptr;
        if (!inside) {
            // first sample is already out of bounds for every beam (:632-636): step 0
            for (int j = 0; j < R; ++j) {
                acc = __dmul_rn(acc, __ldg(slice + static_cast<unsigned>(j * tw)));
                if (steps) steps[j] = 0;
            }
        } else {
            const int fqx = static_cast<int>(floor(qx)), fqy = static_cast<int>(floor(qy));
            const RayStart st = make_ray_start(qx, qy, fqx, fqy);
            const bool in_win = (mp.ww > 0) && fqx >= vx0 && fqx < vx1 && fqy >= vy0 && fqy < vy1;
            const double cths = cth * static_cast<double>(kOne), sths = sth * static_cast<double>(kOne);
            WindowV8S wacc8{0, 0};
            WindowV4S wacc4{0, 0, 0, 0};
            if (WBITS == 8)
                wacc8 = WindowV8S{static_cast<uint32_t>(pin_reg(static_cast<int>(win_saddr) + (st.by - wy0) * pitch + (st.bx - wx0))), pitch};
            else
                wacc4 = WindowV4S{win_saddr, pin_reg(st.bx - wx0), pin_reg(st.by - wy0), pitch};
            const GlobalV8 gacc = make_global_v8(mp.v8, mp.PW, st.bx, st.by);
            for (int j = 0; j < R; ++j) {
                int dxf, dyf;
                beam_direction_prescaled(cths, sths, a.beams.cosa[j], a.beams.sina[j], &dxf, &dyf);
                const ReplayArgs ra{x, y, th, a.beams.angle[j]};
                int r;
                if (in_win)
                    r = WBITS == 8 ? march_ray(wacc8, st, dxf, dyf, M, rg, ra, &replays)
                                   : march_ray(wacc4, st, dxf, dyf, M, rg, ra, &replays);
                else
                    r = march_ray(gacc, st, dxf, dyf, M, rg, ra, &replays);
                acc = __dmul_rn(acc, __ldg(slice + static_cast<unsigned>(j * tw + r)));
                if (steps) steps[j] = static_cast<uint8_t>(r);
            }
        }
        a.w_raw[fo + i] = squash_pow(acc, a.inv_squash);
    }
    if (a.replay_count && replays) atomicAdd(reinterpret_cast<unsigned long long*>(a.replay_count), static_cast<unsigned long long>(replays));
}

// ------------------------------------------------------------------------------------------
// normalisation (:679-686) + expected pose (:696-716)
// ------------------------------------------------------------------------------------------
struct NormArgs {
    int64_t N;
    const double* w_raw;      // [F][N]
    const double* total;      // [F] exact sequential sum of w_raw (nullptr: weights already normalised)
    double* wn;               // [F][N]
    const double* px;
    const double* py;
    const double* pt;
    double* partial;          // [F][nblk][4]
    int nblk;
    unsigned int* done;       // [F] block-completion counters (zero on entry, reset on exit) or nullptr
    double* pose_out;         // [F][3]: written by the last block to finish when `done` is given
    double* pose_host;        // same values into mapped pinned host memory (nullable): saves the D2H copy
    unsigned long long* update_no;   // incremented once when the update's pose is written (nullable)
};
constexpr int kNormThreads = 256;

__global__ void __launch_bounds__(kNormThreads) k_normalize_pose(NormArgs a) {
    pdl_enter();
    __shared__ double sm[kNormThreads / 32];
    const int f = blockIdx.y;
    const int64_t fo = static_cast<int64_t>(f) * a.N;
    const double tot = a.total ? a.total[f] : 0.0;
    const bool do_div = a.total && tot > 0.0;   // `if (sum_weights > 0)` (:680)
    double ax = 0.0, ay = 0.0, as = 0.0, ac = 0.0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kNormThreads + threadIdx.x; i < a.N;
         i += static_cast<int64_t>(gridDim.x) * kNormThreads) {
        double w = a.w_raw[fo + i];
        if (do_div) w = __ddiv_rn(w, tot);
        if (a.wn) a.wn[fo + i] = w;
        double s, c;
        sincos(a.pt[fo + i], &s, &c);
        ax += w * a.px[fo + i];
        ay += w * a.py[fo + i];
        as += w * s;
        ac += w * c;
    }
    ax = block_sum<kNormThreads>(ax, sm);
    ay = block_sum<kNormThreads>(ay, sm);
    as = block_sum<kNormThreads>(as, sm);
    ac = block_sum<kNormThreads>(ac, sm);
    if (threadIdx.x == 0) {
        double* p = a.partial + (static_cast<int64_t>(f) * a.nblk + blockIdx.x) * 4;
        p[0] = ax;
        p[1] = ay;
        p[2] = as;
        p[3] = ac;
    }
    if (!a.done) return;
    // the last block of this filter folds the per-block partial sums in block order
    // (deterministic) and writes the pose -- no second launch
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(a.done + f, 1u) == static_cast<unsigned>(gridDim.x) - 1u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double v[4] = {0, 0, 0, 0};
    for (int b = threadIdx.x; b < a.nblk; b += kNormThreads) {
        const volatile double* p = a.partial + (static_cast<int64_t>(f) * a.nblk + b) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] += p[q];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = block_sum<kNormThreads>(v[q], sm);
    if (threadIdx.x == 0) {
        const double th = atan2(v[2], v[3]);
        a.pose_out[3 * f + 0] = v[0];
        a.pose_out[3 * f + 1] = v[1];
        a.pose_out[3 * f + 2] = th;
        if (a.pose_host) {
            a.pose_host[3 * f + 0] = v[0];
            a.pose_host[3 * f + 1] = v[1];
            a.pose_host[3 * f + 2] = th;
        }
        a.done[f] = 0;
        if (a.update_no && f == 0) *a.update_no += 1ull;   // this update is complete: the next one draws fresh noise
    }
}

// ------------------------------------------------------------------------------------------
// initialisers (:382-446), batch ray queries (:586-609), weighted sub-sampling (:946-958)
// ------------------------------------------------------------------------------------------
struct InitArgs {
    int64_t N;
    double* px;
    double* py;
    double* pt;
    double* wn;
    const double* pose;        // [F][3] (init_pose)
    const double* normals;     // [F][3N] or nullptr
    const int32_t* cell;       // [F][N] or nullptr (init_global)
    const double* theta;       // [F][N] or nullptr
    const int32_t* free_cells;
    int n_free, W;
    double res, ox, oy;
    double w0;                 // initial weight 1 / max_particles (:107, :397, :444)
    uint64_t seed, stream_no;
    int filter0;               // first filter this launch applies to
    int64_t glo;               // global index of local particle 0 (sharded filter): keys the device RNG; injected arrays are local
};

__global__ void k_init_pose(InitArgs a) {
    const int f = a.filter0 + blockIdx.y;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const int64_t fo = static_cast<int64_t>(f) * a.N;
    double z0, z1, z2;
    if (a.normals) {
        const double* z = a.normals + 3 * (static_cast<int64_t>(blockIdx.y) * a.N + i);
        z0 = z[0];
        z1 = z[1];
        z2 = z[2];
    } else {
        const int64_t gi = a.glo + i;
        const Philox4 r = philox4x32_10(static_cast<uint32_t>(gi), static_cast<uint32_t>(gi >> 32), static_cast<uint32_t>(f), 7u,
                                        static_cast<uint32_t>(a.seed) ^ static_cast<uint32_t>(a.stream_no),
                                        static_cast<uint32_t>(a.seed >> 32) ^ 0x9e3779b9u);
        double t;
        normal_pair(r.v[0], r.v[1], &z0, &z1);
        normal_pair(r.v[2], r.v[3], &z2, &t);
    }
    const double* pose = a.pose + 3 * blockIdx.y;
    a.px[fo + i] = __dadd_rn(pose[0], __dmul_rn(z0, 0.5));
    a.py[fo + i] = __dadd_rn(pose[1], __dmul_rn(z1, 0.5));
    a.pt[fo + i] = wrap_angle_dev(__dadd_rn(pose[2], __dmul_rn(z2, 0.4)));
    a.wn[fo + i] = a.w0;
}

__global__ void k_init_global(InitArgs a) {
    const int f = a.filter0 + blockIdx.y;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const int64_t fo = static_cast<int64_t>(f) * a.N;
    int32_t ord;
    double th;
    if (a.cell) {
        ord = a.cell[static_cast<int64_t>(blockIdx.y) * a.N + i];
        th = a.theta[static_cast<int64_t>(blockIdx.y) * a.N + i];
    } else {
        const int64_t gi = a.glo + i;
        const Philox4 r = philox4x32_10(static_cast<uint32_t>(gi), static_cast<uint32_t>(gi >> 32), static_cast<uint32_t>(f), 11u,
                                        static_cast<uint32_t>(a.seed) ^ static_cast<uint32_t>(a.stream_no),
                                        static_cast<uint32_t>(a.seed >> 32) ^ 0x85ebca6bu);
        ord = static_cast<int32_t>(__umulhi(r.v[0], static_cast<uint32_t>(a.n_free)));
        th = canonical_from_words(r.v[1], r.v[2]) * (2.0 * 3.14159265358979323846);
    }
    ord = max(0, min(ord, a.n_free - 1));
    const int32_t lin = a.free_cells[ord];
    const int row = lin / a.W, col = lin - row * a.W;
    a.px[fo + i] = __dadd_rn(__dmul_rn(static_cast<double>(col), a.res), a.ox);   // :438
    a.py[fo + i] = __dadd_rn(__dmul_rn(static_cast<double>(row), a.res), a.oy);   // :439
    a.pt[fo + i] = th;
    a.wn[fo + i] = a.w0;
}

struct QueryArgs {
    MapDev map;
    const double* q;     // column-major n x 3 on the device
    int64_t n;
    float* out;
    double max_range;
};

__global__ void __launch_bounds__(256) k_range_queries(QueryArgs a) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= a.n) return;
    const MapDev& mp = a.map;
    const double x = a.q[i], y = a.q[a.n + i], ang = a.q[2 * a.n + i];
    const double qx = p_coord(x, mp.ox, mp.res, kPadL);
    const double qy = p_coord(y, mp.oy, mp.res, kPadL);
    const bool inside = p_inside(qx, qy, mp.PW, mp.PH);
    int r = 0;
    if (inside) {
        double s, c;
        sincos(ang, &s, &c);
        const RayStart st = make_ray_start(qx, qy, static_cast<int>(floor(qx)), static_cast<int>(floor(qy)));
        const GlobalV8 gacc = make_global_v8(mp.v8, mp.PW, st.bx, st.by);
        const RefGrid rg{mp.grid, mp.W, mp.H, mp.res, mp.ox, mp.oy};
        const ReplayArgs ra{x, y, ang, 0.0f};
        int dxf, dyf;
        beam_direction_fixed(c, s, 1.0, 0.0, &dxf, &dyf);
        r = march_ray(gacc, st, dxf, dyf, mp.M, rg, ra, nullptr);
    }
    // `return step * map_resolution_` / `return MAX_RANGE_METERS` as float (:635, :644, :649)
    a.out[i] = (r >= mp.M) ? static_cast<float>(a.max_range) : static_cast<float>(__dmul_rn(static_cast<double>(r), mp.res));
}

__global__ void k_steps_to_ranges(const uint8_t* steps, int64_t n, int M, double res, double max_range, float* out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = steps[i];
    out[i] = (r >= M) ? static_cast<float>(max_range) : static_cast<float>(__dmul_rn(static_cast<double>(r), res));
}

// weighted sub-sample for visualisation: k draws from the CDF of the current weights
// (visualize() :946-958: discrete_distribution(weights_) drawn max_viz_particles times; u_in = the
// canonical uniforms it would draw, or nullptr for the device RNG)
__global__ void k_sample_particles(const double* cdf, int64_t N, const double* px, const double* py, const double* pt,
                                   int k, uint64_t seed, uint64_t stream_no, const double* u_in, double* out, int32_t* idx_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    double u;
    if (u_in) {
        u = u_in[i];
    } else {
        const Philox4 r = philox4x32_10(static_cast<uint32_t>(i), 0u, 0u, 13u, static_cast<uint32_t>(seed) ^ static_cast<uint32_t>(stream_no),
                                        static_cast<uint32_t>(seed >> 32) ^ 0xc2b2ae35u);
        u = canonical_from_words(r.v[0], r.v[1]);
    }
    int64_t lo = 0, hi = N;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cdf[mid] < u)
            lo = mid + 1;
        else
            hi = mid;
    }
    if (lo >= N) lo = N - 1;
    if (N < 2) lo = 0;        // libstdc++ clears the table for fewer than 2 weights
    out[i] = px[lo];
    out[k + i] = py[lo];
    out[2 * k + i] = pt[lo];
    if (idx_out) idx_out[i] = static_cast<int32_t>(lo);
}

// ------------------------------------------------------------------------------------------
// gather micro-benchmark: random single-byte reads from an L2-resident array or
// from a shared-memory window -- the access pattern of the ray march, without its arithmetic
// ------------------------------------------------------------------------------------------
template <bool SHARED>
__global__ void __launch_bounds__(1024, 1) k_gather_bench(const uint8_t* __restrict__ arr, uint32_t mask, int iters,
                                                         unsigned long long* sink) {
    extern __shared__ __align__(16) uint8_t smem_win[];
    if (SHARED) {
        for (uint32_t i = threadIdx.x * 16u; i <= mask; i += blockDim.x * 16u)
            *reinterpret_cast<uint4*>(smem_win + i) = __ldg(reinterpret_cast<const uint4*>(arr + i));
        __syncthreads();
    }
    uint32_t s0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t s1 = s0 ^ 0x9e3779b9u, s2 = s0 + 0x85ebca6bu, s3 = s0 * 0xc2b2ae35u + 1u;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        // four independent address streams per thread (xorshift), like four rays in flight
        s0 ^= s0 << 13; s0 ^= s0 >> 17; s0 ^= s0 << 5;
        s1 ^= s1 << 13; s1 ^= s1 >> 17; s1 ^= s1 << 5;
        s2 ^= s2 << 13; s2 ^= s2 >> 17; s2 ^= s2 << 5;
        s3 ^= s3 << 13; s3 ^= s3 >> 17; s3 ^= s3 << 5;
        if (SHARED) {
            acc += smem_win[s0 & mask] + smem_win[s1 & mask] + smem_win[s2 & mask] + smem_win[s3 & mask];
        } else {
            acc += __ldg(arr + (s0 & mask)) + __ldg(arr + (s1 & mask)) + __ldg(arr + (s2 & mask)) + __ldg(arr + (s3 & mask));
        }
    }
    if (acc == 0xffffffffu) atomicAdd(sink, 1ull);   // keeps the loads alive
}

__global__ void k_fill(double* p, int64_t n, double v) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace mclb200

namespace mclb200 {

// ------------------------------------------------------------------------------------------
// Wide configurations: MAX_RANGE_PX > 254 (e.g. max_range 15 m at 0.05 m/cell) or more than 128
// beams (angle_step < 9 on a 1080-beam scan).  The reference has neither limit (:195, :307-310); the
// skip-map stages do (9.23 fixed point, one-byte step indices, beam tables in the kernel parameters).
// Such contexts march every ray exactly as cast_ray does (:611-650) -- FP64 position accumulated
// sample by sample, truncating quotients, int8 grid -- with 16-bit step indices.  Correct by
// construction and still one thread per ray, but without skipping: a compatibility path.
// ------------------------------------------------------------------------------------------
constexpr int kMaxBeamsWide = 4096;

__device__ __forceinline__ int ref_cast_steps(const RefGrid& g, double x, double y, double angle, int M) {
    double sn, cs;
    sincos(angle, &sn, &cs);
    const double dx = nf_mul(cs, g.res), dy = nf_mul(sn, g.res);   // :619-620
    double cx = x, cy = y;
    for (int step = 0; step < M; ++step) {
        cx = nf_add(cx, dx);                                        // :624-625
        cy = nf_add(cy, dy);
        const double qx = nf_div(nf_sub(cx, g.ox), g.res), qy = nf_div(nf_sub(cy, g.oy), g.res);   // :628-629
        // (int) of a NaN or of a value beyond int range is INT_MIN on the reference's platform: out of bounds
        if (!(fabs(qx) < 2147483648.0) || !(fabs(qy) < 2147483648.0)) return step;
        const int gx = __double2int_rz(qx), gy = __double2int_rz(qy);
        if (gx < 0 || gx >= g.W || gy < 0 || gy >= g.H) return step;                                 // :632-636
        if (g.data[static_cast<int64_t>(gy) * g.W + gx] > 50) return step;                           // :639-645
    }
    return M;   // MAX_RANGE_METERS (:649)
}

struct WideRayArgs {
    RefGrid grid;
    int M, R;
    int64_t N;                 // particles per filter
    const double* px;          // [F][N]
    const double* py;
    const double* pt;
    const float* beam;         // [R] downsampled_angles_
    uint16_t* steps;           // [F][N][R]
};

__global__ void __launch_bounds__(256) k_raycast_wide(WideRayArgs a) {
    const int f = blockIdx.y;
    const int64_t rays = a.N * a.R;
    for (int64_t g = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; g < rays; g += static_cast<int64_t>(gridDim.x) * 256) {
        const int64_t i = g / a.R;
        const int j = static_cast<int>(g - i * a.R);
        const int64_t p = static_cast<int64_t>(f) * a.N + i;
        const double angle = nf_add(a.pt[p], static_cast<double>(a.beam[j]));   // theta + angle (:533)
        a.steps[p * a.R + j] = static_cast<uint16_t>(ref_cast_steps(a.grid, a.px[p], a.py[p], angle, a.M));
    }
}

struct WideWeightArgs {
    int M, R;
    int64_t N;
    const uint16_t* steps;     // [F][N][R]
    const double* slice;       // [F][R][M+1]
    double* w_raw;             // [F][N]
    double inv_squash;
};

// w = pow(prod_j table(obs_j, range_ij), 1/squash), entries multiplied in beam order (:564-579)
__global__ void __launch_bounds__(256) k_weight_wide(WideWeightArgs a) {
    const int f = blockIdx.y;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= a.N) return;
    const int tw = a.M + 1;
    const uint16_t* st = a.steps + (static_cast<int64_t>(f) * a.N + i) * a.R;
    const double* row = a.slice + static_cast<int64_t>(f) * a.R * tw;
    double acc = 1.0;
    for (int j = 0; j < a.R; ++j) acc = __dmul_rn(acc, __ldg(row + static_cast<int64_t>(j) * tw + st[j]));
    a.w_raw[static_cast<int64_t>(f) * a.N + i] = squash_pow(acc, a.inv_squash);
}

struct WideQueryArgs {
    RefGrid grid;
    int M;
    const double* q;     // column-major n x 3
    int64_t n;
    float* out;
    double max_range;
};

__global__ void __launch_bounds__(256) k_range_queries_wide(WideQueryArgs a) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (i >= a.n) return;
    const int r = ref_cast_steps(a.grid, a.q[i], a.q[a.n + i], a.q[2 * a.n + i], a.M);
    a.out[i] = (r >= a.M) ? static_cast<float>(a.max_range) : static_cast<float>(__dmul_rn(static_cast<double>(r), a.grid.res));
}

__global__ void k_steps16_to_ranges(const uint16_t* steps, int64_t n, int M, double res, double max_range, float* out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = steps[i];
    out[i] = (r >= M) ? static_cast<float>(max_range) : static_cast<float>(__dmul_rn(static_cast<double>(r), res));
}

__global__ void k_widen_steps(const uint8_t* in, int64_t n, uint16_t* out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

}  // namespace mclb200
