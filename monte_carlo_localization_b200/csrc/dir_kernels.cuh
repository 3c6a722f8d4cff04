// csrc/dir_kernels.cuh -- the directional ray stage (single filters of >= 1024 particles, batches on small maps).
//
// Replaces sensor_model's ray cast for every particle x beam (src/particle_filter.cpp:524-540,
// :586-650) and its table product (:564-579) when one filter is large enough for the heading
// sort to resolve sectors (>= kDirMinBuckets buckets):
//   k_build_dir_maps   once per map: the kDirSectors directional skip maps (dirmap.cuh)
//   (k_resample_motion) per update: every particle's ray-start record, in slot order
//   k_dir_gather       per update: the records moved to heading-sorted order (one 32 B gather each)
//   k_dir_plan         per update: per sector, the range of 1024-slot chunks of the heading-sorted
//                      particles that cast rays into it; a work unit = (sector, chunk)
//   k_raycast_dir      persistent CTAs pull units; the sector's window of the map is staged in
//                      shared memory with cp.async.bulk; one lane = one particle, marching the ~5 beams
//                      of it that fall into the sector (16 sectors, 60 beams over 270 degrees); step indices out
//   k_weight_steps     table product in the reference's beam order + pow -> raw weights
// A BATCH of filters whose whole padded map fits one window (small maps) runs the same stage over
// the pool of all filters' particles ("pool mode"): slots are (filter, heading-sorted) order, every
// sector's window is the whole sector map, every (sector, chunk) pair is a unit.
// Lanes of a warp hold heading-neighbours casting the same beam, so their rays share a sector,
// a window and nearly a trip count.  Particles outside the window box (all of them right after
// initialize_global) march the same sector maps in global memory: the 16 maps of a 2000 x 2000 grid are 65 MB and
// stay in the 126 MB L2, and a scattered cloud still needs a third of the isotropic kernel's lookups (measured on
// 2 M uniformly scattered particles: 2.4 ms against 5.4 ms on Spielberg_map, 2.9 against 8.9 ms on basement_fixed),
// so the stage no longer hands scattered clouds back to k_raycast_weight (round 1's 90 %-in-the-box rule).
#pragma once

namespace mclb200 {

constexpr int kDirThreads = 1024;     // threads of a ray CTA == sorted slots of one unit
constexpr int kDirGrab = 16;          // units a CTA takes per visit to the global counter

// 1: rows of the shared-memory window are padded to an odd multiple of 16 bytes, so that the rows a warp's lanes
// touch (heading-neighbours: a few cells apart in x and y) fall into different bank groups whatever the window width
#ifndef MCL_DIR_PAD
#define MCL_DIR_PAD 1
#endif

constexpr int kDirMaxRanges = 256;    // CTAs of the ray kernel (diagnostics array)

// Development builds only (-DMCL_DIR_DIAG=1, scripts/build_variant.sh): per-CTA counters of the last k_raycast_dir launch
// [cycles, units, pieces, windows staged, cycles in the scheduler, cycles staging windows, SM clock at start, at end]
#ifndef MCL_DIR_DIAG
#define MCL_DIR_DIAG 0
#endif
#if MCL_DIR_DIAG
__device__ unsigned long long g_dir_diag[kDirMaxRanges][8];
#endif

// ints of the plan buffer
enum { kPlanMode = 0, kPlanUnits = 1, kPlanCounter = 2, kPlanInBox = 3, kPlanBoxX = 4, kPlanBoxY = 5, kPlanInts = 8 };

struct DirBuildArgs {
    const uint8_t* v8;
    const float* gap;
    const DirSector* sectors;
    uint8_t* out;          // [S][PH*PW]
    int PW, PH;
};

__global__ void __launch_bounds__(256) k_build_dir_maps(DirBuildArgs a) {
    const int64_t cell = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    const int64_t n = static_cast<int64_t>(a.PW) * a.PH;
    if (cell >= n) return;
    const int s = blockIdx.y;
    const int cy = static_cast<int>(cell / a.PW), cx = static_cast<int>(cell - static_cast<int64_t>(cy) * a.PW);
    a.out[static_cast<int64_t>(s) * n + cell] = dir_code(a.v8, a.gap, a.PW, a.PH, cx, cy, a.sectors[s]);
}

// Ray-start record of one particle (32 bytes = one memory sector):
//   a = p0x, p0y (9.23 fixed-point start), bx | by << 16 (P-cell of local coordinate 0),
//       heading bucket | flags << 16 (bit 0: inside the map, bit 1: inside the window box)
//   b = cos(theta) * 2^23, sin(theta) * 2^23
struct __align__(32) DirRec {
    uint4 a;
    double2 b;
};

// k_dir_gather: ray-start records from slot order into heading-sorted order
struct DirPrepArgs {
    MapDev map;
    const double* centre;      // [2] sums of x and y over the shard's particles
    const DirRec* rec_in;      // [cnt] records in slot order (k_resample_motion)
    DirRec* rec;               // [cnt] records in heading-sorted order
    const int32_t* perm;       // sorted slot -> slot index (both relative to the shard / to the slot's filter)
    int* plan;
    int64_t cnt;
    int64_t nfil;              // particles per filter (pool mode: slot / nfil is the filter); >= cnt otherwise
    int box;
    int whole;                 // pool mode: the window is the whole map, every particle inside the map is "in the box"
};

// P-cell of the box corner: the box is centred on the cloud's mean position
__device__ __forceinline__ void dir_box_origin(const double* centre, int64_t cnt, const MapDev& mp, int box, int* bx0, int* by0) {
    const double mx = centre[0] / static_cast<double>(cnt), my = centre[1] / static_cast<double>(cnt);
    const double qx = (mx - mp.ox) / mp.res + kPadL, qy = (my - mp.oy) / mp.res + kPadL;
    const int cx = (qx > -1e9 && qx < 1e9) ? static_cast<int>(floor(qx)) : 0;
    const int cy = (qy > -1e9 && qy < 1e9) ? static_cast<int>(floor(qy)) : 0;
    *bx0 = cx - box / 2;
    *by0 = cy - box / 2;
}

__device__ __forceinline__ void dir_write_record(const MapDev& mp, DirRec* rec, int64_t slot, double x, double y, double th,
                                                 int bucket) {
    double sth, cth;
    sincos(th, &sth, &cth);
    const double qx = p_coord(x, mp.ox, mp.res, kPadL);
    const double qy = p_coord(y, mp.oy, mp.res, kPadL);
    uint4 r0 = make_uint4(0u, 0u, 0u, static_cast<uint32_t>(bucket));
    if (p_inside(qx, qy, mp.PW, mp.PH)) {
        const int fqx = static_cast<int>(floor(qx)), fqy = static_cast<int>(floor(qy));
        const RayStart st = make_ray_start(qx, qy, fqx, fqy);
        r0.x = st.p0x;
        r0.y = st.p0y;
        r0.z = (static_cast<uint32_t>(st.bx) & 0xffffu) | (static_cast<uint32_t>(st.by) << 16);
        r0.w |= 1u << 16;
    }
    rec[slot].a = r0;
    rec[slot].b = make_double2(cth * static_cast<double>(kOne), sth * static_cast<double>(kOne));
}

// Moves every particle's record (written in slot order by k_resample_motion) to its heading-
// sorted slot, one 32-byte gather per particle, and sets the window-box flag -- the box needs the
// cloud centre, which is complete only after the motion kernel.
__global__ void __launch_bounds__(256) k_dir_gather(DirPrepArgs a) {
    pdl_enter();
    __shared__ int s_cnt[8];
    const int64_t pos = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    int bx0, by0;
    dir_box_origin(a.centre, a.cnt, a.map, a.box, &bx0, &by0);
    int in_box = 0;
    if (pos < a.cnt) {
        const int64_t i = (pos / a.nfil) * a.nfil + a.perm[pos];
        uint4 r0 = a.rec_in[i].a;
        const double2 r1 = a.rec_in[i].b;
        if (r0.w >> 16) {
            const int fqx = static_cast<int>(static_cast<int16_t>(r0.z & 0xffffu)) + kLocalOrigin;
            const int fqy = static_cast<int>(static_cast<int16_t>(r0.z >> 16)) + kLocalOrigin;
            in_box = a.whole || (fqx >= bx0 && fqx < bx0 + a.box && fqy >= by0 && fqy < by0 + a.box);
            if (in_box) r0.w |= 2u << 16;
        }
        a.rec[pos].a = r0;
        a.rec[pos].b = r1;
    }
    const unsigned m = __ballot_sync(kFullMask, in_box);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_cnt[w];
        if (t) atomicAdd(a.plan + kPlanInBox, t);
    }
}

struct DirPlanArgs {
    const int* hist;           // [B] heading histogram of the sort
    int io[kMaxBeams];         // per-beam bucket offset (dir_beam_offset)
    int* plan;
    int* sec_tab;              // [S + 1] first unit of every sector | [S] first chunk of every sector
    int64_t cnt;
    int B, R, force;           // force: 0 decide by the in-box fraction, 2 always directional
    int all_chunks;            // pool mode: every sector covers every chunk (slots are not globally heading-sorted)
};

constexpr int kPlanThreads = 1024;

// Work decomposition of one update.  For sector s and beam j the particles whose beam-j ray lies
// in s occupy one cyclic run of B/S heading buckets, i.e. one or two runs of sorted slots.  The
// chunks (1024 slots) touched by any beam's run form the sector's chunk range; a UNIT is one
// (sector, chunk) pair, numbered sector by sector so that consecutive units share a window.
__global__ void __launch_bounds__(kPlanThreads) k_dir_plan(DirPlanArgs a) {
    pdl_enter();
    __shared__ int off[kMaxBuckets + 1];
    __shared__ int wsum[kPlanThreads / 32];
    __shared__ int cmin[kDirSectors], cmax[kDirSectors];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < kDirSectors) {
        cmin[tid] = 0x7fffffff;
        cmax[tid] = 0;
    }
    // ---- exclusive scan of the histogram: off[b] = first sorted slot of bucket b ----------
    constexpr int kPer = kMaxBuckets / kPlanThreads;
    int h[kPer], loc = 0;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int b = tid * kPer + q;
        h[q] = (b < a.B && !a.all_chunks) ? a.hist[b] : 0;   // pool mode: the sort's histograms are per filter, not used
        loc += h[q];
    }
    int v = loc;
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
        off[tid * kPer + q] = excl;
        excl += h[q];
    }
    if (tid == kPlanThreads - 1) off[kMaxBuckets] = excl;
    __syncthreads();
    // ---- chunk range of every sector: union over (beam, half) of the chunks its slots touch
    const int K = a.B / kDirSectors;
    const int npieces = 2 * kDirSectors * a.R;
    for (int p = tid; p < npieces; p += kPlanThreads) {
        const int half = p & 1, sj = p >> 1, j = sj % a.R, s = sj / a.R;
        const int bs = (s * K - a.io[j]) & (a.B - 1);          // first bucket of the run
        const int be = bs + K;                                 // one past, may exceed B (wraps)
        int lo_b, hi_b;
        if (half == 0) {
            lo_b = bs;
            hi_b = be < a.B ? be : a.B;
        } else {
            lo_b = 0;
            hi_b = be > a.B ? be - a.B : 0;
        }
        const int lo_s = off[lo_b], hi_s = off[hi_b];
        if (hi_s > lo_s) {
            atomicMin(&cmin[s], lo_s / kDirThreads);
            atomicMax(&cmax[s], (hi_s - 1) / kDirThreads + 1);
        }
    }
    __syncthreads();
    if (a.all_chunks && tid < kDirSectors) {
        cmin[tid] = 0;
        cmax[tid] = static_cast<int>((a.cnt + kDirThreads - 1) / kDirThreads);
    }
    __syncthreads();
    if (warp == 0) {
        const int ls = lane < kDirSectors ? lane : kDirSectors - 1;            // one lane per sector (kDirSectors <= 32)
        const int n = (lane < kDirSectors && cmax[ls] > cmin[ls]) ? cmax[ls] - cmin[ls] : 0;
        int w = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(kFullMask, w, d);
            if (lane >= d) w += o;
        }
        if (lane < kDirSectors) {
            a.sec_tab[lane] = w - n;
            a.sec_tab[kDirSectors + 1 + lane] = n ? cmin[ls] : 0;
        }
        if (lane == 31) {
            a.sec_tab[kDirSectors] = w;
            const bool dir = a.force != 1;   // (plan[kPlanInBox]: particles that march a shared-memory window, diagnostic)
            a.plan[kPlanMode] = dir ? 1 : 0;
            a.plan[kPlanUnits] = dir ? w : 0;
            a.plan[kPlanCounter] = 0;
            a.plan[kPlanInBox] = 0;
        }
    }
}
static_assert(kDirSectors <= 32 && (kDirSectors & (kDirSectors - 1)) == 0, "k_dir_plan scans the sectors with one warp");

// Everything the rare exact-replay path needs, kept in device memory so that the hot path
// carries one pointer: the reference grid and the pose arrays of the update's destination buffer.
struct DirReplayCtx {
    RefGrid grid;
    const double* px;
    const double* py;
    const double* pt;
    const int32_t* perm;       // already offset to the shard: perm[pos] is relative to lo (to the slot's filter in pool mode)
    int64_t lo;
    int64_t nfil;              // particles per filter in pool mode, >= the slot count otherwise
    float angle[kMaxBeams];
};
struct ReplayLazy {
    const DirReplayCtx* rc;
    int pos, j;
    __device__ __forceinline__ ReplayArgs load() const {
        const int64_t i = rc->lo + (pos / rc->nfil) * rc->nfil + rc->perm[pos];
        return ReplayArgs{rc->px[i], rc->py[i], rc->pt[i], rc->angle[j]};
    }
    __device__ __forceinline__ RefGrid grid() const { return rc->grid; }
};

struct DirRayArgs {
    MapDev map;
    BeamDev beams;
    int io[kMaxBeams];          // per-beam bucket offset
    const DirSector* sectors;   // [S]
    const uint8_t* dirmaps;     // [S][PH*PW]
    const DirRec* rec;          // heading-sorted ray-start records
    const int* sec_tab;         // [S + 1] first unit | [S] first chunk
    int* plan;
    const double* centre;
    const DirReplayCtx* replay;
    int64_t cnt, stride;
    uint8_t* steps_sorted;      // [R][stride]
    int64_t* replay_count;
    int B, shift, box;
    int win_bytes;              // shared-memory bytes reserved for the sector window
    int whole;                  // pool mode: the window of every sector is its whole map
    int beam_ranges;            // 1: the beam offsets io[] are cyclically non-decreasing (any real scan): the beams of a
                                // (warp, sector) are one or two INDEX RANGES read from a table; 0: found by ballots
};

// dynamic shared memory of k_raycast_dir: window | rec0 prefetch [2][1024] | rec1 prefetch [2][1024]
__host__ __device__ inline size_t dir_ray_smem(int win_bytes) {
    return static_cast<size_t>(win_bytes) + 2 * kDirThreads * (sizeof(uint4) + sizeof(double2));
}

template <int MC>
__global__ void __launch_bounds__(kDirThreads, 1) k_raycast_dir(DirRayArgs a) {
    pdl_enter();
    extern __shared__ __align__(16) uint8_t smem_win[];
    __shared__ __align__(8) unsigned long long win_bar;
    __shared__ unsigned s_u0;
    __shared__ uint32_t s_units[kDirGrab];    // chunk << 6 | sector
    __shared__ int s_sec[2 * kDirSectors + 2];
    __shared__ int s_io[kMaxBeams + 32];          // padded: a warp reads 32 consecutive entries from any beam
    __shared__ uint16_t s_cnt[kMaxBuckets + 1];   // s_cnt[x] = number of beams whose offset, counted from beam 0's, is below x
    if (a.plan[kPlanMode] != 1) return;
#if MCL_DIR_DIAG
    unsigned long long dg_t0 = clock64(), dg_units = 0, dg_pieces = 0, dg_wins = 0, dg_sched = 0, dg_stage = 0;
    unsigned long long dg_g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dg_g0));
#endif
    const unsigned n_units = static_cast<unsigned>(a.plan[kPlanUnits]);
    const MapDev& mp = a.map;
    const int M = MC > 0 ? MC : pin_reg(mp.M);
    const int tid = threadIdx.x, lane = tid & 31;
    const int R = a.beams.R;
    const int cnt = static_cast<int>(a.cnt);
    int bx0, by0;
    dir_box_origin(a.centre, a.cnt, mp, a.box, &bx0, &by0);
    const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&win_bar));
    uint32_t win_saddr;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(win_saddr) : "l"(smem_win));
    // thread-private prefetch slots of the ray-start records (two buffers)
    const uint32_t rec0_s = win_saddr + static_cast<uint32_t>(a.win_bytes) + static_cast<uint32_t>(tid) * 16u;
    const uint32_t rec1_s = rec0_s + 2u * kDirThreads * 16u;
    const uint32_t io_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_io));
    if (tid < 2 * kDirSectors + 1) s_sec[tid] = a.sec_tab[tid];
    if (tid < kMaxBeams + 32) s_io[tid] = tid < R ? a.io[tid] : 0;
    if (a.beam_ranges) {
        __syncthreads();
        const int io0 = s_io[0];
        for (int x = tid; x <= a.B; x += kDirThreads) {   // beams with ((io_j - io_0) mod B) < x: a binary search, the offsets are sorted
            int lo = 0, hi = R;
            while (lo < hi) {
                const int m = (lo + hi) >> 1;
                if (((s_io[m] - io0) & (a.B - 1)) < x)
                    lo = m + 1;
                else
                    hi = m;
            }
            s_cnt[x] = static_cast<uint16_t>(lo);
        }
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int64_t ncell = static_cast<int64_t>(mp.PW) * mp.PH;
    const int K = a.B / kDirSectors, Bmask = a.B - 1;
    uint32_t phase = 0;
    unsigned seen = 0;
    int cur_s = -1, replays = 0;
    int spitch = 0;                     // row pitch of the window in shared memory (dir_smem_pitch)
    DirWindow wg{0, 0, 0, 0};
    const uint8_t* smap = a.dirmaps;

    // asynchronous copy of this thread's records of a unit into prefetch buffer `buf`
    auto prefetch = [&](uint32_t unit, uint32_t buf) {
        // slots beyond the last particle repeat it: every lane always holds a real record (the duplicates recompute
        // and rewrite the last particle's steps), so nothing below needs a validity test
        const int pos = min(static_cast<int>(unit >> 6) * kDirThreads + tid, cnt - 1);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(rec0_s + buf * (kDirThreads * 16u)), "l"(&a.rec[pos].a) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(rec1_s + buf * (kDirThreads * 16u)), "l"(&a.rec[pos].b) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    for (;;) {
        __syncthreads();                    // s_u0 / s_units of the previous round have been read by everyone
        // guided self-scheduling: large grabs first, small ones towards the end of the unit list
        const unsigned left = n_units > seen ? n_units - seen : 0u;
        const unsigned grab = max(2u, min(static_cast<unsigned>(kDirGrab), left / (2u * gridDim.x)));
        if (tid == 0) s_u0 = atomicAdd(reinterpret_cast<unsigned*>(a.plan + kPlanCounter), grab);
        __syncthreads();
        const unsigned u0 = s_u0;
        if (u0 >= n_units) break;
        seen = u0 + grab;
        const unsigned nu = min(grab, n_units - u0);
#if MCL_DIR_DIAG
        dg_units += nu;
        dg_pieces += 1;
#endif
        if (tid < static_cast<int>(nu)) {
            // sector of unit u: the last sector whose first unit is <= u (empty sectors share offsets)
            const int u = static_cast<int>(u0) + tid;
            int s = 0;
#pragma unroll
            for (int d = kDirSectors / 2; d > 0; d >>= 1)
                if (s_sec[s + d] <= u) s += d;
            s_units[tid] = (static_cast<uint32_t>(s_sec[kDirSectors + 1 + s] + (u - s_sec[s])) << 6) | static_cast<uint32_t>(s);
        }
        __syncthreads();
        prefetch(s_units[0], 0u);
        for (unsigned k = 0; k < nu; ++k) {
            const uint32_t unit = s_units[k];
            const int s = static_cast<int>(unit & 63u);
            if (s != cur_s) {
                // ---- stage sector s's window: one bulk copy per row, bytes counted on the mbarrier
#if MCL_DIR_DIAG
                const unsigned long long dg_w0 = clock64();
                dg_wins++;
#endif
                __syncthreads();            // every warp has left the old window
                wg = a.whole ? DirWindow{0, 0, mp.PW, mp.PH} : dir_window(a.sectors[s], bx0, by0, a.box, mp.PW, mp.PH);
                smap = a.dirmaps + static_cast<int64_t>(s) * ncell;
                spitch = (MCL_DIR_PAD && !a.whole) ? dir_smem_pitch(wg.pitch) : wg.pitch;
                if (tid == 0) {
                    const uint32_t total = static_cast<uint32_t>(wg.pitch) * static_cast<uint32_t>(wg.rows);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
                }
                __syncthreads();
                const uint8_t* gsrc = smap + static_cast<int64_t>(wg.wy0) * mp.PW + wg.wx0;
                for (int row = tid; row < wg.rows; row += kDirThreads) {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     win_saddr + static_cast<uint32_t>(row * spitch)),
                                 "l"(gsrc + static_cast<int64_t>(row) * mp.PW), "r"(static_cast<uint32_t>(wg.pitch)), "r"(bar)
                                 : "memory");
                }
                uint32_t done = 0;
                while (!done) {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                        "selp.u32 %0, 1, 0, p;\n\t}"
                        : "=r"(done)
                        : "r"(bar), "r"(phase)
                        : "memory");
                }
                phase ^= 1u;
                cur_s = s;
#if MCL_DIR_DIAG
                dg_stage += clock64() - dg_w0;
#endif
            }
            // this unit's records have landed; start the copy of the next unit's
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            const uint32_t buf = k & 1u;
            uint4 r0;
            double2 cs;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w) : "r"(rec0_s + buf * (kDirThreads * 16u)));
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(cs.x), "=d"(cs.y) : "r"(rec1_s + buf * (kDirThreads * 16u)));
            if (k + 1 < nu) prefetch(s_units[k + 1], buf ^ 1u);
            const int pos = min(static_cast<int>(unit >> 6) * kDirThreads + tid, cnt - 1);
            const int bucket = static_cast<int>(r0.w & 0xffffu), flags = static_cast<int>(r0.w >> 16);
            // ---- beams with rays in this sector for some lane of the warp (lanes are bucket-sorted)
            const int bmin = __reduce_min_sync(kFullMask, bucket);
            const int bmax = __reduce_max_sync(kFullMask, bucket);
            RayStart st;
            st.p0x = r0.x;
            st.p0y = r0.y;
            st.bx = static_cast<int>(static_cast<int16_t>(r0.z & 0xffffu));
            st.by = static_cast<int>(static_cast<int16_t>(r0.z >> 16));
            const WindowV8S wacc{static_cast<uint32_t>(pin_reg(static_cast<int>(win_saddr) + (st.by - wg.wy0) * spitch + (st.bx - wg.wx0))), spitch};
            const GlobalV8 gacc = make_global_v8(smap, mp.PW, st.bx, st.by);
            // A particle outside the map: the first sample of every beam is already out of bounds (:632-636), step index
            // 0.  Its march starts beyond M (no lookup is made) and the result is masked to 0.
            const int inmap = (flags & 1) ? 0xff : 0;
            int k0 = (flags & 1) ? 1 : M + 1;
            // the start cell's code lets all of this particle's rays in the sector begin at sample k0 (march.cuh)
            // (particles outside the window box are rare and simply start at sample 1)
            if ((flags & 3) == 3) k0 = dir_first_sample(wacc, st);
            const RayStartOfs so = offset_ray_start(st);   // the march works on positions offset by kEta (march.cuh)
            uint8_t* const out = a.steps_sorted + pos;
            // one ray: beam j of this lane's particle
            auto cast = [&](int j) {
                int dxf, dyf;
                beam_direction_prescaled(cs.x, cs.y, a.beams.cosa[j], a.beams.sina[j], &dxf, &dyf);
                const ReplayLazy rep{a.replay, pos, j};
                int r;
                if (flags & 2)
                    r = march_ray_dir(wacc, so, dxf, dyf, M, rep, &replays, k0);
                else
                    r = march_ray_dir(gacc, so, dxf, dyf, M, rep, &replays, k0);
                out[static_cast<uint64_t>(static_cast<uint32_t>(j)) * static_cast<uint32_t>(a.stride)] = static_cast<uint8_t>(r & inmap);
            };
            const int span = bmax - bmin;
            if (!a.beam_ranges) {
                // beam offsets in no particular order (no real scan): every beam, every lane tests its own sector
                for (int j = 0; j < R; ++j)
                    if (dir_sector_of(bucket, s_io[j], Bmask, a.shift) == s) cast(j);
                continue;
            }
            // Beam j has rays in sector s for some bucket of [bmin, bmax] iff io_j lies in the cyclic interval
            // [s K - bmax, s K - bmin + K) of length K + span.  The offsets are sorted from beam 0 on, so that is
            // one index range of the table, [j0, j1), plus [0, j3) when the interval wraps past beam 0's offset.
            int j0 = 0, j1 = R, j3 = 0;
            if (K + span < a.B) {
                const int aU = (s * K - bmax - s_io[0]) & Bmask;
                j0 = s_cnt[aU];
                if (aU + K + span <= a.B) {
                    j1 = s_cnt[aU + K + span];
                } else {
                    j3 = s_cnt[aU + K + span - a.B];
                }
            }
            const int n1 = j1 - j0, nb = n1 + j3;      // the (warp, sector)'s beams: list index q -> beam
            const int base_u = bmin - s * K;
            for (int qb = 0; qb < nb; qb += 32) {
                // lane l looks at list entry qb + l: is EVERY lane's ray of that beam in the sector (the usual case: a warp
                // spans a few buckets of the sector's K)?  Then no per-lane sector test is needed for it.
                const int ql = qb + lane;
                const int jl = ql < n1 ? j0 + ql : ql - n1;            // < R + 32: s_io is padded
                uint32_t io_l;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(io_l) : "r"(io_base + static_cast<uint32_t>(jl) * 4u));
                const unsigned whole = __ballot_sync(kFullMask, ((base_u + static_cast<int>(io_l)) & Bmask) + span < K);
                const int qe = min(nb, qb + 32);
                for (int q = qb; q < qe; ++q) {
                    const int j = q < n1 ? j0 + q : q - n1;
                    if (!((whole >> (q - qb)) & 1u)) {
                        const int io_j = __shfl_sync(kFullMask, static_cast<int>(io_l), q - qb);
                        if (dir_sector_of(bucket, io_j, Bmask, a.shift) != s) continue;   // another unit's ray
                    }
                    cast(j);
                }
            }
        }
    }
    if (a.replay_count && replays) atomicAdd(reinterpret_cast<unsigned long long*>(a.replay_count), static_cast<unsigned long long>(replays));
#if MCL_DIR_DIAG
    if (tid == 0) {
        unsigned long long dg_g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dg_g1));
        unsigned long long* d = g_dir_diag[blockIdx.x];
        d[0] = clock64() - dg_t0;
        d[1] = dg_units;
        d[2] = dg_pieces;
        d[3] = dg_wins;
        d[4] = dg_sched;
        d[5] = dg_stage;
        d[6] = dg_g0;
        d[7] = dg_g1;
    }
#endif
}

struct WeightStepsArgs {
    const int* plan;
    const uint8_t* steps_sorted;   // [R][stride]
    const int32_t* perm;
    const double* slice;           // [R][M+1]
    double* w_raw;                 // [N]
    uint8_t* steps;                // [N*R] particle-major copy (get_ranges) or nullptr
    int64_t lo, cnt, stride;
    int64_t nfil;                  // particles per filter in pool mode (slot / nfil is the filter), >= cnt otherwise
    int R, tw;
    double inv_squash;
};

// w = pow(prod_j table(obs_j, range_ij), 1/squash): entries multiplied in beam order like the
// reference's loop (:564-579).  One thread = four consecutive sorted slots: the step bytes arrive
// as 32-bit words (128 B per warp and beam), kBatch beams in flight, four independent products.
constexpr int kWeightThreads = 256;
template <bool POOL>   // POOL: slots of several filters (slot / nfil is the filter, each with its own table slice)
__global__ void __launch_bounds__(kWeightThreads) k_weight_steps(WeightStepsArgs a) {
    pdl_enter();
    if (a.plan[kPlanMode] != 1) return;
    const int64_t pos = (static_cast<int64_t>(blockIdx.x) * kWeightThreads + threadIdx.x) * 4;
    if (pos >= a.cnt) return;
    constexpr int kBatch = 10;
    double acc[4] = {1.0, 1.0, 1.0, 1.0};
    int fil[4] = {0, 0, 0, 0};    // filter of each slot: its rows of the table slice
    int foff[4] = {0, 0, 0, 0};
    if (POOL) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            fil[e] = static_cast<int>(min(pos + e, a.cnt - 1) / a.nfil);
            foff[e] = fil[e] * a.R * a.tw;
        }
    }
    const uint32_t* sp = reinterpret_cast<const uint32_t*>(a.steps_sorted + pos);   // stride and pos are multiples of 4
    const size_t wstride = static_cast<size_t>(a.stride) / 4;
    const double* row = a.slice;
    int j = 0;
    for (; j + kBatch <= a.R; j += kBatch) {
        uint32_t w[kBatch];
#pragma unroll
        for (int q = 0; q < kBatch; ++q) w[q] = __ldcs(sp + q * wstride);   // streamed: the table rows stay in L1
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
#pragma unroll
            for (int e = 0; e < 4; ++e)   // slots beyond cnt hold zeros or stale (valid) steps
                acc[e] = __dmul_rn(acc[e], __ldg(row + foff[e] + q * a.tw + ((w[q] >> (8 * e)) & 255u)));
        }
        sp += kBatch * wstride;
        row += kBatch * a.tw;
    }
    for (; j < a.R; ++j) {
        const uint32_t w = __ldcs(sp);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = __dmul_rn(acc[e], __ldg(row + foff[e] + ((w >> (8 * e)) & 255u)));
        sp += wstride;
        row += a.tw;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (pos + e < a.cnt) {
            const int64_t i = a.lo + (POOL ? fil[e] * a.nfil : int64_t{0}) + a.perm[a.lo + pos + e];
            a.w_raw[i] = squash_pow(acc[e], a.inv_squash);
            if (a.steps)   // diagnostics (mcl_get_ranges): particle-major copy of the step indices
                for (int jj = 0; jj < a.R; ++jj) a.steps[i * a.R + jj] = a.steps_sorted[static_cast<int64_t>(jj) * a.stride + pos + e];
        }
    }
}


// The same product with the table slice staged in shared memory (one filter: all slots share the R x (M+1)
// slice, 100 KB at 60 beams x 208).  The lookup is then 4 instructions (byte -> byte offset, LDS.64 with the
// row's base as a uniform offset, DMUL) instead of ~7 through L1 with 64-bit addresses, and costs no L1 tag
// lookups; the next batch of step words is in flight while the current one is multiplied.  Persistent CTAs,
// NT threads each, as many per SM as the table allows.  Results are bit-identical to k_weight_steps<false>.
template <int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) k_weight_steps_sm(WeightStepsArgs a) {
    pdl_enter();
    extern __shared__ __align__(16) double tab[];   // [R][tw]
    if (a.plan[kPlanMode] != 1) return;
    // one filter: persistent CTAs stride over all slots with the filter's slice.  POOL of filters (a.nfil < a.cnt; nfil a
    // multiple of 4): a CTA takes whole filters -- their sorted slots are contiguous --, staging each filter's own slice
    const bool pool = a.nfil < a.cnt;
    const int nf = pool ? static_cast<int>(a.cnt / a.nfil) : 1;
    uint32_t tab_s;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(tab_s) : "l"(tab));
    constexpr int kBatch = 10;
    const size_t wstride = static_cast<size_t>(a.stride) / 4;
    const uint32_t row_bytes = static_cast<uint32_t>(a.tw) * 8u;
    for (int f = pool ? blockIdx.x : 0; f < nf; f += pool ? gridDim.x : 1) {
        if (pool && f != static_cast<int>(blockIdx.x)) __syncthreads();   // every thread has left the previous filter's table
        {
            const int n2 = (a.R * a.tw) / 2;            // (R * tw is even or the tail entry is copied alone)
            const double* slice = a.slice + static_cast<size_t>(f) * a.R * a.tw;
            const bool vec = (reinterpret_cast<uintptr_t>(slice) & 15u) == 0;
            if (vec) {
                const double2* src = reinterpret_cast<const double2*>(slice);
                double2* dst = reinterpret_cast<double2*>(tab);
                for (int i = threadIdx.x; i < n2; i += NT) dst[i] = __ldg(src + i);
                if (threadIdx.x == 0 && ((a.R * a.tw) & 1)) tab[a.R * a.tw - 1] = slice[a.R * a.tw - 1];
            } else {
                for (int i = threadIdx.x; i < a.R * a.tw; i += NT) tab[i] = __ldg(slice + i);
            }
        }
        __syncthreads();
        const int64_t p_begin = pool ? static_cast<int64_t>(f) * a.nfil : 0;
        const int64_t p_end = pool ? p_begin + a.nfil : a.cnt;
        const int64_t p_step = (pool ? int64_t{1} : static_cast<int64_t>(gridDim.x)) * NT * 4;
        for (int64_t pos = p_begin + ((pool ? int64_t{0} : static_cast<int64_t>(blockIdx.x) * NT) + threadIdx.x) * 4; pos < p_end; pos += p_step) {
            const uint32_t* sp = reinterpret_cast<const uint32_t*>(a.steps_sorted + pos);   // stride and pos are multiples of 4
            double acc[4] = {1.0, 1.0, 1.0, 1.0};
            uint32_t row_s = tab_s;
            uint32_t w[kBatch], wn[kBatch];
            int j = 0;
            if (kBatch <= a.R) {
#pragma unroll
                for (int q = 0; q < kBatch; ++q) w[q] = __ldcs(sp + q * wstride);
            }
            for (; j + kBatch <= a.R; j += kBatch) {
                sp += kBatch * wstride;
                const bool more = j + 2 * kBatch <= a.R;
                if (more) {
#pragma unroll
                    for (int q = 0; q < kBatch; ++q) wn[q] = __ldcs(sp + q * wstride);   // the next batch, in flight during the products
                }
#pragma unroll
                for (int q = 0; q < kBatch; ++q) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {   // slots beyond the end hold zeros or stale (valid) steps
                        const uint32_t off = e == 0 ? (w[q] << 3) & 0x7f8u : (w[q] >> (8 * e - 3)) & 0x7f8u;
                        double v;
                        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(row_s + static_cast<uint32_t>(q) * row_bytes + off));
                        acc[e] = __dmul_rn(acc[e], v);
                    }
                }
                row_s += kBatch * row_bytes;
                if (more) {
#pragma unroll
                    for (int q = 0; q < kBatch; ++q) w[q] = wn[q];
                }
            }
            for (; j < a.R; ++j) {
                const uint32_t ww = __ldcs(sp);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t off = e == 0 ? (ww << 3) & 0x7f8u : (ww >> (8 * e - 3)) & 0x7f8u;
                    double v;
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(row_s + off));
                    acc[e] = __dmul_rn(acc[e], v);
                }
                sp += wstride;
                row_s += row_bytes;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (pos + e < p_end) {
                    const int64_t i = a.lo + p_begin + a.perm[a.lo + pos + e];
                    a.w_raw[i] = squash_pow(acc[e], a.inv_squash);
                    if (a.steps)   // diagnostics (mcl_get_ranges): particle-major copy of the step indices
                        for (int jj = 0; jj < a.R; ++jj) a.steps[i * a.R + jj] = a.steps_sorted[static_cast<int64_t>(jj) * a.stride + pos + e];
                }
            }
        }
    }
}

}  // namespace mclb200
