// csrc/mcl_b200.cu -- C ABI (include/mcl_b200.h) over the sm_100a kernels in kernels.cuh.
//
// Host side of the drop-in boundary: owns the device buffers of one ParticleFilter (or a
// batch of independent ones), uploads map / table / beams, and sequences the kernels of one
// MCL update on a CUDA stream.  There is deliberately no CPU implementation behind these
// entry points: without a CUDA device mcl_create fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mcl_b200.h"
#include "kernels.cuh"

using namespace mclb200;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(MCL_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

template <class T>
cudaError_t dalloc(T** p, size_t n) {
    return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T));
}

constexpr size_t kWindowBudget = 226 * 1024;   // shared memory for the skip-map window (227 KB/CTA - 1 KB reserved)

}  // namespace

struct mcl_ctx {
    mcl_params prm{};
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    int F = 1;
    int64_t N = 0;
    int R = 0, M = 0;
    bool have_map = false, have_beams = false;
    double res = 0, ox = 0, oy = 0, oyaw = 0;
    SkipMap skip;
    std::vector<double> table;       // (M+1)^2 column-major, host copy
    MapDev map{};
    BeamDev beams{};
    int8_t* d_grid = nullptr;
    uint8_t* d_v8 = nullptr;
    uint8_t* d_v4 = nullptr;
    int32_t* d_free = nullptr;
    double* d_tabT = nullptr;
    int32_t* d_step2idx = nullptr;
    // state
    double* d_px[2] = {nullptr, nullptr};
    double* d_py[2] = {nullptr, nullptr};
    double* d_pt[2] = {nullptr, nullptr};
    double4* d_pose4[2] = {nullptr, nullptr};   // packed (x, y, theta, 0) copy of the state written by k_resample_motion
    bool pose4_ok[2] = {false, false};          // the packed copy of that buffer matches the SoA arrays
    int cur = 0;
    double* d_wraw = nullptr;
    double* d_wn = nullptr;
    double* d_cdf = nullptr;
    int32_t* d_idx = nullptr;
    uint8_t* d_steps = nullptr;
    bool keep_ranges = false;
    double* d_u = nullptr;
    double* d_z = nullptr;
    double* d_action = nullptr;
    float* d_obs = nullptr;
    double* d_slice = nullptr;
    size_t slice_elems = 0;
    // exact-sum workspaces
    int T = 0, C = 0;
    double* d_tile_sum = nullptr;
    StepFn* d_chunk_fn = nullptr;
    StepFn* d_opq_pre = nullptr;
    int* d_opq_idx = nullptr;
    int* d_tile_opq = nullptr;
    int64_t* d_tile_elem = nullptr;
    int* d_list_chunk = nullptr;
    StepFn* d_list_fn = nullptr;
    double* d_anchors = nullptr;
    double* d_anchor_val = nullptr;
    double* d_tile_start = nullptr;
    double* d_coarse = nullptr;      // [F][coarse_n] coarse level of the CDF search (exact values at segment ends)
    int coarse_n = 0, coarse_shift = 0;
    double* d_S1 = nullptr;
    double* d_S2 = nullptr;
    double* d_scratch_total = nullptr;
    // pose
    int norm_blocks = 1;
    double* d_partial = nullptr;
    double* d_pose = nullptr;
    double* d_centre = nullptr;
    int64_t* d_replays = nullptr;
    // heading sort (coherent warps in the ray kernel)
    int B = 0;
    int* d_hist = nullptr;      // [F][2B]: histogram | scatter cursors
    int32_t* d_perm = nullptr;
    bool sort_enabled = true;
    // directional ray stage (dir_kernels.cuh): one large filter, heading sort fine enough
    bool dir_ready = false;
    bool dir_pool = false;            // batch of filters on a map that fits one window: the stage runs over the pool
    int dir_B = 0;                    // heading buckets of the sector arithmetic (the sort's B for one filter)
    int ray_mode = 0;                 // 0 auto, 1 isotropic kernel only, 2 directional forced
    DirSector sectors[kDirSectors];
    uint8_t* d_dirmaps = nullptr;     // [S][PH*PW]
    DirSector* d_sectors = nullptr;
    int beam_io[kMaxBeams] = {};
    int* d_sec_tab = nullptr;         // [S+1] first unit | [S] first chunk, per sector
    DirReplayCtx* d_replay_ctx = nullptr;   // [2]: one per state buffer
    DirRec* d_rec = nullptr;          // [2][N]: slot order | heading-sorted order
    int* d_plan = nullptr;
    uint8_t* d_steps_sorted = nullptr;   // [R][stride]
    int64_t dir_stride = 0;
    int dir_box = 0;
    size_t dir_smem = 0;
    // particle shard: this context computes output slots [lo, lo+cnt) of the filter
    int64_t lo = 0, cnt = 0;
    bool local_pending = false;
    // peer-to-peer sharding
    bool p2p = false;
    int world = 1, rank = 0;
    const double** d_peer_tab = nullptr;     // device: [buf 2][array 3][world] pointers
    std::vector<void*> ipc_opened;
    double* d_partials = nullptr;            // [world][4] pose partial sums (rank's own at [rank])
    unsigned int* d_done = nullptr;          // [F] block-completion counters of k_normalize_pose
    // what d_tile_sum currently holds: 0 nothing usable, 1 tile sums of w_norm, 2 tile sums of
    // w_raw with w_norm = w_raw / S1 (rescaled on the fly for the approximate prefix)
    int tile_state = 0;
    // pinned staging for the host-facing update
    // pinned staging of the host-facing update: [F][3] doubles of action followed by [F][R] floats of
    // scan in ONE buffer (one H2D copy per update); d_action / d_obs alias the device twin
    double* h_action = nullptr;
    float* h_obs = nullptr;
    double* h_pose = nullptr;
    double* d_pose_mapped = nullptr;   // device alias of h_pose (mapped pinned memory): the pose kernel writes it directly
    uint64_t update_no = 0, init_no = 0;
    unsigned long long* d_update_no = nullptr;   // device twin of update_no, read by k_resample_motion
    // CUDA graphs of the steady-state host-facing update, one per state-buffer parity
    bool graphs_enabled = true;
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    int64_t graph_launches = 0;
    int64_t launches = 0;
    bool profiling = false;
    bool ev_valid = false;            // the stage events of a whole update have been recorded
    cudaEvent_t ev[7] = {};
    mcl_stage_ms last_ms{};
    bool cdf_valid = false;
};

namespace {

// the captured update graphs bake in buffer pointers, launch shapes and the stream's kernels:
// every setter that can change one of them drops the graphs (they are re-captured on demand)
void drop_graphs(mcl_ctx* c) {
    for (auto& g : c->gexec) {
        if (g) cudaGraphExecDestroy(g);
        g = nullptr;
    }
}

ExactArgs exact_args(mcl_ctx* c, const double* src, const double* div, const double* approx_div, double* total,
                     double* out, int force_one) {
    ExactArgs a{};
    a.src = src;
    a.div = div;
    a.approx_div = approx_div;
    a.N = c->N;
    a.T = c->T;
    a.C = c->C;
    a.tile_sum = c->d_tile_sum;
    a.chunk_fn = c->d_chunk_fn;
    a.opq_pre = c->d_opq_pre;
    a.opq_idx = c->d_opq_idx;
    a.tile_opq = c->d_tile_opq;
    a.tile_elem = c->d_tile_elem;
    a.list_chunk = c->d_list_chunk;
    a.list_fn = c->d_list_fn;
    a.anchors = c->d_anchors;
    a.anchor_val = c->d_anchor_val;
    a.tile_start = c->d_tile_start;
    a.total = total;
    a.out = out;
    a.force_last_one = force_one;
    // the pass that emits the CDF also publishes the coarse level of its search
    a.coarse = (out && c->coarse_n > 0) ? c->d_coarse : nullptr;
    a.coarse_m = a.coarse ? (1 << c->coarse_shift) / kChunk : 0;
    a.coarse_n = c->coarse_n;
    return a;
}

// sequential-order sum (and optionally prefix sums) of src[k] (/ div), see exact_sum.cuh
int run_exact(mcl_ctx* c, const double* src, const double* div, double* total, double* out, int force_one,
              bool need_tile_sums, const double* approx_div = nullptr) {
    ExactArgs a = exact_args(c, src, div, approx_div, total, out, force_one);
    const dim3 gt(c->T, c->F);
    if (c->T == 1) {   // a filter of one tile: the whole pass is one CTA per filter
        k_exact_single<<<gt, kTileChunks, 0, c->stream>>>(a);
        c->launches++;
        CK(cudaGetLastError());
        return MCL_OK;
    }
    if (need_tile_sums) {
        k_tile_sums<<<gt, kTileChunks, 0, c->stream>>>(a);
        c->launches++;
    }
    k_exact_chunks<<<gt, kTileChunks, 0, c->stream>>>(a);
    k_exact_walk<<<dim3(1, c->F), kWalkThreads, 0, c->stream>>>(a);
    c->launches += 2;
    if (out) {
        k_exact_emit<<<gt, kTileChunks, 0, c->stream>>>(a);
        c->launches++;
    }
    CK(cudaGetLastError());
    return MCL_OK;
}

// the reference's conversion of a returned range back to a table index (:556-561, :571-574)
void build_step2idx(const mcl_ctx* c, std::vector<int32_t>& out) {
    const int M = c->M;
    out.resize(M + 1);
    for (int r = 0; r <= M; ++r) {
        const double rd = (r >= M) ? c->prm.max_range : static_cast<double>(r) * c->res;
        const float range_f = static_cast<float>(rd);                 // cast_ray returns float
        float px = static_cast<float>(static_cast<double>(range_f) / c->res);   // ranges_px[i] = ranges_[i] / res
        if (px > static_cast<float>(M)) px = static_cast<float>(M);
        int idx = static_cast<int>(std::round(px));
        out[r] = std::max(0, std::min(idx, M));
    }
}

// precompute_sensor_model (:233-292), column-major (r + d*(M+1)).  Host code is compiled
// with -ffp-contract=off so the expression below rounds like the reference's build.
void build_sensor_table(const mcl_ctx* c, std::vector<double>& tab) {
    const int M = c->M, tw = M + 1;
    const double zs = c->prm.z_short, zm = c->prm.z_max, zr = c->prm.z_rand, zh = c->prm.z_hit, sg = c->prm.sigma_hit;
    tab.assign(static_cast<size_t>(tw) * tw, 0.0);
    for (int d = 0; d < tw; ++d) {
        double* col = &tab[static_cast<size_t>(d) * tw];
        double norm = 0.0;
        for (int r = 0; r < tw; ++r) {
            const double z = static_cast<double>(r - d);
            double prob = 0.0;
            prob += zh * std::exp(-(z * z) / (2.0 * sg * sg)) / (sg * std::sqrt(2.0 * M_PI));
            if (r < d) prob += 2.0 * zs * (d - r) / static_cast<double>(d);
            if (r == M) prob += zm;
            if (r < M) prob += zr * 1.0 / static_cast<double>(M);
            norm += prob;
            col[r] = prob;
        }
        if (norm > 0)
            for (int r = 0; r < tw; ++r) col[r] /= norm;
    }
}

int upload_table(mcl_ctx* c) {
    drop_graphs(c);
    const int tw = c->M + 1;
    std::vector<double> tabT(static_cast<size_t>(tw) * tw);
    for (int d = 0; d < tw; ++d)
        for (int r = 0; r < tw; ++r) tabT[static_cast<size_t>(r) * tw + d] = c->table[static_cast<size_t>(d) * tw + r];
    // table(obs=r, range=d) lives at r + d*tw; tabT[obs*tw + range]
    if (c->d_tabT) cudaFree(c->d_tabT);
    c->d_tabT = nullptr;
    CK(dalloc(&c->d_tabT, tabT.size()));
    CK(cudaMemcpy(c->d_tabT, tabT.data(), tabT.size() * sizeof(double), cudaMemcpyHostToDevice));
    return MCL_OK;
}

int ensure_slice(mcl_ctx* c) {
    drop_graphs(c);
    if (!c->have_map || !c->have_beams) return MCL_OK;
    const size_t need = static_cast<size_t>(c->F) * c->R * (c->M + 1);
    if (need > c->slice_elems) {
        if (c->d_slice) cudaFree(c->d_slice);
        c->d_slice = nullptr;
        CK(dalloc(&c->d_slice, need));
        c->slice_elems = need;
    }
    return MCL_OK;
}

constexpr size_t kDirWindowBudget = 112 * 1024;   // one sector window (half of shared memory: room to double-buffer)

void free_dir(mcl_ctx* c) {
    drop_graphs(c);
    for (void* p : {static_cast<void*>(c->d_dirmaps), static_cast<void*>(c->d_sectors), static_cast<void*>(c->d_sec_tab), static_cast<void*>(c->d_replay_ctx),
                    static_cast<void*>(c->d_rec),
                    static_cast<void*>(c->d_plan), static_cast<void*>(c->d_steps_sorted)})
        if (p) cudaFree(p);
    c->d_dirmaps = nullptr;
    c->d_sectors = nullptr;
    c->d_sec_tab = nullptr;
    c->d_replay_ctx = nullptr;
    c->d_rec = nullptr;
    c->d_plan = nullptr;
    c->d_steps_sorted = nullptr;
    c->dir_ready = false;
}

// the exact-replay context of the directional ray kernel, one per state buffer
int upload_replay_ctx(mcl_ctx* c) {
    if (!c->d_replay_ctx) return MCL_OK;
    DirReplayCtx h[2];
    for (int b = 0; b < 2; ++b) {
        h[b].grid = RefGrid{c->d_grid, c->map.W, c->map.H, c->res, c->ox, c->oy};
        h[b].px = c->d_px[b];
        h[b].py = c->d_py[b];
        h[b].pt = c->d_pt[b];
        h[b].perm = c->d_perm + c->lo;
        h[b].lo = c->lo;
        h[b].nfil = c->dir_pool ? c->N : (int64_t{1} << 40);
        for (int j = 0; j < kMaxBeams; ++j) h[b].angle[j] = j < c->R ? c->beams.angle[j] : 0.0f;
    }
    CK(cudaMemcpy(c->d_replay_ctx, h, sizeof(h), cudaMemcpyHostToDevice));
    return MCL_OK;
}

// Directional ray stage: eligible for ONE large filter whose heading sort has at least
// kDirMinBuckets buckets.  Builds the sector maps on the device (k_build_dir_maps) and the
// per-update work buffers.  Called whenever the map or the beam table changes.
int ensure_dir(mcl_ctx* c, bool map_changed) {
    if (!c->have_map || !c->have_beams) return MCL_OK;
    const size_t ncell = static_cast<size_t>(c->skip.PW) * c->skip.PH;
    // one filter: windows around the cloud; a batch: only when the whole padded map is one window
    const bool pool = c->F > 1 && ncell <= kDirWindowBudget && static_cast<int64_t>(c->F) * c->N < (int64_t{1} << 31) - 2048;
    const bool eligible = ((c->F == 1 && c->B >= kDirMinBuckets) || pool) && c->skip.PW <= 32767 && c->skip.PH <= 32767 && c->R <= 127;
    if (!eligible) {
        free_dir(c);
        return MCL_OK;
    }
    CK(cudaStreamSynchronize(c->stream));
    const int64_t pool_n = static_cast<int64_t>(c->F) * c->N;   // slots of the stage (== N for one filter)
    if (map_changed || !c->d_dirmaps) {
        free_dir(c);
        make_dir_sectors(c->M, c->sectors);
        c->dir_box = dir_choose_box(c->sectors, kDirWindowBudget);
        if (c->dir_box == 0 || ncell * kDirSectors > (size_t{8} << 30)) return MCL_OK;   // stays on the isotropic kernel
        size_t smem = 0;
        for (int s = 0; s < kDirSectors; ++s) {
            const DirSector& sc = c->sectors[s];
            smem = std::max(smem, static_cast<size_t>((c->dir_box + sc.exh - sc.exl + 30) & ~15) *
                                      static_cast<size_t>(c->dir_box + sc.eyh - sc.eyl));
        }
        c->dir_smem = pool ? ncell : smem;
        std::vector<float> gap;
        build_gap_map(c->skip, gap);
        float* d_gap = nullptr;
        CK(dalloc(&d_gap, ncell));
        CK(cudaMemcpy(d_gap, gap.data(), ncell * sizeof(float), cudaMemcpyHostToDevice));
        CK(dalloc(&c->d_sectors, static_cast<size_t>(kDirSectors)));
        CK(cudaMemcpy(c->d_sectors, c->sectors, sizeof(c->sectors), cudaMemcpyHostToDevice));
        CK(dalloc(&c->d_dirmaps, ncell * kDirSectors));
        DirBuildArgs ba{c->d_v8, d_gap, c->d_sectors, c->d_dirmaps, c->skip.PW, c->skip.PH};
        k_build_dir_maps<<<dim3(static_cast<unsigned>((ncell + 255) / 256), kDirSectors), 256, 0, c->stream>>>(ba);
        c->launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(d_gap);
        CK(dalloc(&c->d_plan, static_cast<size_t>(kPlanInts)));
        CK(cudaMemset(c->d_plan, 0, sizeof(int) * kPlanInts));
        CK(dalloc(&c->d_rec, static_cast<size_t>(2 * pool_n)));
    }
    c->dir_pool = pool;
    c->dir_B = pool ? kMaxBuckets : c->B;
    // per-beam-table buffers
    for (void* p : {static_cast<void*>(c->d_sec_tab), static_cast<void*>(c->d_steps_sorted), static_cast<void*>(c->d_replay_ctx)})
        if (p) cudaFree(p);
    c->d_sec_tab = nullptr;
    c->d_steps_sorted = nullptr;
    c->d_replay_ctx = nullptr;
    for (int j = 0; j < c->R; ++j) c->beam_io[j] = dir_beam_offset(c->beams.angle[j], c->dir_B);
    CK(dalloc(&c->d_sec_tab, static_cast<size_t>(2 * kDirSectors + 2)));
    CK(dalloc(&c->d_replay_ctx, size_t{2}));
    const int64_t nchunks = (pool_n + kDirThreads - 1) / kDirThreads;
    c->dir_stride = nchunks * kDirThreads;
    CK(dalloc(&c->d_steps_sorted, static_cast<size_t>(c->R) * c->dir_stride));
    CK(cudaMemset(c->d_steps_sorted, 0, static_cast<size_t>(c->R) * c->dir_stride));   // slots beyond the shard stay valid steps
    if (dir_ray_smem(static_cast<int>((c->dir_smem + 15) & ~size_t{15})) > kWindowBudget) return MCL_OK;   // stays on the isotropic kernel
    c->dir_ready = true;
    return upload_replay_ctx(c);
}

int check_filter(const mcl_ctx* c, int filter, bool allow_all) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    if (filter == -1 && allow_all) return MCL_OK;
    if (filter < 0 || filter >= c->F) return fail(MCL_ERR_INVALID, "filter %d out of range [0,%d)", filter, c->F);
    return MCL_OK;
}

int launch_pose(mcl_ctx* c, const double* w, const double* total, double* wn_out, int buf, bool count_update) {
    NormArgs na{};
    na.N = c->N;
    na.w_raw = w;
    na.total = total;
    na.wn = wn_out;
    na.px = c->d_px[buf];
    na.py = c->d_py[buf];
    na.pt = c->d_pt[buf];
    na.partial = c->d_partial;
    na.nblk = c->norm_blocks;
    na.done = c->d_done;
    na.pose_out = c->d_pose;
    na.pose_host = c->d_pose_mapped;
    na.update_no = count_update ? c->d_update_no : nullptr;
    k_normalize_pose<<<dim3(c->norm_blocks, c->F), kNormThreads, 0, c->stream>>>(na);
    c->launches += 1;
    CK(cudaGetLastError());
    return MCL_OK;
}

// discrete_distribution(weights_): sum, normalise, partial_sum (random.tcc:2657-2678) of the
// current normalised weights.  The approximate tile sums the exact kernels need are reused from
// the last weight-sum pass whenever the weights have not been touched since.
int build_cdf(mcl_ctx* c) {
    const double* approx = c->tile_state == 2 ? c->d_S1 : nullptr;
    int rc = run_exact(c, c->d_wn, nullptr, c->d_S2, nullptr, 0, c->tile_state == 0, approx);
    if (rc) return rc;
    if (c->tile_state == 0) c->tile_state = 1;
    rc = run_exact(c, c->d_wn, c->d_S2, c->d_scratch_total, c->d_cdf, 1, false, approx);
    if (rc) return rc;
    c->cdf_valid = true;
    return MCL_OK;
}

// First half of MCL(): CDF of the current weights, then resample / motion / ray cast / weight
// for this context's slots [lo, lo+cnt) into the destination buffers.
int update_local(mcl_ctx* c, const double* action_dev, const float* obs_dev, const double* u_dev, const double* z_dev) {
    if (!c->have_map) return fail(MCL_ERR_NO_MAP, "mcl_set_map has not been called");
    if (!c->have_beams) return fail(MCL_ERR_INVALID, "mcl_set_beam_angles has not been called");
    if (c->local_pending) return fail(MCL_ERR_INVALID, "mcl_update_finish must follow mcl_update_local");
    const int src = c->cur, dst = c->cur ^ 1;
    cudaStream_t s = c->stream;
    if (c->profiling) CK(cudaEventRecord(c->ev[0], s));

    ObsArgs oa{};
    oa.obs = obs_dev;
    oa.tabT = c->d_tabT;
    oa.step2idx = c->d_step2idx;
    oa.slice = c->d_slice;
    oa.R = c->R;
    oa.M = c->M;
    oa.res = c->res;
    // the same launch clears the per-update accumulators (cloud centre, heading histogram + cursors)
    oa.centre = c->d_centre;
    oa.hist = c->sort_enabled ? c->d_hist : nullptr;
    oa.nhist = 2 * c->B * c->F;
    k_prepare_obs<<<dim3(c->R, c->F), 256, 0, s>>>(oa);
    c->launches++;

    int rc = build_cdf(c);
    if (rc) return rc;
    if (c->profiling) CK(cudaEventRecord(c->ev[1], s));

    // a batch's pool mode is opt-in (ray mode 2): measured slower than the isotropic kernel on small maps
    const bool dir = c->dir_ready && c->sort_enabled && c->ray_mode != 1 && (!c->dir_pool || c->ray_mode == 2);
    MotionArgs ma{};
    ma.rec = dir ? c->d_rec : nullptr;
    ma.map = c->map;
    ma.B = c->dir_B;
    ma.N = c->N;
    ma.lo = c->lo;
    ma.cnt = c->cnt;
    ma.cdf = c->d_cdf;
    ma.sx = c->d_px[src];
    ma.sy = c->d_py[src];
    ma.st = c->d_pt[src];
    ma.dx = c->d_px[dst];
    ma.dy = c->d_py[dst];
    ma.dt = c->d_pt[dst];
    ma.idx_out = c->d_idx;
    ma.u = u_dev;
    ma.z = z_dev;
    // the packed source copy is usable if k_resample_motion wrote it (no set_particles / init since) and,
    // for a shard, if the other ranks' slices are reachable too (p2p; the all-gather mode only moves the SoA arrays)
    // Only the unsharded filter reads the packed copy.  For p2p shards the packed PEER reads measured
    // 20 % faster than three SoA reads (one NVLink transaction per pose) but were not bit-exact against
    // one GPU for a 1 M-particle filter on real 2- and 4-GPU runs (exact for 512 k particles;
    // scripts/check_sharded_equals_single.py), so they stay off until that is understood; the SoA peer
    // reads are verified exact at 2 and 4 GPUs.
    static const bool no_packed = std::getenv("MCL_NO_PACKED") != nullptr;   // debugging knob
    static const bool packed_peers = std::getenv("MCL_PACKED_PEERS") != nullptr;   // experiment
    const bool packed = !no_packed && c->pose4_ok[src] && ((c->p2p && packed_peers) || (!c->p2p && c->cnt == c->N));
    ma.spose4 = packed ? c->d_pose4[src] : nullptr;
    ma.dpose4 = c->d_pose4[dst];
    if (c->p2p) {
        ma.peer_pose4 = packed ? reinterpret_cast<const double4* const*>(c->d_peer_tab + static_cast<size_t>(6 + src) * c->world) : nullptr;
        ma.peer_x = c->d_peer_tab + static_cast<size_t>(src * 3 + 0) * c->world;
        ma.peer_y = c->d_peer_tab + static_cast<size_t>(src * 3 + 1) * c->world;
        ma.peer_t = c->d_peer_tab + static_cast<size_t>(src * 3 + 2) * c->world;
        ma.n_local = c->N / c->world;
    }
    ma.coarse = c->coarse_n > 0 ? c->d_coarse : nullptr;
    ma.nc = c->coarse_n;
    ma.cshift = c->coarse_shift;
    ma.action = action_dev;
    ma.disp_x = c->prm.motion_dispersion_x;
    ma.disp_y = c->prm.motion_dispersion_y;
    ma.disp_t = c->prm.motion_dispersion_theta;
    ma.seed = c->prm.seed;
    ma.update_no = c->d_update_no;
    ma.centre = c->d_centre;
    int mblocks = static_cast<int>((c->cnt + kMotionThreads - 1) / kMotionThreads);
    const size_t msmem = sizeof(double) * static_cast<size_t>(c->coarse_n);
    // a large table is staged once per SM by persistent blocks; a small one by every block
    if (msmem > 16 * 1024) mblocks = std::min(mblocks, std::max(1, c->num_sms / std::min(c->F, c->num_sms)));
    k_resample_motion<<<dim3(mblocks, c->F), kMotionThreads, msmem, s>>>(ma);
    c->launches++;
    c->pose4_ok[dst] = true;
    if (c->sort_enabled) {
        SortArgs sa{};
        sa.N = c->N;
        sa.lo = c->lo;
        sa.cnt = c->cnt;
        sa.pt = c->d_pt[dst];
        sa.hist = c->d_hist;
        sa.cursor = c->d_hist + static_cast<size_t>(c->B) * c->F;
        sa.perm = c->d_perm;
        sa.B = c->B;
        // a few fat blocks per filter: ~one per SM for a single big filter
        const int64_t per_filter = std::max<int64_t>(1, c->num_sms / std::min(c->F, c->num_sms));
        int64_t chunk = (c->cnt + per_filter - 1) / per_filter;
        chunk = std::max<int64_t>(kSortThreads, (chunk + kSortThreads - 1) / kSortThreads * kSortThreads);
        sa.chunk = chunk;
        const dim3 gs(static_cast<unsigned>((c->cnt + chunk - 1) / chunk), c->F);
        k_sort_hist<<<gs, kSortThreads, 0, s>>>(sa);
        k_sort_scatter<<<gs, kSortThreads, 0, s>>>(sa);
        c->launches += 2;
    }
    if (dir) {
        DirPrepArgs pa{};
        pa.map = c->map;
        pa.centre = c->d_centre;
        const int64_t slots = c->dir_pool ? static_cast<int64_t>(c->F) * c->N : c->cnt;
        pa.rec_in = c->d_rec;
        pa.rec = c->d_rec + static_cast<int64_t>(c->F) * c->N;
        pa.perm = c->d_perm + c->lo;
        pa.plan = c->d_plan;
        pa.cnt = slots;
        pa.nfil = c->dir_pool ? c->N : (int64_t{1} << 40);
        pa.box = c->dir_box;
        pa.whole = c->dir_pool ? 1 : 0;
        k_dir_gather<<<static_cast<unsigned>((slots + 255) / 256), 256, 0, s>>>(pa);
        DirPlanArgs la{};
        la.hist = c->d_hist;
        std::memcpy(la.io, c->beam_io, sizeof(la.io));
        la.plan = c->d_plan;
        la.sec_tab = c->d_sec_tab;
        la.cnt = slots;
        la.B = c->dir_B;
        la.R = c->R;
        la.force = c->dir_pool ? 2 : c->ray_mode;   // the pool has no cloud box to be outside of
        la.all_chunks = c->dir_pool ? 1 : 0;
        k_dir_plan<<<1, kPlanThreads, 0, s>>>(la);
        c->launches += 2;
    }
    if (c->profiling) CK(cudaEventRecord(c->ev[2], s));

    if (!(dir && c->dir_pool)) {   // the pool mode of the directional stage has no fallback to hand the update to
        RayArgs ra{};
        ra.map = c->map;
        ra.beams = c->beams;
        ra.N = c->N;
        ra.lo = c->lo;
        ra.cnt = c->cnt;
        ra.px = c->d_px[dst];
        ra.py = c->d_py[dst];
        ra.pt = c->d_pt[dst];
        ra.perm = c->sort_enabled ? c->d_perm : nullptr;
        ra.slice = c->d_slice;
        ra.w_raw = c->d_wraw;
        ra.steps = c->keep_ranges ? c->d_steps : nullptr;
        ra.centre = c->d_centre;
        ra.inv_squash = 1.0 / c->prm.squash_factor;
        ra.replay_count = c->d_replays;
        ra.plan = dir ? c->d_plan : nullptr;
        const size_t smem = static_cast<size_t>(c->map.wbits == 8 ? c->map.ww : c->map.ww / 2) * c->map.wh;
        // persistent blocks: one per SM when the window fills shared memory, a few otherwise
        const int per_sm = smem > 100 * 1024 ? 1 : 2;
        const int budget = std::max(1, (c->num_sms * per_sm) / std::min(c->F, c->num_sms * per_sm));
        const int rblocks = static_cast<int>(std::min<int64_t>((c->cnt + kRayThreads - 1) / kRayThreads, budget));
        // MAX_RANGE_PX of the usual map resolutions is baked into specialised instances
        // (0.05 m -> 239, 0.0504 m -> 238, 0.05796 m -> 207); anything else takes the generic one
        const dim3 rgrid(rblocks, c->F);
#define MCL_LAUNCH_RAY(WB, MCV) k_raycast_weight<WB, MCV><<<rgrid, kRayThreads, smem, s>>>(ra)
        if (c->map.wbits == 8) {
            switch (c->M) {
                case 207: MCL_LAUNCH_RAY(8, 207); break;
                case 238: MCL_LAUNCH_RAY(8, 238); break;
                case 239: MCL_LAUNCH_RAY(8, 239); break;
                default: MCL_LAUNCH_RAY(8, 0); break;
            }
        } else {
            switch (c->M) {
                case 207: MCL_LAUNCH_RAY(4, 207); break;
                case 238: MCL_LAUNCH_RAY(4, 238); break;
                case 239: MCL_LAUNCH_RAY(4, 239); break;
                default: MCL_LAUNCH_RAY(4, 0); break;
            }
        }
#undef MCL_LAUNCH_RAY
        c->launches++;
    }
    if (dir) {
        DirRayArgs da{};
        da.map = c->map;
        da.beams = c->beams;
        std::memcpy(da.io, c->beam_io, sizeof(da.io));
        da.sectors = c->d_sectors;
        da.dirmaps = c->d_dirmaps;
        const int64_t slots = c->dir_pool ? static_cast<int64_t>(c->F) * c->N : c->cnt;
        da.rec = c->d_rec + static_cast<int64_t>(c->F) * c->N;
        da.sec_tab = c->d_sec_tab;
        da.replay = c->d_replay_ctx + dst;
        da.plan = c->d_plan;
        da.centre = c->d_centre;
        da.cnt = slots;
        da.stride = c->dir_stride;
        da.steps_sorted = c->d_steps_sorted;
        da.replay_count = c->d_replays;
        da.B = c->dir_B;
        da.shift = 0;
        while ((c->dir_B >> da.shift) > kDirSectors) ++da.shift;
        da.box = c->dir_box;
        da.whole = c->dir_pool ? 1 : 0;
        da.win_bytes = static_cast<int>((c->dir_smem + 15) & ~size_t{15});
        const size_t dsmem = dir_ray_smem(da.win_bytes);
        const int dblocks = static_cast<int>(std::min<int64_t>(c->num_sms, (slots * c->R + kDirThreads - 1) / kDirThreads));
        switch (c->M) {
            case 207: k_raycast_dir<207><<<dblocks, kDirThreads, dsmem, s>>>(da); break;
            case 238: k_raycast_dir<238><<<dblocks, kDirThreads, dsmem, s>>>(da); break;
            case 239: k_raycast_dir<239><<<dblocks, kDirThreads, dsmem, s>>>(da); break;
            default: k_raycast_dir<0><<<dblocks, kDirThreads, dsmem, s>>>(da); break;
        }
        if (c->profiling) CK(cudaEventRecord(c->ev[5], s));
        WeightStepsArgs wa{};
        wa.plan = c->d_plan;
        wa.steps_sorted = c->d_steps_sorted;
        wa.perm = c->d_perm;
        wa.slice = c->d_slice;
        wa.w_raw = c->d_wraw;
        wa.steps = c->keep_ranges ? c->d_steps : nullptr;
        wa.lo = c->lo;
        wa.cnt = slots;
        wa.nfil = c->dir_pool ? c->N : (int64_t{1} << 40);
        wa.stride = c->dir_stride;
        wa.R = c->R;
        wa.tw = c->M + 1;
        wa.inv_squash = 1.0 / c->prm.squash_factor;
        const unsigned wblocks = static_cast<unsigned>((slots + 4 * kWeightThreads - 1) / (4 * kWeightThreads));
        if (c->dir_pool)
            k_weight_steps<true><<<wblocks, kWeightThreads, 0, s>>>(wa);
        else
            k_weight_steps<false><<<wblocks, kWeightThreads, 0, s>>>(wa);
        c->launches += 2;
    }
    if (!dir && c->profiling) CK(cudaEventRecord(c->ev[5], s));
    if (c->p2p) {
        // unnormalised pose sums of the rank's own slots (deterministic two-stage reduction)
        NormArgs na{};
        na.N = c->cnt;
        na.w_raw = c->d_wraw + c->lo;
        na.total = nullptr;
        na.wn = nullptr;
        na.px = c->d_px[dst] + c->lo;
        na.py = c->d_py[dst] + c->lo;
        na.pt = c->d_pt[dst] + c->lo;
        na.partial = c->d_partial;
        na.nblk = c->norm_blocks;
        k_normalize_pose<<<dim3(c->norm_blocks, 1), kNormThreads, 0, s>>>(na);
        k_sum_partials<<<1, 256, 0, s>>>(c->d_partial, c->norm_blocks, c->d_partials + 4 * c->rank);
        c->launches += 2;
    }
    if (c->profiling) CK(cudaEventRecord(c->ev[3], s));
    CK(cudaGetLastError());
    c->local_pending = true;
    return MCL_OK;
}

// Second half: sum_weights = accumulate(weights_); w /= sum (:679-686); particles_ = proposal
// (:689); expected_pose (:696-716) -- over ALL particles of the filter.
int update_finish(mcl_ctx* c) {
    if (!c->local_pending) return fail(MCL_ERR_INVALID, "mcl_update_finish without mcl_update_local");
    const int dst = c->cur ^ 1;
    if (c->profiling) CK(cudaEventRecord(c->ev[6], c->stream));
    int rc = run_exact(c, c->d_wraw, nullptr, c->d_S1, nullptr, 0, true);
    if (rc) return rc;
    c->tile_state = 2;   // d_tile_sum = tile sums of w_raw; w_norm = w_raw / S1 follows
    if (c->p2p) {
        // poses of other ranks are not local: normalise all weights, pose from the gathered partials
        const int64_t n = c->N;
        k_normalize_only<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->stream>>>(c->d_wraw, c->d_S1, c->d_wn, n);
        k_pose_from_partials<<<1, 32, 0, c->stream>>>(c->d_partials, c->world, c->d_S1, c->d_pose, c->d_pose_mapped, c->d_update_no);
        c->launches += 2;
    } else {
        rc = launch_pose(c, c->d_wraw, c->d_S1, c->d_wn, dst, true);
        if (rc) return rc;
    }
    if (c->profiling) {
        CK(cudaEventRecord(c->ev[4], c->stream));
        c->ev_valid = true;
    }
    CK(cudaGetLastError());
    c->cur = dst;
    c->update_no++;
    c->local_pending = false;
    return MCL_OK;
}

int update_device(mcl_ctx* c, const double* action_dev, const float* obs_dev, const double* u_dev, const double* z_dev) {
    int rc = update_local(c, action_dev, obs_dev, u_dev, z_dev);
    if (rc) return rc;
    return update_finish(c);
}

// per-stage device times of the last profiled update, from its CUDA events (the stream must have
// passed the last of them)
void read_stage_times(mcl_ctx* c) {
    float t = 0;
    cudaEventElapsedTime(&t, c->ev[0], c->ev[1]);
    c->last_ms.cdf = t;
    cudaEventElapsedTime(&t, c->ev[1], c->ev[2]);
    c->last_ms.resample_motion = t;
    cudaEventElapsedTime(&t, c->ev[2], c->ev[3]);
    c->last_ms.raycast_weight = t;
    cudaEventElapsedTime(&t, c->ev[6], c->ev[4]);
    c->last_ms.normalize_pose = t;
    cudaEventElapsedTime(&t, c->ev[0], c->ev[4]);
    c->last_ms.total = t;
    cudaEventElapsedTime(&t, c->ev[2], c->ev[5]);
    c->last_ms.ray_march = t;
    cudaEventElapsedTime(&t, c->ev[3], c->ev[6]);
    c->last_ms.exchange = t;
}

// One update from device-resident inputs, replayed as a CUDA graph when the update is in its steady
// state (no diagnostics, whole filter on this GPU, weights untouched since the last update): the
// ~18 launches become one.  The graph reads action and scan from the context's own staging
// buffers, so foreign device pointers are first copied there (264 bytes, device to device).
int update_steady(mcl_ctx* c, const double* action_dev, const float* obs_dev) {
    cudaStream_t s = c->stream;
    // (a captured graph reads the packed copy of the state: an update whose packed source is stale runs directly)
    const bool graph_ok = c->graphs_enabled && !c->profiling && !c->keep_ranges && !c->p2p && c->lo == 0 && c->cnt == c->N &&
                          c->tile_state == 2 && !c->local_pending && c->pose4_ok[c->cur];
    if (!graph_ok) return update_device(c, action_dev, obs_dev, nullptr, nullptr);
    if (action_dev != c->d_action)
        CK(cudaMemcpyAsync(c->d_action, action_dev, sizeof(double) * 3 * c->F, cudaMemcpyDeviceToDevice, s));
    if (obs_dev != c->d_obs)
        CK(cudaMemcpyAsync(c->d_obs, obs_dev, sizeof(float) * c->R * c->F, cudaMemcpyDeviceToDevice, s));
    if (c->gexec[c->cur]) {
        CK(cudaGraphLaunch(c->gexec[c->cur], s));
        c->cur ^= 1;             // what update_device does on the host side
        c->update_no++;
        c->launches += c->graph_launches;
        return MCL_OK;
    }
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        c->graphs_enabled = false;
        return update_device(c, c->d_action, c->d_obs, nullptr, nullptr);
    }
    const int parity = c->cur;
    const int64_t before = c->launches;
    int rc = update_device(c, c->d_action, c->d_obs, nullptr, nullptr);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s, &g);
    if (rc == MCL_OK && e == cudaSuccess && g && cudaGraphInstantiate(&c->gexec[parity], g, 0) == cudaSuccess) {
        c->graph_launches = c->launches - before;
        cudaGraphDestroy(g);
        CK(cudaGraphLaunch(c->gexec[parity], s));
        return MCL_OK;
    }
    // capture refused (e.g. an enclosing capture): run this update directly and stop trying
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    c->gexec[parity] = nullptr;
    c->graphs_enabled = false;
    if (rc != MCL_OK) return rc;
    c->cur ^= 1;                 // undo the host-side bookkeeping of the captured (never executed) update
    c->update_no--;
    c->launches = before;
    return update_device(c, c->d_action, c->d_obs, nullptr, nullptr);
}

}  // namespace

extern "C" {

void mcl_default_params(mcl_params* p) {
    if (!p) return;
    std::memset(p, 0, sizeof *p);
    p->max_particles = 2000;
    p->max_viz_particles = 60;
    p->angle_step = 18;
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
    p->seed = 0x9E3779B97F4A7C15ull;
    p->num_filters = 1;
}

const char* mcl_last_error(void) { return g_err.c_str(); }

const char* mcl_status_str(int s) {
    switch (s) {
        case MCL_OK: return "ok";
        case MCL_ERR_INVALID: return "invalid argument";
        case MCL_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
        case MCL_ERR_CUDA: return "CUDA error";
        case MCL_ERR_NO_MAP: return "map not set";
        case MCL_ERR_UNSUPPORTED: return "unsupported configuration";
        case MCL_ERR_NO_FREE_SPACE: return "no free space in map";
        default: return "unknown";
    }
}

int mcl_abi_version(void) { return MCL_B200_ABI_VERSION; }

int mcl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mcl_destroy(mcl_ctx* c);

static int create_buffers(mcl_ctx* c, const mcl_params* p, int device) {
    c->prm = *p;
    c->device = device;
    c->F = p->num_filters;
    c->N = p->max_particles;
    c->cnt = c->N;
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (auto& e : c->ev) CK(cudaEventCreate(&e));

    const size_t FN = static_cast<size_t>(c->F) * c->N;
    for (int b = 0; b < 2; ++b) {
        CK(dalloc(&c->d_px[b], FN));
        CK(dalloc(&c->d_py[b], FN));
        CK(dalloc(&c->d_pt[b], FN));
        CK(dalloc(&c->d_pose4[b], FN));
        CK(cudaMemset(c->d_pose4[b], 0, FN * sizeof(double4)));
        CK(cudaMemset(c->d_px[b], 0, FN * sizeof(double)));   // particles_ = Zero (:106)
        CK(cudaMemset(c->d_py[b], 0, FN * sizeof(double)));
        CK(cudaMemset(c->d_pt[b], 0, FN * sizeof(double)));
    }
    CK(dalloc(&c->d_wraw, FN));
    CK(dalloc(&c->d_wn, FN));
    CK(dalloc(&c->d_cdf, FN));
    CK(dalloc(&c->d_idx, FN));
    CK(cudaMemset(c->d_idx, 0, FN * sizeof(int32_t)));
    {   // weights_ = 1/N (:107)
        const int64_t n = static_cast<int64_t>(FN);
        k_fill<<<static_cast<unsigned>((n + 255) / 256), 256>>>(c->d_wn, n, 1.0 / static_cast<double>(c->N));
        k_fill<<<static_cast<unsigned>((n + 255) / 256), 256>>>(c->d_wraw, n, 1.0 / static_cast<double>(c->N));
        c->launches += 2;
    }
    c->T = static_cast<int>((c->N + kTile - 1) / kTile);
    c->C = c->T * kTileChunks;
    const size_t FT = static_cast<size_t>(c->F) * c->T, FC = static_cast<size_t>(c->F) * c->C;
    CK(dalloc(&c->d_tile_sum, FT));
    CK(dalloc(&c->d_chunk_fn, FC));
    CK(dalloc(&c->d_opq_pre, FC));
    CK(dalloc(&c->d_opq_idx, FC));
    CK(dalloc(&c->d_tile_opq, FT));
    CK(dalloc(&c->d_tile_elem, FT * 3));
    CK(dalloc(&c->d_list_chunk, FC));
    CK(dalloc(&c->d_list_fn, FC));
    CK(dalloc(&c->d_anchors, FC));
    CK(dalloc(&c->d_anchor_val, FC));
    CK(dalloc(&c->d_tile_start, FT));
    {   // coarse CDF level: segments of 64 particles, doubled until at most 16384 entries (128 KB of shared memory)
        c->coarse_shift = 6;
        while ((c->N >> c->coarse_shift) > 16384) ++c->coarse_shift;
        c->coarse_n = static_cast<int>(c->N >> c->coarse_shift);
        CK(dalloc(&c->d_coarse, static_cast<size_t>(c->F) * std::max(c->coarse_n, 1)));
    }
    CK(dalloc(&c->d_S1, static_cast<size_t>(c->F)));
    CK(dalloc(&c->d_S2, static_cast<size_t>(c->F)));
    CK(dalloc(&c->d_scratch_total, static_cast<size_t>(c->F)));
    c->norm_blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((c->N + kNormThreads * 4 - 1) / (kNormThreads * 4),
                                                                               std::max(1, 4 * c->num_sms / std::min(c->F, 4 * c->num_sms)))));
    CK(dalloc(&c->d_partial, static_cast<size_t>(c->F) * c->norm_blocks * 4));
    CK(dalloc(&c->d_pose, static_cast<size_t>(c->F) * 3));
    CK(cudaMemset(c->d_pose, 0, sizeof(double) * 3 * c->F));
    CK(dalloc(&c->d_centre, static_cast<size_t>(c->F) * 2));
    CK(dalloc(&c->d_done, static_cast<size_t>(c->F)));
    CK(cudaMemset(c->d_done, 0, sizeof(unsigned int) * c->F));
    {   // heading buckets: ~16 particles per bucket, power of two in [32, 4096]
        int B = 32;
        while (B < kMaxBuckets && static_cast<int64_t>(B) * 16 < c->N) B <<= 1;
        // a single filter of at least kDirMinParticles runs the directional ray stage, whose sector
        // arithmetic needs buckets narrower than the sector maps' margin
        if (c->F == 1 && c->N >= kDirMinParticles && B < kDirMinBuckets) B = kDirMinBuckets;
        c->B = B;
        CK(dalloc(&c->d_hist, static_cast<size_t>(2) * B * c->F));
        CK(dalloc(&c->d_perm, FN));
    }
    CK(dalloc(&c->d_update_no, size_t{1}));
    CK(cudaMemset(c->d_update_no, 0, sizeof(unsigned long long)));
    CK(dalloc(&c->d_replays, size_t{1}));
    CK(cudaMemset(c->d_replays, 0, sizeof(int64_t)));
    {
        const size_t in_bytes = sizeof(double) * 3 * c->F + sizeof(float) * kMaxBeams * c->F;
        void* d_in = nullptr;
        CK(cudaMalloc(&d_in, in_bytes));
        c->d_action = static_cast<double*>(d_in);
        c->d_obs = reinterpret_cast<float*>(c->d_action + 3 * c->F);
        CK(cudaMallocHost(reinterpret_cast<void**>(&c->h_action), in_bytes));
        c->h_obs = reinterpret_cast<float*>(c->h_action + 3 * c->F);
        CK(cudaHostAlloc(reinterpret_cast<void**>(&c->h_pose), sizeof(double) * 3 * c->F, cudaHostAllocMapped));
        std::memset(c->h_pose, 0, sizeof(double) * 3 * c->F);
        CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->d_pose_mapped), c->h_pose, 0));
    }
#define MCL_RAY_SMEM(WB, MCV) \
    CK(cudaFuncSetAttribute(k_raycast_weight<WB, MCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWindowBudget)))
    MCL_RAY_SMEM(8, 0);
    MCL_RAY_SMEM(8, 207);
    MCL_RAY_SMEM(8, 238);
    MCL_RAY_SMEM(8, 239);
    MCL_RAY_SMEM(4, 0);
    MCL_RAY_SMEM(4, 207);
    MCL_RAY_SMEM(4, 238);
    MCL_RAY_SMEM(4, 239);
#undef MCL_RAY_SMEM
    CK(cudaFuncSetAttribute(k_resample_motion, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * static_cast<int>(sizeof(double))));
    CK(cudaFuncSetAttribute(k_raycast_dir<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWindowBudget)));
    CK(cudaFuncSetAttribute(k_raycast_dir<207>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWindowBudget)));
    CK(cudaFuncSetAttribute(k_raycast_dir<238>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWindowBudget)));
    CK(cudaFuncSetAttribute(k_raycast_dir<239>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWindowBudget)));
    CK(cudaDeviceSynchronize());
    return MCL_OK;
}

int mcl_create(const mcl_params* p, int device, mcl_ctx** out) {
    if (!p || !out) return fail(MCL_ERR_INVALID, "null argument");
    *out = nullptr;
    if (p->max_particles < 1) return fail(MCL_ERR_INVALID, "max_particles must be >= 1");
    if (p->num_filters < 1) return fail(MCL_ERR_INVALID, "num_filters must be >= 1");
    if (!(p->squash_factor > 0) || !(p->max_range > 0)) return fail(MCL_ERR_INVALID, "squash_factor / max_range must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MCL_ERR_NO_DEVICE, "no CUDA device visible; the MCL update has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(MCL_ERR_INVALID, "device %d not in [0,%d)", device, ndev);
    CK(cudaSetDevice(device));
    auto* c = new mcl_ctx();
    const int rc = create_buffers(c, p, device);
    if (rc != MCL_OK) {
        const std::string keep = g_err;   // mcl_destroy must not clobber the reason
        mcl_destroy(c);
        g_err = keep;
        return rc;
    }
    *out = c;
    return MCL_OK;
}

int mcl_destroy(mcl_ctx* c) {
    if (!c) return MCL_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    void* ptrs[] = {c->d_pose4[0], c->d_pose4[1],
                    c->d_grid, c->d_v8, c->d_v4, c->d_free, c->d_tabT, c->d_step2idx, c->d_px[0], c->d_px[1], c->d_py[0],
                    c->d_py[1], c->d_pt[0], c->d_pt[1], c->d_wraw, c->d_wn, c->d_cdf, c->d_idx, c->d_steps, c->d_u, c->d_z,
                    c->d_action, c->d_slice, c->d_tile_sum, c->d_chunk_fn, c->d_opq_pre, c->d_opq_idx,
                    c->d_tile_opq, c->d_tile_elem, c->d_list_chunk, c->d_list_fn, c->d_anchors, c->d_anchor_val,
                    c->d_tile_start, c->d_coarse, c->d_S1, c->d_S2, c->d_scratch_total, c->d_partial, c->d_pose, c->d_centre, c->d_replays,
                    c->d_hist, c->d_perm, c->d_done};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    free_dir(c);
    drop_graphs(c);
    if (c->d_update_no) cudaFree(c->d_update_no);
    for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    if (c->d_peer_tab) cudaFree(c->d_peer_tab);
    if (c->d_partials) cudaFree(c->d_partials);
    if (c->h_action) cudaFreeHost(c->h_action);
    if (c->h_pose) cudaFreeHost(c->h_pose);
    for (auto& e : c->ev)
        if (e) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return MCL_OK;
}

int mcl_set_map(mcl_ctx* c, const int8_t* data, int width, int height, float resolution, double ox, double oy,
                double oyaw) {
    if (!c || !data) return fail(MCL_ERR_INVALID, "null argument");
    if (width <= 0 || height <= 0) return fail(MCL_ERR_INVALID, "bad grid size %dx%d", width, height);
    CK(cudaSetDevice(c->device));
    const double res = static_cast<double>(resolution);   // map_resolution_ = info.resolution (:191)
    if (!(res > 0.0)) return fail(MCL_ERR_INVALID, "Invalid map resolution: %.6f", res);   // :236-240
    const int M = static_cast<int>(c->prm.max_range / res);   // :195
    if (M < 1 || M > kMaxRangePxSupported)
        return fail(MCL_ERR_UNSUPPORTED, "MAX_RANGE_PX = %d outside [1,%d] (max_range %.3f / resolution %.6f)", M,
                    kMaxRangePxSupported, c->prm.max_range, res);
    if (!build_skip_map(data, width, height, c->skip)) return fail(MCL_ERR_UNSUPPORTED, "grid %dx%d too large", width, height);
    CK(cudaStreamSynchronize(c->stream));
    c->res = res;
    c->ox = ox;
    c->oy = oy;
    c->oyaw = oyaw;   // read by the reference (:192-193) but never used by the march (:628-629)
    c->M = M;
    for (void* p : {static_cast<void*>(c->d_grid), static_cast<void*>(c->d_v8), static_cast<void*>(c->d_v4),
                    static_cast<void*>(c->d_free), static_cast<void*>(c->d_step2idx)})
        if (p) cudaFree(p);
    c->d_grid = nullptr;
    c->d_v8 = c->d_v4 = nullptr;
    c->d_free = nullptr;
    c->d_step2idx = nullptr;
    const size_t cells = static_cast<size_t>(width) * height;
    CK(dalloc(&c->d_grid, cells));
    CK(cudaMemcpy(c->d_grid, data, cells, cudaMemcpyHostToDevice));
    CK(dalloc(&c->d_v8, c->skip.v8.size()));
    CK(cudaMemcpy(c->d_v8, c->skip.v8.data(), c->skip.v8.size(), cudaMemcpyHostToDevice));
    CK(dalloc(&c->d_v4, c->skip.v4.size()));
    CK(cudaMemcpy(c->d_v4, c->skip.v4.data(), c->skip.v4.size(), cudaMemcpyHostToDevice));
    CK(dalloc(&c->d_free, c->skip.free_cells.size()));
    if (!c->skip.free_cells.empty())
        CK(cudaMemcpy(c->d_free, c->skip.free_cells.data(), c->skip.free_cells.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    std::vector<int32_t> s2i;
    build_step2idx(c, s2i);
    CK(dalloc(&c->d_step2idx, s2i.size()));
    CK(cudaMemcpy(c->d_step2idx, s2i.data(), s2i.size() * sizeof(int32_t), cudaMemcpyHostToDevice));

    MapDev& m = c->map;
    m.grid = c->d_grid;
    m.v8 = c->d_v8;
    m.v4 = c->d_v4;
    m.W = width;
    m.H = height;
    m.PW = c->skip.PW;
    m.PH = c->skip.PH;
    m.res = res;
    m.ox = ox;
    m.oy = oy;
    m.M = M;
    // Shared-memory window of the skip map.  Preference: 8-bit codes (one LDS, longest skips)
    // if the whole P-grid fits or a window with at least kMinSpan cells of particle room
    // around two ray lengths does; else 4-bit codes (two cells per byte).
    constexpr int kMinSpan = 48;
    auto plan = [&](int bits, int* ww, int* wh) {
        const size_t cells = kWindowBudget * (bits == 8 ? 1 : 2);
        if (static_cast<size_t>(m.PW) * m.PH <= cells) {
            *ww = m.PW;
            *wh = m.PH;
            return true;
        }
        int w = static_cast<int>(std::sqrt(static_cast<double>(cells))) & ~31;
        w = std::min(w, m.PW);
        int h = std::min(m.PH, static_cast<int>(cells / w));
        if (h == m.PH) w = std::min(m.PW, static_cast<int>(cells / h) & ~31);
        const int need = 2 * (M + 2) + kMinSpan;
        if ((w < need && w < m.PW) || (h < need && h < m.PH)) return false;
        *ww = w;
        *wh = h;
        return true;
    };
    m.ww = m.wh = 0;
    m.wbits = 8;
    if (!plan(8, &m.ww, &m.wh)) {
        m.wbits = 4;
        if (!plan(4, &m.ww, &m.wh)) m.ww = m.wh = 0;   // window path disabled: global-memory march only
    }
    c->have_map = true;
    build_sensor_table(c, c->table);
    int rc = upload_table(c);
    if (rc) return rc;
    rc = ensure_slice(c);
    if (rc) return rc;
    return ensure_dir(c, true);
}

int mcl_max_range_px(const mcl_ctx* c) { return c ? c->M : 0; }

int mcl_get_sensor_table(const mcl_ctx* c, double* out) {
    if (!c || !out) return fail(MCL_ERR_INVALID, "null argument");
    if (!c->have_map) return fail(MCL_ERR_NO_MAP, "map not set");
    std::memcpy(out, c->table.data(), c->table.size() * sizeof(double));
    return MCL_OK;
}

int mcl_set_sensor_table(mcl_ctx* c, const double* tab, int tw) {
    if (!c || !tab) return fail(MCL_ERR_INVALID, "null argument");
    if (!c->have_map) return fail(MCL_ERR_NO_MAP, "map not set");
    if (tw != c->M + 1) return fail(MCL_ERR_INVALID, "table width %d != MAX_RANGE_PX+1 = %d", tw, c->M + 1);
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->table.assign(tab, tab + static_cast<size_t>(tw) * tw);
    return upload_table(c);
}

int mcl_set_beam_angles(mcl_ctx* c, const float* angles, int n) {
    if (!c || !angles) return fail(MCL_ERR_INVALID, "null argument");
    if (n < 1 || n > kMaxBeams) return fail(MCL_ERR_UNSUPPORTED, "%d beams outside [1,%d]", n, kMaxBeams);
    CK(cudaSetDevice(c->device));
    c->R = n;
    c->beams.R = n;
    for (int j = 0; j < n; ++j) {
        c->beams.angle[j] = angles[j];
        c->beams.cosa[j] = std::cos(static_cast<double>(angles[j]));
        c->beams.sina[j] = std::sin(static_cast<double>(angles[j]));
    }
    c->have_beams = true;
    if (c->d_steps) {
        cudaFree(c->d_steps);
        c->d_steps = nullptr;
    }
    if (c->keep_ranges) CK(dalloc(&c->d_steps, static_cast<size_t>(c->F) * c->N * c->R));
    const int rc = ensure_slice(c);
    if (rc) return rc;
    return ensure_dir(c, false);
}

int mcl_num_free_cells(const mcl_ctx* c) { return c ? static_cast<int>(c->skip.free_cells.size()) : 0; }

int mcl_init_pose(mcl_ctx* c, int filter, const double pose[3], const double* normals) {
    int rc = check_filter(c, filter, true);
    if (rc) return rc;
    if (!pose) return fail(MCL_ERR_INVALID, "null pose");
    CK(cudaSetDevice(c->device));
    const int f0 = filter < 0 ? 0 : filter, nf = filter < 0 ? c->F : 1;
    double* d_pose = nullptr;
    double* d_norm = nullptr;
    std::vector<double> hp(static_cast<size_t>(3) * nf);
    for (int k = 0; k < nf; ++k) std::memcpy(&hp[3 * k], pose, 3 * sizeof(double));
    CK(dalloc(&d_pose, hp.size()));
    CK(cudaMemcpyAsync(d_pose, hp.data(), hp.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (normals) {
        CK(dalloc(&d_norm, static_cast<size_t>(3) * c->N * nf));
        for (int k = 0; k < nf; ++k)   // the same injected stream for every addressed filter
            CK(cudaMemcpyAsync(d_norm + static_cast<size_t>(3) * c->N * k, normals, sizeof(double) * 3 * c->N,
                               cudaMemcpyHostToDevice, c->stream));
    }
    InitArgs a{};
    a.N = c->N;
    a.px = c->d_px[c->cur];
    a.py = c->d_py[c->cur];
    a.pt = c->d_pt[c->cur];
    a.wn = c->d_wn;
    a.pose = d_pose;
    a.normals = d_norm;
    a.seed = c->prm.seed;
    a.stream_no = ++c->init_no;
    a.filter0 = f0;
    k_init_pose<<<dim3(static_cast<unsigned>((c->N + 255) / 256), nf), 256, 0, c->stream>>>(a);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_pose);
    if (d_norm) cudaFree(d_norm);
    c->cdf_valid = false;
    c->tile_state = 0;
    c->pose4_ok[c->cur] = false;   // the packed copy no longer matches the state arrays
    return MCL_OK;
}

int mcl_init_global(mcl_ctx* c, int filter, const int32_t* cell, const double* theta) {
    int rc = check_filter(c, filter, true);
    if (rc) return rc;
    if (!c->have_map) return fail(MCL_ERR_NO_MAP, "map not set");   // :403-404
    if (c->skip.free_cells.empty()) return fail(MCL_ERR_NO_FREE_SPACE, "No free space found in map!");
    if ((cell == nullptr) != (theta == nullptr)) return fail(MCL_ERR_INVALID, "cell_ordinal and theta must both be given or both NULL");
    CK(cudaSetDevice(c->device));
    const int f0 = filter < 0 ? 0 : filter, nf = filter < 0 ? c->F : 1;
    int32_t* d_cell = nullptr;
    double* d_theta = nullptr;
    if (cell) {
        CK(dalloc(&d_cell, static_cast<size_t>(c->N) * nf));
        CK(dalloc(&d_theta, static_cast<size_t>(c->N) * nf));
        for (int k = 0; k < nf; ++k) {
            CK(cudaMemcpyAsync(d_cell + static_cast<size_t>(c->N) * k, cell, sizeof(int32_t) * c->N, cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(d_theta + static_cast<size_t>(c->N) * k, theta, sizeof(double) * c->N, cudaMemcpyHostToDevice, c->stream));
        }
    }
    InitArgs a{};
    a.N = c->N;
    a.px = c->d_px[c->cur];
    a.py = c->d_py[c->cur];
    a.pt = c->d_pt[c->cur];
    a.wn = c->d_wn;
    a.cell = d_cell;
    a.theta = d_theta;
    a.free_cells = c->d_free;
    a.n_free = static_cast<int>(c->skip.free_cells.size());
    a.W = c->map.W;
    a.res = c->res;
    a.ox = c->ox;
    a.oy = c->oy;
    a.seed = c->prm.seed;
    a.stream_no = ++c->init_no;
    a.filter0 = f0;
    k_init_global<<<dim3(static_cast<unsigned>((c->N + 255) / 256), nf), 256, 0, c->stream>>>(a);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    if (d_cell) cudaFree(d_cell);
    if (d_theta) cudaFree(d_theta);
    c->cdf_valid = false;
    c->tile_state = 0;
    c->pose4_ok[c->cur] = false;   // the packed copy no longer matches the state arrays
    return MCL_OK;
}

int mcl_set_particles(mcl_ctx* c, int filter, const double* P, const double* w) {
    int rc = check_filter(c, filter, false);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    const size_t N = static_cast<size_t>(c->N), fo = N * filter;
    if (P) {
        CK(cudaMemcpy(c->d_px[c->cur] + fo, P, N * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->d_py[c->cur] + fo, P + N, N * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->d_pt[c->cur] + fo, P + 2 * N, N * sizeof(double), cudaMemcpyHostToDevice));
        c->pose4_ok[c->cur] = false;   // the packed copy no longer matches the state arrays
    }
    if (w) {
        CK(cudaMemcpy(c->d_wn + fo, w, N * sizeof(double), cudaMemcpyHostToDevice));
        c->tile_state = 0;
    }
    c->cdf_valid = false;
    return MCL_OK;
}

int mcl_get_particles(mcl_ctx* c, int filter, double* P) {
    int rc = check_filter(c, filter, false);
    if (rc) return rc;
    if (!P) return fail(MCL_ERR_INVALID, "null output");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    const size_t N = static_cast<size_t>(c->N), fo = N * filter;
    CK(cudaMemcpy(P, c->d_px[c->cur] + fo, N * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(P + N, c->d_py[c->cur] + fo, N * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(P + 2 * N, c->d_pt[c->cur] + fo, N * sizeof(double), cudaMemcpyDeviceToHost));
    return MCL_OK;
}

static int get_array(mcl_ctx* c, int filter, const void* dev, size_t elem, size_t per_filter, void* out) {
    int rc = check_filter(c, filter, false);
    if (rc) return rc;
    if (!out) return fail(MCL_ERR_INVALID, "null output");
    if (!dev) return fail(MCL_ERR_INVALID, "array not available (option disabled or no update yet)");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(out, static_cast<const char*>(dev) + elem * per_filter * filter, elem * per_filter, cudaMemcpyDeviceToHost));
    return MCL_OK;
}

int mcl_get_weights(mcl_ctx* c, int filter, double* w) {
    return get_array(c, filter, c ? c->d_wn : nullptr, sizeof(double), c ? c->N : 0, w);
}
int mcl_get_raw_weights(mcl_ctx* c, int filter, double* w) {
    return get_array(c, filter, c ? c->d_wraw : nullptr, sizeof(double), c ? c->N : 0, w);
}
int mcl_get_cdf(mcl_ctx* c, int filter, double* out) {
    if (c && !c->cdf_valid) return fail(MCL_ERR_INVALID, "no CDF yet: call mcl_update first");
    return get_array(c, filter, c ? c->d_cdf : nullptr, sizeof(double), c ? c->N : 0, out);
}
int mcl_get_resample_indices(mcl_ctx* c, int filter, int32_t* out) {
    return get_array(c, filter, c ? c->d_idx : nullptr, sizeof(int32_t), c ? c->N : 0, out);
}
int mcl_get_range_steps(mcl_ctx* c, int filter, uint8_t* out) {
    return get_array(c, filter, c ? c->d_steps : nullptr, 1, c ? static_cast<size_t>(c->N) * c->R : 0, out);
}

int mcl_get_ranges(mcl_ctx* c, int filter, float* out) {
    int rc = check_filter(c, filter, false);
    if (rc) return rc;
    if (!out) return fail(MCL_ERR_INVALID, "null output");
    if (!c->d_steps) return fail(MCL_ERR_INVALID, "ranges are not kept: call mcl_set_keep_ranges(ctx, 1) before the update");
    CK(cudaSetDevice(c->device));
    const int64_t n = c->N * c->R;
    float* d_out = nullptr;
    CK(dalloc(&d_out, static_cast<size_t>(n)));
    k_steps_to_ranges<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->stream>>>(c->d_steps + n * filter, n, c->M, c->res,
                                                                                       c->prm.max_range, d_out);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost));
    cudaFree(d_out);
    return MCL_OK;
}

int mcl_update_dev(mcl_ctx* c, const double* action_dev, const float* obs_dev, int num_beams) {
    if (!c || !action_dev || !obs_dev) return fail(MCL_ERR_INVALID, "null argument");
    if (num_beams != c->R) return fail(MCL_ERR_INVALID, "num_beams %d != configured %d", num_beams, c->R);
    CK(cudaSetDevice(c->device));
    return update_steady(c, action_dev, obs_dev);
}

int mcl_read_pose(mcl_ctx* c, double* pose_out) {
    if (!c || !pose_out) return fail(MCL_ERR_INVALID, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->h_pose, c->d_pose, sizeof(double) * 3 * c->F, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::memcpy(pose_out, c->h_pose, sizeof(double) * 3 * c->F);
    return MCL_OK;
}

int mcl_synchronize(mcl_ctx* c) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return MCL_OK;
}

int mcl_update(mcl_ctx* c, const double* action, const float* obs, int num_beams, const mcl_noise* noise, double* pose_out) {
    if (!c || !action || !obs) return fail(MCL_ERR_INVALID, "null argument");
    if (!c->have_map) return fail(MCL_ERR_NO_MAP, "mcl_set_map has not been called");
    if (!c->have_beams) return fail(MCL_ERR_INVALID, "mcl_set_beam_angles has not been called");
    if (num_beams != c->R) return fail(MCL_ERR_INVALID, "num_beams %d != configured %d", num_beams, c->R);
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    std::memcpy(c->h_action, action, sizeof(double) * 3 * c->F);
    std::memcpy(c->h_obs, obs, sizeof(float) * c->R * c->F);
    CK(cudaMemcpyAsync(c->d_action, c->h_action, sizeof(double) * 3 * c->F + sizeof(float) * c->R * c->F, cudaMemcpyHostToDevice, s));
    const double* u_dev = nullptr;
    const double* z_dev = nullptr;
    if (noise) {
        bool any_u = false, any_z = false, all_u = true, all_z = true;
        for (int f = 0; f < c->F; ++f) {
            any_u |= noise[f].u_resample != nullptr;
            all_u &= noise[f].u_resample != nullptr;
            any_z |= noise[f].z_motion != nullptr;
            all_z &= noise[f].z_motion != nullptr;
        }
        if (any_u != all_u || any_z != all_z) return fail(MCL_ERR_INVALID, "noise must be injected for all filters of a batch or none");
        const size_t N = static_cast<size_t>(c->N);
        if (any_u) {
            if (!c->d_u) CK(dalloc(&c->d_u, N * c->F));
            for (int f = 0; f < c->F; ++f)
                CK(cudaMemcpyAsync(c->d_u + N * f, noise[f].u_resample, N * sizeof(double), cudaMemcpyHostToDevice, s));
            u_dev = c->d_u;
        }
        if (any_z) {
            if (!c->d_z) CK(dalloc(&c->d_z, 3 * N * c->F));
            for (int f = 0; f < c->F; ++f)
                CK(cudaMemcpyAsync(c->d_z + 3 * N * f, noise[f].z_motion, 3 * N * sizeof(double), cudaMemcpyHostToDevice, s));
            z_dev = c->d_z;
        }
    }
    int rc = noise ? update_device(c, c->d_action, c->d_obs, u_dev, z_dev) : update_steady(c, c->d_action, c->d_obs);
    if (rc) return rc;
    // the pose kernel has written the pose into h_pose (mapped pinned memory): no D2H copy call
    CK(cudaStreamSynchronize(s));
    if (pose_out) std::memcpy(pose_out, c->h_pose, sizeof(double) * 3 * c->F);
    if (c->profiling && c->ev_valid) read_stage_times(c);
    return MCL_OK;
}

int mcl_expected_pose(mcl_ctx* c, int filter, double pose_out[3]) {
    int rc = check_filter(c, filter, false);
    if (rc) return rc;
    if (!pose_out) return fail(MCL_ERR_INVALID, "null output");
    CK(cudaSetDevice(c->device));
    rc = launch_pose(c, c->d_wn, nullptr, nullptr, c->cur, false);
    if (rc) return rc;
    CK(cudaMemcpyAsync(c->h_pose, c->d_pose, sizeof(double) * 3 * c->F, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::memcpy(pose_out, c->h_pose + 3 * filter, sizeof(double) * 3);
    return MCL_OK;
}

int mcl_calc_range_many(mcl_ctx* c, const double* q, int64_t n, float* out) {
    if (!c || !q || !out) return fail(MCL_ERR_INVALID, "null argument");
    if (n < 0) return fail(MCL_ERR_INVALID, "negative query count");
    if (!c->have_map) return fail(MCL_ERR_NO_MAP, "map not set");   // reference returns MAX_RANGE here (:613)
    if (n == 0) return MCL_OK;
    CK(cudaSetDevice(c->device));
    double* d_q = nullptr;
    float* d_o = nullptr;
    CK(dalloc(&d_q, static_cast<size_t>(3 * n)));
    CK(dalloc(&d_o, static_cast<size_t>(n)));
    CK(cudaMemcpyAsync(d_q, q, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    QueryArgs a{};
    a.map = c->map;
    a.q = d_q;
    a.n = n;
    a.out = d_o;
    a.max_range = c->prm.max_range;
    k_range_queries<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->stream>>>(a);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_o, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_q);
    cudaFree(d_o);
    return MCL_OK;
}

int mcl_cast_ray(mcl_ctx* c, double x, double y, double angle, float* out) {
    const double q[3] = {x, y, angle};
    return mcl_calc_range_many(c, q, 1, out);
}

int mcl_sample_particles(mcl_ctx* c, int filter, int k, double* out) {
    int rc = check_filter(c, filter, false);
    if (rc) return rc;
    if (!out || k < 1) return fail(MCL_ERR_INVALID, "bad arguments");
    CK(cudaSetDevice(c->device));
    // CDF of the current weights, as visualize() builds it (:949)
    rc = build_cdf(c);
    if (rc) return rc;
    double* d_o = nullptr;
    CK(dalloc(&d_o, static_cast<size_t>(3) * k));
    const size_t fo = static_cast<size_t>(c->N) * filter;
    k_sample_particles<<<(k + 127) / 128, 128, 0, c->stream>>>(c->d_cdf + fo, c->N, c->d_px[c->cur] + fo, c->d_py[c->cur] + fo,
                                                               c->d_pt[c->cur] + fo, k, c->prm.seed, ++c->init_no, d_o);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(out, d_o, sizeof(double) * 3 * k, cudaMemcpyDeviceToHost));
    cudaFree(d_o);
    return MCL_OK;
}

int mcl_set_profiling(mcl_ctx* c, int enabled) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    c->profiling = enabled != 0;
    c->ev_valid = false;
    return MCL_OK;
}

int mcl_get_stage_ms(mcl_ctx* c, mcl_stage_ms* out) {
    if (!c || !out) return fail(MCL_ERR_INVALID, "null argument");
    if (c->profiling && c->ev_valid && !c->local_pending) {
        // the device-resident and sharded entry points do not synchronise: wait for the last event here
        CK(cudaSetDevice(c->device));
        CK(cudaEventSynchronize(c->ev[4]));
        read_stage_times(c);
    }
    *out = c->last_ms;
    return MCL_OK;
}

int mcl_set_keep_ranges(mcl_ctx* c, int enabled) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    c->keep_ranges = enabled != 0;
    if (c->keep_ranges && !c->d_steps && c->R > 0) CK(dalloc(&c->d_steps, static_cast<size_t>(c->F) * c->N * c->R));
    return MCL_OK;
}

int mcl_kernel_launches(mcl_ctx* c, int64_t* count) {
    if (!c || !count) return fail(MCL_ERR_INVALID, "null argument");
    *count = c->launches;
    return MCL_OK;
}

int mcl_set_shard(mcl_ctx* c, int64_t lo, int64_t count) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    if (c->F != 1) return fail(MCL_ERR_INVALID, "particle sharding applies to a single filter, not a batch");
    if (lo < 0 || count < 1 || lo + count > c->N) return fail(MCL_ERR_INVALID, "shard [%lld,+%lld) outside [0,%lld)",
                                                               (long long)lo, (long long)count, (long long)c->N);
    if (c->local_pending) return fail(MCL_ERR_INVALID, "update in flight");
    c->lo = lo;
    c->cnt = count;
    c->pose4_ok[0] = c->pose4_ok[1] = false;
    drop_graphs(c);
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return upload_replay_ctx(c);
}

int mcl_update_local_dev(mcl_ctx* c, const double* action_dev, const float* obs_dev, int num_beams, const double* u_dev,
                         const double* z_dev) {
    if (!c || !action_dev || !obs_dev) return fail(MCL_ERR_INVALID, "null argument");
    if (num_beams != c->R) return fail(MCL_ERR_INVALID, "num_beams %d != configured %d", num_beams, c->R);
    CK(cudaSetDevice(c->device));
    return update_local(c, action_dev, obs_dev, u_dev, z_dev);
}

int mcl_exchange_buffers_dev(mcl_ctx* c, void* ptrs_out[4], int64_t* n_total, int64_t* lo, int64_t* count) {
    if (!c || !ptrs_out) return fail(MCL_ERR_INVALID, "null argument");
    if (!c->local_pending) return fail(MCL_ERR_INVALID, "call mcl_update_local_dev first");
    const int dst = c->cur ^ 1;
    ptrs_out[0] = c->d_px[dst];
    ptrs_out[1] = c->d_py[dst];
    ptrs_out[2] = c->d_pt[dst];
    ptrs_out[3] = c->d_wraw;
    if (n_total) *n_total = c->N;
    if (lo) *lo = c->lo;
    if (count) *count = c->cnt;
    return MCL_OK;
}

static int install_peers(mcl_ctx* c, int world, int rank, const std::vector<const double*>& tab) {
    if (c->F != 1) return fail(MCL_ERR_INVALID, "peer-to-peer sharding applies to a single filter");
    if (world < 1 || rank < 0 || rank >= world || c->N % world) return fail(MCL_ERR_INVALID, "bad world/rank %d/%d for %lld particles", world, rank, (long long)c->N);
    if (c->d_peer_tab) cudaFree(c->d_peer_tab);
    if (c->d_partials) cudaFree(c->d_partials);
    c->d_peer_tab = nullptr;
    c->d_partials = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&c->d_peer_tab), sizeof(double*) * tab.size()));
    CK(cudaMemcpy(c->d_peer_tab, tab.data(), sizeof(double*) * tab.size(), cudaMemcpyHostToDevice));
    CK(dalloc(&c->d_partials, static_cast<size_t>(4) * world));
    CK(cudaMemset(c->d_partials, 0, sizeof(double) * 4 * world));
    c->world = world;
    c->rank = rank;
    c->lo = (c->N / world) * rank;
    c->cnt = c->N / world;
    c->p2p = true;
    c->pose4_ok[0] = c->pose4_ok[1] = false;
    drop_graphs(c);
    return upload_replay_ctx(c);   // the shard moved: the exact-replay context of the directional stage follows it
}

int mcl_ipc_export(mcl_ctx* c, void* handles_out, size_t capacity) {
    if (!c || !handles_out) return fail(MCL_ERR_INVALID, "null argument");
    if (capacity < 8 * sizeof(cudaIpcMemHandle_t)) return fail(MCL_ERR_INVALID, "need %zu bytes", 8 * sizeof(cudaIpcMemHandle_t));
    CK(cudaSetDevice(c->device));
    auto* h = static_cast<cudaIpcMemHandle_t*>(handles_out);
    for (int b = 0; b < 2; ++b) {
        CK(cudaIpcGetMemHandle(&h[b * 3 + 0], c->d_px[b]));
        CK(cudaIpcGetMemHandle(&h[b * 3 + 1], c->d_py[b]));
        CK(cudaIpcGetMemHandle(&h[b * 3 + 2], c->d_pt[b]));
        CK(cudaIpcGetMemHandle(&h[6 + b], c->d_pose4[b]));
    }
    return MCL_OK;
}

int mcl_ipc_import(mcl_ctx* c, int world, int rank, const void* handles) {
    if (!c || !handles) return fail(MCL_ERR_INVALID, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    const auto* h = static_cast<const cudaIpcMemHandle_t*>(handles);
    std::vector<const double*> tab(static_cast<size_t>(8) * world, nullptr);   // [x y t of buf 0 | of buf 1 | pose4 buf 0 | buf 1][world]
    for (int q = 0; q < world; ++q) {
        for (int k = 0; k < 8; ++k) {
            const double* ptr;
            if (q == rank) {
                const int b = k / 3, a = k % 3;
                ptr = k >= 6 ? reinterpret_cast<const double*>(c->d_pose4[k - 6])
                             : (a == 0 ? c->d_px[b] : (a == 1 ? c->d_py[b] : c->d_pt[b]));
            } else {
                void* p = nullptr;
                CK(cudaIpcOpenMemHandle(&p, h[q * 8 + k], cudaIpcMemLazyEnablePeerAccess));
                c->ipc_opened.push_back(p);
                ptr = static_cast<const double*>(p);
            }
            tab[static_cast<size_t>(k) * world + q] = ptr;
        }
    }
    return install_peers(c, world, rank, tab);
}

int mcl_set_peer_pointers(mcl_ctx* c, int world, int rank, const void* const* ptrs) {
    if (!c || !ptrs) return fail(MCL_ERR_INVALID, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    std::vector<const double*> tab(static_cast<size_t>(8) * world, nullptr);
    for (int q = 0; q < world; ++q)
        for (int k = 0; k < 8; ++k) tab[static_cast<size_t>(k) * world + q] = static_cast<const double*>(ptrs[q * 8 + k]);
    return install_peers(c, world, rank, tab);
}

int mcl_state_pointers_dev(mcl_ctx* c, void* ptrs_out[8]) {
    if (!c || !ptrs_out) return fail(MCL_ERR_INVALID, "null argument");
    for (int b = 0; b < 2; ++b) {
        ptrs_out[b * 3 + 0] = c->d_px[b];
        ptrs_out[b * 3 + 1] = c->d_py[b];
        ptrs_out[b * 3 + 2] = c->d_pt[b];
        ptrs_out[6 + b] = c->d_pose4[b];
    }
    return MCL_OK;
}

int mcl_p2p_buffers_dev(mcl_ctx* c, void** w_raw_dev, void** partials_dev) {
    if (!c || !w_raw_dev || !partials_dev) return fail(MCL_ERR_INVALID, "null argument");
    if (!c->p2p) return fail(MCL_ERR_INVALID, "peer-to-peer sharding is not set up");
    *w_raw_dev = c->d_wraw;
    *partials_dev = c->d_partials;
    return MCL_OK;
}

int mcl_update_finish_dev(mcl_ctx* c) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    return update_finish(c);
}

int mcl_microbench_gather(int device, int shared, size_t array_bytes, int iters_per_thread, double* gathers_per_second) {
    if (!gathers_per_second || iters_per_thread < 1) return fail(MCL_ERR_INVALID, "bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MCL_ERR_NO_DEVICE, "no CUDA device visible");
    }
    CK(cudaSetDevice(device));
    // power-of-two size; the shared variant is capped by the 227 KB shared memory (128 KB window)
    size_t bytes = 4096;
    while (bytes * 2 <= array_bytes) bytes *= 2;
    if (shared && bytes > 128 * 1024) bytes = 128 * 1024;
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, device));
    uint8_t* d_arr = nullptr;
    unsigned long long* d_sink = nullptr;
    CK(dalloc(&d_arr, bytes));
    CK(dalloc(&d_sink, size_t{1}));
    CK(cudaMemset(d_arr, 1, bytes));
    CK(cudaMemset(d_sink, 0, sizeof(unsigned long long)));
    const int blocks = prop.multiProcessorCount, threads = 1024;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const uint32_t mask = static_cast<uint32_t>(bytes - 1);
    if (shared) CK(cudaFuncSetAttribute(k_gather_bench<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {   // first repetition warms up
        CK(cudaEventRecord(e0));
        if (shared)
            k_gather_bench<true><<<blocks, threads, bytes>>>(d_arr, mask, iters_per_thread, d_sink);
        else
            k_gather_bench<false><<<blocks, threads>>>(d_arr, mask, iters_per_thread, d_sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    *gathers_per_second = 4.0 * iters_per_thread * static_cast<double>(blocks) * threads / (best * 1e-3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_arr);
    cudaFree(d_sink);
    return MCL_OK;
}

int mcl_set_graphs(mcl_ctx* c, int enabled) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->graphs_enabled = enabled != 0;
    drop_graphs(c);
    return MCL_OK;
}

int mcl_set_ray_mode(mcl_ctx* c, int mode) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    if (mode < 0 || mode > 2) return fail(MCL_ERR_INVALID, "ray mode %d not in {0 auto, 1 isotropic, 2 directional}", mode);
    if (mode == 2 && !c->dir_ready)
        return fail(MCL_ERR_UNSUPPORTED, "the directional ray stage needs one filter of at least %d particles, a map and a beam table",
                    kDirMinParticles);
    c->ray_mode = mode;
    drop_graphs(c);
    return MCL_OK;
}

int mcl_ray_stage_info(mcl_ctx* c, int* directional_ready, int* last_mode, int* box_cells, int* units) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    int plan[kPlanInts] = {0};
    if (c->dir_ready) {
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaMemcpy(plan, c->d_plan, sizeof(plan), cudaMemcpyDeviceToHost));
    }
    if (directional_ready) *directional_ready = c->dir_ready ? 1 : 0;
    if (last_mode) *last_mode = plan[kPlanMode];
    if (box_cells) *box_cells = c->dir_ready ? c->dir_box : 0;
    if (units) *units = plan[kPlanUnits];
    return MCL_OK;
}

int mcl_get_dir_map(mcl_ctx* c, int sector, uint8_t* out, int* pw, int* ph) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    if (!c->dir_ready) return fail(MCL_ERR_UNSUPPORTED, "the directional ray stage is not active for this context");
    if (sector < 0 || sector >= kDirSectors) return fail(MCL_ERR_INVALID, "sector %d not in [0,%d)", sector, kDirSectors);
    if (pw) *pw = c->skip.PW;
    if (ph) *ph = c->skip.PH;
    if (out) {
        CK(cudaSetDevice(c->device));
        CK(cudaStreamSynchronize(c->stream));
        const size_t ncell = static_cast<size_t>(c->skip.PW) * c->skip.PH;
        CK(cudaMemcpy(out, c->d_dirmaps + ncell * sector, ncell, cudaMemcpyDeviceToHost));
    }
    return MCL_OK;
}

int mcl_set_stream(mcl_ctx* c, void* stream) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->stream = stream ? static_cast<cudaStream_t>(stream) : c->own_stream;
    drop_graphs(c);
    return MCL_OK;
}

}  // extern "C"
