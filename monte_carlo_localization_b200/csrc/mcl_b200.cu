// csrc/mcl_b200.cu -- C ABI (include/mcl_b200.h) over the sm_100a kernels in kernels.cuh.
//
// Host side of the drop-in boundary: owns the device buffers of one ParticleFilter (or a
// batch of independent ones, or one rank's slice of a particle-sharded one), uploads map /
// table / beams, and sequences the kernels of one MCL update on a CUDA stream.  There is
// deliberately no CPU implementation behind these entry points: without a CUDA device
// mcl_create fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

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
#include "map_kernels.cuh"

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
constexpr int kMaxMarks = 64;

// NCCL is bound at run time (dlopen) and only by sharded contexts: the single-GPU library has no
// link-time dependency on it, and a process that already loaded an NCCL (torch) shares that copy.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return MCL_OK;
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(MCL_ERR_UNSUPPORTED, "libnccl.so.2 not found: %s", dlerror());
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GetErrorString)
        return fail(MCL_ERR_UNSUPPORTED, "libnccl.so.2 lacks a required symbol");
    g_nccl = a;
    return MCL_OK;
}

#define NK(call)                                                                                          \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess) return fail(MCL_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
    } while (0)

}  // namespace

struct mcl_ctx {
    mcl_params prm{};
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    int F = 1;
    int64_t N = 0;                   // particles held by this context (a sharded rank: its slice)
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
    float* d_gap = nullptr;          // Euclidean gap map: lives from mcl_set_map until the sector maps are built
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
    double* d_cdf2[2] = {nullptr, nullptr};   // discrete_distribution's _M_cp for the update that resamples FROM state buffer b
    int cdf_last = 0;                         // which of them the last update drew from
    double* d_mid2[2] = {nullptr, nullptr};   // [F][C] chunk-end values of each CDF: the middle level of the resampling search
    int32_t* d_idx = nullptr;
    uint8_t* d_steps = nullptr;
    bool keep_ranges = false;
    // wide configuration (MAX_RANGE_PX > 254 or more than 128 beams): reference-arithmetic march, 16-bit steps
    bool wide = false;
    std::vector<float> beam_angles;
    float* d_beam = nullptr;
    uint16_t* d_steps16 = nullptr;
    size_t obs_capacity = 0;          // floats per filter the obs staging holds
    double* d_u = nullptr;
    double* d_z = nullptr;
    double* d_action = nullptr;
    float* d_obs = nullptr;
    double* d_slice = nullptr;
    size_t slice_elems = 0;
    // exact-sum workspaces (exact_kernels.cuh)
    int T = 0, C = 0;
    double* d_tile_sum = nullptr;
    StepFn* d_chunk_fn = nullptr;
    StepFn* d_opq_pre = nullptr;
    int* d_opq_idx = nullptr;
    double* d_opq_add = nullptr;
    int* d_tile_opq = nullptr;
    int64_t* d_tile_elem = nullptr;
    int* d_list_chunk = nullptr;
    StepFn* d_list_fn = nullptr;       // [2][F][C]
    double* d_list_add = nullptr;      // [2][F][C][8]
    double* d_anchors = nullptr;
    double* d_anchor_val = nullptr;
    double* d_tile_start = nullptr;
    double* d_coarse2[2] = {nullptr, nullptr};   // [F][coarse_n] coarse level of the CDF search (exact values at segment ends), per CDF
    int coarse_n = 0, coarse_shift = 0;
    double* d_S1 = nullptr;
    double* d_S2 = nullptr;
    double* d_scratch_total = nullptr;
    double* d_slice_sum = nullptr;   // [kMaxWorld] approximate slice sums of all ranks
    double* d_rank_end2[2] = {nullptr, nullptr};   // [kMaxWorld] exact CDF value at the end of every rank's slice, per CDF

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
    bool last_dir = false;            // the last update ran the directional stage (decided on the host since round 2)
    DirSector sectors[kDirSectors];
    uint8_t* d_dirmaps = nullptr;     // [S][PH*PW]
    DirSector* d_sectors = nullptr;
    int beam_io[kMaxBeams] = {};
    bool beam_ranges = false;         // beam offsets sorted from beam 0 on (cyclically): the ray kernel reads beam index ranges from a table
    int* d_sec_tab = nullptr;         // [S+1] first unit | [S] first chunk, per sector
    DirReplayCtx* d_replay_ctx = nullptr;   // [2]: one per state buffer
    DirRec* d_rec = nullptr;          // [2][N]: slot order | heading-sorted order
    int* d_plan = nullptr;
    uint8_t* d_steps_sorted = nullptr;   // [R][stride]
    int64_t dir_stride = 0;
    int dir_box = 0;
    size_t dir_smem = 0;
    // particle-sharded filter: this context holds slots [glo, glo + N) of a filter of NG particles
    int world = 1, rank = 0;
    int64_t NG = 0, glo = 0;
    bool connected = false;           // the peers' exchange buffers are mapped
    int xmode = 1;                    // 1: kernels exchange and wait on their own (one rank per GPU); 0: host-ordered
    mcl_barrier_fn hook = nullptr;    // host-ordered exchange: called where every rank must have published
    void* hook_user = nullptr;
    uint8_t* d_mbox = nullptr;        // [2][world][kMboxSlot]
    unsigned long long* d_flag = nullptr;   // [kMaxWorld]
    unsigned long long* d_xseq = nullptr;
    double4* d_routed = nullptr;      // [N] source pose + index of every own slot, pushed by the sources' owners
    int* h_err = nullptr;             // mapped pinned: first device-side exchange error
    int* d_err = nullptr;
    ShardDev sh{};                    // device view (peer pointers), by value in kernel arguments
    double4* routed_peers[kMaxWorld] = {};
    uint32_t* inbox_peers[kMaxWorld] = {};   // every rank's request inbox [world senders][N] (two-hop routing)
    uint32_t* d_inbox = nullptr;
    uint32_t* d_where = nullptr;             // [N] (server, position) of every own slot's request
    unsigned int* d_req_count = nullptr;     // [kMaxWorld] requests appended per destination in the current update
    bool pdl = true;                         // programmatic dependent launches inside an update (mcl_set_pdl)
    int route_mode = -1;                     // -1 auto (two hops from 3 ranks on), 0 two-hop requests, 1 every rank tests all draws
    const StepFn* peer_list_fn[kMaxWorld] = {};
    const double* peer_list_add[kMaxWorld] = {};
    std::vector<void*> ipc_opened;
    char* arena = nullptr;            // ONE allocation for everything a peer touches: mailbox | flags | routed | overflow lists
    ncclComm_t comm = nullptr;        // mcl_create_sharded: bootstrap, state gathers, optional barrier transport
    unsigned long long* d_nccl_tok = nullptr;   // [world] tokens of the NCCL barrier
    unsigned int* d_done = nullptr;          // [F] block-completion counters
    unsigned int* d_route_done = nullptr;
    // what d_tile_sum currently holds: 0 nothing usable, 1 tile sums of w_norm (S2 valid), 2 tile sums of
    // w_raw with w_norm = w_raw / S1 (S1, S2 valid; rescaled on the fly for the approximate prefix)
    int tile_state = 0;
    // pinned staging of the host-facing update: [F][3] doubles of action followed by [F][R] floats of
    // scan in ONE buffer (one H2D copy per update); d_action / d_obs alias the device twin
    double* h_action = nullptr;
    float* h_obs = nullptr;
    double* h_pose = nullptr;
    double* d_pose_mapped = nullptr;   // device alias of h_pose (mapped pinned memory): the pose kernel writes it directly
    uint64_t update_no = 0, init_no = 0;
    unsigned long long* d_update_no = nullptr;   // device twin of update_no, read by k_resample_motion
    // scratch for the occasional calls (initialisers, range queries, viz sampling): grown on demand, never per call
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    // CUDA graphs of the steady-state update, one per state-buffer parity
    bool graphs_enabled = true;
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    int64_t graph_launches = 0;
    int64_t launches = 0;
    bool profiling = false;
    bool ev_valid = false;            // the stage events of a whole update have been recorded
    cudaEvent_t ev[7] = {};
    mcl_stage_ms last_ms{};
    // per-kernel events of the last profiled update
    cudaEvent_t mark_ev[kMaxMarks + 1] = {};
    const char* mark_name[kMaxMarks] = {};
    int nmarks = 0;
    unsigned long long* d_dbg = nullptr;   // diagnostics: phase cycle counts of the exact passes (mcl_debug_pass_cycles)
    int dbg_pass = -1;                // which pass kind records them (-1: none)
    bool cdf_valid = false;           // d_cdf is the CDF of the CURRENT weights
    bool cdf_any = false;             // d_cdf holds the CDF the last update drew from
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

// scratch buffer of the occasional calls
int ensure_tmp(mcl_ctx* c, size_t bytes) {
    if (bytes <= c->tmp_bytes) return MCL_OK;
    CK(cudaStreamSynchronize(c->stream));
    if (c->d_tmp) cudaFree(c->d_tmp);
    c->d_tmp = nullptr;
    c->tmp_bytes = 0;
    const size_t want = std::max<size_t>(bytes, 1 << 16);
    CK(cudaMalloc(&c->d_tmp, want));
    c->tmp_bytes = want;
    return MCL_OK;
}

// profiling: one event after every kernel of the update (mcl_get_kernel_ms)
// Kernels of the update are launched with programmatic stream serialization (PDL): a kernel's blocks may be scheduled
// while its predecessor drains; every kernel starts with griddepcontrol.wait (pdl_enter, device_utils.cuh), which
// returns once the predecessor has completed and its writes are visible, so only launch latency and block
// scheduling overlap.  Stream capture turns these launches into programmatic graph edges.
template <class... P, class... A>
inline cudaError_t launch_dep(bool pdl, void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, P(std::forward<A>(args))...);
}

void mark(mcl_ctx* c, const char* name) {
    c->launches++;
    if (!c->profiling || c->nmarks >= kMaxMarks) return;
    if (!c->mark_ev[c->nmarks + 1] && cudaEventCreate(&c->mark_ev[c->nmarks + 1]) != cudaSuccess) return;
    cudaEventRecord(c->mark_ev[c->nmarks + 1], c->stream);
    c->mark_name[c->nmarks++] = name;
}
void marks_begin(mcl_ctx* c) {
    c->nmarks = 0;
    if (!c->profiling) return;
    if (!c->mark_ev[0] && cudaEventCreate(&c->mark_ev[0]) != cudaSuccess) return;
    cudaEventRecord(c->mark_ev[0], c->stream);
}

bool sharded(const mcl_ctx* c) { return c->world > 1; }
// T == 1 on one GPU: the one-CTA-per-pass kernels (small filters, batches)
bool single_tile(const mcl_ctx* c) { return c->T == 1 && !sharded(c); }

ExactArgs exact_base(mcl_ctx* c) {
    ExactArgs a{};
    a.N = c->N;
    a.glo = c->glo;
    a.T = c->T;
    a.C = c->C;
    a.tile_sum = c->d_tile_sum;
    a.slice_sum = c->d_slice_sum;
    a.chunk_fn = c->d_chunk_fn;
    a.opq_pre = c->d_opq_pre;
    a.opq_idx = c->d_opq_idx;
    a.opq_add = c->d_opq_add;
    a.tile_opq = c->d_tile_opq;
    a.tile_elem = c->d_tile_elem;
    a.list_chunk = c->d_list_chunk;
    a.list_fn = c->d_list_fn;
    a.list_add = c->d_list_add;
    for (int q = 0; q < kMaxWorld; ++q) {
        a.peer_list_fn[q] = c->peer_list_fn[q];
        a.peer_list_add[q] = c->peer_list_add[q];
    }
    a.anchors = c->d_anchors;
    a.anchor_val = c->d_anchor_val;
    a.tile_start = c->d_tile_start;
    a.done = c->d_done;
    a.partial = c->d_partial;
    a.sh = c->sh;
    a.sh.fused = c->xmode;
    return a;
}

// host-ordered exchange: every rank must have launched (and finished) its publishing kernel before any
// rank's consuming kernel runs.  Either the caller's hook (ranks emulated on one GPU: the hook
// synchronises them) or a stream-ordered NCCL all-gather of one word per rank.
int exchange_barrier(mcl_ctx* c) {
    if (c->hook) {
        CK(cudaStreamSynchronize(c->stream));
        const int rc = c->hook(c->hook_user);
        if (rc != 0) return fail(MCL_ERR_INVALID, "barrier hook failed (%d)", rc);
        return MCL_OK;
    }
    if (c->comm) {
        NK(g_nccl.AllGather(c->d_nccl_tok + c->rank, c->d_nccl_tok, 1, ncclUint64, c->comm, c->stream));
        return MCL_OK;
    }
    return fail(MCL_ERR_INVALID, "host-ordered exchange needs a barrier hook or an NCCL communicator");
}

// approximate tile sums of src (+ the slice sums of all ranks of a sharded filter)
int launch_tile_sums(mcl_ctx* c, const double* src) {
    ExactArgs a = exact_base(c);
    a.src = src;
    launch_dep(c->pdl, k_tile_sums, dim3(dim3(c->T, c->F)), dim3(kTileChunks), 0, c->stream, a);
    mark(c, "k_tile_sums");
    if (sharded(c) && !c->xmode) {
        const int rc = exchange_barrier(c);
        if (rc) return rc;
        launch_dep(c->pdl, k_slice_sums_collect, dim3(1), dim3(kTileChunks), 0, c->stream, a);
        mark(c, "k_slice_sums_collect");
    }
    CK(cudaGetLastError());
    return MCL_OK;
}

enum PassKind { kPassRaw, kPassNormalise, kPassStored, kPassCdf };

// one exact pass (exact_kernels.cuh):
//   kPassRaw        S1 = sum w_raw                                    (:679)
//   kPassNormalise  w_norm = w_raw / S1 stored, pose, S2 = sum w_norm  (:680-686, :696-716, random.tcc:2666)
//   kPassStored     S2 = sum w_norm of weights set from outside
//   kPassCdf        running sums of w_norm / S2 at the tile starts     (random.tcc:2672)
int launch_pass(mcl_ctx* c, PassKind kind, int pose_buf, int cdf_idx) {
    ExactArgs a = exact_base(c);
    bool pose = false;
    switch (kind) {
        case kPassRaw:
            a.src = c->d_wraw;
            a.total = c->d_S1;
            break;
        case kPassNormalise:
            a.src = c->d_wraw;
            a.norm = c->d_S1;
            a.pre_norm = c->d_S1;
            a.store = c->d_wn;
            a.total = c->d_S2;
            a.px = c->d_px[pose_buf];
            a.py = c->d_py[pose_buf];
            a.pt = c->d_pt[pose_buf];
            a.pose_out = c->d_pose;
            a.pose_host = c->d_pose_mapped;
            a.update_no = c->d_update_no;
            pose = true;
            break;
        case kPassStored:
            a.src = c->d_wn;
            a.total = c->d_S2;
            break;
        case kPassCdf:
            a.src = c->d_wn;
            a.div = c->d_S2;
            a.pre_norm = c->tile_state == 2 ? c->d_S1 : nullptr;
            a.total = c->d_scratch_total;
            a.rank_end = sharded(c) ? c->d_rank_end2[cdf_idx] : nullptr;
            break;
    }
    a.dbg = (c->d_dbg && c->dbg_pass == static_cast<int>(kind)) ? c->d_dbg : nullptr;
    const dim3 g(c->T, c->F);
    if (pose)
        launch_dep(c->pdl, k_exact_pass<true>, dim3(g), dim3(kTileChunks), 0, c->stream, a);
    else
        launch_dep(c->pdl, k_exact_pass<false>, dim3(g), dim3(kTileChunks), 0, c->stream, a);
    mark(c, kind == kPassRaw ? "k_exact_pass(S1)" : kind == kPassNormalise ? "k_exact_pass(normalise+pose+S2)"
                                                : kind == kPassStored      ? "k_exact_pass(S2)"
                                                                           : "k_exact_pass(cdf)");
    if (sharded(c) && !c->xmode) {
        const int rc = exchange_barrier(c);
        if (rc) return rc;
        if (pose)
            launch_dep(c->pdl, k_exact_finish<true>, dim3(1), dim3(kTileChunks), 0, c->stream, a);
        else
            launch_dep(c->pdl, k_exact_finish<false>, dim3(1), dim3(kTileChunks), 0, c->stream, a);
        mark(c, "k_exact_finish");
    }
    CK(cudaGetLastError());
    return MCL_OK;
}

int launch_emit(mcl_ctx* c, int cdf_idx) {
    ExactArgs a = exact_base(c);
    a.src = c->d_wn;
    a.div = c->d_S2;
    a.out = c->d_cdf2[cdf_idx];
    a.force_last_one = c->rank == c->world - 1 ? 1 : 0;   // _M_cp.back() = 1.0 is the LAST particle of the whole filter
    a.coarse = c->coarse_n > 0 ? c->d_coarse2[cdf_idx] : nullptr;
    a.coarse_m = a.coarse ? (1 << c->coarse_shift) / kChunk : 0;
    a.coarse_n = c->coarse_n;
    a.mid = c->d_mid2[cdf_idx];
    launch_dep(c->pdl, k_exact_emit, dim3(dim3(c->T, c->F)), dim3(kTileChunks), 0, c->stream, a);
    mark(c, "k_exact_emit");
    CK(cudaGetLastError());
    return MCL_OK;
}

// one-tile filters: a whole pass in one CTA per filter
int launch_single(mcl_ctx* c, const double* src, const double* div, double* total, double* out, int force_one, int cdf_idx) {
    ExactArgs a = exact_base(c);
    a.src = src;
    a.div = div;
    a.total = total;
    a.out = out;
    a.force_last_one = force_one;
    a.coarse = (out && c->coarse_n > 0) ? c->d_coarse2[cdf_idx] : nullptr;
    a.coarse_m = a.coarse ? (1 << c->coarse_shift) / kChunk : 0;
    a.coarse_n = c->coarse_n;
    a.mid = out ? c->d_mid2[cdf_idx] : nullptr;
    launch_dep(c->pdl, k_exact_single, dim3(dim3(1, c->F)), dim3(kTileChunks), 0, c->stream, a);
    mark(c, "k_exact_single");
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

// MAX_RANGE_PX > 254 or more than 128 beams: the skip-map stages cannot run (9.23 fixed point, one-byte step
// indices, beam tables in the kernel parameters); the context marches rays with the reference's own arithmetic
int update_wide(mcl_ctx* c) {
    const bool wide = (c->have_map && c->M > kMaxRangePxSupported) || (c->have_beams && c->R > kMaxBeams);
    if (wide != c->wide) drop_graphs(c);
    c->wide = wide;
    if (c->d_steps16) {
        cudaFree(c->d_steps16);
        c->d_steps16 = nullptr;
    }
    if (wide && c->have_beams) CK(dalloc(&c->d_steps16, static_cast<size_t>(c->F) * c->N * c->R));
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

// one sector window in shared memory (the ray-start prefetch buffers take another 64 KB): wider sectors reach further sideways
constexpr size_t kDirWindowBudget = (kDirSectors >= 32 ? 112 : 120) * 1024;

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
        h[b].perm = c->d_perm;
        h[b].lo = 0;
        h[b].nfil = c->dir_pool ? c->N : (int64_t{1} << 40);
        for (int j = 0; j < kMaxBeams; ++j) h[b].angle[j] = j < c->R ? c->beams.angle[j] : 0.0f;
    }
    CK(cudaMemcpy(c->d_replay_ctx, h, sizeof(h), cudaMemcpyHostToDevice));
    return MCL_OK;
}

// Skip codes (v8, v4) and the Euclidean gap map from the int8 grid, on the device (map_kernels.cuh).  `codes` false:
// only the gap map is wanted again (d_v8 / d_v4 are written with the same bytes).
int device_build_maps(mcl_ctx* c, bool codes) {
    const int PW = c->skip.PW, PH = c->skip.PH;
    const size_t n = static_cast<size_t>(PW) * PH;
    if (codes) {
        CK(dalloc(&c->d_v8, n));
        CK(dalloc(&c->d_v4, n / 2));
    }
    if (!c->d_gap) CK(dalloc(&c->d_gap, n));
    // scratch: blocked | dil (bytes), column distances + envelope stack (ints), envelope bounds (doubles), squared distances
    struct Scratch {
        void* p = nullptr;
        ~Scratch() {
            if (p) cudaFree(p);
        }
    } scr;
    const size_t nz = n + static_cast<size_t>(PH);
    const size_t bytes = 2 * n + 2 * n * sizeof(int) + nz * sizeof(double) + n * sizeof(long long) + 64;
    CK(cudaMalloc(&scr.p, bytes));
    char* b = static_cast<char*>(scr.p);
    long long* d2T = reinterpret_cast<long long*>(b);
    double* z = reinterpret_cast<double*>(b + n * sizeof(long long));
    int* gT = reinterpret_cast<int*>(b + n * sizeof(long long) + nz * sizeof(double));
    int* v = gT + n;
    uint8_t* blocked = reinterpret_cast<uint8_t*>(v + n);
    uint8_t* dil = blocked + n;
    cudaStream_t s = c->stream;
    k_map_masks<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(c->d_grid, c->skip.W, c->skip.H, PW, PH, blocked, dil);
    k_edt_cols<<<static_cast<unsigned>((PW + 127) / 128), 128, 0, s>>>(dil, PW, PH, gT);
    k_edt_rows<<<static_cast<unsigned>((PH + 127) / 128), 128, 0, s>>>(gT, PW, PH, v, z, d2T);
    k_map_codes<<<static_cast<unsigned>((n / 2 + 255) / 256), 256, 0, s>>>(blocked, dil, d2T, PW, PH, c->d_v8, c->d_v4, c->d_gap);
    c->launches += 4;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    return MCL_OK;
}

// Directional ray stage: eligible for ONE large filter whose heading sort has at least
// kDirMinBuckets buckets.  Builds the sector maps on the device (k_build_dir_maps) and the
// per-update work buffers.  Called whenever the map or the beam table changes.
int ensure_dir(mcl_ctx* c, bool map_changed) {
    if (!c->have_map || !c->have_beams) return MCL_OK;
    if (c->wide) {
        free_dir(c);
        return MCL_OK;
    }
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
            smem = std::max(smem, static_cast<size_t>(dir_smem_pitch((c->dir_box + sc.exh - sc.exl + 30) & ~15)) *
                                      static_cast<size_t>(c->dir_box + sc.eyh - sc.eyl));
        }
        c->dir_smem = pool ? ncell : smem;
        if (!c->d_gap) {   // the map's gap field was released after an earlier build: transform again
            const int rc = device_build_maps(c, false);
            if (rc) return rc;
        }
        CK(dalloc(&c->d_sectors, static_cast<size_t>(kDirSectors)));
        CK(cudaMemcpy(c->d_sectors, c->sectors, sizeof(c->sectors), cudaMemcpyHostToDevice));
        CK(dalloc(&c->d_dirmaps, ncell * kDirSectors));
        DirBuildArgs ba{c->d_v8, c->d_gap, c->d_sectors, c->d_dirmaps, c->skip.PW, c->skip.PH};
        k_build_dir_maps<<<dim3(static_cast<unsigned>((ncell + 255) / 256), kDirSectors), 256, 0, c->stream>>>(ba);
        c->launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_gap);
        c->d_gap = nullptr;
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
    c->beam_ranges = true;   // any real scan: angles increase with the beam index and span less than a turn
    for (int j = 1; j < c->R; ++j)
        if (((c->beam_io[j] - c->beam_io[0]) & (c->dir_B - 1)) < ((c->beam_io[j - 1] - c->beam_io[0]) & (c->dir_B - 1))) c->beam_ranges = false;
    CK(dalloc(&c->d_sec_tab, static_cast<size_t>(2 * kDirSectors + 2)));
    CK(dalloc(&c->d_replay_ctx, size_t{2}));
    const int64_t nchunks = (pool_n + kDirThreads - 1) / kDirThreads;
    c->dir_stride = nchunks * kDirThreads;
    CK(dalloc(&c->d_steps_sorted, static_cast<size_t>(c->R) * c->dir_stride));
    CK(cudaMemset(c->d_steps_sorted, 0, static_cast<size_t>(c->R) * c->dir_stride));   // slots beyond the shard stay valid steps
    if (dir_ray_smem(static_cast<int>((c->dir_smem + 15) & ~size_t{15})) > kWindowBudget - 10 * 1024) return MCL_OK;   // stays on the isotropic kernel (the kernel has 9 KB of static shared memory)
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
    launch_dep(c->pdl, k_normalize_pose, dim3(dim3(c->norm_blocks, c->F)), dim3(kNormThreads), 0, c->stream, na);
    c->launches += 1;
    CK(cudaGetLastError());
    return MCL_OK;
}

// discrete_distribution(weights_): sum, normalise, partial_sum (random.tcc:2657-2678) of the
// current normalised weights.  After an update S2 = sum(w_norm) is already known (the normalising
// pass computed it) and the tile sums of w_raw tell the binades of the running sums, so only the CDF pass and
// the emit run; after weights were set from outside, tile sums and S2 are rebuilt first.
int ensure_cdf(mcl_ctx* c) {
    if (c->cdf_valid) return MCL_OK;
    const int ci = c->cur;   // the CDF the next update (which resamples from state buffer `cur`) draws from
    int rc;
    if (single_tile(c)) {
        if (c->tile_state == 0) {
            rc = launch_single(c, c->d_wn, nullptr, c->d_S2, nullptr, 0, ci);
            if (rc) return rc;
            c->tile_state = 1;
        }
        rc = launch_single(c, c->d_wn, c->d_S2, c->d_scratch_total, c->d_cdf2[ci], 1, ci);
        if (rc) return rc;
    } else {
        if (c->tile_state == 0) {
            rc = launch_tile_sums(c, c->d_wn);
            if (rc) return rc;
            rc = launch_pass(c, kPassStored, c->cur, ci);
            if (rc) return rc;
            c->tile_state = 1;
        }
        rc = launch_pass(c, kPassCdf, c->cur, ci);
        if (rc) return rc;
        rc = launch_emit(c, ci);
        if (rc) return rc;
    }
    c->cdf_valid = true;
    return MCL_OK;
}

// MCL() (:652-694) + expected_pose() (:696-716) for this context's particles:
//   CDF of the current weights -> resample + motion -> heading sort -> ray cast -> weights ->
//   exact weight sum -> normalise + pose (+ the sum the next CDF needs).
// u_dev / z_dev: injected noise indexed by the GLOBAL slot (nullptr: device Philox).
int update_device(mcl_ctx* c, const double* action_dev, const float* obs_dev, const double* u_dev, const double* z_dev) {
    if (!c->have_map) return fail(MCL_ERR_NO_MAP, "mcl_set_map has not been called");
    if (!c->have_beams) return fail(MCL_ERR_INVALID, "mcl_set_beam_angles has not been called");
    if (sharded(c) && !c->connected) return fail(MCL_ERR_INVALID, "sharded context: mcl_shard_connect* has not been called");
    const int src = c->cur, dst = c->cur ^ 1;
    cudaStream_t s = c->stream;
    marks_begin(c);
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
    oa.hist = (c->sort_enabled && !c->wide) ? c->d_hist : nullptr;
    oa.nhist = 2 * c->B * c->F;
    launch_dep(c->pdl, k_prepare_obs, dim3(dim3(c->R, c->F)), dim3(256), 0, s, oa);
    mark(c, "k_prepare_obs");

    int rc = ensure_cdf(c);
    if (rc) return rc;
    if (c->profiling) CK(cudaEventRecord(c->ev[1], s));

    // (a batch's pool mode was opt-in while the stage had 32 sectors; with 16 it beats the isotropic kernel: 2.67 against
    // 3.28 ms for 1024 x 4000 particles on sibal1)
    const bool dir = !c->wide && c->dir_ready && c->sort_enabled && c->ray_mode != 1;
    const bool sort = c->sort_enabled && !c->wide;   // (the heading order only serves the skip-map ray kernels)
    c->last_dir = dir;
    static const bool no_packed = std::getenv("MCL_NO_PACKED") != nullptr;   // debugging knob
    const bool packed = !no_packed && c->pose4_ok[src];   // the packed source copy was written by the last update (no set_particles / init since)
    // routing of the resampling draws: two hops from 3 ranks on (work per rank independent of the world size), one hop for
    // 2 ranks (one exchange less; measured 0.757 against 0.765 ms at 2 GPUs, 0.891 against 0.838 ms at 8)
    const bool two_hop = sharded(c) && (c->route_mode == 0 || (c->route_mode < 0 && c->world >= 3));
    if (two_hop) {
        // sender-driven resampling in two hops: the slot owners classify their own draws and append
        // requests to the source ranks' inboxes (k_route_request), one exchange of the counts, then the
        // source ranks search and push the source poses to the slots' owners (k_route_serve)
        RouteReqArgs qa{};
        qa.N = c->N;
        qa.rank_end = c->d_rank_end2[src];
        for (int q = 0; q < kMaxWorld; ++q) qa.inbox[q] = c->inbox_peers[q];
        qa.req_count = c->d_req_count;
        qa.where = c->d_where;
        qa.u = u_dev;
        qa.seed = c->prm.seed;
        qa.update_no = c->d_update_no;
        qa.done = c->d_route_done;
        qa.dbg = (c->d_dbg && c->dbg_pass == 9) ? c->d_dbg : nullptr;
        qa.sh = c->sh;
        qa.sh.fused = c->xmode;
        launch_dep(c->pdl, k_route_request, dim3(static_cast<unsigned>((c->N + kReqBlock - 1) / kReqBlock)), dim3(kReqThreads), 0, s, qa);
        mark(c, "k_route_request");
        ShardDev sd = c->sh;
        sd.fused = 0;
        if (!c->xmode) {
            rc = exchange_barrier(c);
            if (rc) {   // the requests of this update will never be served: the next update must not count them again
                cudaMemsetAsync(c->d_req_count, 0, sizeof(unsigned int) * kMaxWorld, s);
                return rc;
            }
            launch_dep(c->pdl, k_route_check, dim3(1), dim3(32), 0, s, sd);
            mark(c, "k_route_check");
        }
        RouteServeArgs va{};
        va.N = c->N;
        va.cdf = c->d_cdf2[src];
        va.coarse = c->coarse_n > 0 ? c->d_coarse2[src] : nullptr;
        va.nc = c->coarse_n;
        va.cshift = c->coarse_shift;
        va.mid = c->d_mid2[src];
        va.spose4 = packed ? c->d_pose4[src] : nullptr;
        va.sx = c->d_px[src];
        va.sy = c->d_py[src];
        va.st = c->d_pt[src];
        va.inbox = c->d_inbox;
        for (int q = 0; q < kMaxWorld; ++q) va.routed[q] = c->routed_peers[q];
        va.req_count = c->d_req_count;
        va.action = action_dev;
        va.z = z_dev;
        va.noise = MotionNoise{c->prm.motion_dispersion_x, c->prm.motion_dispersion_y, c->prm.motion_dispersion_theta, c->prm.seed};
        va.u = u_dev;
        va.seed = c->prm.seed;
        va.update_no = c->d_update_no;
        va.done = c->d_route_done;
        va.dbg = (c->d_dbg && c->dbg_pass == 8) ? c->d_dbg : nullptr;
        va.sh = c->sh;
        va.sh.fused = c->xmode;
        const size_t vsmem = sizeof(uint32_t) * static_cast<size_t>((coarse_slots(c->coarse_n) + 1) & ~1);
        const int vblocks = static_cast<int>(std::min<int64_t>(c->num_sms, (c->N + kRouteThreads - 1) / kRouteThreads));
        launch_dep(c->pdl, k_route_serve, dim3(vblocks), dim3(kRouteThreads), vsmem, s, va);
        mark(c, "k_route_serve");
        if (!c->xmode) {
            rc = exchange_barrier(c);
            if (rc) return rc;
            launch_dep(c->pdl, k_route_check, dim3(1), dim3(32), 0, s, sd);
            mark(c, "k_route_check");
        }
    } else if (sharded(c)) {
        // one hop: every rank tests all NG draws, serves those that fall into its CDF range and pushes
        // the source poses to the slots' owners (k_route); work per rank grows with the number of ranks
        RouteArgs ra{};
        ra.NG = c->NG;
        ra.N = c->N;
        ra.cdf = c->d_cdf2[src];
        ra.coarse = c->coarse_n > 0 ? c->d_coarse2[src] : nullptr;
        ra.nc = c->coarse_n;
        ra.cshift = c->coarse_shift;
        ra.mid = c->d_mid2[src];
        ra.rank_end = c->d_rank_end2[src];
        ra.spose4 = packed ? c->d_pose4[src] : nullptr;
        ra.sx = c->d_px[src];
        ra.sy = c->d_py[src];
        ra.st = c->d_pt[src];
        for (int q = 0; q < kMaxWorld; ++q) ra.routed[q] = c->routed_peers[q];
        ra.u = u_dev;
        ra.seed = c->prm.seed;
        ra.update_no = c->d_update_no;
        ra.done = c->d_route_done;
        ra.dbg = (c->d_dbg && c->dbg_pass == 8) ? c->d_dbg : nullptr;
        ra.sh = c->sh;
        ra.sh.fused = c->xmode;
        const size_t rsmem = sizeof(uint32_t) * static_cast<size_t>((coarse_slots(c->coarse_n) + 1) & ~1) + (kRouteThreads / 32) * kRouteQueue * (sizeof(double) + sizeof(int));
        const int rblocks = static_cast<int>(std::min<int64_t>(c->num_sms, (c->NG + 2 * kRouteThreads - 1) / (2 * kRouteThreads)));
        launch_dep(c->pdl, k_route, dim3(rblocks), dim3(kRouteThreads), rsmem, s, ra);
        mark(c, "k_route");
        if (!c->xmode) {
            rc = exchange_barrier(c);
            if (rc) return rc;
            ShardDev sd = c->sh;
            sd.fused = 0;
            launch_dep(c->pdl, k_route_check, dim3(1), dim3(32), 0, s, sd);
            mark(c, "k_route_check");
        }
    }
    MotionArgs ma{};
    ma.rec = dir ? c->d_rec : nullptr;
    ma.map = c->map;
    ma.B = c->dir_B;
    ma.N = c->N;
    ma.glo = c->glo;
    ma.cdf = c->d_cdf2[src];
    ma.sx = c->d_px[src];
    ma.sy = c->d_py[src];
    ma.st = c->d_pt[src];
    ma.dx = c->d_px[dst];
    ma.dy = c->d_py[dst];
    ma.dt = c->d_pt[dst];
    ma.idx_out = c->d_idx;
    ma.u = u_dev;
    ma.z = z_dev;
    ma.spose4 = packed ? c->d_pose4[src] : nullptr;
    ma.dpose4 = c->d_pose4[dst];
    ma.routed = sharded(c) ? c->d_routed : nullptr;
    ma.where = two_hop ? c->d_where : nullptr;
    ma.coarse = c->coarse_n > 0 ? c->d_coarse2[src] : nullptr;
    ma.nc = c->coarse_n;
    ma.cshift = c->coarse_shift;
    ma.mid = c->d_mid2[src];
    ma.C = c->C;
    ma.action = action_dev;
    ma.disp_x = c->prm.motion_dispersion_x;
    ma.disp_y = c->prm.motion_dispersion_y;
    ma.disp_t = c->prm.motion_dispersion_theta;
    ma.seed = c->prm.seed;
    ma.update_no = c->d_update_no;
    ma.centre = c->d_centre;
    ma.hist = sort ? c->d_hist : nullptr;   // the counting sort's histogram is accumulated by the motion kernel
    ma.hist_B = c->B;
    int mblocks = static_cast<int>((c->N + kMotionThreads - 1) / kMotionThreads);
    const size_t msmem = sharded(c) ? 0 : sizeof(uint32_t) * static_cast<size_t>(coarse_slots(c->coarse_n));
    // a large table is staged once per SM by persistent blocks; a small one by every block
    if (msmem > 8 * 1024) mblocks = std::min(mblocks, std::max(1, c->num_sms / std::min(c->F, c->num_sms)));
    launch_dep(c->pdl, k_resample_motion, dim3(dim3(mblocks, c->F)), dim3(kMotionThreads), msmem, s, ma);
    mark(c, sharded(c) ? "k_resample_motion(routed)" : "k_resample_motion");
    c->pose4_ok[dst] = true;
    if (sort) {
        SortArgs sa{};
        sa.N = c->N;
        sa.lo = 0;
        sa.cnt = c->N;
        sa.pt = c->d_pt[dst];
        sa.hist = c->d_hist;
        sa.cursor = c->d_hist + static_cast<size_t>(c->B) * c->F;
        sa.perm = c->d_perm;
        sa.B = c->B;
        // a few fat blocks per filter: ~one per SM for a single big filter
        const int64_t per_filter = std::max<int64_t>(1, c->num_sms / std::min(c->F, c->num_sms));
        int64_t chunk = (c->N + per_filter - 1) / per_filter;
        chunk = std::max<int64_t>(kSortThreads, (chunk + kSortThreads - 1) / kSortThreads * kSortThreads);
        sa.chunk = chunk;
        const dim3 gs(static_cast<unsigned>((c->N + chunk - 1) / chunk), c->F);
        launch_dep(c->pdl, k_sort_scatter, dim3(gs), dim3(kSortThreads), 0, s, sa);
        mark(c, "k_sort_scatter");
    }
    const int64_t slots = static_cast<int64_t>(c->F) * c->N;   // slots of the directional stage (== N for one filter)
    if (dir) {
        DirPrepArgs pa{};
        pa.map = c->map;
        pa.centre = c->d_centre;
        pa.rec_in = c->d_rec;
        pa.rec = c->d_rec + slots;
        pa.perm = c->d_perm;
        pa.plan = c->d_plan;
        pa.cnt = slots;
        pa.nfil = c->dir_pool ? c->N : (int64_t{1} << 40);
        pa.box = c->dir_box;
        pa.whole = c->dir_pool ? 1 : 0;
        launch_dep(c->pdl, k_dir_gather, dim3(static_cast<unsigned>((slots + 255) / 256)), dim3(256), 0, s, pa);
        mark(c, "k_dir_gather");
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
        launch_dep(c->pdl, k_dir_plan, dim3(1), dim3(kPlanThreads), 0, s, la);
        mark(c, "k_dir_plan");
    }
    if (c->profiling) CK(cudaEventRecord(c->ev[2], s));

    if (c->wide) {
        // reference-arithmetic march (kernels.cuh, "wide configurations"): any MAX_RANGE_PX, any number of beams
        WideRayArgs wa{};
        wa.grid = RefGrid{c->d_grid, c->map.W, c->map.H, c->res, c->ox, c->oy};
        wa.M = c->M;
        wa.R = c->R;
        wa.N = c->N;
        wa.px = c->d_px[dst];
        wa.py = c->d_py[dst];
        wa.pt = c->d_pt[dst];
        wa.beam = c->d_beam;
        wa.steps = c->d_steps16;
        const int64_t rays = c->N * c->R;
        const unsigned wblocks = static_cast<unsigned>(std::min<int64_t>((rays + 255) / 256, static_cast<int64_t>(c->num_sms) * 32));
        k_raycast_wide<<<dim3(wblocks, c->F), 256, 0, s>>>(wa);
        mark(c, "k_raycast_wide");
        if (c->profiling) CK(cudaEventRecord(c->ev[5], s));
        WideWeightArgs ww{};
        ww.M = c->M;
        ww.R = c->R;
        ww.N = c->N;
        ww.steps = c->d_steps16;
        ww.slice = c->d_slice;
        ww.w_raw = c->d_wraw;
        ww.inv_squash = 1.0 / c->prm.squash_factor;
        k_weight_wide<<<dim3(static_cast<unsigned>((c->N + 255) / 256), c->F), 256, 0, s>>>(ww);
        mark(c, "k_weight_wide");
    } else if (!dir) {   // (the directional stage handles compact and scattered clouds alike: no fallback launch)
        RayArgs ra{};
        ra.map = c->map;
        ra.beams = c->beams;
        ra.N = c->N;
        ra.lo = 0;
        ra.cnt = c->N;
        ra.px = c->d_px[dst];
        ra.py = c->d_py[dst];
        ra.pt = c->d_pt[dst];
        ra.perm = sort ? c->d_perm : nullptr;
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
        const int rblocks = static_cast<int>(std::min<int64_t>((c->N + kRayThreads - 1) / kRayThreads, budget));
        // MAX_RANGE_PX of the usual map resolutions is baked into specialised instances
        // (0.05 m -> 239, 0.0504 m -> 238, 0.05796 m -> 207); anything else takes the generic one
        const dim3 rgrid(rblocks, c->F);
#define MCL_LAUNCH_RAY(WB, MCV) launch_dep(c->pdl, (k_raycast_weight<WB, MCV>), dim3(rgrid), dim3(kRayThreads), smem, s, ra)
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
        mark(c, "k_raycast_weight");
    }
    if (dir) {
        DirRayArgs da{};
        da.map = c->map;
        da.beams = c->beams;
        std::memcpy(da.io, c->beam_io, sizeof(da.io));
        da.sectors = c->d_sectors;
        da.dirmaps = c->d_dirmaps;
        da.rec = c->d_rec + slots;
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
        da.beam_ranges = c->beam_ranges ? 1 : 0;
        const size_t dsmem = dir_ray_smem(da.win_bytes);
        const int dblocks = static_cast<int>(std::min<int64_t>(c->num_sms, (slots * c->R + kDirThreads - 1) / kDirThreads));
        switch (c->M) {
            case 207: launch_dep(c->pdl, k_raycast_dir<207>, dim3(dblocks), dim3(kDirThreads), dsmem, s, da); break;
            case 238: launch_dep(c->pdl, k_raycast_dir<238>, dim3(dblocks), dim3(kDirThreads), dsmem, s, da); break;
            case 239: launch_dep(c->pdl, k_raycast_dir<239>, dim3(dblocks), dim3(kDirThreads), dsmem, s, da); break;
            default: launch_dep(c->pdl, k_raycast_dir<0>, dim3(dblocks), dim3(kDirThreads), dsmem, s, da); break;
        }
        mark(c, "k_raycast_dir");
        if (c->profiling) CK(cudaEventRecord(c->ev[5], s));
        WeightStepsArgs wa{};
        wa.plan = c->d_plan;
        wa.steps_sorted = c->d_steps_sorted;
        wa.perm = c->d_perm;
        wa.slice = c->d_slice;
        wa.w_raw = c->d_wraw;
        wa.steps = c->keep_ranges ? c->d_steps : nullptr;
        wa.lo = 0;
        wa.cnt = slots;
        wa.nfil = c->dir_pool ? c->N : (int64_t{1} << 40);
        wa.stride = c->dir_stride;
        wa.R = c->R;
        wa.tw = c->M + 1;
        wa.inv_squash = 1.0 / c->prm.squash_factor;
        const unsigned wblocks = static_cast<unsigned>((slots + 4 * kWeightThreads - 1) / (4 * kWeightThreads));
        const size_t tab_bytes = sizeof(double) * static_cast<size_t>(c->R) * (c->M + 1);
        static const bool no_sm_table = std::getenv("MCL_NO_SMEM_TABLE") != nullptr;   // debugging knob
        if (c->dir_pool && !no_sm_table && (c->N % 4) == 0 && tab_bytes <= 220 * 1024) {
            // a CTA per filter: the filter's own slice in shared memory, its (contiguous) sorted slots
            const unsigned g = static_cast<unsigned>(std::min<int64_t>(c->num_sms, c->F));
            launch_dep(c->pdl, k_weight_steps_sm<1024>, dim3(g), dim3(1024), tab_bytes, s, wa);
        } else if (c->dir_pool) {
            launch_dep(c->pdl, k_weight_steps<true>, dim3(wblocks), dim3(kWeightThreads), 0, s, wa);
        } else if (!no_sm_table && tab_bytes <= 110 * 1024) {   // two persistent 512-thread CTAs per SM share the SM's shared memory
            const unsigned g = static_cast<unsigned>(std::min<int64_t>(2 * c->num_sms, (slots + 4 * 512 - 1) / (4 * 512)));
            launch_dep(c->pdl, k_weight_steps_sm<512>, dim3(g), dim3(512), tab_bytes, s, wa);
        } else if (!no_sm_table && tab_bytes <= 220 * 1024) {
            const unsigned g = static_cast<unsigned>(std::min<int64_t>(c->num_sms, (slots + 4 * 1024 - 1) / (4 * 1024)));
            launch_dep(c->pdl, k_weight_steps_sm<1024>, dim3(g), dim3(1024), tab_bytes, s, wa);
        } else {
            launch_dep(c->pdl, k_weight_steps<false>, dim3(wblocks), dim3(kWeightThreads), 0, s, wa);
        }
        mark(c, "k_weight_steps");
    }
    if (!dir && !c->wide && c->profiling) CK(cudaEventRecord(c->ev[5], s));
    if (c->profiling) {
        CK(cudaEventRecord(c->ev[3], s));
        CK(cudaEventRecord(c->ev[6], s));
    }
    CK(cudaGetLastError());

    // sum_weights = accumulate(weights_); w /= sum (:679-686); particles_ = proposal (:689);
    // expected_pose (:696-716)
    bool next_cdf_built = false;
    if (single_tile(c)) {
        rc = launch_single(c, c->d_wraw, nullptr, c->d_S1, nullptr, 0, dst);
        if (rc) return rc;
        rc = launch_pose(c, c->d_wraw, c->d_S1, c->d_wn, dst, true);
        if (rc) return rc;
        c->tile_state = 0;   // the one-CTA passes need no tile sums; S2 is rebuilt by ensure_cdf
    } else {
        rc = launch_tile_sums(c, c->d_wraw);
        if (rc) return rc;
        rc = launch_pass(c, kPassRaw, dst, dst);
        if (rc) return rc;
        rc = launch_pass(c, kPassNormalise, dst, dst);
        if (rc) return rc;
        c->tile_state = 2;   // d_tile_sum = tile sums of w_raw; w_norm = w_raw / S1; S2 = sum w_norm
    }
    if (c->profiling) {
        CK(cudaEventRecord(c->ev[4], s));
        c->ev_valid = true;
    }
    CK(cudaGetLastError());
    c->cdf_last = src;
    c->cdf_any = true;
    c->cdf_valid = next_cdf_built;
    c->cur = dst;
    c->update_no++;
    return MCL_OK;
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
// state (no diagnostics, weights untouched since the previous update): the ~18 launches become one.
// The graph reads action and scan from the context's own staging buffers, so foreign device pointers
// are first copied there (264 bytes, device to device).  A sharded filter replays a graph too when its
// kernels exchange on their own (xmode 1): there is no host call between its launches.
int update_steady(mcl_ctx* c, const double* action_dev, const float* obs_dev) {
    cudaStream_t s = c->stream;
    // (a captured graph reads the packed copy of the state: an update whose packed source is stale runs directly)
    // steady state: what the captured update finds is what it leaves behind (weights from the last update, no CDF yet)
    const bool steady = (single_tile(c) || c->tile_state == 2) && !c->cdf_valid;
    const bool graph_ok = c->graphs_enabled && !c->profiling && !c->keep_ranges && (!sharded(c) || (c->xmode == 1 && c->connected)) &&
                          steady && c->pose4_ok[c->cur];
    if (!graph_ok) return update_device(c, action_dev, obs_dev, nullptr, nullptr);
    if (action_dev != c->d_action)
        CK(cudaMemcpyAsync(c->d_action, action_dev, sizeof(double) * 3 * c->F, cudaMemcpyDeviceToDevice, s));
    if (obs_dev != c->d_obs)
        CK(cudaMemcpyAsync(c->d_obs, obs_dev, sizeof(float) * c->R * c->F, cudaMemcpyDeviceToDevice, s));
    if (c->gexec[c->cur]) {
        CK(cudaGraphLaunch(c->gexec[c->cur], s));
        c->cdf_last = c->cur;    // what update_device does on the host side
        c->cdf_any = true;
        c->cur ^= 1;
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
    const int tile_state0 = c->tile_state, cdf_last0 = c->cdf_last;
    const bool p4[2] = {c->pose4_ok[0], c->pose4_ok[1]}, cdf_valid0 = c->cdf_valid, cdf_any0 = c->cdf_any;
    int rc = update_device(c, c->d_action, c->d_obs, nullptr, nullptr);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s, &g);
    if (rc == MCL_OK && e == cudaSuccess && g && cudaGraphInstantiate(&c->gexec[parity], g, 0) == cudaSuccess) {
        c->graph_launches = c->launches - before;
        cudaGraphDestroy(g);
        CK(cudaGraphLaunch(c->gexec[parity], s));
        return MCL_OK;
    }
    // capture refused (e.g. an enclosing capture): nothing of the captured update has executed.  Undo its
    // host-side bookkeeping, stop trying, and run this update directly.
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    c->gexec[parity] = nullptr;
    c->graphs_enabled = false;
    c->launches = before;
    c->tile_state = tile_state0;
    c->pose4_ok[0] = p4[0];
    c->pose4_ok[1] = p4[1];
    c->cdf_valid = cdf_valid0;
    c->cdf_any = cdf_any0;
    c->cdf_last = cdf_last0;
    if (rc == MCL_OK) {
        c->cur ^= 1;
        c->update_no--;
    }
    return update_device(c, c->d_action, c->d_obs, nullptr, nullptr);
}

// device-side exchange errors of a sharded filter (a peer that never published, a missing barrier)
int check_shard_error(mcl_ctx* c) {
    if (!c->h_err || *c->h_err == 0) return MCL_OK;
    const int e = *c->h_err;
    *c->h_err = 0;
    return fail(MCL_ERR_CUDA, "sharded exchange failed on rank %d: %s", c->rank,
                e == kShardTimeout ? "a peer did not publish within the time limit"
                : e == kShardMissing ? "a peer's payload was not there when the host-ordered consumer ran"
                                     : "unexpected error code");
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

namespace {

struct ArenaLayout {
    size_t mbox, flag, routed, list_fn, list_add, inbox, bytes;
};
ArenaLayout arena_layout(int world, int64_t n_local) {
    const size_t C = static_cast<size_t>((n_local + kTile - 1) / kTile) * kTileChunks;
    auto up = [](size_t v) { return (v + 255) & ~size_t{255}; };
    ArenaLayout L{};
    size_t o = 0;
    L.mbox = o;
    o = up(o + static_cast<size_t>(2) * world * kMboxSlot);
    L.flag = o;
    o = up(o + sizeof(unsigned long long) * kMaxWorld);
    L.routed = o;
    o = up(o + sizeof(double4) * static_cast<size_t>(n_local) * static_cast<size_t>(world));   // [world servers][N] answers
    L.list_fn = o;
    o = up(o + sizeof(StepFn) * 2 * C);
    L.list_add = o;
    o = up(o + sizeof(double) * 2 * C * kChunk);
    L.inbox = o;
    o = up(o + sizeof(uint32_t) * static_cast<size_t>(world) * static_cast<size_t>(n_local));
    L.bytes = o;
    return L;
}

}  // namespace

static int create_buffers(mcl_ctx* c, const mcl_params* p, int device, int world, int rank) {
    c->prm = *p;
    c->device = device;
    c->F = p->num_filters;
    c->world = world;
    c->rank = rank;
    c->NG = p->max_particles;
    c->N = c->NG / world;
    c->glo = c->N * rank;
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
    for (int b = 0; b < 2; ++b) CK(dalloc(&c->d_cdf2[b], FN));
    for (int b = 0; b < 2; ++b) CK(dalloc(&c->d_mid2[b], static_cast<size_t>(c->F) * (((c->N + kTile - 1) / kTile) * kTileChunks)));
    CK(dalloc(&c->d_idx, FN));
    CK(cudaMemset(c->d_idx, 0, FN * sizeof(int32_t)));
    {   // weights_ = 1/N (:107)
        const int64_t n = static_cast<int64_t>(FN);
        k_fill<<<static_cast<unsigned>((n + 255) / 256), 256>>>(c->d_wn, n, 1.0 / static_cast<double>(c->NG));
        k_fill<<<static_cast<unsigned>((n + 255) / 256), 256>>>(c->d_wraw, n, 1.0 / static_cast<double>(c->NG));
        c->launches += 2;
    }
    c->T = static_cast<int>((c->N + kTile - 1) / kTile);
    c->C = c->T * kTileChunks;
    const size_t FT = static_cast<size_t>(c->F) * c->T, FC = static_cast<size_t>(c->F) * c->C;
    CK(dalloc(&c->d_tile_sum, FT));
    if (!single_tile(c)) {   // (one-tile filters run k_exact_single, which keeps everything in shared memory)
        CK(dalloc(&c->d_chunk_fn, FC));
        CK(dalloc(&c->d_opq_pre, FC));
        CK(dalloc(&c->d_opq_idx, FC));
        CK(dalloc(&c->d_opq_add, FC * kChunk));
        CK(dalloc(&c->d_tile_opq, FT));
        CK(dalloc(&c->d_tile_elem, FT * 3));
        CK(dalloc(&c->d_list_chunk, FC));
        if (world == 1) {   // (a sharded rank keeps them in its exchange arena)
            CK(dalloc(&c->d_list_fn, 2 * FC));
            CK(dalloc(&c->d_list_add, 2 * FC * kChunk));
        }
        CK(dalloc(&c->d_anchors, FC));
        CK(dalloc(&c->d_anchor_val, FC));
        CK(dalloc(&c->d_tile_start, FT));
        CK(dalloc(&c->d_partial, FT * 4));
    }
    {   // coarse CDF level: segments of 64 particles, doubled until at most 16384 entries (128 KB of shared memory)
        c->coarse_shift = 6;
        while ((c->N >> c->coarse_shift) > 16384) ++c->coarse_shift;
        c->coarse_n = static_cast<int>(c->N >> c->coarse_shift);
        for (int b = 0; b < 2; ++b) CK(dalloc(&c->d_coarse2[b], static_cast<size_t>(c->F) * std::max(c->coarse_n, 1)));
    }
    CK(dalloc(&c->d_S1, static_cast<size_t>(c->F)));
    CK(dalloc(&c->d_S2, static_cast<size_t>(c->F)));
    CK(dalloc(&c->d_scratch_total, static_cast<size_t>(c->F)));
    CK(dalloc(&c->d_slice_sum, static_cast<size_t>(kMaxWorld)));
    CK(cudaMemset(c->d_slice_sum, 0, sizeof(double) * kMaxWorld));
    for (int b = 0; b < 2; ++b) {
        CK(dalloc(&c->d_rank_end2[b], static_cast<size_t>(kMaxWorld)));
        CK(cudaMemset(c->d_rank_end2[b], 0, sizeof(double) * kMaxWorld));
    }
    if (single_tile(c)) {
        c->norm_blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((c->N + kNormThreads * 4 - 1) / (kNormThreads * 4),
                                                                                   std::max(1, 4 * c->num_sms / std::min(c->F, 4 * c->num_sms)))));
        CK(dalloc(&c->d_partial, static_cast<size_t>(c->F) * c->norm_blocks * 4));
    } else {
        c->norm_blocks = static_cast<int>(std::min<int64_t>(4 * c->num_sms, (c->N + kNormThreads * 4 - 1) / (kNormThreads * 4)));   // mcl_expected_pose only
        if (static_cast<size_t>(c->norm_blocks) > static_cast<size_t>(c->T)) c->norm_blocks = c->T;   // d_partial is [F][T][4]
        if (c->norm_blocks < 1) c->norm_blocks = 1;
    }
    CK(dalloc(&c->d_pose, static_cast<size_t>(c->F) * 3));
    CK(cudaMemset(c->d_pose, 0, sizeof(double) * 3 * c->F));
    CK(dalloc(&c->d_centre, static_cast<size_t>(c->F) * 2));
    CK(dalloc(&c->d_done, static_cast<size_t>(c->F)));
    CK(cudaMemset(c->d_done, 0, sizeof(unsigned int) * c->F));
    CK(dalloc(&c->d_route_done, size_t{1}));
    CK(cudaMemset(c->d_route_done, 0, sizeof(unsigned int)));
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
        c->obs_capacity = kMaxBeams;
        void* d_in = nullptr;
        CK(cudaMalloc(&d_in, in_bytes));
        c->d_action = static_cast<double*>(d_in);
        c->d_obs = reinterpret_cast<float*>(c->d_action + 3 * c->F);
        CK(cudaMallocHost(reinterpret_cast<void**>(&c->h_action), in_bytes));
        c->h_obs = reinterpret_cast<float*>(c->h_action + 3 * c->F);
        CK(cudaHostAlloc(reinterpret_cast<void**>(&c->h_pose), sizeof(double) * 3 * c->F, cudaHostAllocMapped));
        std::memset(c->h_pose, 0, sizeof(double) * 3 * c->F);
        CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->d_pose_mapped), c->h_pose, 0));
        CK(cudaHostAlloc(reinterpret_cast<void**>(&c->h_err), sizeof(int), cudaHostAllocMapped));
        *c->h_err = 0;
        CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->d_err), c->h_err, 0));
    }
    // exchange state: a single context is "rank 0 of 1" and never publishes
    c->sh.world = world;
    c->sh.rank = rank;
    c->sh.fused = 1;
    c->sh.err = c->d_err;
    if (world > 1) {
        // everything a peer writes (mailbox, flags, routed poses) or may read (overflow lists) lives in ONE
        // allocation, so that one IPC handle maps it
        const ArenaLayout L = arena_layout(world, c->N);
        CK(cudaMalloc(reinterpret_cast<void**>(&c->arena), L.bytes));
        CK(cudaMemset(c->arena, 0, L.bytes));
        c->d_mbox = reinterpret_cast<uint8_t*>(c->arena + L.mbox);
        c->d_flag = reinterpret_cast<unsigned long long*>(c->arena + L.flag);
        c->d_routed = reinterpret_cast<double4*>(c->arena + L.routed);
        c->d_list_fn = reinterpret_cast<StepFn*>(c->arena + L.list_fn);
        c->d_list_add = reinterpret_cast<double*>(c->arena + L.list_add);
        c->d_inbox = reinterpret_cast<uint32_t*>(c->arena + L.inbox);
        CK(dalloc(&c->d_where, static_cast<size_t>(c->N)));
        CK(dalloc(&c->d_req_count, static_cast<size_t>(kMaxWorld)));
        CK(cudaMemset(c->d_req_count, 0, sizeof(unsigned int) * kMaxWorld));
        CK(dalloc(&c->d_xseq, size_t{1}));
        CK(cudaMemset(c->d_xseq, 0, sizeof(unsigned long long)));
        CK(dalloc(&c->d_nccl_tok, static_cast<size_t>(kMaxWorld)));
        CK(cudaMemset(c->d_nccl_tok, 0, sizeof(unsigned long long) * kMaxWorld));
        c->sh.xseq = c->d_xseq;
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
    CK(cudaFuncSetAttribute(k_route_serve, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    CK(cudaFuncSetAttribute(k_route, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            16384 * static_cast<int>(sizeof(double)) + (kRouteThreads / 32) * kRouteQueue * 12));
    CK(cudaFuncSetAttribute(k_weight_steps_sm<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    CK(cudaFuncSetAttribute(k_weight_steps_sm<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(k_raycast_dir<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWindowBudget) - 10 * 1024));
    CK(cudaFuncSetAttribute(k_raycast_dir<207>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWindowBudget) - 10 * 1024));
    CK(cudaFuncSetAttribute(k_raycast_dir<238>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWindowBudget) - 10 * 1024));
    CK(cudaFuncSetAttribute(k_raycast_dir<239>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWindowBudget) - 10 * 1024));
    CK(cudaDeviceSynchronize());
    return MCL_OK;
}

static int create_any(const mcl_params* p, int device, int world, int rank, mcl_ctx** out) {
    if (!p || !out) return fail(MCL_ERR_INVALID, "null argument");
    *out = nullptr;
    if (p->max_particles < 1) return fail(MCL_ERR_INVALID, "max_particles must be >= 1");
    if (p->num_filters < 1) return fail(MCL_ERR_INVALID, "num_filters must be >= 1");
    if (!(p->squash_factor > 0) || !(p->max_range > 0)) return fail(MCL_ERR_INVALID, "squash_factor / max_range must be positive");
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return fail(MCL_ERR_INVALID, "rank %d / world %d outside [1,%d]", rank, world, kMaxWorld);
    if (world > 1 && p->num_filters != 1) return fail(MCL_ERR_INVALID, "particle sharding applies to a single filter, not a batch");
    if (world > 1 && (p->max_particles % world) != 0)
        return fail(MCL_ERR_INVALID, "max_particles (%d) must be a multiple of the number of ranks (%d)", p->max_particles, world);
    if (world > 1 && static_cast<int64_t>(p->max_particles) / world >= (int64_t{1} << kWhereShift))
        return fail(MCL_ERR_INVALID, "a rank of a sharded filter holds fewer than 2^%d particles", kWhereShift);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MCL_ERR_NO_DEVICE, "no CUDA device visible; the MCL update has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(MCL_ERR_INVALID, "device %d not in [0,%d)", device, ndev);
    CK(cudaSetDevice(device));
    auto* c = new mcl_ctx();
    const int rc = create_buffers(c, p, device, world, rank);
    if (rc != MCL_OK) {
        const std::string keep = g_err;   // mcl_destroy must not clobber the reason
        mcl_destroy(c);
        g_err = keep;
        return rc;
    }
    *out = c;
    return MCL_OK;
}

int mcl_create(const mcl_params* p, int device, mcl_ctx** out) { return create_any(p, device, 1, 0, out); }

int mcl_destroy(mcl_ctx* c) {
    if (!c) return MCL_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    void* ptrs[] = {c->d_pose4[0], c->d_pose4[1],
                    c->d_grid, c->d_v8, c->d_v4, c->d_gap, c->d_free, c->d_tabT, c->d_step2idx, c->d_px[0], c->d_px[1], c->d_py[0],
                    c->d_py[1], c->d_pt[0], c->d_pt[1], c->d_wraw, c->d_wn, c->d_cdf2[0], c->d_cdf2[1], c->d_mid2[0], c->d_mid2[1], c->d_idx, c->d_steps, c->d_u, c->d_z,
                    c->d_action, c->d_slice, c->d_tile_sum, c->d_chunk_fn, c->d_opq_pre, c->d_opq_idx, c->d_opq_add,
                    c->d_tile_opq, c->d_tile_elem, c->d_list_chunk, c->arena ? nullptr : c->d_list_fn, c->arena ? nullptr : c->d_list_add,
                    c->arena, c->d_anchors, c->d_anchor_val,
                    c->d_tile_start, c->d_coarse2[0], c->d_coarse2[1], c->d_S1, c->d_S2, c->d_scratch_total, c->d_slice_sum, c->d_rank_end2[0],
                    c->d_rank_end2[1],
                    c->d_partial, c->d_pose, c->d_centre, c->d_replays, c->d_hist, c->d_perm, c->d_done, c->d_route_done,
                    c->d_xseq, c->d_nccl_tok, c->d_tmp, c->d_update_no, c->d_dbg, c->d_beam, c->d_steps16, c->d_where, c->d_req_count};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    free_dir(c);
    drop_graphs(c);
    for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    if (c->h_action) cudaFreeHost(c->h_action);
    if (c->h_pose) cudaFreeHost(c->h_pose);
    if (c->h_err) cudaFreeHost(c->h_err);
    for (auto& e : c->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->mark_ev)
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
    if (M < 1 || M > 65534)
        return fail(MCL_ERR_UNSUPPORTED, "MAX_RANGE_PX = %d outside [1,65534] (max_range %.3f / resolution %.6f)", M,
                    c->prm.max_range, res);
    if (!skip_map_layout(data, width, height, c->skip)) return fail(MCL_ERR_UNSUPPORTED, "grid %dx%d too large", width, height);
    CK(cudaStreamSynchronize(c->stream));
    c->res = res;
    c->ox = ox;
    c->oy = oy;
    c->oyaw = oyaw;   // read by the reference (:192-193) but never used by the march (:628-629)
    c->M = M;
    for (void* p : {static_cast<void*>(c->d_grid), static_cast<void*>(c->d_v8), static_cast<void*>(c->d_v4),
                    static_cast<void*>(c->d_gap), static_cast<void*>(c->d_free), static_cast<void*>(c->d_step2idx)})
        if (p) cudaFree(p);
    c->d_grid = nullptr;
    c->d_v8 = c->d_v4 = nullptr;
    c->d_gap = nullptr;
    c->d_free = nullptr;
    c->d_step2idx = nullptr;
    const size_t cells = static_cast<size_t>(width) * height;
    CK(dalloc(&c->d_grid, cells));
    CK(cudaMemcpy(c->d_grid, data, cells, cudaMemcpyHostToDevice));
    {   // skip codes and gap map: one exact Euclidean distance transform on the device (host twin: map_prep.cpp)
        const int rc = device_build_maps(c, true);
        if (rc) return rc;
    }
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
    rc = update_wide(c);
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
    if (n < 1 || n > kMaxBeamsWide) return fail(MCL_ERR_UNSUPPORTED, "%d beams outside [1,%d]", n, kMaxBeamsWide);
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    drop_graphs(c);
    c->R = n;
    c->beam_angles.assign(angles, angles + n);
    c->beams.R = n <= kMaxBeams ? n : 0;
    for (int j = 0; j < n && j < kMaxBeams; ++j) {
        c->beams.angle[j] = angles[j];
        c->beams.cosa[j] = std::cos(static_cast<double>(angles[j]));
        c->beams.sina[j] = std::sin(static_cast<double>(angles[j]));
    }
    if (c->d_beam) cudaFree(c->d_beam);
    c->d_beam = nullptr;
    CK(dalloc(&c->d_beam, static_cast<size_t>(n)));
    CK(cudaMemcpy(c->d_beam, angles, sizeof(float) * n, cudaMemcpyHostToDevice));
    if (static_cast<size_t>(n) > c->obs_capacity) {   // staging of the host-facing update: [F][3] doubles | [F][n] floats
        const size_t in_bytes = sizeof(double) * 3 * c->F + sizeof(float) * static_cast<size_t>(n) * c->F;
        if (c->d_action) cudaFree(c->d_action);
        if (c->h_action) cudaFreeHost(c->h_action);
        c->d_action = nullptr;
        c->h_action = nullptr;
        void* d_in = nullptr;
        CK(cudaMalloc(&d_in, in_bytes));
        c->d_action = static_cast<double*>(d_in);
        c->d_obs = reinterpret_cast<float*>(c->d_action + 3 * c->F);
        CK(cudaMallocHost(reinterpret_cast<void**>(&c->h_action), in_bytes));
        c->h_obs = reinterpret_cast<float*>(c->h_action + 3 * c->F);
        c->obs_capacity = static_cast<size_t>(n);
    }
    c->have_beams = true;
    if (c->d_steps) {
        cudaFree(c->d_steps);
        c->d_steps = nullptr;
    }
    int rc = update_wide(c);
    if (rc) return rc;
    if (c->keep_ranges && !c->wide) CK(dalloc(&c->d_steps, static_cast<size_t>(c->F) * c->N * c->R));
    rc = ensure_slice(c);
    if (rc) return rc;
    return ensure_dir(c, false);
}

int mcl_num_free_cells(const mcl_ctx* c) { return c ? static_cast<int>(c->skip.free_cells.size()) : 0; }

// initialize_particles_pose :382-399.  normals_3n: the context's own particles (a sharded rank: its slice).
int mcl_init_pose(mcl_ctx* c, int filter, const double pose[3], const double* normals) {
    int rc = check_filter(c, filter, true);
    if (rc) return rc;
    if (!pose) return fail(MCL_ERR_INVALID, "null pose");
    CK(cudaSetDevice(c->device));
    const int f0 = filter < 0 ? 0 : filter, nf = filter < 0 ? c->F : 1;
    const size_t pose_bytes = sizeof(double) * 3 * nf;
    const size_t norm_bytes = normals ? sizeof(double) * 3 * c->N * nf : 0;
    rc = ensure_tmp(c, 256 + pose_bytes + norm_bytes);
    if (rc) return rc;
    double* d_pose = static_cast<double*>(c->d_tmp);
    double* d_norm = normals ? reinterpret_cast<double*>(static_cast<char*>(c->d_tmp) + ((pose_bytes + 255) & ~size_t{255})) : nullptr;
    std::vector<double> hp(static_cast<size_t>(3) * nf);
    for (int k = 0; k < nf; ++k) std::memcpy(&hp[3 * k], pose, 3 * sizeof(double));
    CK(cudaMemcpyAsync(d_pose, hp.data(), pose_bytes, cudaMemcpyHostToDevice, c->stream));
    if (normals) {
        for (int k = 0; k < nf; ++k)   // the same injected stream for every addressed filter
            CK(cudaMemcpyAsync(d_norm + static_cast<size_t>(3) * c->N * k, normals, sizeof(double) * 3 * c->N,
                               cudaMemcpyHostToDevice, c->stream));
    }
    InitArgs a{};
    a.N = c->N;
    a.glo = c->glo;
    a.w0 = 1.0 / static_cast<double>(c->NG);
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
    CK(cudaStreamSynchronize(c->stream));   // hp is read by the copy above
    c->cdf_valid = false;
    c->cdf_any = false;
    c->tile_state = 0;
    c->pose4_ok[c->cur] = false;   // the packed copy no longer matches the state arrays
    return MCL_OK;
}

// initialize_global :401-446.  Injected draws: the context's own particles.
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
        const size_t n = static_cast<size_t>(c->N) * nf;
        rc = ensure_tmp(c, n * (sizeof(double) + sizeof(int32_t)) + 256);
        if (rc) return rc;
        d_theta = static_cast<double*>(c->d_tmp);
        d_cell = reinterpret_cast<int32_t*>(d_theta + n);
        for (int k = 0; k < nf; ++k) {
            CK(cudaMemcpyAsync(d_cell + static_cast<size_t>(c->N) * k, cell, sizeof(int32_t) * c->N, cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(d_theta + static_cast<size_t>(c->N) * k, theta, sizeof(double) * c->N, cudaMemcpyHostToDevice, c->stream));
        }
    }
    InitArgs a{};
    a.N = c->N;
    a.glo = c->glo;
    a.w0 = 1.0 / static_cast<double>(c->NG);
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
    c->cdf_valid = false;
    c->cdf_any = false;
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
    c->cdf_any = false;
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
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    if (!c->cdf_any) return fail(MCL_ERR_INVALID, "no CDF yet: call mcl_update first");
    return get_array(c, filter, c->d_cdf2[c->cdf_last], sizeof(double), c->N, out);   // the _M_cp the last update drew from
}
int mcl_get_resample_indices(mcl_ctx* c, int filter, int32_t* out) {
    return get_array(c, filter, c ? c->d_idx : nullptr, sizeof(int32_t), c ? c->N : 0, out);
}
int mcl_get_range_steps(mcl_ctx* c, int filter, uint8_t* out) {
    if (c && c->wide) return fail(MCL_ERR_UNSUPPORTED, "MAX_RANGE_PX %d / %d beams: step indices need 16 bits, use mcl_get_range_steps16", c->M, c->R);
    return get_array(c, filter, c ? c->d_steps : nullptr, 1, c ? static_cast<size_t>(c->N) * c->R : 0, out);
}
int mcl_get_range_steps16(mcl_ctx* c, int filter, uint16_t* out) {
    int rc = check_filter(c, filter, false);
    if (rc) return rc;
    const size_t n = static_cast<size_t>(c->N) * c->R;
    if (c->wide) return get_array(c, filter, c->d_steps16, 2, n, out);
    if (!out) return fail(MCL_ERR_INVALID, "null output");
    if (!c->d_steps) return fail(MCL_ERR_INVALID, "ranges are not kept: call mcl_set_keep_ranges(ctx, 1) before the update");
    CK(cudaSetDevice(c->device));
    rc = ensure_tmp(c, 2 * n);
    if (rc) return rc;
    uint16_t* d16 = static_cast<uint16_t*>(c->d_tmp);
    k_widen_steps<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->stream>>>(c->d_steps + n * filter, static_cast<int64_t>(n), d16);
    c->launches++;
    CK(cudaMemcpyAsync(out, d16, 2 * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MCL_OK;
}

int mcl_get_ranges(mcl_ctx* c, int filter, float* out) {
    int rc = check_filter(c, filter, false);
    if (rc) return rc;
    if (!out) return fail(MCL_ERR_INVALID, "null output");
    if (!c->wide && !c->d_steps) return fail(MCL_ERR_INVALID, "ranges are not kept: call mcl_set_keep_ranges(ctx, 1) before the update");
    CK(cudaSetDevice(c->device));
    const int64_t n = c->N * c->R;
    rc = ensure_tmp(c, sizeof(float) * static_cast<size_t>(n));
    if (rc) return rc;
    float* d_out = static_cast<float*>(c->d_tmp);
    if (c->wide)
        k_steps16_to_ranges<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->stream>>>(c->d_steps16 + n * filter, n, c->M, c->res,
                                                                                             c->prm.max_range, d_out);
    else
        k_steps_to_ranges<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->stream>>>(c->d_steps + n * filter, n, c->M, c->res,
                                                                                           c->prm.max_range, d_out);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MCL_OK;
}

int mcl_update_dev(mcl_ctx* c, const double* action_dev, const float* obs_dev, int num_beams) {
    if (!c || !action_dev || !obs_dev) return fail(MCL_ERR_INVALID, "null argument");
    if (num_beams != c->R) return fail(MCL_ERR_INVALID, "num_beams %d != configured %d", num_beams, c->R);
    CK(cudaSetDevice(c->device));
    return update_steady(c, action_dev, obs_dev);
}

int mcl_update_dev_noise(mcl_ctx* c, const double* action_dev, const float* obs_dev, int num_beams, const double* u_dev,
                         const double* z_dev) {
    if (!c || !action_dev || !obs_dev) return fail(MCL_ERR_INVALID, "null argument");
    if (num_beams != c->R) return fail(MCL_ERR_INVALID, "num_beams %d != configured %d", num_beams, c->R);
    CK(cudaSetDevice(c->device));
    return update_device(c, action_dev, obs_dev, u_dev, z_dev);
}

int mcl_read_pose(mcl_ctx* c, double* pose_out) {
    if (!c || !pose_out) return fail(MCL_ERR_INVALID, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->h_pose, c->d_pose, sizeof(double) * 3 * c->F, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::memcpy(pose_out, c->h_pose, sizeof(double) * 3 * c->F);
    return check_shard_error(c);
}

int mcl_synchronize(mcl_ctx* c) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return check_shard_error(c);
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
        // injected noise covers the WHOLE filter (a sharded rank indexes it by the global slot)
        const size_t NG = static_cast<size_t>(sharded(c) ? c->NG : c->N);
        if (any_u) {
            if (!c->d_u) CK(dalloc(&c->d_u, NG * c->F));
            for (int f = 0; f < c->F; ++f)
                CK(cudaMemcpyAsync(c->d_u + NG * f, noise[f].u_resample, NG * sizeof(double), cudaMemcpyHostToDevice, s));
            u_dev = c->d_u;
        }
        if (any_z) {
            if (!c->d_z) CK(dalloc(&c->d_z, 3 * NG * c->F));
            for (int f = 0; f < c->F; ++f)
                CK(cudaMemcpyAsync(c->d_z + 3 * NG * f, noise[f].z_motion, 3 * NG * sizeof(double), cudaMemcpyHostToDevice, s));
            z_dev = c->d_z;
        }
    }
    int rc = noise ? update_device(c, c->d_action, c->d_obs, u_dev, z_dev) : update_steady(c, c->d_action, c->d_obs);
    if (rc) return rc;
    // the pose kernel has written the pose into h_pose (mapped pinned memory): no D2H copy call
    CK(cudaStreamSynchronize(s));
    rc = check_shard_error(c);
    if (rc) return rc;
    if (pose_out) std::memcpy(pose_out, c->h_pose, sizeof(double) * 3 * c->F);
    if (c->profiling && c->ev_valid) read_stage_times(c);
    return MCL_OK;
}

int mcl_expected_pose(mcl_ctx* c, int filter, double pose_out[3]) {
    int rc = check_filter(c, filter, false);
    if (rc) return rc;
    if (!pose_out) return fail(MCL_ERR_INVALID, "null output");
    if (sharded(c)) return fail(MCL_ERR_UNSUPPORTED, "mcl_expected_pose on a sharded context: the pose of the whole filter is the update's result (mcl_read_pose)");
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
    const size_t qbytes = (sizeof(double) * 3 * static_cast<size_t>(n) + 255) & ~size_t{255};
    const int rc = ensure_tmp(c, qbytes + sizeof(float) * static_cast<size_t>(n));
    if (rc) return rc;
    double* d_q = static_cast<double*>(c->d_tmp);
    float* d_o = reinterpret_cast<float*>(static_cast<char*>(c->d_tmp) + qbytes);
    CK(cudaMemcpyAsync(d_q, q, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    if (c->wide) {
        WideQueryArgs wq{RefGrid{c->d_grid, c->map.W, c->map.H, c->res, c->ox, c->oy}, c->M, d_q, n, d_o, c->prm.max_range};
        k_range_queries_wide<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->stream>>>(wq);
    } else {
        QueryArgs a{};
        a.map = c->map;
        a.q = d_q;
        a.n = n;
        a.out = d_o;
        a.max_range = c->prm.max_range;
        k_range_queries<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->stream>>>(a);
    }
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_o, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MCL_OK;
}

int mcl_cast_ray(mcl_ctx* c, double x, double y, double angle, float* out) {
    const double q[3] = {x, y, angle};
    return mcl_calc_range_many(c, q, 1, out);
}

// visualize() :946-958: k draws from discrete_distribution(weights_) over the current particles.
// u: the canonical uniforms the reference's generator would produce (NULL: device RNG);
// idx_out (nullable): the drawn particle indices.
int mcl_sample_particles_u(mcl_ctx* c, int filter, int k, const double* u, double* out, int32_t* idx_out) {
    int rc = check_filter(c, filter, false);
    if (rc) return rc;
    if (!out || k < 1) return fail(MCL_ERR_INVALID, "bad arguments");
    if (sharded(c)) return fail(MCL_ERR_UNSUPPORTED, "viz sampling of a sharded filter: gather the state first (mcl_sharded_gather)");
    CK(cudaSetDevice(c->device));
    rc = ensure_cdf(c);   // the CDF of the current weights, as visualize() builds it (:949); reused by the next update
    if (rc) return rc;
    const size_t obytes = (sizeof(double) * 3 * k + 255) & ~size_t{255}, ubytes = (sizeof(double) * k + 255) & ~size_t{255};
    rc = ensure_tmp(c, obytes + ubytes + sizeof(int32_t) * k);
    if (rc) return rc;
    double* d_o = static_cast<double*>(c->d_tmp);
    double* d_u = reinterpret_cast<double*>(static_cast<char*>(c->d_tmp) + obytes);
    int32_t* d_i = reinterpret_cast<int32_t*>(static_cast<char*>(c->d_tmp) + obytes + ubytes);
    if (u) CK(cudaMemcpyAsync(d_u, u, sizeof(double) * k, cudaMemcpyHostToDevice, c->stream));
    const size_t fo = static_cast<size_t>(c->N) * filter;
    k_sample_particles<<<(k + 127) / 128, 128, 0, c->stream>>>(c->d_cdf2[c->cur] + fo, c->N, c->d_px[c->cur] + fo, c->d_py[c->cur] + fo,
                                                               c->d_pt[c->cur] + fo, k, c->prm.seed, ++c->init_no, u ? d_u : nullptr,
                                                               d_o, d_i);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_o, sizeof(double) * 3 * k, cudaMemcpyDeviceToHost, c->stream));
    if (idx_out) CK(cudaMemcpyAsync(idx_out, d_i, sizeof(int32_t) * k, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MCL_OK;
}

int mcl_sample_particles(mcl_ctx* c, int filter, int k, double* out) { return mcl_sample_particles_u(c, filter, k, nullptr, out, nullptr); }

int mcl_set_profiling(mcl_ctx* c, int enabled) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    c->profiling = enabled != 0;
    c->ev_valid = false;
    c->nmarks = 0;
    return MCL_OK;
}

int mcl_get_stage_ms(mcl_ctx* c, mcl_stage_ms* out) {
    if (!c || !out) return fail(MCL_ERR_INVALID, "null argument");
    if (c->profiling && c->ev_valid) {
        // the device-resident entry points do not synchronise: wait for the last event here
        CK(cudaSetDevice(c->device));
        CK(cudaEventSynchronize(c->ev[4]));
        read_stage_times(c);
    }
    *out = c->last_ms;
    return MCL_OK;
}

// Device time of every kernel of the last profiled update, in launch order.  names_out: capacity x 48
// bytes (NUL-terminated); ms_out: capacity floats; *count = kernels of the update.
int mcl_get_kernel_ms(mcl_ctx* c, char* names_out, float* ms_out, int capacity, int* count) {
    if (!c || !count) return fail(MCL_ERR_INVALID, "null argument");
    *count = c->nmarks;
    if (!c->profiling || c->nmarks == 0) return MCL_OK;
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->mark_ev[c->nmarks]));
    for (int i = 0; i < c->nmarks && i < capacity; ++i) {
        float t = 0;
        cudaEventElapsedTime(&t, c->mark_ev[i], c->mark_ev[i + 1]);
        if (ms_out) ms_out[i] = t;
        if (names_out) {
            std::strncpy(names_out + static_cast<size_t>(i) * 48, c->mark_name[i], 47);
            names_out[static_cast<size_t>(i) * 48 + 47] = 0;
        }
    }
    return MCL_OK;
}

// Diagnostics: SM cycle counts of the phases of ONE kind of exact pass (0 S1, 1 normalise+pose+S2, 2 S2 of stored
// weights, 3 cdf) in the updates that follow; pass_kind < 0 switches it off.  out[8]: slowest CTA's tile phase |
// last CTA until it knows it is last | pose fold | tile scan | ordered opaque list | exchange | serial evaluation +
// tile starts | opaque chunks.  Call with out != NULL after an update to read (and clear) them.
int mcl_debug_pass_cycles(mcl_ctx* c, int pass_kind, unsigned long long* out) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    if (!c->d_dbg) {
        CK(dalloc(&c->d_dbg, size_t{16}));
        CK(cudaMemset(c->d_dbg, 0, 16 * sizeof(unsigned long long)));
    }
    if (out) {
        CK(cudaMemcpy(out, c->d_dbg + (pass_kind >= 8 ? 8 : 0), 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        CK(cudaMemset(c->d_dbg, 0, 16 * sizeof(unsigned long long)));
    }
    c->dbg_pass = pass_kind;
    drop_graphs(c);
    return MCL_OK;
}

int mcl_set_keep_ranges(mcl_ctx* c, int enabled) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    c->keep_ranges = enabled != 0;
    if (c->keep_ranges && !c->wide && !c->d_steps && c->R > 0) CK(dalloc(&c->d_steps, static_cast<size_t>(c->F) * c->N * c->R));
    return MCL_OK;
}

int mcl_kernel_launches(mcl_ctx* c, int64_t* count) {
    if (!c || !count) return fail(MCL_ERR_INVALID, "null argument");
    *count = c->launches;
    return MCL_OK;
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

int mcl_set_pdl(mcl_ctx* c, int enabled) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->pdl = enabled != 0;
    drop_graphs(c);
    return MCL_OK;
}

int mcl_set_ray_mode(mcl_ctx* c, int mode) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    if (mode < 0 || mode > 2) return fail(MCL_ERR_INVALID, "ray mode %d not in {0 auto, 1 isotropic, 2 directional}", mode);
    if (mode == 2 && (!c->dir_ready || c->wide))
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
    if (last_mode) *last_mode = c->last_dir ? 1 : 0;
    if (box_cells) *box_cells = c->dir_ready ? c->dir_box : 0;
    if (units) *units = plan[kPlanUnits];
    return MCL_OK;
}

int mcl_get_dir_map(mcl_ctx* c, int sector, uint8_t* out, int* pw, int* ph) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    if (sector == -1) {   // the isotropic skip codes
        if (!c->have_map) return fail(MCL_ERR_NO_MAP, "map not set");
        if (pw) *pw = c->skip.PW;
        if (ph) *ph = c->skip.PH;
        if (out) {
            CK(cudaSetDevice(c->device));
            CK(cudaStreamSynchronize(c->stream));
            CK(cudaMemcpy(out, c->d_v8, static_cast<size_t>(c->skip.PW) * c->skip.PH, cudaMemcpyDeviceToHost));
        }
        return MCL_OK;
    }
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

#if MCL_DIR_DIAG
extern "C" int mcl_debug_dir_diag(mcl_ctx* c, unsigned long long* out) {   // development builds only
    if (!c || !out) return fail(MCL_ERR_INVALID, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpyFromSymbol(out, g_dir_diag, sizeof(unsigned long long) * kDirMaxRanges * 8));
    return MCL_OK;
}
#endif

int mcl_set_stream(mcl_ctx* c, void* stream) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->stream = stream ? static_cast<cudaStream_t>(stream) : c->own_stream;
    drop_graphs(c);
    return MCL_OK;
}

/* ---- particle-sharded filter ------------------------------------------------------------- */

namespace {

// every rank's exchange arena (own included) -> the pointer tables the kernels take by value
int install_arenas(mcl_ctx* c, void* const* bases) {
    const ArenaLayout L = arena_layout(c->world, c->N);
    for (int q = 0; q < c->world; ++q) {
        char* b = static_cast<char*>(bases[q]);
        c->sh.mbox[q] = reinterpret_cast<uint8_t*>(b + L.mbox);
        c->sh.flag[q] = reinterpret_cast<unsigned long long*>(b + L.flag);
        c->routed_peers[q] = reinterpret_cast<double4*>(b + L.routed);
        c->inbox_peers[q] = reinterpret_cast<uint32_t*>(b + L.inbox);
        c->peer_list_fn[q] = reinterpret_cast<const StepFn*>(b + L.list_fn);
        c->peer_list_add[q] = reinterpret_cast<const double*>(b + L.list_add);
    }
    c->connected = true;
    drop_graphs(c);
    return MCL_OK;
}

}  // namespace

int mcl_shard_create(const mcl_params* p, int device, int world, int rank, mcl_ctx** out) {
    if (world < 2) return fail(MCL_ERR_INVALID, "a sharded filter has at least 2 ranks (use mcl_create for one GPU)");
    return create_any(p, device, world, rank, out);
}

int mcl_shard_export(mcl_ctx* c, void* blob, size_t capacity) {
    if (!c || !blob) return fail(MCL_ERR_INVALID, "null argument");
    if (!c->arena) return fail(MCL_ERR_INVALID, "not a sharded context");
    if (capacity < MCL_SHARD_BLOB_BYTES) return fail(MCL_ERR_INVALID, "need %d bytes", MCL_SHARD_BLOB_BYTES);
    CK(cudaSetDevice(c->device));
    std::memset(blob, 0, MCL_SHARD_BLOB_BYTES);
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->arena));
    std::memcpy(blob, &h, sizeof h);
    const int64_t meta[3] = {c->world, c->N, static_cast<int64_t>(arena_layout(c->world, c->N).bytes)};
    std::memcpy(static_cast<char*>(blob) + sizeof h, meta, sizeof meta);
    return MCL_OK;
}

int mcl_shard_connect(mcl_ctx* c, const void* blobs) {
    if (!c || !blobs) return fail(MCL_ERR_INVALID, "null argument");
    if (!c->arena) return fail(MCL_ERR_INVALID, "not a sharded context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    void* bases[kMaxWorld] = {};
    for (int q = 0; q < c->world; ++q) {
        const char* b = static_cast<const char*>(blobs) + static_cast<size_t>(q) * MCL_SHARD_BLOB_BYTES;
        int64_t meta[3];
        std::memcpy(meta, b + sizeof(cudaIpcMemHandle_t), sizeof meta);
        if (meta[0] != c->world || meta[1] != c->N)
            return fail(MCL_ERR_INVALID, "rank %d was created with world %lld / %lld particles per rank, this rank with %d / %lld", q,
                        (long long)meta[0], (long long)meta[1], c->world, (long long)c->N);
        if (q == c->rank) {
            bases[q] = c->arena;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, b, sizeof h);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->ipc_opened.push_back(p);
        bases[q] = p;
    }
    return install_arenas(c, bases);
}

int mcl_shard_connect_local(mcl_ctx* c, mcl_ctx* const* ranks) {
    if (!c || !ranks) return fail(MCL_ERR_INVALID, "null argument");
    if (!c->arena) return fail(MCL_ERR_INVALID, "not a sharded context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    void* bases[kMaxWorld] = {};
    for (int q = 0; q < c->world; ++q) {
        const mcl_ctx* r = ranks[q];
        if (!r || !r->arena || r->world != c->world || r->N != c->N || r->rank != q)
            return fail(MCL_ERR_INVALID, "ranks[%d] is not rank %d of the same sharded filter", q, q);
        if (r->device != c->device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, c->device, r->device));
            if (!can) return fail(MCL_ERR_UNSUPPORTED, "device %d cannot access device %d", c->device, r->device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(r->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
            cudaGetLastError();
        }
        bases[q] = r->arena;
    }
    return install_arenas(c, bases);
}

int mcl_shard_set_exchange(mcl_ctx* c, int fused, mcl_barrier_fn hook, void* user) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    if (!c->arena) return fail(MCL_ERR_INVALID, "not a sharded context");
    if (!fused && !hook && !c->comm) return fail(MCL_ERR_INVALID, "host-ordered exchange needs a barrier hook or an NCCL communicator");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->xmode = fused ? 1 : 0;
    c->hook = fused ? nullptr : hook;
    c->hook_user = user;
    drop_graphs(c);
    return MCL_OK;
}

int mcl_shard_set_route(mcl_ctx* c, int two_hop) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    if (!c->arena) return fail(MCL_ERR_INVALID, "not a sharded context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->route_mode = two_hop < 0 ? -1 : (two_hop ? 0 : 1);
    drop_graphs(c);
    return MCL_OK;
}

int mcl_shard_info(const mcl_ctx* c, int* world, int* rank, int64_t* n_local, int64_t* n_global) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    if (world) *world = c->world;
    if (rank) *rank = c->rank;
    if (n_local) *n_local = c->N;
    if (n_global) *n_global = c->NG;
    return MCL_OK;
}

int mcl_nccl_unique_id(void* id_out, size_t capacity) {
    if (!id_out || capacity < sizeof(ncclUniqueId)) return fail(MCL_ERR_INVALID, "need %zu bytes", sizeof(ncclUniqueId));
    const int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof id);
    return MCL_OK;
}

int mcl_create_sharded(const mcl_params* p, int device, int world, int rank, const void* nccl_unique_id, mcl_ctx** out) {
    if (!nccl_unique_id) return fail(MCL_ERR_INVALID, "null NCCL id");
    int rc = load_nccl();
    if (rc) return rc;
    rc = mcl_shard_create(p, device, world, rank, out);
    if (rc) return rc;
    mcl_ctx* c = *out;
    auto bail = [&](int code) {
        const std::string keep = g_err;
        mcl_destroy(c);
        *out = nullptr;
        g_err = keep;
        return code;
    };
    ncclUniqueId id;
    std::memcpy(&id, nccl_unique_id, sizeof id);
    if (g_nccl.CommInitRank(&c->comm, world, id, rank) != ncclSuccess) {
        c->comm = nullptr;
        fail(MCL_ERR_CUDA, "ncclCommInitRank failed for rank %d of %d", rank, world);
        return bail(MCL_ERR_CUDA);
    }
    // all-gather the ranks' exchange-arena handles over the communicator it owns, then map them
    std::vector<char> blobs(static_cast<size_t>(world) * MCL_SHARD_BLOB_BYTES);
    rc = mcl_shard_export(c, blobs.data() + static_cast<size_t>(rank) * MCL_SHARD_BLOB_BYTES, MCL_SHARD_BLOB_BYTES);
    if (rc) return bail(rc);
    rc = ensure_tmp(c, blobs.size());
    if (rc) return bail(rc);
    char* d = static_cast<char*>(c->d_tmp);
    if (cudaMemcpyAsync(d + static_cast<size_t>(rank) * MCL_SHARD_BLOB_BYTES, blobs.data() + static_cast<size_t>(rank) * MCL_SHARD_BLOB_BYTES,
                        MCL_SHARD_BLOB_BYTES, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        g_nccl.AllGather(d + static_cast<size_t>(rank) * MCL_SHARD_BLOB_BYTES, d, MCL_SHARD_BLOB_BYTES, ncclChar, c->comm, c->stream) != ncclSuccess ||
        cudaMemcpyAsync(blobs.data(), d, blobs.size(), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        fail(MCL_ERR_CUDA, "exchanging the arena handles over NCCL failed: %s", cudaGetErrorString(cudaGetLastError()));
        return bail(MCL_ERR_CUDA);
    }
    rc = mcl_shard_connect(c, blobs.data());
    if (rc) return bail(rc);
    return MCL_OK;
}

// Whole-filter read-back on every rank (parity harness, visualize()): ncclAllGather of the four
// per-particle arrays on the library's stream.  particles_colmajor: NG x 3; either may be NULL.
int mcl_sharded_gather(mcl_ctx* c, double* particles_colmajor, double* weights) {
    if (!c) return fail(MCL_ERR_INVALID, "null context");
    if (!c->comm) return fail(MCL_ERR_INVALID, "mcl_sharded_gather needs the communicator of mcl_create_sharded");
    CK(cudaSetDevice(c->device));
    const size_t NG = static_cast<size_t>(c->NG), n = static_cast<size_t>(c->N);
    const int rc = ensure_tmp(c, sizeof(double) * 4 * NG);
    if (rc) return rc;
    double* g = static_cast<double*>(c->d_tmp);
    const double* srcs[4] = {c->d_px[c->cur], c->d_py[c->cur], c->d_pt[c->cur], c->d_wn};
    for (int k = 0; k < 4; ++k) NK(g_nccl.AllGather(srcs[k], g + k * NG, n, ncclDouble, c->comm, c->stream));
    if (particles_colmajor) CK(cudaMemcpyAsync(particles_colmajor, g, sizeof(double) * 3 * NG, cudaMemcpyDeviceToHost, c->stream));
    if (weights) CK(cudaMemcpyAsync(weights, g + 3 * NG, sizeof(double) * NG, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return check_shard_error(c);
}

}  // extern "C"
