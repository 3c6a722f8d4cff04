// csrc/exact_kernels.cuh -- kernels of the sequential-order FP64 sums (algebra: exact_sum.cuh).
//
// The reference needs three sequentially rounded reductions per update:
//   S1  = accumulate(weights)                     MCL step 4          (src/particle_filter.cpp:679)
//   S2  = accumulate(weights / S1)                discrete_distribution's own sum (random.tcc:2666)
//   cdf = partial_sum((weights / S1) / S2)        its _M_cp           (random.tcc:2672-2677)
// Each is one PASS over the particles:
//   k_tile_sums     approximate sums of 4096-element tiles (they tell the exact pass which binade a running sum is in)
//   k_exact_pass    one CTA per tile: 8-addend chunk step maps, their block scan, the tile's
//                   opaque chunks compacted.  The LAST CTA to finish then (a) orders the filter's
//                   opaque chunks with their incoming maps, (b) -- sharded filter -- publishes that
//                   summary to every rank and waits for theirs (shard.cuh), (c) evaluates the few
//                   opaque chunks of ALL ranks sequentially (the only serial work: a few dozen
//                   8-addend chains per pass), (d) writes the exact running sum at every tile start
//                   and the total.  The pass that sums weights / S1 also stores the normalised
//                   weights and accumulates the expected pose (:696-716) -- the former
//                   k_normalize_pose -- so normalisation costs no pass of its own.
//   k_exact_emit    prefix sums (the CDF) from the tile starts.
// Steady state: tile sums + 3 passes + emit = 5 launches (round 1: 9), and on a sharded filter every
// pass is ONE exchange of < 2 KB per rank instead of an all-gather of all weights.
// k_exact_single keeps a whole pass of a one-tile filter in one CTA (small filters, batches).
#pragma once

namespace mclb200 {

constexpr int kOpqCap = 160;            // opaque chunks per rank that travel inside the mailbox payload
constexpr int kItemWords = 10;          // step map (2) + 8 addends
constexpr int kHdrWords = 16;           // [0] K  [1..2] tail map  [3..6] pose partial sums  [7] approximate slice sum
static_assert((kHdrWords + kOpqCap * kItemWords) * 8 <= kMboxSlot, "exact-pass payload must fit a mailbox slot");
constexpr int kEvalBatch = 128;

struct ExactArgs {
    // addend_i = src[i]; divided by *norm when norm != nullptr and *norm > 0 (the `if (sum_weights > 0)`
    // of :680) -- that quotient is what `store` receives and what the pose sums use --; then divided
    // by *div when div != nullptr (discrete_distribution divides unconditionally)
    const double* src;       // [F][N]
    const double* norm;      // [F] or nullptr
    const double* div;       // [F] or nullptr
    double* store;           // [F][N] or nullptr
    // tile_sum holds sums of the RAW src values (or of the stored quotients when pre_norm == nullptr);
    // the approximate prefix is rescaled by *pre_norm (if > 0) and *div
    const double* pre_norm;  // [F] or nullptr
    int64_t N;               // elements of this rank's slice (per filter)
    int64_t glo;             // elements of lower ranks (global index of local element 0)
    int T;                   // tiles per filter
    int C;                   // chunks per filter = T * kTileChunks
    double* tile_sum;        // [F][T] approximate tile sums
    double* slice_sum;       // [world] approximate slice sums of all ranks (sharded; written by k_tile_sums' exchange)
    StepFn* chunk_fn;        // [F][C]
    StepFn* opq_pre;         // [F][C]  tile-compacted: step map from the previous anchor in the tile (or the tile start)
    int* opq_idx;            // [F][C]  tile-compacted: index of the opaque chunk inside its tile
    double* opq_add;         // [F][C][8] tile-compacted: the opaque chunk's addends (after the divisions)
    int* tile_opq;           // [F][T]  opaque chunks per tile
    int64_t* tile_elem;      // [F][T][3]  (a0, a1, reset)
    int* list_chunk;         // [F][C]  opaque chunks in order
    StepFn* list_fn;         // [2][F][C] (epoch parity) incoming step map of every opaque chunk, in order
    double* list_add;        // [2][F][C][8] their addends
    const StepFn* peer_list_fn[kMaxWorld];    // the same arrays of every rank (entries beyond kOpqCap are read from the owner)
    const double* peer_list_add[kMaxWorld];
    double* anchors;         // [F][C]  exact running sum after each opaque chunk, by order
    double* anchor_val;      // [F][C]  same, by chunk index
    double* tile_start;      // [F][T]  exact running sum before each tile
    double* total;           // [F]     exact sequential sum over ALL ranks
    double* rank_end;        // [world] exact running sum after each rank's slice (sharded) or nullptr
    double* out;             // [F][N]  prefix sums (emit) or nullptr
    int force_last_one;      // discrete_distribution sets _M_cp.back() = 1.0 (random.tcc:2677)
    double* coarse;          // coarse level of the CDF search: coarse[f][k] = out[(k+1)*8*coarse_m - 1]
    int coarse_m, coarse_n;
    double* mid;             // [F][C] middle level: mid[f][c] = the last prefix sum of chunk c (out[8c+7]); nullable
    unsigned int* done;      // [F] block-completion counters (zero on entry, reset on exit)
    // pose (the pass that normalises): expected_pose over the stored weights
    const double* px;
    const double* py;
    const double* pt;
    double* partial;         // [F][T][4]
    double* pose_out;        // [F][3]
    double* pose_host;       // mapped pinned twin (nullable)
    unsigned long long* update_no;   // bumped when the pose of the update is written (nullable)
    unsigned long long* dbg;         // diagnostics (nullable): clock64 stamps of the pass's phases, see mcl_debug_pass_cycles
    ShardDev sh;
};

// emit step: the thread that holds the last chunk of a coarse segment publishes its final prefix sum
__device__ __forceinline__ void emit_coarse(const ExactArgs& a, int f, int64_t chunk_in_filter, int64_t base, double last) {
    if (a.coarse_m > 0 && ((chunk_in_filter + 1) & (a.coarse_m - 1)) == 0 && base + kChunk <= a.N) {
        const int64_t k = (chunk_in_filter + 1) / a.coarse_m - 1;
        if (k < a.coarse_n) a.coarse[static_cast<int64_t>(f) * a.coarse_n + k] = last;
    }
}

struct RFn {
    StepFn f;
    int64_t reset;
};
struct RFnOp {
    __device__ __forceinline__ RFn operator()(const RFn& l, const RFn& r) const {
        if (r.reset) return r;
        return RFn{fn_compose(l.f, r.f), l.reset};
    }
};
struct SEOp {
    __device__ __forceinline__ ScanElem operator()(const ScanElem& l, const ScanElem& r) const { return se_combine(l, r); }
};
struct AddOp {
    __device__ __forceinline__ double operator()(double a, double b) const { return a + b; }
};

__device__ __forceinline__ void load_raw_chunk(double (&v)[kChunk], const double* __restrict__ src, int64_t base, int64_t N) {
    if (base + kChunk <= N) {
        const double2* p = reinterpret_cast<const double2*>(src + base);
#pragma unroll
        for (int i = 0; i < kChunk / 2; ++i) {
            const double2 t = __ldg(p + i);
            v[2 * i] = t.x;
            v[2 * i + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < kChunk; ++i) v[i] = (base + i < N) ? __ldg(src + base + i) : 0.0;
    }
}

__device__ __forceinline__ void load_chunk(double (&v)[kChunk], const double* __restrict__ src, int64_t base,
                                           int64_t N, bool use_div, double div) {
    load_raw_chunk(v, src, base, N);
    if (use_div) {
#pragma unroll
        for (int i = 0; i < kChunk; ++i) v[i] = __ddiv_rn(v[i], div);
    }
}

// integer inclusive block scan (NT threads); sm holds NT/32 ints
template <int NT>
__device__ __forceinline__ int block_scan_int(int v, int* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(kFullMask, v, d);
        if (lane >= d) v += o;
    }
    if (lane == 31) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int w = lane < NT / 32 ? sm[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(kFullMask, w, d);
            if (lane >= d) w += o;
        }
        if (lane < NT / 32) sm[lane] = w;
    }
    __syncthreads();
    if (warp > 0) v += sm[warp - 1];
    __syncthreads();
    return v;
}

// ---- approximate per-tile sums (no divisions) --------------------------------------------------
// Sharded filter: the last CTA also publishes the slice's sum and collects every rank's.
__device__ __forceinline__ void slice_sums_consume(const ExactArgs& a, unsigned long long epoch) {
    if (static_cast<int>(threadIdx.x) < a.sh.world)
        a.slice_sum[threadIdx.x] = ld_sys_f64(reinterpret_cast<const double*>(mbox_slot(a.sh, a.sh.rank, epoch, threadIdx.x) + 7));
    __syncthreads();
    if (threadIdx.x == 0) *a.sh.xseq = epoch;
}

__global__ void __launch_bounds__(kTileChunks) k_tile_sums(ExactArgs a) {
    pdl_enter();
    __shared__ double sm[kTileChunks / 32];
    __shared__ unsigned long long pay[kHdrWords];
    __shared__ bool is_last;
    const int f = blockIdx.y, t = blockIdx.x;
    const double* src = a.src + static_cast<int64_t>(f) * a.N;
    const int64_t base = (static_cast<int64_t>(t) * kTileChunks + threadIdx.x) * kChunk;
    double v[kChunk];
    load_raw_chunk(v, src, base, a.N);
    double c = 0.0;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) c += v[i];
    const double s = block_sum<kTileChunks>(c, sm);
    if (threadIdx.x == 0) a.tile_sum[static_cast<int64_t>(f) * a.T + t] = s;
    if (a.sh.world <= 1) return;
    // ---- sharded: slice sum -> every rank ----
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(a.done + f, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double tot = 0.0;
    for (int tt = threadIdx.x; tt < a.T; tt += kTileChunks) tot += a.tile_sum[tt];
    tot = block_sum<kTileChunks>(tot, sm);
    if (threadIdx.x < kHdrWords) pay[threadIdx.x] = 0ull;
    __syncthreads();
    if (threadIdx.x == 0) {
        pay[7] = static_cast<unsigned long long>(__double_as_longlong(tot));
        a.done[f] = 0;
    }
    __syncthreads();
    const unsigned long long epoch = *a.sh.xseq + 1ull;
    shard_publish(a.sh, epoch, pay, kHdrWords);
    if (!a.sh.fused) return;
    if (!shard_wait(a.sh, epoch)) return;
    slice_sums_consume(a, epoch);
}

// host-ordered ranks: the consume half of k_tile_sums' exchange
__global__ void __launch_bounds__(kTileChunks) k_slice_sums_collect(ExactArgs a) {
    pdl_enter();
    const unsigned long long epoch = *a.sh.xseq + 1ull;
    if (!shard_wait(a.sh, epoch)) return;
    slice_sums_consume(a, epoch);
}

// ---- the last CTA of an exact pass ------------------------------------------------------------
struct FinishShared {
    RFn smr[kTileChunks / 32];
    RFn inc[kTileChunks];
    int smi[kTileChunks / 32];
    int cinc[kTileChunks];
    unsigned long long pay[kHdrWords + kOpqCap * kItemWords];
    unsigned long long items[kEvalBatch][kItemWords];
    double item_v[kEvalBatch];              // running sum after each item of the batch
    int item_q[kEvalBatch], item_r[kEvalBatch];
    int pay_chunk[kOpqCap];                 // chunk index of this rank's first kOpqCap opaque chunks
    int start[kMaxWorld + 1], kq[kMaxWorld];
    double vin, vcur;
    double pose[4];
    double sd[kTileChunks / 32];
};

struct TileScan {
    int t0, t1;
    RFn exc;
    int cexc, K;
    RFn all;
};

// (tile records, opaque-chunk records and partial sums are written by OTHER CTAs -- in the fused kernel several
// times per launch -- so the last CTA reads them past its L1: __ldcg)
__device__ __forceinline__ RFn tile_rfn(const int64_t* tile_elem, int t) {
    return RFn{StepFn{static_cast<int64_t>(__ldcg(reinterpret_cast<const long long*>(tile_elem + 3 * t))),
                      static_cast<int64_t>(__ldcg(reinterpret_cast<const long long*>(tile_elem + 3 * t + 1)))},
               static_cast<int64_t>(__ldcg(reinterpret_cast<const long long*>(tile_elem + 3 * t + 2)))};
}

// every thread folds its range of tiles; block scans give the state entering the range
__device__ __forceinline__ TileScan scan_tiles(const ExactArgs& a, int f, FinishShared& S) {
    const int tid = threadIdx.x, T = a.T;
    const int64_t* tile_elem = a.tile_elem + static_cast<int64_t>(f) * T * 3;
    const int* tile_opq = a.tile_opq + static_cast<int64_t>(f) * T;
    TileScan r;
    const int tpt = (T + kTileChunks - 1) / kTileChunks;
    r.t0 = min(T, tid * tpt);
    r.t1 = min(T, r.t0 + tpt);
    const RFn ident{fn_identity(), 0};
    RFn loc = ident;
    int cnt = 0;
    for (int t = r.t0; t < r.t1; ++t) {
        loc = RFnOp()(loc, tile_rfn(tile_elem, t));
        cnt += __ldcg(tile_opq + t);
    }
    const RFn inc = block_scan_inclusive<kTileChunks>(loc, RFnOp(), S.smr, ident);
    S.inc[tid] = inc;
    const int ci = block_scan_int<kTileChunks>(cnt, S.smi);
    S.cinc[tid] = ci;
    __syncthreads();
    r.exc = tid ? S.inc[tid - 1] : ident;
    r.cexc = tid ? S.cinc[tid - 1] : 0;
    r.K = S.cinc[kTileChunks - 1];
    r.all = S.inc[kTileChunks - 1];
    __syncthreads();
    return r;
}

// (a) the slice's opaque chunks in order, each with the step map that leads to it from the previous
// anchor (or from the slice start), into the epoch-parity lists and -- the first kOpqCap -- into the
// payload; header = count, tail map, pose partial sums
__device__ __forceinline__ void finish_local(const ExactArgs& a, int f, FinishShared& S, const TileScan& ts, unsigned long long epoch,
                                             bool pose) {
    const int par = static_cast<int>(epoch & 1ull);
    const int64_t* tile_elem = a.tile_elem + static_cast<int64_t>(f) * a.T * 3;
    const int* tile_opq = a.tile_opq + static_cast<int64_t>(f) * a.T;
    const int64_t fc = static_cast<int64_t>(f) * a.C;
    const int64_t lc = (static_cast<int64_t>(par) * gridDim.y + f) * a.C;
    // one thread per opaque chunk (they are few, and mostly in the filter's first tiles: a loop of the tiles'
    // owners would serialise on one thread): entry e belongs to the thread range whose inclusive count exceeds e
    const int tpt = (a.T + kTileChunks - 1) / kTileChunks;
    for (int e = threadIdx.x; e < ts.K; e += kTileChunks) {
        int lo = 0, hi = kTileChunks - 1;
        while (lo < hi) {   // first thread range th with cinc[th] > e
            const int mid = (lo + hi) >> 1;
            if (S.cinc[mid] > e)
                hi = mid;
            else
                lo = mid + 1;
        }
        const int th = lo;
        int local = e - (th ? S.cinc[th - 1] : 0);
        RFn run = th ? S.inc[th - 1] : RFn{fn_identity(), 0};   // state entering the range's first tile
        int t = th * tpt;
        for (;;) {
            const int n_t = __ldcg(tile_opq + t);
            if (local < n_t) break;
            local -= n_t;
            run = RFnOp()(run, tile_rfn(tile_elem, t));
            ++t;
        }
        const int64_t slot = fc + static_cast<int64_t>(t) * kTileChunks + local;
        const StepFn pre{static_cast<int64_t>(__ldcg(reinterpret_cast<const long long*>(&a.opq_pre[slot].a0))),
                         static_cast<int64_t>(__ldcg(reinterpret_cast<const long long*>(&a.opq_pre[slot].a1)))};
        const StepFn fn = local == 0 ? fn_compose(run.f, pre) : pre;   // the tile's first opaque chunk continues the run that entered the tile
        const int chunk = t * kTileChunks + __ldcg(a.opq_idx + slot);
        a.list_chunk[fc + e] = chunk;
        a.list_fn[lc + e] = fn;
        unsigned long long* pw = e < kOpqCap ? S.pay + kHdrWords + e * kItemWords : nullptr;
        if (pw) {
            S.pay_chunk[e] = chunk;
            pw[0] = static_cast<unsigned long long>(fn.a0);
            pw[1] = static_cast<unsigned long long>(fn.a1);
        }
        double add[kChunk];
#pragma unroll
        for (int q = 0; q < kChunk; ++q) add[q] = __ldcg(a.opq_add + slot * kChunk + q);
#pragma unroll
        for (int q = 0; q < kChunk; ++q) {
            a.list_add[(lc + e) * kChunk + q] = add[q];
            if (pw) pw[2 + q] = static_cast<unsigned long long>(__double_as_longlong(add[q]));
        }
    }
    if (threadIdx.x < kHdrWords) S.pay[threadIdx.x] = 0ull;
    __syncthreads();
    if (threadIdx.x == 0) {
        S.pay[0] = static_cast<unsigned long long>(ts.K);
        S.pay[1] = static_cast<unsigned long long>(ts.all.f.a0);
        S.pay[2] = static_cast<unsigned long long>(ts.all.f.a1);
        if (pose)
            for (int k = 0; k < 4; ++k) S.pay[3 + k] = static_cast<unsigned long long>(__double_as_longlong(S.pose[k]));
    }
    __syncthreads();
}

// (c) + (d): sequential evaluation of every rank's opaque chunks and tails, tile starts, total, pose
// own_in_smem: S.pay / S.pay_chunk hold this rank's payload (the block that ran finish_local)
__device__ __forceinline__ void finish_global(const ExactArgs& a, int f, FinishShared& S, const TileScan& ts, unsigned long long epoch,
                                              bool pose, bool own_in_smem) {
    const int tid = threadIdx.x;
    const int world = a.sh.world > 1 ? a.sh.world : 1, me = a.sh.world > 1 ? a.sh.rank : 0;
    const int par = static_cast<int>(epoch & 1ull);
    const int64_t fc = static_cast<int64_t>(f) * a.C;
    const int64_t lc = (static_cast<int64_t>(par) * gridDim.y + f) * a.C;
    // per-rank counts: own from the scan, the others from their headers
    if (tid < world) S.kq[tid] = tid == me ? ts.K : static_cast<int>(ld_sys_u64(mbox_slot(a.sh, me, epoch, tid) + 0));
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int q = 0; q < world; ++q) {
            S.start[q] = s;
            s += S.kq[q] + 1;   // the rank's opaque chunks, then its tail map
        }
        S.start[world] = s;
        S.vcur = 0.0;
        S.vin = 0.0;
    }
    __syncthreads();
    const int G = S.start[world];
    for (int b0 = 0; b0 < G; b0 += kEvalBatch) {
        const int nb = min(kEvalBatch, G - b0);
        if (tid < nb) {
            const int g = b0 + tid;
            int q = 0;
            while (q + 1 < world && S.start[q + 1] <= g) ++q;
            S.item_q[tid] = q;
            S.item_r[tid] = g - S.start[q];
        }
        __syncthreads();
        for (int w = tid; w < nb * kItemWords; w += kTileChunks) {
            const int it = w / kItemWords, k = w - it * kItemWords;
            const int q = S.item_q[it], r = S.item_r[it];
            unsigned long long x;
            if (r == S.kq[q]) {   // the rank's tail: a step map and no addends
                if (k >= 2)
                    x = 0ull;     // +0.0
                else if (q == me)
                    x = static_cast<unsigned long long>(k == 0 ? ts.all.f.a0 : ts.all.f.a1);
                else
                    x = ld_sys_u64(mbox_slot(a.sh, me, epoch, q) + 1 + k);
            } else if (q == me && own_in_smem && r < kOpqCap) {
                x = S.pay[kHdrWords + r * kItemWords + k];
            } else if (q == me) {
                x = k == 0   ? static_cast<unsigned long long>(a.list_fn[lc + r].a0)
                    : k == 1 ? static_cast<unsigned long long>(a.list_fn[lc + r].a1)
                             : static_cast<unsigned long long>(__double_as_longlong(a.list_add[(lc + r) * kChunk + (k - 2)]));
            } else if (r < kOpqCap) {
                x = ld_sys_u64(mbox_slot(a.sh, me, epoch, q) + kHdrWords + r * kItemWords + k);
            } else {   // beyond the payload: the owner's lists (same epoch parity), system-scope loads over NVLink
                const int64_t pl = static_cast<int64_t>(par) * a.C + r;
                x = k < 2 ? ld_sys_u64(reinterpret_cast<const unsigned long long*>(a.peer_list_fn[q] + pl) + k)
                          : ld_sys_u64(a.peer_list_add[q] + pl * kChunk + (k - 2));
            }
            S.items[it][k] = x;
        }
        __syncthreads();
        if (tid == 0) {
            double V = S.vcur;
            for (int it = 0; it < nb; ++it) {
                const StepFn fn{static_cast<int64_t>(S.items[it][0]), static_cast<int64_t>(S.items[it][1])};
                double s = fn_apply(fn, V);
#pragma unroll
                for (int e = 0; e < kChunk; ++e) s = __dadd_rn(s, __longlong_as_double(static_cast<long long>(S.items[it][2 + e])));
                V = s;
                S.item_v[it] = V;   // (no global access inside this dependent chain)
            }
            S.vcur = V;
        }
        __syncthreads();
        if (tid < nb) {   // what each item's value is: a rank's end, this rank's incoming sum, or an anchor of this rank
            const int q = S.item_q[tid], r = S.item_r[tid];
            const double V = S.item_v[tid];
            if (r == S.kq[q]) {
                if (a.rank_end) a.rank_end[q] = V;
                if (q + 1 == me) S.vin = V;
            } else if (q == me) {
                a.anchors[fc + r] = V;
                const int chunk = (own_in_smem && r < kOpqCap) ? S.pay_chunk[r] : a.list_chunk[fc + r];
                a.anchor_val[fc + chunk] = V;
            }
        }
        __syncthreads();
    }
    // exact running sum at every tile start of this rank's slice, and the total over all ranks
    {
        const int64_t* tile_elem = a.tile_elem + static_cast<int64_t>(f) * a.T * 3;
        const int* tile_opq = a.tile_opq + static_cast<int64_t>(f) * a.T;
        const double vin = S.vin;
        RFn run = ts.exc;
        int rank = ts.cexc;
        for (int t = ts.t0; t < ts.t1; ++t) {
            a.tile_start[static_cast<int64_t>(f) * a.T + t] = fn_apply(run.f, run.reset ? a.anchors[fc + rank - 1] : vin);
            rank += __ldcg(tile_opq + t);
            run = RFnOp()(run, tile_rfn(tile_elem, t));
        }
        if (tid == 0) a.total[f] = S.vcur;
    }
    if (pose && tid == 0) {
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        for (int q = 0; q < world; ++q)
            for (int k = 0; k < 4; ++k)
                v[k] += q == me ? S.pose[k] : ld_sys_f64(reinterpret_cast<const double*>(mbox_slot(a.sh, me, epoch, q) + 3 + k));
        const double th = atan2(v[2], v[3]);
        a.pose_out[3 * f + 0] = v[0];
        a.pose_out[3 * f + 1] = v[1];
        a.pose_out[3 * f + 2] = th;
        if (a.pose_host) {
            a.pose_host[3 * f + 0] = v[0];
            a.pose_host[3 * f + 1] = v[1];
            a.pose_host[3 * f + 2] = th;
        }
        if (a.update_no && f == 0) *a.update_no += 1ull;   // this update is complete: the next one draws fresh noise
    }
    if (a.sh.world > 1 && tid == 0) *a.sh.xseq = epoch;
}

// the per-block pose partial sums of this rank, folded in block order (deterministic)
__device__ __forceinline__ void fold_pose_partials(const ExactArgs& a, int f, FinishShared& S) {
    double v[4] = {0, 0, 0, 0};
    for (int b = threadIdx.x; b < a.T; b += kTileChunks) {
        const double* p = a.partial + (static_cast<int64_t>(f) * a.T + b) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] += __ldcg(p + q);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = block_sum<kTileChunks>(v[q], S.sd);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) S.pose[q] = v[q];
    }
    __syncthreads();
}

// One tile of an exact pass: the 8-addend chunk step maps of tile t, their block scan, the tile's summary
// (tile_elem) and its opaque chunks compacted; optionally the normalised weights stored and the tile's
// expected-pose partial sums.  All threads of the CTA call; ends with the CTA's global writes issued.
template <bool POSE>
__device__ __forceinline__ void tile_phase(const ExactArgs& a, FinishShared& S, int* wcnt, int f, int t) {
    const int tid = threadIdx.x;
    const double* src = a.src + static_cast<int64_t>(f) * a.N;
    // (sums and tile sums may have been written by another CTA earlier in the SAME launch -- the fused kernel --
    // so they are read past this SM's L1)
    const double nrm = a.norm ? __ldcg(a.norm + f) : 0.0;
    const bool use_norm = a.norm != nullptr && nrm > 0.0;
    const bool use_div = a.div != nullptr;
    const double div = use_div ? __ldcg(a.div + f) : 1.0;

    // approximate running sum before this tile: the lower ranks' slices + this rank's earlier tiles
    double pre = 0.0;
    for (int tt = tid; tt < t; tt += kTileChunks) pre += __ldcg(a.tile_sum + static_cast<int64_t>(f) * a.T + tt);
    if (a.sh.world > 1 && tid < a.sh.rank) pre += __ldcg(a.slice_sum + tid);
    pre = block_sum<kTileChunks>(pre, S.sd);
    if (a.pre_norm) {
        const double pn = __ldcg(a.pre_norm + f);
        if (pn > 0.0) pre = pre / pn;   // `if (sum_weights > 0)`: otherwise the weights were left as they were
    }
    if (use_div) pre = pre / div;

    const int64_t base = (static_cast<int64_t>(t) * kTileChunks + tid) * kChunk;
    double v[kChunk];
    load_raw_chunk(v, src, base, a.N);   // (read-only for the whole launch: the non-coherent path is fine)
    if (use_norm) {
#pragma unroll
        for (int i = 0; i < kChunk; ++i) v[i] = __ddiv_rn(v[i], nrm);
    }
    if (a.store) {
        double* st = a.store + static_cast<int64_t>(f) * a.N;
        if (base + kChunk <= a.N) {
            double2* p = reinterpret_cast<double2*>(st + base);
#pragma unroll
            for (int i = 0; i < kChunk / 2; ++i) p[i] = make_double2(v[2 * i], v[2 * i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < kChunk; ++i)
                if (base + i < a.N) st[base + i] = v[i];
        }
    }
    if (POSE) {
        // expected_pose (:696-716) over the weights just normalised
        const int64_t fo = static_cast<int64_t>(f) * a.N;
        double ax = 0.0, ay = 0.0, as = 0.0, ac = 0.0;
        double q[kChunk];   // slots beyond N load 0 and carry weight 0
        load_raw_chunk(q, a.pt + fo, base, a.N);
#pragma unroll
        for (int i = 0; i < kChunk; ++i) {
            double s, c;
            sincos(q[i], &s, &c);
            as += v[i] * s;
            ac += v[i] * c;
        }
        load_raw_chunk(q, a.px + fo, base, a.N);
#pragma unroll
        for (int i = 0; i < kChunk; ++i) ax += v[i] * q[i];
        load_raw_chunk(q, a.py + fo, base, a.N);
#pragma unroll
        for (int i = 0; i < kChunk; ++i) ay += v[i] * q[i];
        ax = block_sum<kTileChunks>(ax, S.sd);
        ay = block_sum<kTileChunks>(ay, S.sd);
        as = block_sum<kTileChunks>(as, S.sd);
        ac = block_sum<kTileChunks>(ac, S.sd);
        if (tid == 0) {
            double* p = a.partial + (static_cast<int64_t>(f) * a.T + t) * 4;
            p[0] = ax;
            p[1] = ay;
            p[2] = as;
            p[3] = ac;
        }
    }
    if (use_div) {
#pragma unroll
        for (int i = 0; i < kChunk; ++i) v[i] = __ddiv_rn(v[i], div);
    }
    double c = 0.0;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) c += v[i];
    const double incl = block_scan_inclusive<kTileChunks>(c, AddOp(), S.sd, 0.0);
    const double s_in = pre + (incl - c);
    const double s_out = s_in + c;

    StepFn fn = fn_identity();
    int opaque = 0;
    if (base < a.N) {
        const int64_t cnt = a.glo + ((base + kChunk < a.N) ? base + kChunk : a.N);
        const int e = chunk_safe_binade(s_in, s_out, cnt);
        if (e < 0) {
            fn = fn_opaque();
            opaque = 1;
        } else {
            fn = chunk_step_fn(v, kChunk, e);
        }
    }
    const int64_t cidx = static_cast<int64_t>(f) * a.C + static_cast<int64_t>(t) * kTileChunks + tid;
    a.chunk_fn[cidx] = fn;

    const RFn ident{fn_identity(), 0};
    const RFn el = opaque ? RFn{fn_identity(), 1} : RFn{fn, 0};
    const RFn inc = block_scan_inclusive<kTileChunks>(el, RFnOp(), S.smr, ident);
    S.inc[tid] = inc;
    // compact the tile's opaque chunks (in order): the last CTA never scans chunk records
    const unsigned bal = __ballot_sync(kFullMask, opaque);
    if ((tid & 31) == 0) wcnt[tid >> 5] = __popc(bal);
    __syncthreads();
    const RFn exc = tid ? S.inc[tid - 1] : ident;
    int orank = __popc(bal & ((1u << (tid & 31)) - 1u));
    int nopq = 0;
#pragma unroll
    for (int w = 0; w < kTileChunks / 32; ++w) {
        if (w < (tid >> 5)) orank += wcnt[w];
        nopq += wcnt[w];
    }
    if (opaque) {
        const int64_t slot = static_cast<int64_t>(f) * a.C + static_cast<int64_t>(t) * kTileChunks + orank;
        a.opq_pre[slot] = exc.f;
        a.opq_idx[slot] = tid;
#pragma unroll
        for (int i = 0; i < kChunk; ++i) a.opq_add[slot * kChunk + i] = v[i];
    }
    if (tid == kTileChunks - 1) {
        int64_t* te = a.tile_elem + (static_cast<int64_t>(f) * a.T + t) * 3;
        te[0] = inc.f.a0;
        te[1] = inc.f.a1;
        te[2] = inc.reset;
        a.tile_opq[static_cast<int64_t>(f) * a.T + t] = nopq;
    }
}

// The last CTA of a pass: order the opaque chunks, exchange with the other ranks, evaluate, tile starts.
// Returns false when the pass continues in another kernel (host-ordered exchange) or a peer went missing.
template <bool POSE>
__device__ __forceinline__ bool pass_finish(const ExactArgs& a, FinishShared& S, int f, long long c_begin) {
    const int tid = threadIdx.x;
    const long long c0 = clock64();
    if (POSE) fold_pose_partials(a, f, S);
    const bool sharded = a.sh.world > 1;
    const unsigned long long epoch = sharded ? *a.sh.xseq + 1ull : 0ull;
    const long long c1 = clock64();
    const TileScan ts = scan_tiles(a, f, S);
    const long long c2 = clock64();
    finish_local(a, f, S, ts, epoch, POSE);
    const long long c3 = clock64();
    if (sharded) {
        const int nw = kHdrWords + min(static_cast<int>(S.pay[0]), kOpqCap) * kItemWords;
        shard_publish(a.sh, epoch, S.pay, nw);
        if (!a.sh.fused) return false;
        if (!shard_wait(a.sh, epoch)) return false;
    }
    const long long c4 = clock64();
    finish_global(a, f, S, ts, epoch, POSE, true);
    if (a.dbg && tid == 0) {
        const long long c5 = clock64();
        a.dbg[1] = static_cast<unsigned long long>(c0 - c_begin);   // last CTA: its own tile phase + waiting to be last
        a.dbg[2] = static_cast<unsigned long long>(c1 - c0);        // pose fold
        a.dbg[3] = static_cast<unsigned long long>(c2 - c1);        // tile scan
        a.dbg[4] = static_cast<unsigned long long>(c3 - c2);        // ordered opaque list
        a.dbg[5] = static_cast<unsigned long long>(c4 - c3);        // exchange
        a.dbg[6] = static_cast<unsigned long long>(c5 - c4);        // serial evaluation + tile starts
        a.dbg[7] = S.pay[0];                                        // opaque chunks of this rank
    }
    return true;
}

template <bool POSE>
__global__ void __launch_bounds__(kTileChunks) k_exact_pass(ExactArgs a) {
    pdl_enter();
    __shared__ FinishShared S;
    __shared__ int wcnt[kTileChunks / 32];
    __shared__ bool is_last;
    const int f = blockIdx.y, t = blockIdx.x, tid = threadIdx.x;
    const long long c_begin = clock64();
    tile_phase<POSE>(a, S, wcnt, f, t);
    // ---- the last CTA of the filter finishes the pass ----
    __threadfence();
    __syncthreads();
    if (a.dbg && tid == 0) atomicMax(a.dbg + 0, static_cast<unsigned long long>(clock64() - c_begin));   // slowest CTA's tile phase
    if (tid == 0) is_last = atomicAdd(a.done + f, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (tid == 0) a.done[f] = 0;
    pass_finish<POSE>(a, S, f, c_begin);
}

// host-ordered ranks: the consume half of an exact pass (one CTA)
template <bool POSE>
__global__ void __launch_bounds__(kTileChunks) k_exact_finish(ExactArgs a) {
    pdl_enter();
    __shared__ FinishShared S;
    const unsigned long long epoch = *a.sh.xseq + 1ull;
    if (!shard_wait(a.sh, epoch)) return;
    if (POSE) {   // this rank's own pose sums travelled in its own mailbox slot too
        if (threadIdx.x < 4) S.pose[threadIdx.x] = ld_sys_f64(reinterpret_cast<const double*>(mbox_slot(a.sh, a.sh.rank, epoch, a.sh.rank) + 3 + threadIdx.x));
        __syncthreads();
    }
    const TileScan ts = scan_tiles(a, 0, S);
    finish_global(a, 0, S, ts, epoch, POSE, false);
}

// prefix sums of one tile from its exact start value: a scan of the chunks' step maps and anchors gives every
// chunk's exact input, then 8 sequential adds per thread
__device__ __forceinline__ void emit_tile(const ExactArgs& a, ScanElem* sms, ScanElem* sm_inc, int f, int t) {
    const int tid = threadIdx.x;
    const double* src = a.src + static_cast<int64_t>(f) * a.N;
    const double nrm = a.norm ? __ldcg(a.norm + f) : 0.0;
    const bool use_norm = a.norm != nullptr && nrm > 0.0;
    const bool use_div = a.div != nullptr;
    const double div = use_div ? __ldcg(a.div + f) : 1.0;
    const int64_t cidx = static_cast<int64_t>(f) * a.C + static_cast<int64_t>(t) * kTileChunks + tid;
    const double tstart = __ldcg(a.tile_start + static_cast<int64_t>(f) * a.T + t);

    const StepFn cf = a.chunk_fn[cidx];
    ScanElem el = fn_is_opaque(cf) ? se_abs(__ldcg(a.anchor_val + cidx)) : se_fn(cf);
    if (tid == 0) el = se_combine(se_abs(tstart), el);
    const ScanElem ident = se_fn(fn_identity());
    const ScanElem inc = block_scan_inclusive<kTileChunks>(el, SEOp(), sms, ident);
    sm_inc[tid] = inc;
    __syncthreads();
    double s = tid ? bits_dbl(sm_inc[tid - 1].a0) : tstart;

    const int64_t base = (static_cast<int64_t>(t) * kTileChunks + tid) * kChunk;
    if (base >= a.N) return;
    double v[kChunk];
    load_raw_chunk(v, src, base, a.N);
    if (use_norm) {
#pragma unroll
        for (int i = 0; i < kChunk; ++i) v[i] = __ddiv_rn(v[i], nrm);
    }
    if (use_div) {
#pragma unroll
        for (int i = 0; i < kChunk; ++i) v[i] = __ddiv_rn(v[i], div);
    }
    double* out = a.out + static_cast<int64_t>(f) * a.N;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
        s = __dadd_rn(s, v[i]);
        v[i] = s;
    }
    if (a.force_last_one && a.N - 1 >= base && a.N - 1 < base + kChunk) v[a.N - 1 - base] = 1.0;
    emit_coarse(a, f, static_cast<int64_t>(t) * kTileChunks + tid, base, v[kChunk - 1]);
    if (a.mid) {   // the chunk's last VALID prefix sum (the filter's last chunk may be partial)
        double last = v[kChunk - 1];
        if (base + kChunk > a.N) {
#pragma unroll
            for (int i = 0; i < kChunk; ++i)
                if (base + i == a.N - 1) last = v[i];
        }
        a.mid[static_cast<int64_t>(f) * a.C + static_cast<int64_t>(t) * kTileChunks + tid] = last;
    }
    if (base + kChunk <= a.N) {
        double2* p = reinterpret_cast<double2*>(out + base);
#pragma unroll
        for (int i = 0; i < kChunk / 2; ++i) p[i] = make_double2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < kChunk; ++i)
            if (base + i < a.N) out[base + i] = v[i];
    }
}

__global__ void __launch_bounds__(kTileChunks) k_exact_emit(ExactArgs a) {
    pdl_enter();
    __shared__ ScanElem sms[kTileChunks / 32];
    __shared__ ScanElem sm_inc[kTileChunks];
    emit_tile(a, sms, sm_inc, blockIdx.y, blockIdx.x);
}

// T == 1 (a filter of at most kTile = 4096 particles is ONE tile): chunks, walk and emit of the
// exact sequential sum in a single CTA per filter -- one launch per pass, which is what a small
// filter's update time is made of.  Same arithmetic as the kernels above with the tile prefix fixed
// at 0: chunk step maps, the ordered serial pass over the opaque chunks (their addends re-read from
// global memory by thread 0: they are rare), the total, and the prefix sums.
__global__ void __launch_bounds__(kTileChunks) k_exact_single(ExactArgs a) {
    pdl_enter();
    __shared__ double smd[kTileChunks / 32];
    __shared__ RFn smr[kTileChunks / 32];
    __shared__ RFn sm_inc[kTileChunks];
    __shared__ ScanElem sms[kTileChunks / 32];
    __shared__ ScanElem sm_se[kTileChunks];
    __shared__ double sm_anchor[kTileChunks];   // exact running sum after an opaque chunk, by chunk index
    __shared__ StepFn sm_pre[kTileChunks];      // step map from the previous anchor (or 0) to each opaque chunk, by rank
    __shared__ int sm_opq[kTileChunks];         // opaque chunk indices in order
    __shared__ int wcnt[kTileChunks / 32];
    const int f = blockIdx.y, tid = threadIdx.x;
    const double* src = a.src + static_cast<int64_t>(f) * a.N;
    const bool use_div = a.div != nullptr;
    const double div = use_div ? a.div[f] : 1.0;

    const int64_t base = static_cast<int64_t>(tid) * kChunk;
    double v[kChunk];
    load_chunk(v, src, base, a.N, use_div, div);
    double c = 0.0;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) c += v[i];
    const double incl = block_scan_inclusive<kTileChunks>(c, AddOp(), smd, 0.0);
    const double s_in = 0.0 + (incl - c);
    const double s_out = s_in + c;
    StepFn fn = fn_identity();
    int opaque = 0;
    if (base < a.N) {
        const int64_t cnt = (base + kChunk < a.N) ? base + kChunk : a.N;
        const int e = chunk_safe_binade(s_in, s_out, cnt);
        if (e < 0) {
            fn = fn_opaque();
            opaque = 1;
        } else {
            fn = chunk_step_fn(v, kChunk, e);
        }
    }
    const RFn ident{fn_identity(), 0};
    const RFn el = opaque ? RFn{fn_identity(), 1} : RFn{fn, 0};
    const RFn inc = block_scan_inclusive<kTileChunks>(el, RFnOp(), smr, ident);
    sm_inc[tid] = inc;
    const unsigned bal = __ballot_sync(kFullMask, opaque);
    if ((tid & 31) == 0) wcnt[tid >> 5] = __popc(bal);
    __syncthreads();
    const RFn exc = tid ? sm_inc[tid - 1] : ident;
    int orank = __popc(bal & ((1u << (tid & 31)) - 1u));
    int nopq = 0;
#pragma unroll
    for (int w = 0; w < kTileChunks / 32; ++w) {
        if (w < (tid >> 5)) orank += wcnt[w];
        nopq += wcnt[w];
    }
    if (opaque) {
        sm_opq[orank] = tid;
        sm_pre[orank] = exc.f;
    }
    __syncthreads();
    // serial pass over the opaque chunks in order, each from its exact input
    if (tid == 0) {
        double V = 0.0;
        for (int r = 0; r < nopq; ++r) {
            const int ch = sm_opq[r];
            double w8[kChunk];
            load_chunk(w8, src, static_cast<int64_t>(ch) * kChunk, a.N, use_div, div);
            V = chunk_seq_eval(w8, kChunk, fn_apply(sm_pre[r], V));
            sm_anchor[ch] = V;
        }
        const RFn run = sm_inc[kTileChunks - 1];
        a.total[f] = fn_apply(run.f, run.reset ? V : 0.0);
    }
    if (!a.out) return;
    __syncthreads();
    // prefix sums: every chunk's exact input from a scan of step maps and anchors, then 8 adds
    ScanElem se = opaque ? se_abs(sm_anchor[tid]) : se_fn(fn);
    if (tid == 0) se = se_combine(se_abs(0.0), se);
    const ScanElem seid = se_fn(fn_identity());
    const ScanElem sinc = block_scan_inclusive<kTileChunks>(se, SEOp(), sms, seid);
    sm_se[tid] = sinc;
    __syncthreads();
    double s = tid ? bits_dbl(sm_se[tid - 1].a0) : 0.0;
    if (base >= a.N) return;
    double* out = a.out + static_cast<int64_t>(f) * a.N;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
        s = __dadd_rn(s, v[i]);
        v[i] = s;
    }
    if (a.force_last_one && a.N - 1 >= base && a.N - 1 < base + kChunk) v[a.N - 1 - base] = 1.0;
    emit_coarse(a, f, tid, base, v[kChunk - 1]);
    if (a.mid) {
        double last = v[kChunk - 1];
        if (base + kChunk > a.N) {
#pragma unroll
            for (int i = 0; i < kChunk; ++i)
                if (base + i == a.N - 1) last = v[i];
        }
        a.mid[static_cast<int64_t>(f) * a.C + tid] = last;
    }
    if (base + kChunk <= a.N) {
        double2* p = reinterpret_cast<double2*>(out + base);
#pragma unroll
        for (int i = 0; i < kChunk / 2; ++i) p[i] = make_double2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < kChunk; ++i)
            if (base + i < a.N) out[base + i] = v[i];
    }
}

}  // namespace mclb200
