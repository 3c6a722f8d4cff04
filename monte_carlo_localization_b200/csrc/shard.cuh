// csrc/shard.cuh -- one-sided exchange between the ranks of a particle-sharded filter.
//
// The reference is one process; the coupling points of its update are the weight sum
// (src/particle_filter.cpp:679), the CDF of std::discrete_distribution (:658) and the pose sums
// (:702-710).  A sharded filter keeps every per-particle array LOCAL to the rank that owns the
// slot range and exchanges only small summaries at those points.  The exchange is done by the
// kernels themselves over NVLink peer mappings:
//
//   publish   every thread of ONE block (the last block of the producing kernel) stores its part of
//             the payload into slot [epoch & 1][sender] of EVERY rank's mailbox (plain stores on
//             peer-mapped addresses), fences at system scope, and one thread per destination
//             release-stores the epoch into that rank's flag word.
//   wait      one thread per source acquire-loads its own flag word until it carries the epoch.
//
// No rank ever READS peer memory on this path (round 1 found peer loads served stale lines from
// the reader's L1; scripts/peer_probe.cu), except for the overflow lists below, which use
// system-scope loads.  Two mailbox slots per sender suffice: a rank can publish epoch e + 2 only
// after it has seen every peer's epoch e + 1, which the peer publishes after it has read epoch e.
//
// `fused == 0` (emulated ranks on one GPU, or the NCCL transport): the producing kernel only
// publishes; the host orders the ranks (or all-gathers the mailboxes) and a separate one-block
// kernel consumes.  Kernels of different ranks must never spin on each other on ONE device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mclb200 {

constexpr int kMaxWorld = 16;
constexpr int kMboxSlot = 16384;                  // bytes per (parity, sender) mailbox slot
constexpr int kMboxWords = kMboxSlot / 8;
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;   // a dead peer must not hang the GPU

enum ShardErr { kShardOk = 0, kShardTimeout = 1, kShardOverflow = 2, kShardMissing = 3 };

struct ShardDev {
    int world, rank;
    int fused;                                   // 1: the publishing block also waits (one rank per GPU)
    uint8_t* mbox[kMaxWorld];                    // every rank's mailbox [2][world][kMboxSlot] (own included)
    unsigned long long* flag[kMaxWorld];         // every rank's flag words [kMaxWorld] (indexed by sender)
    unsigned long long* xseq;                    // own: number of exchanges completed
    int* err;                                    // own: first error (mapped pinned memory)
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const void* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_sys_u32(const void* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_sys_f64(const double* p) { return __longlong_as_double(static_cast<long long>(ld_sys_u64(p))); }
__device__ __forceinline__ void st_sys_u64(void* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned long long* mbox_slot(const ShardDev& s, int owner, unsigned long long epoch, int sender) {
    return reinterpret_cast<unsigned long long*>(s.mbox[owner] + (static_cast<size_t>(epoch & 1ull) * s.world + sender) * kMboxSlot);
}

// Called by ALL threads of one block.  payload: nwords 8-byte words in shared memory.
__device__ __forceinline__ void shard_publish(const ShardDev& s, unsigned long long epoch, const unsigned long long* payload, int nwords) {
    for (int q = 0; q < s.world; ++q) {
        unsigned long long* dst = mbox_slot(s, q, epoch, s.rank);
        for (int w = threadIdx.x; w < nwords; w += blockDim.x) st_sys_u64(dst + w, payload[w]);
    }
    // the barrier orders the block's payload stores before the releasing threads; their fence + release store are
    // cumulative over them at system scope
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < s.world) {
        __threadfence_system();
        st_release_sys(s.flag[threadIdx.x] + s.rank, epoch);
    }
}

// Called by ALL threads of one block.  Returns false (and records the error) if a peer's payload
// did not arrive: after kSpinTimeoutNs when fused, immediately when the host orders the ranks.
__device__ __forceinline__ bool shard_wait(const ShardDev& s, unsigned long long epoch) {
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < s.world) {
        const unsigned long long* f = s.flag[s.rank] + threadIdx.x;
        // relaxed polls, then ONE acquire: an acquire load invalidates the SM's L1 every time it is issued
        if (ld_sys_u64(f) < epoch) {
            if (!s.fused) {
                s_bad = kShardMissing;
            } else {
                const unsigned long long t0 = global_timer_ns();
                while (ld_sys_u64(f) < epoch) {
                    if (global_timer_ns() - t0 > kSpinTimeoutNs) {
                        s_bad = kShardTimeout;
                        break;
                    }
                    __nanosleep(64);
                }
            }
        }
        (void)ld_acquire_sys(f);
    }
    __syncthreads();
    const int bad = s_bad;
    if (bad && threadIdx.x == 0) atomicCAS(s.err, 0, bad);
    __syncthreads();
    return bad == 0;
}

}  // namespace mclb200
