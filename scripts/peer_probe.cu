// scripts/peer_probe.cu -- which loads return stale data when one GPU reads a buffer that its
// NVLink peer rewrites every other iteration?  (Round-1 finding: the resampling kernel's peer
// reads of source poses were stale with plain loads, and the packed 32-byte variant was not exact
// even with cache-volatile loads.)  This probe reproduces the access pattern without the filter:
//
//   two GPUs, each owns two buffers of n 32-byte records (double-buffered state).  Iteration t:
//   every GPU runs ONE persistent kernel (148 x 1024 threads) that reads records at random indices
//   of the PEER's buffer [t & 1] with the load flavour under test, checks the tag the record must
//   carry (the iteration that wrote it), and writes its own buffer [(t + 1) & 1] with tag t + 1.
//   Between iterations both devices are synchronised by the host (cudaDeviceSynchronize on both),
//   which is STRONGER than the stream-ordered all-gather of the product: a wrong tag seen here is a
//   cache effect, not a race.
//
// Transport: "peer" = one process, cudaDeviceEnablePeerAccess; "ipc" = two processes, the buffers
// mapped with cudaIpcOpenMemHandle (what the product did), a pipe barrier between iterations.
// Prints one JSON line per (transport, records, flavour).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/peer_probe scripts/peer_probe.cu
//   gpurun --gpus 2 -- 'scripts/peer_probe > gpurun_out/peer_probe.jsonl'
#include <cuda_runtime.h>
#include <stdint.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(3);                                                                           \
        }                                                                                      \
    } while (0)

struct __align__(32) Rec {
    double tag, idx, owner, pad;
};

enum Flavour { kPlainSoA = 0, kLdgSoA, kCvSoA, kSysSoA, kPlain16, kLdg16, kCv16, kSys16, kVol16, kSys32, kNumFlavours };
static const char* kNames[kNumFlavours] = {"ld.global 3x8B",       "ld.global.nc 3x8B",   "ld.global.cv 3x8B", "ld.relaxed.sys 3x8B",
                                           "ld.global 2x16B",      "ld.global.nc 2x16B",  "ld.global.cv 2x16B", "ld.relaxed.sys 2x16B",
                                           "ld.volatile 2x16B",    "ld.relaxed.sys 1x32B"};

__device__ __forceinline__ double ld_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_sys2(const double2* p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_vol2(const double2* p) {
    double2 v;
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ld_sys4(const Rec* p, double* a, double* b, double* c, double* d) {
    asm volatile("ld.relaxed.sys.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(*a), "=d"(*b), "=d"(*c), "=d"(*d) : "l"(p) : "memory");
}

// counts[0] = reads, [1] = wrong tag that equals the tag of two iterations ago (stale line),
// [2] = any other wrong tag / index, [3] = torn record (fields of different iterations)
__global__ void __launch_bounds__(1024, 1) k_probe(const Rec* peer_src, Rec* own_dst, int64_t n, int flavour, int t, int me,
                                                   int reads_per_thread, unsigned long long* counts) {
    unsigned long long stale = 0, other = 0, torn = 0, reads = 0;
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u + 977u * t;
    const double want = static_cast<double>(t);   // buffer [t & 1] was written during iteration t - 1 with tag t
    for (int k = 0; k < reads_per_thread; ++k) {
        s ^= s << 13;
        s ^= s >> 17;
        s ^= s << 5;
        const int64_t i = static_cast<int64_t>(s % static_cast<uint32_t>(n));
        const Rec* r = peer_src + i;
        const double* d = reinterpret_cast<const double*>(r);
        double tag, idx, own;
        switch (flavour) {
            case kPlainSoA: tag = d[0]; idx = d[1]; own = d[2]; break;
            case kLdgSoA: tag = __ldg(d); idx = __ldg(d + 1); own = __ldg(d + 2); break;
            case kCvSoA: tag = __ldcv(d); idx = __ldcv(d + 1); own = __ldcv(d + 2); break;
            case kSysSoA: tag = ld_sys(d); idx = ld_sys(d + 1); own = ld_sys(d + 2); break;
            case kPlain16: { const double2 a = reinterpret_cast<const double2*>(d)[0], b = reinterpret_cast<const double2*>(d)[1]; tag = a.x; idx = a.y; own = b.x; break; }
            case kLdg16: { const double2 a = __ldg(reinterpret_cast<const double2*>(d)), b = __ldg(reinterpret_cast<const double2*>(d) + 1); tag = a.x; idx = a.y; own = b.x; break; }
            case kCv16: { const double2 a = __ldcv(reinterpret_cast<const double2*>(d)), b = __ldcv(reinterpret_cast<const double2*>(d) + 1); tag = a.x; idx = a.y; own = b.x; break; }
            case kSys16: { const double2 a = ld_sys2(reinterpret_cast<const double2*>(d)), b = ld_sys2(reinterpret_cast<const double2*>(d) + 1); tag = a.x; idx = a.y; own = b.x; break; }
            case kVol16: { const double2 a = ld_vol2(reinterpret_cast<const double2*>(d)), b = ld_vol2(reinterpret_cast<const double2*>(d) + 1); tag = a.x; idx = a.y; own = b.x; break; }
            default: { double pad; ld_sys4(r, &tag, &idx, &own, &pad); break; }
        }
        ++reads;
        const bool idx_ok = idx == static_cast<double>(i) && own == static_cast<double>(1 - me);
        if (tag == want && idx_ok) continue;
        if (tag == want - 2.0 && idx_ok)
            ++stale;
        else if (idx_ok)
            ++other;
        else
            ++torn;
    }
    // the writer side of the product: this GPU's slice of the OTHER buffer, tagged for the next iteration
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        double2* w = reinterpret_cast<double2*>(own_dst + i);
        w[0] = make_double2(static_cast<double>(t + 1), static_cast<double>(i));
        w[1] = make_double2(static_cast<double>(me), 0.0);
    }
    atomicAdd(counts + 0, reads);
    if (stale) atomicAdd(counts + 1, stale);
    if (other) atomicAdd(counts + 2, other);
    if (torn) atomicAdd(counts + 3, torn);
}

__global__ void k_fill(Rec* buf, int64_t n, double tag, int me) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        buf[i] = Rec{tag, static_cast<double>(i), static_cast<double>(me), 0.0};
}

static const int kIters = 12, kReads = 8;

static void report(const char* transport, int64_t n, int fl, const unsigned long long c[4]) {
    printf("{\"transport\": \"%s\", \"records\": %lld, \"buffer_mb\": %.1f, \"load\": \"%s\", \"iterations\": %d, \"reads\": %llu, "
           "\"stale_two_iterations_old\": %llu, \"other_wrong_tag\": %llu, \"torn\": %llu}\n",
           transport, (long long)n, n * 32.0 / 1048576.0, kNames[fl], kIters, c[0], c[1], c[2], c[3]);
    fflush(stdout);
}

// ---- one process, peer access ---------------------------------------------------------------
static void run_peer(int64_t n) {
    Rec* buf[2][2];
    unsigned long long* cnt[2];
    for (int g = 0; g < 2; ++g) {
        CK(cudaSetDevice(g));
        cudaError_t e = cudaDeviceEnablePeerAccess(1 - g, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        cudaGetLastError();
        for (int b = 0; b < 2; ++b) CK(cudaMalloc(&buf[g][b], n * sizeof(Rec)));
        CK(cudaMalloc(&cnt[g], 4 * sizeof(unsigned long long)));
    }
    for (int fl = 0; fl < kNumFlavours; ++fl) {
        for (int g = 0; g < 2; ++g) {
            CK(cudaSetDevice(g));
            k_fill<<<148, 1024>>>(buf[g][0], n, 0.0, g);    // buffer 0 holds tag 0 for iteration 0
            k_fill<<<148, 1024>>>(buf[g][1], n, -1.0, g);
            CK(cudaMemset(cnt[g], 0, 4 * sizeof(unsigned long long)));
            CK(cudaDeviceSynchronize());
        }
        for (int t = 0; t < kIters; ++t) {
            for (int g = 0; g < 2; ++g) {
                CK(cudaSetDevice(g));
                k_probe<<<148, 1024>>>(buf[1 - g][t & 1], buf[g][(t + 1) & 1], n, fl, t, g, kReads, cnt[g]);
            }
            for (int g = 0; g < 2; ++g) {
                CK(cudaSetDevice(g));
                CK(cudaDeviceSynchronize());
            }
        }
        unsigned long long c[4] = {0, 0, 0, 0}, h[4];
        for (int g = 0; g < 2; ++g) {
            CK(cudaSetDevice(g));
            CK(cudaMemcpy(h, cnt[g], sizeof h, cudaMemcpyDeviceToHost));
            for (int k = 0; k < 4; ++k) c[k] += h[k];
        }
        report("peer-access (one process)", n, fl, c);
    }
    for (int g = 0; g < 2; ++g) {
        CK(cudaSetDevice(g));
        for (int b = 0; b < 2; ++b) cudaFree(buf[g][b]);
        cudaFree(cnt[g]);
    }
}

// ---- two processes, CUDA IPC ----------------------------------------------------------------
static void xfer(int wfd, int rfd, const void* out, void* in, size_t bytes) {
    if (write(wfd, out, bytes) != (ssize_t)bytes) exit(4);
    size_t got = 0;
    while (got < bytes) {
        ssize_t r = read(rfd, static_cast<char*>(in) + got, bytes - got);
        if (r <= 0) exit(5);
        got += r;
    }
}

static void ipc_child(int me, int wfd, int rfd, const std::vector<int64_t>& sizes) {
    CK(cudaSetDevice(me));
    for (int64_t n : sizes) {
        Rec* own[2];
        Rec* peer[2];
        unsigned long long* cnt;
        cudaIpcMemHandle_t mine[2], theirs[2];
        for (int b = 0; b < 2; ++b) {
            CK(cudaMalloc(&own[b], n * sizeof(Rec)));
            CK(cudaIpcGetMemHandle(&mine[b], own[b]));
        }
        CK(cudaMalloc(&cnt, 4 * sizeof(unsigned long long)));
        xfer(wfd, rfd, mine, theirs, sizeof mine);
        for (int b = 0; b < 2; ++b) CK(cudaIpcOpenMemHandle(reinterpret_cast<void**>(&peer[b]), theirs[b], cudaIpcMemLazyEnablePeerAccess));
        for (int fl = 0; fl < kNumFlavours; ++fl) {
            k_fill<<<148, 1024>>>(own[0], n, 0.0, me);
            k_fill<<<148, 1024>>>(own[1], n, -1.0, me);
            CK(cudaMemset(cnt, 0, 4 * sizeof(unsigned long long)));
            CK(cudaDeviceSynchronize());
            char tok = 1, got;
            xfer(wfd, rfd, &tok, &got, 1);
            for (int t = 0; t < kIters; ++t) {
                k_probe<<<148, 1024>>>(peer[t & 1], own[(t + 1) & 1], n, fl, t, me, kReads, cnt);
                CK(cudaDeviceSynchronize());
                xfer(wfd, rfd, &tok, &got, 1);   // both processes have finished iteration t
            }
            unsigned long long h[4], o[4];
            CK(cudaMemcpy(h, cnt, sizeof h, cudaMemcpyDeviceToHost));
            xfer(wfd, rfd, h, o, sizeof h);
            if (me == 0) {
                for (int k = 0; k < 4; ++k) h[k] += o[k];
                report("cuda-ipc (two processes)", n, fl, h);
            }
        }
        for (int b = 0; b < 2; ++b) {
            cudaIpcCloseMemHandle(peer[b]);
        }
        char tok = 1, got;
        xfer(wfd, rfd, &tok, &got, 1);
        for (int b = 0; b < 2; ++b) cudaFree(own[b]);
        cudaFree(cnt);
    }
}

int main(int argc, char** argv) {
    std::vector<int64_t> sizes = {262144, 524288, 1048576, 2097152};
    const bool ipc_only = argc > 1 && !strcmp(argv[1], "ipc");
    const bool peer_only = argc > 1 && !strcmp(argv[1], "peer");
    if (!peer_only) {
        // fork BEFORE any CUDA call: each child owns one GPU
        int ab[2], ba[2];
        if (pipe(ab) || pipe(ba)) return 2;
        fflush(stdout);
        pid_t p0 = fork();
        if (p0 == 0) {
            ipc_child(0, ab[1], ba[0], sizes);
            _exit(0);
        }
        pid_t p1 = fork();
        if (p1 == 0) {
            ipc_child(1, ba[1], ab[0], sizes);
            _exit(0);
        }
        int st0 = 0, st1 = 0;
        waitpid(p0, &st0, 0);
        waitpid(p1, &st1, 0);
        if (st0 || st1) fprintf(stderr, "ipc children exited with %d / %d\n", st0, st1);
    }
    if (!ipc_only) {
        int ndev = 0;
        CK(cudaGetDeviceCount(&ndev));
        if (ndev < 2) {
            fprintf(stderr, "needs 2 GPUs\n");
            return 1;
        }
        for (int64_t n : sizes) run_peer(n);
    }
    return 0;
}
