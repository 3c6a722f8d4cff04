// csrc/device_utils.cuh -- warp/block primitives and the counter-based RNG used by the
// MCL kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mclb200 {

constexpr unsigned kFullMask = 0xffffffffu;

// First statement of every kernel of the update.  Launched with programmatic stream serialization (launch_dep in
// mcl_b200.cu) the kernel's blocks may already be resident while the preceding kernel drains: griddepcontrol.wait
// returns when that kernel has COMPLETED and its memory operations are visible (so completion is transitive along
// the chain), launch_dependents lets the next kernel's blocks be scheduled as soon as every block of this one has
// started.  Both are no-ops for a normal launch.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---- shuffles for arbitrary trivially-copyable structs (multiples of 4 bytes) ----------
template <class T>
__device__ __forceinline__ T shfl_up_any(const T& v, int delta) {
    static_assert(sizeof(T) % 4 == 0, "size");
    T r;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < static_cast<int>(sizeof(T) / 4); ++i) dst[i] = __shfl_up_sync(kFullMask, src[i], delta);
    return r;
}

// Inclusive block scan with a (possibly non-commutative) associative operator.
// op(left, right).  sm must hold NT/32 elements.  All NT threads must call.
template <int NT, class T, class Op>
__device__ __forceinline__ T block_scan_inclusive(T v, Op op, T* sm, const T& identity) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = shfl_up_any(v, d);
        if (lane >= d) v = op(o, v);
    }
    if (lane == 31) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        T t = lane < NT / 32 ? sm[lane] : identity;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            T o = shfl_up_any(t, d);
            if (lane >= d) t = op(o, t);
        }
        if (lane < NT / 32) sm[lane] = t;
    }
    __syncthreads();
    if (warp > 0) v = op(sm[warp - 1], v);
    __syncthreads();
    return v;
}

// w = p^(1/squash) of the sensor model (src/particle_filter.cpp:577-579, std::pow).  MCL_FAST_POW: exp(y * log p), about half
// the instructions of pow(); its error (~6e-14 relative for p >= 1e-230) is nine orders below the weight tolerance.
#ifndef MCL_FAST_POW
#define MCL_FAST_POW 0
#endif
__device__ __forceinline__ double squash_pow(double p, double inv_squash) {
#if MCL_FAST_POW
    return exp(inv_squash * log(p));
#else
    return pow(p, inv_squash);
#endif
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFullMask, v, d);
    return v;
}

// Sum over the block, result valid in every thread.  sm holds NT/32 doubles.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double t = lane < NT / 32 ? sm[lane] : 0.0;
    t = warp_sum(t);
    __syncthreads();
    return t;
}

// ---- Philox4x32-10 (Salmon et al. 2011): counter-based, one call = 4 x 32 random bits ----
struct Philox4 {
    uint32_t v[4];
};
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += W0;
        k1 += W1;
    }
    return Philox4{{c0, c1, c2, c3}};
}

// generate_canonical<double,53> over a 32-bit engine (bits/random.tcc:3349-3381): two words.
__device__ __forceinline__ double canonical_from_words(uint32_t a, uint32_t b) {
    const double s = static_cast<double>(a) + static_cast<double>(b) * 4294967296.0;
    double r = s / 18446744073709551616.0;
    if (r >= 1.0) r = 0.99999999999999988897769753748;  // nextafter(1, 0)
    return r;
}

// Two standard normals from two 32-bit words (Box-Muller).  The transcendentals run in FP32:
// this is the production noise source (the reference draws from std::mt19937 + Marsaglia polar,
// so no bit pattern has to be matched) and FP32 resolution is far below the noise it models.
__device__ __forceinline__ void normal_pair(uint32_t a, uint32_t b, double* n0, double* n1) {
    const float u1 = (static_cast<float>(a >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1), 24 bits
    const float u2 = (static_cast<float>(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    *n0 = static_cast<double>(r * c);
    *n1 = static_cast<double>(r * s);
}

}  // namespace mclb200
