// csrc/map_kernels.cuh -- the skip map and the gap map of map_prep.h, built on the device.
//
// mcl_set_map prepares, from the int8 grid get_omap receives (src/particle_filter.cpp:190-213), the isotropic
// skip codes (v8 / v4) and the Euclidean gap map the cone tracing of dirmap.cuh queries.  Both come from ONE
// exact squared Euclidean distance transform of the dilated blocked mask on the padded P-lattice.  The host
// version (map_prep.cpp; it stays the CPU twin the emulation tests use) takes 0.37 s for a 2000 x 2000 map;
// here the transform is two separable passes with one thread per column / per row:
//   k_map_masks   blocked / dilated masks of the P-lattice (the trunc-quirk duplicate row and column included)
//   k_edt_cols    per column: distance to the nearest seed along the column (two sweeps)
//   k_edt_rows    per row: lower envelope of the parabolas (Felzenszwalb-Huttenlocher), the same algorithm and
//                 the same FP64 operations as map_prep.cpp::dt1d, scratch arrays laid out so that the threads
//                 of a warp (adjacent rows) touch adjacent words
//   k_map_codes   v8, v4 and gap from the squared distances
// The squared distances are integers and the minimum over the parabolas is exact, so the result equals the
// host build bit for bit (tests compare v8 and the sector maps built from the gap map).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "map_prep.h"

namespace mclb200 {

constexpr int kEdtInf = 1 << 29;

__device__ __forceinline__ int p_blocked(const int8_t* __restrict__ grid, int W, int H, int PW, int PH, int px, int py) {
    if (px < 0 || px >= PW || py < 0 || py >= PH) return 1;   // beyond the P-grid is out of bounds, hence blocked
    const int fx = px - kPadL, fy = py - kPadL;               // floor(q)
    if (fx < -1 || fx >= W || fy < -1 || fy >= H) return 1;
    const int rx = fx < 0 ? 0 : fx, ry = fy < 0 ? 0 : fy;     // quotients in (-1, 0) truncate to cell 0 (:628-629)
    return grid[static_cast<int64_t>(ry) * W + rx] > 50 ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_map_masks(const int8_t* __restrict__ grid, int W, int H, int PW, int PH,
                                                   uint8_t* __restrict__ blocked, uint8_t* __restrict__ dil) {
    const int64_t cell = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (cell >= static_cast<int64_t>(PW) * PH) return;
    const int py = static_cast<int>(cell / PW), px = static_cast<int>(cell - static_cast<int64_t>(py) * PW);
    int any = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) any |= p_blocked(grid, W, H, PW, PH, px + dx, py + dy);
    blocked[cell] = static_cast<uint8_t>(p_blocked(grid, W, H, PW, PH, px, py));
    dil[cell] = static_cast<uint8_t>(any);   // 3x3 dilation: the gap between two cells' squares is the centre distance to this set
}

// gT[x * PH + y] = distance along column x from row y to the nearest seed of the column (kEdtInf: none)
__global__ void __launch_bounds__(128) k_edt_cols(const uint8_t* __restrict__ dil, int PW, int PH, int* __restrict__ gT) {
    const int x = blockIdx.x * 128 + threadIdx.x;
    if (x >= PW) return;
    int d = kEdtInf;
    for (int y = 0; y < PH; ++y) {
        d = dil[static_cast<int64_t>(y) * PW + x] ? 0 : (d >= kEdtInf ? kEdtInf : d + 1);
        gT[static_cast<int64_t>(x) * PH + y] = d;
    }
    d = kEdtInf;
    for (int y = PH - 1; y >= 0; --y) {
        d = dil[static_cast<int64_t>(y) * PW + x] ? 0 : (d >= kEdtInf ? kEdtInf : d + 1);
        const int64_t o = static_cast<int64_t>(x) * PH + y;
        if (d < gT[o]) gT[o] = d;
    }
}

// One thread per row y: d2T[x * PH + y] = min over q of (x - q)^2 + g(q, y)^2 -- map_prep.cpp::dt1d, operation
// for operation.  v / z are the envelope's stacks, [k * PH + y].
__global__ void __launch_bounds__(128) k_edt_rows(const int* __restrict__ gT, int PW, int PH, int* __restrict__ v,
                                                  double* __restrict__ z, long long* __restrict__ d2T) {
    const int y = blockIdx.x * 128 + threadIdx.x;
    if (y >= PH) return;
    constexpr long long kInf64 = 0x1fffffffffffffffll;
    auto f = [&](int q) -> long long {
        const int g = gT[static_cast<int64_t>(q) * PH + y];
        return g >= kEdtInf ? kInf64 : static_cast<long long>(g) * g;
    };
    auto V = [&](int k) -> int& { return v[static_cast<int64_t>(k) * PH + y]; };
    auto Z = [&](int k) -> double& { return z[static_cast<int64_t>(k) * PH + y]; };
    int k = -1;
    for (int q = 0; q < PW; ++q) {
        const long long fq = f(q);
        if (fq >= kInf64) continue;
        if (k < 0) {
            k = 0;
            V(0) = q;
            Z(0) = -1e300;
            Z(1) = 1e300;
            continue;
        }
        double s;
        for (;;) {
            const int p = V(k);
            s = __ddiv_rn(__dsub_rn(static_cast<double>(fq + static_cast<long long>(q) * q),
                                    static_cast<double>(f(p) + static_cast<long long>(p) * p)),
                          __dmul_rn(2.0, static_cast<double>(q - p)));
            if (s <= Z(k) && k > 0)
                --k;
            else
                break;
        }
        if (s <= Z(k)) {   // k == 0 and the new parabola dominates everywhere
            V(0) = q;
            Z(0) = -1e300;
            Z(1) = 1e300;
        } else {
            ++k;
            V(k) = q;
            Z(k) = s;
            Z(k + 1) = 1e300;
        }
    }
    if (k < 0) {
        for (int q = 0; q < PW; ++q) d2T[static_cast<int64_t>(q) * PH + y] = kInf64;
        return;
    }
    int j = 0;
    for (int q = 0; q < PW; ++q) {
        while (Z(j + 1) < static_cast<double>(q)) ++j;
        const long long dq = q - V(j);
        d2T[static_cast<int64_t>(q) * PH + y] = dq * dq + f(V(j));
    }
}

// skip codes (map_prep.h) and the gap map from the squared distances; one thread per PAIR of cells (v4 packs two)
__global__ void __launch_bounds__(256) k_map_codes(const uint8_t* __restrict__ blocked, const uint8_t* __restrict__ dil,
                                                   const long long* __restrict__ d2T, int PW, int PH, uint8_t* __restrict__ v8,
                                                   uint8_t* __restrict__ v4, float* __restrict__ gap) {
    const int64_t pair = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (pair >= static_cast<int64_t>(PW) * PH / 2) return;
    uint8_t code[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int64_t cell = 2 * pair + h;
        const int py = static_cast<int>(cell / PW), px = static_cast<int>(cell - static_cast<int64_t>(py) * PW);
        const long long d2 = d2T[static_cast<int64_t>(px) * PH + py];
        uint8_t c;
        if (blocked[cell]) {
            c = 0;
        } else if (dil[cell]) {
            c = 1;
        } else {
            const double d = sqrt(static_cast<double>(d2));
            int adv = static_cast<int>(ceil(d - 1e-3));
            adv = adv < 1 ? 1 : (adv > 254 ? 254 : adv);
            c = static_cast<uint8_t>(1 + adv);
        }
        code[h] = c;
        v8[cell] = c;
        if (gap) {   // never above the true gap: the float below sqrt(d2) (0 stays 0)
            const long long dc = d2 < (1ll << 40) ? d2 : (1ll << 40);
            const float g = static_cast<float>(sqrt(static_cast<double>(dc)));
            gap[cell] = g > 0.0f ? __uint_as_float(__float_as_uint(g) - 1u) : 0.0f;
        }
    }
    v4[pair] = static_cast<uint8_t>((code[0] < 15 ? code[0] : 15) | ((code[1] < 15 ? code[1] : 15) << 4));
}

}  // namespace mclb200
