// csrc/march.cuh -- the ray march on the reference's sample lattice, with skipping.
//
// Replaces ParticleFilter::cast_ray (src/particle_filter.cpp:611-650).  The reference visits
// lattice samples k = 1..M at c_k = c_{k-1} + (cos a, sin a)*res (accumulated in double) and
// returns at the first sample whose cell is out of bounds or occupied (> 50).  Here the
// same lattice is walked in 9.23 fixed point relative to the ray's start cell; the skip map
// (map_prep.h) says how many following samples cannot be hits, so most samples are never
// touched.  The answer is the reference's step index r (hit at sample r+1) or M (no hit).
//
// Exactness.  The fixed-point position of sample k differs from the reference's computed
// quotient by less than kEta cells (kEtaFix units): 2^-24 for the start, k*2^-24 for the
// rounded direction, ~1e-10 for the reference's own accumulated rounding.  A sample is only
// ever *classified* (hit / not hit) in a cell of code 0 or 1; if it then lies within kEta of
// a cell edge and the cell across that edge has the other class, the reference's arithmetic
// is replayed operation for operation in FP64 (replay_sample_is_hit) and decides.  Skips are
// conservative by construction (map_prep.h), so every classification equals the reference's.
//
// All functions are MCL_HD so the same code is exercised on the CPU by tests/emu (test
// harness only -- the product has no CPU path).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MCL_HD __host__ __device__ __forceinline__
#define MCL_D __device__ __forceinline__
#else
#define MCL_HD inline
#define MCL_D inline
#endif

namespace mclb200 {

constexpr int kFrac = 23;                         // fraction bits of the ray-local position
constexpr uint32_t kOne = 1u << kFrac;
constexpr uint32_t kFracMask = kOne - 1u;
constexpr int kLocalOrigin = 256;                 // ray-local coordinate of the start cell
constexpr uint32_t kEtaFix = 160;                 // > (M+2) * 2^-24 in 2^-23 units, M <= 254
constexpr int kMaxRangePxSupported = 254;

// Geometry of the reference grid, needed by the exact replay.
struct RefGrid {
    const int8_t* data;   // row-major int8 occupancy (device or host pointer)
    int W, H;
    double res, ox, oy;
};

// FP64 helpers that must not be contracted into FMAs: the reference is built without FMA
// (CMakeLists.txt:5-11, no -march), see SURVEY F11.
#if defined(__CUDA_ARCH__)
MCL_D double nf_add(double a, double b) { return __dadd_rn(a, b); }
MCL_D double nf_sub(double a, double b) { return __dsub_rn(a, b); }
MCL_D double nf_mul(double a, double b) { return __dmul_rn(a, b); }
MCL_D double nf_div(double a, double b) { return __ddiv_rn(a, b); }
MCL_D int d2i_trunc(double a) { return __double2int_rz(a); }
#else
inline double nf_add(double a, double b) { volatile double r = a + b; return r; }
inline double nf_sub(double a, double b) { volatile double r = a - b; return r; }
inline double nf_mul(double a, double b) { volatile double r = a * b; return r; }
inline double nf_div(double a, double b) { volatile double r = a / b; return r; }
inline int d2i_trunc(double a) { return static_cast<int>(a); }
#endif

// The reference's arithmetic for sample k (1-based) of the ray from (x, y) with step
// (dx, dy) = (cos a * res, sin a * res): :619-646.  O(k); only called on ambiguous samples.
MCL_HD bool replay_sample_is_hit(const RefGrid& g, double x, double y, double dx, double dy, int k) {
    double cx = x, cy = y;
    for (int s = 0; s < k; ++s) {
        cx = nf_add(cx, dx);
        cy = nf_add(cy, dy);
    }
    const int gx = d2i_trunc(nf_div(nf_sub(cx, g.ox), g.res));
    const int gy = d2i_trunc(nf_div(nf_sub(cy, g.oy), g.res));
    if (gx < 0 || gx >= g.W || gy < 0 || gy >= g.H) return true;
    return g.data[static_cast<int64_t>(gy) * g.W + gx] > 50;
}

// Fixed-point start of a ray inside P-cell coordinates.  q = (x - origin)/res + kPadL.
struct RayStart {
    uint32_t p0x, p0y;   // ray-local 9.23 position of the start (integer part == kLocalOrigin)
    int bx, by;          // P-cell of local coordinate 0:  cell = b + (p >> kFrac)
};

MCL_HD RayStart make_ray_start(double qx, double qy, int fqx, int fqy) {
    // fqx = floor(qx) as int; frac in [0,1)
    RayStart s;
    const double fx = qx - static_cast<double>(fqx);
    const double fy = qy - static_cast<double>(fqy);
    uint32_t ux = static_cast<uint32_t>(fx * static_cast<double>(kOne) + 0.5);
    uint32_t uy = static_cast<uint32_t>(fy * static_cast<double>(kOne) + 0.5);
    s.p0x = (static_cast<uint32_t>(kLocalOrigin) << kFrac) + ux;   // ux may equal kOne: carries into the cell
    s.p0y = (static_cast<uint32_t>(kLocalOrigin) << kFrac) + uy;
    s.bx = fqx - kLocalOrigin;
    s.by = fqy - kLocalOrigin;
    return s;
}

// Skip-map accessors.  get(lx, ly) returns the skip code of the cell at ray-local integer
// coordinates (lx, ly) = (p >> kFrac); the particle's base cell is folded into the accessor
// when it is built (make_*), so the march spends no instructions on it.  Callers guarantee
// that every cell a ray can reach is in range.
struct GlobalV8 {
    const uint8_t* base;   // &v8[by * PW + bx]
    int PW;
    MCL_HD int get_p(uint32_t px, uint32_t py) const { return get(static_cast<int>(px >> kFrac), static_cast<int>(py >> kFrac)); }
    MCL_HD int get(int lx, int ly) const {
#if defined(__CUDA_ARCH__)
        return __ldg(base + static_cast<int64_t>(ly) * PW + lx);
#else
        return base[static_cast<int64_t>(ly) * PW + lx];
#endif
    }
};
MCL_HD GlobalV8 make_global_v8(const uint8_t* v8, int PW, int bx, int by) {
    return GlobalV8{v8 + static_cast<int64_t>(by) * PW + bx, PW};
}

// 4-bit window (host pointer form, used by the CPU emulation harness): P-cells
// [wx0, wx0+ww) x [wy0, wy0+wh), wx0 even, two cells per byte.
struct WindowV4 {
    const uint8_t* w4;   // wh rows of pitch bytes
    int offx, offy, pitch;   // bx - wx0, by - wy0
    MCL_HD int get_p(uint32_t px, uint32_t py) const { return get(static_cast<int>(px >> kFrac), static_cast<int>(py >> kFrac)); }
    MCL_HD int get(int lx, int ly) const {
        const int x = lx + offx, y = ly + offy;
        const int b = w4[y * pitch + (x >> 1)];
        return (b >> ((x & 1) << 2)) & 15;
    }
};
MCL_HD WindowV4 make_window_v4(const uint8_t* w4, int wx0, int wy0, int pitch, int bx, int by) {
    return WindowV4{w4, bx - wx0, by - wy0, pitch};
}

struct ReplayArgs {
    double x, y;        // particle position (metres)
    double theta;       // particle heading
    float beam;         // beam angle (float32, as downsampled_angles_ holds it)
};

#if defined(__CUDA_ARCH__)
MCL_D void sincos_ref(double a, double* s, double* c) { sincos(a, s, c); }
#else
}  // namespace mclb200
#include <cmath>
namespace mclb200 {
inline void sincos_ref(double a, double* s, double* c) { *s = std::sin(a); *c = std::cos(a); }
#endif


// P-lattice coordinate of a world coordinate: the reference's quotient (:628-629) + kPadL.
MCL_HD double p_coord(double c, double origin, double res, int pad_l) {
    return nf_div(nf_sub(c, origin), res) + static_cast<double>(pad_l);
}

// Fixed-point direction of beam (ca, sa) = (cos, sin)(beam angle) for a particle with heading
// (cth, sth): cos/sin(theta + alpha) by angle addition, rounded to 2^-23.
MCL_HD void beam_direction_fixed(double cth, double sth, double ca, double sa, int* dxf, int* dyf) {
    // callers may pass (cth, sth) already multiplied by 2^23 (scale == 1): scaling by a power of
    // two commutes with the roundings, so both forms give the same integers
    const double c = cth * ca - sth * sa;
    const double s = sth * ca + cth * sa;
#if defined(__CUDA_ARCH__)
    *dxf = __double2int_rn(c * static_cast<double>(kOne));
    *dyf = __double2int_rn(s * static_cast<double>(kOne));
#else
    *dxf = static_cast<int>(__builtin_lrint(c * static_cast<double>(kOne)));
    *dyf = static_cast<int>(__builtin_lrint(s * static_cast<double>(kOne)));
#endif
}

#if defined(__CUDACC__)
// (cths, sths) = (cos, sin)(theta) * 2^23, computed once per particle
__device__ __forceinline__ void beam_direction_prescaled(double cths, double sths, double ca, double sa, int* dxf,
                                                         int* dyf) {
    *dxf = __double2int_rn(cths * ca - sths * sa);
    *dyf = __double2int_rn(sths * ca + cths * sa);
}
#endif

// Can a particle at P-coordinates (qx, qy) be marched at all?  Outside, its first sample is
// already out of bounds for every beam (:632-636) and the step index is 0.
MCL_HD bool p_inside(double qx, double qy, int PW, int PH) {
    return (qx >= 2.0) && (qx < static_cast<double>(PW - 2)) && (qy >= 2.0) && (qy < static_cast<double>(PH - 2));
}

#if defined(__CUDA_ARCH__)
#define MCL_NOINLINE __device__ __noinline__
#else
#define MCL_NOINLINE
#endif

// Rare path of the march: sample k lies within kEta of a cell edge while sitting in a cell of
// code 0/1.  If a cell across a close edge has the other class, the reference's FP64
// arithmetic decides.  Kept out of line so none of it (FP64 sincos, replay loop) is hoisted
// into the per-ray or per-sample code.
template <class Acc>
MCL_NOINLINE int resolve_uncertain(const Acc& acc, uint32_t px, uint32_t py, int v, const RefGrid& g,
                                   const ReplayArgs& ra, int k, int* replays) {
    const bool hit = (v == 0);
    const int cx = static_cast<int>(px >> kFrac), cy = static_cast<int>(py >> kFrac);
    const uint32_t fx = px & kFracMask, fy = py & kFracMask;
    const bool ux = ((fx + kEtaFix) & kFracMask) < 2u * kEtaFix;
    const bool uy = ((fy + kEtaFix) & kFracMask) < 2u * kEtaFix;
    const int nx = cx + (fx < kEtaFix ? -1 : 1);
    const int ny = cy + (fy < kEtaFix ? -1 : 1);
    bool differs = false;
    if (ux) differs |= ((acc.get(nx, cy) == 0) != hit);
    if (uy) differs |= ((acc.get(cx, ny) == 0) != hit);
    if (ux && uy) differs |= ((acc.get(nx, ny) == 0) != hit);
    if (!differs) return v;
    double sn, cs;
    sincos_ref(nf_add(ra.theta, static_cast<double>(ra.beam)), &sn, &cs);   // theta + angle (:533)
    if (replays) ++*replays;
    return replay_sample_is_hit(g, ra.x, ra.y, nf_mul(cs, g.res), nf_mul(sn, g.res), k) ? 0 : 1;
}

// March one ray.  (dxf, dyf) = round(cos a * 2^23), round(sin a * 2^23).  Returns the step
// index r in [0, M] (M == no hit).  `replays` (nullable) counts exact replays for diagnostics.
//
// The hot loop contains no call and no break: a sample that needs the exact path parks the
// loop counter beyond M, the loop ends through its own condition, the out-of-line resolver
// runs and the loop is re-entered.  (A call inside the loop forces the loop invariants to be
// reloaded every iteration; a break out of the near-wall branch makes the branch unstructured,
// and the lanes of a warp then stop reconverging inside the loop.)
template <class Acc>
MCL_HD int march_ray(const Acc& acc, const RayStart& st, int dxf, int dyf, int M, const RefGrid& g,
                     const ReplayArgs& ra, int* replays) {
    constexpr int kPark = 1 << 20;
    int k = 1, r = M;
    for (;;) {
        int pending_k = 0;   // sample to resolve exactly (0 = none)
        do {
            const uint32_t px = st.p0x + static_cast<uint32_t>(k * dxf);
            const uint32_t py = st.p0y + static_cast<uint32_t>(k * dyf);
            int v = acc.get_p(px, py);   // skip code of the sample's cell
            if (v < 2) {
                // code 0 (blocked) or 1 (next to blocked): the class of this very sample matters
                const uint32_t tx = (px + kEtaFix) & kFracMask, ty = (py + kEtaFix) & kFracMask;
                if ((tx < ty ? tx : ty) < 2u * kEtaFix) {
                    pending_k = k;   // within kEta of a cell edge
                    k = kPark;
                } else if (v == 0) {
                    r = k - 1;
                    k = kPark;
                }
                v = 2;
            }
            k += v - 1;
        } while (k <= M);
        if (pending_k == 0) return r;
        k = pending_k;
        const uint32_t px = st.p0x + static_cast<uint32_t>(k * dxf);
        const uint32_t py = st.p0y + static_cast<uint32_t>(k * dyf);
        const int v = resolve_uncertain(acc, px, py, acc.get_p(px, py), g, ra, k, replays);
        if (v == 0) return k - 1;
        k += 1;
        if (k > M) return M;
    }
}

// ------------------------------------------------------------------------------------------
// March over a DIRECTIONAL skip map (dirmap.cuh): code 0x80 = blocked, bit 7 = "a neighbour is
// blocked" (the landing sample needs the cell-edge test), low 7 bits = advance.  Same lattice
// and the same exactness rule as march_ray; `rep` loads the particle pose only if a sample
// must be replayed in FP64 (rep.load() -> ReplayArgs).
// ------------------------------------------------------------------------------------------
// Returns bit 0 = the sample is a hit, bit 1 = the FP64 replay was needed.  `rep` is passed by
// value (a few registers): rep.grid() is the reference grid, rep.load() the particle's pose.
template <class Acc, class Rep>
MCL_NOINLINE int resolve_uncertain_dir(const Acc acc, uint32_t px, uint32_t py, int v, const Rep rep, int k) {
    const bool hit = (v == 0x80);
    const int cx = static_cast<int>(px >> kFrac), cy = static_cast<int>(py >> kFrac);
    const uint32_t fx = px & kFracMask, fy = py & kFracMask;
    const bool ux = ((fx + kEtaFix) & kFracMask) < 2u * kEtaFix;
    const bool uy = ((fy + kEtaFix) & kFracMask) < 2u * kEtaFix;
    const int nx = cx + (fx < kEtaFix ? -1 : 1);
    const int ny = cy + (fy < kEtaFix ? -1 : 1);
    bool differs = false;
    if (ux) differs |= ((acc.get(nx, cy) == 0x80) != hit);
    if (uy) differs |= ((acc.get(cx, ny) == 0x80) != hit);
    if (ux && uy) differs |= ((acc.get(nx, ny) == 0x80) != hit);
    if (!differs) return hit ? 1 : 0;
    const ReplayArgs ra = rep.load();
    const RefGrid g = rep.grid();
    double sn, cs;
    sincos_ref(nf_add(ra.theta, static_cast<double>(ra.beam)), &sn, &cs);   // theta + angle (:533)
    return 2 | (replay_sample_is_hit(g, ra.x, ra.y, nf_mul(cs, g.res), nf_mul(sn, g.res), k) ? 1 : 0);
}

// A sample is "near an edge" for the hot loop if its fraction lies within 2^-14 cell of a cell
// edge -- a superset of the kEta band that costs two logic operations; the resolver applies the
// exact band.
constexpr uint32_t kEdgeMask = kFracMask & ~511u;
static_assert(2u * kEtaFix <= 512u, "the coarse edge band must contain the kEta band");

// First lattice sample worth looking at.  The ray starts (k = 0, not a sample) somewhere in cell c0; c0's code says
// how many samples a ray of THIS SECTOR can skip from anywhere inside c0, so samples 1 .. adv0 - 1 cannot be hits
// and the march may begin at sample adv0 -- the same guarantee every later jump relies on, used once more.  One
// lookup per (particle, sector) replaces the first lookup of each of the particle's 2-3 rays in that sector.
template <class Acc>
MCL_HD int dir_first_sample(const Acc& acc, const RayStart& st) {
    const int adv = acc.get_p(st.p0x, st.p0y) & 0x7f;   // blocked start cell: adv 0 -> the march begins at sample 1
    return adv > 1 ? adv : 1;
}

// k0: first sample to look at (1, or dir_first_sample's answer)
//
// The hot loop is branch-free (15 instructions per lookup, no divergence inside a trip).  Positions carry a constant
// +kEta offset, so ONE mask test per axis tells whether a sample lies within kEta of a cell edge (the band
// [-kEta, 2^-14 - kEta) around the edge, a superset of the +-kEta band the resolver applies).  The looked-up cell is
// the cell of the OFFSET position: for a sample clear of edges that is the sample's own cell; for a sample inside
// the band it may be the neighbour across the edge, which is harmless --
//   * a free-space code (< 0x80) is only used to skip: the true position is then within 2 kEta < kDirEta of the
//     looked-up cell, which the cone tracing of dirmap.cuh covers (its boxes and balls carry kDirEta of slack), and
//     the sample cannot be a hit (a blocked cell's neighbours all carry the near flag);
//   * a near or blocked code (>= 0x80) inside the band stops the loop, and the resolver below works on the
//     un-offset position exactly as before.
struct RayStartOfs {
    uint32_t ox, oy;   // RayStart::p0x / p0y + kEtaFix
};
MCL_HD RayStartOfs offset_ray_start(const RayStart& st) { return RayStartOfs{st.p0x + kEtaFix, st.p0y + kEtaFix}; }

template <class Acc, class Rep>
MCL_HD int march_ray_dir(const Acc& acc, const RayStartOfs& so, int dxf, int dyf, int M, const Rep& rep, int* replays, int k0 = 1) {
    static_assert(2u * kEtaFix < (1u << 9), "the edge band of the hot loop must contain the kEta band after the offset");
    const uint32_t ox = so.ox, oy = so.oy;
    int k = k0;
    if (k > M) return M;
    for (;;) {
        int adv;
        unsigned stop;
        uint32_t pxo, pyo;   // offset position of the sample just looked at
        do {
            pxo = ox + static_cast<uint32_t>(k * dxf);
            pyo = oy + static_cast<uint32_t>(k * dyf);
            const int v = acc.get_p(pxo, pyo);
            const unsigned edge = static_cast<unsigned>((pxo & kEdgeMask) == 0u) | static_cast<unsigned>((pyo & kEdgeMask) == 0u);
            // blocked, or next to a blocked cell and close to an edge: the class of this very sample must be settled
            stop = static_cast<unsigned>(v == 0x80) | (static_cast<unsigned>(v > 0x80) & edge);
            adv = v & 0x7f;
            k += adv;
        } while ((stop == 0u) & (k <= M));
        if (stop == 0u) return M;
        // the sample that stopped the loop (opaque to the compiler, which would otherwise carry a copy of k and the
        // products k * d through every trip to have them ready here)
#if defined(__CUDA_ARCH__)
        asm volatile("sub.s32 %0, %0, %1;" : "+r"(k) : "r"(adv));
        asm volatile("" : "+r"(pxo), "+r"(pyo));
#else
        k -= adv;
#endif
        if (((pxo & kEdgeMask) != 0u) & ((pyo & kEdgeMask) != 0u)) return k - 1;   // blocked, clear of edges
        const uint32_t px = pxo - kEtaFix, py = pyo - kEtaFix;
        const int res = resolve_uncertain_dir(acc, px, py, acc.get_p(px, py), rep, k);
        *replays += res >> 1;
        if (res & 1) return k - 1;
        k += 1;   // the sample's true cell may be the neighbour: take a single step
        if (k > M) return M;
    }
}

}  // namespace mclb200
