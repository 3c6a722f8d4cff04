// csrc/dirmap.cuh -- directional skip maps for the ray march.
//
// The isotropic skip code of map_prep.h lets a ray jump as far as the nearest blocked cell in
// ANY direction, so a ray that runs along a corridor wall, or towards a wall at a shallow
// angle, still advances only a few cells per lookup.  A directional map is built per heading
// sector: its code says how many lattice samples a ray WHOSE DIRECTION LIES IN THAT SECTOR can
// skip from anywhere inside the cell.  Same lattice, same hit rule as the reference's cast_ray
// (src/particle_filter.cpp:611-650); only samples that provably cannot be hits are skipped.
//
// Code of one P-cell for one sector (one byte):
//     0x80            blocked (occupied or out of bounds)
//     0x80 | adv      not blocked, but an 8-neighbour is blocked: a sample landing here must
//                     be classified with the cell-edge test of march.cuh before it is trusted
//     adv             no blocked 8-neighbour
// adv in [1, 127]: the next adv-1 lattice samples (unit steps along any direction of the
// sector, from any position inside the cell) cannot be hits; the march lands on sample +adv.
//
// adv is found by cone tracing.  All positions a ray of the sector can occupy after t steps
// lie in  cell ⊕ t*arc.  For t <= kDirTex the bounding box of that set is tested cell by cell;
// beyond, the set is covered by a ball of radius rho0 + kappa*t around the point t steps along
// the sector's mid direction, and the ball is tested against the Euclidean gap map of the
// isotropic transform (a point inside cell m is at least gap[m] away from every blocked cell);
// one successful ball test clears a whole run of t (the gap function is 1-Lipschitz).
//
// dir_code() is MCL_HD: the CUDA build kernel and the CPU test harness (tests/emu) run the
// same source, with explicitly rounded FP64 operations so both produce the same bytes.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "march.cuh"

namespace mclb200 {

// Sectors trade march trips against per-(particle, sector) work: 64 / 32 / 16 / 8 sectors need 4.6 / 4.8 / 5.4 / 7.3
// lookups per warp-ray on the bench cloud, but a lane marches 1.2 / 2.5 / 4.9 / 9.8 beams per unit whose fixed cost
// (record, window base, beam range) is ~150 warp instructions: k_raycast_dir 0.481 ms at 32, 0.454 at 16, 0.538 at 8.
#ifndef MCL_DIR_SECTORS
#define MCL_DIR_SECTORS 16
#endif
constexpr int kDirSectors = MCL_DIR_SECTORS;   // heading sectors over [0, 2 pi)
constexpr double kDirMargin = 0.002;     // validity margin on both sides of a sector (rad)
constexpr int kDirTex = 16;              // steps tested with the exact bounding box
constexpr int kDirMaxAdv = 127;
constexpr int kDirBlocked = 0x80;
constexpr int kDirNear = 0x80;
constexpr int kDirMinBuckets = 2048;     // heading buckets of the sort: bucket half-width < margin
constexpr int kDirMinParticles = 1024;   // smaller single filters stay on the isotropic kernel

struct DirSector {
    double ux, uy;                    // unit vector of the sector's mid direction
    double kappa;                     // 2 sin((width/2 + margin)/2): spread per step
    double cmin, cmax, smin, smax;    // range of cos / sin over [s*w - margin, (s+1)*w + margin]
    int exl, exh, eyl, eyh;           // cells a ray of M steps can reach, relative to its start cell
};

constexpr double kDirRho0 = 0.70711 + 4e-5;   // half diagonal of a cell + position slack (> kEta)
constexpr double kDirEta = 4e-5;

MCL_HD int dir_floor(double v) {
#if defined(__CUDA_ARCH__)
    return __double2int_rd(v);
#else
    return static_cast<int>(__builtin_floor(v));
#endif
}
MCL_HD int dir_ceil(double v) {
#if defined(__CUDA_ARCH__)
    return __double2int_ru(v);
#else
    return static_cast<int>(__builtin_ceil(v));
#endif
}

// Directional code of P-cell (cx, cy).  v8 = isotropic skip map (map_prep.h), gap = Euclidean
// gap (cells) between a cell's square and the nearest blocked cell's square, 0 for code < 2.
MCL_HD uint8_t dir_code(const uint8_t* v8, const float* gap, int PW, int PH, int cx, int cy, const DirSector& sc) {
    const int iso = v8[static_cast<int64_t>(cy) * PW + cx];
    if (iso == 0) return static_cast<uint8_t>(kDirBlocked);
    const double fx = static_cast<double>(cx), fy = static_cast<double>(cy);
    int t = 1;
    while (t <= kDirMaxAdv) {
        const double ft = static_cast<double>(t);
        bool safe = false;
        if (t <= kDirTex) {
            const int xl = dir_floor(nf_add(fx - kDirEta, nf_mul(ft, sc.cmin)));
            const int xh = dir_floor(nf_add(fx + (1.0 + kDirEta), nf_mul(ft, sc.cmax)));
            const int yl = dir_floor(nf_add(fy - kDirEta, nf_mul(ft, sc.smin)));
            const int yh = dir_floor(nf_add(fy + (1.0 + kDirEta), nf_mul(ft, sc.smax)));
            safe = xl >= 0 && yl >= 0 && xh < PW && yh < PH;
            for (int yy = yl; yy <= yh && safe; ++yy)
                for (int xx = xl; xx <= xh; ++xx)
                    if (v8[static_cast<int64_t>(yy) * PW + xx] == 0) {
                        safe = false;
                        break;
                    }
        }
        if (safe) {
            ++t;
            continue;
        }
        const int mx = dir_floor(nf_add(fx + 0.5, nf_mul(ft, sc.ux)));
        const int my = dir_floor(nf_add(fy + 0.5, nf_mul(ft, sc.uy)));
        if (mx < 0 || mx >= PW || my < 0 || my >= PH) break;
        const double g = static_cast<double>(gap[static_cast<int64_t>(my) * PW + mx]) - 1e-3;
        const double rho = nf_add(kDirRho0, nf_mul(sc.kappa, ft));
        if (!(g > rho)) break;
        // E(x_t') >= g - (t' - t) > rho0 + kappa t'   <=>   t' < (g + t - rho0) / (1 + kappa)
        const int T = dir_ceil(nf_div(nf_add(g, ft) - kDirRho0, 1.0 + sc.kappa)) - 1;
        t = (T > t ? T : t) + 1;
    }
    int adv = t < kDirMaxAdv ? t : kDirMaxAdv;
    if (iso >= 2) {
        const int ia = (iso - 1) < kDirMaxAdv ? (iso - 1) : kDirMaxAdv;
        adv = adv > ia ? adv : ia;
    }
    return static_cast<uint8_t>(adv | (iso == 1 ? kDirNear : 0));
}

// Heading sector of the rays that beam j casts from particles of heading bucket b (B buckets
// over [-pi, pi), B a power of two and a multiple of the sector count).  io_j =
// floor((alpha_j - pi) / bucket_width + 1/2) mod B is precomputed per beam; the bucket's mid
// direction then lies in sector ((b + io_j) mod B) / (B / S), and every heading of the bucket is
// within half a bucket width (< kDirMargin) of it.
MCL_HD int dir_sector_of(int bucket, int io_j, int Bmask, int shift) { return ((bucket + io_j) & Bmask) >> shift; }


// heading bucket of the coherence sort (B buckets over [-pi, pi)): particles with nearly equal
// headings cast nearly identical rays, so a warp of bucket-neighbours marches in lock step
// B is a power of two; headings outside [-pi, pi) wrap around (theta = pi is bucket 0).
MCL_HD int theta_bucket(double th, int B) {
    const double tb = (th + 3.14159265358979323846) * (static_cast<double>(B) * 0.15915494309189535);
    if (!(tb > -1e9 && tb < 1e9)) return 0;   // NaN / absurd headings: such rays never march
    return dir_floor(tb) & (B - 1);
}

// io_j of dir_sector_of() for beam angle alpha (host side, once per beam table)
inline int dir_beam_offset(float alpha, int B) {
    const double pi = 3.14159265358979323846;
    const double off = (static_cast<double>(alpha) - pi) / (2.0 * pi / B) + 0.5;
    const long long io = static_cast<long long>(__builtin_floor(off));
    return static_cast<int>(((io % B) + B) % B);
}

// Shared-memory window of one sector's map.  Particles whose start cell lies in the box
// [box_x0, box_x0 + box) x [box_y0, box_y0 + box) (and p_inside) only ever read cells of
// [wx0, wx0 + pitch) x [wy0, wy0 + rows): a ray of the sector reaches at most the sector's
// extents beyond its start cell, and it cannot leave the P-grid (the border is blocked).
// wx0 and pitch are multiples of 16 (bulk-copy alignment; PW is a multiple of 32).
struct DirWindow {
    int wx0, wy0, pitch, rows;
};
MCL_HD DirWindow dir_window(const DirSector& sc, int box_x0, int box_y0, int box, int PW, int PH) {
    int x0 = box_x0 + sc.exl;
    x0 = (x0 < 0 ? 0 : x0) & ~15;
    int x1 = (box_x0 + box + sc.exh + 15) & ~15;
    x1 = x1 > PW ? PW : x1;
    int y0 = box_y0 + sc.eyl;
    y0 = y0 < 0 ? 0 : y0;
    int y1 = box_y0 + box + sc.eyh;
    y1 = y1 > PH ? PH : y1;
    DirWindow w;
    w.wx0 = x0 < PW ? x0 : PW;
    w.wy0 = y0 < PH ? y0 : PH;
    w.pitch = x1 > x0 ? x1 - x0 : 0;
    w.rows = y1 > y0 ? y1 - y0 : 0;
    if (w.pitch == 0 || w.rows == 0) w.pitch = w.rows = 0;   // the box lies beside the grid: no particle uses the window
    return w;
}

// Row pitch of a window in shared memory: the copied width padded to an ODD multiple of 16 bytes.  A warp's lanes
// are heading-neighbours casting the same beam, a few cells apart in x and y: with 4 (mod 8) words per row eight
// consecutive rows start in eight different bank groups, whereas a width that is a multiple of 128 bytes puts every
// row on the same banks.  (16-byte granularity: the rows are the destinations of bulk copies.)
MCL_HD int dir_smem_pitch(int pitch) { return pitch | 16; }

// Largest particle box (multiple of 16 cells, <= 192) whose windows fit `capacity` bytes for
// every sector; 0 if none does.
inline int dir_choose_box(const DirSector* sectors, size_t capacity) {
    for (int box = 192; box >= 16; box -= 16) {
        bool ok = true;
        for (int s = 0; s < kDirSectors && ok; ++s) {
            const DirSector& sc = sectors[s];
            const size_t pitch = static_cast<size_t>(dir_smem_pitch((box + sc.exh - sc.exl + 30) & ~15));
            const size_t rows = static_cast<size_t>(box + sc.eyh - sc.eyl);
            ok = pitch * rows <= capacity;
        }
        if (ok) return box;
    }
    return 0;
}

}  // namespace mclb200
