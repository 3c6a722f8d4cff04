// csrc/map_prep.cpp -- see map_prep.h.
#include "map_prep.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace mclb200 {

namespace {
constexpr int64_t kInf = std::numeric_limits<int64_t>::max() / 4;

// 1-D squared distance transform of f (lower envelope of parabolas).
void dt1d(const int64_t* f, int n, int64_t* d, int* v, double* z) {
    int k = -1;
    for (int q = 0; q < n; ++q) {
        if (f[q] >= kInf) continue;
        if (k < 0) {
            k = 0;
            v[0] = q;
            z[0] = -1e300;
            z[1] = 1e300;
            continue;
        }
        double s;
        for (;;) {
            const int p = v[k];
            s = (static_cast<double>(f[q] + static_cast<int64_t>(q) * q) -
                 static_cast<double>(f[p] + static_cast<int64_t>(p) * p)) /
                (2.0 * (q - p));
            if (s <= z[k] && k > 0) {
                --k;
            } else {
                break;
            }
        }
        if (s <= z[k]) {  // k == 0 and new parabola dominates everywhere
            v[0] = q;
            z[0] = -1e300;
            z[1] = 1e300;
        } else {
            ++k;
            v[k] = q;
            z[k] = s;
            z[k + 1] = 1e300;
        }
    }
    if (k < 0) {
        for (int q = 0; q < n; ++q) d[q] = kInf;
        return;
    }
    int j = 0;
    for (int q = 0; q < n; ++q) {
        while (z[j + 1] < q) ++j;
        const int64_t dq = q - v[j];
        d[q] = dq * dq + f[v[j]];
    }
}
}  // namespace

void edt_squared(const std::vector<uint8_t>& mask, int W, int H, std::vector<int64_t>& out) {
    out.assign(static_cast<size_t>(W) * H, kInf);
    const int n = std::max(W, H);
    std::vector<int64_t> f(n), d(n);
    std::vector<int> v(n);
    std::vector<double> z(n + 1);
    // columns
    for (int x = 0; x < W; ++x) {
        for (int y = 0; y < H; ++y) f[y] = mask[static_cast<size_t>(y) * W + x] ? 0 : kInf;
        dt1d(f.data(), H, d.data(), v.data(), z.data());
        for (int y = 0; y < H; ++y) out[static_cast<size_t>(y) * W + x] = d[y];
    }
    // rows
    for (int y = 0; y < H; ++y) {
        int64_t* row = &out[static_cast<size_t>(y) * W];
        for (int x = 0; x < W; ++x) f[x] = row[x];
        dt1d(f.data(), W, d.data(), v.data(), z.data());
        for (int x = 0; x < W; ++x) row[x] = d[x];
    }
}

bool skip_map_layout(const int8_t* data, int W, int H, SkipMap& out) {
    if (W <= 0 || H <= 0) return false;
    if (static_cast<int64_t>(W) * H > (int64_t{1} << 30)) return false;
    out.W = W;
    out.H = H;
    out.PW = ((W + kPadL + kPadR) + 31) / 32 * 32;
    out.PH = H + kPadL + kPadR;
    out.v8.clear();
    out.v4.clear();
    out.free_cells.clear();
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c)
            if (data[static_cast<size_t>(r) * W + c] == 0) out.free_cells.push_back(r * W + c);
    return true;
}

bool build_skip_map(const int8_t* data, int W, int H, SkipMap& out) {
    if (!skip_map_layout(data, W, H, out)) return false;
    const int PW = out.PW, PH = out.PH;
    const size_t n = static_cast<size_t>(PW) * PH;

    // blocked mask on the P-grid
    std::vector<uint8_t> blocked(n, 1);
    for (int py = 0; py < PH; ++py) {
        const int fy = py - kPadL;  // floor(q_y)
        if (fy < -1 || fy >= H) continue;
        const int ry = std::max(fy, 0);
        for (int px = 0; px < PW; ++px) {
            const int fx = px - kPadL;
            if (fx < -1 || fx >= W) continue;
            const int rx = std::max(fx, 0);
            blocked[static_cast<size_t>(py) * PW + px] = data[static_cast<size_t>(ry) * W + rx] > 50 ? 1 : 0;
        }
    }
    // 3x3 dilation: the gap between two cells' squares is the centre distance to the dilated set
    std::vector<uint8_t> dil(n, 0);
    for (int py = 0; py < PH; ++py) {
        for (int px = 0; px < PW; ++px) {
            uint8_t any = 0;
            for (int dy = -1; dy <= 1 && !any; ++dy) {
                const int yy = py + dy;
                if (yy < 0 || yy >= PH) {
                    any = 1;  // beyond the P-grid is out of bounds, hence blocked
                    break;
                }
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = px + dx;
                    if (xx < 0 || xx >= PW || blocked[static_cast<size_t>(yy) * PW + xx]) {
                        any = 1;
                        break;
                    }
                }
            }
            dil[static_cast<size_t>(py) * PW + px] = any;
        }
    }
    std::vector<int64_t> d2;
    edt_squared(dil, PW, PH, d2);

    out.v8.assign(n, 0);
    for (size_t i = 0; i < n; ++i) {
        if (blocked[i]) {
            out.v8[i] = 0;
        } else if (dil[i]) {
            out.v8[i] = 1;
        } else {
            const double d = std::sqrt(static_cast<double>(d2[i]));
            int adv = static_cast<int>(std::ceil(d - 1e-3));
            adv = std::max(1, std::min(adv, 254));
            out.v8[i] = static_cast<uint8_t>(1 + adv);
        }
    }
    out.v4.assign(n / 2, 0);
    for (size_t i = 0; i < n; i += 2) {
        const uint8_t lo = std::min<uint8_t>(out.v8[i], 15), hi = std::min<uint8_t>(out.v8[i + 1], 15);
        out.v4[i / 2] = static_cast<uint8_t>(lo | (hi << 4));
    }
    return true;
}

void build_gap_map(const SkipMap& sk, std::vector<float>& gap) {
    const size_t n = static_cast<size_t>(sk.PW) * sk.PH;
    std::vector<uint8_t> dil(n);
    for (size_t i = 0; i < n; ++i) dil[i] = sk.v8[i] < 2;
    std::vector<int64_t> d2;
    edt_squared(dil, sk.PW, sk.PH, d2);
    gap.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const double d = std::sqrt(static_cast<double>(std::min<int64_t>(d2[i], int64_t{1} << 40)));
        gap[i] = std::nextafterf(static_cast<float>(d), 0.0f);   // never above the true gap
    }
}

void make_dir_sectors(int M, DirSector* out) {
    const double pi = 3.14159265358979323846;
    const double width = 2.0 * pi / kDirSectors;
    for (int s = 0; s < kDirSectors; ++s) {
        DirSector& sc = out[s];
        const double a0 = s * width - kDirMargin, a1 = (s + 1) * width + kDirMargin;
        const double mid = (s + 0.5) * width, half = width / 2 + kDirMargin;
        sc.ux = std::cos(mid);
        sc.uy = std::sin(mid);
        sc.kappa = 2.0 * std::sin(half / 2) * (1.0 + 1e-9);
        auto range = [&](bool sine, double* lo, double* hi) {
            auto f = [&](double a) { return sine ? std::sin(a) : std::cos(a); };
            *lo = std::min(f(a0), f(a1));
            *hi = std::max(f(a0), f(a1));
            for (int k = -4; k <= 12; ++k) {   // interior extrema at multiples of pi/2
                const double e = k * pi / 2;
                if (e > a0 && e < a1) {
                    *lo = std::min(*lo, f(e));
                    *hi = std::max(*hi, f(e));
                }
            }
            *lo -= 1e-12;
            *hi += 1e-12;
        };
        range(false, &sc.cmin, &sc.cmax);
        range(true, &sc.smin, &sc.smax);
        // cells reachable by samples 0..M of a ray starting anywhere in a cell, +-2 for the
        // neighbour lookups of the edge test
        sc.exl = static_cast<int>(std::floor(std::min(0.0, M * sc.cmin))) - 2;
        sc.exh = static_cast<int>(std::floor(1.0 + std::max(0.0, M * sc.cmax))) + 2;
        sc.eyl = static_cast<int>(std::floor(std::min(0.0, M * sc.smin))) - 2;
        sc.eyh = static_cast<int>(std::floor(1.0 + std::max(0.0, M * sc.smax))) + 2;
    }
}

}  // namespace mclb200
