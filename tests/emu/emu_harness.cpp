// tests/emu/emu_harness.cpp -- TEST HARNESS ONLY (never linked into the product).
//
// Compiles the MCL_HD algorithm headers of the CUDA path (march.cuh, exact_sum.cuh) and the
// host map preprocessing (map_prep.cpp) for the CPU, so that the *logic* the kernels run --
// the skip-map march with its exact-replay rule, and the sequential-order sum algebra -- can
// be checked against the oracle in this GPU-less container (`pytest -m "not gpu"`).  The
// block-level scans of the kernels are replaced by plain loops here; everything per-ray and
// per-chunk is the same source the GPU compiles.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../monte_carlo_localization_b200/csrc/exact_sum.cuh"
#include "../../monte_carlo_localization_b200/csrc/map_prep.h"
#include "../../monte_carlo_localization_b200/csrc/march.cuh"

using namespace mclb200;

struct emu_map {
    SkipMap skip;
    std::vector<int8_t> grid;
    double res, ox, oy;
    int M;
    // directional skip maps (dirmap.cuh), built sector by sector on first use
    std::vector<float> gap;
    DirSector sectors[kDirSectors];
    std::vector<uint8_t> dir[kDirSectors];
};

// accessor over a whole sector map (the kernel's global-memory fallback)
struct DirGlobal {
    const uint8_t* base;
    int PW;
    long long* lookups;
    int get_p(uint32_t px, uint32_t py) const {
        if (lookups) ++*lookups;
        return get(static_cast<int>(px >> kFrac), static_cast<int>(py >> kFrac));
    }
    int get(int lx, int ly) const { return base[static_cast<int64_t>(ly) * PW + lx]; }
};
struct ReplayNow {
    ReplayArgs ra;
    RefGrid rg;
    ReplayArgs load() const { return ra; }
    RefGrid grid() const { return rg; }
};

static const std::vector<uint8_t>& emu_dir_sector(emu_map* m, int s) {
    const SkipMap& sk = m->skip;
    if (m->gap.empty()) {
        build_gap_map(sk, m->gap);
        make_dir_sectors(m->M, m->sectors);
    }
    std::vector<uint8_t>& d = m->dir[s];
    if (d.empty()) {
        d.resize(static_cast<size_t>(sk.PW) * sk.PH);
#pragma omp parallel for schedule(dynamic, 8)   // test harness only: rows of one sector map in parallel
        for (int cy = 0; cy < sk.PH; ++cy)
            for (int cx = 0; cx < sk.PW; ++cx)
                d[static_cast<size_t>(cy) * sk.PW + cx] = dir_code(sk.v8.data(), m->gap.data(), sk.PW, sk.PH, cx, cy, m->sectors[s]);
    }
    return d;
}

extern "C" {

emu_map* emu_map_create(const int8_t* data, int W, int H, float resolution, double ox, double oy, double max_range) {
    auto* m = new emu_map();
    if (!build_skip_map(data, W, H, m->skip)) {
        delete m;
        return nullptr;
    }
    m->grid.assign(data, data + static_cast<size_t>(W) * H);
    m->res = resolution;
    m->ox = ox;
    m->oy = oy;
    m->M = static_cast<int>(max_range / m->res);
    return m;
}
void emu_map_destroy(emu_map* m) { delete m; }
int emu_map_dims(const emu_map* m, int* PW, int* PH, int* M) {
    *PW = m->skip.PW;
    *PH = m->skip.PH;
    *M = m->M;
    return 0;
}
void emu_map_v8(const emu_map* m, uint8_t* out) { std::memcpy(out, m->skip.v8.data(), m->skip.v8.size()); }
void emu_map_v4(const emu_map* m, uint8_t* out) { std::memcpy(out, m->skip.v4.data(), m->skip.v4.size()); }

// Step indices for n particles x R beams, the way k_raycast_weight computes them.
// mode 0: global v8 accessor; mode 1: v4 window [wx0, wx0+ww) x [wy0, wy0+wh) (particles that
// are not window-safe fall back to v8, like the kernel).  Returns the number of exact replays.
long long emu_range_steps(const emu_map* m, const double* px, const double* py, const double* pt, long long n,
                          const float* angles, int R, int mode, int wx0, int wy0, int ww, int wh, uint8_t* steps_out,
                          long long* iters_out) {
    const SkipMap& sk = m->skip;
    const int M = m->M;
    std::vector<double> ca(R), sa(R);
    for (int j = 0; j < R; ++j) {
        ca[j] = std::cos(static_cast<double>(angles[j]));
        sa[j] = std::sin(static_cast<double>(angles[j]));
    }
    // window copy of the nibble map
    std::vector<uint8_t> win;
    int vx0 = 0, vx1 = 0, vy0 = 0, vy1 = 0;
    const int pitch = ww / 2;
    if (mode == 1) {
        win.resize(static_cast<size_t>(pitch) * wh);
        for (int r = 0; r < wh; ++r)
            std::memcpy(&win[static_cast<size_t>(r) * pitch], &sk.v4[static_cast<size_t>(wy0 + r) * (sk.PW / 2) + wx0 / 2], pitch);
        vx0 = (wx0 == 0) ? 2 : wx0 + M + 2;
        vx1 = (wx0 + ww >= sk.PW) ? sk.PW - 2 : wx0 + ww - M - 2;
        vy0 = (wy0 == 0) ? 2 : wy0 + M + 2;
        vy1 = (wy0 + wh >= sk.PH) ? sk.PH - 2 : wy0 + wh - M - 2;
    }
    const RefGrid rg{m->grid.data(), sk.W, sk.H, m->res, m->ox, m->oy};
    int replays = 0;
    long long iters = 0;
    (void)iters;
    for (long long i = 0; i < n; ++i) {
        const double x = px[i], y = py[i], th = pt[i];
        const double sth = std::sin(th), cth = std::cos(th);
        const double qx = p_coord(x, m->ox, m->res, kPadL), qy = p_coord(y, m->oy, m->res, kPadL);
        uint8_t* out = steps_out + i * R;
        if (!p_inside(qx, qy, sk.PW, sk.PH)) {
            for (int j = 0; j < R; ++j) out[j] = 0;
            continue;
        }
        const int fqx = static_cast<int>(std::floor(qx)), fqy = static_cast<int>(std::floor(qy));
        const RayStart st = make_ray_start(qx, qy, fqx, fqy);
        const bool in_win = mode == 1 && fqx >= vx0 && fqx < vx1 && fqy >= vy0 && fqy < vy1;
        const WindowV4 wacc = make_window_v4(win.data(), wx0, wy0, pitch, st.bx, st.by);
        const GlobalV8 gacc = make_global_v8(sk.v8.data(), sk.PW, st.bx, st.by);
        for (int j = 0; j < R; ++j) {
            int dxf, dyf;
            beam_direction_fixed(cth, sth, ca[j], sa[j], &dxf, &dyf);
            const ReplayArgs ra{x, y, th, angles[j]};
            const int r = in_win ? march_ray(wacc, st, dxf, dyf, M, rg, ra, &replays)
                                 : march_ray(gacc, st, dxf, dyf, M, rg, ra, &replays);
            out[j] = static_cast<uint8_t>(r);
        }
    }
    if (iters_out) *iters_out = iters;
    return replays;
}

// Host-side helpers of the directional stage, exposed for property tests.
int emu_dir_constants(int* sectors, double* margin) {
    *sectors = kDirSectors;
    *margin = kDirMargin;
    return kDirMinBuckets;
}
int emu_dir_sector(double theta, float alpha, int B) {
    int shift = 0;
    while ((B >> shift) > kDirSectors) ++shift;
    return dir_sector_of(theta_bucket(theta, B), dir_beam_offset(alpha, B), B - 1, shift);
}
// window geometry for a box at (bx0, by0): out = wx0, wy0, pitch, rows per sector; returns the box side
// dir_choose_box picks for `capacity` bytes
int emu_dir_windows(emu_map* m, int bx0, int by0, long long capacity, int* out) {
    emu_dir_sector(m, 0);
    const int box = dir_choose_box(m->sectors, static_cast<size_t>(capacity));
    for (int s = 0; s < kDirSectors; ++s) {
        const DirWindow w = dir_window(m->sectors[s], bx0, by0, box, m->skip.PW, m->skip.PH);
        out[4 * s + 0] = w.wx0;
        out[4 * s + 1] = w.wy0;
        out[4 * s + 2] = w.pitch;
        out[4 * s + 3] = w.rows;
    }
    return box;
}

// One sector's directional map (PH*PW bytes), as the build kernel writes it.
void emu_dir_map(emu_map* m, int sector, uint8_t* out) {
    const std::vector<uint8_t>& d = emu_dir_sector(m, sector);
    std::memcpy(out, d.data(), d.size());
}

// Step indices the way k_raycast_dir computes them: heading bucket -> sector per beam, march
// over that sector's map; use_window: particles inside the box around the cloud centre read
// a copy of the sector's window (the shared-memory path), the others the whole map.
// lookups_out (nullable): skip-map lookups per ray.
long long emu_range_steps_dir(emu_map* m, const double* px, const double* py, const double* pt, long long n,
                              const float* angles, int R, int B, int use_window, int box, uint8_t* steps_out,
                              int32_t* lookups_out) {
    const SkipMap& sk = m->skip;
    const int M = m->M;
    emu_dir_sector(m, 0);
    std::vector<double> ca(R), sa(R);
    std::vector<int> io(R);
    for (int j = 0; j < R; ++j) {
        ca[j] = std::cos(static_cast<double>(angles[j]));
        sa[j] = std::sin(static_cast<double>(angles[j]));
        io[j] = dir_beam_offset(angles[j], B);
    }
    int shift = 0;
    while ((B >> shift) > kDirSectors) ++shift;
    // cloud centre -> box, like the kernel
    double mx = 0, my = 0;
    for (long long i = 0; i < n; ++i) {
        mx += px[i];
        my += py[i];
    }
    mx /= static_cast<double>(n);
    my /= static_cast<double>(n);
    const int bcx = static_cast<int>(std::floor((mx - m->ox) / m->res + kPadL));
    const int bcy = static_cast<int>(std::floor((my - m->oy) / m->res + kPadL));
    const int box_x0 = bcx - box / 2, box_y0 = bcy - box / 2;
    std::vector<std::vector<uint8_t>> win(kDirSectors);
    std::vector<DirWindow> wgeo(kDirSectors);
    const RefGrid rg{m->grid.data(), sk.W, sk.H, m->res, m->ox, m->oy};
    int replays = 0;
    long long oob = 0;
    for (long long i = 0; i < n; ++i) {
        const double x = px[i], y = py[i], th = pt[i];
        const double sth = std::sin(th), cth = std::cos(th);
        const double qx = p_coord(x, m->ox, m->res, kPadL), qy = p_coord(y, m->oy, m->res, kPadL);
        uint8_t* out = steps_out + i * R;
        if (!p_inside(qx, qy, sk.PW, sk.PH)) {
            for (int j = 0; j < R; ++j) {
                out[j] = 0;
                if (lookups_out) lookups_out[i * R + j] = 0;
            }
            continue;
        }
        const int fqx = static_cast<int>(std::floor(qx)), fqy = static_cast<int>(std::floor(qy));
        const RayStart st = make_ray_start(qx, qy, fqx, fqy);
        const int bucket = theta_bucket(th, B);
        const bool in_box = use_window && fqx >= box_x0 && fqx < box_x0 + box && fqy >= box_y0 && fqy < box_y0 + box;
        int k0_sector = -1, k0 = 1;   // like the kernel: one start-cell lookup per (particle, sector)
        for (int j = 0; j < R; ++j) {
            int dxf, dyf;
            beam_direction_fixed(cth, sth, ca[j], sa[j], &dxf, &dyf);
            const int s = dir_sector_of(bucket, io[j], B - 1, shift);
            const std::vector<uint8_t>& d = emu_dir_sector(m, s);
            const ReplayNow rep{ReplayArgs{x, y, th, angles[j]}, rg};
            long long lk = 0;
            int r;
            if (in_box) {
                if (win[s].empty()) {
                    wgeo[s] = dir_window(m->sectors[s], box_x0, box_y0, box, sk.PW, sk.PH);
                    // poison outside the window so that a read beyond it would change the result
                    win[s].assign(static_cast<size_t>(wgeo[s].pitch) * wgeo[s].rows, 0x80);
                    for (int row = 0; row < wgeo[s].rows; ++row)
                        std::memcpy(&win[s][static_cast<size_t>(row) * wgeo[s].pitch],
                                    &d[static_cast<size_t>(wgeo[s].wy0 + row) * sk.PW + wgeo[s].wx0], wgeo[s].pitch);
                }
                const DirWindow& w = wgeo[s];
                // bounds-checked accessor: the window must contain every cell the march reads
                struct Checked {
                    const uint8_t* w;
                    int offx, offy, pitch, rows;
                    long long* lookups;
                    long long* oob;
                    int get_p(uint32_t px_, uint32_t py_) const {
                        ++*lookups;
                        return get(static_cast<int>(px_ >> kFrac), static_cast<int>(py_ >> kFrac));
                    }
                    int get(int lx, int ly) const {
                        const int xx = lx + offx, yy = ly + offy;
                        if (xx < 0 || xx >= pitch || yy < 0 || yy >= rows) {
                            ++*oob;   // a read beyond the window: reported as a failure
                            return 0x80;
                        }
                        return w[static_cast<size_t>(yy) * pitch + xx];
                    }
                };
                const Checked acc{win[s].data(), st.bx - w.wx0, st.by - w.wy0, w.pitch, w.rows, &lk, &oob};
                if (s != k0_sector) {
                    k0 = dir_first_sample(acc, st);
                    k0_sector = s;
                }
                r = march_ray_dir(acc, offset_ray_start(st), dxf, dyf, M, rep, &replays, k0);
            } else {
                const DirGlobal acc{d.data() + static_cast<int64_t>(st.by) * sk.PW + st.bx, sk.PW, &lk};
                if (s != k0_sector) {
                    k0 = dir_first_sample(acc, st);
                    k0_sector = s;
                }
                r = march_ray_dir(acc, offset_ray_start(st), dxf, dyf, M, rep, &replays, k0);
            }
            out[j] = static_cast<uint8_t>(r);
            if (lookups_out) lookups_out[i * R + j] = static_cast<int32_t>(lk);
        }
    }
    if (oob) return -oob;
    return replays;
}

// shared-memory slot of coarse key k (exact_sum.cuh::coarse_slot) and the table size for nc keys
int emu_coarse_slot(int k) { return coarse_slot(k); }
int emu_coarse_slots(int nc) { return coarse_slots(nc); }

// Sequential-order sum / prefix sums of src[k] (/ div if use_div), following the kernels:
// k_tile_sums -> k_exact_chunks -> k_exact_walk -> k_exact_emit.  prefix_out nullable.
// Returns the number of opaque chunks (diagnostic).
long long emu_exact_scan(const double* src, long long N, int use_div, double div, double* total_out, double* prefix_out,
                         int force_last_one) {
    const int T = static_cast<int>((N + kTile - 1) / kTile);
    const long long C = static_cast<long long>(T) * kTileChunks;
    auto addend = [&](long long k) -> double {
        if (k >= N) return 0.0;
        return use_div ? nf_div(src[k], div) : src[k];
    };
    // k_tile_sums
    std::vector<double> tile_sum(T, 0.0);
    for (int t = 0; t < T; ++t) {
        double s = 0.0;
        for (long long k = static_cast<long long>(t) * kTile; k < std::min<long long>(N, static_cast<long long>(t + 1) * kTile); ++k) s += src[k];
        tile_sum[t] = s;
    }
    // k_exact_chunks
    std::vector<StepFn> chunk_fn(C), chunk_pre(C);
    std::vector<uint8_t> chunk_flag(C, 0);
    struct RF { StepFn f; int64_t reset; };
    auto rf_op = [](const RF& l, const RF& r) { return r.reset ? r : RF{fn_compose(l.f, r.f), l.reset}; };
    std::vector<RF> tile_elem(T);
    std::vector<int> tile_opq(T, 0);
    double pre_sum = 0.0;
    for (int t = 0; t < T; ++t) {
        double pre = pre_sum;
        if (use_div) pre = pre / div;
        double run = 0.0;
        RF inc{fn_identity(), 0};
        for (int c = 0; c < kTileChunks; ++c) {
            const long long cidx = static_cast<long long>(t) * kTileChunks + c;
            const long long base = cidx * kChunk;
            double v[kChunk];
            double csum = 0.0;
            for (int i = 0; i < kChunk; ++i) {
                v[i] = addend(base + i);
                csum += v[i];
            }
            const double s_in = pre + run, s_out = s_in + csum;
            run += csum;
            StepFn fn = fn_identity();
            int opaque = 0;
            if (base < N) {
                const long long cnt = std::min<long long>(N, base + kChunk);
                const int e = chunk_safe_binade(s_in, s_out, cnt);
                if (e < 0) {
                    fn = fn_opaque();
                    opaque = 1;
                } else {
                    fn = chunk_step_fn(v, kChunk, e);
                }
            }
            chunk_fn[cidx] = fn;
            const RF exc = inc;
            if (opaque) {
                chunk_pre[cidx] = exc.f;
                chunk_flag[cidx] = exc.reset ? 1 : 3;
                tile_opq[t]++;
            }
            inc = rf_op(inc, opaque ? RF{fn_identity(), 1} : RF{fn, 0});
        }
        tile_elem[t] = inc;
        pre_sum += tile_sum[t];
    }
    // k_exact_walk
    std::vector<long long> list_chunk;
    std::vector<StepFn> list_fn;
    {
        RF run{fn_identity(), 0};
        for (int t = 0; t < T; ++t) {
            if (tile_opq[t] > 0) {
                for (long long c = static_cast<long long>(t) * kTileChunks; c < static_cast<long long>(t + 1) * kTileChunks; ++c) {
                    if (chunk_flag[c] & 1) {
                        list_chunk.push_back(c);
                        list_fn.push_back((chunk_flag[c] & 2) ? fn_compose(run.f, chunk_pre[c]) : chunk_pre[c]);
                    }
                }
            }
            run = rf_op(run, tile_elem[t]);
        }
    }
    std::vector<double> anchors(list_chunk.size()), anchor_val(C, 0.0);
    {
        double V = 0.0;
        for (size_t r = 0; r < list_chunk.size(); ++r) {
            double v[kChunk];
            for (int i = 0; i < kChunk; ++i) v[i] = addend(list_chunk[r] * kChunk + i);
            const double vin = fn_apply(list_fn[r], V);
            V = chunk_seq_eval(v, kChunk, vin);
            anchors[r] = V;
            anchor_val[list_chunk[r]] = V;
        }
    }
    std::vector<double> tile_start(T);
    double total = 0.0;
    {
        RF run{fn_identity(), 0};
        int rank = 0;
        for (int t = 0; t < T; ++t) {
            tile_start[t] = fn_apply(run.f, run.reset ? anchors[rank - 1] : 0.0);
            rank += tile_opq[t];
            run = rf_op(run, tile_elem[t]);
        }
        total = fn_apply(run.f, run.reset ? anchors[rank - 1] : 0.0);
    }
    if (total_out) *total_out = total;
    // k_exact_emit
    if (prefix_out) {
        for (int t = 0; t < T; ++t) {
            ScanElem inc = se_abs(tile_start[t]);
            for (int c = 0; c < kTileChunks; ++c) {
                const long long cidx = static_cast<long long>(t) * kTileChunks + c;
                double s = bits_dbl(inc.a0);   // value before this chunk
                const ScanElem el = (chunk_flag[cidx] & 1) ? se_abs(anchor_val[cidx]) : se_fn(chunk_fn[cidx]);
                inc = se_combine(inc, el);
                const long long base = cidx * kChunk;
                for (int i = 0; i < kChunk && base + i < N; ++i) {
                    s = nf_add(s, addend(base + i));
                    prefix_out[base + i] = s;
                }
            }
        }
        if (force_last_one && N > 0) prefix_out[N - 1] = 1.0;
    }
    return static_cast<long long>(list_chunk.size());
}

}  // extern "C"
