// csrc/map_prep.h -- host-side preprocessing of the occupancy grid into the skip map the
// ray-march kernels read.  Replaces nothing in the reference one-to-one: the reference marches
// the raw int8 grid cell by cell (src/particle_filter.cpp:611-650); this builds an equivalent
// structure that lets the GPU skip samples that provably cannot be hits while keeping the
// reference's sample lattice and hit rule.
//
// P-lattice ("padded floor lattice").  The reference maps a sample to a cell with
// static_cast<int>((c - origin) / res), i.e. truncation toward zero, so quotients in (-1, 0)
// land in cell 0.  With q the real quotient, P-cell index p = floor(q) + PADL:
//     floor(q) == -1      -> same content as reference cell 0   (trunc quirk)
//     floor(q) in [0, W)  -> reference cell floor(q)
//     anything else       -> out of bounds == blocked (reference returns on OOB, :632-636)
// A cell is "blocked" iff OOB or occupancy > 50 (:642); unknown (-1) is transparent.
//
// Skip code per P-cell (one byte, v8):
//     0        blocked
//     1        not blocked but an 8-neighbour is blocked ("near"; the sample's cell must
//              be known exactly)
//     2..255   1 + adv, adv = ceil(d - 1e-3) >= 1 where d is the Euclidean gap between this
//              cell's square and the nearest blocked cell's square: a sample inside this
//              cell is not a hit, and neither are the next adv-1 lattice samples (unit step).
// v4 is the same code clamped to 15, two cells per byte (low nibble = even column), used for
// the shared-memory window.
#pragma once
#include <cstdint>
#include <vector>

#include "dirmap.cuh"

namespace mclb200 {

constexpr int kPadL = 8;   // P-cells left of / below reference cell 0 (incl. the trunc-duplicate)
constexpr int kPadR = 8;   // minimum P-cells right of / above the last reference cell

struct SkipMap {
    int W = 0, H = 0;      // reference grid
    int PW = 0, PH = 0;    // P-grid (PW is a multiple of 32)
    std::vector<uint8_t> v8;   // PH * PW
    std::vector<uint8_t> v4;   // PH * PW/2
    std::vector<int32_t> free_cells;  // row*W+col of cells == 0, row-major order (:411-421)
};

// Build the skip map.  Returns false if the grid is too large for the fixed-point march.
// (CPU twin of the device build in map_kernels.cuh: the product calls skip_map_layout() and builds the
// codes on the GPU; tests/emu builds them here.)
bool build_skip_map(const int8_t* data, int W, int H, SkipMap& out);

// Dimensions of the P-grid and the free-cell list only (v8 / v4 stay empty).
bool skip_map_layout(const int8_t* data, int W, int H, SkipMap& out);

// Exact squared Euclidean distance transform to the set {mask != 0} (Felzenszwalb-
// Huttenlocher lower envelopes); out[i] = squared distance in cells, big if no seed.
void edt_squared(const std::vector<uint8_t>& mask, int W, int H, std::vector<int64_t>& out);

// Euclidean gap (cells) between every P-cell's square and the nearest blocked cell's square
// (0 for blocked cells and their 8-neighbours): the distance field the cone tracing of
// dirmap.cuh queries.
void build_gap_map(const SkipMap& sk, std::vector<float>& gap);

// Geometry of the kDirSectors heading sectors for rays of at most M steps.
void make_dir_sectors(int M, DirSector* out);

}  // namespace mclb200
