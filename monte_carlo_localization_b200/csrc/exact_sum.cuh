// csrc/exact_sum.cuh -- sequential-order FP64 sums and prefix sums, computed in parallel.
//
// Why.  The reference resamples with std::discrete_distribution (src/particle_filter.cpp:658):
// libstdc++ sums the weights with a sequential std::accumulate, divides each weight by that
// sum, and builds the CDF with a sequential std::partial_sum (bits/random.tcc:2657-2678).
// Step 4 of MCL() normalises with another sequential accumulate (:679).  FP64 addition is not
// associative, so a tree reduction / parallel scan produces values that differ in the last
// bits, and a draw u that falls between the two roundings of a CDF edge picks a different
// particle.  "Resample indices bit-exact" therefore needs the *sequentially rounded* values.
//
// How.  Let s be the running sum (a positive double) and a >= 0 the next addend.  As long as
// fl(s + a) stays in the binade of s, fl(s + a) = s + round_to_ulp(a), where the only
// dependence on s is the parity of its last mantissa bit (round-half-to-even).  Hence within
// one binade the map "s_in -> s_out" of ANY run of consecutive addends is
//        bits(s_out) = bits(s_in) + A[bits(s_in) & 1]
// for a pair of integers (A[0], A[1]), and such maps compose associatively:
//        (B o A)[p] = A[p] + B[(p + A[p]) & 1].
// So: (1) a plain parallel scan gives every running sum to ~1e-10, which fixes its binade
// except within a rigorous error band of a power of two; (2) every kChunk-addend chunk that is
// safely inside one binade is summarised by its pair (two short sequential chains from the
// even and the odd bottom of the binade); (3) pairs are combined with a parallel scan;
// (4) the few chunks that may cross a binade ("opaque", a few dozen per million addends) are
// evaluated sequentially from their now-exact input.  The result is bit-identical to the
// sequential loop.  This header holds the algebra; the kernels are in mcl_b200.cu.
#pragma once
#include <stdint.h>

#include "march.cuh"  // MCL_HD, nf_add

namespace mclb200 {

constexpr int kChunk = 8;               // addends per chunk (one thread)
constexpr int kTileChunks = 512;        // chunks per tile (one CTA)
constexpr int kTile = kChunk * kTileChunks;   // 4096 addends per tile

// Step map of a run of addends inside one binade.  opaque => not representable; a[] unused.
struct StepFn {
    int64_t a0, a1;
};
constexpr int64_t kOpaqueMark = INT64_MIN;

MCL_HD bool fn_is_opaque(const StepFn& f) { return f.a0 == kOpaqueMark; }
MCL_HD StepFn fn_identity() { return StepFn{0, 0}; }
MCL_HD StepFn fn_opaque() { return StepFn{kOpaqueMark, 0}; }
// first f, then g
MCL_HD StepFn fn_compose(const StepFn& f, const StepFn& g) {
    StepFn r;
    r.a0 = f.a0 + ((f.a0 & 1) ? g.a1 : g.a0);
    r.a1 = f.a1 + (((1 + f.a1) & 1) ? g.a1 : g.a0);
    return r;
}

MCL_HD int64_t dbl_bits(double v) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(v);
#else
    union { double d; int64_t i; } u;
    u.d = v;
    return u.i;
#endif
}
MCL_HD double bits_dbl(int64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    union { double d; int64_t i; } u;
    u.i = b;
    return u.d;
#endif
}
MCL_HD double fn_apply(const StepFn& f, double s) {
    const int64_t b = dbl_bits(s);
    return bits_dbl(b + ((b & 1) ? f.a1 : f.a0));
}

// Scan element: either an absolute value (the exact running sum after an opaque chunk or at
// a tile start) or a step map to be applied to whatever precedes it.
struct ScanElem {
    int64_t a0, a1;   // if is_abs: a0 = bits of the value
    int is_abs;
};
MCL_HD ScanElem se_fn(const StepFn& f) { return ScanElem{f.a0, f.a1, 0}; }
MCL_HD ScanElem se_abs(double v) { return ScanElem{dbl_bits(v), 0, 1}; }
// left then right
MCL_HD ScanElem se_combine(const ScanElem& l, const ScanElem& r) {
    if (r.is_abs) return r;
    if (l.is_abs) {
        const int64_t b = l.a0;
        return ScanElem{b + ((b & 1) ? r.a1 : r.a0), 0, 1};
    }
    const StepFn c = fn_compose(StepFn{l.a0, l.a1}, StepFn{r.a0, r.a1});
    return ScanElem{c.a0, c.a1, 0};
}

// Rigorous band within which the sequentially rounded running sum after `count` addends lies
// around an approximately computed one (both are within count*eps*sum of the real sum; the
// +512 absorbs the reordering/rounding of the approximate scan itself).
MCL_HD double sum_band(double approx, int64_t count) {
    return approx * (static_cast<double>(count + 512) * 2.5e-16);
}

MCL_HD int dbl_exponent(double v) { return static_cast<int>((dbl_bits(v) >> 52) & 0x7ff); }

// Is a chunk whose approximate running sum goes s_in -> s_out (after `count_out` addends in
// total) certainly inside one binade?  Returns the biased exponent or -1.
MCL_HD int chunk_safe_binade(double s_in, double s_out, int64_t count_out) {
    const double band = sum_band(s_out, count_out);
    const double lo = s_in - band, hi = s_out + band;
    if (!(lo > 0.0)) return -1;
    const int e_lo = dbl_exponent(lo), e_hi = dbl_exponent(hi);
    if (e_lo != e_hi || e_lo == 0 || e_lo == 0x7ff) return -1;
    return e_lo;
}

// Step map of n addends for running sums in the binade with biased exponent e.
MCL_HD StepFn chunk_step_fn(const double* a, int n, int e) {
    const int64_t b0 = static_cast<int64_t>(e) << 52;   // 2^(e-1023), even mantissa
    double y0 = bits_dbl(b0), y1 = bits_dbl(b0 + 1);
    for (int i = 0; i < n; ++i) {
        y0 = nf_add(y0, a[i]);
        y1 = nf_add(y1, a[i]);
    }
    return StepFn{dbl_bits(y0) - b0, dbl_bits(y1) - (b0 + 1)};
}

// Sequential evaluation of n addends from an exact input.
MCL_HD double chunk_seq_eval(const double* a, int n, double s) {
    for (int i = 0; i < n; ++i) s = nf_add(s, a[i]);
    return s;
}

// Position of key k of the resampling search's coarse level in shared memory (kernels.cuh::cdf_lower_bound).  A binary
// search over a power-of-two table probes multiples of large powers of two first -- ALL in bank 0: with the keys stored
// at their index the 2^s candidates of step s <= 8 were 2^s-way bank conflicts (ncu: 216 wavefronts per warp and search,
// 88 % of them replays).  One pad word per 32 keys and another per 1024 spread every step's candidates over the banks
// (~36 wavefronts per warp and search).  tests/test_emu_logic.py checks the spread level by level.
MCL_HD int coarse_slot(int k) { return k + (k >> 5) + (k >> 10); }
MCL_HD int coarse_slots(int nc) { return nc > 0 ? coarse_slot(nc - 1) + 1 : 0; }

}  // namespace mclb200
