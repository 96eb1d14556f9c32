// Fast level lookup for the log-grid d² quantiser (_grid_quantize_safe, quantization.py:91-127) — shared by the
// table builder (quantize.cu), the force kernel (accel.cu) and the exhaustive self-check (quantize.cu).
//
// The exact level of a clamped d² value t is k(t) = #{ j in 1..L-1 : T_j <= t } with T_j the exact thresholds
// the table builder finds by bisection over float bit patterns (reference op order).  The force kernel does not
// evaluate that per pair; it computes ONE fused multiply-add
//     W(t) = bits( fma( lg2.approx(t), scale, cm ) ) - bits(M),      M = 768,
// whose result lands in the binade [512, 1024) where a float has FB = 14 fractional bits: W is the
// fixed-point value (position of t on the level axis + ½)·2^FB + mgp, rounded ONCE to the 2^-FB grid, so
//     k_fast = (W >> FB) & (P-1),   P = pow2ceil(L) <= 256.
// The builder measures, at every threshold T_j and at its predecessor float, how far W is from the ideal
// j·2^FB, takes the worst case + 2 as the margin, rounds it up to a power of two mgp and folds it into cm.
// The margin also covers the force kernel evaluating W on the FUSED d² (fma chain, 2-3 roundings) instead of the
// reference's op-by-op d² (4-6 roundings): the two differ by < 7 ulp, the builder measures 16 floats further out.
// Then (W mod 2^FB) >= 2·mgp proves k_fast == k(t) (W is monotone in t); the few t closer than that to a level
// boundary ("doubt", ~5e-4 of all pairs at L = 256) take a slow path that compares against T_j itself.
// nb_lut_selfcheck proves the claim for EVERY float in [t_lo, t_hi] of a given table.
#pragma once
#include "common.cuh"

namespace nb {

constexpr int kLutFastMaxLevels = 256;          // P·128 B of shared memory for the lane-replicated factor table
constexpr int kLutFb = 14;                      // fractional bits of W: the grid of the binade [512, 1024), whatever L is
constexpr uint32_t kLutD2Slack = 16;            // floats between the fused d² of the fast path and the exact-order d² (< 7 ulp, see accel.cu)
constexpr int kLutGridP = 256;                  // W = (level position + ½)·2^14 + bits(M), M = 3·256

struct LutFast {
    float scale;        // (L-1) / (log2 t_hi - log2 t_lo)     (0 when the grid is degenerate)
    float cm;           // ½ - lo2·scale + M + mgp·2^-FB, rounded to the 2^-FB grid
    uint32_t zmask;     // (2^FB - 1) & ~(2·mgp - 1): (W & zmask) == 0  <=>  doubt;  0 => every t takes the slow path
    int fb;             // fractional bits of W
    int p;              // pow2ceil(L)
    int single_ok;      // slow path may decide with ONE threshold compare (mgp < 2^(FB-2)); else binary search
};

__host__ __device__ inline int lut_pow2ceil(int v) { int p = 2; while (p < v) p <<= 1; return p; }
__host__ __device__ inline int lut_log2(int p) { int l = 0; while ((1 << l) < p) ++l; return l; }

// W(t) + bits(M): the raw float bits the kernels work on
__device__ __forceinline__ uint32_t lut_wbits(float t, float scale, float cm) {
    return __float_as_uint(fmaf(lg2_approx(t), scale, cm));
}

// last 16-byte record of the level table: { cm, zmask, fb | single_ok << 8, p }
__device__ __forceinline__ LutFast lut_fast_load(const float4* table, int levels) {
    const float4 h0 = table[0], h1 = table[1 + levels];
    LutFast f;
    f.scale = h0.y;
    f.cm = h1.x;
    f.zmask = __float_as_uint(h1.y);
    const int w = __float_as_int(h1.z);
    f.fb = w & 0xff;
    f.single_ok = (w >> 8) & 1;
    f.p = __float_as_int(h1.w);
    return f;
}

// Exact level from the thresholds (slow path; `thr[j]` = T_j for j = 1..L-1, thr[0] <= every t, thr[j >= L] = +inf)
template <typename ThrPtr>
__device__ __noinline__ int lut_search_level(float t, int levels, ThrPtr thr) {
    int lo = 0, hi = levels;                                           // invariant: T_lo <= t < T_hi
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (t >= thr[mid]) lo = mid; else hi = mid;
    }
    return lo;
}
template <typename ThrPtr>
__device__ __forceinline__ int lut_exact_level(float t, uint32_t wbits, const LutFast& f, int levels, ThrPtr thr) {
    int k;
    if (f.single_ok) {
        const int j = (int)((wbits >> f.fb) & (uint32_t)(f.p - 1));   // the only boundary within the margin of W
        k = t >= thr[j] ? j : j - 1;
    } else {
        k = lut_search_level(t, levels, thr);                          // pathological grids only (levels denser than floats)
    }
    return min(max(k, 0), levels - 1);
}

}  // namespace nb
