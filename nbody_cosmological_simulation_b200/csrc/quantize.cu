// Quantisers of quantization.py as device code:
//   * pass 1 of the int modes: global max of d² over all pairs (quantization.py:112-113),
//   * the level table that collapses _grid_quantize_safe + pow/reciprocal/G into ≤ L entries,
//   * the free-standing tensor functions (_grid_quantize, _grid_quantize_safe, fp16/bf16 round trips).
// Every arithmetic step that feeds a round() is executed in the reference's op order with one IEEE
// rounding per op (no FMA contraction), so level indices are bit-exact given identical inputs.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "stream.cuh"
#include "lut.cuh"

namespace nb {

__device__ __forceinline__ float q_log(float x) { return logf(x); }
__device__ __forceinline__ double q_log(double x) { return log(x); }
__device__ __forceinline__ float q_exp(float x) { return expf(x); }
__device__ __forceinline__ double q_exp(double x) { return exp(x); }
__device__ __forceinline__ float q_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float q_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float q_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float q_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double q_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double q_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double q_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double q_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float q_rint(float a) { return rintf(a); }
__device__ __forceinline__ double q_rint(double a) { return rint(a); }
__device__ __forceinline__ float q_max(float a, float b) { return (a != a) ? a : fmaxf(a, b); }     // torch.clamp keeps NaN
__device__ __forceinline__ double q_max(double a, double b) { return (a != a) ? a : fmax(a, b); }

// normalized = (log(t) - lo) / (hi - lo) * (L-1)      quantization.py:119
template <typename T>
__device__ __forceinline__ T log_grid_normalized(T t, T lo, T span, T lm1) {
    return q_mul(q_div(q_sub(q_log(t), lo), span), lm1);
}
// exp(k/(L-1)*(hi-lo)+lo).clamp(min)                  quantization.py:121-127
template <typename T>
__device__ __forceinline__ T log_grid_value(T k, T lo, T span, T lm1, T min_val) {
    return q_max(q_exp(q_add(q_mul(q_div(k, lm1), span), lo)), min_val);
}

// ---- level table -------------------------------------------------------------------------------------
// Layout (float4 records): [0] header { lo2, scale, min_val, degenerate }, [1 + k] level k = { T_{k+1}, g_k, g_{k+1}, 0 },
// [1 + L] fast-lookup record (lut.cuh).  Launched as ONE block when L <= kLutFastMaxLevels so that the margin of
// the fast lookup can be reduced over all thresholds in the same launch.
struct LevelGrid {
    float t_lo, t_hi, lo, hi, span, lm1, lo2, scale;
    bool degenerate;
};

__device__ __forceinline__ LevelGrid level_grid(const int64_t* __restrict__ scalars, float eps2, float min_val, int levels) {
    LevelGrid g;
    g.t_lo = fmaxf(eps2, min_val);                                             // the diagonal: smallest clamped d²
    g.t_hi = fmaxf((float)double_from_key(scalars[NB_SLOT_MAX_D2]), min_val);
    g.lo = q_log(g.t_lo); g.hi = q_log(g.t_hi);
    g.span = q_sub(g.hi, g.lo);
    g.lm1 = (float)(levels - 1);
    g.degenerate = g.span < 1e-10f;                                            // quantization.py:115
    g.lo2 = log2f(g.t_lo);
    g.scale = g.degenerate ? 0.f : g.lm1 / (log2f(g.t_hi) - g.lo2);
    return g;
}

__global__ void __launch_bounds__(256) build_level_table_kernel(const int64_t* __restrict__ scalars, float eps2, float min_val,
                                                                float G, int levels, float4* __restrict__ table) {
    __shared__ int red[32];
    const LevelGrid gr = level_grid(scalars, eps2, min_val, levels);
    const float lo = gr.lo, span = gr.span, lm1 = gr.lm1;
    const bool degenerate = gr.degenerate;
    if (threadIdx.x == 0 && blockIdx.x == 0)
        table[0] = make_float4(gr.lo2, gr.scale, min_val, degenerate ? 1.f : 0.f);
    // fast lookup (lut.cuh): W = fma(lg2 t, scale, cm0) lands on a 2^-fb grid in [2P, 4P)
    const bool fast = levels <= kLutFastMaxLevels;                             // then gridDim.x == 1
    const int P = lut_pow2ceil(levels), fb = kLutFb;
    const float M = 3.0f * (float)kLutGridP;
    const float cm0 = q_add(fmaf(-gr.lo2, gr.scale, 0.5f), M);
    int need = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < levels; k += gridDim.x * blockDim.x) {
        auto factor = [&](int kk) -> float {
            // u_k -> G / u_k^1.5 as torch evaluates it: pow, reciprocal, mul by G   (simulation.py:97-101)
            const float u = degenerate ? gr.t_hi : log_grid_value((float)kk, lo, span, lm1, min_val);
            return q_mul(__frcp_rn(powf(u, 1.5f)), G);
        };
        float thr = __int_as_float(0x7f800000);                                // T_L = +inf
        if (!degenerate && k + 1 < levels) {
            // smallest float t in [t_lo, t_hi] whose level index is >= k+1 (index is monotone in t)
            unsigned lo_b = __float_as_uint(gr.t_lo), hi_b = __float_as_uint(gr.t_hi);
            const float target = (float)(k + 1);
            while (hi_b - lo_b > 1u) {
                const unsigned mid = lo_b + ((hi_b - lo_b) >> 1);
                const float idx = q_rint(log_grid_normalized(__uint_as_float(mid), lo, span, lm1));
                if (idx >= target) hi_b = mid; else lo_b = mid;
            }
            thr = __uint_as_float(hi_b);
            if (fast) {
                // distance of the fast lookup's fixed-point position from the ideal (k+1)·2^fb at the threshold and
                // just below it: W(T_j) must not be below the boundary by more than the margin, W(prev T_j) not above
                // — evaluated kLutD2Slack floats further out, because the force kernel feeds the lookup the FUSED d²
                // (2-3 roundings) while the level is defined on the reference's op-by-op d² (4-6 roundings)
                const int ideal = (k + 1) << fb;
                const int v_at = (int)(lut_wbits(__uint_as_float(hi_b - kLutD2Slack), gr.scale, cm0) - __float_as_uint(M));
                const int v_below = (int)(lut_wbits(__uint_as_float(hi_b - 1u + kLutD2Slack), gr.scale, cm0) - __float_as_uint(M));
                need = max(need, max(ideal - v_at, v_below - ideal + 1));
            }
        }
        table[1 + k] = make_float4(thr, factor(k), k + 1 < levels ? factor(k + 1) : 0.f, 0.f);
    }
    if (fast) {
        need = block_reduce(need, OpMax(), 0, red);
        if (threadIdx.x == 0) {
            const long long mg = (long long)need + 2;                          // + slack for lg2.approx non-monotonicity
            long long mgp = 2;
            while (mgp < mg) mgp <<= 1;
            const long long span_fb = 1ll << fb;
            const uint32_t zmask = 2 * mgp >= span_fb ? 0u : (uint32_t)((span_fb - 1) & ~(2 * mgp - 1));
            const int single_ok = mgp <= (span_fb >> 2) ? 1 : 0;
            const float cm = zmask ? q_add(cm0, (float)mgp * exp2f((float)-fb)) : cm0;
            table[1 + levels] = make_float4(cm, __uint_as_float(zmask), __int_as_float(fb | (single_ok << 8)), __int_as_float(P));
        }
    } else if (threadIdx.x == 0 && blockIdx.x == 0) {
        table[1 + levels] = make_float4(cm0, 0.f, __int_as_float(fb), __int_as_float(P));
    }
}

// Exhaustive proof of the fast lookup for one table: every float t in [t_lo, t_hi] is pushed through (a) the
// reference's op sequence round((log t - lo)/(hi - lo)·(L-1)), (b) the fast lookup, (c) the slow path.
// counters: [0] floats tested, [1] floats in doubt, [2] fast-path mismatches outside doubt, [3] slow-path mismatches.
struct TableThresholds {
    const float4* table; int levels;
    __device__ __forceinline__ float operator[](int j) const {
        return j <= 0 ? 0.f : (j <= levels ? table[j].x : __int_as_float(0x7f800000));
    }
};

__global__ void __launch_bounds__(256) lut_selfcheck_kernel(const int64_t* __restrict__ scalars, float eps2, float min_val, int levels,
                                                            const float4* __restrict__ table, unsigned long long* __restrict__ counters) {
    __shared__ unsigned long long red[32];
    const LevelGrid gr = level_grid(scalars, eps2, min_val, levels);
    const LutFast f = lut_fast_load(table, levels);
    const TableThresholds thr{table, levels};
    const uint32_t b0 = __float_as_uint(gr.t_lo), b1 = __float_as_uint(gr.t_hi);
    unsigned long long tested = 0, doubt = 0, bad_fast = 0, bad_slow = 0;
    if (!gr.degenerate) {
        for (uint64_t b = (uint64_t)b0 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b <= b1; b += (uint64_t)gridDim.x * blockDim.x) {
            const float t = __uint_as_float((uint32_t)b);
            const int k_ref = (int)q_rint(log_grid_normalized(t, gr.lo, gr.span, gr.lm1));
            ++tested;
            doubt += (lut_wbits(t, f.scale, f.cm) & f.zmask) == 0u;
            // the lookup sees a d² that may sit up to kLutD2Slack floats away from the exact-order d² = t
            for (int o = -(int)kLutD2Slack; o <= (int)kLutD2Slack; ++o) {
                const uint32_t wf = lut_wbits(__uint_as_float((uint32_t)((int64_t)b + o)), f.scale, f.cm);
                const int k_fast = (int)((wf >> f.fb) & (uint32_t)(f.p - 1));
                bad_fast += ((wf & f.zmask) != 0u && k_fast != k_ref);
                bad_slow += (lut_exact_level(t, wf, f, levels, thr) != k_ref);
            }
        }
    }
    tested = block_reduce(tested, OpAdd(), 0ull, red);
    doubt = block_reduce(doubt, OpAdd(), 0ull, red);
    bad_fast = block_reduce(bad_fast, OpAdd(), 0ull, red);
    bad_slow = block_reduce(bad_slow, OpAdd(), 0ull, red);
    if (threadIdx.x == 0) {
        atomicAdd(counters + 0, tested); atomicAdd(counters + 1, doubt);
        atomicAdd(counters + 2, bad_fast); atomicAdd(counters + 3, bad_slow);
    }
}

// ---- elementwise helpers -----------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) tensor_minmax_kernel(const T* __restrict__ in, int64_t count, int log_space, T clamp_min,
                                                            int64_t* __restrict__ scalars) {
    __shared__ long long red[32];
    long long kmin = kKeyHighest, kmax = kKeyLowest;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x) {
        T v = in[e];
        if (log_space) v = q_log(q_max(v, clamp_min));
        const long long k = key_from_double((double)v);
        kmin = k < kmin ? k : kmin;
        kmax = k > kmax ? k : kmax;
    }
    kmin = block_reduce(kmin, OpMin(), (long long)kKeyHighest, red);
    kmax = block_reduce(kmax, OpMax(), (long long)kKeyLowest, red);
    if (threadIdx.x == 0) {
        atomicMin(reinterpret_cast<long long*>(scalars + NB_SLOT_VAL_MIN), kmin);
        atomicMax(reinterpret_cast<long long*>(scalars + NB_SLOT_VAL_MAX), kmax);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) grid_quantize_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t count, int levels,
                                                            const int64_t* __restrict__ scalars) {
    const T lo = (T)double_from_key(scalars[NB_SLOT_VAL_MIN]);
    const T hi = (T)double_from_key(scalars[NB_SLOT_VAL_MAX]);
    const T span = q_sub(hi, lo), lm1 = (T)(levels - 1);
    const bool active = !(span < (T)1e-10);                                    // quantization.py:81
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x) {
        const T v = in[e];
        if (!active) { out[e] = v; continue; }
        const T k = q_rint(q_mul(q_div(q_sub(v, lo), span), lm1));            // :84-85
        out[e] = q_add(q_mul(q_div(k, lm1), span), lo);                        // :86
    }
}

template <typename T>
__global__ void __launch_bounds__(256) grid_quantize_safe_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                                 int32_t* __restrict__ index_out, int64_t count, int levels,
                                                                 T min_val, const int64_t* __restrict__ scalars) {
    const T lo = (T)double_from_key(scalars[NB_SLOT_VAL_MIN]);
    const T hi = (T)double_from_key(scalars[NB_SLOT_VAL_MAX]);
    const T span = q_sub(hi, lo), lm1 = (T)(levels - 1);
    const bool active = !(span < (T)1e-10);                                    // quantization.py:115
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x) {
        const T t = q_max(in[e], min_val);                                     // :106
        if (!active) { out[e] = t; if (index_out) index_out[e] = 0; continue; }
        const T k = q_rint(log_grid_normalized(t, lo, span, lm1));             // :119-120
        out[e] = log_grid_value(k, lo, span, lm1, min_val);                    // :121-127
        if (index_out) index_out[e] = (int32_t)k;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) snap_index_kernel(const T* __restrict__ normalized, int32_t* __restrict__ index_out, int64_t count) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x)
        index_out[e] = (int32_t)q_rint(normalized[e]);
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256) round_trip_kernel(const T* __restrict__ in, float* __restrict__ out, int64_t count) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x) {
        if (MODE == NB_MODE_FLOAT16) {
            // torch casts double -> half directly (one rounding)
            const __half h = sizeof(T) == 8 ? __double2half((double)in[e]) : __float2half_rn((float)in[e]);
            out[e] = __half2float(h);
        } else {
            const __nv_bfloat16 h = sizeof(T) == 8 ? __double2bfloat16((double)in[e]) : __float2bfloat16_rn((float)in[e]);
            out[e] = __bfloat162float(h);
        }
    }
}

__global__ void reset_scalars_kernel(int64_t* scalars) {
    if (threadIdx.x == 0) {
        scalars[NB_SLOT_MAX_D2] = kKeyLowest;
        scalars[NB_SLOT_ACC_MIN] = kKeyHighest;
        scalars[NB_SLOT_ACC_MAX] = kKeyLowest;
        scalars[NB_SLOT_VAL_MIN] = kKeyHighest;
        scalars[NB_SLOT_VAL_MAX] = kKeyLowest;
        scalars[NB_SLOT_RADIUS_MAX] = kKeyLowest;
        scalars[6] = 0;
        scalars[7] = 0;
    }
}

inline int ew_grid(int64_t count) {
    int64_t b = (count + 255) / 256;
    if (b > (int64_t)kNumSMsB200 * 16) b = (int64_t)kNumSMsB200 * 16;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace nb

using namespace nb;

extern "C" int nb_reset_scalars(int64_t* scalars, void* stream) {
    if (!scalars) return NB_ERR_INVALID_ARGUMENT;
    reset_scalars_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(scalars);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int64_t nb_level_table_bytes(int levels) { return levels < 2 ? 0 : (int64_t)(levels + 2) * 16; }

extern "C" int nb_build_level_table(const int64_t* scalars, int dtype, double eps_sq, double min_dist_sq, double G, int levels,
                                    void* table, void* stream) {
    if (!scalars || !table || levels < 2) return NB_ERR_INVALID_ARGUMENT;
    if (dtype != NB_F32) return NB_ERR_UNSUPPORTED;
    const int blocks = levels <= kLutFastMaxLevels ? 1 : (levels + 255) / 256;
    build_level_table_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(scalars, (float)eps_sq, (float)min_dist_sq, (float)G, levels,
                                                                        (float4*)table);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_lut_selfcheck(const int64_t* scalars, double eps_sq, double min_dist_sq, int levels, const void* table,
                                uint64_t* counters, void* stream) {
    if (!scalars || !table || !counters || levels < 2) return NB_ERR_INVALID_ARGUMENT;
    if (levels > kLutFastMaxLevels) return NB_ERR_UNSUPPORTED;
    lut_selfcheck_kernel<<<kNumSMsB200 * 8, 256, 0, (cudaStream_t)stream>>>(scalars, (float)eps_sq, (float)min_dist_sq, levels,
                                                                            (const float4*)table, (unsigned long long*)counters);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_tensor_minmax(const void* in, int64_t count, int dtype, int log_space, double clamp_min, int64_t* scalars,
                                void* stream) {
    if (!in || !scalars || count <= 0) return NB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NB_F32) tensor_minmax_kernel<float><<<ew_grid(count), 256, 0, st>>>((const float*)in, count, log_space, (float)clamp_min, scalars);
    else if (dtype == NB_F64) tensor_minmax_kernel<double><<<ew_grid(count), 256, 0, st>>>((const double*)in, count, log_space, clamp_min, scalars);
    else return NB_ERR_INVALID_ARGUMENT;
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_grid_quantize(const void* in, void* out, int64_t count, int dtype, int levels, const int64_t* scalars, void* stream) {
    if (!in || !out || !scalars || count <= 0 || levels < 2) return NB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NB_F32) grid_quantize_kernel<float><<<ew_grid(count), 256, 0, st>>>((const float*)in, (float*)out, count, levels, scalars);
    else if (dtype == NB_F64) grid_quantize_kernel<double><<<ew_grid(count), 256, 0, st>>>((const double*)in, (double*)out, count, levels, scalars);
    else return NB_ERR_INVALID_ARGUMENT;
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_grid_quantize_safe(const void* in, void* out, int32_t* index_out, int64_t count, int dtype, int levels,
                                     double min_val, const int64_t* scalars, void* stream) {
    if (!in || !out || !scalars || count <= 0 || levels < 2) return NB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NB_F32)
        grid_quantize_safe_kernel<float><<<ew_grid(count), 256, 0, st>>>((const float*)in, (float*)out, index_out, count, levels, (float)min_val, scalars);
    else if (dtype == NB_F64)
        grid_quantize_safe_kernel<double><<<ew_grid(count), 256, 0, st>>>((const double*)in, (double*)out, index_out, count, levels, min_val, scalars);
    else return NB_ERR_INVALID_ARGUMENT;
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_snap_index(const void* normalized, int32_t* index_out, int64_t count, int dtype, void* stream) {
    if (!normalized || !index_out || count <= 0) return NB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NB_F32) snap_index_kernel<float><<<ew_grid(count), 256, 0, st>>>((const float*)normalized, index_out, count);
    else if (dtype == NB_F64) snap_index_kernel<double><<<ew_grid(count), 256, 0, st>>>((const double*)normalized, index_out, count);
    else return NB_ERR_INVALID_ARGUMENT;
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_round_trip(const void* in, void* out, int64_t count, int dtype, int mode, void* stream) {
    if (!in || !out || count <= 0) return NB_ERR_INVALID_ARGUMENT;
    if (mode != NB_MODE_FLOAT16 && mode != NB_MODE_BFLOAT16) return NB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    const int g = ew_grid(count);
    if (dtype == NB_F32 && mode == NB_MODE_FLOAT16) round_trip_kernel<float, NB_MODE_FLOAT16><<<g, 256, 0, st>>>((const float*)in, (float*)out, count);
    else if (dtype == NB_F32) round_trip_kernel<float, NB_MODE_BFLOAT16><<<g, 256, 0, st>>>((const float*)in, (float*)out, count);
    else if (dtype == NB_F64 && mode == NB_MODE_FLOAT16) round_trip_kernel<double, NB_MODE_FLOAT16><<<g, 256, 0, st>>>((const double*)in, (float*)out, count);
    else if (dtype == NB_F64) round_trip_kernel<double, NB_MODE_BFLOAT16><<<g, 256, 0, st>>>((const double*)in, (float*)out, count);
    else return NB_ERR_INVALID_ARGUMENT;
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}
