// Quantisers of quantization.py as device code:
//   * pass 1 of the int modes: global max of d² over all pairs (quantization.py:112-113),
//   * the level table that collapses _grid_quantize_safe + pow/reciprocal/G into ≤ L entries,
//   * the free-standing tensor functions (_grid_quantize, _grid_quantize_safe, fp16/bf16 round trips).
// Every arithmetic step that feeds a round() is executed in the reference's op order with one IEEE
// rounding per op (no FMA contraction), so level indices are bit-exact given identical inputs.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "stream.cuh"

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
struct LevelHeader { float lo2, scale, min_val, degenerate; };

__global__ void __launch_bounds__(256) build_level_table_kernel(const int64_t* __restrict__ scalars, float eps2, float min_val,
                                                                float G, int levels, float4* __restrict__ table) {
    const float t_lo = fmaxf(eps2, min_val);                                   // the diagonal: smallest clamped d²
    const float t_hi = fmaxf((float)double_from_key(scalars[NB_SLOT_MAX_D2]), min_val);
    const float lo = q_log(t_lo), hi = q_log(t_hi);
    const float span = q_sub(hi, lo);
    const float lm1 = (float)(levels - 1);
    const bool degenerate = span < 1e-10f;                                     // quantization.py:115
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const float lo2 = log2f(t_lo), hi2 = log2f(t_hi);
        LevelHeader h;
        h.lo2 = lo2;
        h.scale = degenerate ? 0.f : lm1 / (hi2 - lo2);
        h.min_val = min_val;
        h.degenerate = degenerate ? 1.f : 0.f;
        table[0] = make_float4(h.lo2, h.scale, h.min_val, h.degenerate);
    }
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < levels; k += gridDim.x * blockDim.x) {
        auto factor = [&](int kk) -> float {
            // u_k -> G / u_k^1.5 as torch evaluates it: pow, reciprocal, mul by G   (simulation.py:97-101)
            const float u = degenerate ? t_hi : log_grid_value((float)kk, lo, span, lm1, min_val);
            return q_mul(__frcp_rn(powf(u, 1.5f)), G);
        };
        float thr = __int_as_float(0x7f800000);                                // T_L = +inf
        if (!degenerate && k + 1 < levels) {
            // smallest float t in [t_lo, t_hi] whose level index is >= k+1 (index is monotone in t)
            unsigned lo_b = __float_as_uint(t_lo), hi_b = __float_as_uint(t_hi);
            const float target = (float)(k + 1);
            while (hi_b - lo_b > 1u) {
                const unsigned mid = lo_b + ((hi_b - lo_b) >> 1);
                const float idx = q_rint(log_grid_normalized(__uint_as_float(mid), lo, span, lm1));
                if (idx >= target) hi_b = mid; else lo_b = mid;
            }
            thr = __uint_as_float(hi_b);
        }
        table[1 + k] = make_float4(thr, factor(k), k + 1 < levels ? factor(k + 1) : 0.f, 0.f);
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

extern "C" int64_t nb_level_table_bytes(int levels) { return levels < 2 ? 0 : (int64_t)(levels + 1) * 16; }

extern "C" int nb_build_level_table(const int64_t* scalars, int dtype, double eps_sq, double min_dist_sq, double G, int levels,
                                    void* table, void* stream) {
    if (!scalars || !table || levels < 2) return NB_ERR_INVALID_ARGUMENT;
    if (dtype != NB_F32) return NB_ERR_UNSUPPORTED;
    const int blocks = (levels + 255) / 256;
    build_level_table_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(scalars, (float)eps_sq, (float)min_dist_sq, (float)G, levels,
                                                                        (float4*)table);
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
