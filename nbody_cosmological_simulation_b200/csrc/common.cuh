// Shared device/host helpers for libnbody_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/nbody_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libnbody_b200 targets sm_100a (B200) only"
#endif

namespace nb {

constexpr int kNumSMsB200 = 148;
constexpr int kChunkUnits = NB_CHUNK_UNITS;           // 128 units / chunk
constexpr int kChunkABytes = kChunkUnits * 16;        // {x0,x1,y0,y1} (fp32 pair) or {x,y} (fp64)

// Padding records (fill of the last chunk / whole padding chunks): mass 0 and a far-away position, chosen so
// that d2^-3/2 underflows to exactly 0 in the state dtype: a pad contributes 0 force even in the kernels that
// do not read per-source masses (uniform-mass fast path), and 0 potential because its mass is 0.
constexpr float kPadCoordF32 = 1.0e18f;                // d² ~ 3e36 (finite); r³ ~ 2e-55 -> 0
constexpr double kPadCoordF64 = 1.0e150;               // d² ~ 3e300 (finite); r³ ~ 2e-451 -> 0
constexpr float kPadDetectF32 = 2.5e17f;               // |x| above this marks a pad (max-d² pass skips it)

__host__ __device__ inline int chunk_b_bytes(int dim) { return kChunkUnits * (dim == 3 ? 16 : 8); }
__host__ __device__ inline int chunk_bytes(int dim) { return kChunkABytes + chunk_b_bytes(dim); }
__host__ __device__ inline int chunk_sources(int dtype) { return dtype == NB_F32 ? 2 * kChunkUnits : kChunkUnits; }

inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? NB_OK : NB_ERR_CUDA_BASE + (int)e; }
#define NB_CUDA_LAUNCH_CHECK() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return nb::cuda_status(e__); } while (0)

// ---- order-preserving 64-bit keys of doubles (cross-rank MIN/MAX on raw int64 works) -------------
// signed compare of the key == IEEE compare of the value; NaN (positive payload) sorts above +inf so a
// NaN anywhere makes the max NaN, which reproduces torch.max NaN propagation well enough for the
// `max - min` span (quantization.py:78-81).
__host__ __device__ inline int64_t key_from_double(double v) {
#ifdef __CUDA_ARCH__
    long long b = __double_as_longlong(v);
#else
    long long b; memcpy(&b, &v, 8);
#endif
    return (b < 0) ? (long long)(b ^ 0x7fffffffffffffffLL) : b;
}
__host__ __device__ inline double double_from_key(int64_t k) {
    long long b = (k < 0) ? (long long)(k ^ 0x7fffffffffffffffLL) : (long long)k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double(b);
#else
    double v; memcpy(&v, &b, 8); return v;
#endif
}
constexpr int64_t kKeyLowest = INT64_MIN;    // below -inf
constexpr int64_t kKeyHighest = INT64_MAX;   // above every number (NaN payload region)

#ifdef __CUDACC__
// ---- mbarrier / TMA bulk-copy PTX (sm_90+; on sm_100a SASS shows SYNCS.* and UBLKCP) -------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on `bar` (16 B aligned, size % 16 == 0)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));   // not volatile: read-only tables, free to schedule
    return v;
}

// ---- packed fp32x2 math (sm_100 FFMA2 / FADD2 / FMUL2) ------------------------------------------------
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); // one MUFU.RSQ (non-ftz adds FSETP + 2 predicated FMUL per call)
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // one MUFU.LG2 (non-ftz adds FSETP + FMUL + FADD per call)
    return y;
}
__device__ __forceinline__ double rsqrt64h(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); // MUFU.RSQ64H: ~20-bit seed from the high word
    return y;
}

// ---- warp / block reductions ---------------------------------------------------------------------
template <typename T, typename Op>
__device__ __forceinline__ T warp_reduce(T v, Op op) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
struct OpAdd { template <typename T> __device__ T operator()(T a, T b) const { return a + b; } };
struct OpMin { template <typename T> __device__ T operator()(T a, T b) const { return a < b ? a : b; } };
struct OpMax { template <typename T> __device__ T operator()(T a, T b) const { return a > b ? a : b; } };

// Block-wide reduction; result valid in thread 0.  `scratch` holds >= 32 elements of T.
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T identity, T* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    v = warp_reduce(v, op);
    __syncthreads();                 // scratch may be reused between calls
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nwarps ? scratch[lane] : identity;
        v = warp_reduce(v, op);
    }
    return v;
}
#endif  // __CUDACC__

}  // namespace nb
