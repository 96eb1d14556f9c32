// The O(N) remainder of metrics.collect_metrics, as device code without sorting:
//   compute_galaxy_radius       metrics.py:81-95    k-th smallest radius by MSB-first radix select (exact)
//   compute_velocity_dispersion metrics.py:148-156  unbiased std of |v| from fp64 power sums
// Radii / speeds are formed with the reference's rounding sequence (separately rounded squares and adds, IEEE sqrt),
// so the selected order statistic is bit-identical to `torch.sort(radii)[0][k]`.
#include "common.cuh"

namespace nb {

template <typename T> struct KeyOf;
template <> struct KeyOf<float> { using type = unsigned; static constexpr int kBits = 32; };
template <> struct KeyOf<double> { using type = unsigned long long; static constexpr int kBits = 64; };

__device__ __forceinline__ unsigned bits_of(float v) { return __float_as_uint(v); }
__device__ __forceinline__ unsigned long long bits_of(double v) { return (unsigned long long)__double_as_longlong(v); }
__device__ __forceinline__ float from_bits(unsigned b) { return __uint_as_float(b); }
__device__ __forceinline__ double from_bits(unsigned long long b) { return __longlong_as_double((long long)b); }

template <typename T> __device__ __forceinline__ T m_mul(T a, T b);
template <> __device__ __forceinline__ float m_mul(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double m_mul(double a, double b) { return __dmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T m_add(T a, T b);
template <> __device__ __forceinline__ float m_add(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double m_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float m_sqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double m_sqrt(double a) { return __dsqrt_rn(a); }

// sqrt((v ** 2).sum(dim=-1)) of row i — ((x²+y²)+z²), each op rounded once (metrics.py:92,155)
template <typename T, int DIM>
__device__ __forceinline__ T row_norm(const T* __restrict__ a, int64_t i) {
    T s = m_add(m_mul(a[i * DIM], a[i * DIM]), m_mul(a[i * DIM + 1], a[i * DIM + 1]));
    if (DIM == 3) s = m_add(s, m_mul(a[i * DIM + 2], a[i * DIM + 2]));
    return m_sqrt(s);
}

// state block: [0] = prefix key found so far, [1] = remaining rank k inside the current prefix, [2..257] = histogram
template <typename T, int DIM>
__global__ void __launch_bounds__(256) radix_hist_kernel(const T* __restrict__ pos, int64_t n, int shift,
                                                         unsigned long long* __restrict__ state) {
    using K = typename KeyOf<T>::type;
    __shared__ unsigned hist[256];
    hist[threadIdx.x] = 0u;
    __syncthreads();
    const K prefix = (K)state[0];
    const K high_mask = (shift + 8 >= KeyOf<T>::kBits) ? (K)0 : (K)(~(K)0 << (shift + 8));
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const K key = bits_of(row_norm<T, DIM>(pos, i));            // radii are >= 0: bit order == value order
        if ((key & high_mask) == (prefix & high_mask)) atomicAdd(&hist[(unsigned)((key >> shift) & 0xff)], 1u);
    }
    __syncthreads();
    if (hist[threadIdx.x]) atomicAdd(&state[2 + threadIdx.x], (unsigned long long)hist[threadIdx.x]);
}

template <typename T>
__global__ void radix_pick_kernel(int shift, unsigned long long* __restrict__ state, T* __restrict__ out, int last) {
    using K = typename KeyOf<T>::type;
    if (threadIdx.x == 0) {
        unsigned long long k = state[1];
        int digit = 255;
        for (int d = 0; d < 256; ++d) {
            const unsigned long long c = state[2 + d];
            if (k < c) { digit = d; break; }
            k -= c;
        }
        state[0] = (unsigned long long)((K)state[0] | ((K)digit << shift));
        state[1] = k;
        for (int d = 0; d < 256; ++d) state[2 + d] = 0ull;
        if (last) out[0] = from_bits((K)state[0]);
    }
}

__global__ void radix_init_kernel(unsigned long long* state, unsigned long long k) {
    for (int i = threadIdx.x; i < 258; i += blockDim.x) state[i] = i == 1 ? k : 0ull;
}

template <typename T, int DIM>
__global__ void __launch_bounds__(256) speed_moments_kernel(const T* __restrict__ vel, int64_t n, double* __restrict__ partials) {
    __shared__ double red[32];
    double s1 = 0.0, s2 = 0.0;
    // power sums about a pivot (the first star's speed): the variance is shift-invariant and Σ(s−p)² − (Σ(s−p))²/n
    // cancels far less than the raw moments when the spread is small against the mean
    const double pivot = (double)row_norm<T, DIM>(vel, 0);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double s = (double)row_norm<T, DIM>(vel, i) - pivot;
        s1 += s;
        s2 = fma(s, s, s2);
    }
    s1 = block_reduce(s1, OpAdd(), 0.0, red);
    s2 = block_reduce(s2, OpAdd(), 0.0, red);
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = s1; partials[2 * blockIdx.x + 1] = s2; }
}

__global__ void __launch_bounds__(256) moments_final_kernel(const double* __restrict__ partials, int blocks, double* __restrict__ out) {
    __shared__ double red[32];
    double s1 = 0.0, s2 = 0.0;
    for (int b = threadIdx.x; b < blocks; b += blockDim.x) { s1 += partials[2 * b]; s2 += partials[2 * b + 1]; }
    s1 = block_reduce(s1, OpAdd(), 0.0, red);
    s2 = block_reduce(s2, OpAdd(), 0.0, red);
    if (threadIdx.x == 0) { out[0] = s1; out[1] = s2; }
}

template <typename T, int DIM>
int launch_radius_kth(const void* pos, int64_t n, int64_t k, void* out, void* ws, cudaStream_t st) {
    unsigned long long* state = reinterpret_cast<unsigned long long*>(ws);
    int64_t blocks = (n + 255) / 256;
    if (blocks > kNumSMsB200 * 8) blocks = kNumSMsB200 * 8;
    radix_init_kernel<<<1, 256, 0, st>>>(state, (unsigned long long)k);
    for (int shift = KeyOf<T>::kBits - 8; shift >= 0; shift -= 8) {
        radix_hist_kernel<T, DIM><<<(int)blocks, 256, 0, st>>>((const T*)pos, n, shift, state);
        radix_pick_kernel<T><<<1, 32, 0, st>>>(shift, state, (T*)out, shift == 0);
    }
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

}  // namespace nb

using namespace nb;

extern "C" int64_t nb_metrics_workspace_bytes(void) { return 258 * 8 + (int64_t)kNumSMsB200 * 8 * 2 * 8; }

extern "C" int nb_radius_kth(const void* pos, int64_t n, int dim, int dtype, int64_t k, void* out, void* workspace,
                             int64_t workspace_bytes, void* stream) {
    if (!pos || !out || !workspace || n <= 0 || k < 0 || k >= n || (dim != 2 && dim != 3)) return NB_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < nb_metrics_workspace_bytes()) return NB_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NB_F32 && dim == 2) return launch_radius_kth<float, 2>(pos, n, k, out, workspace, st);
    if (dtype == NB_F32 && dim == 3) return launch_radius_kth<float, 3>(pos, n, k, out, workspace, st);
    if (dtype == NB_F64 && dim == 2) return launch_radius_kth<double, 2>(pos, n, k, out, workspace, st);
    if (dtype == NB_F64 && dim == 3) return launch_radius_kth<double, 3>(pos, n, k, out, workspace, st);
    return NB_ERR_INVALID_ARGUMENT;
}

extern "C" int nb_speed_moments(const void* vel, int64_t n, int dim, int dtype, double* out, void* workspace,
                                int64_t workspace_bytes, void* stream) {
    if (!vel || !out || !workspace || n <= 0 || (dim != 2 && dim != 3)) return NB_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < nb_metrics_workspace_bytes()) return NB_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (n + 255) / 256;
    if (blocks > kNumSMsB200 * 8) blocks = kNumSMsB200 * 8;
    double* part = reinterpret_cast<double*>(workspace);
    if (dtype == NB_F32 && dim == 2) speed_moments_kernel<float, 2><<<(int)blocks, 256, 0, st>>>((const float*)vel, n, part);
    else if (dtype == NB_F32 && dim == 3) speed_moments_kernel<float, 3><<<(int)blocks, 256, 0, st>>>((const float*)vel, n, part);
    else if (dtype == NB_F64 && dim == 2) speed_moments_kernel<double, 2><<<(int)blocks, 256, 0, st>>>((const double*)vel, n, part);
    else if (dtype == NB_F64 && dim == 3) speed_moments_kernel<double, 3><<<(int)blocks, 256, 0, st>>>((const double*)vel, n, part);
    else return NB_ERR_INVALID_ARGUMENT;
    NB_CUDA_LAUNCH_CHECK();
    moments_final_kernel<<<1, 256, 0, st>>>(part, (int)blocks, out);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}
