// The O(N) remainder of metrics.collect_metrics, as device code without sorting:
//   compute_galaxy_radius       metrics.py:81-95    k-th smallest radius by MSB-first radix select (exact)
//   compute_velocity_dispersion metrics.py:148-156  unbiased std of |v| from fp64 power sums
// Radii / speeds are formed with the reference's rounding sequence (separately rounded squares and adds, IEEE sqrt),
// so the selected order statistic is bit-identical to `torch.sort(radii)[0][k]`.
#include <type_traits>
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


// ======================================================================================================
// compute_bound_fraction (metrics.py:98-145) without a sort and without gathering state
// ======================================================================================================
// The reference ranks the stars by distance from the centre of mass (argsort), takes the cumulative mass in that order
// (cumsum) and calls a star bound when |v| < sqrt(2 G M_enclosed / max(r, 0.1)).  Only the VERDICT per star is needed,
// and it is monotone in M_enclosed.  So: a mass histogram over 2^19 monotone radius bins (the top bits of the fp32
// pattern of r) brackets every star's enclosed mass between "everything in lower bins + itself" and "everything up to
// the end of its own bin"; a star whose verdict is the same at both ends is counted at once, the few others ("doubt":
// their speed is within ~1e-4 of the escape speed) get their exact enclosed mass from a brute-force sweep.  Every stage
// is additive over i-range shards: a sharded run all-reduces the histogram (4 MB) and the doubt partial sums.
constexpr int kRadiusBinShift = 12;                       // bin = float_bits((float)r) >> 12  (r >= 0: < 2^19 bins)

template <typename T> __device__ __forceinline__ T m_sub(T a, T b);
template <> __device__ __forceinline__ float m_sub(float a, float b) { return __fsub_rn(a, b); }
template <> __device__ __forceinline__ double m_sub(double a, double b) { return __dsub_rn(a, b); }
template <typename T> __device__ __forceinline__ T m_div(T a, T b);
template <> __device__ __forceinline__ float m_div(float a, float b) { return __fdiv_rn(a, b); }
template <> __device__ __forceinline__ double m_div(double a, double b) { return __ddiv_rn(a, b); }

// sqrt(((positions - com) ** 2).sum(dim=-1)) of row i (metrics.py:122), each op rounded once
template <typename T, int DIM>
__device__ __forceinline__ T radius_about(const T* __restrict__ pos, const T* __restrict__ c, int64_t i) {
    const T dx = m_sub(pos[i * DIM], c[0]), dy = m_sub(pos[i * DIM + 1], c[1]);
    T s = m_add(m_mul(dx, dx), m_mul(dy, dy));
    if (DIM == 3) { const T dz = m_sub(pos[i * DIM + 2], c[2]); s = m_add(s, m_mul(dz, dz)); }
    return m_sqrt(s);
}
// monotone coarse bin of a radius: rounding to fp32 is monotone, and non-negative floats order like their bit patterns
template <typename T> __device__ __forceinline__ unsigned radius_bin(T r) {
    const float f = (float)r;
    return f == f ? (__float_as_uint(f < 0.f ? 0.f : f) >> kRadiusBinShift) : (0x7f800000u >> kRadiusBinShift);     // NaN -> the last bin
}

template <typename T, typename TM, int DIM>
__global__ void __launch_bounds__(256) mass_moments_kernel(const T* __restrict__ pos, const TM* __restrict__ mass, int64_t n,
                                                           double* __restrict__ partials) {
    __shared__ double red[32];
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double m = (double)mass[i];
#pragma unroll
        for (int k = 0; k < DIM; ++k) s[k] = fma(m, (double)pos[i * DIM + k], s[k]);
        s[3] += m;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double t = block_reduce(s[k], OpAdd(), 0.0, red);
        if (threadIdx.x == 0) partials[4 * blockIdx.x + k] = t;
    }
}
__global__ void __launch_bounds__(256) moments4_final_kernel(const double* __restrict__ partials, int blocks, int dim, double* __restrict__ out) {
    __shared__ double red[32];
    for (int k = 0; k < 4; ++k) {
        double s = 0.0;
        for (int b = threadIdx.x; b < blocks; b += blockDim.x) s += partials[4 * b + k];
        s = block_reduce(s, OpAdd(), 0.0, red);
        if (threadIdx.x == 0) { if (k < dim) out[k] = s; else if (k == 3) out[dim] = s; }
    }
}

template <typename T, typename TM, int DIM>
__global__ void __launch_bounds__(256) radius_mass_hist_kernel(const T* __restrict__ pos, const TM* __restrict__ mass,
                                                               const T* __restrict__ centre, int64_t n, double* __restrict__ hist) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&hist[radius_bin(radius_about<T, DIM>(pos, centre, i))], (double)mass[i]);
}

// exclusive prefix sum of `count` doubles by ONE CTA (count <= a few million: the 2^19-bin histogram takes ~20 us)
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t count) {
    __shared__ double warp_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per = (count + blockDim.x - 1) / blockDim.x;          // contiguous slice per thread
    const int64_t b = threadIdx.x * per, e = min(count, b + per);
    double s = 0.0;
    for (int64_t i = b; i < e; ++i) s += in[i];
    double inc = s;                                                     // inclusive scan of the thread totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        double w = warp_tot[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        warp_tot[lane] = wi - w;                                        // exclusive
    }
    __syncthreads();
    double run = warp_tot[warp] + (inc - s);
    for (int64_t i = b; i < e; ++i) { const double v = in[i]; out[i] = run; run += v; }
}

// a "doubt" star: everything stage 5/6 need, 40 bytes
struct DoubtRecord { double r, r_clamped, v, m; long long index; };

// verdict of metrics.py:133-139 for one star given its enclosed mass, in the dtype torch computes it in (P)
template <typename P> __device__ __forceinline__ bool bound_verdict(P v, P r_clamped, double enclosed, double two_g) {
    const P e = m_div(m_mul((P)two_g, (P)enclosed), r_clamped);          // 2 * G * enclosed_mass / r.clamp(min=0.1)
    return v < m_sqrt(e);                                               // v_mag < sqrt(...)
}

template <typename T, typename TM, int DIM>
__global__ void __launch_bounds__(256) bound_classify_kernel(const T* __restrict__ pos, const T* __restrict__ vel,
                                                             const TM* __restrict__ mass, const T* __restrict__ centre, int64_t n,
                                                             int64_t index_base, double two_g, const double* __restrict__ hist,
                                                             const double* __restrict__ prefix, unsigned long long* __restrict__ counters,
                                                             DoubtRecord* __restrict__ doubt, int64_t capacity) {
    using P = typename std::conditional<(sizeof(T) == 8 || sizeof(TM) == 8), double, float>::type;
    __shared__ unsigned long long red[32];
    unsigned long long sure = 0ull;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T r = radius_about<T, DIM>(pos, centre, i);
        const T rc = (r != r) ? r : (r < (T)0.1 ? (T)0.1 : r);          // clamp(min=0.1) keeps NaN
        const T v = row_norm<T, DIM>(vel, i);
        const unsigned b = radius_bin(r);
        const double m = (double)mass[i];
        const double lo = prefix[b] + m, hi = prefix[b] + hist[b];
        if (bound_verdict<P>((P)v, (P)rc, lo, two_g)) ++sure;                        // bound whatever the order inside the bin
        else if (bound_verdict<P>((P)v, (P)rc, hi, two_g)) {                         // depends on the order inside the bin
            const unsigned long long slot = atomicAdd(&counters[1], 1ull);
            if ((int64_t)slot < capacity) doubt[slot] = DoubtRecord{(double)r, (double)rc, (double)v, m, (long long)(index_base + i)};
        }
    }
    sure = block_reduce(sure, OpAdd(), 0ull, red);
    if (threadIdx.x == 0 && sure) atomicAdd(&counters[0], sure);
}

// partial[d] += Σ_{local j} m_j [ r_j < r_d  or  (r_j == r_d and index_j <= index_d) ]   (argsort order; the star itself counts)
template <typename T, typename TM, int DIM>
__global__ void __launch_bounds__(256) bound_resolve_kernel(const T* __restrict__ pos, const TM* __restrict__ mass,
                                                            const T* __restrict__ centre, int64_t n, int64_t index_base,
                                                            const DoubtRecord* __restrict__ doubt, int64_t n_doubt,
                                                            double* __restrict__ partial) {
    constexpr int TILE = 1024;
    __shared__ double tr[TILE], tm[TILE];
    const int64_t j0 = (int64_t)blockIdx.x * TILE;
    const int cnt = (int)min((int64_t)TILE, n - j0);
    for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
        tr[t] = (double)radius_about<T, DIM>(pos, centre, j0 + t);
        tm[t] = (double)mass[j0 + t];
    }
    __syncthreads();
    for (int64_t d = threadIdx.x; d < n_doubt; d += blockDim.x) {
        const double rd = doubt[d].r;
        const long long rel = doubt[d].index - (index_base + j0);       // tie-break position of d inside this tile's index range
        double s = 0.0;
        for (int t = 0; t < cnt; ++t) {
            const double rj = tr[t];
            if (rj < rd || (rj == rd && (long long)t <= rel)) s += tm[t];
        }
        if (s != 0.0) atomicAdd(&partial[d], s);
    }
}

template <typename P>
__global__ void __launch_bounds__(256) bound_finish_kernel(const DoubtRecord* __restrict__ doubt, int64_t n_doubt,
                                                           const double* __restrict__ enclosed, double two_g,
                                                           unsigned long long* __restrict__ counters) {
    __shared__ unsigned long long red[32];
    unsigned long long c = 0ull;
    for (int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; d < n_doubt; d += (int64_t)gridDim.x * blockDim.x)
        if (bound_verdict<P>((P)doubt[d].v, (P)doubt[d].r_clamped, enclosed[d], two_g)) ++c;
    c = block_reduce(c, OpAdd(), 0ull, red);
    if (threadIdx.x == 0 && c) atomicAdd(&counters[0], c);
}

// one 8-bit digit pass of the radix select over radii about the origin, exposed so that a sharded run can all-reduce
// the 256 counts between passes (compute_galaxy_radius across ranks)
template <typename T, int DIM>
__global__ void __launch_bounds__(256) radius_digit_hist_kernel(const T* __restrict__ pos, int64_t n, int shift,
                                                                unsigned long long prefix, unsigned long long* __restrict__ counts) {
    using K = typename KeyOf<T>::type;
    __shared__ unsigned hist[256];
    hist[threadIdx.x] = 0u;
    __syncthreads();
    const K high_mask = (shift + 8 >= KeyOf<T>::kBits) ? (K)0 : (K)(~(K)0 << (shift + 8));
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const K key = bits_of(row_norm<T, DIM>(pos, i));
        if ((key & high_mask) == ((K)prefix & high_mask)) atomicAdd(&hist[(unsigned)((key >> shift) & 0xff)], 1u);
    }
    __syncthreads();
    if (hist[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)hist[threadIdx.x]);
}

inline int grid_for_n(int64_t n) {
    int64_t blocks = (n + 255) / 256;
    if (blocks > kNumSMsB200 * 8) blocks = kNumSMsB200 * 8;
    return (int)(blocks < 1 ? 1 : blocks);
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

// ---- compute_bound_fraction stages (metrics.py:98-145) -------------------------------------------------------
#define NB_DISPATCH_TTD(CALL)                                                                                              \
    if (dtype == NB_F32 && mass_dtype == NB_F32 && dim == 2) { CALL(float, float, 2); }                                     \
    else if (dtype == NB_F32 && mass_dtype == NB_F32 && dim == 3) { CALL(float, float, 3); }                                \
    else if (dtype == NB_F32 && mass_dtype == NB_F64 && dim == 2) { CALL(float, double, 2); }                               \
    else if (dtype == NB_F32 && mass_dtype == NB_F64 && dim == 3) { CALL(float, double, 3); }                               \
    else if (dtype == NB_F64 && mass_dtype == NB_F32 && dim == 2) { CALL(double, float, 2); }                               \
    else if (dtype == NB_F64 && mass_dtype == NB_F32 && dim == 3) { CALL(double, float, 3); }                               \
    else if (dtype == NB_F64 && mass_dtype == NB_F64 && dim == 2) { CALL(double, double, 2); }                              \
    else if (dtype == NB_F64 && mass_dtype == NB_F64 && dim == 3) { CALL(double, double, 3); }                              \
    else return NB_ERR_INVALID_ARGUMENT;

extern "C" int64_t nb_radius_bins(void) { return (int64_t)1 << (32 - kRadiusBinShift - 1); }
extern "C" int64_t nb_doubt_record_bytes(void) { return (int64_t)sizeof(DoubtRecord); }

extern "C" int nb_mass_moments(const void* pos, const void* mass, int64_t n, int dim, int dtype, int mass_dtype, double* out,
                               void* workspace, int64_t workspace_bytes, void* stream) {
    if (!pos || !mass || !out || !workspace || n <= 0 || (dim != 2 && dim != 3)) return NB_ERR_INVALID_ARGUMENT;
    const int blocks = grid_for_n(n);
    if (workspace_bytes < (int64_t)blocks * 4 * 8) return NB_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t st = (cudaStream_t)stream;
    double* part = reinterpret_cast<double*>(workspace);
#define NB_CALL(T, TM, D) mass_moments_kernel<T, TM, D><<<blocks, 256, 0, st>>>((const T*)pos, (const TM*)mass, n, part)
    NB_DISPATCH_TTD(NB_CALL)
#undef NB_CALL
    NB_CUDA_LAUNCH_CHECK();
    moments4_final_kernel<<<1, 256, 0, st>>>(part, blocks, dim, out);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_radius_mass_histogram(const void* pos, const void* mass, const void* centre, int64_t n, int dim, int dtype,
                                        int mass_dtype, double* hist, void* stream) {
    if (!pos || !mass || !centre || !hist || n <= 0 || (dim != 2 && dim != 3)) return NB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = grid_for_n(n);
#define NB_CALL(T, TM, D) radius_mass_hist_kernel<T, TM, D><<<blocks, 256, 0, st>>>((const T*)pos, (const TM*)mass, (const T*)centre, n, hist)
    NB_DISPATCH_TTD(NB_CALL)
#undef NB_CALL
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_exclusive_scan_f64(const double* in, double* out, int64_t count, void* stream) {
    if (!in || !out || count <= 0) return NB_ERR_INVALID_ARGUMENT;
    exclusive_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(in, out, count);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_bound_classify(const void* pos, const void* vel, const void* mass, const void* centre, int64_t n, int64_t index_base,
                                 int dim, int dtype, int mass_dtype, double G, const double* hist, const double* prefix,
                                 uint64_t* counters, void* doubt_records, int64_t doubt_capacity, void* stream) {
    if (!pos || !vel || !mass || !centre || !hist || !prefix || !counters || !doubt_records || n <= 0 || (dim != 2 && dim != 3))
        return NB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = grid_for_n(n);
    const double two_g = 2 * G;                                  // `2 * G` is a Python float product before it meets the tensor
#define NB_CALL(T, TM, D) bound_classify_kernel<T, TM, D><<<blocks, 256, 0, st>>>((const T*)pos, (const T*)vel, (const TM*)mass, (const T*)centre, \
        n, index_base, two_g, hist, prefix, (unsigned long long*)counters, (DoubtRecord*)doubt_records, doubt_capacity)
    NB_DISPATCH_TTD(NB_CALL)
#undef NB_CALL
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_bound_resolve(const void* pos, const void* mass, const void* centre, int64_t n, int64_t index_base, int dim, int dtype,
                                int mass_dtype, const void* doubt_records, int64_t n_doubt, double* partial, void* stream) {
    if (!pos || !mass || !centre || !doubt_records || !partial || n <= 0 || n_doubt < 0 || (dim != 2 && dim != 3))
        return NB_ERR_INVALID_ARGUMENT;
    if (n_doubt == 0) return NB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t blocks = (n + 1023) / 1024;
#define NB_CALL(T, TM, D) bound_resolve_kernel<T, TM, D><<<(unsigned)blocks, 256, 0, st>>>((const T*)pos, (const TM*)mass, (const T*)centre, n, \
        index_base, (const DoubtRecord*)doubt_records, n_doubt, partial)
    NB_DISPATCH_TTD(NB_CALL)
#undef NB_CALL
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_bound_finish(const void* doubt_records, int64_t n_doubt, const double* enclosed, int dtype, int mass_dtype, double G,
                               uint64_t* counters, void* stream) {
    if (!doubt_records || !enclosed || !counters || n_doubt < 0) return NB_ERR_INVALID_ARGUMENT;
    if (n_doubt == 0) return NB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = grid_for_n(n_doubt);
    if (dtype == NB_F64 || mass_dtype == NB_F64)
        bound_finish_kernel<double><<<blocks, 256, 0, st>>>((const DoubtRecord*)doubt_records, n_doubt, enclosed, 2 * G, (unsigned long long*)counters);
    else
        bound_finish_kernel<float><<<blocks, 256, 0, st>>>((const DoubtRecord*)doubt_records, n_doubt, enclosed, 2 * G, (unsigned long long*)counters);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_radius_digit_histogram(const void* pos, int64_t n, int dim, int dtype, int shift, uint64_t prefix, uint64_t* counts,
                                         void* stream) {
    if (!pos || !counts || n <= 0 || (dim != 2 && dim != 3) || shift < 0 || (shift & 7)) return NB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = grid_for_n(n);
    if (dtype == NB_F32 && shift > 24) return NB_ERR_INVALID_ARGUMENT;
    if (dtype == NB_F64 && shift > 56) return NB_ERR_INVALID_ARGUMENT;
    if (dtype == NB_F32 && dim == 2) radius_digit_hist_kernel<float, 2><<<blocks, 256, 0, st>>>((const float*)pos, n, shift, prefix, (unsigned long long*)counts);
    else if (dtype == NB_F32) radius_digit_hist_kernel<float, 3><<<blocks, 256, 0, st>>>((const float*)pos, n, shift, prefix, (unsigned long long*)counts);
    else if (dtype == NB_F64 && dim == 2) radius_digit_hist_kernel<double, 2><<<blocks, 256, 0, st>>>((const double*)pos, n, shift, prefix, (unsigned long long*)counts);
    else if (dtype == NB_F64) radius_digit_hist_kernel<double, 3><<<blocks, 256, 0, st>>>((const double*)pos, n, shift, prefix, (unsigned long long*)counts);
    else return NB_ERR_INVALID_ARGUMENT;
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}
