// Energy and rotation-curve reductions.
//   potential energy  simulation.py:176-192   O(N²), same source-streaming skeleton as the force kernel
//   kinetic energy    simulation.py:170-174   O(N)
//   rotation curve    metrics.py:25-78        O(N) one-pass binned sum/count (the reference does
//                                              3·bins masked passes with a host sync each)
// Reductions: per-thread partials -> warp shuffles -> one fp64 value per CTA -> fixed-order final sum
// (deterministic; no floating-point atomics on the energy path).
#include <type_traits>
#include "stream.cuh"

namespace nb {

// ---- potential: Σ_{i∈targets} m_i ( Σ_j m_j / sqrt(d²_ij) − m_i / sqrt(ε²) ) ---------------------------
// The j == i term is removed by subtracting the identical expression (d² == ε² exactly when dx == 0), so
// the pair loop carries no index compare; requires targets ⊂ sources (always true: i-range shards).
// Half-ring partition of the unordered pairs at chunk granularity: with C chunks and δ = (q − p) mod C, target chunk p
// takes source chunk q with weight 1 if 0 < 2δ < C, ½ if δ == 0 (its own square: both orders, self term removed
// by the caller) or 2δ == C (C even: both sides take it), 0 otherwise — every unordered pair is counted exactly once
// and every target chunk has the same amount of work, so an i-range shard on any rank costs the same (the plain
// upper triangle gives rank 0 twice the mean).  ring == 0: no index correspondence between targets and sources;
// every chunk gets ½ (the reference's full-matrix form ½ Σ_{j≠i}).
__device__ __forceinline__ double pair_weight(int64_t own, int64_t q, int64_t ring) {
    if (ring <= 0) return 0.5;
    int64_t d = q - own;
    if (d < 0) d += ring;
    if (d == 0 || 2 * d == ring) return 0.5;
    return 2 * d < ring ? 1.0 : 0.0;
}

template <int DIM_, int IPT, int THREADS_>
struct PotentialF32 {
    static constexpr int DIM = DIM_;
    static constexpr int THREADS = THREADS_;
    float2 nx[IPT], ny[IPT], nz[IPT];
    float2 acc[IPT];
    double sum[IPT];                         // Σ over the streamed chunks of weight(own chunk, source chunk) · Σ_j m_j / r_ij
    int64_t own_chunk[IPT];                  // global chunk index of target t
    int64_t ring;                            // number of chunks of the source set (0: every chunk has weight ½)
    float2 eps2;
    __device__ __forceinline__ void init(const float* pos, int64_t n_tgt, float e2) {
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
            int64_t i = (int64_t)blockIdx.x * (THREADS * IPT) + t * THREADS + threadIdx.x;
            if (i >= n_tgt) i = n_tgt - 1;
            const float x = pos[i * DIM + 0], y = pos[i * DIM + 1], z = DIM == 3 ? pos[i * DIM + 2] : 0.f;
            nx[t] = make_float2(-x, -x); ny[t] = make_float2(-y, -y); nz[t] = make_float2(-z, -z);
            acc[t] = make_float2(0.f, 0.f);
            sum[t] = 0.0;
        }
        eps2 = make_float2(e2, e2);
    }
    __device__ __forceinline__ void chunk(const unsigned char* s, int64_t chunk_index) {
        const float4* A = reinterpret_cast<const float4*>(s);
        const float4* B4 = reinterpret_cast<const float4*>(s + kChunkABytes);
        const float2* B2 = reinterpret_cast<const float2*>(s + kChunkABytes);
#pragma unroll 4
        for (int p = 0; p < kChunkUnits; ++p) {
            const float4 a = A[p];
            const float2 xs = make_float2(a.x, a.y), ys = make_float2(a.z, a.w);
            float2 zs = make_float2(0.f, 0.f), ms;
            if (DIM == 3) { const float4 b = B4[p]; zs = make_float2(b.x, b.y); ms = make_float2(b.z, b.w); }
            else ms = B2[p];
#pragma unroll
            for (int t = 0; t < IPT; ++t) {
                const float2 dx = add2(xs, nx[t]), dy = add2(ys, ny[t]);
                float2 d2 = fma2(dx, dx, eps2);
                d2 = fma2(dy, dy, d2);
                if (DIM == 3) { const float2 dz = add2(zs, nz[t]); d2 = fma2(dz, dz, d2); }
                const float2 r = make_float2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
                acc[t] = fma2(ms, r, acc[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
            sum[t] += pair_weight(own_chunk[t], chunk_index, ring) * (double)(acc[t].x + acc[t].y);
            acc[t] = make_float2(0.f, 0.f);
        }
    }
};

__device__ __forceinline__ double rsqrt_full(double d2) {
    // MUFU.RSQ64H seed + one third-order (Halley) step: y = y0 (1 + e/2 + 3e²/8), e = 1 − d2·y0²
    const double y0 = rsqrt64h(d2);
    const double e = fma(-d2, y0 * y0, 1.0);
    return fma(y0 * e, fma(0.375, e, 0.5), y0);
}

template <int DIM_, int IPT, int THREADS_>
struct PotentialF64 {
    static constexpr int DIM = DIM_;
    static constexpr int THREADS = THREADS_;
    double xi[IPT], yi[IPT], zi[IPT];
    double sum[IPT];
    int64_t own_chunk[IPT];
    int64_t ring;
    double eps2;
    __device__ __forceinline__ void init(const double* pos, int64_t n_tgt, double e2) {
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
            int64_t i = (int64_t)blockIdx.x * (THREADS * IPT) + t * THREADS + threadIdx.x;
            if (i >= n_tgt) i = n_tgt - 1;
            xi[t] = pos[i * DIM + 0]; yi[t] = pos[i * DIM + 1]; zi[t] = DIM == 3 ? pos[i * DIM + 2] : 0.0;
            sum[t] = 0.0;
        }
        eps2 = e2;
    }
    __device__ __forceinline__ void chunk(const unsigned char* s, int64_t chunk_index) {
        const double2* A = reinterpret_cast<const double2*>(s);
        const double2* B2 = reinterpret_cast<const double2*>(s + kChunkABytes);
        const double* B1 = reinterpret_cast<const double*>(s + kChunkABytes);
        double part[IPT];
#pragma unroll
        for (int t = 0; t < IPT; ++t) part[t] = 0.0;
#pragma unroll 2
        for (int p = 0; p < kChunkUnits; ++p) {
            const double2 a = A[p];
            double zs = 0.0, m;
            if (DIM == 3) { const double2 b = B2[p]; zs = b.x; m = b.y; } else m = B1[p];
#pragma unroll
            for (int t = 0; t < IPT; ++t) {
                const double dx = a.x - xi[t], dy = a.y - yi[t];
                double d2 = fma(dx, dx, eps2);
                d2 = fma(dy, dy, d2);
                if (DIM == 3) { const double dz = zs - zi[t]; d2 = fma(dz, dz, d2); }
                part[t] = fma(m, rsqrt_full(d2), part[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < IPT; ++t) sum[t] += pair_weight(own_chunk[t], chunk_index, ring) * part[t];
    }
};

template <typename T, typename TM, int DIM, int IPT, int THREADS>
__global__ void __launch_bounds__(THREADS + 32) potential_kernel(const char* __restrict__ src, int64_t n_chunks,
                                                                 const T* __restrict__ pos_tgt, const TM* __restrict__ mass_tgt,
                                                                 int64_t n_tgt, int64_t tgt_offset, int triangular,   /* targets are an index-aligned slice of the sources: half-ring partition */
                                                                 int chunks_per_split, double eps_sq,
                                                                 double* __restrict__ block_partials) {
    // Unordered pairs: out = Σ_{i<j} m_i m_j / r_ij.  With chunk-aligned targets (triangular != 0) a CTA streams the
    // half ring of source chunks that starts at its own first chunk (pair_weight above); its own squares carry the
    // j == i term, which is removed below.  Otherwise every chunk has weight ½: ½ Σ_{j≠i}, the reference's
    // full-matrix form.
    __shared__ double red[32];
    constexpr int TB = THREADS * IPT;
    constexpr int CS = sizeof(T) == 4 ? 2 * kChunkUnits : kChunkUnits;       // sources per chunk
    static_assert(TB % CS == 0, "a target block must cover whole chunks");
    using Cons = typename std::conditional<sizeof(T) == 4, PotentialF32<DIM, IPT, THREADS>, PotentialF64<DIM, IPT, THREADS>>::type;
    Cons cons;
    const bool is_consumer = threadIdx.x < THREADS;
    if (is_consumer) cons.init(pos_tgt, n_tgt, (T)eps_sq);
    const int64_t first_tgt = tgt_offset + (int64_t)blockIdx.x * TB;
    const int64_t g0 = triangular ? first_tgt / CS : 0;                       // first chunk of this block's targets
    // the block's own chunks end with its REAL targets (a partial last block must not reach into the next rank's slot)
    const int64_t blk_tgts = min((int64_t)TB, n_tgt - (int64_t)blockIdx.x * TB);
    const int64_t own_chunks = (blk_tgts + CS - 1) / CS;
    const int64_t span = triangular ? min(n_chunks, own_chunks + n_chunks / 2) : n_chunks;   // ring offsets this block needs
    const int64_t c0 = min(span, (int64_t)blockIdx.y * chunks_per_split);
    const int64_t c1 = min(span, c0 + (int64_t)chunks_per_split);
    cons.ring = triangular ? n_chunks : 0;
#pragma unroll
    for (int t = 0; t < IPT; ++t) {
        int64_t i = (int64_t)blockIdx.x * TB + t * THREADS + (is_consumer ? threadIdx.x : 0);
        if (i >= n_tgt) i = n_tgt - 1;
        cons.own_chunk[t] = (tgt_offset + i) / CS;
    }
    stream_sources(src, g0 + c0, g0 + c1, cons, triangular ? n_chunks : 0);
    double mine = 0.0;
    if (is_consumer) {
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
            const int64_t i = (int64_t)blockIdx.x * (THREADS * IPT) + t * THREADS + threadIdx.x;
            if (i < n_tgt) {
                const double m = (double)mass_tgt[i];
                double sd = cons.sum[t];
                const int64_t self_off = cons.own_chunk[t] - g0;       // the j == i term lives in exactly one j-split
                if (self_off >= c0 && self_off < c1) {
                    double self;
                    if constexpr (sizeof(T) == 4) self = (double)((float)m * rsqrt_approx((float)eps_sq));
                    else self = m * rsqrt_full(eps_sq);
                    sd -= 0.5 * self;                                  // its square was taken with weight ½
                }
                mine += m * sd;
            }
        }
    }
    const double tot = block_reduce(mine, OpAdd(), 0.0, red);
    if (threadIdx.x == 0) block_partials[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
}

// fixed-order sum of `count` doubles -> out[0]   (single CTA)
__global__ void __launch_bounds__(1024) final_sum_kernel(const double* __restrict__ partials, int64_t count, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t e = threadIdx.x; e < count; e += blockDim.x) s += partials[e];
    s = block_reduce(s, OpAdd(), 0.0, red);
    if (threadIdx.x == 0) out[0] = s;
}

// ---- kinetic: Σ_i m_i Σ_k v_ik² -------------------------------------------------------------------------
template <typename T, typename TM, int DIM>
__global__ void __launch_bounds__(256) kinetic_kernel(const T* __restrict__ vel, const TM* __restrict__ mass, int64_t n,
                                                      double* __restrict__ block_partials) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v2 = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) { const double v = (double)vel[i * DIM + k]; v2 = fma(v, v, v2); }
        s = fma((double)mass[i], v2, s);
    }
    s = block_reduce(s, OpAdd(), 0.0, red);
    if (threadIdx.x == 0) block_partials[blockIdx.x] = s;
}

// ---- rotation curve ----------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T r_mul(T a, T b);
template <> __device__ __forceinline__ float r_mul(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double r_mul(double a, double b) { return __dmul_rn(a, b); }
template <typename T> __device__ __forceinline__ T r_add(T a, T b);
template <> __device__ __forceinline__ float r_add(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double r_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float r_sqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double r_sqrt(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float r_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double r_div(double a, double b) { return __ddiv_rn(a, b); }

// radii = sqrt((positions ** 2).sum(dim=-1))   metrics.py:48 — ((x²+y²)+z²), each op rounded
template <typename T, int DIM>
__device__ __forceinline__ T radius_of(const T* __restrict__ pos, int64_t i) {
    T s = r_add(r_mul(pos[i * DIM], pos[i * DIM]), r_mul(pos[i * DIM + 1], pos[i * DIM + 1]));
    if (DIM == 3) s = r_add(s, r_mul(pos[i * DIM + 2], pos[i * DIM + 2]));
    return r_sqrt(s);
}

template <typename T, int DIM>
__global__ void __launch_bounds__(256) radius_max_kernel(const T* __restrict__ pos, int64_t n, int64_t* __restrict__ scalars) {
    __shared__ long long red[32];
    long long best = kKeyLowest;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const long long k = key_from_double((double)radius_of<T, DIM>(pos, i));
        best = k > best ? k : best;
    }
    best = block_reduce(best, OpMax(), (long long)kKeyLowest, red);
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<long long*>(scalars + NB_SLOT_RADIUS_MAX), best);
}

constexpr int kMaxBins = 1024;

template <typename T, int DIM>
__global__ void __launch_bounds__(256) rotation_curve_kernel(const T* __restrict__ pos, const T* __restrict__ vel, int64_t n,
                                                             const T* __restrict__ edges, int num_bins,
                                                             double* __restrict__ sum_vt, int64_t* __restrict__ count) {
    __shared__ T s_edges[kMaxBins + 1];
    __shared__ double s_sum[kMaxBins];
    __shared__ unsigned long long s_cnt[kMaxBins];
    for (int b = threadIdx.x; b <= num_bins; b += blockDim.x) s_edges[b] = edges[b];
    for (int b = threadIdx.x; b < num_bins; b += blockDim.x) { s_sum[b] = 0.0; s_cnt[b] = 0ull; }
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T r = radius_of<T, DIM>(pos, i);
        // |x·vy − y·vx| / clamp(r, 0.1)     metrics.py:55-57 (components 0 and 1 only, also for D=3)
        const T lz = r_add(r_mul(pos[i * DIM], vel[i * DIM + 1]), -r_mul(pos[i * DIM + 1], vel[i * DIM]));
        const T rc = (r != r) ? r : (r < (T)0.1 ? (T)0.1 : r);
        const T vt = r_div(lz < 0 ? -lz : lz, rc);
        // half-open bins [e_b, e_{b+1}); edges are non-decreasing: binary search for the last edge <= r
        if (!(r >= s_edges[0]) || !(r < s_edges[num_bins])) continue;
        int lo = 0, hi = num_bins;                       // invariant: e[lo] <= r < e[hi]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (r >= s_edges[mid]) lo = mid; else hi = mid; }
        atomicAdd(&s_sum[lo], (double)vt);
        atomicAdd(&s_cnt[lo], 1ull);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < num_bins; b += blockDim.x) {
        if (s_cnt[b]) {
            atomicAdd(&sum_vt[b], s_sum[b]);
            atomicAdd(reinterpret_cast<unsigned long long*>(count) + b, s_cnt[b]);
        }
    }
}

}  // namespace nb

using namespace nb;

extern "C" int64_t nb_energy_workspace_bytes(int64_t n_targets) {
    if (n_targets <= 0) return 0;
    // one double per CTA of the largest grid the planners below can choose
    const int64_t blocks_i = (n_targets + 511) / 512;
    int64_t ctas = blocks_i * 64;                          // planner cap: 64 splits
    if (ctas < kNumSMsB200 * 16) ctas = kNumSMsB200 * 16;
    return ctas * (int64_t)sizeof(double);
}

template <typename T, typename TM, int DIM>
static int launch_potential(const void* packed_src, int64_t n_src, const void* pos_tgt, const void* mass_tgt, int64_t n_tgt,
                            int64_t tgt_offset, int dtype, double eps_sq, double* out, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
    constexpr int TH = 256, IPT = 2;
    const int64_t n_chunks = nb_num_chunks(n_src, dtype);
    auto k = potential_kernel<T, TM, DIM, IPT, TH>;
    const int smem = stream_smem_bytes(DIM);
    int occ = 1;
    const int frc = kernel_occupancy((const void*)k, TH + 32, smem, &occ);
    if (frc != NB_OK) return frc;
    const int triangular = tgt_offset >= 0 && tgt_offset % nb_chunk_sources(dtype) == 0;
    // chunks a target block streams: its own plus half the ring (all of them in the full-matrix form)
    const int64_t own = (TH * IPT) / nb_chunk_sources(dtype);
    const int64_t span = triangular ? (own + n_chunks / 2 < n_chunks ? own + n_chunks / 2 : n_chunks) : n_chunks;
    const SplitPlan sp = plan_splits(n_tgt, span, TH * IPT, occ, 64);
    const int blocks_i = sp.blocks_i, cps = sp.chunks_per_split, splits = sp.splits;
    const int64_t ctas = (int64_t)blocks_i * splits;
    if (workspace_bytes < ctas * (int64_t)sizeof(double)) return NB_ERR_WORKSPACE_TOO_SMALL;
    k<<<dim3(blocks_i, splits), TH + 32, smem, st>>>((const char*)packed_src, n_chunks, (const T*)pos_tgt, (const TM*)mass_tgt, n_tgt,
                                                    triangular ? tgt_offset : 0, triangular, cps, eps_sq, (double*)workspace);
    NB_CUDA_LAUNCH_CHECK();
    final_sum_kernel<<<1, 1024, 0, st>>>((const double*)workspace, ctas, out);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_potential_energy(const void* packed_src, int64_t n_src, const void* pos_tgt, const void* mass_tgt, int64_t n_tgt,
                                   int64_t tgt_offset, int dim, int dtype, int mass_dtype, double eps_sq, double* out, void* workspace,
                                   int64_t workspace_bytes, void* stream) {
    if (!packed_src || !pos_tgt || !mass_tgt || !out || !workspace || n_src <= 0 || n_tgt <= 0 || (dim != 2 && dim != 3))
        return NB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
#define NB_PE_CASE(T, DT, TM, MDT, D) \
    if (dtype == DT && mass_dtype == MDT && dim == D) \
        return launch_potential<T, TM, D>(packed_src, n_src, pos_tgt, mass_tgt, n_tgt, tgt_offset, dtype, eps_sq, out, workspace, workspace_bytes, st);
    NB_PE_CASE(float, NB_F32, float, NB_F32, 2) NB_PE_CASE(float, NB_F32, float, NB_F32, 3)
    NB_PE_CASE(float, NB_F32, double, NB_F64, 2) NB_PE_CASE(float, NB_F32, double, NB_F64, 3)
    NB_PE_CASE(double, NB_F64, float, NB_F32, 2) NB_PE_CASE(double, NB_F64, float, NB_F32, 3)
    NB_PE_CASE(double, NB_F64, double, NB_F64, 2) NB_PE_CASE(double, NB_F64, double, NB_F64, 3)
#undef NB_PE_CASE
    return NB_ERR_INVALID_ARGUMENT;
}

extern "C" int nb_kinetic_energy(const void* vel, const void* mass, int64_t n, int dim, int dtype, int mass_dtype, double* out,
                                 void* workspace, int64_t workspace_bytes, void* stream) {
    if (!vel || !mass || !out || !workspace || n <= 0 || (dim != 2 && dim != 3)) return NB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (n + 255) / 256;
    if (blocks > kNumSMsB200 * 16) blocks = kNumSMsB200 * 16;
    if (workspace_bytes < blocks * (int64_t)sizeof(double)) return NB_ERR_WORKSPACE_TOO_SMALL;
    double* part = (double*)workspace;
#define NB_KE_CASE(T, DT, TM, MDT, D) \
    if (dtype == DT && mass_dtype == MDT && dim == D) kinetic_kernel<T, TM, D><<<(int)blocks, 256, 0, st>>>((const T*)vel, (const TM*)mass, n, part); else
    NB_KE_CASE(float, NB_F32, float, NB_F32, 2) NB_KE_CASE(float, NB_F32, float, NB_F32, 3)
    NB_KE_CASE(float, NB_F32, double, NB_F64, 2) NB_KE_CASE(float, NB_F32, double, NB_F64, 3)
    NB_KE_CASE(double, NB_F64, float, NB_F32, 2) NB_KE_CASE(double, NB_F64, float, NB_F32, 3)
    NB_KE_CASE(double, NB_F64, double, NB_F64, 2) NB_KE_CASE(double, NB_F64, double, NB_F64, 3)
    return NB_ERR_INVALID_ARGUMENT;
#undef NB_KE_CASE
    NB_CUDA_LAUNCH_CHECK();
    final_sum_kernel<<<1, 1024, 0, st>>>(part, blocks, out);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_radius_max(const void* pos, int64_t n, int dim, int dtype, int64_t* scalars, void* stream) {
    if (!pos || !scalars || n <= 0 || (dim != 2 && dim != 3)) return NB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (n + 255) / 256;
    if (blocks > kNumSMsB200 * 8) blocks = kNumSMsB200 * 8;
    if (dtype == NB_F32 && dim == 2) radius_max_kernel<float, 2><<<(int)blocks, 256, 0, st>>>((const float*)pos, n, scalars);
    else if (dtype == NB_F32) radius_max_kernel<float, 3><<<(int)blocks, 256, 0, st>>>((const float*)pos, n, scalars);
    else if (dtype == NB_F64 && dim == 2) radius_max_kernel<double, 2><<<(int)blocks, 256, 0, st>>>((const double*)pos, n, scalars);
    else if (dtype == NB_F64) radius_max_kernel<double, 3><<<(int)blocks, 256, 0, st>>>((const double*)pos, n, scalars);
    else return NB_ERR_INVALID_ARGUMENT;
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_rotation_curve(const void* pos, const void* vel, int64_t n, int dim, int dtype, const void* edges, int num_bins,
                                 double* sum_vt, int64_t* count, void* stream) {
    if (!pos || !vel || !edges || !sum_vt || !count || n <= 0 || (dim != 2 && dim != 3)) return NB_ERR_INVALID_ARGUMENT;
    if (num_bins < 1 || num_bins > kMaxBins) return NB_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (n + 255) / 256;
    if (blocks > kNumSMsB200 * 4) blocks = kNumSMsB200 * 4;
    if (dtype == NB_F32 && dim == 2) rotation_curve_kernel<float, 2><<<(int)blocks, 256, 0, st>>>((const float*)pos, (const float*)vel, n, (const float*)edges, num_bins, sum_vt, count);
    else if (dtype == NB_F32) rotation_curve_kernel<float, 3><<<(int)blocks, 256, 0, st>>>((const float*)pos, (const float*)vel, n, (const float*)edges, num_bins, sum_vt, count);
    else if (dtype == NB_F64 && dim == 2) rotation_curve_kernel<double, 2><<<(int)blocks, 256, 0, st>>>((const double*)pos, (const double*)vel, n, (const double*)edges, num_bins, sum_vt, count);
    else if (dtype == NB_F64) rotation_curve_kernel<double, 3><<<(int)blocks, 256, 0, st>>>((const double*)pos, (const double*)vel, n, (const double*)edges, num_bins, sum_vt, count);
    else return NB_ERR_INVALID_ARGUMENT;
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}
