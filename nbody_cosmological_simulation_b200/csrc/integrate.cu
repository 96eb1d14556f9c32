// Packed-source emission and the fused kick-drift-kick (+ acceleration grid snap) integrator.
// Reference semantics: simulation.py:120-143 (step), quantization.py:74-88 (_grid_quantize).
// HBM-bound elementwise work: one thread per "unit" (two fp32 particles / one fp64 particle), every
// array read once and written once per tick, contiguous per warp.
#include "common.cuh"
#include "internal.cuh"

namespace nb {

// ---- exact (non-contracted) arithmetic in the state dtype ------------------------------------------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float rint_even(float a) { return rintf(a); }
__device__ __forceinline__ double rint_even(double a) { return rint(a); }

// _grid_quantize on one value given the global min/max (quantization.py:84-86); each op rounded alone.
template <typename T>
struct LinearGrid {
    T lo, span, lm1;
    bool active;     // false when (max - min) < 1e-10  (quantization.py:81) -> values pass through
    __device__ LinearGrid(const int64_t* scalars, int slot_min, int slot_max, int levels) {
        lo = (T)double_from_key(scalars[slot_min]);
        T hi = (T)double_from_key(scalars[slot_max]);
        span = sub_rn(hi, lo);
        lm1 = (T)(levels - 1);
        active = !(span < (T)1e-10);
    }
    __device__ __forceinline__ T snap(T a) const {
        if (!active) return a;
        T k = rint_even(mul_rn(div_rn(sub_rn(a, lo), span), lm1));
        return add_rn(mul_rn(div_rn(k, lm1), span), lo);
    }
};

// ---- packed source record writers ------------------------------------------------------------------
// fp32: unit = particles (2u, 2u+1); A = {x0,x1,y0,y1}; B = {z0,z1,m0,m1} (D=3) | {m0,m1} (D=2)
template <int DIM>
__device__ __forceinline__ void emit_unit_f32(char* packed, int64_t unit, const float* p0, const float* p1, float m0, float m1) {
    const int64_t chunk = unit / kChunkUnits;
    const int u = (int)(unit % kChunkUnits);
    char* base = packed + chunk * (int64_t)chunk_bytes(DIM);
    *reinterpret_cast<float4*>(base + u * 16) = make_float4(p0[0], p1[0], p0[1], p1[1]);
    if (DIM == 3) *reinterpret_cast<float4*>(base + kChunkABytes + u * 16) = make_float4(p0[2], p1[2], m0, m1);
    else          *reinterpret_cast<float2*>(base + kChunkABytes + u * 8) = make_float2(m0, m1);
}
// fp64: unit = particle u; A = {x,y}; B = {z,m} (D=3) | {m} (D=2)
template <int DIM>
__device__ __forceinline__ void emit_unit_f64(char* packed, int64_t unit, const double* p, double m) {
    const int64_t chunk = unit / kChunkUnits;
    const int u = (int)(unit % kChunkUnits);
    char* base = packed + chunk * (int64_t)chunk_bytes(DIM);
    *reinterpret_cast<double2*>(base + u * 16) = make_double2(p[0], p[1]);
    if (DIM == 3) *reinterpret_cast<double2*>(base + kChunkABytes + u * 16) = make_double2(p[2], m);
    else          *reinterpret_cast<double*>(base + kChunkABytes + u * 8) = m;
}

template <typename TM>
__device__ __forceinline__ double load_mass(const void* mass, int64_t i) { return (double)reinterpret_cast<const TM*>(mass)[i]; }

// ---- nb_pack_sources -------------------------------------------------------------------------------
template <typename T, int DIM, typename TM>
__global__ void __launch_bounds__(256) pack_kernel(const T* __restrict__ pos, const TM* __restrict__ mass, int64_t n,
                                                   char* __restrict__ packed, int64_t n_units) {
    constexpr int UP = sizeof(T) == 4 ? 2 : 1;      // particles per unit
    for (int64_t unit = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; unit < n_units;
         unit += (int64_t)gridDim.x * blockDim.x) {
        T p[UP][3];
        T m[UP];
#pragma unroll
        for (int h = 0; h < UP; ++h) {
            int64_t i = unit * UP + h;
            bool real = i < n;
            const T far = sizeof(T) == 4 ? (T)kPadCoordF32 : (T)kPadCoordF64;     // padding record: far away, mass 0
#pragma unroll
            for (int k = 0; k < DIM; ++k) p[h][k] = real ? pos[i * DIM + k] : far;
            m[h] = real ? (T)mass[i] : (T)0;
        }
        if constexpr (sizeof(T) == 4) emit_unit_f32<DIM>(packed, unit, p[0], p[UP - 1], m[0], m[UP - 1]);
        else emit_unit_f64<DIM>(packed, unit, p[0], m[0]);
    }
}

// ---- nb_kdk ----------------------------------------------------------------------------------------
// Optional source of the accelerations: the j-split partial sums of the force kernel, reduced on the fly in the
// fixed split order and with the same `(T)(Σ·scale)` rounding as accel_finalize_kernel (bit-identical), and written
// to `acc` so that the attribute stays observable.  partial == nullptr: `acc` is read as before.
struct PartialSrc { const double* partial; int splits; int64_t count; double scale; };

template <typename T>
__device__ __forceinline__ T reduce_partial(const PartialSrc& ps, int64_t e) {
    double s = 0.0;
#pragma unroll 8
    for (int sp = 0; sp < ps.splits; ++sp) s += ps.partial[(int64_t)sp * ps.count + e];   // loads batched, adds in split order
    return (T)(s * ps.scale);
}

template <typename T, int DIM, typename TM, int PHASE>
__global__ void __launch_bounds__(256) kdk_kernel(const T* __restrict__ x_in, const T* __restrict__ v_in, T* __restrict__ acc,
                                                  T* __restrict__ x_out, T* __restrict__ v_out, int64_t n, T half_dt, T dt,
                                                  int snap_levels, const int64_t* __restrict__ scalars,
                                                  const TM* __restrict__ mass, char* __restrict__ packed, int64_t n_units,
                                                  const PartialSrc ps) {
    constexpr int UP = sizeof(T) == 4 ? 2 : 1;
    LinearGrid<T> grid(scalars, NB_SLOT_ACC_MIN, NB_SLOT_ACC_MAX, snap_levels > 0 ? snap_levels : 2);
    const bool snap = snap_levels > 0;
    for (int64_t unit = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; unit < n_units;
         unit += (int64_t)gridDim.x * blockDim.x) {
        T px[UP][3];
        T pm[UP];
#pragma unroll
        for (int h = 0; h < UP; ++h) {
            const int64_t i = unit * UP + h;
            const bool real = i < n;
            if (real) {
#pragma unroll
                for (int k = 0; k < DIM; ++k) {
                    const int64_t e = i * DIM + k;
                    T a = ps.partial ? reduce_partial<T>(ps, e) : acc[e];
                    if (snap) a = grid.snap(a);
                    if (snap || ps.partial) acc[e] = a;
                    T v = v_in[e];
                    const T kick = mul_rn(a, half_dt);
                    v = add_rn(v, kick);                                     // simulation.py:141 / :132
                    if (PHASE == NB_KDK_KICK_KICK_DRIFT) v = add_rn(v, kick); // :132 of the next tick
                    v_out[e] = v;
                    if (PHASE != NB_KDK_KICK) {
                        T x = add_rn(x_in[e], mul_rn(v, dt));                // :135
                        x_out[e] = x;
                        px[h][k] = x;
                    }
                }
                if (PHASE != NB_KDK_KICK && packed) pm[h] = (T)mass[i];
            }
        }
        if (PHASE != NB_KDK_KICK && packed) {
            // padding half of the last unit / padding units of the last chunk(s): far away, mass 0
            const T far = sizeof(T) == 4 ? (T)kPadCoordF32 : (T)kPadCoordF64;
#pragma unroll
            for (int h = 0; h < UP; ++h) {
                if (unit * UP + h >= n) {
#pragma unroll
                    for (int k = 0; k < DIM; ++k) px[h][k] = far;
                    pm[h] = (T)0;
                }
            }
            if constexpr (sizeof(T) == 4) emit_unit_f32<DIM>(packed, unit, px[0], px[UP - 1], pm[0], pm[UP - 1]);
            else emit_unit_f64<DIM>(packed, unit, px[0], pm[0]);
        }
    }
}

// Vectorised variant: one thread owns TWO units (4 fp32 / 2 fp64 particles = 48 or 32 contiguous bytes per array),
// so every array is moved with 16-byte LDG/STG (3 or 2 per array) instead of 4/8-byte accesses — the scalar
// kernel above is LSU-instruction bound (2.6 TB/s at N = 2^20).  Requires 16-byte aligned base pointers; the last,
// partially filled group and all padding units fall back to element-wise code inside the same kernel.
template <typename T> struct Vec16;
template <> struct Vec16<float> { using type = float4; };
template <> struct Vec16<double> { using type = double2; };

template <typename T, int DIM, typename TM, int PHASE>
__global__ void __launch_bounds__(256) kdk_vec_kernel(const T* __restrict__ x_in, const T* __restrict__ v_in, T* __restrict__ acc,
                                                      T* __restrict__ x_out, T* __restrict__ v_out, int64_t n, T half_dt, T dt,
                                                      int snap_levels, const int64_t* __restrict__ scalars,
                                                      const TM* __restrict__ mass, char* __restrict__ packed, int64_t n_units,
                                                      const PartialSrc ps) {
    constexpr int UP = sizeof(T) == 4 ? 2 : 1;       // particles per unit
    constexpr int PPT = 2 * UP;                      // particles per thread
    constexpr int V = PPT * DIM;                     // values per thread and array
    constexpr int EPV = 16 / sizeof(T);              // elements per 16-byte vector
    constexpr int NV = V / EPV;                      // 3 (D=3) or 2 (D=2) vectors
    using VT = typename Vec16<T>::type;
    LinearGrid<T> grid(scalars, NB_SLOT_ACC_MIN, NB_SLOT_ACC_MAX, snap_levels > 0 ? snap_levels : 2);
    const bool snap = snap_levels > 0;
    const int64_t n_groups = (n_units + 1) / 2;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p0 = g * PPT;
        const int64_t left = n - p0;
        const int nreal = left >= PPT ? PPT : (left > 0 ? (int)left : 0);
        const int64_t e0 = p0 * DIM;
        T x[V], v[V], a[V];
        if (nreal == PPT) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                if (!ps.partial) reinterpret_cast<VT*>(a)[q] = reinterpret_cast<const VT*>(acc + e0)[q];
                reinterpret_cast<VT*>(v)[q] = reinterpret_cast<const VT*>(v_in + e0)[q];
                if (PHASE != NB_KDK_KICK) reinterpret_cast<VT*>(x)[q] = reinterpret_cast<const VT*>(x_in + e0)[q];
            }
        } else {
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const bool ok = e < nreal * DIM;
                a[e] = (ok && !ps.partial) ? acc[e0 + e] : (T)0;
                v[e] = ok ? v_in[e0 + e] : (T)0;
                x[e] = (ok && PHASE != NB_KDK_KICK) ? x_in[e0 + e] : (T)0;
            }
        }
        if (ps.partial) {
            // Σ over the j-splits, two doubles per load (e0 is even: V values per thread, V ∈ {4, 6, 8, 12})
            double s[V];
#pragma unroll
            for (int e = 0; e < V; ++e) s[e] = 0.0;
            if (nreal == PPT && (ps.count & 1) == 0) {          // rows of an even length keep the 16-byte alignment
#pragma unroll 4
                for (int sp = 0; sp < ps.splits; ++sp) {                   // loads of 4 splits in flight, adds in split order
                    const double2* row = reinterpret_cast<const double2*>(ps.partial + (int64_t)sp * ps.count + e0);
#pragma unroll
                    for (int q = 0; q < V / 2; ++q) { const double2 d = row[q]; s[2 * q] += d.x; s[2 * q + 1] += d.y; }
                }
            } else {
                for (int sp = 0; sp < ps.splits; ++sp) {
#pragma unroll
                    for (int e = 0; e < V; ++e)
                        if (e < nreal * DIM) s[e] += ps.partial[(int64_t)sp * ps.count + e0 + e];
                }
            }
#pragma unroll
            for (int e = 0; e < V; ++e) a[e] = (T)(s[e] * ps.scale);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
            if (snap) a[e] = grid.snap(a[e]);
            const T kick = mul_rn(a[e], half_dt);
            v[e] = add_rn(v[e], kick);                                        // simulation.py:141 / :132
            if (PHASE == NB_KDK_KICK_KICK_DRIFT) v[e] = add_rn(v[e], kick);   // :132 of the next tick
            if (PHASE != NB_KDK_KICK) x[e] = add_rn(x[e], mul_rn(v[e], dt)); // :135
        }
        if (nreal == PPT) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                if (snap || ps.partial) reinterpret_cast<VT*>(acc + e0)[q] = reinterpret_cast<VT*>(a)[q];
                reinterpret_cast<VT*>(v_out + e0)[q] = reinterpret_cast<VT*>(v)[q];
                if (PHASE != NB_KDK_KICK) reinterpret_cast<VT*>(x_out + e0)[q] = reinterpret_cast<VT*>(x)[q];
            }
        } else {
#pragma unroll
            for (int e = 0; e < V; ++e) {
                if (e < nreal * DIM) {
                    if (snap || ps.partial) acc[e0 + e] = a[e];
                    v_out[e0 + e] = v[e];
                    if (PHASE != NB_KDK_KICK) x_out[e0 + e] = x[e];
                }
            }
        }
        if (PHASE != NB_KDK_KICK && packed) {
            const T far = sizeof(T) == 4 ? (T)kPadCoordF32 : (T)kPadCoordF64;
            T pm[PPT];
#pragma unroll
            for (int h = 0; h < PPT; ++h) {
                const bool real = h < nreal;
                pm[h] = real ? (T)mass[p0 + h] : (T)0;
                if (!real) {
#pragma unroll
                    for (int k = 0; k < DIM; ++k) x[h * DIM + k] = far;
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int64_t unit = 2 * g + u;
                if (unit < n_units) {
                    if constexpr (sizeof(T) == 4) emit_unit_f32<DIM>(packed, unit, &x[(2 * u) * DIM], &x[(2 * u + 1) * DIM], pm[2 * u], pm[2 * u + 1]);
                    else emit_unit_f64<DIM>(packed, unit, &x[u * DIM], pm[u]);
                }
            }
        }
    }
}

// ---- nb_snap_accelerations -------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) snap_kernel(T* __restrict__ acc, int64_t count, int levels,
                                                   const int64_t* __restrict__ scalars) {
    LinearGrid<T> grid(scalars, NB_SLOT_ACC_MIN, NB_SLOT_ACC_MAX, levels);
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x)
        acc[e] = grid.snap(acc[e]);
}

inline int grid_for(int64_t work_items, int threads) {
    int64_t blocks = (work_items + threads - 1) / threads;
    const int64_t cap = (int64_t)kNumSMsB200 * 16;          // grid-stride beyond 16 CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename T, int DIM, typename TM>
int launch_pack(const void* pos, const void* mass, int64_t n, void* packed, int64_t total_chunks, cudaStream_t st) {
    const int64_t natural = nb_num_chunks(n, sizeof(T) == 4 ? NB_F32 : NB_F64);
    if (total_chunks != 0 && total_chunks < natural) return NB_ERR_INVALID_ARGUMENT;
    const int64_t n_units = (total_chunks ? total_chunks : natural) * kChunkUnits;
    pack_kernel<T, DIM, TM><<<grid_for(n_units, 256), 256, 0, st>>>((const T*)pos, (const TM*)mass, n, (char*)packed, n_units);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

template <typename T, int DIM, typename TM, int PHASE>
int launch_kdk(const void* x_in, const void* v_in, void* acc, void* x_out, void* v_out, int64_t n, double dt,
               int snap_levels, const int64_t* scalars, const void* mass, void* packed, int64_t total_chunks, cudaStream_t st,
               const PartialSrc ps = PartialSrc{nullptr, 0, 0, 0.0}) {
    constexpr int UP = sizeof(T) == 4 ? 2 : 1;
    // with packed output the padding units (rest of the last chunk, plus whole padding chunks) are written too
    const int64_t natural = nb_num_chunks(n, sizeof(T) == 4 ? NB_F32 : NB_F64);
    if (packed && total_chunks != 0 && total_chunks < natural) return NB_ERR_INVALID_ARGUMENT;
    const int64_t n_units = packed ? (total_chunks ? total_chunks : natural) * kChunkUnits : (n + UP - 1) / UP;
    auto aligned16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    // small systems reducing partial sums on the fly are latency-bound: one unit per thread (twice the threads) there
    const bool prefer_scalar = ps.partial != nullptr && n_units <= (int64_t)kNumSMsB200 * 256;
    if (!prefer_scalar && aligned16(x_in) && aligned16(v_in) && aligned16(acc) && aligned16(x_out) && aligned16(v_out)) {
        kdk_vec_kernel<T, DIM, TM, PHASE><<<grid_for((n_units + 1) / 2, 256), 256, 0, st>>>(
            (const T*)x_in, (const T*)v_in, (T*)acc, (T*)x_out, (T*)v_out, n, (T)(dt / 2), (T)dt, snap_levels, scalars,
            (const TM*)mass, (char*)packed, n_units, ps);
    } else {
        kdk_kernel<T, DIM, TM, PHASE><<<grid_for(n_units, 256), 256, 0, st>>>(
            (const T*)x_in, (const T*)v_in, (T*)acc, (T*)x_out, (T*)v_out, n, (T)(dt / 2), (T)dt, snap_levels, scalars,
            (const TM*)mass, (char*)packed, n_units, ps);
    }
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

}  // namespace nb

using namespace nb;

extern "C" int nb_pack_sources(const void* pos, const void* mass, int64_t n, int dim, int dtype, int mass_dtype,
                               void* packed, int64_t total_chunks, void* stream) {
    if (!pos || !mass || !packed || n <= 0 || (dim != 2 && dim != 3)) return NB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
#define NB_PACK_CASE(T, DT, TM, MDT, D) \
    if (dtype == DT && mass_dtype == MDT && dim == D) return launch_pack<T, D, TM>(pos, mass, n, packed, total_chunks, st);
    NB_PACK_CASE(float, NB_F32, float, NB_F32, 2) NB_PACK_CASE(float, NB_F32, float, NB_F32, 3)
    NB_PACK_CASE(float, NB_F32, double, NB_F64, 2) NB_PACK_CASE(float, NB_F32, double, NB_F64, 3)
    NB_PACK_CASE(double, NB_F64, float, NB_F32, 2) NB_PACK_CASE(double, NB_F64, float, NB_F32, 3)
    NB_PACK_CASE(double, NB_F64, double, NB_F64, 2) NB_PACK_CASE(double, NB_F64, double, NB_F64, 3)
#undef NB_PACK_CASE
    return NB_ERR_INVALID_ARGUMENT;
}

static int kdk_dispatch(const void* x_in, const void* v_in, void* acc, void* x_out, void* v_out, int64_t n, int dim, int dtype,
                        double dt, int phase, int snap_levels, const int64_t* scalars, const void* mass, int mass_dtype,
                        void* packed_out, int64_t total_chunks, void* stream, const PartialSrc ps) {
    if (!v_in || !acc || !v_out || n <= 0 || (dim != 2 && dim != 3)) return NB_ERR_INVALID_ARGUMENT;
    if (phase != NB_KDK_KICK && (!x_in || !x_out)) return NB_ERR_INVALID_ARGUMENT;
    if (snap_levels < 0 || snap_levels == 1 || (snap_levels > 0 && !scalars)) return NB_ERR_INVALID_ARGUMENT;
    if (packed_out && (!mass || phase == NB_KDK_KICK)) return NB_ERR_INVALID_ARGUMENT;
    if (!scalars) return NB_ERR_INVALID_ARGUMENT;       // the snap grid constructor always reads the block
    cudaStream_t st = (cudaStream_t)stream;
    if (!mass) mass_dtype = dtype;
#define NB_KDK_CASE(T, DT, TM, MDT, D, PH)                                  \
    if (dtype == DT && mass_dtype == MDT && dim == D && phase == PH)        \
        return launch_kdk<T, D, TM, PH>(x_in, v_in, acc, x_out, v_out, n, dt, snap_levels, scalars, mass, packed_out, total_chunks, st, ps);
#define NB_KDK_PHASES(T, DT, TM, MDT, D) \
    NB_KDK_CASE(T, DT, TM, MDT, D, NB_KDK_KICK_DRIFT) NB_KDK_CASE(T, DT, TM, MDT, D, NB_KDK_KICK) NB_KDK_CASE(T, DT, TM, MDT, D, NB_KDK_KICK_KICK_DRIFT)
    NB_KDK_PHASES(float, NB_F32, float, NB_F32, 2) NB_KDK_PHASES(float, NB_F32, float, NB_F32, 3)
    NB_KDK_PHASES(float, NB_F32, double, NB_F64, 2) NB_KDK_PHASES(float, NB_F32, double, NB_F64, 3)
    NB_KDK_PHASES(double, NB_F64, float, NB_F32, 2) NB_KDK_PHASES(double, NB_F64, float, NB_F32, 3)
    NB_KDK_PHASES(double, NB_F64, double, NB_F64, 2) NB_KDK_PHASES(double, NB_F64, double, NB_F64, 3)
#undef NB_KDK_PHASES
#undef NB_KDK_CASE
    return NB_ERR_INVALID_ARGUMENT;
}

extern "C" int nb_kdk(const void* x_in, const void* v_in, void* acc, void* x_out, void* v_out, int64_t n, int dim, int dtype,
                      double dt, int phase, int snap_levels, const int64_t* scalars, const void* mass, int mass_dtype,
                      void* packed_out, int64_t total_chunks, void* stream) {
    return kdk_dispatch(x_in, v_in, acc, x_out, v_out, n, dim, dtype, dt, phase, snap_levels, scalars, mass, mass_dtype, packed_out,
                        total_chunks, stream, PartialSrc{nullptr, 0, 0, 0.0});
}

int nb::kdk_from_partials(const void* x_in, const void* v_in, void* acc, void* x_out, void* v_out, int64_t n, int dim, int dtype,
                          double dt, int phase, const int64_t* scalars, const void* mass, int mass_dtype, void* packed_out,
                          int64_t total_chunks, const PartialSums& p, cudaStream_t st) {
    // the state dtype must be the dtype the reduction would have written (fp64 accelerations on fp32 state — FLOAT64
    // mode, first tick — promote the state first; the caller handles that)
    if (p.minmax || p.count != n * dim || p.out_f64 != (dtype == NB_F64)) return NB_ERR_INVALID_ARGUMENT;
    return kdk_dispatch(x_in, v_in, acc, x_out, v_out, n, dim, dtype, dt, phase, 0, scalars, mass, mass_dtype, packed_out, total_chunks,
                        (void*)st, PartialSrc{p.partial, p.splits, p.count, p.scale});
}

extern "C" int nb_snap_accelerations(void* acc, int64_t count, int acc_dtype, int levels, const int64_t* scalars, void* stream) {
    if (!acc || !scalars || count <= 0 || levels < 2) return NB_ERR_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    if (acc_dtype == NB_F32) snap_kernel<float><<<grid_for(count, 256), 256, 0, st>>>((float*)acc, count, levels, scalars);
    else if (acc_dtype == NB_F64) snap_kernel<double><<<grid_for(count, 256), 256, 0, st>>>((double*)acc, count, levels, scalars);
    else return NB_ERR_INVALID_ARGUMENT;
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}
