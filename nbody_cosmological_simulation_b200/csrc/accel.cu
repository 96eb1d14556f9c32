// All-pairs softened-gravity force evaluation — GalaxySimulation._compute_accelerations,
// simulation.py:74-118, with the d² quantisers of quantization.py:21-127 fused in.
//
// Roofline: FP32 (FMA pipe) / FP64 (DFMA pipe) issue-bound, 0 algorithmic HBM bytes per pair: the
// packed source set (16 B per fp32 particle) is L2-resident and streamed through shared memory.
//   fp32, D=3: per source PAIR and target 12 packed fp32x2 ops (11 when all masses are equal) + 2 MUFU.RSQ;
//   D=2: 9 (8) packed ops + 2 MUFU.RSQ.  Ops with <= 2 distinct register pairs issue in 2 cycles, the three
//   accumulate FFMA2 (3 distinct pairs) in 3: floor 25 cycles per 64 interactions per SM sub-partition.
//   fp64, D=3: 16 (15) DFMA-class ops + 1 MUFU.RSQ64H per pair (seed + one cubic-corrected Newton step).
// Accumulation: per-thread fp32x2 partial sums over one chunk (256 sources), flushed into fp64
// accumulators per chunk => the Σ_j error does not grow with N (SURVEY.md §7 "hard parts").
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cooperative_groups.h>
#include "stream.cuh"
#include "lut.cuh"
#include "internal.cuh"

namespace nb {

// tuning knobs of the fast-lookup kernel (tools/tune_lutf.sh builds variants)
#ifndef NB_LUTF_MINB
#define NB_LUTF_MINB 2
#endif
#ifndef NB_LUTF_UNROLL
#define NB_LUTF_UNROLL 4
#endif
#ifndef NB_LUTF_IPT
#define NB_LUTF_IPT 2
#endif
#ifndef NB_LUTF_REDO_INLINE
#define NB_LUTF_REDO_INLINE __noinline__
#endif
#ifndef NB_LUTF_THREADS
#define NB_LUTF_THREADS 256
#endif
#ifndef NB_F32_PERTURB
#define NB_F32_PERTURB 0              // source-order perturbations of the float pair loop (same arithmetic up to the d² summation order)
#endif
#ifndef NB_F32_ACC_VARIANT
#define NB_F32_ACC_VARIANT 0          // how the float kernels write the three accumulates (tools/f32_loop_probe.sh compares)
#endif
enum QMode { Q_F32 = 0, Q_F16 = 1, Q_BF16 = 2, Q_LUT = 3, Q_F64 = 4, Q_LUTF = 5 };   // Q_LUTF: fast level lookup (lut.cuh), L <= 256

struct AccelArgs {
    const char* src;          // packed sources
    int64_t n_chunks;         // chunks this launch streams (the whole packed set, or a contiguous window of it starting at src)
    const void* pos_tgt;      // (n_tgt, DIM) state dtype
    int64_t n_tgt;
    int chunks_per_split;
    double* partial;          // [splits][n_tgt][DIM]
    double* partial_phi;      // PHI variants: [splits][n_tgt] sums of m_j / r_ij (the potential at each target), else unused
    double eps_sq;
    const void* table;        // level table (Q_LUT)
    int levels;
    int lut_rep;              // shared-memory replication of the level table (1, 2, 4 or 8 copies, see accel_kernel)
    float neg_zero;           // -0.0f, passed at run time so that the compiler cannot fold fma(d, d, -0) back into a mul
    float uniform_mass;       // Q_LUTF: != 0 when every real source has this mass (the per-pair mass multiply is dropped)
    int splits_before;        // windowed evaluation: split slots already used by earlier windows
    int max_splits;           // > 0: cap on the split count of this launch
};

// Level table layout (Q_LUT): float4 entry[k] = { T_{k+1}, g_k, g_{k+1}, 0 } for k = 0..L-1, preceded by a
// 16-byte header { lo2 (log2 of lower bound), scale (levels-1)/(hi2-lo2), min_val, degenerate flag }
// (written by build_level_table_kernel in quantize.cu).

// ======================================================================================================
// fp32 state, packed-pair arithmetic
// ======================================================================================================
// UNI: all source masses are equal (checked by the caller): the per-pair `·m_j` is dropped and the common mass is
// applied once per target in the finalize pass; padding records then rely on their far-away position (w == 0).
// PHI (Q_F32 only): the pair loop also accumulates Σ_j m_j / r_ij per target — r = rsqrt(d²) is already in a register,
// so the potential costs ONE more packed op per source pair (FFMA2; FADD2 with uniform masses) instead of a second
// O(N²) pass (simulation.py:176-192 evaluated on the state the force pass has just seen).
template <int DIM_, int QMODE, int IPT, int THREADS_, bool UNI = false, int UNROLL = 4, bool PHI = false>
struct ForceF32 {
    static constexpr int DIM = DIM_;
    static constexpr int THREADS = THREADS_;
    static constexpr int TARGETS_PER_THREAD = IPT;
    static constexpr bool HAS_PHI = PHI;
    static_assert(!PHI || QMODE == Q_F32, "the potential is defined on the unquantised d² (simulation.py:181-186)");
    float2 nx[IPT], ny[IPT], nz[IPT];      // {-x_i, -x_i}: targets, negated and duplicated for packed adds
    float2 ax[IPT], ay[IPT], az[IPT];      // chunk-local sums; .x = even sources, .y = odd sources
    double sx[IPT], sy[IPT], sz[IPT];      // running fp64 sums
    float2 ap[IPT];                        // PHI: chunk-local Σ m_j / r_ij
    double sp[IPT];
    float2 eps2;
    float neg_zero;
    const float4* lut;                     // shared-memory copy of the level table (Q_LUT)
    float lo2, scale, min_val;
    // Q_LUTF (lut.cuh): lane-replicated factor table g[k][lane] (an LDS.32 at k·128 + lane·4 never bank-conflicts
    // and moves 4× fewer shared-memory wavefronts than the 16-byte entries of Q_LUT), thresholds for the slow path
    float nxs[IPT], nys[IPT], nzs[IPT];    // −x_i as scalars: packed ops take them in the broadcast form (one register, not a pair)
    float scale_s, cm_s, eps_s;
    uint32_t lutg_lane;                    // shared-window address of g[0][lane]
    const float* thr_smem;                 // T_j, j = 0..P
    LutFast lf;
    float2 scale2, cm2;
    uint32_t kmask;                        // (P-1) << kLutFb: the level bits of W
    int n_levels;
    bool clamp_lo;                         // eps² < min_val: d² must be clamped from below (quantization.py:106)
    float inv_uniform_mass;                // 1/m when all real sources have mass m (0: general masses)

    __device__ __forceinline__ void init(const AccelArgs& a, const float4* lut_smem, int tile) {
        const float* pos = reinterpret_cast<const float*>(a.pos_tgt);
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
            int64_t i = (int64_t)tile * (THREADS * IPT) + t * THREADS + threadIdx.x;
            if (i >= a.n_tgt) i = a.n_tgt - 1;
            const float x = pos[i * DIM + 0], y = pos[i * DIM + 1], z = DIM == 3 ? pos[i * DIM + 2] : 0.f;
            nx[t] = make_float2(-x, -x); ny[t] = make_float2(-y, -y); nz[t] = make_float2(-z, -z);
            if (QMODE == Q_LUTF) { nxs[t] = -x; nys[t] = -y; nzs[t] = -z; }
            ax[t] = ay[t] = az[t] = make_float2(0.f, 0.f);
            sx[t] = sy[t] = sz[t] = 0.0;
            if (PHI) { ap[t] = make_float2(0.f, 0.f); sp[t] = 0.0; }
        }
        const float e = (float)a.eps_sq;                // softening_sq cast to the tensor dtype (simulation.py:86)
        eps2 = make_float2(e, e);
        neg_zero = a.neg_zero;
        lut = lut_smem;
        if (QMODE == Q_LUT) {
            // copy (lane mod rep) of every entry: the 8 lanes of an LDS.128 quarter-warp then hit 8 different 16-byte
            // bank groups whatever their levels are (ncu: 3.0e9 bank conflicts per launch at L=256 without this)
            lut = lut_smem + 1 + (threadIdx.x & (a.lut_rep - 1));
            lut_stride = a.lut_rep;
            const float4 h = lut_smem[0];
            scale = h.y; min_val = h.z;
            lo2 = fmaf(-h.x, h.y, -0.5f + 1.0f / 64.0f);     // additive constant of the level estimate (see lut_factor)
        }
        if (QMODE == Q_LUTF) {
            const float4* tab = reinterpret_cast<const float4*>(a.table);
            lf = lut_fast_load(tab, a.levels);
            min_val = tab[0].z;
            clamp_lo = e < min_val;
            inv_uniform_mass = a.uniform_mass != 0.f ? 1.0f / a.uniform_mass : 0.f;
            scale2 = make_float2(lf.scale, lf.scale);
            cm2 = make_float2(lf.cm, lf.cm);
            scale_s = lf.scale; cm_s = lf.cm; eps_s = e;
            kmask = (uint32_t)(lf.p - 1) << kLutFb;
            n_levels = a.levels;
            lutg_lane = smem_u32(lut_smem) + 4u * (threadIdx.x & 31);
            thr_smem = reinterpret_cast<const float*>(lut_smem) + 32 * lf.p;
        }
    }

    // Q_LUTF: force factors of a source pair against one target.  Per scalar: MUFU.LG2, half an FFMA2, LOP3 + LEA.HI
    // (address), LDS.32, LOP3 (doubt predicate).
    __device__ __forceinline__ float2 lutf_lookup(float2 t, bool& doubt) const {
        const float2 w = fma2(make_float2(lg2_approx(t.x), lg2_approx(t.y)), make_float2(scale_s, scale_s), make_float2(cm_s, cm_s));
        const uint32_t wx = __float_as_uint(w.x), wy = __float_as_uint(w.y);
        doubt = doubt || ((wx & lf.zmask) == 0u) || ((wy & lf.zmask) == 0u);
        float2 g;
        // address = &g[0][lane] + k·128 in the 32-bit shared window: LOP3 (mask the level bits) + LEA.HI (shift, add)
        g.x = lds_f32(lutg_lane + ((wx & kmask) >> (kLutFb - 7)));
        g.y = lds_f32(lutg_lane + ((wy & kmask) >> (kLutFb - 7)));
        return g;
    }
    // Exact factor of one scalar pair (slow path): FULL = the factor itself, else the correction g_exact − g_fast that
    // turns the fast path's contribution into the exact one (0 unless the pair is in doubt AND its level differs).
    // t = d² in the reference's op order (defines the level), tf = the fused d² the main loop looked up
    template <bool FULL>
    __device__ __forceinline__ float lutf_exact(float t, float tf) const {
        const uint32_t w = lut_wbits(tf, lf.scale, lf.cm);
        const int k = lut_exact_level(t, w, lf, n_levels, thr_smem);
        const float ge = lds_f32(lutg_lane + (uint32_t)(k * 128));
        if (FULL) return ge;
        const float gf = lds_f32(lutg_lane + ((w & kmask) >> (kLutFb - 7)));
        return ge - gf;
    }
    template <bool FULL, bool CLAMP, bool MUL_MASS>
    __device__ NB_LUTF_REDO_INLINE void lutf_redo(const unsigned char* s, int p) {
        const float4 a = reinterpret_cast<const float4*>(s)[p];
        const float2 xs = make_float2(a.x, a.y), ys = make_float2(a.z, a.w);
        float2 zs = make_float2(0.f, 0.f), ms;
        if (DIM == 3) { const float4 b = reinterpret_cast<const float4*>(s + kChunkABytes)[p]; zs = make_float2(b.x, b.y); ms = make_float2(b.z, b.w); }
        else ms = reinterpret_cast<const float2*>(s + kChunkABytes)[p];
        // massless records (the padding of the last chunk: 240 identical far-away points whose common W may well sit in
        // the doubt zone and then overflow every thread's queue) contribute exactly 0 whatever their level is
        if (MUL_MASS && ms.x == 0.f && ms.y == 0.f) return;
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
            const float2 dx = add2(xs, nx[t]), dy = add2(ys, ny[t]);
            float2 dz = make_float2(0.f, 0.f);
            if (DIM == 3) dz = add2(zs, nz[t]);
            float2 tq = dist_sq<false>(dx, dy, dz), tf = dist_sq<true>(dx, dy, dz);
            if (CLAMP) {
                tq = make_float2(fmaxf(tq.x, min_val), fmaxf(tq.y, min_val));
                tf = make_float2(fmaxf(tf.x, min_val), fmaxf(tf.y, min_val));
            }
            float2 w = make_float2(lutf_exact<FULL>(tq.x, tf.x), lutf_exact<FULL>(tq.y, tf.y));
            if (MUL_MASS) w = mul2(w, ms);
            ax[t] = fma2(w, dx, ax[t]);
            ay[t] = fma2(w, dy, ay[t]);
            if (DIM == 3) az[t] = fma2(w, dz, az[t]);
        }
    }

    // d² for a source pair: fused (fp32 mode, tolerance 1e-5) or the reference's exact rounding
    // sequence (modes whose next step is a snap: fp16/bf16 round trip, log-grid index).
    template <bool FUSED = (QMODE == Q_F32)>
    __device__ __forceinline__ float2 dist_sq(float2 dx, float2 dy, float2 dz) const {
        if (FUSED) {
#if NB_F32_PERTURB & 2
            float2 d2 = DIM == 3 ? fma2(dz, dz, eps2) : eps2;
            d2 = fma2(dy, dy, d2);
            d2 = fma2(dx, dx, d2);
            return d2;
#else
            float2 d2 = fma2(dx, dx, eps2);
            d2 = fma2(dy, dy, d2);
            if (DIM == 3) d2 = fma2(dz, dz, d2);
            return d2;
#endif
        } else {
            // The reference's exact rounding sequence rn(rn(rn(dx²)+rn(dy²))[+rn(dz²)])+ε²) with packed ops.  ptxas
            // contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (seen in SASS; changes d² by an ulp and flips fp16/bf16/
            // grid snaps), so each square is written as fma(d, d, −0) — exactly rn(d²) — with the −0 a run-time kernel
            // argument: a literal −0 is folded back into a mul and contracted again (also seen in SASS).
            const float2 nz2 = make_float2(neg_zero, neg_zero);
            float2 s = add2(fma2(dx, dx, nz2), fma2(dy, dy, nz2));
            if (DIM == 3) s = add2(s, fma2(dz, dz, nz2));
            return add2(s, eps2);
        }
    }

    __device__ __forceinline__ float lut_factor(float d2) const {
        // t = clamp(d², min); k = round((log t − lo)/(hi − lo)·(L−1)) found exactly from a MUFU.LG2 estimate biased to
        // k−1 ≤ k_lo ≤ k plus one comparison against the exact threshold T_{k_lo+1}.  t ≥ t_lo by construction, so only
        // the upper clamp is needed (padding records / Inf / NaN); rint() is done with the 1.5·2²³ magic add (FMA
        // pipe + LOP3) instead of F2I, which would share the 16-lane XU pipe with MUFU.LG2.
        const float t = fmaxf(d2, min_val);
        float kf = fmaf(lg2_approx(t), scale, lo2);                      // lo2 holds (−lo2·scale − ½ + margin), see init()
        kf = fminf(kf, levels_m1_f);
        const int k_lo = __float_as_int(kf + 12582912.0f) & 0x3fffff;
        const float4 e = lut[k_lo * lut_stride];
        return t >= e.x ? e.z : e.y;
    }
    float levels_m1_f;
    int lut_stride;

    __device__ __forceinline__ void chunk(const unsigned char* s, int64_t c) {
        if (QMODE == Q_LUTF) {
            // Uniform masses: the pair loop drops the mass multiply (the common mass is applied once per target in the
            // reduction) — except in a chunk that ends with padding records (mass 0, always at the end of a chunk), which
            // takes the multiplying loop and is rescaled by 1/m when its partial sums are flushed.
            const float last_mass = DIM == 3 ? reinterpret_cast<const float4*>(s + kChunkABytes)[kChunkUnits - 1].w
                                             : reinterpret_cast<const float2*>(s + kChunkABytes)[kChunkUnits - 1].y;
            const bool plain = inv_uniform_mass != 0.f && last_mass != 0.f;
            if (plain) { if (clamp_lo) chunk_lutf<true, false>(s, 1.f); else chunk_lutf<false, false>(s, 1.f); }
            else {
                const float rescale = inv_uniform_mass != 0.f ? inv_uniform_mass : 1.f;
                if (clamp_lo) chunk_lutf<true, true>(s, rescale); else chunk_lutf<false, true>(s, rescale);
            }
            return;
        }
        if (!UNI && QMODE != Q_LUT) {
            // General masses, but THIS chunk's 256 sources may still share one mass (mass classes laid out in blocks:
            // jitter_test.py:45-86 nested levels, reality_glitch_tests.py:366-397 wall galaxy): then the pair loop is the
            // uniform-mass one (11 packed ops instead of 12) and the common mass scales the chunk's flush.  Decided per
            // warp from shared memory (4 LDS + a vote per 128 iterations); chunks that end in padding stay general.
            const int lane = threadIdx.x & 31;
            float m_first, eq = 1.f;
            if (DIM == 3) {
                const float4* B = reinterpret_cast<const float4*>(s + kChunkABytes);
                m_first = B[0].z;
#pragma unroll
                for (int q = 0; q < kChunkUnits / 32; ++q) { const float4 b = B[lane + 32 * q]; if (b.z != m_first || b.w != m_first) eq = 0.f; }
            } else {
                const float2* B = reinterpret_cast<const float2*>(s + kChunkABytes);
                m_first = B[0].x;
#pragma unroll
                for (int q = 0; q < kChunkUnits / 32; ++q) { const float2 b = B[lane + 32 * q]; if (b.x != m_first || b.y != m_first) eq = 0.f; }
            }
            if (__all_sync(0xffffffffu, eq != 0.f) && m_first != 0.f) { chunk_direct<true>(s, m_first); return; }
        }
        chunk_direct<UNI>(s, 1.f);
    }

    // Main loop without a branch: every iteration (one source pair x IPT targets) takes the fast lookup; an iteration
    // with a d² in doubt pushes its index into an 8-entry queue held in TWO registers (predicated funnel shift, LEA,
    // IADD).  The queue is drained after the chunk, adding (g_exact − g_fast)·m·d for the queued iterations.  A thread
    // sees 0.26 doubtful iterations per chunk on average (L = 256), so more than 8 happen about once in 10¹¹
    // thread-chunks for ordinary tables (a 4-entry queue overflowed 1.6 times per launch at N = 10⁴ and the straggling
    // CTA doubled the kernel time); tables whose levels are denser than floats overflow always and redo every chunk
    // on the slow path.
    template <bool CLAMP, bool MUL_MASS>
    __device__ __forceinline__ void chunk_lutf(const unsigned char* s, float rescale) {
        const float4* A = reinterpret_cast<const float4*>(s);
        const float4* B4 = reinterpret_cast<const float4*>(s + kChunkABytes);
        const float2* B2 = reinterpret_cast<const float2*>(s + kChunkABytes);
        uint32_t q0 = 0, q1 = 0, queued = 0;
#pragma unroll UNROLL
        for (int p = 0; p < kChunkUnits; ++p) {
            const float4 a = A[p];
            const float2 xs = make_float2(a.x, a.y), ys = make_float2(a.z, a.w);
            float2 zs = make_float2(0.f, 0.f), ms;
            if (DIM == 3) { const float4 b = B4[p]; zs = make_float2(b.x, b.y); ms = make_float2(b.z, b.w); }
            else ms = B2[p];
            bool doubt = false;
#pragma unroll
            for (int t = 0; t < IPT; ++t) {
                const float2 dx = add2(xs, make_float2(nxs[t], nxs[t])), dy = add2(ys, make_float2(nys[t], nys[t]));
                float2 dz = make_float2(0.f, 0.f);
                if (DIM == 3) dz = add2(zs, make_float2(nzs[t], nzs[t]));
                const float2 e2 = make_float2(eps_s, eps_s);
                float2 tq = fma2(dx, dx, e2);                // fused d²: < 7 ulp from the exact-order value, inside the lookup margin (lut.cuh)
                tq = fma2(dy, dy, tq);
                if (DIM == 3) tq = fma2(dz, dz, tq);   // fused d²: < 7 ulp from the exact-order value, inside the lookup margin (lut.cuh)
                if (CLAMP) tq = make_float2(fmaxf(tq.x, min_val), fmaxf(tq.y, min_val));
                float2 w = lutf_lookup(tq, doubt);
                if (MUL_MASS) w = mul2(w, ms);
                ax[t] = fma2(w, dx, ax[t]);
                ay[t] = fma2(w, dy, ay[t]);
                if (DIM == 3) az[t] = fma2(w, dz, az[t]);
            }
            if (doubt) { q1 = __funnelshift_l(q0, q1, 8); q0 = (q0 << 8) + (uint32_t)p; ++queued; }
        }
        if (queued) {
            if (queued > 8u) {
#pragma unroll
                for (int t = 0; t < IPT; ++t) ax[t] = ay[t] = az[t] = make_float2(0.f, 0.f);
                for (int p = 0; p < kChunkUnits; ++p) lutf_redo<true, CLAMP, MUL_MASS>(s, p);
            } else {
                for (; queued; --queued, q0 = __funnelshift_r(q0, q1, 8), q1 >>= 8) lutf_redo<false, CLAMP, MUL_MASS>(s, (int)(q0 & 0xffu));
            }
        }
        if (MUL_MASS && rescale != 1.f) {
#pragma unroll
            for (int t = 0; t < IPT; ++t) {
                ax[t] = make_float2(ax[t].x * rescale, ax[t].y * rescale);
                ay[t] = make_float2(ay[t].x * rescale, ay[t].y * rescale);
                az[t] = make_float2(az[t].x * rescale, az[t].y * rescale);
            }
        }
        flush();
    }

    // `scale`: common mass of a chunk that ran the uniform-mass loop inside a general-mass launch (else 1)
    __device__ __forceinline__ void flush(float scale = 1.f) {
        // flush the chunk-local fp32 sums into the fp64 accumulators
        const double sc = (double)scale;
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
            sx[t] = fma((double)(ax[t].x + ax[t].y), sc, sx[t]); ax[t] = make_float2(0.f, 0.f);
            sy[t] = fma((double)(ay[t].x + ay[t].y), sc, sy[t]); ay[t] = make_float2(0.f, 0.f);
            if (DIM == 3) { sz[t] = fma((double)(az[t].x + az[t].y), sc, sz[t]); az[t] = make_float2(0.f, 0.f); }
            if (PHI) { sp[t] = fma((double)(ap[t].x + ap[t].y), sc, sp[t]); ap[t] = make_float2(0.f, 0.f); }
        }
    }

    // ULOOP: every source of this chunk has the same mass — it is left out of the pair loop (applied by the caller's
    // finalize pass when the whole launch is uniform (UNI), by flush(chunk_mass) otherwise)
    template <bool ULOOP>
    __device__ __forceinline__ void chunk_direct(const unsigned char* s, float chunk_mass) {
        const float4* A = reinterpret_cast<const float4*>(s);
        const float4* B4 = reinterpret_cast<const float4*>(s + kChunkABytes);
        const float2* B2 = reinterpret_cast<const float2*>(s + kChunkABytes);
#pragma unroll UNROLL
        for (int p = 0; p < kChunkUnits; ++p) {
            const float4 a = A[p];
            const float2 xs = make_float2(a.x, a.y), ys = make_float2(a.z, a.w);
            float2 zs = make_float2(0.f, 0.f), ms;
            if (DIM == 3) { const float4 b = B4[p]; zs = make_float2(b.x, b.y); ms = make_float2(b.z, b.w); }
            else ms = B2[p];
#if NB_F32_ACC_VARIANT == 2
            // Two phases per source pair: first the weights of ALL targets, then ALL accumulates.  With the accumulates of a
            // target adjacent and nothing independent left to slot between them, ptxas keeps `w` in the operand-reuse cache
            // for the 2nd and 3rd FFMA2 of each triple (2 issue cycles instead of 3: an FFMA2 reads three register PAIRS).
            float2 wv[IPT], dxv[IPT], dyv[IPT], dzv[IPT];
#pragma unroll
            for (int t = 0; t < IPT; ++t) {
                dxv[t] = add2(xs, nx[t]);
                dyv[t] = add2(ys, ny[t]);
                dzv[t] = DIM == 3 ? add2(zs, nz[t]) : make_float2(0.f, 0.f);
                float2 d2 = dist_sq(dxv[t], dyv[t], dzv[t]);
                if (QMODE == Q_LUT) {
                    wv[t] = mul2(make_float2(lut_factor(d2.x), lut_factor(d2.y)), ms);
                } else {
                    if (QMODE == Q_F16) { const __half2 h = __floats2half2_rn(d2.x, d2.y); d2 = __half22float2(h); }
                    else if (QMODE == Q_BF16) { const __nv_bfloat162 h = __floats2bfloat162_rn(d2.x, d2.y); d2 = __bfloat1622float2(h); }
                    const float2 r = make_float2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
                    wv[t] = ULOOP ? mul2(mul2(r, r), r) : mul2(mul2(r, r), mul2(r, ms));
                    if (PHI) ap[t] = ULOOP ? add2(r, ap[t]) : fma2(ms, r, ap[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < IPT; ++t) {
                ax[t] = fma2(wv[t], dxv[t], ax[t]);
                ay[t] = fma2(wv[t], dyv[t], ay[t]);
                if (DIM == 3) az[t] = fma2(wv[t], dzv[t], az[t]);
            }
        }
#else
#pragma unroll
            for (int tt = 0; tt < IPT; ++tt) {
                const int t = (NB_F32_PERTURB & 4) ? IPT - 1 - tt : tt;
#if NB_F32_PERTURB & 1
                const float2 dy = add2(ys, ny[t]);
                const float2 dx = add2(xs, nx[t]);              // diff = pos[j] − pos[i]   simulation.py:83
#else
                const float2 dx = add2(xs, nx[t]);              // diff = pos[j] − pos[i]   simulation.py:83
                const float2 dy = add2(ys, ny[t]);
#endif
                float2 dz = make_float2(0.f, 0.f);
                if (DIM == 3) dz = add2(zs, nz[t]);
                float2 d2 = dist_sq(dx, dy, dz);
                float2 w;                                       // m_j · f(d²) without G (hoisted), or with G (LUT)
                if (QMODE == Q_LUT) {
                    w = mul2(make_float2(lut_factor(d2.x), lut_factor(d2.y)), ms);
                } else {
                    if (QMODE == Q_F16) {                       // dist_sq.half().float()    quantization.py:56
                        const __half2 h = __floats2half2_rn(d2.x, d2.y);
                        d2 = __half22float2(h);
                    } else if (QMODE == Q_BF16) {               // dist_sq.bfloat16().float() quantization.py:53
                        const __nv_bfloat162 h = __floats2bfloat162_rn(d2.x, d2.y);
                        d2 = __bfloat1622float2(h);
                    }
                    const float2 r = make_float2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
                    if (ULOOP) w = (NB_F32_PERTURB & 8) ? mul2(r, mul2(r, r)) : mul2(mul2(r, r), r);   // 1 / d²^1.5 (common mass applied at the flush / in finalize)
                    else w = mul2(mul2(r, r), mul2(r, ms));     // m_j / d²^1.5             simulation.py:97-105
                    if (PHI) ap[t] = ULOOP ? add2(r, ap[t]) : fma2(ms, r, ap[t]);    // Σ_j m_j / r_ij   simulation.py:185-188
                }
                // Σ_j w·diff (simulation.py:112).  Register-bank note (tools/regbank.cu): an FFMA2 with three distinct
                // register pairs issues in 3 cycles, not 2 (two banks, one 64-lane read each per cycle), so this loop's
                // floor is 6+6+4+9 = 25 cycles per 64 interactions, not 22.  Tried and measured slower: scalar FFMAs on
                // the halves (x-halves are all even registers: 2-3 cycles each) and a 3-instruction asm block.
#if NB_F32_ACC_VARIANT == 1
                // the three accumulates as ONE asm statement: ptxas keeps them adjacent, so `w` stays in the operand-reuse
                // cache for the second and third (an FFMA2 with three distinct register pairs costs 3 issue cycles, one with
                // a reused operand 2)
                if (DIM == 3) {
                    asm("{\n\t.reg .b64 w, a, b, c, x, y, z;\n\t"
                        "mov.b64 w, {%6, %7};\n\tmov.b64 a, {%8, %9};\n\tmov.b64 b, {%10, %11};\n\tmov.b64 c, {%12, %13};\n\t"
                        "mov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tmov.b64 z, {%4, %5};\n\t"
                        "fma.rn.f32x2 x, w, a, x;\n\tfma.rn.f32x2 y, w, b, y;\n\tfma.rn.f32x2 z, w, c, z;\n\t"
                        "mov.b64 {%0, %1}, x;\n\tmov.b64 {%2, %3}, y;\n\tmov.b64 {%4, %5}, z;\n\t}"
                        : "+f"(ax[t].x), "+f"(ax[t].y), "+f"(ay[t].x), "+f"(ay[t].y), "+f"(az[t].x), "+f"(az[t].y)
                        : "f"(w.x), "f"(w.y), "f"(dx.x), "f"(dx.y), "f"(dy.x), "f"(dy.y), "f"(dz.x), "f"(dz.y));
                } else {
                    ax[t] = fma2(w, dx, ax[t]);
                    ay[t] = fma2(w, dy, ay[t]);
                }
#elif NB_F32_ACC_VARIANT == 3
                if (DIM == 3) az[t] = fma2(w, dz, az[t]);
                ay[t] = fma2(w, dy, ay[t]);
                ax[t] = fma2(w, dx, ax[t]);
#else
                ax[t] = fma2(w, dx, ax[t]);
                ay[t] = fma2(w, dy, ay[t]);
                if (DIM == 3) az[t] = fma2(w, dz, az[t]);
#endif
            }
        }
#endif
        flush(chunk_mass);
    }

    __device__ __forceinline__ void store(const AccelArgs& a, int tile, int split) const {
        double* out = a.partial + (int64_t)split * a.n_tgt * DIM;
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
            const int64_t i = (int64_t)tile * (THREADS * IPT) + t * THREADS + threadIdx.x;
            if (i < a.n_tgt) {
                out[i * DIM + 0] = sx[t]; out[i * DIM + 1] = sy[t];
                if (DIM == 3) out[i * DIM + 2] = sz[t];
                if (PHI) a.partial_phi[(int64_t)split * a.n_tgt + i] = sp[t];
            }
        }
    }
};

// ======================================================================================================
// fp64 state (QMODE Q_F64: all-double; Q_F32/F16/BF16: d² cast down as quantization.py:48-56 does)
// and "mixed": fp32 state evaluated in FLOAT64 mode (d² formed in fp32 with the reference's exact
// rounding sequence, then widened — quantization.py:45 applied to an fp32 dist_sq).
// ======================================================================================================
__device__ __forceinline__ double mass_over_dist_cubed(double d2, double m) {
    // m · d2^(-3/2): MUFU.RSQ64H seed y0 (rel. err δ≈2^-22), e = 1 − d2·y0² ≈ 2δ,
    // d2^(-3/2) = y0³ (1−e)^(-3/2) = y0³ (1 + e(3/2 + 15/8 e) + O(e³))   — 7 DFMA-class ops, error ~1e-19+ulp
    const double y0 = rsqrt64h(d2);
    const double t = y0 * y0;
    const double e = fma(-d2, t, 1.0);
    const double ce = fma(1.875, e, 1.5) * e;
    const double w = (m * y0) * t;
    return fma(w, ce, w);
}

__device__ __forceinline__ double inv_dist_cubed(double d2) {     // mass_over_dist_cubed with m == 1: one DMUL less
    const double y0 = rsqrt64h(d2);
    const double t = y0 * y0;
    const double e = fma(-d2, t, 1.0);
    const double ce = fma(1.875, e, 1.5) * e;
    const double w = y0 * t;
    return fma(w, ce, w);
}

// PHI (Q_F64 only): Σ_j m_j / r_ij from the same pass — 1/r = d² · d²^(-3/2), one more DFMA per pair.
template <int DIM_, int QMODE, int IPT, int THREADS_, bool UNI = false, int UNROLL = 2, bool PHI = false>
struct ForceF64 {
    static constexpr int DIM = DIM_;
    static constexpr int THREADS = THREADS_;
    static constexpr int TARGETS_PER_THREAD = IPT;
    static constexpr bool HAS_PHI = PHI;
    static_assert(!PHI || QMODE == Q_F64, "the potential is defined on the unquantised d² (simulation.py:181-186)");
    double xi[IPT], yi[IPT], zi[IPT];
    double sx[IPT], sy[IPT], sz[IPT];
    double sp[IPT];
    double eps2;

    __device__ __forceinline__ void init(const AccelArgs& a, const float4*, int tile) {
        const double* pos = reinterpret_cast<const double*>(a.pos_tgt);
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
            int64_t i = (int64_t)tile * (THREADS * IPT) + t * THREADS + threadIdx.x;
            if (i >= a.n_tgt) i = a.n_tgt - 1;
            xi[t] = pos[i * DIM + 0]; yi[t] = pos[i * DIM + 1]; zi[t] = DIM == 3 ? pos[i * DIM + 2] : 0.0;
            sx[t] = sy[t] = sz[t] = 0.0;
            if (PHI) sp[t] = 0.0;
        }
        eps2 = a.eps_sq;
    }
    __device__ __forceinline__ void chunk(const unsigned char* s, int64_t) {
        if (!UNI && QMODE == Q_F64) {
            // general masses, but this chunk's 128 sources may share one mass: uniform-mass loop (15 ops instead of 16)
            // into chunk-local sums, scaled by the common mass when they are added to the running sums (see ForceF32)
            const int lane = threadIdx.x & 31;
            double m_first;
            bool eq = true;
            if (DIM == 3) {
                const double2* B = reinterpret_cast<const double2*>(s + kChunkABytes);
                m_first = B[0].y;
#pragma unroll
                for (int q = 0; q < kChunkUnits / 32; ++q) eq = eq && B[lane + 32 * q].y == m_first;
            } else {
                const double* B = reinterpret_cast<const double*>(s + kChunkABytes);
                m_first = B[0];
#pragma unroll
                for (int q = 0; q < kChunkUnits / 32; ++q) eq = eq && B[lane + 32 * q] == m_first;
            }
            if (__all_sync(0xffffffffu, eq) && m_first != 0.0) {
                double lx[IPT], ly[IPT], lz[IPT], lp[IPT];
#pragma unroll
                for (int t = 0; t < IPT; ++t) lx[t] = ly[t] = lz[t] = lp[t] = 0.0;
                pair_loop<true>(s, lx, ly, lz, lp);
#pragma unroll
                for (int t = 0; t < IPT; ++t) {
                    sx[t] = fma(lx[t], m_first, sx[t]); sy[t] = fma(ly[t], m_first, sy[t]);
                    if (DIM == 3) sz[t] = fma(lz[t], m_first, sz[t]);
                    if (PHI) sp[t] = fma(lp[t], m_first, sp[t]);
                }
                return;
            }
        }
        pair_loop<UNI>(s, sx, sy, sz, sp);
    }
    template <bool ULOOP>
    __device__ __forceinline__ void pair_loop(const unsigned char* s, double (&ox)[IPT], double (&oy)[IPT], double (&oz)[IPT], double (&op)[IPT]) {
        const double2* A = reinterpret_cast<const double2*>(s);
        const double2* B2 = reinterpret_cast<const double2*>(s + kChunkABytes);
        const double* B1 = reinterpret_cast<const double*>(s + kChunkABytes);
#pragma unroll UNROLL
        for (int p = 0; p < kChunkUnits; ++p) {
            const double2 a = A[p];
            double zs = 0.0, m;
            if (DIM == 3) { const double2 b = B2[p]; zs = b.x; m = b.y; } else m = B1[p];
#pragma unroll
            for (int t = 0; t < IPT; ++t) {
                const double dx = a.x - xi[t], dy = a.y - yi[t], dz = DIM == 3 ? zs - zi[t] : 0.0;
                double d2 = fma(dx, dx, eps2);
                d2 = fma(dy, dy, d2);
                if (DIM == 3) d2 = fma(dz, dz, d2);
                double w;
                if (QMODE == Q_F64) {
                    w = ULOOP ? inv_dist_cubed(d2) : mass_over_dist_cubed(d2, m);
                    if (PHI) op[t] = fma(d2, w, op[t]);                           // [m_j] · d² · d²^(-3/2) = [m_j] / r_ij
                } else {
                    float u = (float)d2;                                         // dist_sq.float()
                    if (QMODE == Q_F16) u = __half2float(__float2half_rn(u));
                    if (QMODE == Q_BF16) u = __bfloat162float(__float2bfloat16_rn(u));
                    const float r = rsqrt_approx(u);
                    w = (double)((r * r) * (r * (float)m));
                }
                ox[t] = fma(w, dx, ox[t]);
                oy[t] = fma(w, dy, oy[t]);
                if (DIM == 3) oz[t] = fma(w, dz, oz[t]);
            }
        }
    }
    __device__ __forceinline__ void store(const AccelArgs& a, int tile, int split) const {
        double* out = a.partial + (int64_t)split * a.n_tgt * DIM;
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
            const int64_t i = (int64_t)tile * (THREADS * IPT) + t * THREADS + threadIdx.x;
            if (i < a.n_tgt) {
                out[i * DIM + 0] = sx[t]; out[i * DIM + 1] = sy[t];
                if (DIM == 3) out[i * DIM + 2] = sz[t];
                if (PHI) a.partial_phi[(int64_t)split * a.n_tgt + i] = sp[t];
            }
        }
    }
};

template <int DIM_, int IPT, int THREADS_>
struct ForceMixed {       // fp32 sources/targets, FLOAT64 mode
    static constexpr bool HAS_PHI = false;
    static constexpr int DIM = DIM_;
    static constexpr int THREADS = THREADS_;
    static constexpr int TARGETS_PER_THREAD = IPT;
    float xi[IPT], yi[IPT], zi[IPT];
    double sx[IPT], sy[IPT], sz[IPT];
    float eps2;

    __device__ __forceinline__ void init(const AccelArgs& a, const float4*, int tile) {
        const float* pos = reinterpret_cast<const float*>(a.pos_tgt);
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
            int64_t i = (int64_t)tile * (THREADS * IPT) + t * THREADS + threadIdx.x;
            if (i >= a.n_tgt) i = a.n_tgt - 1;
            xi[t] = pos[i * DIM + 0]; yi[t] = pos[i * DIM + 1]; zi[t] = DIM == 3 ? pos[i * DIM + 2] : 0.f;
            sx[t] = sy[t] = sz[t] = 0.0;
        }
        eps2 = (float)a.eps_sq;
    }
    __device__ __forceinline__ void one(float xs, float ys, float zs, float m, int t) {
        const float dx = __fsub_rn(xs, xi[t]), dy = __fsub_rn(ys, yi[t]);
        const float dz = DIM == 3 ? __fsub_rn(zs, zi[t]) : 0.f;
        float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (DIM == 3) s = __fadd_rn(s, __fmul_rn(dz, dz));
        const double d2 = (double)__fadd_rn(s, eps2);                      // fp32 d², widened (quantization.py:45)
        const double w = mass_over_dist_cubed(d2, (double)m);
        sx[t] = fma(w, (double)dx, sx[t]);
        sy[t] = fma(w, (double)dy, sy[t]);
        if (DIM == 3) sz[t] = fma(w, (double)dz, sz[t]);
    }
    __device__ __forceinline__ void chunk(const unsigned char* s, int64_t) {
        const float4* A = reinterpret_cast<const float4*>(s);
        const float4* B4 = reinterpret_cast<const float4*>(s + kChunkABytes);
        const float2* B2 = reinterpret_cast<const float2*>(s + kChunkABytes);
#pragma unroll 2
        for (int p = 0; p < kChunkUnits; ++p) {
            const float4 a = A[p];
            float z0 = 0.f, z1 = 0.f, m0, m1;
            if (DIM == 3) { const float4 b = B4[p]; z0 = b.x; z1 = b.y; m0 = b.z; m1 = b.w; }
            else { const float2 b = B2[p]; m0 = b.x; m1 = b.y; }
#pragma unroll
            for (int t = 0; t < IPT; ++t) {
                one(a.x, a.z, z0, m0, t);
                one(a.y, a.w, z1, m1, t);
            }
        }
    }
    __device__ __forceinline__ void store(const AccelArgs& a, int tile, int split) const {
        double* out = a.partial + (int64_t)split * a.n_tgt * DIM;
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
            const int64_t i = (int64_t)tile * (THREADS * IPT) + t * THREADS + threadIdx.x;
            if (i < a.n_tgt) {
                out[i * DIM + 0] = sx[t]; out[i * DIM + 1] = sy[t];
                if (DIM == 3) out[i * DIM + 2] = sz[t];
            }
        }
    }
};

// ======================================================================================================
// kernels
// ======================================================================================================
constexpr int kMaxLevelsSmem = 4096;      // level table entries staged in shared memory (64 KB + header)

// LUTKIND: 0 = no level table, 1 = 16-byte entries replicated per bank group (Q_LUT), 2 = fast lookup (Q_LUTF)
template <class Consumer, int LUTKIND>
__global__ void __launch_bounds__(Consumer::THREADS + 32, LUTKIND == 2 ? NB_LUTF_MINB : (Consumer::HAS_PHI ? 3 : 0)) accel_kernel(const AccelArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const bool is_consumer = threadIdx.x < Consumer::THREADS;
    const float4* lut_smem = nullptr;
    if (LUTKIND == 1) {
        // the level table sits behind the streaming stages
        float4* dst = reinterpret_cast<float4*>(smem + stream_smem_bytes(Consumer::DIM));
        const float4* srcT = reinterpret_cast<const float4*>(a.table);
        if (threadIdx.x == 0) dst[0] = srcT[0];
        for (int e = threadIdx.x; e < a.levels * a.lut_rep; e += blockDim.x) dst[1 + e] = srcT[1 + e / a.lut_rep];
        lut_smem = dst;
        __syncthreads();
    }
    if (LUTKIND == 2) {
        // g[k][lane] for k < P (0 beyond the last level), then T_j for j = 0..P (lut.cuh)
        float* dst = reinterpret_cast<float*>(smem + stream_smem_bytes(Consumer::DIM));
        const float4* srcT = reinterpret_cast<const float4*>(a.table);
        const int P = lut_pow2ceil(a.levels);
        for (int e = threadIdx.x; e < P * 32; e += blockDim.x) {
            const int k = e >> 5;
            dst[e] = k < a.levels ? srcT[1 + k].y : 0.f;
        }
        for (int j = threadIdx.x; j <= P; j += blockDim.x)
            dst[P * 32 + j] = j == 0 ? 0.f : (j <= a.levels ? srcT[j].x : __int_as_float(0x7f800000));
        lut_smem = reinterpret_cast<const float4*>(dst);
        __syncthreads();
    }
    Consumer cons;
    if (is_consumer) cons.init(a, lut_smem, blockIdx.x);
    if constexpr (LUTKIND == 1) cons.levels_m1_f = (float)(a.levels - 1);
    const int64_t c0 = (int64_t)blockIdx.y * a.chunks_per_split;
    const int64_t c1 = min(a.n_chunks, c0 + (int64_t)a.chunks_per_split);
    stream_sources(a.src, c0, c1, cons);
    if (is_consumer) cons.store(a, blockIdx.x, blockIdx.y);
}

// acc_out[i,k] = scale · Σ_splits partial;  optional min/max of the outputs -> scalars (quantization.py:78-79)
template <typename TOUT, bool MINMAX>
__global__ void __launch_bounds__(256) accel_finalize_kernel(const double* __restrict__ partial, int splits, int64_t count,
                                                             double scale, TOUT* __restrict__ out, int64_t* __restrict__ scalars) {
    __shared__ long long red[32];
    long long kmin = kKeyHighest, kmax = kKeyLowest;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
#pragma unroll 8
        for (int sp = 0; sp < splits; ++sp) s += partial[(int64_t)sp * count + e];        // loads batched, adds in split order
        const TOUT v = (TOUT)(s * scale);
        out[e] = v;
        if (MINMAX) {
            const long long k = key_from_double((double)v);
            kmin = k < kmin ? k : kmin;
            kmax = k > kmax ? k : kmax;
        }
    }
    if (MINMAX) {
        kmin = block_reduce(kmin, OpMin(), (long long)kKeyHighest, red);
        kmax = block_reduce(kmax, OpMax(), (long long)kKeyLowest, red);
        if (threadIdx.x == 0) {
            atomicMin(reinterpret_cast<long long*>(scalars + NB_SLOT_ACC_MIN), kmin);
            atomicMax(reinterpret_cast<long long*>(scalars + NB_SLOT_ACC_MAX), kmax);
        }
    }
}

// Potential energy from the per-target potentials a PHI force pass left behind:
//   out[0] = Σ_{i<j} m_i m_j / r_ij = ½ Σ_i m_i (φ_i − self_i),   φ_i = phi_scale · Σ_splits partial_phi[s][i],
// where self_i is the j == i term of the pair loop (d² == ε² exactly), removed by subtracting the identical expression.
template <typename T, typename TM>
__global__ void __launch_bounds__(256) phi_energy_kernel(const double* __restrict__ partial_phi, int splits, int64_t n_tgt,
                                                         double phi_scale, bool uniform, const TM* __restrict__ mass, double eps_sq,
                                                         double* __restrict__ block_partials) {
    __shared__ double red[32];
    double self_unit;                      // 1 / r of the self pair as the pair loop evaluates it
    if constexpr (sizeof(T) == 4) self_unit = (double)rsqrt_approx((float)eps_sq);
    else self_unit = eps_sq * inv_dist_cubed(eps_sq);
    double mine = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_tgt; i += (int64_t)gridDim.x * blockDim.x) {
        double phi = 0.0;
        for (int sp = 0; sp < splits; ++sp) phi += partial_phi[(int64_t)sp * n_tgt + i];
        const double m = (double)mass[i];
        // uniform masses: the loop summed 1/r and phi_scale is the common mass; general: it summed m_j / r
        const double self = uniform ? self_unit : m * self_unit;
        mine = fma(m, (phi - self) * phi_scale, mine);
    }
    const double tot = block_reduce(mine, OpAdd(), 0.0, red);
    if (threadIdx.x == 0) block_partials[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(1024) phi_final_kernel(const double* __restrict__ partials, int count, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int e = threadIdx.x; e < count; e += blockDim.x) s += partials[e];
    s = block_reduce(s, OpAdd(), 0.0, red);
    if (threadIdx.x == 0) out[0] = 0.5 * s;
}

// ======================================================================================================
// Persistent whole-tick kernel for small systems (N <= ~16K): GalaxySimulation.run, simulation.py:145-158
// ======================================================================================================
// At N <= 10^4 a tick is a few microseconds of arithmetic; launched as two kernels per tick (even replayed from a CUDA
// graph) it costs 13-16 us, most of it launch gaps and kernel ramp.  Here ONE cooperative launch runs `ticks` ticks: all
// CTAs stay resident and alternate between
//   phase 1 (every thread, grid-stride over particle pairs): reduce the j-split partial sums of the previous force pass,
//           closing kick + opening kick + drift with separately rounded mul/add, write x, v, a and the packed records;
//   phase 2 (CTA = one task: a tile of targets x a range of source chunks): the pair loop of accel_kernel, same consumer,
//           same TMA ring, leaving partial sums;
// separated by grid-wide barriers.  The arithmetic per value is that of kdk_kernel / accel_kernel (integrate.cu), so a
// run is bit-identical to issuing the kernels one by one.  cudaLaunchCooperativeKernel refuses a grid that is not
// co-resident, so the barriers cannot deadlock; the host falls back to the graph replay when it refuses.
struct PersistentArgs {
    AccelArgs a;                 // chunks_per_split / partial / n_tgt ... of the in-kernel force passes
    float* x; float* v; float* acc; const void* mass; int mass_f64;
    int64_t n, n_units; float half_dt, dt; char* packed;
    int ticks, tiles, splits;    // tasks = tiles x splits (<= gridDim.x)
    double scale;                // G or G*m (uniform masses)
    const double* partial_in; int splits_in;     // partial sums left by the force pass that preceded the launch
};

template <int DIM>
__device__ __forceinline__ void persistent_kdk(const PersistentArgs& p, const double* partial, int splits) {
    const int64_t count = p.n * DIM;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t unit = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; unit < p.n_units; unit += stride) {
        float px[2][3], pm[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t i = unit * 2 + h;
            if (i < p.n) {
#pragma unroll
                for (int k = 0; k < DIM; ++k) {
                    const int64_t e = i * DIM + k;
                    double s = 0.0;
                    for (int sp = 0; sp < splits; ++sp) s += partial[(int64_t)sp * count + e];        // split order, as accel_finalize
                    const float a = (float)(s * p.scale);
                    p.acc[e] = a;
                    const float kick = __fmul_rn(a, p.half_dt);
                    float v = __fadd_rn(p.v[e], kick);                        // simulation.py:141 (closing kick)
                    v = __fadd_rn(v, kick);                                   // :132 (opening kick of the next tick)
                    p.v[e] = v;
                    const float x = __fadd_rn(p.x[e], __fmul_rn(v, p.dt));    // :135
                    p.x[e] = x;
                    px[h][k] = x;
                }
                pm[h] = p.mass_f64 ? (float)reinterpret_cast<const double*>(p.mass)[i] : reinterpret_cast<const float*>(p.mass)[i];
            } else {
#pragma unroll
                for (int k = 0; k < DIM; ++k) px[h][k] = kPadCoordF32;         // padding record: far away, mass 0
                pm[h] = 0.f;
            }
        }
        const int64_t chunk = unit / kChunkUnits;
        const int u = (int)(unit % kChunkUnits);
        char* base = p.packed + chunk * (int64_t)chunk_bytes(DIM);
        *reinterpret_cast<float4*>(base + u * 16) = make_float4(px[0][0], px[1][0], px[0][1], px[1][1]);
        if (DIM == 3) *reinterpret_cast<float4*>(base + kChunkABytes + u * 16) = make_float4(px[0][2], px[1][2], pm[0], pm[1]);
        else          *reinterpret_cast<float2*>(base + kChunkABytes + u * 8) = make_float2(pm[0], pm[1]);
    }
}

template <class Consumer>
__global__ void __launch_bounds__(Consumer::THREADS + 32, 2) persistent_ticks_kernel(const PersistentArgs p) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const bool is_consumer = threadIdx.x < Consumer::THREADS;
    const int tasks = p.tiles * p.splits;
    const AccelArgs& a = p.a;
    const int tile = (int)(blockIdx.x % p.tiles), split = (int)(blockIdx.x / p.tiles);
    const double* partial = p.partial_in;
    int splits = p.splits_in;
    for (int t = 0; t < p.ticks; ++t) {
        persistent_kdk<Consumer::DIM>(p, partial, splits);
        __threadfence();
        grid.sync();
        // the packed records were written through the generic proxy; the TMA engine reads them through the async proxy
        asm volatile("fence.proxy.async;" ::: "memory");
        if ((int)blockIdx.x < tasks) {
            Consumer cons;
            if (is_consumer) cons.init(a, nullptr, tile);
            const int64_t c0 = (int64_t)split * a.chunks_per_split;
            const int64_t c1 = min(a.n_chunks, c0 + (int64_t)a.chunks_per_split);
            stream_sources(a.src, c0, c1, cons);
            if (is_consumer) cons.store(a, tile, split);
        }
        __threadfence();
        grid.sync();
        partial = a.partial;
        splits = p.splits;
    }
}

// ======================================================================================================
// One-barrier ticks for the smallest systems (N <= 4096)
// ======================================================================================================
// persistent_ticks_kernel pays two grid barriers per tick and the start-up of the TMA ring in every force phase.  Here a
// tick has ONE barrier: CTA (tile, split) integrates the particles it needs ITSELF — its 256 targets and the 256·cps
// sources of its split — from the previous state (x, v ping-pong buffers; Σ_j sums in three rotating fp64 buffers), puts
// the source records into shared memory in the packed layout, runs the same pair loop as every other force kernel
// (ForceF32::chunk on shared memory), and adds its per-target sums to the fp64 buffer with atomics.  The integrator is
// redone by every CTA that needs a particle (≈ 2·cps+... particle updates per thread), which is cheaper than a second
// barrier; the CTAs of split 0 own the tile and write its new state.  fp64 atomic adds of fp32-valued chunk sums are exact
// (24-bit mantissas, tens of addends), so their order cannot change the result: runs are bit-identical to the kernels
// issued one by one.  Three sum buffers: tick t reads A, adds into B and clears C (read in tick t−1, idle now).
struct SmallTicksArgs {
    float* xv[2][2];             // [buffer][0 = x, 1 = v]; buffer 0 is the caller's state
    float* acc;                  // accelerations (float), written with the final state
    double* sums[3];             // rotating Σ_j buffers, n·DIM doubles each
    const double* partial_in; int splits_in;     // partial sums of the force pass that preceded the launch (tick 0 reads these)
    const void* mass; int mass_f64;
    int64_t n; int ticks, tiles, splits, cps; int64_t n_chunks;
    float half_dt, dt, eps_sq, neg_zero; double scale;
};

template <int DIM, bool UNI>
__global__ void __launch_bounds__(256, 1) small_ticks_kernel(const SmallTicksArgs p) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(128) unsigned char smem[];
    using Cons = ForceF32<DIM, Q_F32, 1, 256, UNI>;
    const int tile = (int)(blockIdx.x % p.tiles), split = (int)(blockIdx.x / p.tiles);
    const bool active = (int)blockIdx.x < p.tiles * p.splits;
    const bool owner = active && split == 0;
    const int64_t tgt = (int64_t)tile * 256 + threadIdx.x;
    const int64_t count = p.n * DIM;
    const int cb = chunk_bytes(DIM);
    const int64_t c0 = (int64_t)split * p.cps, c1 = min(p.n_chunks, c0 + (int64_t)p.cps);

    // state of particle i after this tick's kick(s) and drift, from the previous state and its Σ_j sums
    auto advance = [&](int64_t i, int t, const float* x, const float* v, const double* sums, float* xo, float* vo, float* ao) {
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            const int64_t e = i * DIM + k;
            double sacc;
            // (other SMs rewrite these buffers every second / third tick: read them past the non-coherent L1)
            if (t == 0) { sacc = 0.0; for (int sp = 0; sp < p.splits_in; ++sp) sacc += p.partial_in[(int64_t)sp * count + e]; }
            else sacc = __ldcg(sums + e);
            const float a = (float)(sacc * p.scale);
            const float kick = __fmul_rn(a, p.half_dt);
            float vn = __fadd_rn(__ldcg(v + e), kick);            // simulation.py:141 (closing kick of the previous tick)
            vn = __fadd_rn(vn, kick);                             // :132 (opening kick)
            xo[k] = __fadd_rn(__ldcg(x + e), __fmul_rn(vn, p.dt));   // :135
            vo[k] = vn;
            ao[k] = a;
        }
    };

    for (int t = 0; t < p.ticks; ++t) {
        const float* x = p.xv[t & 1][0];
        const float* v = p.xv[t & 1][1];
        float* xn = p.xv[(t + 1) & 1][0];
        float* vn = p.xv[(t + 1) & 1][1];
        const double* s_read = p.sums[t % 3];
        double* s_add = p.sums[(t + 1) % 3];
        double* s_clear = p.sums[(t + 2) % 3];
        if (active) {
            Cons cons;
            // --- own target: integrate, keep the new position in registers; the tile's owner CTA writes the new state
            float xt[3] = {0.f, 0.f, 0.f}, vt[3], at[3];
            const int64_t ti = tgt < p.n ? tgt : p.n - 1;
            advance(ti, t, x, v, s_read, xt, vt, at);
            if (owner && tgt < p.n) {
#pragma unroll
                for (int k = 0; k < DIM; ++k) { xn[tgt * DIM + k] = xt[k]; vn[tgt * DIM + k] = vt[k]; p.acc[tgt * DIM + k] = at[k]; s_clear[tgt * DIM + k] = 0.0; }
            }
            cons.nx[0] = make_float2(-xt[0], -xt[0]); cons.ny[0] = make_float2(-xt[1], -xt[1]); cons.nz[0] = make_float2(-xt[2], -xt[2]);
            cons.ax[0] = cons.ay[0] = cons.az[0] = make_float2(0.f, 0.f);
            cons.sx[0] = cons.sy[0] = cons.sz[0] = 0.0;
            cons.eps2 = make_float2(p.eps_sq, p.eps_sq);
            cons.neg_zero = p.neg_zero;
            // --- the sources of this split: integrate and write their packed records into shared memory
            for (int64_t c = c0; c < c1; ++c) {
                const int64_t sidx = c * 256 + threadIdx.x;
                float xs[3] = {kPadCoordF32, kPadCoordF32, kPadCoordF32}, vs[3], as[3], ms = 0.f;
                if (sidx < p.n) {
                    advance(sidx, t, x, v, s_read, xs, vs, as);
                    ms = p.mass_f64 ? (float)reinterpret_cast<const double*>(p.mass)[sidx] : reinterpret_cast<const float*>(p.mass)[sidx];
                }
                float* A = reinterpret_cast<float*>(smem + (c - c0) * cb);
                float* B = reinterpret_cast<float*>(smem + (c - c0) * cb + kChunkABytes);
                const int u = threadIdx.x >> 1, h = threadIdx.x & 1;
                A[u * 4 + h] = xs[0]; A[u * 4 + 2 + h] = xs[1];
                if (DIM == 3) { B[u * 4 + h] = xs[2]; B[u * 4 + 2 + h] = ms; }
                else B[u * 2 + h] = ms;
            }
            __syncthreads();
            for (int64_t c = c0; c < c1; ++c) cons.chunk(smem + (c - c0) * cb, c);
            if (tgt < p.n) {
                atomicAdd(&s_add[tgt * DIM + 0], cons.sx[0]);
                atomicAdd(&s_add[tgt * DIM + 1], cons.sy[0]);
                if (DIM == 3) atomicAdd(&s_add[tgt * DIM + 2], cons.sz[0]);
            }
        }
        __threadfence();
        grid.sync();
    }
    // the state of the last tick is in buffer (ticks & 1): bring it home if that is the scratch buffer (no CTA reads it any more)
    if ((p.ticks & 1) && owner && tgt < p.n) {
#pragma unroll
        for (int k = 0; k < DIM; ++k) { p.xv[0][0][tgt * DIM + k] = p.xv[1][0][tgt * DIM + k]; p.xv[0][1][tgt * DIM + k] = p.xv[1][1][tgt * DIM + k]; }
    }
}

// ---- host side -----------------------------------------------------------------------------------------
constexpr int kForceThreads = 256;
constexpr int kForceIPT = 2;
constexpr int kTargetsPerBlock = kForceThreads * kForceIPT;

constexpr int64_t kPhiBlockBytes = 16 * 1024;        // per-CTA partial sums of the potential reduction (phi_energy_kernel)

// Instrumentation hook (nb_profile_next_force): CUDA events recorded around the NEXT pair-kernel launch of this host
// thread, then forgotten.  bench.py uses it to time the dominant kernel inside nb_run_ticks without changing the path.
struct ForceProfile { void* start; void* stop; int launches; };       // stop is recorded after the `launches`-th launch
inline ForceProfile& force_profile() { static thread_local ForceProfile p{nullptr, nullptr, 0}; return p; }

template <class Consumer, int LUTKIND>
int launch_accel(const AccelArgs& a0, int64_t workspace_bytes, cudaStream_t st, int* splits_out, double** phi_out = nullptr) {
    AccelArgs a = a0;
    if (a.n_tgt <= 0 || a.n_chunks <= 0 || !a.partial) return NB_ERR_INVALID_ARGUMENT;
    // a windowed launch appends its split slots behind the ones earlier windows of the same evaluation wrote
    const int splits_before = a.splits_before;
    a.partial += (int64_t)splits_before * a.n_tgt * Consumer::DIM;
    // a PHI consumer keeps one more double per target and split (the potential), behind the acceleration partials
    const int64_t per_split = a.n_tgt * (Consumer::DIM + (Consumer::HAS_PHI ? 1 : 0)) * (int64_t)sizeof(double);
    int64_t max_by_ws = (workspace_bytes - (Consumer::HAS_PHI ? kPhiBlockBytes : 0)) / per_split - splits_before;
    const int cap = max_splits_for(a.n_tgt, Consumer::DIM) - splits_before;
    if (max_by_ws > cap) max_by_ws = cap;
    if (max_by_ws < 1) return NB_ERR_WORKSPACE_TOO_SMALL;
    if (a.max_splits > 0 && max_by_ws > a.max_splits) max_by_ws = a.max_splits;
    int smem = stream_smem_bytes(Consumer::DIM);
    if (LUTKIND == 1) {
        a.lut_rep = 8;                                        // as many copies as fit in 32 KB
        while (a.lut_rep > 1 && a.levels * a.lut_rep * 16 > 32 * 1024) a.lut_rep >>= 1;
        smem += (a.levels * a.lut_rep + 1) * 16;
    }
    if (LUTKIND == 2) {
        const int P = lut_pow2ceil(a.levels);
        smem += P * 128 + (P + 1 + 3) / 4 * 16;
    }
    auto kern = accel_kernel<Consumer, LUTKIND>;
    int ctas_per_sm = 1;
    const int frc = kernel_occupancy((const void*)kern, Consumer::THREADS + 32, smem, &ctas_per_sm);
    if (frc != NB_OK) return frc;
    const SplitPlan p = plan_splits(a.n_tgt, a.n_chunks, Consumer::THREADS * Consumer::TARGETS_PER_THREAD, ctas_per_sm, (int)max_by_ws);
    a.chunks_per_split = p.chunks_per_split;
    if (Consumer::HAS_PHI) a.partial_phi = a.partial + (int64_t)p.splits * a.n_tgt * Consumer::DIM;
    ForceProfile& prof = force_profile();
    if (prof.start) { cudaEventRecord((cudaEvent_t)prof.start, st); prof.start = nullptr; }
    kern<<<dim3(p.blocks_i, p.splits), Consumer::THREADS + 32, smem, st>>>(a);
    NB_CUDA_LAUNCH_CHECK();
    if (prof.stop && --prof.launches <= 0) { cudaEventRecord((cudaEvent_t)prof.stop, st); prof.stop = nullptr; }   // one-shot
    *splits_out = p.splits;
    if (phi_out) *phi_out = a.partial_phi;
    return NB_OK;
}

}  // namespace nb

using namespace nb;

#ifndef NB_TUNE_HARNESS
extern "C" int nb_accel_max_splits(int64_t n_targets, int dim) {
    return (n_targets <= 0 || (dim != 2 && dim != 3)) ? 0 : max_splits_for(n_targets, dim);
}

extern "C" int64_t nb_accel_workspace_bytes(int64_t n_targets, int dim) {
    if (n_targets <= 0 || (dim != 2 && dim != 3)) return 0;
    // room for the largest j-split the planner may choose for this many targets (+ one potential per target and split
    // and the block partials of its reduction, for a pass that also returns the potential energy)
    return (int64_t)max_splits_for(n_targets, dim) * n_targets * (dim + 1) * (int64_t)sizeof(double) + kPhiBlockBytes;
}

int nb::accel_pairs(const void* packed_src, int64_t n_src, const void* pos_tgt, int64_t n_tgt, int dim, int dtype, int mode,
                    double G, double eps_sq, const void* level_table, int levels, int uniform_mass, double mass_value,
                    int64_t* scalars, void* workspace, int64_t workspace_bytes, cudaStream_t st, PartialSums* out, bool want_phi,
                    const SourceWindow* window) {
    if (!packed_src || !pos_tgt || !workspace || n_src <= 0 || n_tgt <= 0 || (dim != 2 && dim != 3))
        return NB_ERR_INVALID_ARGUMENT;
    if (dtype != NB_F32 && dtype != NB_F64) return NB_ERR_INVALID_ARGUMENT;
    if (mode < NB_MODE_FLOAT64 || mode > NB_MODE_CUSTOM) return NB_ERR_INVALID_ARGUMENT;
    const bool lut = mode == NB_MODE_INT8_SIM || mode == NB_MODE_INT4_SIM || mode == NB_MODE_CUSTOM;
    if (lut && (!level_table || levels < 2 || !scalars)) return NB_ERR_INVALID_ARGUMENT;
    if (lut && levels > kMaxLevelsSmem) return NB_ERR_UNSUPPORTED;
    if (lut && dtype == NB_F64) return NB_ERR_UNSUPPORTED;       // fp64 state + int modes: no caller in the reference
    // the fused potential exists where the pair loop sees the unquantised d² in the state dtype (simulation.py:181-186)
    if (want_phi && !((dtype == NB_F32 && mode == NB_MODE_FLOAT32) || (dtype == NB_F64 && mode == NB_MODE_FLOAT64)))
        return NB_ERR_UNSUPPORTED;

    AccelArgs a{};
    a.src = (const char*)packed_src;
    a.n_chunks = nb_num_chunks(n_src, dtype);
    if (window) {
        // a contiguous window [first, first + count) of the packed set: the SAME kernels, started at an offset source pointer
        // (one kernel image for single- and multi-GPU runs: the packed-FMA loop's ptxas schedule is worth several per cent
        // and differs between otherwise equivalent instantiations — profiles/r02/README.md)
        if (lut || want_phi || window->first_chunk < 0 || window->n_chunks <= 0 || window->splits_before < 0 ||
            window->first_chunk + window->n_chunks > a.n_chunks)
            return NB_ERR_INVALID_ARGUMENT;
        if (!((dtype == NB_F32 && mode == NB_MODE_FLOAT32) || (dtype == NB_F64 && mode == NB_MODE_FLOAT64))) return NB_ERR_UNSUPPORTED;
        a.src += window->first_chunk * (int64_t)chunk_bytes(dim);
        a.n_chunks = window->n_chunks;
        a.splits_before = window->splits_before;
        a.max_splits = window->max_splits;
    }
    a.pos_tgt = pos_tgt;
    a.n_tgt = n_tgt;
    a.partial = (double*)workspace;
    a.eps_sq = eps_sq;
    a.table = level_table;
    a.levels = levels;
    a.neg_zero = -0.0f;
    a.uniform_mass = (lut && levels <= kLutFastMaxLevels && uniform_mass != 0 && mass_value != 0.0) ? (float)mass_value : 0.f;

    int splits = 0, rc = NB_ERR_INVALID_ARGUMENT;
    double* phi = nullptr;
    constexpr int TH = kForceThreads, IPT = kForceIPT;
    // uniform-mass fast path: fp32 state in FLOAT32 / FLOAT16 / BFLOAT16 mode, fp64 state in FLOAT64 mode
    const bool uni = uniform_mass != 0 && ((dtype == NB_F32 && (mode == NB_MODE_FLOAT32 || mode == NB_MODE_FLOAT16 || mode == NB_MODE_BFLOAT16)) ||
                                           (dtype == NB_F64 && mode == NB_MODE_FLOAT64));
#define NB_F32_UNI_CASE(D, Q) rc = launch_accel<ForceF32<D, Q, IPT, TH, true>, 0>(a, workspace_bytes, st, &splits)
#define NB_F32_CASE(D, Q, LUT) rc = launch_accel<ForceF32<D, Q, IPT, TH>, LUT>(a, workspace_bytes, st, &splits)
#define NB_F64_CASE(D, Q) rc = launch_accel<ForceF64<D, Q, IPT, TH>, 0>(a, workspace_bytes, st, &splits)
#define NB_PHI_CASE(F, D, Q, U, UNR) rc = launch_accel<F<D, Q, IPT, TH, U, UNR, true>, 0>(a, workspace_bytes, st, &splits, &phi)
    if (want_phi) {
        if (dtype == NB_F32) {
            if (uni) { if (dim == 2) NB_PHI_CASE(ForceF32, 2, Q_F32, true, 4); else NB_PHI_CASE(ForceF32, 3, Q_F32, true, 4); }
            else { if (dim == 2) NB_PHI_CASE(ForceF32, 2, Q_F32, false, 4); else NB_PHI_CASE(ForceF32, 3, Q_F32, false, 4); }
        } else {
            if (uni) { if (dim == 2) NB_PHI_CASE(ForceF64, 2, Q_F64, true, 2); else NB_PHI_CASE(ForceF64, 3, Q_F64, true, 2); }
            else { if (dim == 2) NB_PHI_CASE(ForceF64, 2, Q_F64, false, 2); else NB_PHI_CASE(ForceF64, 3, Q_F64, false, 2); }
        }
    } else
#undef NB_PHI_CASE
    if (dtype == NB_F32) {
        if (mode == NB_MODE_FLOAT64) {
            if (dim == 2) rc = launch_accel<ForceMixed<2, IPT, TH>, 0>(a, workspace_bytes, st, &splits);
            else rc = launch_accel<ForceMixed<3, IPT, TH>, 0>(a, workspace_bytes, st, &splits);
        } else if (mode == NB_MODE_FLOAT32 && uni) { if (dim == 2) NB_F32_UNI_CASE(2, Q_F32); else NB_F32_UNI_CASE(3, Q_F32); }
        else if (mode == NB_MODE_FLOAT16 && uni) { if (dim == 2) NB_F32_UNI_CASE(2, Q_F16); else NB_F32_UNI_CASE(3, Q_F16); }
        else if (mode == NB_MODE_BFLOAT16 && uni) { if (dim == 2) NB_F32_UNI_CASE(2, Q_BF16); else NB_F32_UNI_CASE(3, Q_BF16); }
        else if (mode == NB_MODE_FLOAT32) { if (dim == 2) NB_F32_CASE(2, Q_F32, 0); else NB_F32_CASE(3, Q_F32, 0); }
        else if (mode == NB_MODE_FLOAT16) { if (dim == 2) NB_F32_CASE(2, Q_F16, 0); else NB_F32_CASE(3, Q_F16, 0); }
        else if (mode == NB_MODE_BFLOAT16) { if (dim == 2) NB_F32_CASE(2, Q_BF16, 0); else NB_F32_CASE(3, Q_BF16, 0); }
        else if (levels <= kLutFastMaxLevels) {
            if (dim == 2) rc = launch_accel<ForceF32<2, Q_LUTF, NB_LUTF_IPT, NB_LUTF_THREADS, false, NB_LUTF_UNROLL>, 2>(a, workspace_bytes, st, &splits);
            else rc = launch_accel<ForceF32<3, Q_LUTF, NB_LUTF_IPT, NB_LUTF_THREADS, false, NB_LUTF_UNROLL>, 2>(a, workspace_bytes, st, &splits);
        }
        else { if (dim == 2) NB_F32_CASE(2, Q_LUT, 1); else NB_F32_CASE(3, Q_LUT, 1); }
    } else {
        if (mode == NB_MODE_FLOAT64 && uni) {
            if (dim == 2) rc = launch_accel<ForceF64<2, Q_F64, IPT, TH, true>, 0>(a, workspace_bytes, st, &splits);
            else rc = launch_accel<ForceF64<3, Q_F64, IPT, TH, true>, 0>(a, workspace_bytes, st, &splits);
        } else if (mode == NB_MODE_FLOAT64) { if (dim == 2) NB_F64_CASE(2, Q_F64); else NB_F64_CASE(3, Q_F64); }
        else if (mode == NB_MODE_FLOAT32) { if (dim == 2) NB_F64_CASE(2, Q_F32); else NB_F64_CASE(3, Q_F32); }
        else if (mode == NB_MODE_FLOAT16) { if (dim == 2) NB_F64_CASE(2, Q_F16); else NB_F64_CASE(3, Q_F16); }
        else if (mode == NB_MODE_BFLOAT16) { if (dim == 2) NB_F64_CASE(2, Q_BF16); else NB_F64_CASE(3, Q_BF16); }
    }
#undef NB_F32_CASE
#undef NB_F32_UNI_CASE
#undef NB_F64_CASE
    if (rc != NB_OK) return rc;
    // what the reduction needs: Σ splits, ×G (float modes: G was hoisted out of the pair loop; LUT factors carry G)
    out->partial = a.partial;
    out->splits = splits + a.splits_before;
    out->count = n_tgt * dim;
    out->scale = lut ? (a.uniform_mass != 0.f ? (double)a.uniform_mass : 1.0) : (uni ? G * mass_value : G);
    out->out_f64 = dtype == NB_F64 || mode == NB_MODE_FLOAT64;
    out->minmax = mode == NB_MODE_INT8_SIM || mode == NB_MODE_INT4_SIM;
    out->partial_phi = phi;
    out->phi_scale = uni ? mass_value : 1.0;
    out->phi_uniform = uni;
    out->n_tgt = n_tgt;
    return NB_OK;
}

int nb::potential_from_phi(const PartialSums& p, int dtype, const void* mass_tgt, int mass_dtype, double eps_sq, double* out,
                           cudaStream_t st) {
    if (!p.partial_phi || !mass_tgt || !out) return NB_ERR_INVALID_ARGUMENT;
    // block partials live behind the potentials (launch_accel reserved kPhiBlockBytes there)
    double* blocks = const_cast<double*>(p.partial_phi) + (int64_t)p.splits * p.n_tgt;
    int64_t nb = (p.n_tgt + 255) / 256;
    const int64_t cap = kPhiBlockBytes / (int64_t)sizeof(double);
    if (nb > cap) nb = cap;
    if (nb > kNumSMsB200 * 8) nb = kNumSMsB200 * 8;
#define NB_PHI_E(T, TM) phi_energy_kernel<T, TM><<<(int)nb, 256, 0, st>>>(p.partial_phi, p.splits, p.n_tgt, p.phi_scale, p.phi_uniform, (const TM*)mass_tgt, eps_sq, blocks)
    if (dtype == NB_F32 && mass_dtype == NB_F32) NB_PHI_E(float, float);
    else if (dtype == NB_F32 && mass_dtype == NB_F64) NB_PHI_E(float, double);
    else if (dtype == NB_F64 && mass_dtype == NB_F32) NB_PHI_E(double, float);
    else if (dtype == NB_F64 && mass_dtype == NB_F64) NB_PHI_E(double, double);
    else return NB_ERR_INVALID_ARGUMENT;
#undef NB_PHI_E
    NB_CUDA_LAUNCH_CHECK();
    phi_final_kernel<<<1, 1024, 0, st>>>(blocks, (int)nb, out);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

// `ticks` steady-state ticks (closing kick of the previous tick fused with kick-drift, then the force) in ONE cooperative
// launch; `ps` describes the partial sums of the force pass before the launch on entry and those of the last in-kernel
// pass on return.  NB_ERR_UNSUPPORTED when the configuration is outside the kernel's scope or the grid cannot be
// co-resident — the caller then replays the tick body from a CUDA graph instead.
template <class Consumer>
static int launch_persistent(PersistentArgs& p, int64_t workspace_bytes, cudaStream_t st, int* splits_out) {
    auto kern = persistent_ticks_kernel<Consumer>;
    const int threads = Consumer::THREADS + 32;
    const int smem = stream_smem_bytes(Consumer::DIM);
    int occ = 0;
    const int frc = kernel_occupancy((const void*)kern, threads, smem, &occ);
    if (frc != NB_OK) return frc;
    const int cap = device_sm_count() * occ;
    const int tpb = Consumer::THREADS * Consumer::TARGETS_PER_THREAD;
    const int tiles = (int)((p.n + tpb - 1) / tpb);
    if (tiles > cap) return NB_ERR_UNSUPPORTED;
    int64_t max_splits = cap / tiles;
    const int64_t per_split = p.n * Consumer::DIM * (int64_t)sizeof(double);
    if (max_splits > workspace_bytes / per_split) max_splits = workspace_bytes / per_split;
    if (max_splits > max_splits_for(p.n, Consumer::DIM)) max_splits = max_splits_for(p.n, Consumer::DIM);
    if (max_splits > p.a.n_chunks) max_splits = p.a.n_chunks;
    if (max_splits < 1) return NB_ERR_WORKSPACE_TOO_SMALL;
    const int cps = (int)((p.a.n_chunks + max_splits - 1) / max_splits);
    const int splits = (int)((p.a.n_chunks + cps - 1) / cps);
    p.a.chunks_per_split = cps;
    p.tiles = tiles;
    p.splits = splits;
    // every SM takes part in the elementwise phase even when there are fewer tasks than SMs
    int grid = tiles * splits;
    if (grid < device_sm_count()) grid = device_sm_count();
    void* params[] = {(void*)&p};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(threads), params, (size_t)smem, st);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported || e == cudaErrorLaunchOutOfResources) {
        cudaGetLastError();
        return NB_ERR_UNSUPPORTED;
    }
    if (e != cudaSuccess) return cuda_status(e);
    *splits_out = splits;
    return NB_OK;
}

// One-barrier ticks (small_ticks_kernel): same contract as persistent_ticks below; N <= 4096.
template <int DIM, bool UNI>
static int launch_small_ticks(SmallTicksArgs& p, cudaStream_t st) {
    auto kern = small_ticks_kernel<DIM, UNI>;
    const int smem = p.cps * chunk_bytes(DIM);
    int occ = 0;
    const int frc = kernel_occupancy((const void*)kern, 256, smem, &occ);
    if (frc != NB_OK) return frc;
    const int grid = p.tiles * p.splits;
    if (grid > device_sm_count() * occ) return NB_ERR_UNSUPPORTED;
    void* params[] = {(void*)&p};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(256), params, (size_t)smem, st);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported || e == cudaErrorLaunchOutOfResources) {
        cudaGetLastError();
        return NB_ERR_UNSUPPORTED;
    }
    return e == cudaSuccess ? NB_OK : cuda_status(e);
}

static int small_ticks(void* x, void* v, void* acc, const void* mass, int mass_dtype, int64_t n, int dim, double G, double eps_sq,
                       double dt, int uniform_mass, double mass_value, void* workspace, int64_t workspace_bytes, int64_t ticks,
                       PartialSums* ps, cudaStream_t st) {
    if (n > 4096 || ticks > 0x7fffffff) return NB_ERR_UNSUPPORTED;
    const bool uni = uniform_mass != 0;
    const int64_t count = n * dim;
    SmallTicksArgs p{};
    p.n = n; p.ticks = (int)ticks;
    p.n_chunks = nb_num_chunks(n, NB_F32);
    p.tiles = (int)((n + 255) / 256);
    int max_splits = device_sm_count() / p.tiles;                     // one CTA per SM
    if (max_splits < 1) return NB_ERR_UNSUPPORTED;
    if (max_splits > p.n_chunks) max_splits = (int)p.n_chunks;
    p.cps = (int)((p.n_chunks + max_splits - 1) / max_splits);
    p.splits = (int)((p.n_chunks + p.cps - 1) / p.cps);
    // scratch behind the partial sums the launch starts from: three fp64 sum buffers + the second x, v buffer
    const int64_t in_bytes = (int64_t)ps->splits * count * (int64_t)sizeof(double);
    const int64_t need = in_bytes + 3 * count * (int64_t)sizeof(double) + 2 * count * (int64_t)sizeof(float) + 256;
    if (workspace_bytes < need) return NB_ERR_UNSUPPORTED;
    char* base = (char*)workspace + ((in_bytes + 127) / 128) * 128;
    for (int b = 0; b < 3; ++b) p.sums[b] = (double*)(base + b * count * (int64_t)sizeof(double));
    float* x1 = (float*)(base + 3 * count * (int64_t)sizeof(double));
    p.xv[0][0] = (float*)x; p.xv[0][1] = (float*)v; p.xv[1][0] = x1; p.xv[1][1] = x1 + count;
    p.acc = (float*)acc;
    p.partial_in = ps->partial; p.splits_in = ps->splits;
    p.mass = mass; p.mass_f64 = mass_dtype == NB_F64;
    p.half_dt = (float)(dt / 2); p.dt = (float)dt; p.eps_sq = (float)eps_sq; p.neg_zero = -0.0f;
    p.scale = uni ? G * mass_value : G;
    if (ps->scale != p.scale || ps->count != count) return NB_ERR_UNSUPPORTED;
    // tick 0 adds into sums[1], tick 1 into sums[2] (cleared by tick 0's owners), ...: only sums[1] must start at zero
    cudaError_t e = cudaMemsetAsync(p.sums[1], 0, count * sizeof(double), st);
    if (e != cudaSuccess) return cuda_status(e);
    int rc;
    if (dim == 2) rc = uni ? launch_small_ticks<2, true>(p, st) : launch_small_ticks<2, false>(p, st);
    else rc = uni ? launch_small_ticks<3, true>(p, st) : launch_small_ticks<3, false>(p, st);
    if (rc != NB_OK) return rc;
    // the sums of the last force pass: one "split"; the packed records of the final positions exist only in shared memory
    ps->partial = p.sums[ticks % 3];
    ps->splits = 1;
    ps->packed_stale = true;
    return NB_OK;
}

int nb::persistent_ticks(void* x, void* v, void* acc, const void* mass, int mass_dtype, int64_t n, int dim, int dtype, int mode, double G,
                         double eps_sq, double dt, int uniform_mass, double mass_value, void* packed, void* workspace,
                         int64_t workspace_bytes, int64_t ticks, PartialSums* ps, cudaStream_t st) {
    if (dtype != NB_F32 || mode != NB_MODE_FLOAT32 || ticks < 1 || ticks > 0x7fffffff || n > 32768 || !ps || !ps->partial ||
        ps->out_f64 || ps->minmax)
        return NB_ERR_UNSUPPORTED;
    static const bool one_barrier = [] { const char* e = getenv("NB_B200_ONE_BARRIER"); return !e || atoi(e) != 0; }();
    if (one_barrier && n <= 4096) {
        const int rc = small_ticks(x, v, acc, mass, mass_dtype, n, dim, G, eps_sq, dt, uniform_mass, mass_value, workspace, workspace_bytes,
                                   ticks, ps, st);
        if (rc != NB_ERR_UNSUPPORTED) return rc;
    }
    const bool uni = uniform_mass != 0;
    PersistentArgs p{};
    p.a.src = (const char*)packed;
    p.a.n_chunks = nb_num_chunks(n, dtype);
    p.a.pos_tgt = x;
    p.a.n_tgt = n;
    p.a.partial = (double*)workspace;
    p.a.eps_sq = eps_sq;
    p.a.neg_zero = -0.0f;
    p.x = (float*)x; p.v = (float*)v; p.acc = (float*)acc; p.mass = mass; p.mass_f64 = mass_dtype == NB_F64;
    p.n = n; p.n_units = p.a.n_chunks * kChunkUnits;
    p.half_dt = (float)(dt / 2); p.dt = (float)dt;
    p.packed = (char*)packed;
    p.ticks = (int)ticks;
    p.scale = uni ? G * mass_value : G;
    p.partial_in = ps->partial; p.splits_in = ps->splits;
    if (ps->scale != p.scale || ps->count != n * dim) return NB_ERR_UNSUPPORTED;
    int splits = 0, rc;
    // few targets: one target per thread doubles the number of tiles (shorter tasks, more SMs busy)
    const bool small = (n + 511) / 512 * p.a.n_chunks <= device_sm_count();
#define NB_PERSIST(D, U, I) rc = launch_persistent<ForceF32<D, Q_F32, I, kForceThreads, U>>(p, workspace_bytes, st, &splits)
    if (dim == 2) { if (uni) { if (small) NB_PERSIST(2, true, 1); else NB_PERSIST(2, true, 2); } else { if (small) NB_PERSIST(2, false, 1); else NB_PERSIST(2, false, 2); } }
    else          { if (uni) { if (small) NB_PERSIST(3, true, 1); else NB_PERSIST(3, true, 2); } else { if (small) NB_PERSIST(3, false, 1); else NB_PERSIST(3, false, 2); } }
#undef NB_PERSIST
    if (rc != NB_OK) return rc;
    ps->partial = (const double*)workspace;
    ps->splits = splits;
    return NB_OK;
}

extern "C" int nb_profile_next_force(void* start_event, void* stop_event, int launches) {
    ForceProfile& p = force_profile();
    p.start = start_event;
    p.stop = stop_event;
    p.launches = launches < 1 ? 1 : launches;
    return NB_OK;
}

int nb::accel_reduce(const PartialSums& p, void* acc_out, int64_t* scalars, cudaStream_t st) {
    if (!acc_out) return NB_ERR_INVALID_ARGUMENT;
    int64_t blocks = (p.count + 255) / 256;
    if (blocks > kNumSMsB200 * 8) blocks = kNumSMsB200 * 8;
    if (p.out_f64) accel_finalize_kernel<double, false><<<(int)blocks, 256, 0, st>>>(p.partial, p.splits, p.count, p.scale, (double*)acc_out, scalars);
    else if (p.minmax) accel_finalize_kernel<float, true><<<(int)blocks, 256, 0, st>>>(p.partial, p.splits, p.count, p.scale, (float*)acc_out, scalars);
    else accel_finalize_kernel<float, false><<<(int)blocks, 256, 0, st>>>(p.partial, p.splits, p.count, p.scale, (float*)acc_out, scalars);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_accel(const void* packed_src, int64_t n_src, const void* pos_tgt, int64_t n_tgt, int dim, int dtype,
                        int mode, double G, double eps_sq, const void* level_table, int levels, int uniform_mass,
                        double mass_value, void* acc_out,
                        int64_t* scalars, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!acc_out) return NB_ERR_INVALID_ARGUMENT;
    PartialSums p{};
    const int rc = accel_pairs(packed_src, n_src, pos_tgt, n_tgt, dim, dtype, mode, G, eps_sq, level_table, levels, uniform_mass,
                               mass_value, scalars, workspace, workspace_bytes, (cudaStream_t)stream, &p, false, nullptr);
    if (rc != NB_OK) return rc;
    return accel_reduce(p, acc_out, scalars, (cudaStream_t)stream);
}

// Windowed evaluation (i-range-sharded ticks): the window of the rank's OWN packed slot runs while the all-gather of the
// other slots is still in flight, then the windows over the slots after and before it; nb_accel_finish reduces them all.
extern "C" int nb_accel_window(const void* packed_src, int64_t n_src, int64_t first_chunk, int64_t n_chunks,
                               const void* pos_tgt, int64_t n_tgt, int dim, int dtype, int mode, double G, double eps_sq,
                               int uniform_mass, double mass_value, void* workspace, int64_t workspace_bytes, int splits_before,
                               int max_splits, int* splits_total_out, void* stream) {
    if (!splits_total_out) return NB_ERR_INVALID_ARGUMENT;
    PartialSums p{};
    const SourceWindow w{first_chunk, n_chunks, splits_before, max_splits};
    const int rc = accel_pairs(packed_src, n_src, pos_tgt, n_tgt, dim, dtype, mode, G, eps_sq, nullptr, 0, uniform_mass, mass_value,
                               nullptr, workspace, workspace_bytes, (cudaStream_t)stream, &p, false, &w);
    if (rc != NB_OK) return rc;
    *splits_total_out = p.splits;
    return NB_OK;
}

extern "C" int nb_accel_finish(const void* workspace, int splits_total, int64_t n_tgt, int dim, int dtype, int mode, double G,
                               int uniform_mass, double mass_value, void* acc_out, void* stream) {
    if (!workspace || !acc_out || splits_total < 1 || n_tgt <= 0 || (dim != 2 && dim != 3)) return NB_ERR_INVALID_ARGUMENT;
    if (!((dtype == NB_F32 && mode == NB_MODE_FLOAT32) || (dtype == NB_F64 && mode == NB_MODE_FLOAT64))) return NB_ERR_UNSUPPORTED;
    const bool uni = uniform_mass != 0;
    PartialSums p{};
    p.partial = (const double*)workspace;
    p.splits = splits_total;
    p.count = n_tgt * dim;
    p.scale = uni ? G * mass_value : G;
    p.out_f64 = dtype == NB_F64 || mode == NB_MODE_FLOAT64;
    p.minmax = false;
    return accel_reduce(p, acc_out, nullptr, (cudaStream_t)stream);
}

extern "C" int nb_accel_potential(const void* packed_src, int64_t n_src, const void* pos_tgt, const void* mass_tgt, int64_t n_tgt,
                                  int dim, int dtype, int mass_dtype, int mode, double G, double eps_sq, int uniform_mass,
                                  double mass_value, void* acc_out, double* pe_out, void* workspace, int64_t workspace_bytes,
                                  void* stream) {
    if (!acc_out || !pe_out || !mass_tgt) return NB_ERR_INVALID_ARGUMENT;
    PartialSums p{};
    int rc = accel_pairs(packed_src, n_src, pos_tgt, n_tgt, dim, dtype, mode, G, eps_sq, nullptr, 0, uniform_mass, mass_value,
                         nullptr, workspace, workspace_bytes, (cudaStream_t)stream, &p, true, nullptr);
    if (rc != NB_OK) return rc;
    if ((rc = potential_from_phi(p, dtype, mass_tgt, mass_dtype, eps_sq, pe_out, (cudaStream_t)stream))) return rc;
    return accel_reduce(p, acc_out, nullptr, (cudaStream_t)stream);
}
#endif  // NB_TUNE_HARNESS
