// Source-streaming skeleton shared by every O(N²) kernel (force, max-d², potential energy).
//
// A CTA owns a block of target particles (registers) and a contiguous range of packed source chunks.
// One extra producer warp feeds a ring of shared-memory stages with TMA bulk copies (cp.async.bulk +
// mbarrier complete_tx); the consumer warps walk each stage chunk by chunk with warp-broadcast
// LDS.128 reads and release it through an "empty" mbarrier.  Warps may drift apart by up to
// kStages-1 stages, so there is no CTA-wide barrier in the steady state.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <mutex>
#include "common.cuh"

namespace nb {

constexpr int kStages = 3;
constexpr int kStageChunks = 4;                 // 16 KB (D=3) / 12 KB (D=2) per stage
constexpr int kBarrierBytes = 128;              // full[kStages] + empty[kStages] mbarriers, padded

__host__ __device__ inline int stream_smem_bytes(int dim) { return kBarrierBytes + kStages * kStageChunks * chunk_bytes(dim); }

// ---- j-split planning (host) ------------------------------------------------------------------------
// grid = (blocks_i target blocks) x (splits source ranges).  All CTAs of a launch cost about the same, so
// the makespan is ceil(CTAs / resident slots) waves: pick the split count that wastes the least of the
// last wave, charging a small fixed cost per CTA (pipeline fill, target load, partial-sum store).
struct SplitPlan { int blocks_i; int splits; int chunks_per_split; };

inline int device_sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = kNumSMsB200;
        cudaGetLastError();
    }
    return sms;
}

inline SplitPlan plan_splits(int64_t n_tgt, int64_t n_chunks, int targets_per_block, int ctas_per_sm, int max_splits) {
    SplitPlan best{};
    best.blocks_i = (int)((n_tgt + targets_per_block - 1) / targets_per_block);
    const double slots = (double)device_sm_count() * (ctas_per_sm > 0 ? ctas_per_sm : 1);
    if (max_splits > 65535) max_splits = 65535;
    double best_eff = -1.0;
    for (int64_t s = 1; s <= max_splits && s <= n_chunks; ++s) {
        const int64_t cps = (n_chunks + s - 1) / s;
        const int64_t splits = (n_chunks + cps - 1) / cps;
        if (splits != s) continue;                              // same plan as a smaller s
        const double ctas = (double)best.blocks_i * (double)splits;
        const double waves = ceil(ctas / slots);
        // useful work / capacity of the waves, each CTA paying ~0.2 chunk-equivalents of fixed cost
        const double eff = ((double)best.blocks_i * (double)n_chunks) / (waves * slots * ((double)cps + 0.2));
        // (a further split costs one more fp64 partial-sum slot — tens of MB written and read once, microseconds — which the
        // fixed cost above already over-charges; r01 demanded a 1 % gain here and left 0.9 % on the table at N = 2^20)
        if (eff > best_eff * 1.0005) { best_eff = eff; best.splits = (int)splits; best.chunks_per_split = (int)cps; }
    }
    if (best_eff < 0) { best.splits = 1; best.chunks_per_split = (int)n_chunks; }
    // The wave model is blind to the ragged END of a launch (the SMs do not finish together: about 0.2 CTA lifetimes of
    // idle time, fitted on the fp32 and fast-lookup kernels).  That matters exactly when it picked ONE split for a multi-wave
    // grid — every CTA then streams the whole source set (84 ms at N = 2^20 in the fast-lookup kernel, whose 2048 x s CTAs fill
    // 296 slots equally badly for every s): take the largest split count the model rates within 0.1 % of its choice whose CTAs
    // still stream >= 256 chunks.  Measured: 585.0 -> 571.7 ms (INT8_SIM, D = 3, N = 2^20; profiles/r02/int_splits_ab.log).
    if (best.splits == 1 && (double)best.blocks_i > slots) {
        for (int64_t s = max_splits < n_chunks ? max_splits : n_chunks; s > 1; --s) {
            const int64_t cps = (n_chunks + s - 1) / s;
            const int64_t splits = (n_chunks + cps - 1) / cps;
            if (splits != s || cps < 256) continue;
            const double waves = ceil((double)best.blocks_i * (double)splits / slots);
            const double eff = ((double)best.blocks_i * (double)n_chunks) / (waves * slots * ((double)cps + 0.2));
            if (eff >= best_eff * 0.999) { best.splits = (int)splits; best.chunks_per_split = (int)cps; break; }
        }
    }
    // developer override for A/B timing of the split count (tools/time_splits.py); never set in production
    static const int forced = [] { const char* e = getenv("NB_B200_SPLITS"); return e ? atoi(e) : 0; }();
    if (forced > 0 && forced <= max_splits && forced <= n_chunks) {
        best.chunks_per_split = (int)((n_chunks + forced - 1) / forced);
        best.splits = (int)((n_chunks + best.chunks_per_split - 1) / best.chunks_per_split);
    }
    return best;
}

// upper bound on the split count used for sizing workspaces (<= 256 MiB of partial sums, <= 32 splits;
// NB_B200_SPLIT_WORKSPACE_MB / NB_B200_SPLIT_CAP: developer overrides for A/B timing, never set in production)
inline int max_splits_for(int64_t n_tgt, int dim) {
    static const int64_t mb = [] { const char* e = getenv("NB_B200_SPLIT_WORKSPACE_MB"); return e && atoi(e) > 0 ? (int64_t)atoi(e) : (int64_t)256; }();
    static const int64_t hard = [] { const char* e = getenv("NB_B200_SPLIT_CAP"); return e && atoi(e) > 0 ? (int64_t)atoi(e) : (int64_t)32; }();
    const int64_t per_split = n_tgt * dim * (int64_t)sizeof(double);
    int64_t s = (mb << 20) / (per_split > 0 ? per_split : 1);
    if (s > hard) s = hard;
    if (s < 1) s = 1;
    return (int)s;
}

// ---- per-kernel launch facts (host) -----------------------------------------------------------------
// Dynamic-smem opt-in and resident CTAs/SM are immutable properties of a kernel image on a device; they are
// looked up once per (kernel, device, smem) and then served from a small table, because the tick loop launches
// the same kernels thousands of times (and may be under stream capture, where attribute calls are best avoided).
struct KernelFacts { const void* fn; int dev; int smem; int threads; int occ; };

inline int kernel_occupancy(const void* fn, int threads, int smem, int* occ_out) {
    static std::mutex mu;
    static KernelFacts table[96];
    static int used = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    int opted_in = -1;                                  // largest dynamic-smem size already enabled for this kernel
    for (int i = 0; i < used; ++i) {
        if (table[i].fn != fn || table[i].dev != dev) continue;
        if (table[i].smem > opted_in) opted_in = table[i].smem;
        if (table[i].smem == smem && table[i].threads == threads) {
            *occ_out = table[i].occ;
            return NB_OK;
        }
    }
    if (smem > opted_in) {                              // only ever raise the opt-in (level tables vary in size)
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return cuda_status(e);
    }
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, smem);
    if (e != cudaSuccess) return cuda_status(e);
    if (occ < 1) occ = 1;
    if (used < 96) table[used++] = KernelFacts{fn, dev, smem, threads, occ};
    else if (smem > opted_in) table[95] = KernelFacts{fn, dev, smem, threads, occ};
    *occ_out = occ;
    return NB_OK;
}

// Consumer concept:
//   static constexpr int DIM, THREADS (consumer threads; the CTA has THREADS + 32);
//   __device__ void chunk(const unsigned char* smem_chunk, int64_t chunk_index);   // all consumer threads
// `ring` > 0: chunk indices are taken modulo `ring` (the range [c0, c1) may run past the last chunk and continue at
// chunk 0 — the half-ring partition of the potential-energy pairs); a stage that straddles the wrap point is filled
// by two bulk copies counted on the same barrier.
template <class Consumer>
__device__ __forceinline__ void stream_sources(const char* __restrict__ src, int64_t c0, int64_t c1, Consumer& cons,
                                               int64_t ring = 0) {
    constexpr int DIM = Consumer::DIM;
    constexpr int NCW = Consumer::THREADS / 32;     // consumer warps
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t bar0 = smem_u32(smem);
    unsigned char* data = smem + kBarrierBytes;
    const int cb = chunk_bytes(DIM);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = (int)(c1 - c0);
    const int n_iters = (n_chunks + kStageChunks - 1) / kStageChunks;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar0 + 8 * s, 1);                  // full: one arrive (+tx bytes) by the producer
            mbar_init(bar0 + 8 * (kStages + s), NCW);    // empty: one arrive per consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == NCW) {
        // ---------------- producer warp: one elected lane issues the bulk copies ----------------
        if (lane == 0) {
            for (int it = 0; it < n_iters; ++it) {
                const int s = it % kStages;
                if (it >= kStages) {
                    // the producer has ~a stage of compute time to spare: back off instead of spinning on the
                    // issue port of its SMSP (ncu: the bare try_wait loop was 19 % of all executed instructions)
                    while (!mbar_try_wait(bar0 + 8 * (kStages + s), ((it / kStages) - 1) & 1)) __nanosleep(256);
                }
                const int first = it * kStageChunks;
                const int cnt = min(kStageChunks, n_chunks - first);
                const uint32_t bytes = (uint32_t)(cnt * cb);
                mbar_arrive_expect_tx(bar0 + 8 * s, bytes);
                int64_t g = c0 + first;
                int head = cnt;
                if (ring > 0) { g %= ring; if (g + cnt > ring) head = (int)(ring - g); }
                tma_bulk_g2s(smem_u32(data + s * kStageChunks * cb), src + g * (int64_t)cb, (uint32_t)(head * cb), bar0 + 8 * s);
                if (head < cnt)
                    tma_bulk_g2s(smem_u32(data + s * kStageChunks * cb + head * cb), src, (uint32_t)((cnt - head) * cb), bar0 + 8 * s);
            }
        }
    } else {
        // ---------------- consumer warps ----------------
        for (int it = 0; it < n_iters; ++it) {
            const int s = it % kStages;
            mbar_wait(bar0 + 8 * s, (it / kStages) & 1);
            const int first = it * kStageChunks;
            const int cnt = min(kStageChunks, n_chunks - first);
            const unsigned char* stage = data + s * kStageChunks * cb;
            for (int c = 0; c < cnt; ++c) {
                int64_t g = c0 + first + c;
                if (ring > 0) g %= ring;
                cons.chunk(stage + c * cb, g);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (kStages + s));
        }
    }
}

}  // namespace nb
