// Pass 1 of the log-grid modes: hi = log(max over ALL pairs of d²)   (quantization.py:112-113).
//
// The reference reduces the N×N d² tensor.  Here the exact same maximum is found in O(N) + O(C²):
// a pair can only attain the maximum if both ends lie in the outer shell of the point set —
//   R_i + R_j >= d_ij   and   max d >= D_lb   =>   R_i >= D_lb − R_max
// with R = distance from the bounding-box centre, D_lb = distance from the farthest point to its farthest
// partner.  Only the C shell candidates are compared pairwise, with the reference's exact rounding sequence
// rn(rn(rn(dx²)+rn(dy²))[+rn(dz²)]) (and max rn(s+ε²) = rn(max s + ε²) by monotonicity of rn).  Safety margins
// (1e-4·D) dwarf the fp32 error of the pruning quantities, so the result is bit-identical to the full N² scan;
// a spherical shell of points degrades to the O(N²) scan (C = N), nothing worse.
// Every rank of a sharded run holds the full packed source set, so each computes the global value locally.
#include "common.cuh"

namespace nb {

struct MdHeader {                     // 128-byte workspace header, reset by md_init
    unsigned lo[3], hi[3];            // bounding box, order-preserving float keys
    unsigned dlb_bits;                // max d² from the far point (non-negative float bits)
    unsigned count;                   // number of candidates
    unsigned flags;                   // 1: NaN coordinate seen, 2: ±Inf coordinate seen
    unsigned found;                   // candidates found (kept when `count` is zeroed by the single-CTA path)
    unsigned long long far_key;       // (float bits of R² << 32) | index of the farthest point
    unsigned best_bits;               // max over candidate pairs of s (non-negative float bits)
    unsigned pad1[19];
};
static_assert(sizeof(MdHeader) == 128, "header layout");

__device__ __forceinline__ unsigned fkey(float v) { unsigned b = __float_as_uint(v); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float funkey(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

template <int DIM>
__device__ __forceinline__ void load_source(const char* __restrict__ packed, int64_t j, float& x, float& y, float& z) {
    const int64_t chunk = j / (2 * kChunkUnits);
    const int r = (int)(j % (2 * kChunkUnits)), u = r >> 1, h = r & 1;
    const float* A = reinterpret_cast<const float*>(packed + chunk * (int64_t)chunk_bytes(DIM) + u * 16);
    x = A[h]; y = A[2 + h]; z = 0.f;
    if (DIM == 3) z = reinterpret_cast<const float*>(packed + chunk * (int64_t)chunk_bytes(DIM) + kChunkABytes + u * 16)[h];
}

__global__ void md_init(MdHeader* h) {
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) { h->lo[k] = 0xffffffffu; h->hi[k] = 0u; }
        h->dlb_bits = 0u; h->count = 0u; h->flags = 0u; h->found = 0u; h->far_key = 0ull; h->best_bits = 0u;
    }
}

template <int DIM>
__global__ void __launch_bounds__(256) md_bbox(const char* __restrict__ packed, int64_t n, MdHeader* __restrict__ h) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    unsigned flags = 0;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        float p[3];
        load_source<DIM>(packed, j, p[0], p[1], p[2]);
        if (p[0] > kPadDetectF32) continue;                          // padding record
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            if (p[k] != p[k]) flags |= 1u; else if (fabsf(p[k]) == INFINITY) flags |= 2u;
            lo[k] = fminf(lo[k], p[k]); hi[k] = fmaxf(hi[k], p[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        lo[k] = warp_reduce(lo[k], OpMin()); hi[k] = warp_reduce(hi[k], OpMax());
    }
    flags = __reduce_or_sync(0xffffffffu, flags);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < DIM; ++k) { atomicMin(&h->lo[k], fkey(lo[k])); atomicMax(&h->hi[k], fkey(hi[k])); }
        if (flags) atomicOr(&h->flags, flags);
    }
}

template <int DIM>
__device__ __forceinline__ void centre_of(const MdHeader* h, float* c) {
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = k < DIM ? 0.5f * funkey(h->lo[k]) + 0.5f * funkey(h->hi[k]) : 0.f;
}

template <int DIM>
__global__ void __launch_bounds__(256) md_far_point(const char* __restrict__ packed, int64_t n, MdHeader* __restrict__ h) {
    float c[3];
    centre_of<DIM>(h, c);
    unsigned long long best = 0ull;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        float x, y, z;
        load_source<DIM>(packed, j, x, y, z);
        if (x > kPadDetectF32) continue;
        const float r2 = (x - c[0]) * (x - c[0]) + (y - c[1]) * (y - c[1]) + (z - c[2]) * (z - c[2]);
        const unsigned long long key = ((unsigned long long)__float_as_uint(r2) << 32) | (unsigned long long)(unsigned)j;
        if (r2 == r2 && key > best) best = key;
    }
    best = warp_reduce(best, OpMax());
    if ((threadIdx.x & 31) == 0 && best) atomicMax(&h->far_key, best);
}

template <int DIM>
__global__ void __launch_bounds__(256) md_lower_bound(const char* __restrict__ packed, int64_t n, MdHeader* __restrict__ h) {
    float px, py, pz;
    load_source<DIM>(packed, (int64_t)(h->far_key & 0xffffffffull), px, py, pz);
    float best = 0.f;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        float x, y, z;
        load_source<DIM>(packed, j, x, y, z);
        if (x > kPadDetectF32) continue;
        best = fmaxf(best, (x - px) * (x - px) + (y - py) * (y - py) + (z - pz) * (z - pz));
    }
    best = warp_reduce(best, OpMax());
    if ((threadIdx.x & 31) == 0) atomicMax(&h->dlb_bits, __float_as_uint(best));
}

template <int DIM>
__global__ void __launch_bounds__(256) md_compact(const char* __restrict__ packed, int64_t n, MdHeader* __restrict__ h,
                                                  float4* __restrict__ cand) {
    float c[3];
    centre_of<DIM>(h, c);
    const float rmax = sqrtf(__uint_as_float((unsigned)(h->far_key >> 32)));
    const float dlb = sqrtf(__uint_as_float(h->dlb_bits));
    const float thr = dlb * (1.f - 1e-4f) - rmax * (1.f + 1e-5f) - 1e-30f;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (n + 31) / 32 * 32;                     // whole warps iterate together (ballot below)
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_round; j += stride) {
        float x = 0.f, y = 0.f, z = 0.f;
        bool keep = false;
        if (j < n) {
            load_source<DIM>(packed, j, x, y, z);
            if (!(x > kPadDetectF32)) {
                const float r = sqrtf((x - c[0]) * (x - c[0]) + (y - c[1]) * (y - c[1]) + (z - c[2]) * (z - c[2]));
                keep = r >= thr;
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (mask) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&h->count, (unsigned)__popc(mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) cand[base + __popc(mask & ((1u << lane) - 1u))] = make_float4(x, y, z, 0.f);
        }
    }
}

// exact pairwise maximum over the candidates (shared-memory tiles of 256)
template <int DIM>
__global__ void __launch_bounds__(256) md_pairs(const float4* __restrict__ cand, MdHeader* __restrict__ h) {
    __shared__ float4 tile[256];
    const unsigned count = h->count;
    float best = 0.f;
    for (unsigned ib = blockIdx.x * 256u; ib < count; ib += gridDim.x * 256u) {
        const unsigned i = ib + threadIdx.x;
        const float4 me = cand[i < count ? i : count - 1];
        for (unsigned jb = 0; jb < count; jb += 256u) {
            __syncthreads();
            const unsigned j = jb + threadIdx.x;
            tile[threadIdx.x] = cand[j < count ? j : count - 1];
            __syncthreads();
#pragma unroll 8
            for (int t = 0; t < 256; ++t) {
                const float4 o = tile[t];
                const float dx = __fsub_rn(o.x, me.x), dy = __fsub_rn(o.y, me.y);
                float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));             // simulation.py:86 rounding order
                if (DIM == 3) { const float dz = __fsub_rn(o.z, me.z); s = __fadd_rn(s, __fmul_rn(dz, dz)); }
                best = fmaxf(best, s);
            }
        }
    }
    best = warp_reduce(best, OpMax());
    if ((threadIdx.x & 31) == 0) atomicMax(&h->best_bits, __float_as_uint(best));
}

__global__ void md_publish(MdHeader* __restrict__ h, float eps2, int64_t* __restrict__ scalars) {
    if (threadIdx.x == 0) {
        if (h->count) h->found = h->count;
        float v = __fadd_rn(__uint_as_float(h->best_bits), eps2);      // rn(max s + ε²)
        if (h->flags & 2u) v = INFINITY;                                // torch.max semantics for non-finite inputs
        if (h->flags & 1u) v = __int_as_float(0x7fffffff);
        atomicMax(reinterpret_cast<long long*>(scalars + NB_SLOT_MAX_D2), (long long)key_from_double((double)v));
    }
}

// Small systems (n <= kMdSmallMax): the four O(n) phases above in ONE CTA with block barriers in between instead of
// four launches + an init launch (a tick of a 10⁴-star int-mode run spent 25 µs in these seven tiny kernels).  The
// arithmetic is the same, so the candidate SET is the same; if it is small the exact pairwise maximum is taken here
// as well and `count` is zeroed so that md_pairs, launched afterwards as always, finds nothing left to do.
constexpr int64_t kMdSmallMax = 32768;
constexpr unsigned kMdSmallPairs = 1024;

template <int DIM>
__global__ void __launch_bounds__(1024) md_small(const char* __restrict__ packed, int64_t n, MdHeader* __restrict__ gh,
                                                 float4* __restrict__ cand) {
    __shared__ MdHeader h;
    __shared__ float4 tile[256];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) {
        for (int k = 0; k < 3; ++k) { h.lo[k] = 0xffffffffu; h.hi[k] = 0u; }
        h.dlb_bits = 0u; h.count = 0u; h.flags = 0u; h.found = 0u; h.far_key = 0ull; h.best_bits = 0u;
    }
    __syncthreads();
    {   // bounding box + non-finite flags (md_bbox)
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        unsigned flags = 0;
        for (int64_t j = tid; j < n; j += blockDim.x) {
            float p[3];
            load_source<DIM>(packed, j, p[0], p[1], p[2]);
            if (p[0] > kPadDetectF32) continue;
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                if (p[k] != p[k]) flags |= 1u; else if (fabsf(p[k]) == INFINITY) flags |= 2u;
                lo[k] = fminf(lo[k], p[k]); hi[k] = fmaxf(hi[k], p[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < DIM; ++k) { lo[k] = warp_reduce(lo[k], OpMin()); hi[k] = warp_reduce(hi[k], OpMax()); }
        flags = __reduce_or_sync(0xffffffffu, flags);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < DIM; ++k) { atomicMin(&h.lo[k], fkey(lo[k])); atomicMax(&h.hi[k], fkey(hi[k])); }
            if (flags) atomicOr(&h.flags, flags);
        }
    }
    __syncthreads();
    float c[3];
    centre_of<DIM>(&h, c);
    {   // farthest point from the centre (md_far_point)
        unsigned long long best = 0ull;
        for (int64_t j = tid; j < n; j += blockDim.x) {
            float x, y, z;
            load_source<DIM>(packed, j, x, y, z);
            if (x > kPadDetectF32) continue;
            const float r2 = (x - c[0]) * (x - c[0]) + (y - c[1]) * (y - c[1]) + (z - c[2]) * (z - c[2]);
            const unsigned long long key = ((unsigned long long)__float_as_uint(r2) << 32) | (unsigned long long)(unsigned)j;
            if (r2 == r2 && key > best) best = key;
        }
        best = warp_reduce(best, OpMax());
        if (lane == 0 && best) atomicMax(&h.far_key, best);
    }
    __syncthreads();
    {   // its farthest partner: lower bound of the maximum (md_lower_bound)
        float px, py, pz;
        load_source<DIM>(packed, (int64_t)(h.far_key & 0xffffffffull), px, py, pz);
        float best = 0.f;
        for (int64_t j = tid; j < n; j += blockDim.x) {
            float x, y, z;
            load_source<DIM>(packed, j, x, y, z);
            if (x > kPadDetectF32) continue;
            best = fmaxf(best, (x - px) * (x - px) + (y - py) * (y - py) + (z - pz) * (z - pz));
        }
        best = warp_reduce(best, OpMax());
        if (lane == 0) atomicMax(&h.dlb_bits, __float_as_uint(best));
    }
    __syncthreads();
    {   // outer-shell candidates (md_compact)
        const float rmax = sqrtf(__uint_as_float((unsigned)(h.far_key >> 32)));
        const float dlb = sqrtf(__uint_as_float(h.dlb_bits));
        const float thr = dlb * (1.f - 1e-4f) - rmax * (1.f + 1e-5f) - 1e-30f;
        const int64_t n_round = (n + 31) / 32 * 32;
        for (int64_t j = tid; j < n_round; j += blockDim.x) {
            float x = 0.f, y = 0.f, z = 0.f;
            bool keep = false;
            if (j < n) {
                load_source<DIM>(packed, j, x, y, z);
                if (!(x > kPadDetectF32)) {
                    const float r = sqrtf((x - c[0]) * (x - c[0]) + (y - c[1]) * (y - c[1]) + (z - c[2]) * (z - c[2]));
                    keep = r >= thr;
                }
            }
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (mask) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&h.count, (unsigned)__popc(mask));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (keep) cand[base + __popc(mask & ((1u << lane) - 1u))] = make_float4(x, y, z, 0.f);
            }
        }
    }
    __syncthreads();                                   // candidates written by this block are visible to it
    const unsigned count = h.count;
    if (tid == 0) h.found = count;
    if (count <= kMdSmallPairs) {
        // exact pairwise maximum over the candidates (md_pairs), rows strided over the block
        float best = 0.f;
        for (unsigned ib = 0; ib < count; ib += blockDim.x) {
            const unsigned i = ib + tid;
            const float4 me = cand[i < count ? i : count - 1];
            for (unsigned jb = 0; jb < count; jb += 256u) {
                __syncthreads();
                if (tid < 256) { const unsigned j = jb + tid; tile[tid] = cand[j < count ? j : count - 1]; }
                __syncthreads();
                const int lim = (int)min(256u, count - jb);
                for (int t = 0; t < lim; ++t) {
                    const float4 o = tile[t];
                    const float dx = __fsub_rn(o.x, me.x), dy = __fsub_rn(o.y, me.y);
                    float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));         // simulation.py:86 rounding order
                    if (DIM == 3) { const float dz = __fsub_rn(o.z, me.z); s = __fadd_rn(s, __fmul_rn(dz, dz)); }
                    best = fmaxf(best, s);
                }
            }
        }
        best = warp_reduce(best, OpMax());
        if (lane == 0) atomicMax(&h.best_bits, __float_as_uint(best));
        __syncthreads();
        if (tid == 0) h.count = 0u;                    // nothing left for md_pairs
        __syncthreads();
    }
    if (tid < (int)(sizeof(MdHeader) / 4)) reinterpret_cast<unsigned*>(gh)[tid] = reinterpret_cast<const unsigned*>(&h)[tid];
}

template <int DIM>
int launch_max_dist(const char* packed, int64_t n_src, float eps2, int64_t* scalars, void* ws, cudaStream_t st) {
    MdHeader* h = reinterpret_cast<MdHeader*>(ws);
    float4* cand = reinterpret_cast<float4*>(reinterpret_cast<char*>(ws) + sizeof(MdHeader));
    int64_t blocks = (n_src + 255) / 256;
    if (blocks > kNumSMsB200 * 8) blocks = kNumSMsB200 * 8;
    if (n_src <= kMdSmallMax) {
        md_small<DIM><<<1, 1024, 0, st>>>(packed, n_src, h, cand);
    } else {
        md_init<<<1, 32, 0, st>>>(h);
        md_bbox<DIM><<<(int)blocks, 256, 0, st>>>(packed, n_src, h);
        md_far_point<DIM><<<(int)blocks, 256, 0, st>>>(packed, n_src, h);
        md_lower_bound<DIM><<<(int)blocks, 256, 0, st>>>(packed, n_src, h);
        md_compact<DIM><<<(int)blocks, 256, 0, st>>>(packed, n_src, h, cand);
    }
    md_pairs<DIM><<<(int)blocks, 256, 0, st>>>(cand, h);
    md_publish<<<1, 32, 0, st>>>(h, eps2, scalars);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

}  // namespace nb

using namespace nb;

extern "C" int64_t nb_max_dist_workspace_bytes(int64_t n_src) {
    return n_src <= 0 ? 0 : (int64_t)sizeof(MdHeader) + ((n_src + 255) / 256 * 256) * (int64_t)sizeof(float4);
}

extern "C" int nb_max_dist_sq(const void* packed_src, int64_t n_src, int dim, int dtype, double eps_sq, int64_t* scalars,
                              void* workspace, int64_t workspace_bytes, void* stream) {
    if (!packed_src || !scalars || !workspace || n_src <= 0 || (dim != 2 && dim != 3)) return NB_ERR_INVALID_ARGUMENT;
    if (dtype != NB_F32) return NB_ERR_UNSUPPORTED;       // int modes on fp64 state: no caller in the reference
    if (workspace_bytes < nb_max_dist_workspace_bytes(n_src)) return NB_ERR_WORKSPACE_TOO_SMALL;
    if (n_src > 0xffffffffll) return NB_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    return dim == 2 ? launch_max_dist<2>((const char*)packed_src, n_src, (float)eps_sq, scalars, workspace, st)
                    : launch_max_dist<3>((const char*)packed_src, n_src, (float)eps_sq, scalars, workspace, st);
}
