// Initial conditions at scale: galaxy.py:10-92 (`create_disk_galaxy`) and :142-211
// (`create_galaxy_with_halo`) as COUNTER-BASED generators.
//
// The reference draws from torch's sequential generator, so a rank that owns stars [start, start+count) would have to
// generate (and hold) all N stars to get its slice.  Here star i's draws are Philox4x32-10(key = seed,
// counter = (i, stream)): any partition of [0, N) produces the same galaxy bit for bit, each rank generates only its own
// slice, and the two global quantities the recipes need are made partition-independent too:
//   * mean circular speed (velocity dispersion = 0.1 / 0.05 x mean, galaxy.py:82,206): Σ round(v·2^36) as int64 —
//     integer addition is associative, so an all-reduce over any sharding gives the same bits;
//   * enclosed visible mass of the halo recipe (argsort + cumsum of unit masses = rank in radius order,
//     galaxy.py:185-194): every rank regenerates the N radii (O(N) arithmetic, 8 B per star), counting-sorts them into
//     2^19 monotone radius bins and ranks its own stars exactly inside their bin (ties by index, as a stable argsort).
// Same distributions and formulas as the reference (fp32, same operation structure); NOT the same random stream as
// torch.manual_seed (the default create_disk_galaxy keeps that and stays fixture-identical).
#include <math.h>
#include "common.cuh"

namespace nb {

constexpr double kVsumScale = 68719476736.0;          // 2^36: v_circ ~ 0.01..10 -> 16M stars stay far below 2^63
constexpr float kGInit = 0.001f;                      // galaxy.py:59,181
constexpr int kInitBinShift = 12;                     // radius bin = float bits >> 12 (r >= 0.1: monotone, < 2^19 bins)

struct Philox { uint32_t x, y, z, w; };
__device__ __forceinline__ Philox philox4x32_10(uint64_t index, uint32_t stream, uint64_t seed) {
    uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = stream, c3 = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox{c0, c1, c2, c3};
}
__device__ __forceinline__ float u01(uint32_t bits) { return (float)(bits >> 8) * (1.0f / 16777216.0f); }            // [0, 1)
__device__ __forceinline__ float u01_open(uint32_t bits) { return (float)((bits >> 8) + 1u) * (1.0f / 16777216.0f); } // (0, 1]

struct DiskConsts {
    float neg_scale, c1, max_r, core_radius, inv_scale, inner_coef, outer_base, outer_coef, inv_norm;
};
__host__ inline DiskConsts disk_consts(int64_t n, double R, double cmf) {
    const double scale = R / 3.0, max_r = R * 2.0, total = (double)n;
    DiskConsts c;
    c.neg_scale = (float)(-scale);
    c.c1 = (float)(1.0 - exp(-max_r / scale));
    c.max_r = (float)max_r;
    c.core_radius = (float)(R * 0.2);
    c.inv_scale = (float)scale;                      // used as a divisor (r / scale), kept as the value itself
    c.inner_coef = (float)(cmf * total);
    c.outer_base = (float)(cmf * total);
    c.outer_coef = (float)((1.0 - cmf) * total);
    c.inv_norm = (float)(1.0 - 2.0 * exp(-max_r / scale));
    return c;
}

struct DiskStar { float r, angle, x, y, v_circ; };
// galaxy.py:33-79 for one star from its two uniforms
__device__ __forceinline__ DiskStar disk_star(const DiskConsts& c, float u, float ua) {
    DiskStar s;
    float r = __fmul_rn(c.neg_scale, logf(__fsub_rn(1.0f, __fmul_rn(u, c.c1))));           // -scale * log(1 - u * (1 - exp(..)))
    r = fminf(fmaxf(r, 0.1f), c.max_r);                                                   // clamp(min=0.1, max=max_r)
    s.r = r;
    s.angle = __fmul_rn(__fmul_rn(ua, 2.0f), 3.14159265358979323846f);                    // rand * 2 * pi
    float sn, cs;
    sincosf(s.angle, &sn, &cs);
    s.x = __fmul_rn(r, cs);
    s.y = __fmul_rn(r, sn);
    float enc;
    if (r < c.core_radius) {
        const float q = __fdiv_rn(r, c.core_radius);
        enc = __fmul_rn(c.inner_coef, __fmul_rn(q, q));                                   // cmf * M * (r / r_core)^2
    } else {
        const float t = __fdiv_rn(r, c.inv_scale);
        const float disk = __fdiv_rn(__fmul_rn(c.outer_coef, __fsub_rn(1.0f, __fmul_rn(__fadd_rn(1.0f, t), expf(-t)))), c.inv_norm);
        enc = __fadd_rn(c.outer_base, disk);
    }
    s.v_circ = sqrtf(__fdiv_rn(__fmul_rn(kGInit, enc), fmaxf(r, 0.1f)));                  // sqrt(G * M_enc / clamp(r, 0.1))
    return s;
}

__device__ __forceinline__ unsigned init_bin(float r) { return __float_as_uint(r) >> kInitBinShift; }
__device__ __forceinline__ float xy_radius(float x, float y) { return sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))); }

__global__ void __launch_bounds__(256) disk_phase1_kernel(DiskConsts c, uint64_t seed, int64_t start, int64_t count,
                                                          float* __restrict__ pos, float* __restrict__ vel, float* __restrict__ mass,
                                                          long long* __restrict__ vsum_fixed) {
    __shared__ long long red[32];
    long long fixed = 0;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        const Philox p = philox4x32_10((uint64_t)(start + k), 0u, seed);
        const DiskStar s = disk_star(c, u01(p.x), u01(p.y));
        fixed += llrint((double)s.v_circ * kVsumScale);
        if (pos) { pos[2 * k] = s.x; pos[2 * k + 1] = s.y; }
        if (vel) {
            float sn, cs;
            sincosf(s.angle, &sn, &cs);
            vel[2 * k] = -__fmul_rn(s.v_circ, sn);                                        // -v * sin(theta)
            vel[2 * k + 1] = __fmul_rn(s.v_circ, cs);                                     //  v * cos(theta)
        }
        if (mass) mass[k] = 1.0f;
    }
    fixed = block_reduce(fixed, OpAdd(), 0ll, red);
    if (threadIdx.x == 0 && fixed) atomicAdd(reinterpret_cast<unsigned long long*>(vsum_fixed), (unsigned long long)fixed);
}

// velocities += randn_like(velocities) * dispersion (galaxy.py:90,207): Box-Muller on the star's 3rd and 4th Philox words
__global__ void __launch_bounds__(256) add_dispersion_kernel(uint64_t seed, uint32_t stream, int64_t start, int64_t count, float dispersion,
                                                             float* __restrict__ vel) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        const Philox p = philox4x32_10((uint64_t)(start + k), stream, seed);
        const float rad = sqrtf(-2.0f * logf(u01_open(p.z)));
        float sn, cs;
        sincosf(6.28318530717958647692f * u01(p.w), &sn, &cs);
        vel[2 * k] = __fadd_rn(vel[2 * k], __fmul_rn(__fmul_rn(rad, cs), dispersion));
        vel[2 * k + 1] = __fadd_rn(vel[2 * k + 1], __fmul_rn(__fmul_rn(rad, sn), dispersion));
    }
}

// ---- halo recipe: rank of every star in radius order without a sort -------------------------------------------
// radius as the halo recipe sees it: sqrt((pos ** 2).sum()) of the GENERATED position (galaxy.py:182)
__device__ __forceinline__ float regenerated_radius(const DiskConsts& c, uint64_t seed, int64_t i) {
    const Philox p = philox4x32_10((uint64_t)i, 0u, seed);
    const DiskStar s = disk_star(c, u01(p.x), u01(p.y));
    return xy_radius(s.x, s.y);
}
__global__ void __launch_bounds__(256) radius_count_kernel(DiskConsts c, uint64_t seed, int64_t n, double* __restrict__ hist) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&hist[init_bin(regenerated_radius(c, seed, i))], 1.0);
}
__global__ void __launch_bounds__(256) radius_scatter_kernel(DiskConsts c, uint64_t seed, int64_t n, const double* __restrict__ prefix,
                                                             unsigned* __restrict__ cursor, float* __restrict__ sorted_r,
                                                             unsigned* __restrict__ sorted_idx) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float r = regenerated_radius(c, seed, i);
        const unsigned b = init_bin(r);
        const int64_t slot = (int64_t)prefix[b] + (int64_t)atomicAdd(&cursor[b], 1u);
        sorted_r[slot] = r;
        sorted_idx[slot] = (unsigned)i;
    }
}
// galaxy.py:176-204 for the local stars: exact rank inside the bin, NFW dark matter, circular speed, tangential velocity
__global__ void __launch_bounds__(256) halo_phase1_kernel(int64_t n, float halo_radius, float dm_total, int64_t start, int64_t count,
                                                          const float* __restrict__ pos, const double* __restrict__ hist,
                                                          const double* __restrict__ prefix, const float* __restrict__ sorted_r,
                                                          const unsigned* __restrict__ sorted_idx, float* __restrict__ vel,
                                                          long long* __restrict__ vsum_fixed) {
    __shared__ long long red[32];
    long long fixed = 0;
    const float f_norm = (float)(log(11.0) - 10.0 / 11.0);
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        const float x = pos[2 * k], y = pos[2 * k + 1];
        const float r = xy_radius(x, y);
        const unsigned b = init_bin(r);
        const int64_t seg0 = (int64_t)prefix[b], seg = (int64_t)hist[b];
        const unsigned me = (unsigned)(start + k);
        int64_t below = 0;
        for (int64_t q = 0; q < seg; ++q) {
            const float rj = sorted_r[seg0 + q];
            below += (rj < r || (rj == r && sorted_idx[seg0 + q] <= me)) ? 1 : 0;
        }
        const float enclosed_visible = (float)(seg0 + below);                             // cumsum of unit masses at its rank
        const float xs = __fdiv_rn(r, halo_radius);
        const float f_x = __fsub_rn(logf(__fadd_rn(1.0f, xs)), __fdiv_rn(xs, __fadd_rn(1.0f, xs)));
        const float enclosed_dm = __fdiv_rn(__fmul_rn(dm_total, f_x), f_norm);             // nfw_enclosed_mass
        const float v = sqrtf(__fdiv_rn(__fmul_rn(kGInit, __fadd_rn(enclosed_visible, enclosed_dm)), fmaxf(r, 0.1f)));
        const float th = atan2f(y, x);
        float sn, cs;
        sincosf(th, &sn, &cs);
        vel[2 * k] = -__fmul_rn(v, sn);
        vel[2 * k + 1] = __fmul_rn(v, cs);
        fixed += llrint((double)v * kVsumScale);
    }
    fixed = block_reduce(fixed, OpAdd(), 0ll, red);
    if (threadIdx.x == 0 && fixed) atomicAdd(reinterpret_cast<unsigned long long*>(vsum_fixed), (unsigned long long)fixed);
}

inline int init_grid(int64_t n) {
    int64_t blocks = (n + 255) / 256;
    if (blocks > kNumSMsB200 * 8) blocks = kNumSMsB200 * 8;
    return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace nb

using namespace nb;

extern "C" double nb_init_vsum_scale(void) { return kVsumScale; }

extern "C" int nb_disk_galaxy_phase1(int64_t num_stars, double galaxy_radius, double core_mass_fraction, uint64_t seed, int64_t start,
                                     int64_t count, float* pos, float* vel, float* mass, int64_t* vsum_fixed, void* stream) {
    if (num_stars <= 0 || start < 0 || count <= 0 || start + count > num_stars || !vsum_fixed || !(galaxy_radius > 0.0))
        return NB_ERR_INVALID_ARGUMENT;
    disk_phase1_kernel<<<init_grid(count), 256, 0, (cudaStream_t)stream>>>(disk_consts(num_stars, galaxy_radius, core_mass_fraction), seed,
                                                                             start, count, pos, vel, mass, (long long*)vsum_fixed);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_galaxy_add_dispersion(uint64_t seed, int stream_id, int64_t start, int64_t count, double dispersion, float* vel,
                                        void* stream) {
    if (start < 0 || count <= 0 || !vel || stream_id < 0) return NB_ERR_INVALID_ARGUMENT;
    add_dispersion_kernel<<<init_grid(count), 256, 0, (cudaStream_t)stream>>>(seed, (uint32_t)stream_id, start, count, (float)dispersion, vel);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_disk_radius_histogram(int64_t num_stars, double galaxy_radius, double core_mass_fraction, uint64_t seed, double* hist,
                                        void* stream) {
    if (num_stars <= 0 || !hist || !(galaxy_radius > 0.0)) return NB_ERR_INVALID_ARGUMENT;
    radius_count_kernel<<<init_grid(num_stars), 256, 0, (cudaStream_t)stream>>>(disk_consts(num_stars, galaxy_radius, core_mass_fraction),
                                                                                 seed, num_stars, hist);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_disk_radius_scatter(int64_t num_stars, double galaxy_radius, double core_mass_fraction, uint64_t seed, const double* prefix,
                                      uint32_t* cursor, float* sorted_r, uint32_t* sorted_idx, void* stream) {
    if (num_stars <= 0 || num_stars > 0xffffffffll || !prefix || !cursor || !sorted_r || !sorted_idx) return NB_ERR_INVALID_ARGUMENT;
    radius_scatter_kernel<<<init_grid(num_stars), 256, 0, (cudaStream_t)stream>>>(disk_consts(num_stars, galaxy_radius, core_mass_fraction),
                                                                                   seed, num_stars, prefix, cursor, sorted_r, sorted_idx);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}

extern "C" int nb_halo_phase1(int64_t num_stars, double halo_radius, double dm_mass_ratio, int64_t start, int64_t count, const float* pos,
                              const double* hist, const double* prefix, const float* sorted_r, const uint32_t* sorted_idx, float* vel,
                              int64_t* vsum_fixed, void* stream) {
    if (num_stars <= 0 || start < 0 || count <= 0 || start + count > num_stars || !pos || !hist || !prefix || !sorted_r || !sorted_idx ||
        !vel || !vsum_fixed)
        return NB_ERR_INVALID_ARGUMENT;
    const float dm_total = (float)((double)num_stars * dm_mass_ratio);                     // mass.sum().item() * dm_mass_ratio
    halo_phase1_kernel<<<init_grid(count), 256, 0, (cudaStream_t)stream>>>(num_stars, (float)halo_radius, dm_total, start, count, pos, hist,
                                                                            prefix, sorted_r, sorted_idx, vel, (long long*)vsum_fixed);
    NB_CUDA_LAUNCH_CHECK();
    return NB_OK;
}
