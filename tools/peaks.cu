// Pipe-throughput microbenchmarks for the roofline denominators the N-body kernels are judged
// against (FP32 FFMA / packed FFMA2, MUFU.RSQ, FP64 DFMA, conversions, LDS broadcast).
// MEASURED_PEAKS.json only carries HBM and bf16-tensor peaks; SURVEY.md §7 asks for these.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/peaks tools/peaks.cu
//   ./tools/peaks > gpurun_out/peaks.json
//
// Every kernel runs ILP independent dependency chains per thread for ITERS iterations; the result
// is reported as lane-operations per clock per SM (from the wall time of a full-occupancy grid and
// the SM clock sampled by clock64/globaltimer) and as T(FL)OP/s.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1);} } while (0)

constexpr int ILP = 8;
constexpr int ITERS = 32768;     // ~1-10 ms per kernel: launch ramp and tail are < 1 % of the time
constexpr int UNROLL = 16;

struct Result { unsigned long long cycles; };

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <int OP>
__global__ void __launch_bounds__(1024, 1) k_fp32(float* out, float seed, unsigned long long* cyc) {
    float a[ILP], b = seed, c = seed * 0.5f;
    #pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = seed + i + threadIdx.x;
    const unsigned long long g0 = gtimer();
    long long t0 = clock64();
    for (int it = 0; it < ITERS / UNROLL; ++it) {
        #pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            #pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == 0) a[i] = fmaf(a[i], b, c);
                if (OP == 1) a[i] = __fmul_rn(a[i], b);
                if (OP == 2) a[i] = __fadd_rn(a[i], c);
                if (OP == 3) asm volatile("rsqrt.approx.f32 %0, %0;" : "+f"(a[i]));
                if (OP == 4) a[i] = fmaxf(a[i], c) + 0.0f * b;     // FMNMX (alu pipe) — the add folds away
                if (OP == 5) { float x = fabsf(a[i]); asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(x)); }   // MUFU.LG2 R, |R|
                if (OP == 6) { float x = -fabsf(a[i]); asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(x)); }  // MUFU.EX2 R, -|R|
                if (OP == 7) { float x = fabsf(a[i]); asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(x)); } // MUFU.RSQ R, |R|
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    #pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (unsigned long long)(t1 - t0); cyc[1] = gtimer() - g0; }
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) k_fp32x2(float* out, float seed, unsigned long long* cyc) {
    float2 a[ILP], b = make_float2(seed, seed * 1.0001f), c = make_float2(seed * 0.5f, seed * 0.25f);
    #pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
    const unsigned long long g0 = gtimer();
    long long t0 = clock64();
    for (int it = 0; it < ITERS / UNROLL; ++it) {
        #pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            #pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == 0) a[i] = __ffma2_rn(a[i], b, c);
                if (OP == 1) a[i] = __fmul2_rn(a[i], b);
                if (OP == 2) a[i] = __fadd2_rn(a[i], c);
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    #pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (unsigned long long)(t1 - t0); cyc[1] = gtimer() - g0; }
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) k_fp64(double* out, double seed, unsigned long long* cyc) {
    double a[ILP], b = seed, c = seed * 0.5;
    #pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = seed + i + threadIdx.x;
    const unsigned long long g0 = gtimer();
    long long t0 = clock64();
    for (int it = 0; it < ITERS / UNROLL; ++it) {
        #pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            #pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == 0) a[i] = fma(a[i], b, c);
                if (OP == 1) a[i] = __dmul_rn(a[i], b);
                if (OP == 2) a[i] = __dadd_rn(a[i], c);
                if (OP == 3) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(a[i]));   // MUFU.RSQ64H
                if (OP == 4) { float f = (float)a[i]; a[i] = (double)f + c; }                 // F2F.F32.F64 + F2F.F64.F32 (+DADD)
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    #pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (unsigned long long)(t1 - t0); cyc[1] = gtimer() - g0; }
}

// LDS.128 broadcast (all lanes read the same 16 B) — the source-tile access pattern of the force kernel
__global__ void __launch_bounds__(1024, 1) k_lds(float* out, unsigned long long* cyc) {
    __shared__ float4 tile[1024];
    tile[threadIdx.x] = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
    __syncthreads();
    float4 acc[4] = {};
    const unsigned long long g0 = gtimer();
    long long t0 = clock64();
    for (int it = 0; it < ITERS / UNROLL; ++it) {
        #pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            #pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 v = tile[(it * UNROLL + u * 4 + i) & 1023];
                acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0].x + acc[1].y + acc[2].z + acc[3].w;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (unsigned long long)(t1 - t0); cyc[1] = gtimer() - g0; }
}

struct Row { std::string name; double ops_per_thread; double flop_per_op; };

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    int clock_khz = 0; CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    // ONE wave: every kernel here is __launch_bounds__(1024, 1) and may use up to 64 registers, so exactly one CTA of
    // 1024 threads (32 warps, 8 per SM sub-partition x ILP 8 chains) is resident per SM.  (Round 1 launched 2 x SMs
    // CTAs and assumed both were co-resident: its per-SM and implied-clock columns were off by 2x.)
    const int threads = 1024, blocks = sms;
    void* out; CK(cudaMalloc(&out, (size_t)blocks * threads * 8));
    unsigned long long* dcyc; CK(cudaMalloc(&dcyc, 16));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f, \"results\": {\n", prop.name, sms, clock_khz / 1000.0);
    bool first = true;
    auto run = [&](const char* name, auto launch, double ops_per_thread, double flop_per_op) {
        for (int w = 0; w < 3; ++w) launch();
        CK(cudaDeviceSynchronize());
        float best = 1e30f; unsigned long long cyc[2] = {0, 0};
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) { best = ms; CK(cudaMemcpy(cyc, dcyc, 16, cudaMemcpyDeviceToHost)); }
        }
        CK(cudaGetLastError());
        double total_ops = ops_per_thread * threads * (double)blocks;
        double ops_per_s = total_ops / (best * 1e-3);
        // block 0 times its own loop with clock64 (SM cycles) and globaltimer (ns): per-SM rate from its cycles (one CTA
        // per SM), SM clock under this load = cycles / ns
        double lane_ops_per_clk_sm = ops_per_thread * threads / (double)cyc[0];
        double eff_mhz = cyc[1] ? (double)cyc[0] / (double)cyc[1] * 1e3 : 0.0;
        printf("%s  \"%s\": {\"ms\": %.4f, \"lane_ops_per_clk_per_sm\": %.2f, \"Tops_per_s\": %.3f, \"Tflops\": %.3f, \"implied_sm_mhz\": %.0f}",
               first ? "" : ",\n", name, best, lane_ops_per_clk_sm, ops_per_s / 1e12, ops_per_s * flop_per_op / 1e12, eff_mhz);
        first = false;
    };
    const double n = (double)ITERS * ILP;
    run("ffma",      [&] { k_fp32<0><<<blocks, threads>>>((float*)out, 1.0001f, dcyc); }, n, 2);
    run("fmul",      [&] { k_fp32<1><<<blocks, threads>>>((float*)out, 1.0001f, dcyc); }, n, 1);
    run("fadd",      [&] { k_fp32<2><<<blocks, threads>>>((float*)out, 1.0001f, dcyc); }, n, 1);
    run("mufu_rsq",  [&] { k_fp32<3><<<blocks, threads>>>((float*)out, 1.0001f, dcyc); }, n, 1);
    run("mufu_lg2",  [&] { k_fp32<5><<<blocks, threads>>>((float*)out, 1.0001f, dcyc); }, n, 1);
    run("mufu_ex2",  [&] { k_fp32<6><<<blocks, threads>>>((float*)out, 1.0001f, dcyc); }, n, 1);
    run("mufu_rsq_nochain", [&] { k_fp32<7><<<blocks, threads>>>((float*)out, 1.0001f, dcyc); }, n, 1);
    run("fmnmx",     [&] { k_fp32<4><<<blocks, threads>>>((float*)out, 1.0001f, dcyc); }, n, 1);
    run("ffma2",     [&] { k_fp32x2<0><<<blocks, threads>>>((float*)out, 1.0001f, dcyc); }, 2 * n, 2);
    run("fmul2",     [&] { k_fp32x2<1><<<blocks, threads>>>((float*)out, 1.0001f, dcyc); }, 2 * n, 1);
    run("fadd2",     [&] { k_fp32x2<2><<<blocks, threads>>>((float*)out, 1.0001f, dcyc); }, 2 * n, 1);
    run("dfma",      [&] { k_fp64<0><<<blocks, threads>>>((double*)out, 1.0001, dcyc); }, n, 2);
    run("dmul",      [&] { k_fp64<1><<<blocks, threads>>>((double*)out, 1.0001, dcyc); }, n, 1);
    run("dadd",      [&] { k_fp64<2><<<blocks, threads>>>((double*)out, 1.0001, dcyc); }, n, 1);
    run("mufu_rsq64h", [&] { k_fp64<3><<<blocks, threads>>>((double*)out, 1.0001, dcyc); }, n, 1);
    run("f2f_f64_f32_roundtrip", [&] { k_fp64<4><<<blocks, threads>>>((double*)out, 1.0001, dcyc); }, 2 * n, 1);
    run("lds128_broadcast", [&] { k_lds<<<blocks, threads>>>((float*)out, dcyc); }, (double)ITERS * 4, 1);
    printf("\n}}\n");
    return 0;
}
