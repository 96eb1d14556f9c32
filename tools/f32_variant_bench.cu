// Developer harness: time ONE source-level variant (-DNB_F32_ACC_VARIANT, -DNB_F32_PERTURB, -DNB_F32_UNROLL) of the benchmark
// force kernels at N = 2^20 (D = 3 and D = 2, uniform masses) — the ptxas schedule lottery of DESIGN.md §5.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -fmad=false -DNB_F32_PERTURB=5 -o /tmp/v5 tools/f32_variant_bench.cu
#define NB_TUNE_HARNESS
#include "../nbody_cosmological_simulation_b200/csrc/accel.cu"
#include <cstdio>
#include <vector>
#include <random>
#ifndef NB_F32_UNROLL
#define NB_F32_UNROLL 4
#endif
extern "C" int64_t nb_chunk_sources(int dtype) { return dtype == NB_F32 ? 256 : 128; }
extern "C" int64_t nb_num_chunks(int64_t n, int dtype) { int64_t cs = nb_chunk_sources(dtype); return (n + cs - 1) / cs; }

template <class Consumer>
static float run(const AccelArgs& a, int64_t ws_bytes) {
    int splits = 0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (launch_accel<Consumer, 0>(a, ws_bytes, 0, &splits)) return -1.f;
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); launch_accel<Consumer, 0>(a, ws_bytes, 0, &splits); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

template <int DIM>
static float bench(int64_t n) {
    std::mt19937 rng(42); std::uniform_real_distribution<float> u(-10.f, 10.f);
    const int64_t chunks = nb_num_chunks(n, NB_F32);
    std::vector<char> packed(chunks * chunk_bytes(DIM)); std::vector<float> pos(n * DIM);
    for (int64_t i = 0; i < n; ++i) {
        float p[3] = {u(rng), u(rng), u(rng)};
        char* base = packed.data() + (i / 256) * chunk_bytes(DIM);
        const int un = (i % 256) / 2, h = i & 1;
        float* A = (float*)(base + un * 16); A[h] = p[0]; A[2 + h] = p[1];
        if (DIM == 3) { float* B = (float*)(base + kChunkABytes + un * 16); B[h] = p[2]; B[2 + h] = 1e-3f; }
        else { float* B = (float*)(base + kChunkABytes + un * 8); B[h] = 1e-3f; }
        for (int k = 0; k < DIM; ++k) pos[i * DIM + k] = p[k];
    }
    char *dpacked, *dpos; double* ws; const int64_t ws_bytes = (int64_t)10 * n * DIM * 8;
    cudaMalloc(&dpacked, packed.size()); cudaMalloc(&dpos, pos.size() * 4); cudaMalloc(&ws, ws_bytes);
    cudaMemcpy(dpacked, packed.data(), packed.size(), cudaMemcpyHostToDevice); cudaMemcpy(dpos, pos.data(), pos.size() * 4, cudaMemcpyHostToDevice);
    AccelArgs a{}; a.src = dpacked; a.n_chunks = chunks; a.pos_tgt = dpos; a.n_tgt = n; a.partial = ws; a.eps_sq = 0.01; a.neg_zero = -0.0f;
    const float ms = run<ForceF32<DIM, Q_F32, 2, 256, true, NB_F32_UNROLL>>(a, ws_bytes);
    cudaFree(dpacked); cudaFree(dpos); cudaFree(ws);
    return ms;
}

int main() {
    const int64_t n = 1 << 20;
    const float m3 = bench<3>(n), m2 = bench<2>(n);
    printf("acc_variant %d perturb %2d unroll %d :  D=3 %8.3f ms (%.3f T inter/s)   D=2 %8.3f ms (%.3f T inter/s)\n", NB_F32_ACC_VARIANT,
           NB_F32_PERTURB, NB_F32_UNROLL, m3, (double)n * n / (m3 * 1e-3) / 1e12, m2, (double)n * n / (m2 * 1e-3) / 1e12);
    return 0;
}
