// Developer harness: time force-kernel configurations (threads x targets/thread x unroll) on one GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DNB_TUNE_HARNESS -o tools/tune_force tools/tune_force.cu
#define NB_TUNE_HARNESS
#include "../nbody_cosmological_simulation_b200/csrc/accel.cu"
#include <cstdio>
#include <vector>
#include <random>

extern "C" int64_t nb_chunk_sources(int dtype) { return dtype == NB_F32 ? 256 : 128; }
extern "C" int64_t nb_num_chunks(int64_t n, int dtype) { int64_t cs = nb_chunk_sources(dtype); return (n + cs - 1) / cs; }

template <class Consumer>
static void run(const char* name, const AccelArgs& a, int64_t ws_bytes, double n_inter) {
    int splits = 0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = launch_accel<Consumer, 0>(a, ws_bytes, 0, &splits);
    if (rc) { printf("%-34s launch failed rc=%d\n", name, rc); return; }
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); launch_accel<Consumer, 0>(a, ws_bytes, 0, &splits); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, accel_kernel<Consumer, 0>);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, accel_kernel<Consumer, 0>, Consumer::THREADS + 32, stream_smem_bytes(Consumer::DIM));
    printf("%-34s regs %3d occ %d splits %2d  %8.3f ms  %6.3f T inter/s  %6.2f TFLOP/s@20\n", name, fa.numRegs, occ, splits, best,
           n_inter / (best * 1e-3) / 1e12, 20 * n_inter / (best * 1e-3) / 1e12);
}

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 262144;
    const bool f64 = argc > 2 && atoi(argv[2]) == 64;
    const int dim = 3;
    std::mt19937 rng(42); std::uniform_real_distribution<double> u(-10.0, 10.0);
    const int dtype = f64 ? NB_F64 : NB_F32;
    const int64_t chunks = nb_num_chunks(n, dtype);
    const size_t esz = f64 ? 8 : 4;
    std::vector<char> packed(chunks * chunk_bytes(dim)), pos(n * dim * esz);
    for (int64_t i = 0; i < n; ++i) {
        double p[3] = {u(rng), u(rng), u(rng)};
        const int64_t c = f64 ? i / 128 : i / 256; char* base = packed.data() + c * chunk_bytes(dim);
        if (f64) {
            const int un = i % 128; double* A = (double*)(base + un * 16); double* B = (double*)(base + kChunkABytes + un * 16);
            A[0] = p[0]; A[1] = p[1]; B[0] = p[2]; B[1] = 1e-3; double* P = (double*)pos.data(); P[i*3]=p[0]; P[i*3+1]=p[1]; P[i*3+2]=p[2];
        } else {
            const int un = (i % 256) / 2, h = i & 1; float* A = (float*)(base + un * 16); float* B = (float*)(base + kChunkABytes + un * 16);
            A[h] = (float)p[0]; A[2 + h] = (float)p[1]; B[h] = (float)p[2]; B[2 + h] = 1e-3f; float* P = (float*)pos.data(); P[i*3]=(float)p[0]; P[i*3+1]=(float)p[1]; P[i*3+2]=(float)p[2];
        }
    }
    char *dpacked, *dpos; double* ws; const int64_t ws_bytes = (int64_t)32 * n * dim * 8;
    cudaMalloc(&dpacked, packed.size()); cudaMalloc(&dpos, pos.size()); cudaMalloc(&ws, ws_bytes);
    cudaMemcpy(dpacked, packed.data(), packed.size(), cudaMemcpyHostToDevice); cudaMemcpy(dpos, pos.data(), pos.size(), cudaMemcpyHostToDevice);
    AccelArgs a{}; a.src = dpacked; a.n_chunks = chunks; a.pos_tgt = dpos; a.n_tgt = n; a.partial = ws; a.eps_sq = 0.01;
    const double ni = (double)n * (double)n;
    printf("N=%lld D=3 %s uniform-mass\n", (long long)n, f64 ? "fp64" : "fp32");
#define CFG(TH, IPT, UN) if (f64) run<ForceF64<3, Q_F64, IPT, TH, true, UN>>("f64 th" #TH " ipt" #IPT " unroll" #UN, a, ws_bytes, ni); \
                         else run<ForceF32<3, Q_F32, IPT, TH, true, UN>>("f32 th" #TH " ipt" #IPT " unroll" #UN, a, ws_bytes, ni);
    CFG(256, 2, 4) CFG(256, 2, 2) CFG(256, 2, 8) CFG(256, 1, 4) CFG(256, 1, 8) CFG(256, 4, 2) CFG(256, 4, 4)
    CFG(128, 2, 4) CFG(128, 4, 4) CFG(128, 4, 2) CFG(128, 1, 8) CFG(512, 1, 4) CFG(512, 2, 4) CFG(512, 2, 2) CFG(384, 2, 4)
    CFG(256, 3, 4) CFG(128, 3, 4) CFG(256, 2, 16) CFG(128, 2, 8)
    return 0;
}
