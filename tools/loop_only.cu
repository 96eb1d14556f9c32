// Upper bound of the force inner loop: run ForceF32/ForceF64::chunk() on ONE resident shared-memory chunk over and over
// (no TMA, no mbarrier traffic, no tail) and report interactions/s.  Compare with the streaming kernel to see what the
// pipeline costs, and with 148*4*64/(2*ops) per clock to see what the instruction mix itself costs.
#define NB_TUNE_HARNESS
#include "../nbody_cosmological_simulation_b200/csrc/accel.cu"
#include <cstdio>
extern "C" int64_t nb_chunk_sources(int dtype) { return dtype == NB_F32 ? 256 : 128; }
extern "C" int64_t nb_num_chunks(int64_t n, int dtype) { int64_t cs = nb_chunk_sources(dtype); return (n + cs - 1) / cs; }

template <class Consumer>
__global__ void __launch_bounds__(Consumer::THREADS + 32) loop_kernel(AccelArgs a, int reps) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < chunk_bytes(Consumer::DIM) / 4; i += blockDim.x)
        reinterpret_cast<float*>(smem)[i] = 0.001f * (i % 977) + 0.5f;
    __syncthreads();
    if (threadIdx.x >= Consumer::THREADS) return;
    Consumer cons;
    cons.init(a, nullptr);
    for (int r = 0; r < reps; ++r) cons.chunk(smem, r);
    cons.store(a);
}

template <class Consumer>
void run(const char* name, AccelArgs a, int reps, double per_chunk_sources) {
    int sms = 148, occ = 0;
    auto k = loop_kernel<Consumer>;
    const int smem = chunk_bytes(Consumer::DIM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, Consumer::THREADS + 32, smem + 36 * 1024);   // same residency as the streaming kernel
    const int grid = sms * (occ > 3 ? 3 : occ);
    a.n_tgt = (int64_t)grid * Consumer::THREADS * Consumer::TARGETS_PER_THREAD;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<grid, Consumer::THREADS + 32, smem>>>(a, reps); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); k<<<grid, Consumer::THREADS + 32, smem>>>(a, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    const double inter = (double)a.n_tgt * reps * per_chunk_sources;
    printf("%-40s grid %4d  %8.3f ms  %6.3f T inter/s  (%s)\n", name, grid, best, inter / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    AccelArgs a{}; a.eps_sq = 0.01; a.n_chunks = 1;
    const int64_t cap = 148 * 4 * 1024 * 4;
    void *pos, *ws; cudaMalloc(&pos, cap * 3 * 8); cudaMemset(pos, 0, cap * 3 * 8); cudaMalloc(&ws, cap * 3 * 8);
    a.pos_tgt = pos; a.partial = (double*)ws;
    run<ForceF32<3, Q_F32, 2, 256, true, 4>>("f32 D3 uniform (11 packed ops)", a, 4000, 256);
    run<ForceF32<3, Q_F32, 2, 256, false, 4>>("f32 D3 general (12 packed ops)", a, 4000, 256);
    run<ForceF32<2, Q_F32, 2, 256, true, 4>>("f32 D2 uniform (8 packed ops)", a, 4000, 256);
    run<ForceF32<3, Q_F32, 4, 256, true, 2>>("f32 D3 uniform ipt4", a, 2000, 256);
    run<ForceF32<3, Q_F32, 1, 256, true, 8>>("f32 D3 uniform ipt1", a, 4000, 256);
    run<ForceF64<3, Q_F64, 2, 256, true, 2>>("f64 D3 uniform (15 ops)", a, 2000, 128);
    run<ForceF64<3, Q_F64, 2, 256, false, 2>>("f64 D3 general (16 ops)", a, 2000, 128);
    return 0;
}
