// Register-bank cost of packed fp32x2 ops: cycles per warp instruction for operand patterns of the force loop.
#include <cuda_runtime.h>
#include <cstdio>
constexpr int ILP = 6, ITERS = 2048, UNROLL = 8;
template <int OP>
__global__ void __launch_bounds__(256, 3) k(float* out, float seed, unsigned long long* cyc) {
    float2 a[ILP], b[ILP], c[ILP];
    float s1 = seed * 0.25f;
    #pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = make_float2(seed + i, seed - i); b[i] = make_float2(1.0f + 1e-6f * i + 1e-7f * threadIdx.x, 1.0f - 1e-6f * i); c[i] = make_float2(1e-3f * i, -1e-3f * i + 1e-7f * threadIdx.x); }
    long long t0 = clock64();
    for (int it = 0; it < ITERS / UNROLL; ++it) {
        #pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            #pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == 0) a[i] = __ffma2_rn(a[i], b[i], c[i]);                 // 3 distinct 64-bit operands
                if (OP == 1) a[i] = __ffma2_rn(b[i], b[i], a[i]);                 // 2 distinct (d² chain)
                if (OP == 2) a[i] = __fadd2_rn(a[i], make_float2(s1, s1));        // pair + broadcast scalar
                if (OP == 3) a[i] = __fadd2_rn(a[i], c[i]);                       // 2 distinct pairs
                if (OP == 4) a[i] = __fmul2_rn(a[i], b[i]);                       // 2 distinct pairs
                if (OP == 5) a[i] = __ffma2_rn(a[i], b[(i + 1) % ILP], c[(i + 2) % ILP]);   // 3 distinct, rotating
                if (OP == 6) { a[i].x = fmaf(a[i].x, b[i].x, c[i].x); a[i].y = fmaf(a[i].y, b[i].y, c[i].y); }   // 2 scalar FFMA, 3 distinct each
                if (OP == 7) { a[i].x = fmaf(b[i].x, b[i].x, a[i].x); a[i].y = fmaf(b[i].y, b[i].y, a[i].y); }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    #pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (unsigned long long)(t1 - t0);
}
int main() {
    float* out; unsigned long long* d; cudaMalloc(&out, 148 * 3 * 256 * 4); cudaMalloc(&d, 8);
    const char* names[] = {"FFMA2 a*b+c (3 distinct pairs)", "FFMA2 b*b+a (2 distinct pairs)", "FADD2 a+{s,s} (pair + scalar)", "FADD2 a+c (2 pairs)",
                           "FMUL2 a*b (2 pairs)", "FFMA2 rotating 3 distinct", "2x FFMA scalar a*b+c", "2x FFMA scalar b*b+a"};
    for (int op = 0; op < 8; ++op) {
        unsigned long long cyc = 0;
        for (int r = 0; r < 2; ++r) {
            switch (op) { case 0: k<0><<<444, 256>>>(out, 1.0001f, d); break; case 1: k<1><<<444, 256>>>(out, 1.0001f, d); break; case 2: k<2><<<444, 256>>>(out, 1.0001f, d); break;
                case 3: k<3><<<444, 256>>>(out, 1.0001f, d); break; case 4: k<4><<<444, 256>>>(out, 1.0001f, d); break; case 5: k<5><<<444, 256>>>(out, 1.0001f, d); break;
                case 6: k<6><<<444, 256>>>(out, 1.0001f, d); break; case 7: k<7><<<444, 256>>>(out, 1.0001f, d); break; }
            cudaDeviceSynchronize(); cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        }
        // 3 CTAs x 8 warps = 24 warps per SM = 6 per SMSP; each issues ITERS*ILP packed ops (2 scalar ops count as one "pair op")
        double per_op = (double)cyc / ((double)ITERS * ILP * 6);
        printf("%-36s %7.3f cycles per (64-lane-op) instruction per SMSP   [%s]\n", names[op], per_op, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
