// Microbenchmark: issue cost of Blackwell's packed FP32 instructions (FFMA2 / FADD2 / FMUL2) vs scalar FFMA,
// alone and mixed with other work (ALU / MUFU) to see whether they relieve an issue-bound instruction stream.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void kern(float* out, int iters, float a, float b) {
    float2 acc[8];
    float s[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = threadIdx.x * 0.002f + i;
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    float mu = 0.f;
    int ia[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) ia[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = fmaf(s[i], a, b);
        } else if (MODE == 1) {  // 8 FFMA2 = 16 FMAs
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(acc[i], a2, b2);
        } else if (MODE == 2) {  // 16 scalar FFMA + 8 integer ops (issue-slot competition)
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = fmaf(s[i], a, b);
#pragma unroll
            for (int i = 0; i < 16; ++i) ia[i] = (ia[i] ^ it) + i;  // 2 independent ALU ops per accumulator
        } else if (MODE == 3) {  // 8 FFMA2 + 8 integer ops
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(acc[i], a2, b2);
#pragma unroll
            for (int i = 0; i < 16; ++i) ia[i] = (ia[i] ^ it) + i;
        } else if (MODE == 4 || MODE == 5) {  // representative mix: 16 FMAs + 4 ALU + 2 MUFU per iteration
            if (MODE == 4) {
#pragma unroll
                for (int i = 0; i < 16; ++i) s[i] = fmaf(s[i], a, b);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(acc[i], a2, b2);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) ia[i] = (ia[i] ^ it);
            float m0, m1;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(m0) : "f"(s[0]));
            asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(m1) : "f"(acc[0].x));
            mu += m0 + m1;
        }
    }
    float r = mu;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += (float)ia[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) r += acc[i].x + acc[i].y;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> float run(float* d, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<MODE><<<148 * 8, 128>>>(d, 100, 0.999f, 0.001f);
    cudaEventRecord(e0);
    kern<MODE><<<148 * 8, 128>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 128 * 4);
    const int iters = 200000;
    const char* names[6] = {"16 FFMA", "8 FFMA2 (16 FMA)", "16 FFMA + 32 ALU", "8 FFMA2 + 32 ALU", "16 FFMA + 4 ALU + 2 MUFU + 2 FADD", "8 FFMA2 + 4 ALU + 2 MUFU + 2 FADD"};
    float ms[6] = {run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters), run<4>(d, iters), run<5>(d, iters)};
    for (int m = 0; m < 6; ++m) {
        // warp-instructions issued per SM per cycle assuming 1965 MHz: 32 warps/SM
        double cyc = ms[m] * 1e-3 * 1.965e9;
        printf("%-36s %8.3f ms  cycles/iter/warp-slot %.2f\n", names[m], ms[m], cyc / iters / 8.0);
    }
    return 0;
}
