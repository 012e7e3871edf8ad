// Does compute-sanitizer racecheck model the bar.arrive / bar.sync producer-consumer pattern?  A correctly synchronised
// hand-over between two warps through shared memory; any hazard it reports here is a limitation of the tool.
#include <cstdio>
__global__ void k(double* out, int n) {
    __shared__ double buf[2][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        if (warp == 0) {  // producer
            if (i >= 2) { if ((i - 2) & 1) asm volatile("bar.sync 4, 64;" ::: "memory"); else asm volatile("bar.sync 3, 64;" ::: "memory"); }
            buf[i & 1][lane] = i + lane;
            __threadfence_block();
            __syncwarp();
            if (i & 1) asm volatile("bar.arrive 2, 64;" ::: "memory"); else asm volatile("bar.arrive 1, 64;" ::: "memory");
        } else {          // consumer
            if (i & 1) asm volatile("bar.sync 2, 64;" ::: "memory"); else asm volatile("bar.sync 1, 64;" ::: "memory");
            acc += buf[i & 1][31 - lane];
            __threadfence_block();
            if (i & 1) asm volatile("bar.arrive 4, 64;" ::: "memory"); else asm volatile("bar.arrive 3, 64;" ::: "memory");
        }
    }
    if (warp == 0) { for (int i = max(n - 2, 0); i < n; ++i) { if (i & 1) asm volatile("bar.sync 4, 64;" ::: "memory"); else asm volatile("bar.sync 3, 64;" ::: "memory"); } }
    if (warp == 1) out[lane] = acc;
}
int main() {
    double* d; cudaMalloc(&d, 32 * sizeof(double));
    k<<<1, 64>>>(d, 10);
    double h[32]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s out[0]=%g (expect %g)\n", cudaGetErrorString(cudaGetLastError()), h[0], 10 * 31.0 + 45.0);
    return 0;
}
