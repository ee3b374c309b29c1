// Probe: does a short burst of DMMAs pay a warm-up penalty after the FP64 tensor path has been idle?
// One warp (and 8 warps) time a burst of 2 dependent + 4 independent DMMAs after spinning for `idle` cycles on
// (a) integer work only, (b) DFMA work.  Prints cycles per burst.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>  // 0: integer spin, 1: DFMA spin, 2: spin with one dummy DMMA every ~200 cycles
__global__ void k(long long* out, int idle, double seed) {
    double c[4][2], x = seed + threadIdx.x * 1e-9, keep0 = 0, keep1 = 0;
    for (int s = 0; s < 4; s++) c[s][0] = c[s][1] = seed;
    long long total = 0;
    for (int rep = 0; rep < 20; rep++) {
        long long t0 = clock64();
        if (MODE == 0) { while (clock64() - t0 < idle) { } }
        else if (MODE == 1) { while (clock64() - t0 < idle) { x = fma(x, 0.999999, 1e-9); } }
        else { long long last = t0; while (clock64() - t0 < idle) { if (clock64() - last > 200) { dmma(keep0, keep1, 1e-3, 1e-3); last = clock64(); } } }
        __syncthreads();
        long long t1 = clock64();
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int s = 0; s < 4; s++) dmma(c[s][0], c[s][1], 1e-3 + x * 1e-12, 1e-3);
        double sink = 0;
        for (int s = 0; s < 4; s++) sink += c[s][0] + c[s][1];
        long long t2 = clock64() + (long long)(sink == 1.2345);
        if (rep >= 4) total += t2 - t1;
    }
    if (threadIdx.x == 0) out[0] = total / 16;
    if (threadIdx.x == 1) out[1] = (long long)(x + keep0 + keep1);
}
int main() {
    long long* d; cudaMalloc(&d, 64); long long h[2];
    for (int threads : {32, 256})
        for (int idle : {0, 100, 300, 1000, 3000, 10000}) {
            k<0><<<1, threads>>>(d, idle, 1.0); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); long long a = h[0];
            k<1><<<1, threads>>>(d, idle, 1.0); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); long long b = h[0];
            k<2><<<1, threads>>>(d, idle, 1.0); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); long long c = h[0];
            printf("threads %3d idle %5d cycles: burst of 8 DMMAs (4 chains x 2) takes %lld cycles after an integer spin, %lld after a DFMA spin, %lld with a dummy DMMA every 200 cycles\n", threads, idle, a, b, c);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
