// Probe: dependent-issue latencies on the FP64 path of sm_100a (cycles per dependent instruction, one warp).
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k_lat(double* out, long long* cyc, double seed) {
    double x = seed + threadIdx.x * 1e-9, y = seed * 0.5;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) x = fma(x, 0.999999, y);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) x = x * 1.0000001;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // MUFU.RCP64H chain (seed only)
    double r = x;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(r));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    x += r;
    // DMMA chain
    double c0 = x, c1 = y;
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(1e-3), "d"(1e-3));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    x += c0 + c1;
    // independent DFMA throughput (8 chains)
    double z[8];
    for (int j = 0; j < 8; j++) z[j] = x + j;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) z[j] = fma(z[j], 0.999999, y);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    for (int j = 0; j < 8; j++) x += z[j];
    // DSETP + select chain
    t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) x = (x > 0.5) ? x * 0.99 : y;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // shared-memory round trip (store + load dependent)
    __shared__ double sm[64];
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { sm[threadIdx.x & 31] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31] + 1.0; __syncwarp(); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // __syncthreads
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = t1 - t0;
    out[threadIdx.x] = x;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64);
    for (int threads : {32, 64, 256}) {
        k_lat<<<1, threads>>>(out, cyc, 1.25); cudaDeviceSynchronize();
        k_lat<<<1, threads>>>(out, cyc, 1.25); cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads %3d: cycles per op — DFMA dep %.1f | DMUL dep %.1f | MUFU.RCP64H dep %.1f | DMMA dep %.1f | DFMA x8 indep %.1f per 8 | DSETP+sel+DMUL %.1f | STS+LDS+DADD %.1f | bar.sync %.1f\n",
               threads, h[0] / (double)N, h[1] / (double)N, h[2] / (double)N, h[3] / (double)N, h[4] / (double)N, h[5] / (double)N, h[6] / (double)N, h[7] / (double)N);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
