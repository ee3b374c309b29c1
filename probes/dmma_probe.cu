// Probe: FP64 peak on B200 — cuBLAS DGEMM, cuSOLVER potrf, raw DMMA / DFMA issue rates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probes/dmma_probe probes/dmma_probe.cu -lcublas -lcusolver
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int SHAPE>
__global__ void dmma_loop(double* out, int iters) {
    // 8 independent accumulator chains per warp
    double c[8][4];
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = 0.0;
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    double b0 = threadIdx.x * 2e-3, b1 = b0 + 1, b2 = b0 + 2, b3 = b0 + 3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (SHAPE == 0) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a0), "d"(b0));
            } else if (SHAPE == 1) {
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b0));
            } else if (SHAPE == 2) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
            } else if (SHAPE == 3) {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(a4), "d"(a5), "d"(a6), "d"(a7), "d"(b0), "d"(b1), "d"(b2), "d"(b3));
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_loop(double* out, int iters) {
    double c[16];
    for (int i = 0; i < 16; i++) c[i] = i;
    double a = threadIdx.x * 1e-3, b = 1.0000001;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = fma(c[i], b, a);
    }
    double s = 0;
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s sms=%d clock=%d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    double* out; CK(cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double)));
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32 > 1024 ? 1024 : warps * 32;
        int blocks = 148 * (warps * 32 / threads);
        double fl;
        float ms;
        ms = time_ms([&] { dmma_loop<0><<<blocks, threads>>>(out, iters); });
        fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)blocks * threads / 32;
        printf("warps/SM=%2d m8n8k4   : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
        ms = time_ms([&] { dmma_loop<1><<<blocks, threads>>>(out, iters); });
        fl = 2.0 * 16 * 8 * 4 * 8 * iters * (double)blocks * threads / 32;
        printf("warps/SM=%2d m16n8k4  : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
        ms = time_ms([&] { dmma_loop<2><<<blocks, threads>>>(out, iters); });
        fl = 2.0 * 16 * 8 * 8 * 8 * iters * (double)blocks * threads / 32;
        printf("warps/SM=%2d m16n8k8  : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
        ms = time_ms([&] { dmma_loop<3><<<blocks, threads>>>(out, iters); });
        fl = 2.0 * 16 * 8 * 16 * 8 * iters * (double)blocks * threads / 32;
        printf("warps/SM=%2d m16n8k16 : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
        ms = time_ms([&] { dfma_loop<<<blocks, threads>>>(out, iters); });
        fl = 2.0 * 16 * iters * (double)blocks * threads;
        printf("warps/SM=%2d DFMA     : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    }
    // cuBLAS DGEMM
    cublasHandle_t h; cublasCreate(&h);
    for (int n : {2048, 4096, 8192}) {
        double *A, *B, *C; size_t bytes = (size_t)n * n * 8;
        CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
        CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes)); CK(cudaMemset(C, 0, bytes));
        double al = 1.0, be = 0.0;
        float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n); }, 5);
        printf("cublasDgemm TN n=%d: %.3f ms  %.2f TFLOP/s\n", n, ms, 2.0 * n * n * (double)n / ms * 1e-9);
        ms = time_ms([&] { cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, n, n, &al, A, n, &be, C, n); }, 5);
        printf("cublasDsyrk n=%d: %.3f ms  %.2f TFLOP/s\n", n, ms, 1.0 * n * n * (double)n / ms * 1e-9);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    // cuSOLVER potrf / potri for context
    cusolverDnHandle_t sh; cusolverDnCreate(&sh);
    for (int n : {1024, 4096, 16384}) {
        size_t bytes = (size_t)n * n * 8;
        double *A, *A0; CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&A0, bytes));
        std::vector<double> hA((size_t)n * n, 0.0);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) hA[(size_t)i * n + j] = (i == j) ? n + 1.0 : 1.0 / (1.0 + abs(i - j));
        CK(cudaMemcpy(A0, hA.data(), bytes, cudaMemcpyHostToDevice));
        int lwork = 0, lwork2 = 0; cusolverDnDpotrf_bufferSize(sh, CUBLAS_FILL_MODE_LOWER, n, A, n, &lwork);
        cusolverDnDpotri_bufferSize(sh, CUBLAS_FILL_MODE_LOWER, n, A, n, &lwork2);
        if (lwork2 > lwork) lwork = lwork2;
        double* work; CK(cudaMalloc(&work, (size_t)lwork * 8)); int* info; CK(cudaMalloc(&info, 4));
        cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
        float best1 = 1e30f, best2 = 1e30f;
        for (int r = 0; r < 3; r++) {
            CK(cudaMemcpy(A, A0, bytes, cudaMemcpyDeviceToDevice));
            cudaEventRecord(e0);
            cusolverDnDpotrf(sh, CUBLAS_FILL_MODE_LOWER, n, A, n, work, lwork, info);
            cudaEventRecord(e1);
            cusolverDnDpotri(sh, CUBLAS_FILL_MODE_LOWER, n, A, n, work, lwork, info);
            cudaEventRecord(e2); CK(cudaEventSynchronize(e2));
            float m1, m2; cudaEventElapsedTime(&m1, e0, e1); cudaEventElapsedTime(&m2, e1, e2);
            if (m1 < best1) best1 = m1; if (m2 < best2) best2 = m2;
        }
        int hinfo; cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost);
        printf("cusolver n=%d potrf %.3f ms (%.2f TF)  potri %.3f ms (%.2f TF) info=%d\n", n, best1,
               (double)n * n * n / 3 / best1 * 1e-9, best2, 2.0 * n * n * (double)n / 3 / best2 * 1e-9, hinfo);
        cudaFree(A); cudaFree(A0); cudaFree(work); cudaFree(info);
    }
    return 0;
}
