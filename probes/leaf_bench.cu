// Probe: where does the 64x64 leaf spend its time?  Times k_leaf (and phase-disabled variants) in isolation.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../hbetune_rs_b200/csrc/kernels.cuh"
using namespace hbegp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_empty() {}

template <int DBG>
float time_leaf(int B, double* A0, double* A, double* W, int np, double* ldp, int* st, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    size_t bytes = (size_t)B * np * np * 8;
    CK(cudaFuncSetAttribute(k_leaf<double, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)leaf_smem_bytes<double>()));
    float total = 0;
    for (int r = 0; r < reps + 1; r++) {
        CK(cudaMemcpy(A, A0, bytes, cudaMemcpyDeviceToDevice));
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        k_leaf<double, DBG><<<dim3(1, 1, B), 256, leaf_smem_bytes<double>()>>>(A, W, (long)np * np, np, 0, ldp, np / 64, st);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0) total += ms;
    }
    return total / reps * 1e3f;
}

int main() {
    const int np = 1024;
    for (int B : {1, 33, 148, 296}) {
        size_t bytes = (size_t)B * np * np * 8;
        double *A0, *A, *W, *ldp; int* st;
        CK(cudaMalloc(&A0, bytes)); CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&W, bytes)); CK(cudaMalloc(&ldp, B * 16 * 8)); CK(cudaMalloc(&st, B * 4));
        std::vector<double> h((size_t)np * np, 0.0);
        for (int i = 0; i < 64; i++) for (int j = 0; j <= i; j++) h[(size_t)i * np + j] = (i == j) ? 2.0 : 0.5 * exp(-0.1 * (i - j));
        for (int b = 0; b < B; b++) CK(cudaMemcpy(A0 + (size_t)b * np * np, h.data(), (size_t)np * np * 8, cudaMemcpyHostToDevice));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_empty<<<1, 32>>>(); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_empty<<<1, 32>>>(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("B=%3d empty %.1f us | full %.1f us | no-chol %.1f us | no-gj %.1f us | io-only %.1f us\n", B, ms * 1e3,
               time_leaf<0>(B, A0, A, W, np, ldp, st, 10), time_leaf<1>(B, A0, A, W, np, ldp, st, 10),
               time_leaf<2>(B, A0, A, W, np, ldp, st, 10), time_leaf<3>(B, A0, A, W, np, ldp, st, 10));
        cudaFree(A0); cudaFree(A); cudaFree(W); cudaFree(ldp); cudaFree(st);
    }
    return 0;
}
