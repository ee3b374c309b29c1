// Probe: the early-fragment GEMM main loop (-DHBEGP_EARLY=1) against the production loop, same tile, 3 and 4 stages.
// Build both: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DHBEGP_EARLY=1] -o probes/gemm_early[_1] probes/gemm_early.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../hbetune_rs_b200/csrc/gemm.cuh"
using namespace hbegp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <bool AK, bool BK, int STG>
void run(const char* name, int M, int N, int K, int batch, int kmode, int lower, double* A, double* B, double* C) {
    using Cfg = GemmCfg<double, 64, 64, 32, 32, AK, BK, 16, STG>;
    auto kern = gemm_kernel<double, 64, 64, 32, 32, AK, BK, 16, STG>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
    GemmArgs<double> g{};
    g.A = A; g.B = B; g.C = C; g.lda = AK ? K : M; g.ldb = BK ? K : N; g.ldc = N;
    g.sA = (long)M * K; g.sB = (long)N * K; g.sC = (long)M * N;
    g.M = M; g.N = N; g.K = K; g.kmode = kmode; g.lower_only = lower; g.alpha = 1.0; g.beta = 0.0; g.rowsumsq = nullptr;
    long tm = M / 64, tn = N / 64;
    long tiles = lower ? tm * (tm + 1) / 2 : tm * tn;
    dim3 grid((unsigned)tiles, 1, batch);
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::THREADS, Cfg::SMEM_BYTES));
    CK(cudaMemset(C, 0, (size_t)M * N * 8 * batch));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES>>>(g); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES>>>(g); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    std::vector<double> h((size_t)M * N);
    CK(cudaMemcpy(h.data(), C, h.size() * 8, cudaMemcpyDeviceToHost));
    double cs = 0; for (size_t i = 0; i < h.size(); i += 7) cs += h[i] * (double)((i % 13) + 1);
    double fl = 2.0 * M * N * (double)K * batch;
    if (kmode != K_FULL) fl *= 0.5;
    if (lower) fl *= 0.5;
    printf("EARLY=%d %-10s s%d M=%d K=%d b=%d kmode=%d lower=%d: %8.3f ms %6.2f TF (occ %d) checksum %.10e\n", HBEGP_EARLY, name, STG, M, K, batch,
           kmode, lower, best, fl / best * 1e-9, occ, cs);
}

int main() {
    const int Mx = 4096; const int batch = 4;
    size_t bytes = (size_t)Mx * Mx * 8 * batch;
    double *A, *B, *C; CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
    std::vector<double> h((size_t)Mx * Mx * batch);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)((i * 2654435761u) % 1000) * 1e-3 - 0.5;
    CK(cudaMemcpy(A, h.data(), bytes, cudaMemcpyHostToDevice)); CK(cudaMemcpy(B, h.data(), bytes, cudaMemcpyHostToDevice));
    run<true, true, 3>("KK", 2048, 2048, 2048, 4, K_FULL, 0, A, B, C);
    run<true, true, 4>("KK", 2048, 2048, 2048, 4, K_FULL, 0, A, B, C);
    run<true, false, 3>("KN", 2048, 2048, 2048, 4, K_FULL, 0, A, B, C);
    run<true, false, 4>("KN", 2048, 2048, 2048, 4, K_FULL, 0, A, B, C);
    run<false, false, 3>("NN", 2048, 2048, 2048, 4, K_FULL, 0, A, B, C);
    run<false, false, 4>("NN", 2048, 2048, 2048, 4, K_FULL, 0, A, B, C);
    run<false, false, 3>("NN lauum", 4096, 4096, 4096, 4, K_GE_M, 1, A, B, C);
    run<false, false, 4>("NN lauum", 4096, 4096, 4096, 4, K_GE_M, 1, A, B, C);
    run<true, true, 3>("KK trsm", 2048, 2048, 2048, 4, K_LE_N, 0, A, B, C);
    run<true, true, 4>("KK trsm", 2048, 2048, 2048, 4, K_LE_N, 0, A, B, C);
    return 0;
}
