// Probe: DMMA GEMM tile configurations (NT, batched) vs cuBLAS, to pick the production tile shapes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probes/gemm_bench probes/gemm_bench.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cublas_v2.h>
#include "../hbetune_rs_b200/csrc/gemm.cuh"
using namespace hbegp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int BM, int BN, int WM, int WN, bool AK, bool BK, int BKK, int STG>
void run(const char* name, int M, int N, int K, int batch, int kmode, int lower, double* A, double* B, double* C) {
    using Cfg = GemmCfg<double, BM, BN, WM, WN, AK, BK, BKK, STG>;
    auto kern = gemm_kernel<double, BM, BN, WM, WN, AK, BK, BKK, STG>;
    if (Cfg::SMEM_BYTES > 227 * 1024) { printf("%-28s smem too large\n", name); return; }
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
    GemmArgs<double> g{};
    g.A = A; g.B = B; g.C = C; g.lda = K; g.ldb = K; g.ldc = N;
    if (!AK) g.lda = M;
    if (!BK) g.ldb = N;
    g.sA = (long)M * K; g.sB = (long)N * K; g.sC = (long)M * N;
    g.M = M; g.N = N; g.K = K; g.kmode = kmode; g.lower_only = lower; g.alpha = 1.0; g.beta = 0.0; g.rowsumsq = nullptr;
    long tm = M / BM, tn = N / BN;
    long tiles = lower ? tm * (tm + 1) / 2 : tm * tn;
    dim3 grid((unsigned)tiles, 1, batch);
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::THREADS, Cfg::SMEM_BYTES));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES>>>(g); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES>>>(g); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double fl = 2.0 * M * N * (double)K * batch;
    if (kmode != K_FULL) fl *= 0.5;
    if (lower) fl *= 0.5;
    printf("%-28s M=%d N=%d K=%d b=%d kmode=%d lower=%d: %8.3f ms %6.2f TF  (occ %d CTA/SM, %d thr, %zu KB smem, %ld CTAs)\n", name, M, N, K, batch, kmode,
           lower, best, fl / best * 1e-9, occ, Cfg::THREADS, Cfg::SMEM_BYTES / 1024, tiles * batch);
}

int main() {
    const int Mx = 4096; const int batch = 4;
    size_t bytes = (size_t)Mx * Mx * 8 * batch;
    double *A, *B, *C; CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
    std::vector<double> h((size_t)Mx * Mx * batch);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)((i * 2654435761u) % 1000) * 1e-3 - 0.5;
    CK(cudaMemcpy(A, h.data(), bytes, cudaMemcpyHostToDevice)); CK(cudaMemcpy(B, h.data(), bytes, cudaMemcpyHostToDevice));
    struct Shape { int M, N, K, b, kmode, lower; };
    Shape shapes[] = {{2048, 2048, 2048, 4, K_FULL, 0}, {4096, 4096, 4096, 1, K_FULL, 0}, {2048, 2048, 2048, 4, K_LE_N, 0},
                      {1536, 1536, 1536, 8, K_FULL, 0}};
    for (auto s : shapes) {
        printf("---- shape M=%d N=%d K=%d batch=%d kmode=%d lower=%d\n", s.M, s.N, s.K, s.b, s.kmode, s.lower);
#define RUN(BM, BN, WM, WN, BKK, STG) run<BM, BN, WM, WN, true, true, BKK, STG>(#BM "x" #BN " w" #WM "x" #WN " bk" #BKK " s" #STG, s.M, s.N, s.K, s.b, s.kmode, s.lower, A, B, C)
        RUN(64, 64, 32, 32, 16, 3);
        RUN(64, 64, 32, 32, 32, 2);
        RUN(64, 64, 32, 32, 8, 4);
        RUN(64, 64, 32, 32, 8, 6);
        RUN(64, 64, 32, 64, 16, 3);
        RUN(64, 64, 64, 32, 16, 3);
        RUN(64, 128, 32, 64, 16, 3);
        RUN(128, 64, 64, 32, 16, 3);
        RUN(128, 64, 64, 32, 32, 2);
        RUN(128, 128, 64, 64, 16, 3);
        RUN(96, 96, 48, 48, 16, 3);
    }
    // TN layout check on one shape (A m-major, B n-major)
    printf("---- layouts (128x128 w64x32)\n");
    run<128, 128, 64, 32, true, false, 16, 3>("KN", 2048, 2048, 2048, 4, K_FULL, 0, A, B, C);
    run<128, 128, 64, 32, false, false, 16, 3>("NN(lauum)", 2048, 2048, 2048, 4, K_FULL, 0, A, B, C);
    cublasHandle_t hd; cublasCreate(&hd);
    double al = 1, be = 0;
    for (auto s : {2048, 4096}) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        int b = s == 2048 ? 4 : 1;
        cublasDgemmStridedBatched(hd, CUBLAS_OP_T, CUBLAS_OP_N, s, s, s, &al, A, s, (long long)s * s, B, s, (long long)s * s, &be, C, s, (long long)s * s, b);
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int r = 0; r < 5; r++) {
            cudaEventRecord(e0);
            cublasDgemmStridedBatched(hd, CUBLAS_OP_T, CUBLAS_OP_N, s, s, s, &al, A, s, (long long)s * s, B, s, (long long)s * s, &be, C, s, (long long)s * s, b);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("cublas strided batched n=%d b=%d: %.3f ms %.2f TF\n", s, b, best, 2.0 * s * s * (double)s * b / best * 1e-9);
    }
    return 0;
}
