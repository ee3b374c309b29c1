// Probe: the register-resident 8-wide-panel leaf (leaf_core_nb8) against the barrier-per-pivot leaf (leaf_core):
// agreement of W = L^-1 and of the log-determinant partials, and time per launch of k_leaf / k_node128.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probes/leaf2_bench probes/leaf2_bench.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <type_traits>
#include "../hbetune_rs_b200/csrc/kernels.cuh"
#include "../hbetune_rs_b200/csrc/node_mma.cuh"
using namespace hbegp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <typename T, int LV, bool NODE>
float run(int B, const T* A0, T* A, T* W, int np, T* ldp, int* st, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    size_t bytes = (size_t)B * np * np * sizeof(T);
    size_t smem = NODE ? node128_smem_bytes<T>() : leaf_smem_bytes<T>();
    if (NODE) CK(cudaFuncSetAttribute(k_node128<T, LV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CK(cudaFuncSetAttribute(k_leaf<T, 0, LV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float total = 0;
    for (int r = 0; r < reps + 1; r++) {
        CK(cudaMemcpy(A, A0, bytes, cudaMemcpyDeviceToDevice));
        CK(cudaMemset(W, 0xff, bytes));
        CK(cudaMemset(st, 0, B * sizeof(int)));
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        if (NODE) k_node128<T, LV><<<dim3(1, 1, B), 256, smem>>>(A, W, (long)np * np, np, 0, ldp, np / 64, st);
        else k_leaf<T, 0, LV><<<dim3(1, 1, B), 256, smem>>>(A, W, (long)np * np, np, 0, ldp, np / 64, st);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0) total += ms;
    }
    return total / reps * 1e3f;
}

float run_v2(int B, const double* A0, double* A, double* W, int np, double* ldp, int* st, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    size_t bytes = (size_t)B * np * np * sizeof(double), smem = node128_v2_smem_bytes();
    CK(cudaFuncSetAttribute(k_node128_v2<false, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_node128_v2<true, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        long long* prof; CK(cudaMalloc(&prof, B * 64 * 8)); CK(cudaMemset(prof, 0, B * 64 * 8));
        CK(cudaMemcpy(A, A0, bytes, cudaMemcpyDeviceToDevice));
        k_node128_v2<true, double><<<dim3(1, 1, B), 256, smem>>>(A, W, (long)np * np, np, 0, ldp, np / 64, st, prof);
        CK(cudaDeviceSynchronize());
        long long h[64]; CK(cudaMemcpy(h, prof, 64 * 8, cudaMemcpyDeviceToHost));
        const char* names[] = {"load", "chol1", "lvl0_1", "lvls1", "L21", "syrk+T", "chol2", "lvl0_2", "lvls2", "W21", "store+logdet"};
        printf("   v2 phases (cycles, matrix 0):");
        for (int i = 0; i < 11; i++) printf(" %s %lld", names[i], h[i + 1] - h[i]);
        printf(" | total %lld\n", h[11] - h[0]);
        for (int leaf = 0; leaf < 2; leaf++) {
            const long long* q = h + (leaf ? 40 : 16);
            printf("   leaf %d panel iterations (cycles):", leaf + 1);
            for (int P = 0; P < 8; P++) printf(" %lld", q[P + 1] - q[P]);  // prologue (panel 0), then iterations 0..6
            printf("\n");
        }
        cudaFree(prof);
    }
    float total = 0;
    for (int r = 0; r < reps + 1; r++) {
        CK(cudaMemcpy(A, A0, bytes, cudaMemcpyDeviceToDevice));
        CK(cudaMemset(W, 0xff, bytes));
        CK(cudaMemset(st, 0, B * sizeof(int)));
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        k_node128_v2<false, double><<<dim3(1, 1, B), 256, smem>>>(A, W, (long)np * np, np, 0, ldp, np / 64, st, nullptr);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0) total += ms;
    }
    return total / reps * 1e3f;
}

template <typename T, int DBG, int LV>
float time_dbg(int B, const T* A0, T* A, T* W, int np, T* ldp, int* st, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    size_t smem = leaf_smem_bytes<T>();
    CK(cudaFuncSetAttribute(k_leaf<T, DBG, LV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float total = 0;
    for (int r = 0; r < reps + 1; r++) {
        CK(cudaMemcpy(A, A0, (size_t)B * np * np * sizeof(T), cudaMemcpyDeviceToDevice));
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        k_leaf<T, DBG, LV><<<dim3(1, 1, B), 256, smem>>>(A, W, (long)np * np, np, 0, ldp, np / 64, st);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0) total += ms;
    }
    return total / reps * 1e3f;
}

template <typename T>
void study(const char* name) {
    const int np = 128;
    printf("==== %s\n", name);
    for (int B : {1, 33, 148}) {
        size_t elems = (size_t)B * np * np, bytes = elems * sizeof(T);
        T *A0, *A, *W0, *W1, *ldp0, *ldp1; int *st0, *st1;
        CK(cudaMalloc(&A0, bytes)); CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&W0, bytes)); CK(cudaMalloc(&W1, bytes));
        CK(cudaMalloc(&ldp0, B * 2 * sizeof(T))); CK(cudaMalloc(&ldp1, B * 2 * sizeof(T)));
        CK(cudaMalloc(&st0, B * 4)); CK(cudaMalloc(&st1, B * 4));
        // SPD test blocks: Matern-like kernel matrix of random 3-d points + noise (condition number ~1e3..1e5), one
        // indefinite matrix (b == 5) to check the failure flag
        std::vector<T> h(elems);
        srand(7);
        for (int b = 0; b < B; b++) {
            std::vector<double> pts(np * 3);
            for (auto& v : pts) v = rand() / (double)RAND_MAX;
            for (int i = 0; i < np; i++)
                for (int j = 0; j < np; j++) {
                    double d2 = 0;
                    for (int k = 0; k < 3; k++) { double t = (pts[i * 3 + k] - pts[j * 3 + k]) / 0.4; d2 += t * t; }
                    double r = sqrt(5.0 * d2);
                    double v = 1.7 * (1 + r + r * r / 3) * exp(-r) + (i == j ? 0.01 : 0.0);
                    if (b == 5 && i == j && i == 70) v = -1.0;
                    h[(size_t)b * np * np + (size_t)i * np + j] = (T)v;
                }
        }
        CK(cudaMemcpy(A0, h.data(), bytes, cudaMemcpyHostToDevice));
        printf("B=%3d k_leaf phases: old full/no-chol/no-inv/io %.1f %.1f %.1f %.1f | nb8 %.1f %.1f %.1f %.1f us\n", B,
               time_dbg<T, 0, 0>(B, A0, A, W0, np, ldp0, st0, 10), time_dbg<T, 1, 0>(B, A0, A, W0, np, ldp0, st0, 10),
               time_dbg<T, 2, 0>(B, A0, A, W0, np, ldp0, st0, 10), time_dbg<T, 3, 0>(B, A0, A, W0, np, ldp0, st0, 10),
               time_dbg<T, 0, 1>(B, A0, A, W0, np, ldp0, st0, 10), time_dbg<T, 1, 1>(B, A0, A, W0, np, ldp0, st0, 10),
               time_dbg<T, 2, 1>(B, A0, A, W0, np, ldp0, st0, 10), time_dbg<T, 3, 1>(B, A0, A, W0, np, ldp0, st0, 10));
        for (int node = 0; node < 3; node++) {
            float t0, t1;
            if (node == 2) {
                if constexpr (std::is_same<T, double>::value) { t0 = run<T, 0, true>(B, A0, A, W0, np, ldp0, st0, 10); t1 = run_v2(B, A0, A, W1, np, ldp1, st1, 10); }
                else continue;
            } else if (node) { t0 = run<T, 0, true>(B, A0, A, W0, np, ldp0, st0, 10); t1 = run<T, 1, true>(B, A0, A, W1, np, ldp1, st1, 10); }
            else { t0 = run<T, 0, false>(B, A0, A, W0, np, ldp0, st0, 10); t1 = run<T, 1, false>(B, A0, A, W1, np, ldp1, st1, 10); }
            std::vector<T> w0(elems), w1(elems), l0(B * 2), l1(B * 2);
            std::vector<int> s0(B), s1(B);
            CK(cudaMemcpy(w0.data(), W0, bytes, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(w1.data(), W1, bytes, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(l0.data(), ldp0, B * 2 * sizeof(T), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(l1.data(), ldp1, B * 2 * sizeof(T), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(s0.data(), st0, B * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(s1.data(), st1, B * 4, cudaMemcpyDeviceToHost));
            const int ext = node ? 128 : 64;
            double maxd = 0, maxw = 0, maxl = 0; int stdiff = 0, nfail = 0;
            for (int b = 0; b < B; b++) {
                stdiff += s0[b] != s1[b]; nfail += s1[b];
                if (s0[b]) continue;
                for (int i = 0; i < ext; i++)
                    for (int j = 0; j < ext; j++) {
                        double a = w0[(size_t)b * np * np + (size_t)i * np + j], c = w1[(size_t)b * np * np + (size_t)i * np + j];
                        if (!(fabs(a - c) <= maxd)) maxd = fabs(a - c);
                        if (fabs(a) > maxw) maxw = fabs(a);
                    }
                for (int q = 0; q < (node ? 2 : 1); q++) maxl = fmax(maxl, fabs((double)l0[b * 2 + q] - (double)l1[b * 2 + q]));
            }
            printf("B=%3d %-9s old %.1f us | nb8 %.1f us | max|dW| %.3g (max|W| %.3g) max|d logdet| %.3g status diffs %d (failed %d)\n", B,
                   node == 2 ? "node_v2" : node ? "k_node128" : "k_leaf", t0, t1, maxd, maxw, maxl, stdiff, nfail);
        }
        cudaFree(A0); cudaFree(A); cudaFree(W0); cudaFree(W1); cudaFree(ldp0); cudaFree(ldp1); cudaFree(st0); cudaFree(st1);
    }
}

int main() {
    study<double>("f64");
    study<float>("f32");
    return 0;
}
