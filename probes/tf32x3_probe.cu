// Probe for the tcgen05 3xTF32 GEMM (hbetune_rs_b200/csrc/gemm_tf32.cuh): correctness of every operand layout /
// k-range / epilogue variant against the FFMA kernel of gemm.cuh, accuracy of both against an FP64 reference, speed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o probes/tf32x3_probe probes/tf32x3_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../hbetune_rs_b200/csrc/gemm_tf32.cuh"

using namespace hbegp;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(2);                                                                            \
        }                                                                                       \
    } while (0)

// C[m][n] = sum_k A(m,k) B(n,k) in FP64 from FP32 inputs; also the sum of |a b| (error scale)
__global__ void ref_kernel(const float* A, const float* B, double* C, double* S, int M, int N, int K, long lda, long ldb, bool ak, bool bk,
                           int kmode) {
    int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
    if (n >= N || m >= M) return;
    double acc = 0, s = 0;
    for (int k = 0; k < K; k++) {
        double a = ak ? A[(long)m * lda + k] : A[(long)k * lda + m];
        double b = bk ? B[(long)n * ldb + k] : B[(long)k * ldb + n];
        acc += a * b;
        s += fabs(a * b);
    }
    C[(long)m * N + n] = acc;
    S[(long)m * N + n] = s;
}

static unsigned long long rng_state = 88172645463325252ULL;
static double urand() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) / 9007199254740992.0;
}

template <bool AK, bool BKM>
static void run_ffma(const GemmArgs<float>& g, int batch) {
    CK((launch_gemm_cfg<float, 64, 64, 32, 32, AK, BKM>(g, batch, 0)));
}

struct Stat {
    double max_rel = 0, rms_rel = 0;
};

static Stat compare(const std::vector<float>& c, const std::vector<double>& ref, const std::vector<double>& scale, bool lower, int M, int N) {
    Stat s;
    double acc = 0;
    long cnt = 0;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            if (lower && (n / 128) > (m / 128)) continue;
            double e = fabs((double)c[(long)m * N + n] - ref[(long)m * N + n]) / (scale[(long)m * N + n] + 1e-300);
            if (e > s.max_rel) s.max_rel = e;
            acc += e * e;
            cnt++;
        }
    s.rms_rel = sqrt(acc / (cnt ? cnt : 1));
    return s;
}

template <bool AK, bool BKM>
static int accuracy_and_speed(int n, bool positive) {
    const int M = n, N = n, K = n;
    std::vector<float> hA((size_t)n * n), hB((size_t)n * n);
    for (auto& v : hA) v = (float)(positive ? urand() : 2 * urand() - 1);
    for (auto& v : hB) v = (float)(positive ? urand() : 2 * urand() - 1);
    float *dA, *dB, *dC1, *dC2;
    double *dR, *dS;
    CK(cudaMalloc(&dA, sizeof(float) * n * n));
    CK(cudaMalloc(&dB, sizeof(float) * n * n));
    CK(cudaMalloc(&dC1, sizeof(float) * n * n));
    CK(cudaMalloc(&dC2, sizeof(float) * n * n));
    CK(cudaMalloc(&dR, sizeof(double) * n * n));
    CK(cudaMalloc(&dS, sizeof(double) * n * n));
    CK(cudaMemcpy(dA, hA.data(), sizeof(float) * n * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), sizeof(float) * n * n, cudaMemcpyHostToDevice));
    GemmArgs<float> g{};
    g.A = dA; g.B = dB; g.lda = g.ldb = g.ldc = n; g.M = M; g.N = N; g.K = K; g.kmode = K_FULL; g.alpha = 1.f; g.beta = 0.f;
    g.C = dC1;
    run_ffma<AK, BKM>(g, 1);
    g.C = dC2;
    CK((tf32::launch<AK, BKM>(g, 1, 0)));
    ref_kernel<<<dim3((N + 127) / 128, M), 128>>>(dA, dB, dR, dS, M, N, K, n, n, AK, BKM, 0);
    CK(cudaDeviceSynchronize());
    std::vector<float> c1((size_t)n * n), c2((size_t)n * n);
    std::vector<double> r((size_t)n * n), sc((size_t)n * n);
    CK(cudaMemcpy(c1.data(), dC1, sizeof(float) * n * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c2.data(), dC2, sizeof(float) * n * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r.data(), dR, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sc.data(), dS, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    Stat s1 = compare(c1, r, sc, false, M, N), s2 = compare(c2, r, sc, false, M, N);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms1 = 0, ms2 = 0;
    const int reps = 5;
    g.C = dC1;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) run_ffma<AK, BKM>(g, 1);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms1, e0, e1));
    g.C = dC2;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) CK((tf32::launch<AK, BKM>(g, 1, 0)));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms2, e0, e1));
    const double fl = 2.0 * n * n * (double)n * reps;
    printf("n=%5d A_k=%d B_k=%d %s | FFMA max %.2e rms %.2e %7.1f TF | 3xTF32 max %.2e rms %.2e %7.1f TF\n", n, (int)AK, (int)BKM,
           positive ? "pos " : "sym ", s1.max_rel, s1.rms_rel, fl / (ms1 * 1e-3) * 1e-12, s2.max_rel, s2.rms_rel, fl / (ms2 * 1e-3) * 1e-12);
    fflush(stdout);
    cudaFree(dA); cudaFree(dB); cudaFree(dC1); cudaFree(dC2); cudaFree(dR); cudaFree(dS);
    return s2.max_rel < 1e-5 ? 0 : 1;
}

// every variant against the FFMA kernel on the same inputs (different summation order: tolerance 2e-6 of sum |a b|)
template <bool AK, bool BKM>
static int variant(const char* name, int M, int N, int K, int kmode, int lower, float alpha, float beta, int batch, bool rowsum) {
    const int ld = 704;  // sub-matrix of a larger allocation (like the recursion's blocks)
    const long stride = (long)ld * ld;
    const size_t tot = (size_t)stride * batch;
    std::vector<float> hA(tot), hB(tot), hC(tot);
    for (auto& v : hA) v = (float)(2 * urand() - 1);
    for (auto& v : hB) v = (float)(2 * urand() - 1);
    for (auto& v : hC) v = (float)(2 * urand() - 1);
    // the triangular k ranges assume a triangular operand (W = L^-1 with exact zeros above the diagonal): the 64- and
    // the 128-wide tiles cut k at different places, which only agrees when the skipped entries are zero
    const long org = 64L * ld + 128;
    for (int b = 0; b < batch; b++)
        for (int r = 0; r < 576 && r < ld - 64; r++)
            for (int c = 0; c < 576 && c < ld - 128; c++) {
                const size_t i = (size_t)b * stride + org + (long)r * ld + c;
                if (kmode == K_LE_N && c > r) hB[i] = 0.f;  // B[n][k], k <= n
                if (kmode == K_GE_N && r < c) hB[i] = 0.f;  // B[k][n], k >= n
                if (kmode == K_LE_M && c > r) hA[i] = 0.f;  // A[m][k], k <= m
                if (kmode == K_GE_M && r < c) hA[i] = 0.f;  // A[k][m], k >= m
            }
    float *dA, *dB, *dC1, *dC2, *dR1, *dR2;
    CK(cudaMalloc(&dA, tot * 4)); CK(cudaMalloc(&dB, tot * 4)); CK(cudaMalloc(&dC1, tot * 4)); CK(cudaMalloc(&dC2, tot * 4));
    const int nt64 = N / 64, nt128 = (N + 127) / 128;
    CK(cudaMalloc(&dR1, (size_t)M * nt64 * batch * 4)); CK(cudaMalloc(&dR2, (size_t)M * nt128 * batch * 4));
    CK(cudaMemcpy(dA, hA.data(), tot * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), tot * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC1, hC.data(), tot * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC2, hC.data(), tot * 4, cudaMemcpyHostToDevice));
    GemmArgs<float> g{};
    const long off = 64L * ld + 128;  // block origin inside the allocation
    g.A = dA + off; g.B = dB + off; g.lda = g.ldb = g.ldc = ld; g.sA = g.sB = g.sC = stride;
    g.M = M; g.N = N; g.K = K; g.kmode = kmode; g.lower_only = lower; g.alpha = alpha; g.beta = beta;
    g.C = dC1 + off;
    if (rowsum) { g.rowsumsq = dR1; g.ld_rs = nt64; g.s_rs = (long)M * nt64; }
    run_ffma<AK, BKM>(g, batch);
    g.C = dC2 + off;
    if (rowsum) { g.rowsumsq = dR2; g.ld_rs = nt128; g.s_rs = (long)M * nt128; }
    cudaError_t le = tf32::launch<AK, BKM>(g, batch, 0);
    if (le != cudaSuccess) { printf("%-28s launch failed: %s\n", name, cudaGetErrorString(le)); return 1; }
    CK(cudaDeviceSynchronize());
    double worst = 0;
    if (!rowsum) {
        std::vector<float> c1(tot), c2(tot);
        CK(cudaMemcpy(c1.data(), dC1, tot * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(c2.data(), dC2, tot * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < tot; i++) {  // includes everything OUTSIDE the block / the lower tiles: must be untouched
            if (lower) {
                // the two kernels cover different tile sets above the diagonal (64- vs 128-wide tiles): compare the lower 128-tiles
                long b = i / stride, r = (i % stride) / ld, c = (i % stride) % ld;
                long rr = r - 64, cc = c - 128;
                (void)b;
                if (rr >= 0 && rr < M && cc >= 0 && cc < N && cc > rr) continue;
            }
            double e = fabs((double)c1[i] - c2[i]);
            if (e > worst) worst = e;
        }
        worst /= sqrt((double)K);
    } else {
        std::vector<float> r1((size_t)M * nt64 * batch), r2((size_t)M * nt128 * batch);
        CK(cudaMemcpy(r1.data(), dR1, r1.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(r2.data(), dR2, r2.size() * 4, cudaMemcpyDeviceToHost));
        for (int b = 0; b < batch; b++)
            for (int m = 0; m < M; m++) {
                double s1 = 0, s2 = 0;
                for (int t = 0; t < nt64; t++) s1 += r1[((size_t)b * M + m) * nt64 + t];
                for (int t = 0; t < nt128; t++) s2 += r2[((size_t)b * M + m) * nt128 + t];
                double e = fabs(s1 - s2) / (fabs(s1) + 1e-30);
                if (e > worst) worst = e;
            }
    }
    const bool ok = worst < 5e-6;
    printf("%-28s A_k=%d B_k=%d M=%d N=%d K=%d batch=%d : %s (%.2e)\n", name, (int)AK, (int)BKM, M, N, K, batch, ok ? "ok" : "MISMATCH", worst);
    fflush(stdout);
    cudaFree(dA); cudaFree(dB); cudaFree(dC1); cudaFree(dC2); cudaFree(dR1); cudaFree(dR2);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    int bad = 0;
    CK((tf32::configure<true, true>()));
    CK((tf32::configure<true, false>()));
    CK((tf32::configure<false, false>()));
    bad += accuracy_and_speed<true, true>(256, false);
    bad += accuracy_and_speed<true, false>(256, false);
    bad += accuracy_and_speed<false, false>(256, false);
    if (bad) { printf("basic layouts wrong: stopping\n"); return 1; }
    // the shapes of the recursion (hbegp.cu merge_node / lauum) and of the predictive variance
    bad += variant<true, true>("panel solve K_LE_N", 256, 384, 384, K_LE_N, 0, 1.f, 0.f, 3, false);
    bad += variant<true, false>("T = L21 W11 K_GE_N", 256, 384, 384, K_GE_N, 0, 1.f, 0.f, 3, false);
    bad += variant<true, true>("trailing update lower", 384, 384, 256, K_FULL, 1, -1.f, 1.f, 3, false);
    bad += variant<true, false>("W21 = -W22 T K_LE_M", 384, 256, 384, K_LE_M, 0, -1.f, 0.f, 3, false);
    bad += variant<false, false>("K^-1 = W^T W K_GE_M lower", 576, 576, 576, K_GE_M, 1, 1.f, 0.f, 2, false);
    bad += variant<true, true>("partial tiles (M, N % 128 = 64)", 192, 320, 320, K_LE_N, 0, 1.f, 0.f, 2, false);
    bad += variant<true, false>("partial tiles K_LE_M", 192, 320, 192, K_LE_M, 0, -1.f, 0.f, 1, false);
    bad += variant<true, true>("row sums of squares", 256, 576, 576, K_LE_N, 0, 1.f, 0.f, 1, true);
    for (int n : {1024, 4096}) {
        bad += accuracy_and_speed<true, true>(n, false);
        bad += accuracy_and_speed<true, true>(n, true);
        bad += accuracy_and_speed<true, false>(n, true);
        bad += accuracy_and_speed<false, false>(n, true);
    }
    if (argc > 1) bad += accuracy_and_speed<false, false>(8192, true);
    printf(bad ? "FAILED (%d)\n" : "all ok\n", bad);
    return bad ? 1 : 0;
}
