// Multi-GPU paths of libhbegp.so driven through the C ABI alone (no Python, no torch): restarts and candidate rows
// sharded over 2+ GPUs must give results bit-identical to one GPU.
//   1. single-process handle (hbegp_multi_*): the form a drop-in for the one-process reference needs;
//   2. communicator mode (hbegp_comm_init + *_sharded): one context per rank, here one host thread per rank.
// Exit code 0 = pass, 1 = fail, 77 = skipped (fewer than 2 GPUs).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "hbegp.h"

static int failures = 0;
#define EXPECT(cond, ...)                                         \
    do {                                                          \
        if (!(cond)) {                                            \
            failures++;                                           \
            printf("FAILED %s:%d: %s -- ", __FILE__, __LINE__, #cond); \
            printf(__VA_ARGS__);                                  \
            printf("\n");                                         \
        }                                                         \
    } while (0)
#define OK(call)                                                                        \
    do {                                                                                \
        int rc_ = (call);                                                               \
        if (rc_ != HBEGP_OK) {                                                          \
            printf("%s:%d: %s -> %d: %s\n", __FILE__, __LINE__, #call, rc_, hbegp_last_error()); \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

static bool same_results(int n_runs, int p, const std::vector<hbegp_run_result>& a, const std::vector<double>& ta,
                         const std::vector<hbegp_run_result>& b, const std::vector<double>& tb, const char* what) {
    bool ok = true;
    for (int r = 0; r < n_runs; r++) {
        if (a[r].best_lml != b[r].best_lml || a[r].best_eval != b[r].best_eval || a[r].n_evals != b[r].n_evals ||
            a[r].final_f != b[r].final_f || a[r].status != b[r].status) {
            printf("%s: run %d differs: lml %.17g / %.17g, evals %lld / %lld\n", what, r, a[r].best_lml, b[r].best_lml, a[r].n_evals,
                   b[r].n_evals);
            ok = false;
        }
    }
    if (std::memcmp(ta.data(), tb.data(), sizeof(double) * n_runs * p) != 0) {
        printf("%s: best thetas differ\n", what);
        ok = false;
    }
    return ok;
}

int main(int argc, char** argv) {
    const int G = (argc > 1) ? std::atoi(argv[1]) : 2;
    {   // is device G - 1 there?  (the library reports a bad device index without touching anything)
        hbegp_ctx* probe = nullptr;
        if (hbegp_ctx_create(G - 1, HBEGP_F64, nullptr, &probe) != HBEGP_OK) {
            printf("skipped: the multi-GPU test needs %d GPUs (%s)\n", G, hbegp_last_error());
            return 77;
        }
        hbegp_ctx_destroy(probe);
    }
    const int n = 1500, d = 5, p = d + 2, n_runs = 9, maxeval = 40;
    const long m = 40001;
    unsigned long long rng[4];
    hbegp_rng_seed(2024, rng);
    std::vector<double> x((size_t)n * d), y(n), xs((size_t)m * d);
    for (auto& v : x) v = hbegp_rng_uniform(rng, 0.0, 1.0);
    for (auto& v : xs) v = hbegp_rng_uniform(rng, 0.0, 1.0);
    for (int i = 0; i < n; i++) {
        double s = 0;
        for (int k = 0; k < d; k++) s += std::sin(6.283185307179586 * x[(size_t)i * d + k]);
        y[i] = 1.05 + 0.25 * s + 0.05 * hbegp_rng_uniform(rng, -1.0, 1.0);
    }
    std::vector<double> lo(p, 1e-3), hi(p, 1e3), starts((size_t)n_runs * p, 0.0);
    lo[0] = 1e-2; hi[0] = 1e1; lo[1] = 1e-2; hi[1] = 1e2;
    for (int r = 1; r < n_runs; r++)
        for (int k = 0; k < p; k++) starts[(size_t)r * p + k] = hbegp_rng_uniform(rng, std::log(lo[k]), std::log(hi[k]));

    // ---- reference: one GPU
    hbegp_ctx* c1 = nullptr;
    OK(hbegp_ctx_create(0, HBEGP_F64, nullptr, &c1));
    OK(hbegp_set_data(c1, n, d, x.data(), y.data()));
    std::vector<hbegp_run_result> res1(n_runs);
    std::vector<double> th1((size_t)n_runs * p);
    OK(hbegp_fit_runs(c1, 2.5, n_runs, starts.data(), lo.data(), hi.data(), maxeval, res1.data(), th1.data()));
    const int best = hbegp_pick_best_run(n_runs, res1.data());
    EXPECT(best >= 0, "no run succeeded");
    hbegp_model* m1 = nullptr;
    double lml1 = 0;
    OK(hbegp_model_create(c1, 2.5, &th1[(size_t)best * p], lo.data(), hi.data(), &m1, &lml1, nullptr, nullptr));
    std::vector<double> mean1(m), var1(m);
    long nb1 = 0;
    OK(hbegp_predict(m1, m, xs.data(), mean1.data(), var1.data(), &nb1));
    std::vector<double> lml_b(n_runs), grad_b((size_t)n_runs * p);
    std::vector<int> st_b(n_runs);
    OK(hbegp_lml_grad_batch(c1, 2.5, n_runs, starts.data(), lo.data(), hi.data(), lml_b.data(), grad_b.data(), st_b.data()));

    // ---- 1. single-process handle over G GPUs
    {
        hbegp_multi* mm = nullptr;
        OK(hbegp_multi_create(G, nullptr, HBEGP_F64, &mm));
        EXPECT(hbegp_multi_n_gpus(mm) == G, "n_gpus");
        OK(hbegp_multi_set_data(mm, n, d, x.data(), y.data()));
        std::vector<double> l2(n_runs), g2((size_t)n_runs * p);
        std::vector<int> s2(n_runs);
        OK(hbegp_multi_lml_grad_batch(mm, 2.5, n_runs, starts.data(), lo.data(), hi.data(), l2.data(), g2.data(), s2.data()));
        EXPECT(std::memcmp(l2.data(), lml_b.data(), sizeof(double) * n_runs) == 0 && std::memcmp(g2.data(), grad_b.data(), sizeof(double) * n_runs * p) == 0,
               "multi_lml_grad_batch differs from one GPU");
        std::vector<hbegp_run_result> res2(n_runs);
        std::vector<double> th2((size_t)n_runs * p);
        OK(hbegp_multi_fit_runs(mm, 2.5, n_runs, starts.data(), lo.data(), hi.data(), maxeval, res2.data(), th2.data()));
        EXPECT(same_results(n_runs, p, res1, th1, res2, th2, "multi_fit_runs"), "fit over %d GPUs differs from one GPU", G);
        hbegp_multi_model* mmod = nullptr;
        double lml2 = 0;
        OK(hbegp_multi_model_create(mm, 2.5, &th2[(size_t)best * p], lo.data(), hi.data(), &mmod, &lml2, nullptr, nullptr));
        EXPECT(lml2 == lml1, "model lml %.17g vs %.17g", lml2, lml1);
        std::vector<double> mean2(m), var2(m);
        long nb2 = 0;
        OK(hbegp_multi_predict(mmod, m, xs.data(), mean2.data(), var2.data(), &nb2));
        // the replica on the LAST GPU alone must also give the single-GPU answer (the broadcast model is complete)
        std::vector<double> mean3(1000), var3(1000);
        OK(hbegp_predict(hbegp_multi_model_replica(mmod, G - 1), 1000, xs.data(), mean3.data(), var3.data(), nullptr));
        EXPECT(std::memcmp(mean3.data(), mean1.data(), sizeof(double) * 1000) == 0 && std::memcmp(var3.data(), var1.data(), sizeof(double) * 1000) == 0,
               "broadcast replica differs");
        // row blocks change which 64-row tile a candidate sits in, not its value: every reduction is per candidate row
        EXPECT(std::memcmp(mean2.data(), mean1.data(), sizeof(double) * m) == 0, "sharded mean differs from one GPU");
        EXPECT(std::memcmp(var2.data(), var1.data(), sizeof(double) * m) == 0, "sharded variance differs from one GPU");
        EXPECT(nb2 == nb1, "below-warning count");
        double coll_ms = 0;
        long long ncoll = 0;
        int ver = 0, world = 0;
        OK(hbegp_comm_info(hbegp_multi_ctx(mm, 0), nullptr, &world, &ver, &coll_ms, &ncoll));
        printf("single-process handle over %d GPUs: ok so far (%d failures); NCCL %d, %lld broadcasts, %.3f ms on the device\n", world, failures, ver,
               ncoll, coll_ms);
        hbegp_multi_model_destroy(mmod);
        hbegp_multi_destroy(mm);
    }

    // ---- 2. communicator mode: one context per rank (a thread each), exchange inside the library
    {
        unsigned char id[HBEGP_COMM_ID_BYTES];
        OK(hbegp_comm_unique_id(id));
        std::vector<int> rcs(G, 0);
        std::vector<std::string> why(G);
        std::vector<std::thread> th;
        for (int r = 0; r < G; r++)
            th.emplace_back([&, r] {
                auto bad = [&](const char* what) {
                    rcs[r] = 1;
                    why[r] = std::string(what) + ": " + hbegp_last_error();
                };
                hbegp_ctx* c = nullptr;
                if (hbegp_ctx_create(r, HBEGP_F64, nullptr, &c)) return bad("ctx_create");
                if (hbegp_comm_init(c, G, r, id)) return bad("comm_init");
                if (hbegp_set_data(c, n, d, x.data(), y.data())) return bad("set_data");
                std::vector<double> l(n_runs), g((size_t)n_runs * p);
                std::vector<int> s(n_runs);
                if (hbegp_lml_grad_batch_sharded(c, 2.5, n_runs, starts.data(), lo.data(), hi.data(), l.data(), g.data(), s.data()))
                    return bad("lml_grad_batch_sharded");
                if (std::memcmp(l.data(), lml_b.data(), sizeof(double) * n_runs) || std::memcmp(g.data(), grad_b.data(), sizeof(double) * n_runs * p) ||
                    std::memcmp(s.data(), st_b.data(), sizeof(int) * n_runs)) {
                    rcs[r] = 1;
                    why[r] = "lml_grad_batch_sharded differs from one GPU";
                    return;
                }
                std::vector<hbegp_run_result> res(n_runs);
                std::vector<double> tht((size_t)n_runs * p);
                if (hbegp_fit_runs_sharded(c, 2.5, n_runs, starts.data(), lo.data(), hi.data(), maxeval, r, G, nullptr, nullptr, res.data(), tht.data()))
                    return bad("fit_runs_sharded");
                if (!same_results(n_runs, p, res1, th1, res, tht, "fit_runs_sharded (NCCL)")) {
                    rcs[r] = 1;
                    why[r] = "fit_runs_sharded differs from one GPU";
                    return;
                }
                hbegp_model* mod = nullptr;
                if (hbegp_model_create(c, 2.5, &tht[(size_t)best * p], lo.data(), hi.data(), &mod, nullptr, nullptr, nullptr)) return bad("model_create");
                std::vector<double> mean(m), var(m);
                long nb = 0;
                if (hbegp_predict_sharded(mod, m, xs.data(), mean.data(), var.data(), &nb)) return bad("predict_sharded");
                if (std::memcmp(mean.data(), mean1.data(), sizeof(double) * m) || std::memcmp(var.data(), var1.data(), sizeof(double) * m) || nb != nb1) {
                    rcs[r] = 1;
                    why[r] = "predict_sharded differs from one GPU";
                }
                double ms = 0;
                long long nc = 0;
                hbegp_comm_info(c, nullptr, nullptr, nullptr, &ms, &nc);
                if (r == 0) printf("communicator mode, rank 0 of %d: %lld collectives, %.3f ms on the device\n", G, nc, ms);
                hbegp_model_destroy(mod);
                hbegp_ctx_destroy(c);
            });
        for (auto& t : th) t.join();
        for (int r = 0; r < G; r++) EXPECT(rcs[r] == 0, "rank %d: %s", r, why[r].c_str());
    }
    hbegp_model_destroy(m1);
    hbegp_ctx_destroy(c1);
    printf("%s (%d failures)\n", failures ? "FAILED" : "all ok", failures);
    return failures ? 1 : 0;
}
