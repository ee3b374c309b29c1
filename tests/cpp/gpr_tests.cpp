// C++ port of the reference's library-level GPR tests, run against libhbegp.so through include/hbegp.hpp:
//   src/gpr/predict.rs:54-99           it_works_on_a_simple_case
//   tests/gpr_tests.rs:72-225          describe_gpr (1-D behaviour: fit, interpolation, extrapolation, uncertainty)
// Same data, seeds, bounds and tolerances as the reference.  (The optimiser is the library's bounded L-BFGS,
// not NLopt, so these are behavioural checks exactly like the originals.)
#include <cstdio>
#include <cmath>
#include <functional>
#include <string>
#include <vector>

#include "hbegp.hpp"

using namespace hbegp_cpp;

static int failures = 0, checks = 0;
#define EXPECT(cond, ...)                                              \
    do {                                                               \
        checks++;                                                      \
        if (!(cond)) {                                                 \
            failures++;                                                \
            printf("  FAILED %s:%d: %s -- ", __FILE__, __LINE__, #cond); \
            printf(__VA_ARGS__);                                       \
            printf("\n");                                              \
        }                                                              \
    } while (0)
#define EXPECT_CLOSE(a, b, eps) EXPECT(std::fabs((double)(a) - (double)(b)) <= (eps), "%.6g vs %.6g (eps %.3g)", (double)(a), (double)(b), (double)(eps))

static void run(const char* name, const std::function<void()>& f) {
    printf("test %s\n", name);
    try {
        f();
    } catch (const std::exception& e) {
        failures++;
        printf("  EXCEPTION: %s\n", e.what());
    }
}

struct SimpleModel {  // tests/gpr_tests.rs:11-35 (the space is [lo, hi] -> [0, 1])
    SurrogateModelGPR<double> model;
    double lo, hi;
    double feature(double x) const { return (x - lo) / (hi - lo); }
    double predict(double x) const { return model.predict_mean({feature(x)}); }
    double uncertainty(double x) const { return model.predict_statistics({feature(x)}).std; }
};

static int run_all();

int main() {
    try {
        return run_all();
    } catch (const std::exception& e) {
        printf("fatal: %s\n", e.what());
        return 2;
    }
}

static int run_all() {
    Context ctx(0, HBEGP_F64);

    run("predict.rs::it_works_on_a_simple_case", [&] {
        std::vector<double> xs{0.0, 0.5, 0.5, 1.0}, ys{0.0, 0.8, 1.2, 2.0};
        Product kernel{{BoundedValue(3.0, 0.1, 4.0)}, {2.5, {BoundedValue(1.5, 0.1, 2.0)}}};
        RNG rng = RNG::new_with_seed(938274);
        auto fk = FittedKernel<double>::create(ctx, kernel, xs, 4, 1, ys, rng, 4, BoundedValue(1.0, 0.001, 1.0));
        std::vector<double> px{0.0, 0.25, 0.5, 0.75, 1.0}, var;
        auto mean = predict(fk, px, 5, &var);
        const double want[5] = {0.0, 0.5, 1.0, 1.5, 2.0};
        for (int i = 0; i < 5; i++) {
            EXPECT_CLOSE(mean[i], want[i], 0.1);
            EXPECT_CLOSE(var[i], 0.03, 0.03);
        }
    });

    auto density_model = [&] {  // gpr_tests.rs:84-94
        std::vector<double> xs{0.1, 0.5, 0.5, 0.9}, ys{1.0, 1.8, 2.2, 3.0};
        RNG rng = RNG::new_with_seed(123);
        return SimpleModel{EstimatorGPR(1).estimate<double>(ctx, xs, 4, ys, nullptr, rng), 0.0, 1.0};
    };
    run("with_differing_sampling_density::should_roughly_fit_the_data", [&] {
        auto m = density_model();
        EXPECT_CLOSE(m.predict(0.1), 1.0, 0.1);
        EXPECT_CLOSE(m.predict(0.5), 2.0, 0.1);
        EXPECT_CLOSE(m.predict(0.9), 3.0, 0.1);
    });
    run("with_differing_sampling_density::should_provie_a_reasonable_interpolation", [&] {
        auto m = density_model();
        EXPECT_CLOSE(m.predict(0.3), 1.5, 0.1);
        EXPECT_CLOSE(m.predict(0.7), 2.5, 0.1);
    });
    run("with_differing_sampling_density::should_prefer_a_conservative_extrapolation", [&] {
        auto m = density_model();
        EXPECT_CLOSE(m.predict(0.0), 0.9, 0.1);
        EXPECT_CLOSE(m.predict(1.0), 3.1, 0.1);
    });
    run("with_differing_sampling_density::uncertainty", [&] {
        auto m = density_model();
        EXPECT_CLOSE(m.uncertainty(0.1), m.uncertainty(0.9), 0.05);
        EXPECT(m.uncertainty(0.5) < m.uncertainty(0.1), "%.4g !< %.4g", m.uncertainty(0.5), m.uncertainty(0.1));
    });

    auto unsampled_model = [&] {  // gpr_tests.rs:130-142
        std::vector<double> xs{0.3, 0.5, 0.7}, ys{1.0, 2.0, 1.5};
        RNG rng = RNG::new_with_seed(9372);
        EstimatorGPR est(1);
        est.noise_bounds(1e-5, 1e0).length_scale_bounds({{0.1, 1.0}});
        return SimpleModel{est.estimate<double>(ctx, xs, 3, ys, nullptr, rng), 0.0, 1.0};
    };
    run("with_unsampled_regions::has_low_uncertainty_at_samples", [&] {
        auto m = unsampled_model();
        for (double x : {0.3, 0.5, 0.7}) EXPECT(m.uncertainty(x) < 0.01, "uncertainty(%.1f) = %.4g", x, m.uncertainty(x));
    });
    run("with_unsampled_regions::more_uncertainty_away_from_samples", [&] {
        auto m = unsampled_model();
        const double base = m.uncertainty(0.3);
        for (double x : {0.4, 0.6, 0.0, 1.0}) EXPECT(m.uncertainty(x) > 10.0 * base, "uncertainty(%.1f) = %.4g vs base %.4g", x, m.uncertainty(x), base);
    });

    run("works_in_1d", [&] {  // gpr_tests.rs:173-224: sphere on [-2, 2], 5 points
        std::vector<double> raw{-2.0, -1.0, 0.0, 1.0, 2.0}, xs, ys;
        for (double x : raw) {
            xs.push_back((x + 2.0) / 4.0);
            ys.push_back(x * x);
        }
        RNG rng = RNG::new_with_seed(4531);
        EstimatorGPR est(1);
        est.length_scale_bounds({{1e-2, 1e1}}).noise_bounds(1e-2, 1e1);
        SimpleModel m{est.estimate<double>(ctx, xs, 5, ys, nullptr, rng), -2.0, 2.0};
        for (double x : {-2.0, -1.0, 0.0, 1.0, 2.0, -1.5, -0.5, 1.5}) {
            auto st = m.model.predict_statistics({m.feature(x)});
            const double expected = x * x;
            EXPECT(expected - 0.6 * st.std < st.mean && st.mean < expected + st.std, "x=%.1f expected %.3f predicted %.3f +- %.3f", x,
                   expected, st.mean, st.std);
        }
    });

    run("extend_and_device_acquisition", [&] {
        auto m = density_model();
        std::vector<double> xs{0.1, 0.5, 0.5, 0.9, 0.3}, ys{1.0, 1.8, 2.2, 3.0, 1.5};
        auto ext = EstimatorGPR(1).extend<double>(ctx, xs, 5, ys, m.model);
        EXPECT(ext.length_scales() == m.model.length_scales(), "extend must keep the prior's hyper-parameters");
        std::vector<double> cand;
        for (int i = 0; i <= 100; i++) cand.push_back(i / 100.0);
        long best = -1;
        auto r = ext.predict_mean_ei_a(cand, 101, 1.0, &best);
        EXPECT(best >= 0 && best <= 100, "best = %ld", best);
        for (int i = 0; i <= 100; i++) EXPECT(r.second[i] >= 0.0 && r.second[i] <= r.second[best], "EI[%d] = %g", i, r.second[i]);
        auto one = ext.predict_mean_ei({cand[best]}, 1.0);
        EXPECT_CLOSE(one.second, r.second[best], 1e-12);
    });

    run("bounds_errors", [&] {
        std::vector<double> xs{0.1, 0.9}, ys{1.0, 2.0};
        RNG rng = RNG::new_with_seed(1);
        bool threw = false;
        try {
            EstimatorGPR est(1);
            est.noise_bounds(2.0, 5.0);
            est.estimate<double>(ctx, xs, 2, ys, nullptr, rng);
        } catch (const EstimatorGPR::Error& e) {
            threw = std::string(e.what()).find("noise level") == 0;
        }
        EXPECT(threw, "noise start value outside its bounds must raise Error::NoiseBounds");
    });

    printf("%d checks, %d failures\n", checks, failures);
    return failures ? 1 : 0;
}
