// Bounded limited-memory BFGS with reverse communication (host side).
//
// Stands in for NLopt's `Algorithm::Lbfgs` as configured by src/util/gradmin.rs:35-60 (box bounds,
// maxeval evaluations, no x/f tolerances): the NLopt C sources (nlopt crate 0.5.1, Luksan PLIS) are
// absent from the reference tree, so the trajectory of the real optimiser cannot be reproduced here.
// The interface is ask/tell so that many independent runs can advance in lockstep and share one batched
// GPU evaluation per step (the reference drives its runs one after another, gradmin.rs:19-30).
//
// Method: projected quasi-Newton.  Variables sitting on a bound with the gradient pushing outwards are
// frozen; the L-BFGS two-loop recursion runs on the free ones; trial points are projected back into the
// box; the step is accepted on the Armijo condition measured along the projected displacement, otherwise
// shrunk by safeguarded quadratic interpolation.  An objective value of +inf (kernel matrix not positive
// definite, src/gpr/fit.rs:103-113) is an ordinary failed trial.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>

namespace hbegp {

// Stopping tolerances of every BoundedLbfgs created afterwards (process-wide; hbegp_lbfgs_set_tolerances).  The
// reference sets only maxeval on NLopt (gradmin.rs:52-54); NLopt's own L-BFGS still stops on a stalled objective
// (Luksan PLIS: two consecutive iterations without progress) and on a vanishing gradient, which these defaults stand
// in for.  A value <= 0 switches the rule off (maxeval-only stopping).
struct LbfgsTolerances {
    double ftol = 1e-11, gtol = 1e-8;
};
inline LbfgsTolerances& lbfgs_tolerances() {
    static LbfgsTolerances t;
    return t;
}

class BoundedLbfgs {
public:
    BoundedLbfgs(int n, const double* x0, const double* lo, const double* hi, int maxeval, int memory = 10)
        : n_(n), m_(memory), maxeval_(maxeval), lo_(lo, lo + n), hi_(hi, hi + n), x_(n), g_(n), xt_(x0, x0 + n),
          d_(n), free_(n, 1), ftol_(lbfgs_tolerances().ftol), gtol_(lbfgs_tolerances().gtol) {
        for (int i = 0; i < n_; i++) xt_[i] = std::min(std::max(xt_[i], lo_[i]), hi_[i]);
        done_ = (maxeval_ <= 0);
    }

    bool done() const { return done_; }
    int evals() const { return evals_; }
    // Point to evaluate next (inside the box).
    const double* ask() const { return xt_.data(); }
    const double* x() const { return have_x_ ? x_.data() : xt_.data(); }
    double f() const { return f_; }

    // Feeds f(ask()) and its gradient; returns true when the run has finished.
    bool tell(double ft, const double* gt) {
        if (done_) return true;
        evals_++;
        const bool finite = std::isfinite(ft);
        if (!have_x_) {
            // first evaluation
            x_ = xt_;
            f_ = ft;
            have_x_ = true;
            if (!finite) return finish();  // nothing to descend from
            g_.assign(gt, gt + n_);
            return next_direction(true);
        }
        // line-search trial
        double dec = 0.0;  // g . (xt - x)
        for (int i = 0; i < n_; i++) dec += g_[i] * (xt_[i] - x_[i]);
        if (!(dec < 0.0)) {
            // the projection bent the quasi-Newton step into a non-descent displacement: Armijo and the quadratic
            // interpolation below both assume a negative slope, so this is a failed direction, not a trial to judge
            if (!S_.empty() || !steepest_) {
                S_.clear();
                Y_.clear();
                rho_.clear();
                return next_direction(true);
            }
            return finish();
        }
        if (finite && ft <= f_ + c1_ * dec) {
            // accept
            std::vector<double> s(n_), y(n_);
            double sy = 0, ss = 0, yy = 0;
            for (int i = 0; i < n_; i++) {
                s[i] = xt_[i] - x_[i];
                y[i] = gt[i] - g_[i];
                sy += s[i] * y[i];
                ss += s[i] * s[i];
                yy += y[i] * y[i];
            }
            if (sy > 1e-10 * std::sqrt(ss * yy) && sy > 0) {
                if ((int)S_.size() == m_) {
                    S_.erase(S_.begin());
                    Y_.erase(Y_.begin());
                    rho_.erase(rho_.begin());
                }
                S_.push_back(s);
                Y_.push_back(y);
                rho_.push_back(1.0 / sy);
            }
            const double fprev = f_;
            x_ = xt_;
            f_ = ft;
            g_.assign(gt, gt + n_);
            const double scale = std::max(std::max(std::fabs(fprev), std::fabs(f_)), 1.0);
            if (ftol_ > 0 && fprev - f_ <= ftol_ * scale) stall_++;
            else stall_ = 0;
            if (stall_ >= 2) return finish();
            return next_direction(false);
        }
        // reject: shrink
        ls_iter_++;
        double anew;
        if (finite) {
            const double denom = 2.0 * (ft - f_ - dec);
            anew = (denom > 0) ? -dec * alpha_ / denom : 0.5 * alpha_;
            anew = std::min(std::max(anew, 0.1 * alpha_), 0.5 * alpha_);
        } else {
            anew = 0.25 * alpha_;
        }
        alpha_ = anew;
        if (ls_iter_ > 25 || alpha_ < 1e-18) {
            if (!S_.empty() || !steepest_) {
                // quasi-Newton direction failed: drop the history and retry along the projected gradient
                S_.clear();
                Y_.clear();
                rho_.clear();
                return next_direction(true);
            }
            return finish();
        }
        return make_trial();
    }

private:
    bool finish() {
        done_ = true;
        return true;
    }

    bool next_direction(bool force_steepest) {
        if (evals_ >= maxeval_) return finish();
        // active set
        double pgmax = 0.0;
        int nfree = 0;
        for (int i = 0; i < n_; i++) {
            bool at_lo = x_[i] <= lo_[i] && g_[i] > 0.0;
            bool at_hi = x_[i] >= hi_[i] && g_[i] < 0.0;
            free_[i] = !(at_lo || at_hi);
            if (free_[i]) {
                nfree++;
                pgmax = std::max(pgmax, std::fabs(g_[i]));
            }
        }
        if (nfree == 0 || pgmax <= (gtol_ > 0 ? gtol_ : 0.0)) return finish();
        steepest_ = force_steepest || S_.empty();
        // two-loop recursion on the free variables
        std::vector<double> q(n_);
        for (int i = 0; i < n_; i++) q[i] = free_[i] ? g_[i] : 0.0;
        if (!steepest_) {
            const int k = (int)S_.size();
            std::vector<double> a(k);
            for (int j = k - 1; j >= 0; j--) {
                double sq = 0, sy = 0;
                for (int i = 0; i < n_; i++)
                    if (free_[i]) {
                        sq += S_[j][i] * q[i];
                        sy += S_[j][i] * Y_[j][i];
                    }
                if (!(sy > 0)) {
                    a[j] = 0;
                    continue;
                }
                a[j] = sq / sy;
                for (int i = 0; i < n_; i++)
                    if (free_[i]) q[i] -= a[j] * Y_[j][i];
            }
            {
                double sy = 0, yy = 0;
                for (int i = 0; i < n_; i++)
                    if (free_[i]) {
                        sy += S_[k - 1][i] * Y_[k - 1][i];
                        yy += Y_[k - 1][i] * Y_[k - 1][i];
                    }
                const double gamma = (sy > 0 && yy > 0) ? sy / yy : 1.0;
                for (int i = 0; i < n_; i++) q[i] *= gamma;
            }
            for (int j = 0; j < k; j++) {
                double yq = 0, sy = 0;
                for (int i = 0; i < n_; i++)
                    if (free_[i]) {
                        yq += Y_[j][i] * q[i];
                        sy += S_[j][i] * Y_[j][i];
                    }
                if (!(sy > 0)) continue;
                const double beta = yq / sy;
                for (int i = 0; i < n_; i++)
                    if (free_[i]) q[i] += (a[j] - beta) * S_[j][i];
            }
        }
        double gd = 0, gnorm1 = 0;
        for (int i = 0; i < n_; i++) {
            d_[i] = free_[i] ? -q[i] : 0.0;
            gd += g_[i] * d_[i];
            if (free_[i]) gnorm1 += std::fabs(g_[i]);
        }
        if (!(gd < 0) || !std::isfinite(gd)) {
            // not a descent direction: fall back to the projected gradient
            steepest_ = true;
            gd = 0;
            for (int i = 0; i < n_; i++) {
                d_[i] = free_[i] ? -g_[i] : 0.0;
                gd += g_[i] * d_[i];
            }
        }
        alpha_ = steepest_ ? std::min(1.0, 1.0 / gnorm1) : 1.0;
        ls_iter_ = 0;
        return make_trial();
    }

    bool make_trial() {
        if (evals_ >= maxeval_) return finish();
        bool moved = false;
        for (int i = 0; i < n_; i++) {
            double v = x_[i] + alpha_ * d_[i];
            v = std::min(std::max(v, lo_[i]), hi_[i]);
            xt_[i] = v;
            if (v != x_[i]) moved = true;
        }
        if (!moved) return finish();
        return false;
    }

    int n_, m_, maxeval_;
    std::vector<double> lo_, hi_, x_, g_, xt_, d_;
    std::vector<char> free_;
    std::vector<std::vector<double>> S_, Y_;
    std::vector<double> rho_;
    double f_ = std::numeric_limits<double>::infinity();
    double alpha_ = 1.0;
    const double c1_ = 1e-4;
    double ftol_, gtol_;
    int evals_ = 0, ls_iter_ = 0, stall_ = 0;
    bool have_x_ = false, done_ = false, steepest_ = true;
};

// ---- src/core/random.rs restated: Xoshiro256** (rand_xoshiro 0.4.0) + rand 0.7.2 Uniform<f64> inclusive
struct Xoshiro256 {
    unsigned long long s[4];
    static unsigned long long rotl(unsigned long long x, int k) { return (x << k) | (x >> (64 - k)); }
    unsigned long long next() {
        const unsigned long long result = rotl(s[1] * 5ULL, 7) * 9ULL;
        const unsigned long long t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    static void seed(unsigned long long seed, unsigned long long out[4]) {  // SplitMix64 (seed_from_u64)
        unsigned long long x = seed;
        for (int i = 0; i < 4; i++) {
            x += 0x9E3779B97F4A7C15ULL;
            unsigned long long z = x;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
            out[i] = z ^ (z >> 31);
        }
    }
    static double from_bits(unsigned long long b) {
        double d;
        __builtin_memcpy(&d, &b, 8);
        return d;
    }
    static unsigned long long to_bits(double d) {
        unsigned long long b;
        __builtin_memcpy(&b, &d, 8);
        return b;
    }
    double uniform_inclusive(double low, double high) {  // gradmin.rs:23 `rng.uniform(lo..=hi)`
        const double max_rand = from_bits((~0ULL >> 12) | (1023ULL << 52)) - 1.0;
        double scale = (high - low) / max_rand;
        while (scale * max_rand + low > high) scale = from_bits(to_bits(scale) - 1);
        const double v12 = from_bits((next() >> 12) | (1023ULL << 52));
        return (v12 - 1.0) * scale + low;
    }
};

}  // namespace hbegp
