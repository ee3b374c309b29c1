//! `EstimatorGPRCuda` / `SurrogateModelCuda`: the reference's `EstimatorGPR` / `SurrogateModelGPR`
//! (src/core/gpr.rs) with every kernel evaluation and all linear algebra behind `libhbegp.so`.
//! NOT compiled in this repository (no Rust toolchain in the build image) — see README.md next to this file.
//! Intended location: hbetune/src/core/gpr_cuda.rs, with `pub use` next to the existing exports in src/lib.rs:26-40.

use std::ffi::CStr;
use std::os::raw::{c_int, c_long, c_void};
use std::rc::Rc;

use ndarray::prelude::*;

use hbegp_sys as ffi;

use crate::core::gpr::{estimate_amplitude, Error, EstimatorGPR};       // made pub(crate) in gpr.rs
use crate::core::surrogate_model::SummaryStatistics;
use crate::core::ynormalize::YNormalize;
use crate::gpr::{ConstantKernel, Kernel, Matern, Product, Scalar};
use crate::util::BoundedValue;
use crate::{Estimator, Space, SurrogateModel, RNG};

type ConcreteKernel = Product<ConstantKernel, Matern>;

/// `Scalar` gains: `const HBEGP_DTYPE: c_int` (HBEGP_F64 for f64, HBEGP_F32 for f32) in src/gpr/scalar.rs.
fn dtype<A: Scalar>() -> c_int {
    A::HBEGP_DTYPE
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::hbegp_last_error()) }.to_string_lossy().into_owned()
}

fn check(rc: c_int) {
    // usage / CUDA errors have no counterpart in the reference's `Error` enum: they are bugs or a lost device
    assert!(rc >= 0, "libhbegp: {}", last_error());
}

/// One per process and GPU; not `Sync` (src/gpr/fit.rs:92 already assumes single-threaded model calls).
pub struct Context(*mut ffi::hbegp_ctx);

impl Context {
    pub fn new<A: Scalar>(device: i32) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::hbegp_ctx_create(device, dtype::<A>(), std::ptr::null_mut(), &mut h) });
        Context(h)
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::hbegp_ctx_destroy(self.0) };
    }
}

/// Device-resident part of a model (X, alpha, L^-1).  Shared by clones of the model.
struct DeviceModel(*mut ffi::hbegp_model);
impl Drop for DeviceModel {
    fn drop(&mut self) {
        unsafe { ffi::hbegp_model_destroy(self.0) };
    }
}

#[derive(Clone)]
pub struct SurrogateModelCuda<A: Scalar> {
    kernel: ConcreteKernel,
    noise: BoundedValue<f64>,
    y_norm: YNormalize<A>,
    lml: f64,
    n_features: usize,
    device: Rc<DeviceModel>,
}

impl<A: Scalar> SurrogateModelCuda<A> {
    pub fn kernel(&self) -> &ConcreteKernel {
        &self.kernel
    }

    /// src/gpr/predict.rs:7-52 through hbegp_predict
    fn predict(&self, x: ArrayView2<A>, want_variance: bool) -> (Array1<A>, Option<Array1<A>>) {
        let x = x.as_standard_layout();
        let m = x.nrows();
        let mut mean = Array1::<A>::zeros(m);
        let mut var = if want_variance { Some(Array1::<A>::zeros(m)) } else { None };
        let mut below: c_long = 0;
        let var_ptr = var.as_mut().map_or(std::ptr::null_mut(), |v| v.as_mut_ptr() as *mut c_void);
        check(unsafe {
            ffi::hbegp_predict(self.device.0, m as c_long, x.as_ptr() as *const c_void,
                               mean.as_mut_ptr() as *mut c_void, var_ptr, &mut below)
        });
        self.warn_about_negative_variances(below);
        (mean, var)
    }

    /// src/gpr/predict.rs:39-46: the same stderr message, values from hbegp_predict_warn_values (row order)
    fn warn_about_negative_variances(&self, below: c_long) {
        if below <= 0 {
            return;
        }
        let mut vals = vec![0f64; (below as usize).min(4096)];
        let k = unsafe {
            ffi::hbegp_predict_warn_values(self.device.0, vals.len() as c_int, vals.as_mut_ptr(), std::ptr::null_mut())
        };
        vals.truncate(k.max(0) as usize);
        eprintln!(
            "Variances below 0 were predicted and will be corrected: {}",
            vals.iter().map(|v| format!("{:.2e}", v)).collect::<Vec<_>>().join(", ")
        );
    }

    fn y_norm_ffi(&self) -> ffi::hbegp_ynorm {
        self.y_norm.as_ffi() // amplitude, expected, projection, dtype (ynormalize.rs:7-12)
    }

    /// find_best_candidate_by_ei (src/core/acquisition.rs:177-202) for all candidates in ONE call: prediction, EI,
    /// de-normalisation and the arg-max (`max_by`: the LAST maximum wins) on the device.
    pub fn best_candidate_by_ei(&self, candidates: ArrayView2<A>, fmin: A) -> (usize, A, A) {
        let x = candidates.as_standard_layout();
        let m = x.nrows();
        let (mut mean, mut ei) = (Array1::<A>::zeros(m), Array1::<A>::zeros(m));
        let (mut best, mut below): (c_long, c_long) = (-1, 0);
        check(unsafe {
            ffi::hbegp_predict_mean_ei(self.device.0, &self.y_norm_ffi(), m as c_long, x.as_ptr() as *const c_void,
                                       fmin.into(), mean.as_mut_ptr() as *mut c_void, ei.as_mut_ptr() as *mut c_void,
                                       &mut best, &mut below)
        });
        self.warn_about_negative_variances(below);
        let best = best as usize;
        (best, mean[best], ei[best])
    }

    /// find_best_individual_by_confidence_bound (src/core/minimize.rs:680-714) over all samples in ONE call
    /// (strict `<`: the FIRST minimum wins).
    pub fn best_by_confidence_bound(&self, samples: ArrayView2<A>, cb: A) -> usize {
        let x = samples.as_standard_layout();
        let (mut best, mut below): (c_long, c_long) = (-1, 0);
        check(unsafe {
            ffi::hbegp_predict_confidence_bound(self.device.0, &self.y_norm_ffi(), x.nrows() as c_long,
                                                x.as_ptr() as *const c_void, cb.into(), std::ptr::null_mut(), &mut best,
                                                &mut below)
        });
        self.warn_about_negative_variances(below);
        best as usize
    }
}

impl<A: Scalar> SurrogateModel<A> for SurrogateModelCuda<A> {
    fn length_scales(&self) -> Vec<f64> {
        self.kernel.k2().length_scale().iter().map(BoundedValue::value).collect()
    }

    fn predict_mean_a(&self, x: Array2<A>) -> Array1<A> {
        let (y, _) = self.predict(x.view(), false);
        self.y_norm.project_location_from_normalized(y) // gpr.rs:91
    }

    fn predict_confidence_bound(&self, x: Array1<A>, cb: A) -> A {
        let (mnorm, vnorm) = self.predict(x.view().insert_axis(Axis(0)), true);
        let stdnorm = vnorm.unwrap().mapv(|v| v.sqrt());
        *self.y_norm.project_location_from_normalized(mnorm + stdnorm * cb).first().unwrap() // gpr.rs:104-111
    }

    fn predict_statistics(&self, x: Array1<A>) -> SummaryStatistics<A> {
        // gpr.rs:114-177 unchanged except for the source of (mnorm, vnorm)
        let (mnorm, vnorm) = self.predict(x.view().insert_axis(Axis(0)), true);
        crate::core::gpr::statistics_from_normalized(&self.y_norm, mnorm, vnorm.unwrap()) // body of gpr.rs:131-176 factored out
    }

    fn predict_mean_ei_a(&self, x: Array2<A>, fmin: A) -> (Array1<A>, Array1<A>) {
        // gpr.rs:179-212 in ONE call: prediction, projection of fmin into normalised space (:192-196), per-row
        // expected_improvement in f64 (acquisition.rs:141-171) and de-normalisation of the mean (:210) on the device
        let x = x.as_standard_layout();
        let m = x.nrows();
        let (mut mean, mut ei) = (Array1::<A>::zeros(m), Array1::<A>::zeros(m));
        let mut below: c_long = 0;
        check(unsafe {
            ffi::hbegp_predict_mean_ei(self.device.0, &self.y_norm_ffi(), m as c_long, x.as_ptr() as *const c_void,
                                       fmin.into(), mean.as_mut_ptr() as *mut c_void, ei.as_mut_ptr() as *mut c_void,
                                       std::ptr::null_mut(), &mut below)
        });
        self.warn_about_negative_variances(below);
        (mean, ei)
    }
}

pub struct EstimatorGPRCuda {
    inner: EstimatorGPR, // the reference's builder fields and defaults (gpr.rs:219-236), unchanged
    ctx: Rc<Context>,
}

impl EstimatorGPRCuda {
    pub fn with_context(space: &Space, ctx: Rc<Context>) -> Self {
        EstimatorGPRCuda { inner: <EstimatorGPR as Estimator<f64>>::new(space), ctx }
    }

    /// natural-unit bounds in theta order [noise | c | l_1 .. l_d] (fit.rs:140-144)
    fn natural_bounds(noise: &BoundedValue<f64>, kernel: &ConcreteKernel) -> (Vec<f64>, Vec<f64>) {
        let mut lo = vec![noise.min(), kernel.k1().constant().min()];
        let mut hi = vec![noise.max(), kernel.k1().constant().max()];
        for l in kernel.k2().length_scale() {
            lo.push(l.min());
            hi.push(l.max());
        }
        (lo, hi)
    }

    fn finish<A: Scalar>(
        &self, kernel: ConcreteKernel, noise: BoundedValue<f64>, theta: &[f64], bounds: Option<(&[f64], &[f64])>,
        y_norm: YNormalize<A>, n_features: usize,
    ) -> SurrogateModelCuda<A> {
        let (lo, hi) = bounds.map_or((std::ptr::null(), std::ptr::null()), |(l, h)| (l.as_ptr(), h.as_ptr()));
        let mut model = std::ptr::null_mut();
        let mut lml = 0.0;
        let rc = unsafe {
            ffi::hbegp_model_create((self.ctx).0, self.inner.matern_nu, theta.as_ptr(), lo, hi, &mut model, &mut lml,
                                    std::ptr::null_mut(), std::ptr::null_mut())
        };
        assert!(rc != ffi::HBEGP_NOT_PD, "Kernel matrix must be invertible."); // fit.rs:55
        check(rc);
        SurrogateModelCuda { kernel, noise, y_norm, lml, n_features, device: Rc::new(DeviceModel(model)) }
    }
}

impl<A: Scalar> Estimator<A> for EstimatorGPRCuda {
    type Model = SurrogateModelCuda<A>;
    type Error = Error;

    fn new(space: &Space) -> Self {
        Self::with_context(space, Rc::new(Context::new::<A>(0)))
    }

    fn estimate(&self, x: Array2<A>, y: Array1<A>, prior: Option<&Self::Model>, rng: &mut RNG)
        -> Result<Self::Model, Self::Error>
    {
        let (n, d) = x.dim();
        assert!(y.len() == n, "expected y values for {} observations: {}", n, y);
        let (y_train, y_norm) =
            YNormalize::new_project_into_normalized(y, self.inner.y_projection, self.inner.known_optimum.map(A::from_f));
        let amplitude = estimate_amplitude(y_train.view(), self.inner.amplitude_bounds);
        // gpr.rs:402-427 with the prior's kernel / noise taken from the CUDA model
        let (kernel, noise) = match prior {
            Some(p) => (p.kernel.clone(), p.noise.clone()),
            None => self.inner.default_kernel_and_noise(amplitude)?, // body of get_kernel_or_default's `None` arm
        };
        let mut rng = rng.fork_random_state(); // gpr.rs:276

        let (lo, hi) = Self::natural_bounds(&noise, &kernel);
        let p = lo.len();
        let n_runs = 1 + self.inner.n_restarts_optimizer;
        // gradmin.rs:19-24: run 0 from the current theta, then p inclusive-uniform draws per restart, in order
        let mut starts = vec![noise.value().ln()];
        starts.extend(kernel.theta());
        for _ in 0..self.inner.n_restarts_optimizer {
            for k in 0..p {
                starts.push(rng.uniform(lo[k].ln()..=hi[k].ln()));
            }
        }
        let x = x.as_standard_layout();
        let mut results = vec![ffi::hbegp_run_result::default(); n_runs];
        let mut best_theta = vec![0f64; n_runs * p];
        unsafe {
            check(ffi::hbegp_set_data((self.ctx).0, n as c_long, d as c_int, x.as_ptr() as *const c_void,
                                      y_train.as_ptr() as *const c_void));
            check(ffi::hbegp_fit_runs((self.ctx).0, self.inner.matern_nu, n_runs as c_int, starts.as_ptr(), lo.as_ptr(),
                                      hi.as_ptr(), 150, results.as_mut_ptr(), best_theta.as_mut_ptr())); // gradmin.rs:54
        }
        let best = unsafe { ffi::hbegp_pick_best_run(n_runs as c_int, results.as_ptr()) };
        assert!(best >= 0, "called `Option::unwrap()` on a `None` value"); // fit.rs:161
        let theta = &best_theta[best as usize * p..(best as usize + 1) * p];
        let kernel = kernel.with_clamped_theta(&theta[1..]); // fit.rs:163
        let noise = noise.with_clamped_value(A::from_f(theta[0].exp()).into()); // fit.rs:164 (rounded through A)
        Ok(self.finish(kernel, noise, theta, Some((&lo, &hi)), y_norm, d))
    }

    fn extend(&self, x: Array2<A>, y: Array1<A>, prior: &Self::Model, _rng: &mut RNG)
        -> Result<Self::Model, Self::Error>
    {
        // gpr.rs:293-337 / fit.rs:33-68: one evaluation at the prior's theta, no optimisation, no clamping
        let (n, d) = x.dim();
        assert!(y.len() == n, "expected y values for {} observations: {}", n, y);
        let (y_train, y_norm) =
            YNormalize::new_project_into_normalized(y, self.inner.y_projection, self.inner.known_optimum.map(A::from_f));
        let x = x.as_standard_layout();
        check(unsafe {
            ffi::hbegp_set_data((self.ctx).0, n as c_long, d as c_int, x.as_ptr() as *const c_void,
                                y_train.as_ptr() as *const c_void)
        });
        // hbegp_model_extend takes theta from the prior's device model and appends the new rows to its
        // factorisation when the old rows are an unchanged prefix of x (minimize.rs:629-644 chains the validation
        // samples behind the evaluation history); otherwise it runs the full evaluation.
        let mut model = std::ptr::null_mut();
        let mut lml = 0.0;
        let rc = unsafe {
            ffi::hbegp_model_extend((self.ctx).0, prior.device.0, &mut model, &mut lml, std::ptr::null_mut(),
                                    std::ptr::null_mut(), std::ptr::null_mut())
        };
        assert!(rc != ffi::HBEGP_NOT_PD, "Kernel matrix must be invertible."); // fit.rs:55
        check(rc);
        Ok(SurrogateModelCuda {
            kernel: prior.kernel.clone(), noise: prior.noise.clone(), y_norm, lml, n_features: d,
            device: Rc::new(DeviceModel(model)),
        })
    }
}
