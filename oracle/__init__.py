"""CPU oracle for the hbetune GP surrogate hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy + SciPy LAPACK) of the reference's
``src/gpr`` arithmetic and of the thin adapter around it (``src/core/gpr.rs``,
``src/core/ynormalize.rs``, ``src/util/gradmin.rs``).  Every function cites the
reference file:line it follows.  It exists to *check* the CUDA path:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
  / ``--impl reference`` legs may import it;
* the product path (``hbetune_rs_b200`` / ``libhbegp.so``) never imports, links or
  executes anything in here and has no CPU fallback.

Pinning status
--------------
* Kernel matrices and theta-gradients are pinned against the reference's own
  sklearn-derived golden vectors (``src/gpr/matern_kernel.rs:189-253``,
  ``src/gpr/product_kernel.rs:120-169``), ``cdist`` (``matern_kernel.rs:285-305``),
  ``outer`` (``lml.rs:105-119``) and ``clamp_negative_variance``
  (``predict.rs:129-149``); see ``tests/test_oracle_golden.py``.
* LML value, LML gradient, Cholesky, alpha, K^-1, fitted theta and mean/variance
  are NOT pinned by any reference test tighter than +-0.03 ("parity unpinned" by
  the reference itself), and the reference cannot be built
  here (no cargo/rustc).  They are pinned instead against the implementation the
  reference restates and took its own goldens from -- scikit-learn's
  GaussianProcessRegressor with ConstantKernel * Matern + WhiteKernel (part of this
  image): LML to 1e-11, its log-space gradient to 1e-8, predictive mean to 1e-10 and
  variance (after swapping sklearn's noise term for the reference's 1e-5) to 1e-9,
  for nu in {0.5, 1.5, 2.5}; see ``tests/test_sklearn_pin.py`` (which checks the
  CUDA path against sklearn the same way on the GPU).  Plus self-consistency checks
  (finite-difference gradient, K K^-1 = I).  The fitted theta stays unpinned (NLopt).
* Third-party arithmetic that is absent from /root/reference and restated from the
  published algorithms: LAPACK potrf/potrs/potri (OpenBLAS via SciPy; the reference
  pins openblas-src 0.7.0), ndarray 0.13 ``sum`` (8-lane unrolled fold),
  rand 0.7.2 ``Uniform<f64>`` inclusive sampling, rand_xoshiro 0.4.0
  ``Xoshiro256StarStar`` (SplitMix64 seeding).  NLopt's L-BFGS (nlopt 0.5.1,
  Luksan PLIS) is NOT restated: its source is absent; the optimiser is supplied
  by the caller (the host library's own bounded L-BFGS is used on both sides).
"""
