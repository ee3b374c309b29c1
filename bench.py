#!/usr/bin/env python
"""Benchmark of the GP surrogate hot path (BASELINE.json metric: GP fit LML evals/s + Cholesky FP64
TFLOP/s + predict candidates/s), north-star shape: n = 4096, d = 16, 64 restarts (B = 65 thetas per
lockstep optimiser step), m = 2^20 candidates.

One JSON line on stdout (rank 0).  A "step" is ONE lockstep optimiser step of the fit: a batched
LML + gradient evaluation of all B thetas (src/gpr/lml.rs:29-79 x B), sharded over the ranks with an
all-gather of the per-theta results.  `value` = evaluations per second with X, y resident in HBM;
`e2e` = the same through the C ABI with host buffers (X, y, theta uploaded and results read back every
step).  Prediction throughput, per-phase rooflines and the whole fit + predict wall time are reported in
the same line.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # long spellings for use under torchrun, whose own parser treats --n / --m as ambiguous abbreviations
    ap.add_argument("--n", "--train-n", dest="n", type=int, default=4096)
    ap.add_argument("--d", "--feat-d", dest="d", type=int, default=16)
    ap.add_argument("--restarts", type=int, default=64)
    ap.add_argument("--m", "--cand-m", dest="m", type=int, default=1 << 20)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-full-fit", action="store_true", help="skip the whole fit+predict wall-time leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--maxeval", type=int, default=150)
    ap.add_argument("--fit-shard", default="balanced", choices=["balanced", "static"],
                    help="multi-GPU restart loop: per-round balancing (hbegp_fit_runs_sharded) or a static split of the runs")
    ap.add_argument("--only", default="all", choices=["all", "fit-step", "predict"],
                    help="profiling aid: run just the timed fit steps or just the timed prediction, print a short line")
    return ap.parse_args()


def synth(n, d, seed=1, A=np.float64):
    """X ~ U[0,1)^{n x d}; y = sum_k sin(2 pi x_k) + 0.1 N(0,1), normalised like the linear YNormalize
    (ynormalize.rs:168-173) so that mean(y) = 1.05.  (Same generator as tests/util.py; kept here so that the GPU
    arm of the benchmark imports nothing from tests/ or oracle/.)"""
    rng = np.random.default_rng(seed)
    x = rng.random((n, d))
    y = np.sin(2 * np.pi * x).sum(axis=1) + 0.1 * rng.standard_normal(n)
    y = y - y.min()
    y = y / (y.mean() if y.mean() > 0 else 1.0) + 0.05
    return x.astype(A), y.astype(A)


def workload(args):
    """X ~ U[0,1)^{n x d}, y = sum sin(2 pi x_k) + 0.1 N(0,1) normalised like the
    linear YNormalize; bounds as EstimatorGPR::new except noise in [1e-2, 1e1]."""
    A = np.float64 if args.dtype == "f64" else np.float32
    x, y = synth(args.n, args.d, seed=1, A=A)
    p = args.d + 2
    lo = np.array([1e-2] + [max(np.quantile(y.astype(np.float64), 0.1) ** 2 * args.n, 2e-5) / 2] + [1e-3] * args.d)
    hi = np.array([1e1] + [2 * float((y.astype(np.float64) ** 2).sum())] + [1e3] * args.d)
    rng = np.random.default_rng(7)
    B = args.restarts + 1
    # thetas of a typical mid-fit optimiser step: moderate noise / amplitude / length scales
    thetas = np.empty((B, p))
    thetas[:, 0] = rng.uniform(math.log(0.03), math.log(1.0), B)
    thetas[:, 1] = rng.uniform(math.log(0.5), math.log(5.0), B)
    thetas[:, 2:] = rng.uniform(math.log(0.3), math.log(3.0), (B, args.d))
    xs = np.random.default_rng(2).random((args.m, args.d)).astype(A)
    return A, x, y, lo, hi, thetas, xs


def base_config(args, B):
    return {"workload": f"ns_fit_n{args.n}_d{args.d}_B{B}", "n": args.n, "d": args.d, "restarts": args.restarts, "B": B,
            "m": args.m}


def ncu_traffic():
    """DRAM bytes per launch from the committed `ncu --set full` captures (profiles/r01_ncu_traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
    except (OSError, ValueError):
        return {}


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for s in self.samples:
            f = [t.strip() for t in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if smax and v > 0.5 * smax] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def oracle_eval_seconds(x, y, theta, A):
    """One reference-faithful LML + gradient evaluation on the host (oracle: full-square kernel,
    materialised (n, n, d+1) gradient tensor, explicit potri), timed.  The ONLY place bench.py touches oracle/
    (cpu_baseline leg and --impl reference)."""
    from tests.util import oracle_lml
    t0 = time.perf_counter()
    res = oracle_lml(theta, x, y, A=A)
    dt = time.perf_counter() - t0
    return dt, res


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (restatement oracle — the Rust
    crate cannot be built here: no cargo/rustc) on the host cores; each step is ONE of the B evaluations."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    A, x, y, lo, hi, thetas, xs = workload(args)
    cores = os.cpu_count()
    budget = 240.0
    times = []
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        dt, _ = oracle_eval_seconds(x, y, thetas[i % len(thetas)], A)
        if i >= args.warmup:
            times.append(dt)
        # bounded sample: keep the whole run within a few minutes
        if time.perf_counter() - t_start > budget and len(times) >= 1:
            break
    ms = 1e3 * float(np.mean(times))
    value = 1e3 / ms
    sample = (f"{len(times)} timed of {args.steps} requested steps; each step = 1 of the {len(thetas)} LML+gradient "
              f"evaluations of one optimiser step (n={args.n}, d={args.d}), restatement oracle (NumPy + OpenBLAS), "
              "not the Rust binary")
    print(json.dumps({
        "impl": "reference", "metric": "gp_fit_lml_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": base_config(args, len(thetas)),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import hbetune_rs_b200 as h
    from hbetune_rs_b200 import dist as hd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hbetune_rs_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the single JSON line: at NCCL_DEBUG=WARN / VERSION / INFO NCCL prints its version banner
        # on stdout; unset means no banner, and anything it does log goes to stderr
        if "HBEGP_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["HBEGP_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    A, x, y, lo, hi, thetas, xs = workload(args)
    B, p = thetas.shape
    n, d, m = args.n, args.d, args.m
    side = torch.cuda.Stream()  # a real (non-default) stream: torch events on it bracket the library's kernels
    torch.cuda.set_stream(side)
    stream = torch.cuda.current_stream().cuda_stream
    assert stream != 0
    ctx = h.Context(local_rank, h.F64 if args.dtype == "f64" else h.F32, stream=stream)
    ctx.set_data(x, y)
    mine = hd.owned_runs(B, rank, world)
    my_thetas = thetas[mine]
    counts = [len(hd.owned_runs(B, r, world)) for r in range(world)]

    def step_resident():
        lml, grad, status = ctx.lml_grad_batch(my_thetas, lo=lo, hi=hi)
        local = np.concatenate([lml[:, None], grad, status[:, None].astype(np.float64)], axis=1)
        return hd.all_gather_array(local, counts)  # per-restart LML exchange (NCCL over NVLink)

    x_pin = torch.from_numpy(x).pin_memory().numpy()  # pinned host staging for the end-to-end leg
    y_pin = torch.from_numpy(y).pin_memory().numpy()

    def step_e2e():
        ctx.set_data(x_pin, y_pin)  # host -> device copy of X, y every step (theta goes up / results come back inside)
        return step_resident()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps, ctx.launch_count - l0, out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if args.only == "fit-step":
        ms_step, launches, _ = timed(step_resident, args.steps, args.warmup)
        if rank == 0:
            print(json.dumps({"only": "fit-step", "ms_per_step": ms_step, "evals_per_s": B / (ms_step * 1e-3),
                              "gpu_launches": int(launches), "clocks": sampler.stop()}), flush=True)
        ctx.close()
        return
    if args.only != "predict":
        ms_step, launches, gathered = timed(step_resident, args.steps, args.warmup)
        ms_e2e, _, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))

    # ---- prediction: mean + variance of m candidates, rows sharded over ranks
    lo_r, hi_r = hd.row_block(m, rank, world)
    theta_model = np.array([math.log(0.1), 0.0] + [math.log(0.5)] * d)
    model = ctx.model(theta_model, want_alpha=False)
    tdt = torch.float64 if A == np.float64 else torch.float32
    xs_dev = torch.from_numpy(xs[lo_r:hi_r]).cuda()
    mean_dev = torch.empty(hi_r - lo_r, dtype=tdt, device="cuda")
    var_dev = torch.empty(hi_r - lo_r, dtype=tdt, device="cuda")

    def predict_resident():
        model.predict_device(hi_r - lo_r, xs_dev.data_ptr(), mean_dev.data_ptr(), var_dev.data_ptr())

    xs_pin = torch.from_numpy(xs[lo_r:hi_r]).pin_memory()

    def predict_e2e():
        mean, var = model.predict(xs_pin.numpy(), want_variance=True, warn=False)
        return hd.all_gather_array(np.stack([mean, var], axis=1).astype(np.float64), [row_cnt(r) for r in range(world)])

    def row_cnt(r):
        a, b = hd.row_block(m, r, world)
        return b - a

    def predict_mean_resident():
        model.predict_device(hi_r - lo_r, xs_dev.data_ptr(), mean_dev.data_ptr(), None)

    psteps = max(2, min(args.steps, 3))
    ms_pred, pred_launches, _ = timed(predict_resident, psteps, 1)
    ms_pred_mean, _, _ = timed(predict_mean_resident, psteps, 1)
    if args.only == "predict":
        if rank == 0:
            print(json.dumps({"only": "predict", "ms": ms_pred, "candidates_per_s": m / (ms_pred * 1e-3),
                              "tflops_mn2": float(hi_r - lo_r) * n * n / (ms_pred * 1e-3) * 1e-12,
                              "gpu_launches": int(pred_launches), "clocks": sampler.stop()}), flush=True)
        model.close()
        ctx.close()
        return
    ms_pred_e2e, _, _ = timed(predict_e2e, psteps, 1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-phase timings on rank 0's shard (CUDA events on the context's stream, inside the library)
    phases = None
    if rank == 0:
        reps = 2
        t0 = ctx.bench_phase(my_thetas, 0, reps)
        t1 = ctx.bench_phase(my_thetas, 1, reps)
        t2 = ctx.bench_phase(my_thetas, 2, reps)
        t3 = ctx.bench_phase(my_thetas, 3, reps)
        t5 = ctx.bench_phase(my_thetas, 5, max(reps, 3))  # only the K^-1 = W^T W launches (one per stream group)
        phases = {"assemble_ms": t0, "factor_inverse_ms": t1 - t0, "alpha_kinv_ms": t2 - t1, "grad_finish_ms": t3 - t2,
                  "eval_ms": t3, "kinv_gemm_ms": t5, "batch": len(mine)}

    # ---- FP64 tensor peak measured live: cuBLAS DGEMM 8192^3 through torch.matmul
    peak_tf = None
    if rank == 0:
        torch.backends.cuda.matmul.allow_tf32 = False  # the f32 path computes in true FP32 (FFMA), so does its denominator
        a = torch.randn(8192, 8192, dtype=tdt, device="cuda")
        b = torch.randn(8192, 8192, dtype=tdt, device="cuda")
        torch.matmul(a, b)
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        peak_tf = 2 * 8192.0 ** 3 / best * 1e-9
        del a, b

    # ---- the whole north-star job once: fit (B lockstep runs, <= maxeval evaluations each) + predict
    full = None
    if not args.no_full_fit:
        bv = h.BoundedValue
        kernel = h.Product(h.ConstantKernel(bv(math.sqrt(lo[1] * hi[1]), lo[1], hi[1])),
                           h.Matern(2.5, [bv(1.0, 1e-3, 1e3)] * d))
        barrier()
        t0 = time.perf_counter()
        fk = h.FittedKernel.new(ctx, kernel, x, y, h.RNG.new_with_seed(1), args.restarts, bv(1.0, 1e-2, 1e1),
                                maxeval=args.maxeval,
                                shard=None if world == 1 else (hd.BalancedFit() if args.fit_shard == "balanced" else hd.sharded_fit_runs))
        barrier()
        t_fit = time.perf_counter() - t0
        t0 = time.perf_counter()
        mean, var = fk.model.predict(xs_pin.numpy(), want_variance=True, warn=False)
        if world > 1:
            hd.all_gather_array(np.stack([mean, var], axis=1).astype(np.float64), [row_cnt(r) for r in range(world)])
        barrier()
        t_pred = time.perf_counter() - t0
        full = {"fit_s": max_over_ranks(t_fit), "predict_s": max_over_ranks(t_pred), "n_evals": int(fk.n_evals),
                "lml": fk.lml, "noise": fk.noise.value, "amplitude": fk.kernel.k1.constant.value}
        full["fit_predict_s"] = full["fit_s"] + full["predict_s"]

    if rank == 0:
        evals_per_s = B / (ms_step * 1e-3)
        out = {
            "metric": "gp_fit_lml_evals_per_s", "value": evals_per_s, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": dict(base_config(args, B),
                           sharding=f"restarts and candidate rows over {world} rank(s); one n x n factorisation per GPU",
                           l2=f"inputs larger than L2: per-step working set {B // world * 2 * n * n * x.itemsize / 1e9:.1f} GB per GPU vs 126 MB"),
            "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "evals/s",
                    "h2d_bytes_per_step": int(x.nbytes + y.nbytes + thetas.nbytes),
                    "d2h_bytes_per_step": int(B * (p + 1) * 8 + B * 4)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "predict": {"candidates_per_s": m / (ms_pred * 1e-3), "ms": ms_pred, "m": m,
                        "e2e_candidates_per_s": m / (ms_pred_e2e * 1e-3), "e2e_ms": ms_pred_e2e,
                        "h2d_bytes": int(xs.nbytes), "d2h_bytes": int(2 * m * xs.itemsize), "gpu_launches": int(pred_launches)},
            "fit_predict": full,
        }
        nb = len(mine)
        flops_fi = 2.0 * n ** 3 / 3.0 * nb  # L and L^-1 together (potrf n^3/3 + trtri n^3/3)
        tf_fi = flops_fi / (phases["factor_inverse_ms"] * 1e-3) * 1e-12
        tf_eval = 1.0 * n ** 3 * nb / (phases["eval_ms"] * 1e-3) * 1e-12
        tf_var = float(hi_r - lo_r) * n * n / (ms_pred * 1e-3) * 1e-12
        # Headline roofline: the largest single launch of the dominant kernel, gemm_kernel (DMMA.8x8x4) -- the
        # K^-1 = W^T W product (n^3/3 FLOP per matrix, one launch per stream group), timed live with CUDA events on
        # the library's stream (hbegp_bench_phase 5).  `traffic` is the ncu DRAM byte count of one such launch.
        traffic = ncu_traffic()
        tf_kinv = (n ** 3 / 3.0) * nb / (phases["kinv_gemm_ms"] * 1e-3) * 1e-12
        kt = traffic.get("kinv_gemm", {})
        out["roofline"] = {
            "bound": "tensor" if args.dtype == "f64" else "fp32-fma",
            "kernel": ("gemm_kernel<double,64,64,32,32,false,false> (DMMA.8x8x4)" if args.dtype == "f64" else
                       "gemm_kernel<float,64,64,32,32,false,false> (FFMA micro-kernel; TF32 is not parity-safe here)")
                      + ": K^-1 = W^T W on the lower tiles",
            "achieved": tf_kinv, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf_kinv / peak_tf,
            "traffic": kt.get("traffic") if args.dtype == "f64" and n == 4096 else None,
            "algorithmic_bytes": kt.get("algorithmic_bytes") if args.dtype == "f64" and n == 4096 else None,
            "traffic_note": f"DRAM bytes of one ncu --set full launch over {kt.get('matrices', '?')} matrices "
                            "(profiles/r01_kinv_gemm_early_ncu_raw.csv); algorithmic = 8 n (n + 1) bytes per matrix "
                            "(W lower read once, K^-1 lower written once)",
            "launch_ms": phases["kinv_gemm_ms"], "matrices": nb,
            "peak_source": ("cuBLAS DGEMM 8192^3 (torch.matmul f64) measured in this run; MEASURED_PEAKS.json has no FP64 entry"
                            if args.dtype == "f64" else
                            "cuBLAS SGEMM 8192^3 in true FP32 (torch.matmul f32, TF32 off) measured in this run"),
            "algorithmic": "n^3 / 3 FLOP per matrix (lower tiles of W^T W, K restricted to k >= row tile)",
        }
        out["phases"] = phases
        t_dense = (phases["factor_inverse_ms"] + phases["alpha_kinv_ms"]) * 1e-3
        tf_dense = 1.0 * n ** 3 * nb / t_dense * 1e-12
        out["cholesky_fp64_tflops"] = tf_fi
        out["lml_eval_fp64_tflops"] = tf_eval
        out["predict_var_fp64_tflops"] = tf_var
        out["rooflines"] = {
            "factor_inverse_phase": {"bound": "tensor", "achieved": tf_fi, "peak": peak_tf, "frac": tf_fi / peak_tf, "unit": "TFLOP/s",
                                     "what": "2 n^3 / 3 FLOP per evaluation (Cholesky factor n^3/3 + its triangular inverse n^3/3) over the "
                                             "whole recursion phase: k_node128 / k_leaf chains + 4 gemm_kernel launches per node"},
            "lml_eval_n3": {"achieved": tf_eval, "peak": peak_tf, "frac": tf_eval / peak_tf, "unit": "TFLOP/s"},
            "dense_phases_n3": {"achieved": tf_dense, "peak": peak_tf, "frac": tf_dense / peak_tf, "unit": "TFLOP/s",
                                "what": "all gemm_kernel + k_leaf launches (factor, inverse, K^-1 = W^T W): n^3 FLOP per evaluation"},
            "kinv_gemm": dict({"bound": "tensor", "unit": "TFLOP/s", "peak": peak_tf,
                               "algorithmic": "n^3/3 FLOP, 8 n^2 bytes (read W lower, write K^-1 lower) per matrix"},
                              **traffic.get("kinv_gemm", {})),
            "predict_var_gemm": dict({"bound": "tensor", "unit": "TFLOP/s", "peak": peak_tf,
                                      "algorithmic": "m n^2 FLOP per chunk of m candidates"},
                                     **traffic.get("predict_var_gemm", {})),
            "predict_var_mn2": {"achieved": tf_var * world, "peak": peak_tf * world, "frac": tf_var / peak_tf, "unit": "TFLOP/s"},
            "assemble_hbm": {"achieved": nb * (8.0 * n * d + x.itemsize * n * (n + 1) / 2) / (phases["assemble_ms"] * 1e-3) * 1e-9,
                             "peak": peaks().get("hbm_gbs"), "unit": "GB/s"},
        }
        # the element-wise kernels are reported against HBM as BASELINE.json asks, with the FP64 FMA-pipe
        # bound that actually limits them beside it (approximate op counts per matrix entry: distances 3d, Matern +
        # exp + sqrt ~40; the gradient contraction re-forms the distances and adds 3d per entry).
        alu_peak = 36.7 if args.dtype == "f64" else 2 * 36.7  # TFLOP/s, profiles/r01_fp64_peak_probe.log (raw DFMA issue rate)
        isz = x.itemsize
        rl = out["rooflines"]
        rl["assemble_hbm"]["alu"] = {"achieved": nb * (3 * d + 40) * n * (n + 1) / 2 / (phases["assemble_ms"] * 1e-3) * 1e-12,
                                     "peak": alu_peak, "unit": "TFLOP/s (approx. op count)"}
        rl["grad_contract_hbm"] = {
            "achieved": nb * (isz * n * (n + 1) / 2 + isz * n * (d + 1)) / (phases["grad_finish_ms"] * 1e-3) * 1e-9,
            "peak": peaks().get("hbm_gbs"), "unit": "GB/s",
            "what": "k_grad_contract + k_finish: reads the lower tiles of K^-1 once, scaled X and alpha",
            "alu": {"achieved": nb * (6 * d + 60) * n * (n + 1) / 2 / (phases["grad_finish_ms"] * 1e-3) * 1e-12,
                    "peak": alu_peak, "unit": "TFLOP/s (approx. op count)"}}
        mm = hi_r - lo_r
        rl["predict_mean_hbm"] = {
            "achieved": (isz * mm * d + isz * n * d + isz * mm) / (ms_pred_mean * 1e-3) * 1e-9,
            "peak": peaks().get("hbm_gbs"), "unit": "GB/s", "ms": ms_pred_mean, "candidates_per_s": mm / (ms_pred_mean * 1e-3),
            "what": "k_kstar_mean without variance: k* is formed tile by tile and never written (reads X*, X, alpha; writes the mean)",
            "alu": {"achieved": mm * n * (3.0 * d + 42) / (ms_pred_mean * 1e-3) * 1e-12, "peak": alu_peak,
                    "unit": "TFLOP/s (approx. op count)"}}
        for key in ("assemble_hbm", "grad_contract_hbm", "predict_mean_hbm"):
            r = rl[key]
            if r["peak"]:
                r["frac"] = r["achieved"] / r["peak"]
            r["alu"]["frac"] = r["alu"]["achieved"] / r["alu"]["peak"]
        if not args.no_cpu_baseline and world == 1:
            dt, _ = oracle_eval_seconds(x, y, thetas[0], A)
            out["cpu_baseline"] = {
                "value": 1.0 / dt, "unit": "evals/s", "cores": os.cpu_count(), "kind": "port",
                "sample": f"1 of the {B} LML+gradient evaluations of one step (n={n}, d={d}), reference-faithful restatement "
                          f"oracle (NumPy + OpenBLAS potrf/potrs/potri, (n,n,d+1) tensor materialised): {dt:.1f} s",
            }
            try:
                # the reference builds OpenBLAS single-threaded (Makefile:3-4, USE_THREAD=0): same evaluation on 1 thread
                from threadpoolctl import threadpool_limits
                with threadpool_limits(limits=1):
                    dt1, _ = oracle_eval_seconds(x, y, thetas[0], A)
                out["cpu_baseline"]["reference_equivalent_1_thread"] = {"value": 1.0 / dt1, "unit": "evals/s", "cores": 1,
                                                                         "seconds_per_eval": dt1}
            except ImportError:
                pass
        print(json.dumps(out), flush=True)
    model.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {"hbm_gbs": 6650.0, "hbm_source": "fallback (B200_PROFILING.md)"}


if __name__ == "__main__":
    main()
