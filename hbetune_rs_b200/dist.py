"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink; gloo in the CPU
tests).  Only the two parts of the path that shard naturally are sharded:

* restart runs of the fit (``src/util/gradmin.rs:19-30``).  ``BalancedFit``: all ranks drive the same lockstep
  loop, each round's live runs are dealt out round-robin and one small sum all-reduce shares the round's
  (lml, status, gradient) records (``hbegp_fit_runs_sharded``) -- the GPUs stay evenly loaded while runs finish at
  different times.  ``sharded_fit_runs`` is the static alternative: run r goes to rank r mod G, one all-gather of
  the per-run ``(best_lml, best_eval, n_evals, final_f, status, theta[p])`` records at the end.  Either way every
  rank applies the same deterministic pick (``fit.rs:116-117``: strict ``>``, earliest wins) to identical records;
* candidate rows of a prediction: contiguous blocks, all-gather of the (mean, var) shards.

The single n x n factorisation stays on one GPU (replicas only).
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np

from . import _lib


def _dist():
    import torch.distributed as dist
    return dist


def world() -> Tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _device():
    import torch
    dist = _dist()
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def owned_runs(n_runs: int, rank: int, size: int) -> List[int]:
    return list(range(rank, n_runs, size))


def all_gather_array(local: np.ndarray, rows_per_rank: Sequence[int]) -> List[np.ndarray]:
    """All-gathers float64 arrays whose leading dimension differs per rank (padded to the maximum)."""
    import torch
    rank, size = world()
    if size == 1:
        return [local]
    dist = _dist()
    width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
    maxrows = max(rows_per_rank)
    buf = torch.zeros((maxrows, width), dtype=torch.float64, device=_device())
    if local.shape[0]:
        buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64).reshape(local.shape[0], width)).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(size)]
    dist.all_gather(out, buf)
    return [o[: rows_per_rank[r]].cpu().numpy().reshape((rows_per_rank[r],) + local.shape[1:]) for r, o in enumerate(out)]


def sharded_fit_runs(starts: np.ndarray, run_fn: Callable):
    """``run_fn(starts_subset) -> (results, thetas)`` runs the local subset; returns the records of ALL runs
    in run order on every rank (as a ctypes array of ``hbegp_run_result`` plus the theta matrix)."""
    rank, size = world()
    n_runs, p = starts.shape
    mine = owned_runs(n_runs, rank, size)
    if mine:
        res, thetas = run_fn(starts[mine])
        rec = np.array([[r.best_lml, float(r.best_eval), float(r.n_evals), r.final_f, float(r.status)] for r in res])
        local = np.concatenate([rec, np.asarray(thetas, dtype=np.float64)], axis=1)
    else:
        local = np.zeros((0, 5 + p))
    counts = [len(owned_runs(n_runs, r, size)) for r in range(size)]
    parts = all_gather_array(local, counts)
    merged = np.zeros((n_runs, 5 + p))
    for r, part in enumerate(parts):
        for row, run in zip(part, owned_runs(n_runs, r, size)):
            merged[run] = row
    out = (_lib.RunResult * n_runs)()
    for i in range(n_runs):
        out[i].best_lml = merged[i, 0]
        out[i].best_eval = int(merged[i, 1])
        out[i].n_evals = int(merged[i, 2])
        out[i].final_f = merged[i, 3]
        out[i].status = int(merged[i, 4])
    return out, merged[:, 5:].copy()


def all_reduce_sum_inplace(values: np.ndarray) -> None:
    """Sums a float64 array over all ranks in place (NCCL through a device tensor, gloo directly)."""
    import torch
    rank, size = world()
    if size == 1:
        return
    dist = _dist()
    dev = _device()
    t = torch.from_numpy(values)
    if dev.type == "cuda":
        g = t.to(dev)
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        t.copy_(g)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)


def init_library_comm(ctx) -> None:
    """Gives ``ctx`` (this rank's context) the NCCL communicator the library uses for its own exchanges
    (``hbegp_comm_init``): rank 0 creates the id, ``torch.distributed`` only carries those 128 bytes."""
    import torch
    rank, size = world()
    if size == 1:
        return
    dist = _dist()
    dev = _device()
    ident = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        ident = torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(ident, src=0)
    ctx.comm_init(size, rank, bytes(ident.cpu().numpy().tobytes()))


class LibraryFit:
    """``shard=`` argument of ``FittedKernel.new``: the balanced restart loop with the per-round exchange done INSIDE
    libhbegp.so (device-side packing + ncclAllReduce on the communicator of ``init_library_comm``)."""

    def fit_runs(self, ctx, starts, lo, hi, nu=2.5, maxeval=150):
        rank, size = world()
        return ctx.fit_runs(starts, lo, hi, nu, maxeval, rank=rank, world=size, allreduce=None)


class BalancedFit:
    """``shard=`` argument of ``FittedKernel.new`` / ``EstimatorGPR.shard``: the restart loop balanced per round over
    all ranks (``hbegp_fit_runs_sharded``).  Every rank must hold the same training data and call with the same
    arguments; every rank gets the records of all runs."""

    def fit_runs(self, ctx, starts, lo, hi, nu=2.5, maxeval=150):
        rank, size = world()
        return ctx.fit_runs(starts, lo, hi, nu, maxeval, rank=rank, world=size, allreduce=all_reduce_sum_inplace)


def row_block(m: int, rank: int, size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of candidate rows owned by ``rank``."""
    base, extra = divmod(m, size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_predict(predict_fn: Callable, xs: np.ndarray, want_variance: bool = True):
    """``predict_fn(xs_block, want_variance) -> (mean, var)`` on the local block; all-gathers the shards."""
    rank, size = world()
    m = xs.shape[0]
    lo, hi = row_block(m, rank, size)
    mean, var = predict_fn(xs[lo:hi], want_variance)
    cols = [np.asarray(mean, dtype=np.float64)]
    if want_variance:
        cols.append(np.asarray(var, dtype=np.float64))
    local = np.stack(cols, axis=1) if hi > lo else np.zeros((0, len(cols)))
    counts = [row_block(m, r, size)[1] - row_block(m, r, size)[0] for r in range(size)]
    full = np.concatenate(all_gather_array(local, counts), axis=0)
    return full[:, 0].astype(xs.dtype), (full[:, 1].astype(xs.dtype) if want_variance else None)
