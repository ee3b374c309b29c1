"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: restart sharding + deterministic winner
pick, candidate-row sharding + all-gather.  The GPU evaluation is replaced by a deterministic stand-in;
what is under test is the partition / exchange / merge code in hbetune_rs_b200/dist.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_run(starts):
    """Stand-in for Context.fit_runs: a deterministic function of each start point."""
    from hbetune_rs_b200 import _lib
    res = (_lib.RunResult * len(starts))()
    thetas = np.array(starts) * 2.0
    for i, s in enumerate(starts):
        res[i].best_lml = -float(np.sum((s - 0.3) ** 2))
        res[i].best_eval = int(abs(s[0]) * 10) % 7
        res[i].n_evals = 20 + i
        res[i].final_f = -res[i].best_lml
        res[i].status = 1 if s[0] > 0.9 else 0
    return res, thetas


def _worker(rank, size, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from hbetune_rs_b200 import _lib, dist as hd
        rng = np.random.default_rng(5)
        starts = rng.random((7, 4))
        starts[3] = starts[1]  # tie: the earlier run must win
        res, thetas = hd.sharded_fit_runs(starts, _fake_run)
        ref, ref_thetas = _fake_run(starts)
        for i in range(7):
            assert res[i].best_lml == ref[i].best_lml and res[i].best_eval == ref[i].best_eval
            assert res[i].status == ref[i].status and res[i].final_f == ref[i].final_f
        assert res[0].n_evals == 20  # run 0 is rank 0's first local run
        np.testing.assert_array_equal(thetas, ref_thetas)
        best = _lib.lib.hbegp_pick_best_run(7, res)
        ok = [i for i in range(7) if ref[i].status == 0]
        want = max(ok, key=lambda i: (ref[i].best_lml, -i))
        assert best == want
        xs = rng.random((11, 3))
        mean, var = hd.sharded_predict(lambda x, wv: (x.sum(axis=1), x.prod(axis=1)), xs)
        np.testing.assert_allclose(mean, xs.sum(axis=1))
        np.testing.assert_allclose(var, xs.prod(axis=1))
        mean, var = hd.sharded_predict(lambda x, wv: (x.sum(axis=1), None), xs[:1], want_variance=False)
        assert var is None and mean.shape == (1,)
        # balanced restart loop (hbegp_fit_runs_sharded's machinery over a host objective): the per-round sum
        # all-reduce goes through gloo; every rank must get the single-process records bit for bit
        from tests.test_host import _quadratic_objective, _run_with
        centers = np.array([0.3, -0.2, 1.0])
        blo, bhi = np.exp(np.full(3, -3.0)), np.exp(np.full(3, 3.0))
        st = np.random.default_rng(11).uniform(-2.5, 2.5, (7, 3))
        obj1, calls1 = _quadratic_objective(centers, fail_above=2.2)
        rc, ref2, ref2_thetas = _run_with(obj1, st, blo, bhi)
        assert rc == 0
        obj2, calls2 = _quadratic_objective(centers, fail_above=2.2)
        rc, res2, thetas2 = _run_with(obj2, st, blo, bhi, rank=rank, world=size, allreduce=hd.all_reduce_sum_inplace)
        assert rc == 0
        for i in range(7):
            assert (res2[i].best_eval, res2[i].n_evals, res2[i].status) == (ref2[i].best_eval, ref2[i].n_evals, ref2[i].status)
            assert res2[i].best_lml == ref2[i].best_lml or res2[i].status != 0
            assert res2[i].final_f == ref2[i].final_f
        np.testing.assert_array_equal(thetas2, ref2_thetas)
        assert len(calls2) <= len(calls1) and sum(calls2) < sum(calls1)  # only this rank's share was evaluated here
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_partitions_cover_everything():
    from hbetune_rs_b200 import dist as hd
    for size in (1, 2, 3, 8):
        runs = sorted(sum((hd.owned_runs(65, r, size) for r in range(size)), []))
        assert runs == list(range(65))
        blocks = [hd.row_block(1000003, r, size) for r in range(size)]
        assert blocks[0][0] == 0 and blocks[-1][1] == 1000003
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(size - 1))


def test_pick_best_run_ties_and_failures():
    from hbetune_rs_b200 import _lib
    res = (_lib.RunResult * 4)()
    for i, (l, st) in enumerate([(1.0, 1), (2.0, 0), (2.0, 0), (-1.0, 0)]):
        res[i].best_lml, res[i].status, res[i].best_eval = l, st, 0
    assert _lib.lib.hbegp_pick_best_run(4, res) == 1
    for i in range(4):
        res[i].status = 1
    assert _lib.lib.hbegp_pick_best_run(4, res) == -1
