"""N > 1 host logic on CPU: two gloo ranks shard the realization axis, run the ORACLE loop body on
their shards and merge counters; the result must equal the single-process run."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from chest_b200.distributed import shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 5, 16, 25, 1000):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def _fake_run(first, n):
    """deterministic stand-in for the loop body: counts depend on the global realization index only"""
    idx = np.arange(first, first + n, dtype=np.uint32)
    base = (idx[:, None, None, None, None, None] * 7 + 3) % 11
    return (base + np.arange(2 * 3 * 3 * 2 * 2, dtype=np.uint32).reshape(1, 2, 3, 3, 2, 2)).astype(np.uint32)


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    from chest_b200.distributed import run_sharded
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counters, err_all = run_sharded(_fake_run, n_total)
    q.put((rank, counters, err_all))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    n_total = 5                                   # ragged: 3 + 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _fake_run(0, n_total)
    for rank, counters, err_all in res:
        assert np.array_equal(counters, ref.astype(np.int64).sum(axis=0))
        assert np.array_equal(err_all, ref)
