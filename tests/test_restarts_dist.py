"""The N > 1 host path (independent fits sharded over ranks, results gathered) on CPU: gloo, world_size 2."""
import os
import socket

import pytest
import torch.multiprocessing as mp

from additivecausalexpansion_b200 import restarts


def test_assign_fits_round_robin():
    a = restarts.assign_fits(64, 8)
    assert all(len(x) == 8 for x in a) and a[3][:3] == [3, 11, 19]
    assert sorted(i for part in restarts.assign_fits(7, 3) for i in part) == list(range(7))
    assert restarts.assign_fits(2, 4) == [[0], [1], [], []]
    with pytest.raises(ValueError):
        restarts.assign_fits(3, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = restarts.run_restarts(7, lambda i: {"evidence": -100.0 + (i * 37 % 11), "iters": 10 + i}, dist)
    q.put((rank, [(r["fit"], r["rank"], r["evidence"]) for r in res], restarts.best_restart(res)["fit"]))
    dist.destroy_process_group()


def test_run_restarts_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = None
    for rank, res, best in got:
        assert [f for f, _, _ in res] == list(range(7))
        assert all(rk == f % 2 for f, rk, _ in res)  # fit i ran on rank i % 2
        assert best == max(range(7), key=lambda i: -100.0 + (i * 37 % 11))
        ref = ref or res
        assert res == ref  # every rank sees the same gathered list


def test_run_restarts_single_process():
    res = restarts.run_restarts(3, lambda i: {"evidence": float(i)})
    assert [r["fit"] for r in res] == [0, 1, 2] and restarts.best_restart(res)["fit"] == 2


def test_run_restarts_concurrent_threads():
    """`concurrency` fits of a rank's share at a time (one host thread each): same results, same order."""
    import threading
    import time

    seen, lock, active, peak = [], threading.Lock(), [0], [0]

    def fit(i):
        with lock:
            active[0] += 1
            peak[0] = max(peak[0], active[0])
        time.sleep(0.02)
        with lock:
            active[0] -= 1
            seen.append(i)
        return {"evidence": float(-i)}

    res = restarts.run_restarts(6, fit, concurrency=3)
    assert [r["fit"] for r in res] == list(range(6)) and [r["evidence"] for r in res] == [-float(i) for i in range(6)]
    assert peak[0] >= 2 and sorted(seen) == list(range(6))
