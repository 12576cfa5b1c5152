"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: partitions, the collectives of the
row-sharded mode and the stage schedule, run with a numpy stage backend on oracle arithmetic.
The GPU kernels themselves are covered by the -m gpu tests."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, synthetic_pair

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, port, fn_name, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        queue.put((rank, globals()[fn_name](rank)))
    finally:
        dist.destroy_process_group()


def _run(fn_name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, fn_name, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(WORLD))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return results


# ---------------------------------------------------------------------------- worker bodies
def _sharded_idt(rank):
    import color_transfer_b200  # noqa: F401
    from color_transfer_b200 import sharded
    from numpy_backend import NumpyIdtBackend
    from oracle import reference_numpy as oracle
    t, r = synthetic_pair(61, 47, 31, np.float64, ref_shape=(53, 40))
    rot = sharded.predraw_rotations(1, 4, seed=42)[0]
    want, traces = oracle.idt_instrumented(t, r, rotations=rot)
    a, b = sharded.row_partition(t.shape[0], WORLD, rank)
    ra, rb = sharded.row_partition(r.shape[0], WORLD, rank)
    comm = sharded.Comm()
    backend = NumpyIdtBackend(t[a:b], r[ra:rb], rot, 255, 4)
    out = sharded.idt_transfer_sharded(None, None, rot, 255, 4, comm=comm, backend=backend)
    counts_ok = all(np.array_equal(backend.trace[i]["counts_t"], traces[i]["counts_t"]) and
                    np.array_equal(backend.trace[i]["counts_r"], traces[i]["counts_r"]) for i in range(4))
    ranges_ok = all(np.array_equal(backend.trace[i]["lo"], traces[i]["lo"]) and
                    np.array_equal(backend.trace[i]["hi"], traces[i]["hi"]) for i in range(4))
    luts_ok = all(np.array_equal(backend.trace[i]["lut"], traces[i]["lut"]) for i in range(4))
    return {"err": float(np.max(np.abs(out - want[a:b]))), "counts_ok": counts_ok, "ranges_ok": ranges_ok,
            "luts_ok": luts_ok, "collectives": comm.calls, "rows": (a, b)}


def _sharded_moments(rank):
    import color_transfer_b200  # noqa: F401
    from color_transfer_b200 import sharded
    t, _ = synthetic_pair(37, 29, 32, np.float64)
    a, b = sharded.row_partition(t.shape[0], WORLD, rank)
    x = t[a:b].reshape(-1, 3) - 0.5                      # the kernels' fixed shift K = 0.5
    iu = np.triu_indices(3)
    sums = np.concatenate([[x.shape[0]], x.sum(0), (x[:, :, None] * x[:, None, :]).sum(0)[iu]])
    total = sharded.Comm().gather_sum_in_rank_order(torch.from_numpy(sums)).numpy()
    n, s1 = total[0], total[1:4]
    m2 = np.zeros((3, 3))
    m2[iu] = total[4:]
    m2 = m2 + np.triu(m2, 1).T
    mean = 0.5 + s1 / n
    cov = (m2 - np.outer(s1, s1) / n) / (n - 1)
    full = t.reshape(-1, 3)
    return {"mean_err": float(np.max(np.abs(mean - full.mean(0)))), "cov_err": float(np.max(np.abs(cov - np.cov(full.T)))),
            "total": total.tolist()}


def _frame_parallel(rank):
    import color_transfer_b200  # noqa: F401
    from color_transfer_b200 import sharded
    mine = sharded.frame_partition(7, WORLD, rank)
    rot = sharded.predraw_rotations(7, 4, seed=42)       # every rank draws the same stream
    gathered = [None] * WORLD
    dist.all_gather_object(gathered, (mine, rot[mine].tobytes()))
    return {"mine": mine, "all": [g[0] for g in gathered], "same_rot": rot.tobytes()}


# ---------------------------------------------------------------------------- tests
def test_row_sharded_idt_matches_unsharded_oracle():
    res = _run("_sharded_idt")
    assert [res[r]["rows"] for r in range(WORLD)] == [(0, 31), (31, 61)]
    for r in range(WORLD):
        assert res[r]["counts_ok"], "sharded histogram counts differ from the unsharded ones"
        assert res[r]["ranges_ok"] and res[r]["luts_ok"]
        assert res[r]["err"] < 1e-13
        assert res[r]["collectives"] == (1 + 3) + 4      # MIN of keys: iteration 0 + 3 nexts; SUM of counts: 4


def test_row_sharded_moments_are_additive_and_rank_identical():
    res = _run("_sharded_moments")
    assert res[0]["total"] == res[1]["total"]            # bit-identical on every rank
    assert res[0]["mean_err"] < 1e-14 and res[0]["cov_err"] < 1e-15


def test_frame_parallel_partition_and_rotation_order():
    res = _run("_frame_parallel")
    assert res[0]["mine"] == [0, 2, 4, 6] and res[1]["mine"] == [1, 3, 5]
    assert sorted(res[0]["all"][0] + res[0]["all"][1]) == list(range(7))
    assert res[0]["same_rot"] == res[1]["same_rot"]


def test_partitions_single_process():
    import color_transfer_b200  # noqa: F401
    from color_transfer_b200 import sharded
    for h, w in ((16384, 8), (10, 3), (7, 8), (1, 2)):
        blocks = [sharded.row_partition(h, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == h
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
    assert sharded.frame_partition(600, 8, 3)[:3] == [3, 11, 19] and len(sharded.frame_partition(600, 8, 3)) == 75
