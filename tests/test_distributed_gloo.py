"""CPU: the multi-rank host logic of abeille_b200.distributed on a world_size-2 gloo group -- partition plans,
the order-preserving bank rebalance (all_to_all), the scalar gather and the global history-id scan.  The GPU
kernels are not involved (no device here); the same functions run over NCCL on the B200s."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from abeille_b200.distributed import (even_split, gather_vector, global_first_ids, pair_distance_sqrd, rebalance_bank,
                                      rebalance_plan)


def test_even_split_and_plan_cover_everything():
    for world in (1, 2, 3, 8):
        for counts in ([5] * world, list(range(1, world + 1)), [0] * (world - 1) + [1000], [17, 0, 3, 99, 1, 0, 0, 5][:world]):
            total = sum(counts)
            target, bounds = even_split(total, world)
            assert sum(target) == total and max(target) - min(target) <= 1 and bounds[-1] == total
            sends = [rebalance_plan(counts, r)[0] for r in range(world)]
            recvs = [rebalance_plan(counts, r)[1] for r in range(world)]
            for r in range(world):
                assert sum(sends[r]) == counts[r] and sum(recvs[r]) == target[r]
                for q in range(world):
                    assert sends[r][q] == recvs[q][r]


def test_plan_to_the_master_and_back():
    """particles_to_master / distribute_particles of the sharded comb: any target partition, order-preserving."""
    for world in (1, 2, 3, 8):
        for counts in ([5] * world, list(range(1, world + 1)), [17, 0, 3, 99, 1, 0, 0, 5][:world]):
            total = sum(counts)
            for target in ([total] + [0] * (world - 1), even_split(total, world)[0], counts[::-1]):
                sends = [rebalance_plan(counts, r, target)[0] for r in range(world)]
                recvs = [rebalance_plan(counts, r, target)[1] for r in range(world)]
                for r in range(world):
                    assert sum(sends[r]) == counts[r] and sum(recvs[r]) == target[r]
                    for q in range(world):
                        assert sends[r][q] == recvs[q][r]
                # global order: what rank r receives from rank q precedes what it receives from rank q + 1, and the pieces a rank
                # sends go to ranks in increasing order of their target ranges
                old = np.concatenate([[0], np.cumsum(counts)])
                new = np.concatenate([[0], np.cumsum(target)])
                for r in range(world):
                    pos = old[r]
                    for q in range(world):
                        if sends[r][q]:
                            assert new[q] <= pos and pos + sends[r][q] <= new[q + 1]
                            pos += sends[r][q]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, counts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        old = np.concatenate([[0], np.cumsum(counts)])
        n = counts[rank]
        cap = sum(counts)
        # global order = the value itself
        src = {"x": torch.zeros(cap, dtype=torch.float64), "id_a": torch.zeros(cap, dtype=torch.int64)}
        src["x"][:n] = torch.arange(old[rank], old[rank + 1], dtype=torch.float64)
        src["id_a"][:n] = torch.arange(old[rank], old[rank + 1], dtype=torch.int64) * 7
        dst = {k: torch.full((cap,), -1, dtype=v.dtype) for k, v in src.items()}
        m = rebalance_bank(src, dst, ["x", "id_a"], counts, rank)
        allv = gather_vector(np.array([float(m), float(rank), 2.5]), world, torch.device("cpu"))
        new_counts = [int(v) for v in allv[:, 0]]
        first = global_first_ids(new_counts, rank, 1000)
        q.put((rank, m, dst["x"][:m].numpy().copy(), dst["id_a"][:m].numpy().copy(), allv, first))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[10, 30], [0, 41], [25, 25], [1000, 3]])
def test_rebalance_preserves_global_order_on_two_ranks(counts):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, counts, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = sum(counts)
    target, bounds = even_split(total, world)
    assert [r[1] for r in res] == target
    xs = np.concatenate([r[2] for r in res])
    assert np.array_equal(xs, np.arange(total, dtype=np.float64))          # global order preserved
    assert np.array_equal(np.concatenate([r[3] for r in res]), np.arange(total) * 7)
    for r in res:
        assert np.array_equal(r[4][:, 1], [0.0, 1.0]) and np.all(r[4][:, 2] == 2.5)
        assert r[5] == 1000 + int(bounds[r[0]])                             # contiguous global history ids


def _cloud(n=900, seed=3):
    rng = np.random.default_rng(seed)
    r = rng.normal(size=(n, 3)) * np.array([20.0, 20.0, 60.0]) + np.array([5.0, -3.0, 40.0])
    w = rng.uniform(0.2, 1.8, n)
    w[rng.random(n) < 0.05] *= -1.0   # a few negative weights, as carter tracking leaves them
    return r, w


def _moments(r, w, origin):  # what abl_bank_moments_device returns for a slice
    d = r - np.asarray(origin, dtype=np.float64)
    return np.array([w.sum(), *(w[:, None] * d).sum(axis=0), (w * (d * d).sum(axis=1)).sum()])


def _pair_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, w = _cloud()
        lo, hi = (0, 350) if rank == 0 else (350, len(w))    # uneven slices
        got = pair_distance_sqrd(lambda o: _moments(r[lo:hi], w[lo:hi], o), lambda v: gather_vector(v, world, torch.device("cpu")))
        q.put((rank, got))
    finally:
        dist.destroy_process_group()


def test_pair_distance_of_a_sharded_bank_is_the_references_double_sum():
    """settings: pair-distance-sqrd on two ranks: the sums of the slices' moments give the reference's double sum over all pairs of
    the WHOLE bank (PowerIterator::compute_pair_dist_sqrd, src/power_iterator.cpp:637-663, evaluated here term by term), and the
    one-rank value."""
    r, w = _cloud()
    d2 = ((r[:, None, :] - r[None, :, :]) ** 2).sum(axis=2)
    ref = float((d2 * w[:, None] * w[None, :]).sum() / (2.0 * w.sum() ** 2))
    one = pair_distance_sqrd(lambda o: _moments(r, w, o), lambda v: v[None, :])
    assert abs(one - ref) < 1e-12 * abs(ref)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pair_worker, args=(k, world, port, q)) for k in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] and abs(res[0][1] - ref) < 1e-12 * abs(ref)
