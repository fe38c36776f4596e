"""N > 1 host logic on CPU: world_size-2 gloo process groups (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from baofit_b200 import shard


def test_shard_ranges_partition_in_order():
    for total in (0, 1, 7, 2100, 10000):
        for world in (1, 2, 3, 8):
            blocks = [shard.shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f1 == f0 + c0
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeSession:
    """A scan whose 'fit' at each grid point is the library's own minimiser over a cheap CPU
    objective, so that the sharding + gather logic is exercised without a GPU."""
    npar = 3

    def __init__(self):
        from baofit_b200 import host_api as H
        self.H = H
        self.grid = [(a, b) for a in np.linspace(0.8, 1.3, 7) for b in np.linspace(0.6, 1.5, 9)]

    def scan_size(self):
        return len(self.grid)

    def parameter_scan(self, first, count):
        chi2, pts, st = [], [], []
        for a, b in self.grid[first:first + count]:
            r = self.H.minimize(lambda p: (p[:, 0] - a * b) ** 2 + 0.3 * (a - 1) ** 2 + 0.1 * (b - 1) ** 2,
                                [0.0, a, b], [0.1, 0.1, 0.1], [1, 0, 0])
            chi2.append(2 * r["fval"]); pts.append(r["values"]); st.append(r["status"])
        return dict(chi2=np.array(chi2), points=np.array(pts).reshape(-1, 3), status=np.array(st, dtype=np.int32),
                    best_values=np.zeros(3), best_chi2=0.0)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = shard.scan_sharded(FakeSession())
        np.save(os.path.join(out_dir, "scan_%d.npy" % rank), np.column_stack([res["chi2"], res["points"]]))
        uneven = np.arange(shard.shard_range(11, rank, world)[1] * 2, dtype=np.float64).reshape(-1, 2) + 100 * rank
        np.save(os.path.join(out_dir, "uneven_%d.npy" % rank), shard.gather_blocks(uneven, 11))
    finally:
        dist.destroy_process_group()


def test_scan_sharded_over_two_gloo_ranks_equals_single_process(built, tmp_path):
    single = FakeSession().parameter_scan(0, 63)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "scan_0.npy"), np.load(tmp_path / "scan_1.npy")
    assert np.array_equal(r0, r1)
    assert np.array_equal(r0[:, 0], single["chi2"]) and np.array_equal(r0[:, 1:], single["points"])
    u0 = np.load(tmp_path / "uneven_0.npy")
    assert u0.shape == (11, 2) and np.array_equal(u0, np.load(tmp_path / "uneven_1.npy"))
    assert u0[0, 0] == 0 and u0[6, 0] == 100  # rank 0 owns 6 of 11 units, rank 1 the other 5


class FakeToySession:
    """Toy 'fits' keyed by the global toy index, as bf_sample_toys keys its noise: results must not
    depend on how the realisations are split over ranks."""
    npar = 2

    def toymc(self, first, count, seed=1966):
        idx = np.arange(first, first + count)
        rng_vals = np.array([np.random.Generator(np.random.PCG64([seed, int(i)])).standard_normal(3) for i in idx]).reshape(-1, 3)
        return dict(fitted=rng_vals[:, :2], chi2=100 + rng_vals[:, 2], status=(idx % 7 == 0).astype(np.int32))


def _toy_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = shard.toymc_sharded(FakeToySession(), 37, seed=5)
        np.save(os.path.join(out_dir, "toys_%d.npy" % rank), np.column_stack([res["chi2"], res["fitted"], res["status"]]))
    finally:
        dist.destroy_process_group()


def test_toymc_sharded_over_two_gloo_ranks_equals_single_process(tmp_path):
    single = FakeToySession().toymc(0, 37, seed=5)
    port = _free_port()
    mp.spawn(_toy_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "toys_0.npy"), np.load(tmp_path / "toys_1.npy")
    assert np.array_equal(r0, r1) and r0.shape == (37, 4)
    assert np.array_equal(r0[:, 0], single["chi2"]) and np.array_equal(r0[:, 1:3], single["fitted"])
    assert np.array_equal(r0[:, 3].astype(np.int32), single["status"])
