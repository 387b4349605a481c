"""CPU, world_size 2 and 3 over gloo: band partition + halo exchange."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, R, n1, n2, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import maprast_b200 as mr
        D = mr.distributed
        rng = np.random.default_rng(123)
        full = torch.from_numpy(rng.integers(0, 256, size=(n2, n1, 3), dtype=np.uint8))
        a, b = D.band_range(n2, world, rank)
        buf, own = D.alloc_band(n1, b - a, R, device="cpu")
        buf.fill_(0xEE)
        own.copy_(full[a:b])
        hl, hh = D.exchange_halos(buf, R)
        assert hl == (R if rank > 0 else 0) and hh == (R if rank < world - 1 else 0)
        # after the exchange the readable region equals the corresponding slice of the sheet
        got = buf[R - hl:R + (b - a) + hh]
        ok = torch.equal(got, full[a - hl:b + hh])
        q.put((rank, a, b, hl, hh, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,R", [(2, 2), (3, 3), (2, 0)])
def test_halo_exchange(world, R):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world * 7 + R
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, 37, 50, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(r[5] for r in res), res
    # bands tile the sheet exactly
    assert res[0][1] == 0 and res[-1][2] == 50
    for x, y in zip(res, res[1:]):
        assert x[2] == y[1]


def test_partitions(mr):
    D = mr.distributed
    for n, w in [(40000, 8), (2000, 3), (7, 8), (256, 8)]:
        r = [D.band_range(n, w, k) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(x[1] == y[0] for x, y in zip(r, r[1:]))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert [D.sheet_range(256, 8, k) for k in (0, 7)] == [(0, 32), (224, 256)]


def test_banded_oracle_equals_whole(oracle, mr):
    """Semantics of halo_lo/halo_hi: per-band clean with exchanged halos == whole-sheet clean."""
    from maprast_b200 import synth
    seed, pal, stats, _ = synth.config_workload(1)
    scan = oracle.synth_scan(seed, 0, pal, 96, 64)
    lab = oracle.categorize_u8(scan, stats.mean, stats.lo, stats.hi)
    whole = oracle.majority(lab, 2)
    parts = []
    for k in range(4):
        a, b = mr.distributed.band_range(64, 4, k)
        hl, hh = min(2, a), min(2, 64 - b)
        parts.append(oracle.majority(lab[a - hl:b + hh], 2, halo_lo=hl, halo_hi=hh))
    assert np.array_equal(np.concatenate(parts), whole)


class _FakeCtx:
    """records what set_known_for_band uploads (the host-side selection is what is tested here)"""

    def set_known_band(self, idx, cat, first_line):
        self.idx, self.cat, self.first_line = np.asarray(idx), np.asarray(cat), first_line


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("R", [0, 2, 3])
def test_known_points_per_band(world, R):
    """every rank gets exactly the whole-sheet points its windows can see (own lines +- R) and its first line"""
    sys.path.insert(0, ROOT)
    import maprast_b200 as mr
    D = mr.distributed
    n1, n2 = 37, 203
    rng = np.random.default_rng(world * 10 + R)
    idx = np.unique(rng.integers(0, n1 * n2, size=400))
    cat = rng.integers(1, 9, size=idx.size).astype(np.uint8)
    seen = np.zeros(idx.size, dtype=np.int64)
    for rank in range(world):
        a, b = D.band_range(n2, world, rank)
        fake = _FakeCtx()
        n = D.set_known_for_band(fake, idx, cat, n1, n2, R, rank, world)
        j = idx // n1
        want = (j >= a - R) & (j < b + R)
        assert n == int(want.sum()) and fake.first_line == a
        assert np.array_equal(fake.idx, idx[want]) and np.array_equal(fake.cat, cat[want])
        seen += ((j >= a) & (j < b)).astype(np.int64)
    assert (seen == 1).all()          # every point belongs to exactly one band's own lines
