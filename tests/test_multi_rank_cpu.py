"""CPU (gloo, world_size 2) test of the N>1 host path: block sharding + the final reduce of the match counters.
Each rank scans its own shard of a generated data set with the oracle (no GPU here) and the all-reduced counters must equal
a single-process scan of the whole data set."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KW = dict(seed=20250718, total_rows=23 * 512 + 100, rows_per_block=512, hot_block_permille=500, hit_row_permille=200, columns_mask=3)


def _tree(F):
    return F.and_([F.phrase("_msg", "timeout"), F.phrase("level", "error")])


def _scan(lo, hi):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vloracle
    cfg = vloracle.GenConfig(**KW)
    r = vloracle.scan_generated(cfg, _tree(vloracle.Filter), lo, hi, 2, want_counts=True)
    blocks_matched = int((r["counts"] > 0).sum())
    return [int(r["stats"][1]), int(r["matches"]), blocks_matched, int(r["stats"][3])], r["counts"]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from victorialogs_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nb = (KW["total_rows"] + KW["rows_per_block"] - 1) // KW["rows_per_block"]
    lo, hi = shard.shard_range(nb, world, rank)
    local, counts = _scan(lo, hi)
    t = torch.tensor(local, dtype=torch.int64)
    shard.reduce_counters(t)
    before, total = shard.gather_hit_prefix(local[1])
    out.put((rank, lo, hi, local, t.tolist(), before, total))
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    sys.path.insert(0, ROOT)
    from victorialogs_b200 import shard
    for nb in (0, 1, 7, 24, 33334):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard.shard_range(nb, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nb
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_reduce_to_the_single_process_answer():
    nb = (KW["total_rows"] + KW["rows_per_block"] - 1) // KW["rows_per_block"]
    whole, _ = _scan(0, nb)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, loc0, red0, before0, tot0), (r1, lo1, hi1, loc1, red1, before1, tot1) = res
    assert (lo0, hi1) == (0, nb) and hi0 == lo1
    assert red0 == red1 == whole
    assert [a + b for a, b in zip(loc0, loc1)] == whole
    assert (before0, before1) == (0, loc0[1]) and tot0 == tot1 == whole[1]


def test_shard_ranges_by_bytes():
    """Ranges are contiguous, cover every block once, and no rank carries more than its share plus one block."""
    import random
    from victorialogs_b200 import shard
    rng = random.Random(4)
    for trial in range(200):
        n, world = rng.choice([0, 1, 7, 64, 1000]), rng.choice([1, 2, 4, 8])
        sizes = [rng.choice([1, 10, 1000, 250000]) for _ in range(n)]
        ranges = shard.shard_ranges_by_bytes(sizes, world)
        assert len(ranges) == world and ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        if n:
            share = sum(sizes) / world
            for lo, hi in ranges:
                assert sum(sizes[lo:hi]) <= share + max(sizes) + 1e-9
