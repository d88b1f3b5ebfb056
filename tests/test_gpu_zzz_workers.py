"""Concurrent search workers (SURVEY §8(b) "threading"): the reference runs one goroutine per CPU, each with its own blockSearch
(lib/logstorage/storage_search.go:1040-1067); here that is one vlscan_ctx per host thread, all on the same device, sharing read-only
inputs and - like searchOptions.filter - one compiled program.  Every worker must get, on every batch, exactly the bitmaps a lone
worker gets.  (ctypes releases the GIL for the duration of a call, so the calls really overlap.)"""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_concurrent_workers_on_one_device():
    from victorialogs_b200 import scan as vs
    F = vs.Filter
    nworkers, rounds = 4, 6
    trees = [F.and_([F.phrase("_msg", "timeout"), F.phrase("level", "error")]), F.regexp("_msg", "conn.*refused"),
             F.or_([F.in_("status", ["500", "503"]), F.prefix("path", "api")])]
    shared = [vs.Program(t) for t in trees]                       # one program object used by all workers at once
    # inputs: every worker gets its own slice of a generated data set, in on-disk form (ZSTD frames decoded on the device) and decoded form
    setup = vs.Ctx(0)
    slices, want = [], []
    for w in range(nworkers):
        cfg = vs.GenConfig(seed=100 + w, total_rows=90 * 3000, rows_per_block=3000, hot_block_permille=600, hit_row_permille=60, columns_mask=0xF)
        host = vs.DownloadedBlocks(setup, setup.generate(cfg, 0, 90))
        disk = host.compress()
        slices.append((host, disk))
        want.append([[a.copy() for a in setup.scan_batch(p, host)[:2]] for p in shared])
    assert sum(int(c.sum()) for per in want for _, c in per) > 1000
    errors = []
    barrier = threading.Barrier(nworkers)

    def worker(w):
        try:
            ctx = vs.Ctx(0)
            own = [vs.Program(t) for t in trees]                  # and programs of its own
            host, disk = slices[w]
            barrier.wait(timeout=120)
            for r in range(rounds):
                for k in range(len(trees)):
                    prog = shared[k] if (r + w) % 2 == 0 else own[k]
                    blocks = disk if (r + k) % 2 == 0 else host
                    words, counts, st = ctx.scan_batch(prog, blocks)
                    if not (np.array_equal(words, want[w][k][0]) and np.array_equal(counts, want[w][k][1])):
                        errors.append("worker %d round %d tree %d: result differs from the single-worker scan" % (w, r, k))
                # a malformed call on this worker must not disturb the others, and its error text stays on this thread
                try:
                    vs.Program(vs.Filter(b"\xff", "bad"))
                    errors.append("worker %d: malformed tree accepted" % w)
                except vs.VlscanError as e:
                    if "filter" not in str(e):
                        errors.append("worker %d: unexpected error text %r" % (w, str(e)))
            ctx.close()
        except Exception as e:                                     # noqa: BLE001 - reported below
            errors.append("worker %d: %r" % (w, e))

    threads = [threading.Thread(target=worker, args=(w,)) for w in range(nworkers)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in threads), "a worker hangs"
    assert not errors, errors[:5]
    setup.close()
