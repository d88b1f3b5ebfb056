"""A part directory on disk -> vlscan_part_open (metadata inflated by the DEVICE decoder, the default) -> descriptors -> vlscan_scan_batch,
against the oracle run on the blocks the part was written from.  Bar: bit-exact bitmaps and counts; the reader's descriptors are
also compared with those the libzstd-backed open of tests/test_part_reader_cpu.py produces."""
import numpy as np
import pytest

from test_part_reader_cpu import write_part, all_fields, check_block, libzstd_inflate

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(oracle):
    from victorialogs_b200 import scan as vs
    ctx = vs.Ctx(0)
    yield oracle, vs, ctx
    ctx.close()


def test_part_on_disk_scans_like_the_blocks_it_was_written_from(env, tmp_path):
    oracle, vs, ctx = env
    path, files, originals, header = write_part(tmp_path, "part", max_shards=4, max_index_block=400, seed=21, streams=6, blocks_per_stream=5)
    part = vs.Part(path, ctx=ctx)                                   # column_names.bin, metaindex.bin, index.bin blocks: inflated on the device
    cpu = vs.Part(path, inflate=libzstd_inflate)
    assert part.header == cpu.header == header and part.column_names == cpu.column_names
    assert [part.block_header(i) for i in range(part.nblocks)] == [cpu.block_header(i) for i in range(cpu.nblocks)]
    fields = all_fields(originals)
    hb = part.blocks(fields)
    for i, (sid, b, ts) in enumerate(originals):
        check_block(hb, i, fields, b)
    F, OF = vs.Filter, oracle.Filter
    pairs = [
        (F.phrase("_msg", "error"), OF.phrase("_msg", "error")),
        (F.and_([F.phrase("_msg", "GET"), F.or_([F.exact("level", "info"), F.prefix("sparse", "v1")])]), OF.and_([OF.phrase("_msg", "GET"), OF.or_([OF.exact("level", "info"), OF.prefix("sparse", "v1")])])),
        (F.in_("status", ["404", "500", "503"]), OF.in_("status", ["404", "500", "503"])),
        (F.regexp("_msg", "conn.*peer"), OF.regexp("_msg", "conn.*peer")),
        (F.not_(F.exact("host", "host-1")), OF.not_(OF.exact("host", "host-1"))),
        (F.ipv4_range("ip", 0x0A000000, 0x7FFFFFFF), OF.ipv4_range("ip", 0x0A000000, 0x7FFFFFFF)),
        (F.exact_prefix("ts", "2024-0"), OF.exact_prefix("ts", "2024-0")),
        (F.string_range("ratio", "2", "7"), OF.string_range("ratio", "2", "7")),
        (F.phrase("bytes", "12"), OF.phrase("bytes", "12")),
        (F.exact("only_in_1", "u7"), OF.exact("only_in_1", "u7")),
        (F.phrase("delta", "-5"), OF.phrase("delta", "-5")),
    ]
    matched = 0
    for gf, of in pairs:
        prog = vs.Program(gf)
        sub = part.blocks([f.decode() if f != b"_msg" else "_msg" for f in prog.fields()])     # only the fields the program reads
        words, counts, st = ctx.scan_batch(prog, sub)
        per = vs.split_bitmaps(words, sub.rows)
        for i, (sid, b, ts) in enumerate(originals):
            want = oracle.bitmap_rows(b.search(of), b.rows)
            assert oracle.bitmap_rows(np.ascontiguousarray(per[i]), b.rows) == want, (gf, i)
            assert int(counts[i]) == len(want)
            matched += len(want)
    assert matched > 200
    # time pruning at the block level, then the scan on what is left
    lo_t, hi_t = originals[7][2][0], originals[19][2][-1]
    sub = part.blocks(["_msg"], min_timestamp=lo_t, max_timestamp=hi_t)
    want_src = [i for i, (_, _, ts) in enumerate(originals) if not (ts[-1] < lo_t or ts[0] > hi_t)]
    assert sub.source == want_src and len(want_src) < len(originals)
    prog = vs.Program(F.phrase("_msg", "timeout"))
    words, counts, st = ctx.scan_batch(prog, sub)
    of = OF.phrase("_msg", "timeout")
    assert [int(c) for c in counts] == [len(oracle.bitmap_rows(originals[i][1].search(of), originals[i][1].rows)) for i in want_src]


def test_search_part_loop(env, tmp_path):
    """search_part = the block loop of Storage.search for one part: time pruning per block, batches through vlscan_scan_batch"""
    oracle, vs, ctx = env
    path, files, originals, header = write_part(tmp_path, "part2", seed=33, streams=5, blocks_per_stream=4)
    part = vs.Part(path, ctx=ctx)
    gf = vs.Filter.or_([vs.Filter.phrase("_msg", "reset"), vs.Filter.exact("level", "error")])
    of = oracle.Filter.or_([oracle.Filter.phrase("_msg", "reset"), oracle.Filter.exact("level", "error")])
    lo_t, hi_t = originals[3][2][1 if len(originals[3][2]) > 1 else 0], originals[14][2][-1]
    for batch_blocks in (3, 1000):
        hits = vs.search_part(ctx, part, gf, lo_t, hi_t, batch_blocks=batch_blocks)
        want = []
        for i, (sid, b, ts) in enumerate(originals):
            if ts[-1] < lo_t or ts[0] > hi_t:
                continue
            rows = oracle.bitmap_rows(b.search(of), b.rows)
            if rows:
                want.append((i, rows, lo_t <= ts[0] and ts[-1] <= hi_t))
        assert [(src, oracle.bitmap_rows(np.ascontiguousarray(w), originals[src][1].rows), inside) for src, w, c, inside in hits] == want
        assert [c for _, _, c, _ in hits] == [len(r) for _, r, _ in want]
        assert len(want) > 3
