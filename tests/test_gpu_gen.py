"""GPU tests of the synthetic generator and of the BASELINE.json query shapes on generated data.

1. The device generator must emit byte-for-byte what the reference WRITER path (restated in the oracle: valuesEncoder ->
   marshalStringsBlock -> tokenizeHashes -> bloom) produces for the same rows: encodings, dict order, min/max, lens items,
   data bytes, bloom bytes.
2. The C1..C4 filters of BASELINE.json on generated blocks: bit-exact bitmaps vs the oracle, through both the resident
   path and the end-to-end host-buffer path, and identical block-granular accounting (values / bloom / bitmap bytes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SEED = 20250718


@pytest.fixture(scope="module")
def env(oracle):
    from victorialogs_b200 import scan as vs
    ctx = vs.Ctx(0)
    yield oracle, vs, ctx
    ctx.close()


def cfgs(oracle, vs, **kw):
    a = dict(seed=SEED, total_rows=24 * 512 + 100, rows_per_block=512, hot_block_permille=400, hit_row_permille=150, columns_mask=15)
    a.update(kw)
    return oracle.GenConfig(**a), vs.GenConfig(**a)


def test_generator_matches_reference_writer(env):
    oracle, vs, ctx = env
    for kw in (dict(), dict(rows_per_block=64, total_rows=64 * 9), dict(rows_per_block=3000, total_rows=3000 * 3 + 64, hot_block_permille=1000, hit_row_permille=900),
               dict(columns_mask=15 | (4 << 8), hit_row_permille=500), dict(columns_mask=1 | (12 << 8), hot_block_permille=1000, hit_row_permille=1000, rows_per_block=200, total_rows=900)):   # vocabulary focus
        ocfg, gcfg = cfgs(oracle, vs, **kw)
        nb = (ocfg.total_rows + ocfg.rows_per_block - 1) // ocfg.rows_per_block
        batch = ctx.generate(gcfg, 0, nb)
        assert batch.rows == ocfg.total_rows and batch.nblocks == nb
        dl = ctx.download(batch)
        assert dl.field_names == [n for k, n in enumerate([b"_msg", b"level", b"path", b"status"]) if ocfg.columns_mask >> k & 1]
        for b in range(nb):
            ob = oracle.Block.generated(ocfg, b)
            assert dl.rows[b] == ob.rows and not ob.consts
            for oc in ob.columns:
                gc = dl.column(b, oc.name)
                lens, data = oracle.decode_values_block(oc.values_block)
                assert gc["value_type"] == oc.value_type, (b, oc.name)
                assert (gc["min_value"], gc["max_value"]) == (oc.min_value, oc.max_value), (b, oc.name)
                assert gc["dict"] == oc.dict, (b, oc.name)
                assert gc["lens_items"] == lens, (b, oc.name)
                assert gc["data"] == data, (b, oc.name)
                assert gc["bloom"] == oc.bloom, (b, oc.name, len(gc["bloom"]), len(oc.bloom))
        batch.free()


def test_vocabulary_focus_sets_the_selectivity(env):
    """bits 8..11 of columns_mask: every vocabulary row carries ONE entry, so hit_row_permille is the selectivity of that entry's query"""
    oracle, vs, ctx = env
    ocfg, gcfg = cfgs(oracle, vs, columns_mask=1 | (4 << 8), hot_block_permille=1000, hit_row_permille=500, rows_per_block=2000, total_rows=40000)
    batch = ctx.generate(gcfg, 0, 20)
    st = ctx.scan_resident(vs.Program(vs.Filter.regexp("_msg", "conn.*refused")), batch)
    want = sum(len(oracle.bitmap_rows(oracle.Block.generated(ocfg, b).search(oracle.Filter.regexp("_msg", "conn.*refused")), 2000)) for b in range(20))
    assert st.rows_matched == want and 0.45 < want / 40000 < 0.55
    batch.free()
    for bad in (13 << 8, 1 << 12):
        with pytest.raises(vs.VlscanError):
            ctx.generate(vs.GenConfig(seed=SEED, total_rows=100, rows_per_block=100, hot_block_permille=0, hit_row_permille=0, columns_mask=1 | bad), 0, 1)


def test_generator_block_ranges_are_consistent(env):
    oracle, vs, ctx = env
    ocfg, gcfg = cfgs(oracle, vs)
    whole = ctx.download(ctx.generate(gcfg, 0, 25))
    part = ctx.download(ctx.generate(gcfg, 7, 12))
    for j in range(5):
        for f in (b"_msg", b"level", b"path", b"status"):
            assert part.column(j, f) == whole.column(7 + j, f)


def queries(F):
    return {
        "C1 _msg:error": F.phrase("_msg", "error"),
        "C2 _msg:timeout AND level:error": F.and_([F.phrase("_msg", "timeout"), F.phrase("level", "error")]),
        "C3 _msg:~conn.*refused": F.regexp("_msg", "conn.*refused"),
        "C4 _msg:GET AND path:api* AND status:in(500,502,503)": F.and_([F.phrase("_msg", "GET"), F.prefix("path", "api"), F.in_("status", ["500", "502", "503"])]),
        "or/not": F.or_([F.phrase("_msg", "terror"), F.and_([F.exact("level", "FATAL"), F.not_(F.prefix("path", "static"))])]),
        "regex variant with tokens": F.regexp("_msg", "error: conn refused by .*"),
        "prefix on _msg": F.prefix("_msg", "time"),
        "phrase with punctuation": F.phrase("_msg", "GET /api/v1"),
        "in on strings": F.in_("path", ["health", "api/v1/items/7"]),
        "status exact": F.exact("status", "404"),
        "status prefix": F.prefix("status", "50"),
        "level regexp": F.regexp("level", "^(?i)err"),
        "no hits anywhere": F.phrase("_msg", "nosuchtoken"),
    }


def test_baseline_queries_on_generated_blocks(env):
    oracle, vs, ctx = env
    ocfg, gcfg = cfgs(oracle, vs)
    nb = 25
    oblocks = [oracle.Block.generated(ocfg, b) for b in range(nb)]
    batch = ctx.generate(gcfg, 0, nb)
    host = ctx.download(batch)
    oq, gq = queries(oracle.Filter), queries(vs.Filter)
    for name in oq:
        ostats = np.zeros(6, dtype=np.uint64)
        want = [oracle.bitmap_rows(b.search(oq[name], ostats), b.rows) for b in oblocks]
        prog = vs.Program(gq[name])
        st = ctx.scan_resident(prog, batch)
        words, counts = ctx.fetch()
        per = vs.split_bitmaps(words, [b.rows for b in oblocks])
        got = [oracle.bitmap_rows(np.ascontiguousarray(w), b.rows) for w, b in zip(per, oblocks)]
        assert got == want, name
        assert [int(c) for c in counts] == [len(w) for w in want], name
        # accounting parity with the oracle's block-granular, short-circuit-following counters
        assert (st.blocks, st.rows) == (int(ostats[0]), int(ostats[1])), name
        assert st.bloom_probe_bytes == int(ostats[2]), (name, st.bloom_probe_bytes, int(ostats[2]))
        assert st.values_bytes == int(ostats[3]), (name, st.values_bytes, int(ostats[3]))
        assert st.bitmap_bytes == int(ostats[4]), name
        assert st.columns_read == int(ostats[5]), name
        # the end-to-end path on host buffers gives the same bits
        w2, c2, st2 = ctx.scan_batch(prog, host)
        assert np.array_equal(w2, words), name
        assert st2.h2d_bytes > 0 and st2.d2h_bytes > 0
    batch.free()


def test_bloom_first_staging_leaves_pruned_values_on_the_host(env):
    """Clustered data (3 of 10 blocks hold vocabulary rows): staged bloom-first, the values of blocks the bloom filters rule out never cross
    PCIe - fewer bytes, same bits, same accounting - for on-disk and decoded stage; a query without tokens is staged in one go."""
    import os
    import parity_util as pu
    oracle, vs, ctx = env
    _, gcfg = cfgs(oracle, vs, total_rows=400_000, rows_per_block=2000, hot_block_permille=300, hit_row_permille=100)
    nb = 200
    batch = ctx.generate(gcfg, 0, nb)
    decoded = ctx.download(batch)
    batch.free()
    ondisk = decoded.compress(threads=8)
    F = vs.Filter
    trees = {"phrase": F.phrase("_msg", "timeout"), "and": F.and_([F.phrase("_msg", "timeout"), F.phrase("level", "error")]),
             "or": F.or_([F.phrase("_msg", "timeout"), F.phrase("_msg", "terror")]), "not": F.not_(F.phrase("_msg", "timeout")),
             "and-not": F.and_([F.phrase("_msg", "GET"), F.not_(F.prefix("path", "static"))]),
             "or of ands": F.or_([F.and_([F.phrase("_msg", "timeout"), F.exact("level", "error")]), F.and_([F.phrase("_msg", "refused"), F.in_("status", ["500", "503"])])])}
    old = os.environ.get("VLSCAN_BLOOM_FIRST")
    try:
        for name, tree in trees.items():
            prog = vs.Program(tree)
            for host in (ondisk, decoded):
                res = {}
                for mode in ("0", "2"):
                    os.environ["VLSCAN_BLOOM_FIRST"] = mode
                    w, c, st = ctx.scan_batch(prog, host)
                    res[mode] = (w.copy(), c.copy(), st)
                (w0, c0, s0), (w2, c2, s2) = res["0"], res["2"]
                assert np.array_equal(w0, w2) and np.array_equal(c0, c2), name
                for k in pu.ACCOUNTING:
                    assert getattr(s0, k) == getattr(s2, k), (name, k)
                assert s2.staged_columns + s2.pruned_columns > 0, name
                if name in ("phrase", "and", "or", "not"):   # cold blocks fail the bloom probe (under NOT: every row matches unread): most _msg values stay on the host
                    assert s2.pruned_columns > s2.staged_columns and s2.h2d_bytes < 0.7 * s0.h2d_bytes, (name, s2.pruned_columns, s2.staged_columns, s2.h2d_bytes, s0.h2d_bytes)
        # no tokens, no probe: the default (adaptive) mode stages in one go
        os.environ.pop("VLSCAN_BLOOM_FIRST", None)
        w, c, st = ctx.scan_batch(vs.Program(F.regexp("_msg", "conn.*refused")), ondisk)
        assert st.staged_columns == 0 and st.pruned_columns == 0
        # adaptive: a probe that prunes nothing switches the next calls of the same program to one-go staging
        prog = vs.Program(F.phrase("_msg", "stream"))   # a token of every row
        seen = [ctx.scan_batch(prog, ondisk)[2].staged_columns for _ in range(3)]
        assert seen[0] > 0 and seen[1] == 0 and seen[2] == 0, seen
    finally:
        if old is None:
            os.environ.pop("VLSCAN_BLOOM_FIRST", None)
        else:
            os.environ["VLSCAN_BLOOM_FIRST"] = old


def test_full_size_properties(env):
    """Size-independent properties at a larger scale (2M rows): NOT(f) complements f, AND is an intersection, counts add up,
    hit offsets are sorted and agree with the bitmaps, repeated scans are idempotent."""
    oracle, vs, ctx = env
    _, gcfg = cfgs(oracle, vs, total_rows=2_000_000, rows_per_block=4000, hot_block_permille=300, hit_row_permille=100)
    nb = 500
    batch = ctx.generate(gcfg, 0, nb)
    F = vs.Filter
    def run(f):
        st = ctx.scan_resident(vs.Program(f), batch)
        w, c = ctx.fetch()
        return w.copy(), c.copy(), st
    a, ca, sa = run(F.phrase("_msg", "error"))
    na, cna, _ = run(F.not_(F.phrase("_msg", "error")))
    allw, call, _ = run(F.noop())
    assert int(call.sum()) == 2_000_000
    assert np.array_equal(a | na, allw) and not np.any(a & na)
    b, cb, _ = run(F.phrase("level", "error"))
    ab, cab, _ = run(F.and_([F.phrase("_msg", "error"), F.phrase("level", "error")]))
    assert np.array_equal(ab, a & b)
    ob, cob, _ = run(F.or_([F.phrase("_msg", "error"), F.phrase("level", "error")]))
    assert np.array_equal(ob, a | b)
    a2, ca2, sa2 = run(F.phrase("_msg", "error"))
    assert np.array_equal(a, a2) and sa.rows_matched == sa2.rows_matched == int(ca.sum())
    hits, offs = ctx.fetch_hits()
    assert len(hits) == int(ca.sum())
    bits = np.unpackbits(a.view(np.uint8), bitorder="little")
    assert int(bits.sum()) == len(hits)
    for blk in (0, 17, nb - 1):
        h = hits[int(offs[blk]):int(offs[blk + 1])]
        assert np.all(np.diff(h.astype(np.int64)) > 0)
    # bloom pruning happened: cold blocks never had their _msg values read
    assert sa.values_bytes < batch.device_bytes and sa.blocks_matched < nb
    batch.free()
