"""GPU parity for the timestamps column: encoding.UnmarshalTimestamps on the device for all six marshal types
and filterTime (lib/logstorage/filter_time.go:114-137), through the C ABI against the CPU oracle (oracle/vlo_timestamps.h) and the reference's
own table (filter_time_test.go:13-88).  Bar: bit-exact bitmaps and counts.  Also vlscan_gather_timestamps / vlscan_gather_values: the
selected rows' `_time` and column values as blockResult would materialise them (block_result.go:491-507, 529-591)."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MT_NAMES = {1: "zstd nearest delta2", 2: "delta const", 3: "const", 4: "zstd nearest delta", 5: "nearest delta2", 6: "nearest delta"}


@pytest.fixture(scope="module")
def env(oracle):
    from victorialogs_b200 import scan as vs
    import parity_util as pu
    ctx = vs.Ctx(0)
    yield oracle, vs, pu, ctx
    ctx.close()


def check(env, blocks, of, gf, stage="ondisk"):
    oracle, vs, pu, ctx = env
    want = [oracle.bitmap_rows(b.search(of), b.rows) for b in blocks]
    got, counts, st = pu.gpu_rows(ctx, gf, blocks, stage)
    assert got == want, gf
    assert [int(c) for c in counts] == [len(w) for w in want]
    return want


def series(rng, kind, n):
    base = rng.choice([0, 1_700_000_000_000_000_000, -5_000_000_000, (1 << 62)])
    if kind == "const":
        return [base] * n
    if kind == "step":
        d = rng.choice([1, 1000, 123456789])
        return [base + i * d for i in range(n)]
    if kind == "jitter":
        t, out = base, []
        for _ in range(n):
            t += max(0, int(rng.gauss(1_000_000, 200_000)))
            out.append(t)
        return out
    t, out = base, []                                          # bursty: many equal neighbours and a few big gaps
    for _ in range(n):
        t += rng.choice([0, 0, 0, 1, 7, 10 ** rng.randrange(3, 10)])
        out.append(t)
    return out


def test_reference_table(env):
    """filter_time_test.go:13-88 TestFilterTime: one block, five rows; bounds in nanoseconds, both inclusive"""
    oracle, vs, pu, ctx = env
    ts = [1, 9, 123, 456, 789]
    blk = oracle.Block.from_columns([("_msg", [b"some value for row %d" % i for i in range(5)])]).set_timestamps(ts)
    F, G = oracle.Filter, vs.Filter
    for lo, hi, want in [(-10, 1, [0]), (-10, 10, [0, 1]), (1, 1, [0]), (2, 456, [1, 2, 3]), (2, 457, [1, 2, 3]), (120, 788, [2, 3]), (120, 789, [2, 3, 4]), (120, 10000, [2, 3, 4]),
                         (789, 1000, [4]), (-10, -1, []), (790, 1000, []), (1, 789, [0, 1, 2, 3, 4]), (-100, 10000, [0, 1, 2, 3, 4]), (5, 4, []), (10, 122, [])]:
        got = check(env, [blk], F.time(lo, hi), G.time(lo, hi))
        assert got[0] == want, (lo, hi)
    check(env, [blk], F.and_([F.time(2, 500), F.phrase("_msg", "row")]), G.and_([G.time(2, 500), G.phrase("_msg", "row")]))
    check(env, [blk], F.or_([F.time(2, 9), F.not_(F.time(0, 500))]), G.or_([G.time(2, 9), G.not_(G.time(0, 500))]))


def test_all_marshal_types_and_shapes(env):
    oracle, vs, pu, ctx = env
    rng = random.Random(8)
    F, G = oracle.Filter, vs.Filter
    seen = set()
    for trial in range(60):
        kind = rng.choice(["const", "step", "jitter", "bursty"])
        n = rng.choice([1, 2, 3, 17, 63, 64, 65, 300, 3000, 20000])
        ts = series(rng, kind, n)
        blk = oracle.Block.from_columns([("x", [b"v%d" % (i % 7) for i in range(n)]), ("y", [b"w%d" % i for i in range(n)])]).set_timestamps(ts)
        seen.add(blk.timestamps_block()[1])
        probes = [(ts[0], ts[-1]), (ts[0] + 1, ts[-1]), (ts[0], ts[-1] - 1), (ts[n // 2], ts[n // 2]), (ts[n // 3] + 1, ts[(2 * n) // 3]), (ts[-1] + 1, ts[-1] + 5), (ts[0] - 9, ts[0] - 1)]
        for lo, hi in probes:
            check(env, [blk], F.time(lo, hi), G.time(lo, hi))
        lo, hi = ts[n // 4], ts[(3 * n) // 4]
        check(env, [blk], F.and_([F.time(lo, hi), F.phrase("x", "v3")]), G.and_([G.time(lo, hi), G.phrase("x", "v3")]), stage=rng.choice(["ondisk", "decoded"]))
    assert seen >= {1, 2, 3, 5}, {MT_NAMES[m] for m in seen}
    # many blocks in one batch, each with its own encoding; the filter covers some fully, some partly, some not at all
    blocks, t0 = [], 1_700_000_000_000_000_000
    for bi in range(40):
        n = rng.choice([64, 100, 1000, 2500])
        ts = [t0 + v for v in series(rng, rng.choice(["const", "step", "jitter", "bursty"]), n)]
        ts = [v - ts[0] + t0 for v in ts]
        t0 = ts[-1] + rng.choice([0, 1, 10 ** 9])
        blocks.append(oracle.Block.from_columns([("x", [b"v%d" % (i % 5) for i in range(n)])]).set_timestamps(ts))
    lo = blocks[7].timestamps_block()[2] + 5
    hi = blocks[30].timestamps_block()[3] - 5
    check(env, blocks, F.time(lo, hi), G.time(lo, hi))
    check(env, blocks, F.and_([F.phrase("x", "v1"), F.not_(F.time(lo, hi))]), G.and_([G.phrase("x", "v1"), G.not_(G.time(lo, hi))]))


def test_malformed_timestamps_are_rejected(env):
    oracle, vs, pu, ctx = env
    ts = [10 + 7 * i + (i % 5) for i in range(500)]
    blk = oracle.Block.from_columns([("x", [b"v%d" % i for i in range(500)])]).set_timestamps(ts)
    d = pu.oracle_block_to_desc(blk)
    data, mt, mn, mx = d["timestamps"]
    assert mt in (1, 5)
    prog = vs.Program(vs.Filter.time(ts[100], ts[300]))
    if mt == 1:   # work on the plain form: what the device sees after inflating the frame
        import ctypes as C
        z = C.CDLL("libzstd.so.1"); z.ZSTD_getFrameContentSize.restype = C.c_ulonglong; z.ZSTD_decompress.restype = C.c_size_t
        n = z.ZSTD_getFrameContentSize(data, C.c_size_t(len(data))); out = C.create_string_buffer(n)
        assert z.ZSTD_decompress(out, C.c_size_t(n), data, C.c_size_t(len(data))) == n
        data, mt = out.raw[:n], 5
    good = dict(d, timestamps=(data, mt, mn, mx))
    words, counts, st = ctx.scan_batch(prog, vs.HostBlocks([b"x"], [good]))
    assert int(counts[0]) == 201
    for bad in (data[:-1], data + b"\x00", data[:5] + bytes([data[5] | 0x80]) + data[6:], b"\xff" * 11 + b"\x01" + data):
        with pytest.raises(vs.VlscanError):
            ctx.scan_batch(prog, vs.HostBlocks([b"x"], [dict(d, timestamps=(bad, mt, mn, mx))]))
    with pytest.raises(vs.VlscanError):   # a block without timestamps cannot answer a _time filter
        ctx.scan_batch(prog, vs.HostBlocks([b"x"], [{k: v for k, v in d.items() if k != "timestamps"}]))
    words, counts, st = ctx.scan_batch(prog, vs.HostBlocks([b"x"], [good]))   # the ctx stays usable
    assert int(counts[0]) == 201


def test_gather_timestamps_and_values(env):
    """The selected rows' `_time` and column values, as blockResult would read them, for every column kind; hits in block order, rows ascending."""
    oracle, vs, pu, ctx = env
    rng = random.Random(21)
    blocks, rows_all = [], []
    t0 = 1_700_000_000_000_000_000
    for bi in range(9):
        n = rng.choice([1, 64, 65, 300, 2100])
        ts = [t0 + v for v in series(rng, rng.choice(["const", "step", "jitter", "bursty"]), n)]
        ts = [v - ts[0] + t0 for v in ts]
        t0 = ts[-1] + 1
        cols = {
            "msg": [b"row %d of block %d %s" % (i, bi, b"x" * (i % 40)) if i % 7 else b"" for i in range(n)],
            "u16": [b"%d" % (i * 37 % 60000) for i in range(n)],
            "i64": [b"%d" % ((i - n // 2) * 987654321) for i in range(n)],
            "f64": [b"%d.%d" % (i * 7 - 900, 1 + i % 97) for i in range(n)],
            "ip": [b"10.%d.%d.%d" % (i % 3, i % 251, (i * 7) % 256) for i in range(n)],
            "ts": [b"2024-03-%02dT12:%02d:%02d.%03dZ" % (1 + i % 28, i % 60, (i * 7) % 60, i % 1000) for i in range(n)],
            "lvl": [[b"info", b"warn", b"error", b""][i % 4] for i in range(n)],
            "cst": [b"same value"] * n,
        }
        if bi % 3 == 2:
            del cols["ip"]       # a field some blocks do not have
        blk = oracle.Block.from_columns(list(cols.items())).set_timestamps(ts)
        blocks.append(blk)
        rows_all.append((ts, cols))
    hb = pu.host_blocks_from_oracle(blocks)
    batch = ctx.upload(hb)
    # what a reader of the stored block sees (blockResult: decode the values block, then valueType -> string).  Not always the ingested
    # string: "-158.10" passes tryParseFloat64Exact (values_encoder.go:788-848), is stored as the float and reads back as "-158.1".
    stored = []
    for bi, blk in enumerate(blocks):
        d = {name: [value] * blk.rows for name, value in blk.consts}
        for c in blk.columns:
            items = oracle.unmarshal_strings_block(c.values_block, blk.rows)
            d[c.name] = [c.dict[it[0]] if c.value_type == 2 else oracle.encoded_to_string(c.value_type, it) for it in items]   # 2 = valueTypeDict
        for name, vals_in in rows_all[bi][1].items():
            if name != "f64":
                assert d.get(name.encode(), [b""] * blk.rows) == vals_in, name   # a column of empty values is not stored at all
        stored.append(d)
    assert any(a != b for bi in range(len(blocks)) for a, b in zip(stored[bi].get(b"f64", []), rows_all[bi][1]["f64"]))   # the lossy case is covered
    F, G = oracle.Filter, vs.Filter
    for of, gf in [(F.phrase("lvl", "error"), G.phrase("lvl", "error")), (F.prefix("msg", "row 1"), G.prefix("msg", "row 1")), (F.noop(), G.noop()), (F.phrase("msg", "absent"), G.phrase("msg", "absent")),
                   (F.time(rows_all[2][0][0] + 1, rows_all[6][0][-1] - 1), G.time(rows_all[2][0][0] + 1, rows_all[6][0][-1] - 1))]:
        ctx.scan_resident(vs.Program(gf), batch)
        want_rows = [oracle.bitmap_rows(b.search(of), b.rows) for b in blocks]
        ts, offs = ctx.gather_timestamps(batch)
        assert [int(o) for o in offs] == [sum(len(w) for w in want_rows[:k]) for k in range(len(blocks) + 1)]
        assert list(ts) == [rows_all[bi][0][r] for bi, w in enumerate(want_rows) for r in w], gf
        for field in ("msg", "u16", "i64", "f64", "ip", "ts", "lvl", "cst", "nope"):
            vals, offs2 = ctx.gather_values(field, batch)
            assert list(offs2) == list(offs)
            want = [stored[bi].get(field.encode(), [b""] * blocks[bi].rows)[r] for bi, w in enumerate(want_rows) for r in w]
            assert vals == want, (gf, field)
    batch.free()
