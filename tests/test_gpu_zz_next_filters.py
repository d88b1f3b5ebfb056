"""GPU parity for the filter kinds of SURVEY §8(f) rank 3 that libvlscan compiles (exact_prefix, len_range, string_range, ipv4_range,
value_type, i(phrase), i(prefix*), seq(), contains_all(), contains_any(), eq_field(), le_field() / lt_field(), range()), through the C ABI
against the CPU oracle and against the reference's own tables (filter_{exact_prefix,len_range,string_range,ipv4_range,value_type,
any_case_phrase,any_case_prefix,sequence,contains_all,contains_any,eq_field,le_field,range}_test.go: all 994 cases of
tests/golden/filter_cases_next.json).
Bar: bit-exact bitmaps and counts.  (The file name sorts after the parity tests of the round-1 kinds on purpose.)"""
import random

import numpy as np
import pytest

from golden_util import load_filter_cases, build_filter

pytestmark = pytest.mark.gpu

KINDS = ("exact_prefix", "len_range", "string_range", "ipv4_range", "value_type", "any_case_phrase", "any_case_prefix", "sequence", "contains_all", "contains_any", "eq_field", "le_field", "range")
CASES = [c for c in load_filter_cases("filter_cases_next.json") if c["filter"]["kind"] in KINDS]


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
    assert st.rows_matched == sum(len(w) for w in want)
    return want


def test_reference_tables(env):
    oracle, vs, pu, ctx = env
    assert len(CASES) == 994   # every testFilterMatchForColumns case of the thirteen filters' test files
    for c in CASES:
        b = oracle.Block.from_columns(c["columns"])
        gf = build_filter(vs.Filter, c["filter"])
        got, counts, st = pu.gpu_rows(ctx, gf, [b])
        assert got[0] == c["expected"], (c["src"], c["filter"])
        assert int(counts[0]) == len(c["expected"])
    for c in CASES[::5]:
        b = oracle.Block.from_columns(c["columns"])
        got, counts, st = pu.gpu_rows(ctx, build_filter(vs.Filter, c["filter"]), [b], stage="decoded")
        assert got[0] == c["expected"], (c["src"], c["filter"])


def test_every_column_kind(env):
    """Each filter kind against every column encoding (string, dict, const, missing, uint8..uint64, int64, float64, ipv4, iso8601):
    differential against the oracle, which follows the per-type gates of the reference."""
    oracle, vs, pu, ctx = env
    n = 260
    cols = [
        ("u8", [b"%d" % (i % 200) for i in range(n)]),
        ("u16", [b"%d" % (i * 37 % 60000) for i in range(n)]),
        ("u32", [b"%d" % (i * 104729 % 4000000000) for i in range(n)]),
        ("u64", [b"%d" % (i * 1234567890123 + 5000000000) for i in range(n)]),
        ("i64", [b"%d" % ((i - 130) * 987654321) for i in range(n)]),
        ("f64", [b"%d.%d" % (i * 7 - 900, i % 97) for i in range(n)]),
        ("ip", [b"10.%d.%d.%d" % (i % 3, i % 251, (i * 7) % 256) for i in range(n)]),
        ("ts", [b"2024-03-%02dT12:%02d:%02d.%03dZ" % (1 + i % 28, i % 60, (i * 7) % 60, i % 1000) for i in range(n)]),
        ("lvl", [[b"info", b"warn", b"error", b"10.0.0.7", b""][i % 5] for i in range(n)]),
        ("cst", [b"same value"] * n),
        ("msg", [[b"", b"a", "йцук".encode(), b"row %d" % i, b"10.1.2.%d" % (i % 256), b"\xff\xfe", b"zz top"][i % 7] for i in range(n)]),
    ]
    blk = oracle.Block.from_columns(cols)
    vts = {c.name: c.value_type for c in blk.columns}
    assert (vts[b"u8"], vts[b"u16"], vts[b"u32"], vts[b"u64"], vts[b"i64"], vts[b"f64"], vts[b"ip"], vts[b"ts"], vts[b"lvl"], vts[b"msg"]) == (3, 4, 5, 6, 10, 7, 8, 9, 2, 1)
    F, G = oracle.Filter, vs.Filter
    fields = ["u8", "u16", "u32", "u64", "i64", "f64", "ip", "ts", "lvl", "cst", "msg", "missing"]
    probes = []
    for f in fields:
        for p in ["", "1", "10", "10.", "-", "-9", "2024-03", "row ", "same", "e", "9", ":", "0", "йц", "zz"]:
            probes.append(("exact_prefix", f, (p,)))
        for lo, hi in [(0, 0), (0, 1), (1, 1), (2, 4), (3, 3), (4, 2), (5, 12), (7, 15), (10, 10), (15, 24), (24, 24), (25, 30), (0, 1000), (21, 21), (20, 22)]:
            probes.append(("len_range", f, (lo, hi)))
        for lo, hi in [("", ""), ("", "z"), ("0", "9"), ("1", "2"), ("10.", "10.1"), ("-", "0"), ("-5", "-1"), ("+", ","), ("2024", "2025"), (":", "a"), ("a", "zzz"), ("b", "a"),
                       ("9", ":"), ("same", "samf"), (b"\xff", b"\xff\xff")]:
            probes.append(("string_range", f, (lo, hi)))
        for lo, hi in [(0, 0xFFFFFFFF), (0x0A000000, 0x0A00FFFF), (0x0A010000, 0x0A01FFFF), (0x0A000007, 0x0A000007), (5, 4), (0, 0), (0x0A0102FF, 0x0A010300)]:
            probes.append(("ipv4_range", f, (lo, hi)))
        for t in ["string", "dict", "const", "uint8", "uint16", "uint32", "uint64", "int64", "float64", "ipv4", "iso8601", "unknown", ""]:
            probes.append(("value_type", f, (t,)))
        for p in ["", "1", "10", "ROW", "Row 1", "SAME VALUE", "same", "ERROR", "Info", "10.0.0.7", "2024-03-05t12:04", "2024-03-05T12:04:28.004z", "t12", "ЙЦУК", "йц", "Zz", "ZZ TOP", "-", "e", "nan", "100.5", "-130"]:
            probes.append(("any_case_phrase", f, (p,)))
            probes.append(("any_case_prefix", f, (p,)))
        for lst in [[], [""], ["", ""], ["1"], ["10"], ["row", "1"], ["1", "row"], ["same", "value"], ["value", "same"], ["10", "0", "7"], ["10.0.0.7"], ["2024-03-05T12:04:28.004Z"], ["2024", "03"],
                    ["12", "04"], ["zz", "top"], ["", "top"], ["error"], ["info", "warn"], ["100"], ["100", "5"], ["-130"], ["7", "7"], ["1", "1", "1"], ["255"], ["йцук"], ["a", ""], ["no such"]]:
            probes.append(("sequence", f, (lst,)))
            probes.append(("contains_all", f, (lst,)))
            probes.append(("contains_any", f, (lst,)))
        for lo, hi in [(0, 0), (-1e9, 1e9), (1, 100), (100.5, 100.5), (10, 9), (-130 * 987654321, 0), (0.5, 2.5), (167772167, 167772167), (1.7e18, 1.8e18), (float("-inf"), float("inf")), (255, 1e30), (-0.0, 0.0),
                       (1709640000e9, 1709647200e9), (37, 37)]:
            probes.append(("range", f, (lo, hi)))
        for g in fields:
            probes.append(("eq_field", f, (g,)))
            probes.append(("le_field", f, (g, False)))
            probes.append(("le_field", f, (g, True)))
    for kind, field, args in probes:
        check(env, [blk], getattr(F, kind)(field, *args), getattr(G, kind)(field, *args))
    # two-column filters on columns made to agree on some rows: same type on both sides (binary compare), dict against dict, typed against its text
    n2 = 300
    cols2 = [("a", [b"%d" % (i % 50) for i in range(n2)]), ("b", [b"%d" % ((i * 7) % 50) for i in range(n2)]), ("d1", [[b"x", b"y", b"10", b"9"][i % 4] for i in range(n2)]),
             ("d2", [[b"x", b"10", b"y", b"100"][(i // 2) % 4] for i in range(n2)]), ("s", [b"%d" % (i % 50) if i % 3 else b"v%d" % i for i in range(n2)]),
             ("f1", [b"%d.5" % (i % 20) for i in range(n2)]), ("f2", [b"%d.5" % ((i * 3) % 20) for i in range(n2)]), ("i1", [b"%d" % (i % 20 - 10) for i in range(n2)]), ("i2", [b"%d" % ((i * 3) % 20 - 10) for i in range(n2)]),
             ("t1", [b"2024-03-%02dT00:00:00.000Z" % (1 + i % 9) for i in range(n2)]), ("t2", [b"2024-03-%02dT00:00:00.000Z" % (1 + (i * 5) % 9) for i in range(n2)]), ("c", [b"7"] * n2)]
    blk2 = oracle.Block.from_columns(cols2)
    names2 = [c[0] for c in cols2] + ["missing"]
    for f in names2:
        for g in names2:
            check(env, [blk2], F.eq_field(f, g), G.eq_field(f, g))
            check(env, [blk2], F.le_field(f, g, False), G.le_field(f, g, False))
            check(env, [blk2], F.le_field(f, g, True), G.le_field(f, g, True))
    # combinators mixing old and new kinds
    of = F.and_([F.exact_prefix("msg", "row"), F.or_([F.len_range("lvl", 4, 4), F.ipv4_range("ip", 0x0A010000, 0x0A01FFFF)]), F.not_(F.string_range("u8", "1", "2")), F.value_type("ts", "iso8601")])
    gf = G.and_([G.exact_prefix("msg", "row"), G.or_([G.len_range("lvl", 4, 4), G.ipv4_range("ip", 0x0A010000, 0x0A01FFFF)]), G.not_(G.string_range("u8", "1", "2")), G.value_type("ts", "iso8601")])
    check(env, [blk], of, gf)
    check(env, [blk], of, gf, stage="decoded")


def test_random_strings_many_blocks(env):
    oracle, vs, pu, ctx = env
    rng = random.Random(99)
    alphabet = ["a", "b", "0", "1", ".", " ", "-", "й", "日", "é", "Z"]
    blocks = []
    for bi in range(12):
        rows = rng.choice([1, 2, 63, 64, 65, 300])
        vals = []
        for _ in range(rows):
            k = rng.randrange(5)
            if k == 0: v = "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 12))).encode()
            elif k == 1: v = b"%d.%d.%d.%d" % tuple(rng.choice([0, 1, 9, 10, 99, 127, 255, 256]) for _ in range(4))
            elif k == 2: v = bytes(rng.getrandbits(8) for _ in range(rng.randrange(0, 6)))
            elif k == 3: v = b"prefix " + bytes(rng.choice(b"xyz") for _ in range(rng.randrange(0, 40)))
            else: v = b""
            vals.append(v)
        blocks.append(oracle.Block.from_columns([("m", vals), ("k", [b"k%d" % (i % 11) for i in range(rows)])]))
    F, G = oracle.Filter, vs.Filter
    for _ in range(140):
        kind = rng.choice(["exact_prefix", "len_range", "string_range", "ipv4_range", "any_case_phrase", "any_case_prefix", "sequence", "contains_all", "contains_any"])
        word = lambda: "".join(rng.choice(alphabet + ["A", "B", "Й", "É", "z"]) for _ in range(rng.randrange(0, 4)))
        if kind in ("any_case_phrase", "any_case_prefix"): args = (word() if rng.random() < 0.8 else "PREFIX X",)
        elif kind in ("sequence", "contains_all", "contains_any"): args = ([word() if rng.random() < 0.8 else rng.choice(["prefix", "k1", "k10", "255", "x"]) for _ in range(rng.randrange(0, 4))],)
        elif kind == "exact_prefix": args = ("".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 3))) if rng.random() < 0.7 else "prefix x",)
        elif kind == "len_range": args = tuple(sorted([rng.randrange(0, 14), rng.randrange(0, 50)]))
        elif kind == "string_range": args = tuple(sorted(["".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 3))) for _ in range(2)], key=lambda s: s.encode()))
        else: args = tuple(sorted([rng.getrandbits(32) >> rng.choice([0, 8, 24]), rng.getrandbits(32)]))
        for field in ("m", "k"):
            check(env, blocks, getattr(F, kind)(field, *args), getattr(G, kind)(field, *args))
