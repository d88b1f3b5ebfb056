"""The oracle's restatement of the part directory format (oracle/vlo_part.h).

The reference ships no part files, so the record layouts are pinned by the marshaled lengths its own tests assert
(lib/logstorage/block_header_test.go, index_block_header_test.go), by hand-derived byte vectors, and by writer -> files -> reader
round trips whose blocks must come back byte-identical and answer filters identically."""
import random
import struct

import numpy as np
import pytest

from oracle import vloracle as vo

NEAREST_DELTA2, ZSTD_NEAREST_DELTA = 5, 4      # encoding.MarshalType values (vm/lib/encoding/encoding.go:13-30)
VT = {v: k for k, v in vo.VT_NAMES.items()}

BH_GO = dict(account_id=123, project_id=456, id_lo=3443, id_hi=23434, uncompressed_size_bytes=4344, rows_count=1234, ts_block_offset=13234,
             ts_block_size=8843, min_timestamp=-4334, max_timestamp=23434, ts_marshal_type=NEAREST_DELTA2, columns_header_index_offset=8923481,
             columns_header_index_size=8989832, columns_header_offset=4384, columns_header_size=894)


def full(fields, d):
    return {k: d.get(k, 0) for k in fields}


def test_block_header_lengths_of_the_reference_tests():
    # block_header_test.go:29-56
    assert len(vo.marshal_block_header()) == 63
    data = vo.marshal_block_header(**BH_GO)
    assert len(data) == 73
    assert vo.unmarshal_block_headers(data) == [full(vo.BLOCK_HEADER_FIELDS, BH_GO)]
    # hand-derived bytes: streamID big endian, then varuints, then the 33-byte timestampsHeader
    assert data[:24] == struct.pack(">IIQQ", 123, 456, 23434, 3443)
    assert data[24:26] == bytes([0xF8, 0x21]) and data[26:28] == bytes([0xD2, 0x09])          # 4344, 1234
    assert data[28:61] == struct.pack(">QQqqB", 13234, 8843, -4334, 23434, NEAREST_DELTA2)
    # :410-440
    assert vo.unmarshal_block_headers(b"") == []
    second = dict(BH_GO, uncompressed_size_bytes=89894, columns_header_index_offset=1234, columns_header_index_size=89324, columns_header_offset=12332, columns_header_size=234)
    both = vo.marshal_block_header() + vo.marshal_block_header(**second)
    assert len(both) == 134
    assert vo.unmarshal_block_headers(both) == [full(vo.BLOCK_HEADER_FIELDS, {}), full(vo.BLOCK_HEADER_FIELDS, second)]
    # format v0 has no columnsHeaderIndex fields
    v0 = data[:61] + data[-4:]
    got = vo.unmarshal_block_headers(v0, format_version=0)[0]
    assert got["columns_header_offset"] == 4384 and got["columns_header_size"] == 894 and got["columns_header_index_size"] == 0


def test_block_header_unmarshal_failures():
    # block_header_test.go:151-198: nil is an empty list for unmarshalBlockHeaders, every proper prefix of a header fails
    bh = dict(BH_GO, columns_header_index_offset=89434, columns_header_index_size=89123)
    data = vo.marshal_block_header(**bh)
    for n in range(1, len(data)):
        with pytest.raises(RuntimeError):
            vo.unmarshal_block_headers(data[:n])
    with pytest.raises(RuntimeError):
        vo.unmarshal_block_headers(b"foo")
    # limits: rowsCount <= maxRowsPerBlock, columnsHeaderSize <= maxColumnsHeaderSize
    with pytest.raises(RuntimeError, match="rowsCount"):
        vo.unmarshal_block_headers(vo.marshal_block_header(**dict(bh, rows_count=8 * 1024 * 1024 + 1)))
    assert vo.unmarshal_block_headers(vo.marshal_block_header(**dict(bh, rows_count=8 * 1024 * 1024)))
    with pytest.raises(RuntimeError, match="columnsHeaderSize"):
        vo.unmarshal_block_headers(vo.marshal_block_header(**dict(bh, columns_header_size=(8 << 20) + 1)))
    # validateBlockHeaders :186-204: streamIDs ascend; minTimestamp ascends within a stream
    a = vo.marshal_block_header(**bh)
    smaller_sid = vo.marshal_block_header(**dict(bh, id_lo=3442))
    earlier = vo.marshal_block_header(**dict(bh, min_timestamp=-4335))
    later = vo.marshal_block_header(**dict(bh, min_timestamp=-4333))
    assert len(vo.unmarshal_block_headers(a + later)) == 2 and len(vo.unmarshal_block_headers(smaller_sid + a)) == 2
    assert len(vo.unmarshal_block_headers(a + vo.marshal_block_header(**dict(bh, id_hi=23435, min_timestamp=-99999)))) == 2
    for bad in (a + smaller_sid, a + earlier, a + vo.marshal_block_header(**dict(bh, account_id=122))):
        with pytest.raises(RuntimeError, match="smaller"):
            vo.unmarshal_block_headers(bad)


def test_index_block_header_lengths_of_the_reference_tests():
    # index_block_header_test.go:27-44,127-141
    assert len(vo.marshal_index_block_header()) == 56
    f = dict(account_id=123, project_id=456, id_hi=214, id_lo=2111, min_timestamp=1234, max_timestamp=898943, index_block_offset=234, index_block_size=898)
    data = vo.marshal_index_block_header(**f)
    assert data == struct.pack(">IIQQqqQQ", 123, 456, 214, 2111, 1234, 898943, 234, 898)
    assert vo.unmarshal_index_block_headers(data) == [f]
    two = vo.marshal_index_block_header(index_block_offset=234, index_block_size=5432) + vo.marshal_index_block_header(min_timestamp=-123)
    assert len(two) == 112
    got = vo.unmarshal_index_block_headers(two)
    assert [g["index_block_size"] for g in got] == [5432, 0] and got[1]["min_timestamp"] == -123
    assert vo.unmarshal_index_block_headers(b"") == []
    for n in range(1, 56):
        with pytest.raises(RuntimeError):
            vo.unmarshal_index_block_headers(data[:n])
    with pytest.raises(RuntimeError, match="smaller streamID"):
        vo.unmarshal_index_block_headers(data + vo.marshal_index_block_header(**dict(f, project_id=455)))


def test_columns_header_index_lengths_of_the_reference_tests():
    # block_header_test.go:75-94
    assert vo.marshal_columns_header_index([], []) == b"\x00\x00"
    data = vo.marshal_columns_header_index([(234, 123432), (23898, 0)], [(0, 8989)])
    assert len(data) == 14
    assert data == bytes([2, 0xEA, 0x01, 0xA8, 0xC4, 0x07, 0xDA, 0xBA, 0x01, 0x00, 1, 0x00, 0x9D, 0x46])
    assert vo.unmarshal_columns_header_index(data) == ([(234, 123432), (23898, 0)], [(0, 8989)])
    for bad in (b"", b"foo", data[:-1], data + b"\x00", b"\x05\x01"):
        with pytest.raises(RuntimeError):
            vo.unmarshal_columns_header_index(bad)


def test_column_header_lengths_of_the_reference_tests():
    # block_header_test.go:467-479 (the name is not part of the record since format v1)
    data = vo.marshal_column_header(value_type=VT["uint8"])
    assert data == bytes([VT["uint8"], 0, 0, 0, 0, 0, 0])
    d = vo.marshal_column_header(dict_values=[b"abc"], value_type=VT["dict"], values_offset=12345, values_size=254452)
    assert len(d) == 11
    assert d == bytes([VT["dict"], 1, 3]) + b"abc" + bytes([0xB9, 0x60, 0xF4, 0xC3, 0x0F])
    got, used = vo.unmarshal_column_header(d + b"tail")
    assert used == 11 and got["dict"] == [b"abc"] and got["values_offset"] == 12345 and got["values_size"] == 254452 and got["bloom_filter_size"] == 0
    # every value type round-trips its min / max through its own width; int64 is zig-zag coded (encoding.MarshalInt64), float64 keeps the bits
    cases = [("string", 0, 0, 1 + 4), ("uint8", 3, 250, 1 + 2 + 4), ("uint16", 300, 65535, 1 + 4 + 4), ("uint32", 70000, 2 ** 32 - 1, 1 + 8 + 4), ("uint64", 2 ** 40, 2 ** 64 - 1, 1 + 16 + 4),
             ("int64", (-5) & (2 ** 64 - 1), 7, 1 + 16 + 4), ("float64", struct.unpack("<Q", struct.pack("<d", -1.5))[0], struct.unpack("<Q", struct.pack("<d", 2.25))[0], 1 + 16 + 4),
             ("ipv4", 0x01020304, 0xFFFFFFFF, 1 + 8 + 4), ("iso8601", 1, 2 ** 62, 1 + 16 + 4)]
    for name, mn, mx, size in cases:
        f = dict(value_type=VT[name], min_value=mn, max_value=mx, values_offset=1, values_size=2, bloom_filter_offset=3, bloom_filter_size=4)
        data = vo.marshal_column_header(**f)
        assert len(data) == size, name
        got, used = vo.unmarshal_column_header(data)
        got.pop("dict")
        assert used == size and got == f, name
        for n in range(len(data)):
            with pytest.raises(RuntimeError):
                vo.unmarshal_column_header(data[:n])
    assert vo.marshal_column_header(value_type=VT["int64"], min_value=(-5) & (2 ** 64 - 1), max_value=7)[1:17] == struct.pack(">QQ", 9, 14)
    # block_header_test.go:482-507 and the size limits of unmarshalValues / unmarshalBloomFilters
    for bad in (b"", b"foo", bytes([0]), bytes([11]), vo.marshal_column_header(value_type=VT["string"], values_size=(8 << 20) + 1),
                vo.marshal_column_header(value_type=VT["string"], bloom_filter_size=(8 << 20) + 1)):
        with pytest.raises(RuntimeError):
            vo.unmarshal_column_header(bad)
    with pytest.raises(RuntimeError, match="max 8"):
        vo.marshal_column_header(dict_values=[b"%d" % i for i in range(9)], value_type=VT["dict"])


def test_columns_header_lengths_of_the_reference_tests():
    # block_header_test.go:120-149
    assert vo.marshal_columns_header([], []) == (b"\x00\x00", b"\x00\x00")
    cols = [("foobar", dict(value_type=VT["string"], values_offset=12345, values_size=23434, bloom_filter_offset=89843, bloom_filter_size=8934)),
            ("message", dict(value_type=VT["uint16"], min_value=123, max_value=456, values_offset=3412345, values_size=234434, bloom_filter_offset=83, bloom_filter_size=34))]
    csh, idx = vo.marshal_columns_header(cols, [("foo", "bar")])
    assert len(csh) == 31
    # the index holds (columnNameID, offset of the record inside the columnsHeader): ids are handed out in order of first use
    assert vo.unmarshal_columns_header_index(idx) == ([(0, 1), (1, 12)], [(2, 27)])
    assert vo.unmarshal_column_header(csh[1:])[1] == 11 and vo.unmarshal_column_header(csh[12:])[1] == 14
    assert csh[26] == 1 and csh[27:] == b"\x03bar"
    same, names = vo.columns_header_roundtrip(csh, idx, ["foobar", "message", "foo"])
    assert same and names == [b"foobar", b"message", b"foo"]
    # setColumnNames failures :350-380
    with pytest.raises(RuntimeError, match="columnNameID"):
        vo.columns_header_roundtrip(csh, idx, ["foobar", "message"])
    with pytest.raises(RuntimeError, match="number of column headers"):
        vo.columns_header_roundtrip(csh, vo.marshal_columns_header_index([(0, 1)], [(2, 27)]), ["a", "b", "c"])
    with pytest.raises(RuntimeError, match="number of const columns"):
        vo.columns_header_roundtrip(csh, vo.marshal_columns_header_index([(0, 1), (1, 12)], []), ["a", "b", "c"])
    # block_header_test.go:239-288
    for n in range(len(csh)):
        with pytest.raises(RuntimeError):
            vo.columns_header_roundtrip(csh[:n], idx, ["foobar", "message", "foo"])
    with pytest.raises(RuntimeError, match="tail"):
        vo.columns_header_roundtrip(csh + b"\x00", idx, ["foobar", "message", "foo"])


def test_part_header_json():
    f = dict(FormatVersion=3, CompressedSizeBytes=1234, UncompressedSizeBytes=56789, RowsCount=100, BlocksCount=4, MinTimestamp=-5, MaxTimestamp=1700000000000000000, BloomValuesShardsCount=7)
    text = vo.part_header_json(**f)
    # encoding/json: members in declaration order, no whitespace
    assert text == (b'{"FormatVersion":3,"CompressedSizeBytes":1234,"UncompressedSizeBytes":56789,"RowsCount":100,"BlocksCount":4,'
                    b'"MinTimestamp":-5,"MaxTimestamp":1700000000000000000,"BloomValuesShardsCount":7}')
    import json
    assert json.loads(text) == f
    assert vo.part_header_parse(text) == f
    assert vo.part_header_parse(json.dumps(f, indent=2)) == f
    assert vo.part_header_parse(b'{"RowsCount": 5, "Extra": "x\\"y", "Other": 1.5e3, "BlocksCount":5}') == dict(full(vo.PART_HEADER_FIELDS, {}), RowsCount=5, BlocksCount=5)
    assert vo.part_header_parse(b"{}") == full(vo.PART_HEADER_FIELDS, {})
    # part_header.go:62-83
    assert vo.part_header_parse(b'{"FormatVersion":1,"RowsCount":1}')["BloomValuesShardsCount"] == 8
    assert vo.part_header_parse(b'{"FormatVersion":0,"RowsCount":1}')["BloomValuesShardsCount"] == 0
    for bad in (b"", b"[]", b'{"FormatVersion":4}', b'{"FormatVersion":1,"BloomValuesShardsCount":3}', b'{"MinTimestamp":2,"MaxTimestamp":1}', b'{"RowsCount":1,"BlocksCount":2}',
                b'{"RowsCount":-1}', b'{"RowsCount":1.5}', b'{"RowsCount":1} x', b'{"RowsCount":1'):
        with pytest.raises(RuntimeError):
            vo.part_header_parse(bad)


# ---- writer -> files -> reader ------------------------------------------------------------------------------------------------
def make_block(rng, rows, with_msg=True, extra=None):
    words = ["error", "warn", "GET", "POST", "/api/v1/query", "timeout", "connection", "reset", "peer", "10.0.0.1", "abc_def", "кириллица", "x" * 300]
    cols = []
    if with_msg:
        cols.append(("_msg", [" ".join(rng.choice(words) for _ in range(rng.randrange(1, 8))) for _ in range(rows)]))
    cols.append(("level", [rng.choice(["info", "warn", "error"]) for _ in range(rows)]))                       # dict
    cols.append(("status", [str(rng.choice([200, 204, 301, 404, 500, 502, 503, 504, 400, 401])) for _ in range(rows)]))   # uint16 (10 distinct values)
    cols.append(("bytes", [str(rng.randrange(0, 1 << 40)) for _ in range(rows)]))                              # uint64
    cols.append(("delta", [str(rng.randrange(-1000, 1000)) for _ in range(rows)]))                             # int64
    cols.append(("ratio", ["%d.%03d" % (rng.randrange(100), rng.randrange(1000)) for _ in range(rows)]))       # float64
    cols.append(("ip", ["%d.%d.%d.%d" % tuple(rng.randrange(256) for _ in range(4)) for _ in range(rows)]))    # ipv4
    cols.append(("ts", ["2024-%02d-%02dT%02d:00:00.%03dZ" % (rng.randrange(1, 13), rng.randrange(1, 29), rng.randrange(24), rng.randrange(1000)) for _ in range(rows)]))
    cols.append(("host", ["host-1"] * rows))                                                                   # const
    cols.append(("sparse", [rng.choice(["", "", "v%d" % rng.randrange(1000)]) for _ in range(rows)]))         # string with empty values
    for name, vals in extra or []:
        cols.append((name, vals))
    b = vo.Block.from_columns(cols, rows)
    return b, cols


def sorted_ts(rng, rows, base):
    t, out = base, []
    for _ in range(rows):
        t += rng.choice([0, 1, 1000, 12345678])
        out.append(t)
    return out


def block_signature(b):
    cols = sorted(((c.name if c.name != b"_msg" else b"", c.name, c.value_type, c.min_value, c.max_value, tuple(c.dict), c.values_block, c.bloom) for c in b.columns))
    consts = sorted((n if n != b"_msg" else b"", n, v) for n, v in b.consts)
    return [c[1:] for c in cols], [c[1:] for c in consts], b.rows, b.timestamps_block()


def uncompressed_size(cols, rows):
    # block.uncompressedSizeBytes block.go:48-80: 3 + 10 + len(time.RFC3339Nano) per row, 6 + len(name) + len(value) per non-empty field
    n = (3 + 10 + 35) * rows
    for name, vals in cols:
        n += sum(6 + len(name.encode()) + len(v.encode()) for v in vals if v)
    return n


@pytest.fixture(scope="module")
def written_part():
    rng = random.Random(11)
    w = vo.PartWriter(max_index_block=150, max_shards=3)
    originals = []
    base = 1_700_000_000_000_000_000
    sids = [(0, 0, 1, 5), (0, 0, 1, 9), (0, 7, 0, 1), (3, 0, 0, 0)]
    for si, sid in enumerate(sids):
        t = base + si * 10_000
        for k in range(3):
            rows = rng.choice([1, 2, 17, 64, 257])
            extra = [("only_in_%d" % si, ["u%d" % rng.randrange(50) for _ in range(rows)])] if k == 1 else None
            b, cols = make_block(rng, rows, with_msg=not (si == 2 and k == 0), extra=extra)
            ts = sorted_ts(rng, rows, t)
            t = ts[-1]
            b.set_timestamps(ts)
            u = uncompressed_size(cols, rows)
            w.add_block(sid, b, u)
            originals.append((sid, b, cols, ts, u))
    files = w.finalize()
    return w, files, originals


def test_part_round_trip_is_byte_identical(written_part, tmp_path):
    w, files, originals = written_part
    path = tmp_path / "part"
    vo.save_part(files, str(path))
    with pytest.raises(FileExistsError):
        vo.save_part(files, str(path))
    loaded = vo.load_part(str(path))
    assert loaded == files
    r = vo.PartReader(loaded)
    assert r.header == w.header
    hdr = r.header
    assert hdr["FormatVersion"] == 3 and hdr["BlocksCount"] == len(originals) == r.nblocks and hdr["RowsCount"] == sum(o[1].rows for o in originals)
    assert hdr["MinTimestamp"] == min(o[3][0] for o in originals) and hdr["MaxTimestamp"] == max(o[3][-1] for o in originals)
    assert hdr["UncompressedSizeBytes"] == sum(o[4] for o in originals)
    assert hdr["CompressedSizeBytes"] == sum(len(v) for k, v in files.items() if k != "metadata.json")
    assert hdr["BloomValuesShardsCount"] == 3
    assert sorted(files) == sorted(["metadata.json", "column_names.bin", "column_idxs.bin", "metaindex.bin", "index.bin", "columns_header_index.bin", "columns_header.bin",
                                    "timestamps.bin", "message_bloom.bin", "message_values.bin"] + ["bloom.bin%d" % i for i in range(3)] + ["values.bin%d" % i for i in range(3)])
    # the message field is stored under the empty name and never enters the shards
    assert b"_msg" not in r.column_names and len(set(r.column_names)) == len(r.column_names)
    for i, (sid, b, cols, ts, u) in enumerate(originals):
        bh = r.block_header(i)
        assert (bh["account_id"], bh["project_id"], bh["id_hi"], bh["id_lo"]) == sid
        assert bh["rows_count"] == b.rows and bh["uncompressed_size_bytes"] == u and bh["min_timestamp"] == ts[0] and bh["max_timestamp"] == ts[-1]
        got = r.block(i)
        assert block_signature(got) == block_signature(b), i
        data, mt, mn, mx = got.timestamps_block()
        assert list(vo.unmarshal_timestamps(data, mt, mn, got.rows)) == ts
        # columns come back sorted by their on-disk name (block.sortColumnsByName)
        raw = [c.name if c.name != b"_msg" else b"" for c in got.columns]
        assert raw == sorted(raw)
        # blockSearch.getColumnHeader / getConstColumnValue through the columnsHeaderIndex
        for c in b.columns:
            ch = r.column_header(i, c.name)
            assert ch is not None and ch["value_type"] == c.value_type and ch["min_value"] == c.min_value and ch["max_value"] == c.max_value
            assert ch["values_size"] == len(c.values_block) and ch["bloom_filter_size"] == (0 if c.value_type == VT["dict"] else len(c.bloom))
            assert r.const_value(i, c.name) == b""
        for name, value in b.consts:
            assert r.const_value(i, name) == value and r.column_header(i, name) is None
        assert r.column_header(i, "no_such_column") is None and r.const_value(i, "no_such_column") == b""


def test_part_blocks_answer_filters_like_the_originals(written_part):
    w, files, originals = written_part
    r = vo.PartReader(files)
    F = vo.Filter
    filters = [F.phrase("_msg", "error"), F.prefix("_msg", "time"), F.exact("level", "warn"), F.in_("status", ["404", "500"]), F.phrase("bytes", "12"), F.regexp("_msg", "conn.*peer"),
               F.ipv4_range("ip", 0x0A000000, 0x7FFFFFFF), F.exact("host", "host-1"), F.phrase("sparse", ""), F.not_(F.exact_prefix("ts", "2024-0")),
               F.and_([F.phrase("_msg", "GET"), F.or_([F.exact("level", "info"), F.range("delta", -10, 500)])]), F.string_range("ratio", "2", "7"), F.exact("only_in_1", "u7")]
    some = 0
    for i, (sid, b, cols, ts, u) in enumerate(originals):
        got = r.block(i)
        for f in filters:
            a, c = b.search(f), got.search(f)
            assert np.array_equal(a, c), (i, f.tokens())
            some += int(np.count_nonzero(a))
        t0 = ts[len(ts) // 3]
        ft = F.time(t0, ts[-1] - 1)
        assert np.array_equal(b.search(ft), got.search(ft))
    assert some > 100


def test_part_layout_invariants(written_part):
    """what blockStreamReader checks while it streams a part (block_stream_reader.go): every file is consumed front to back without gaps"""
    w, files, originals = written_part
    r = vo.PartReader(files)
    ihs = r.index_block_headers()
    assert len(ihs) > 3                                         # max_index_block=150 forces several index blocks
    off = 0
    for ih in ihs:
        assert ih["index_block_offset"] == off
        off += ih["index_block_size"]
    assert off == len(files["index.bin"])
    bhs = [r.block_header(i) for i in range(r.nblocks)]
    for key, size_key, fname in (("ts_block_offset", "ts_block_size", "timestamps.bin"), ("columns_header_offset", "columns_header_size", "columns_header.bin"),
                                 ("columns_header_index_offset", "columns_header_index_size", "columns_header_index.bin")):
        off = 0
        for bh in bhs:
            assert bh[key] == off, key
            off += bh[size_key]
        assert off == len(files[fname])
    # each index block header carries the first streamID and the time range of its blocks (mustWriteIndexBlock index_block_header.go:38-51)
    k = 0
    import ctypes as C
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
    for ih in ihs:
        frame = files["index.bin"][ih["index_block_offset"]:ih["index_block_offset"] + ih["index_block_size"]]
        n = z.ZSTD_getFrameContentSize(frame, C.c_size_t(len(frame)))
        out = C.create_string_buffer(n)
        z.ZSTD_decompress.restype = C.c_size_t
        assert z.ZSTD_decompress(out, C.c_size_t(n), frame, C.c_size_t(len(frame))) == n
        group = vo.unmarshal_block_headers(out.raw)
        assert group == bhs[k:k + len(group)]
        assert (ih["account_id"], ih["project_id"], ih["id_hi"], ih["id_lo"]) == tuple(group[0][x] for x in ("account_id", "project_id", "id_hi", "id_lo"))
        assert ih["min_timestamp"] == min(g["min_timestamp"] for g in group) and ih["max_timestamp"] == max(g["max_timestamp"] for g in group)
        k += len(group)
    assert k == len(bhs)
    # values and bloom bytes of every column are accounted for exactly once across message_* and the shards
    total_values = sum(len(c.values_block) for o in originals for c in o[1].columns)
    total_bloom = sum(len(c.bloom) for o in originals for c in o[1].columns if c.value_type != VT["dict"])
    assert total_values == len(files["message_values.bin"]) + sum(len(files["values.bin%d" % i]) for i in range(3))
    assert total_bloom == len(files["message_bloom.bin"]) + sum(len(files["bloom.bin%d" % i]) for i in range(3))
    # column -> shard assignment is round robin in order of first appearance (getBloomValuesWriterForColumnName :181-211)
    names = [n for n in r.column_names]
    first_seen = []
    for sid, b, cols, ts, u in originals:
        for raw in sorted((c.name if c.name != b"_msg" else b"") for c in b.columns):
            if raw and raw not in first_seen:
                first_seen.append(raw)
    assert set(first_seen) <= set(names)


def test_writer_rejects_out_of_order_blocks():
    rng = random.Random(5)
    b, _ = make_block(rng, 4)
    b.set_timestamps([10, 20, 30, 40])
    early, _ = make_block(rng, 2)
    early.set_timestamps([5, 50])
    w = vo.PartWriter()
    w.add_block((0, 0, 0, 2), b)
    with pytest.raises(RuntimeError, match="smaller than the previously written sid"):
        w.add_block((0, 0, 0, 1), b)
    with pytest.raises(RuntimeError, match="timestamp smaller"):
        w.add_block((0, 0, 0, 2), early)
    nots, _ = make_block(rng, 2)
    with pytest.raises(RuntimeError, match="needs timestamps"):
        w.add_block((0, 0, 0, 3), nots)
    files = w.finalize()
    r = vo.PartReader(files)
    assert r.nblocks == 1 and r.header["RowsCount"] == 4 and r.header["BloomValuesShardsCount"] >= 1
    # an empty part is still a valid directory
    e = vo.PartWriter().finalize()
    r = vo.PartReader(e)
    assert r.nblocks == 0 and r.header["RowsCount"] == 0 and r.index_block_headers() == []


def test_reader_rejects_damaged_parts(written_part):
    w, files, originals = written_part
    for name in files:
        if name.startswith(("bloom.bin", "values.bin", "message_")) or name in ("timestamps.bin", "columns_header.bin", "columns_header_index.bin"):
            continue                                             # data files are only touched when a block is read
        with pytest.raises(RuntimeError):
            vo.PartReader({k: v for k, v in files.items() if k != name})
        if files[name]:
            with pytest.raises(RuntimeError):
                vo.PartReader(dict(files, **{name: files[name][:-1]}))
    for name in ("timestamps.bin", "columns_header.bin", "columns_header_index.bin", "values.bin0", "bloom.bin2"):
        with pytest.raises(RuntimeError):
            vo.PartReader({k: v for k, v in files.items() if k != name})
    # truncated data files open fine, reading the block whose bytes are gone fails
    r = vo.PartReader(dict(files, **{"timestamps.bin": files["timestamps.bin"][:-1]}))
    r.block(0)
    with pytest.raises(RuntimeError, match="outside the file"):
        r.block(r.nblocks - 1)
    # counters that disagree with the index
    import json
    meta = json.loads(files["metadata.json"])
    for key in ("BlocksCount", "RowsCount"):
        with pytest.raises(RuntimeError, match=key):
            vo.PartReader(dict(files, **{"metadata.json": json.dumps(dict(meta, **{key: meta[key] + 1})).encode()}))
    with pytest.raises(RuntimeError, match="shardIdx"):
        vo.PartReader(dict(files, **{"metadata.json": json.dumps(dict(meta, BloomValuesShardsCount=2)).encode()}))
    # random damage never crashes the reader
    rng = random.Random(3)
    small = [k for k in files if k in ("metadata.json", "column_names.bin", "column_idxs.bin", "metaindex.bin", "index.bin", "columns_header_index.bin", "columns_header.bin")]
    opened = failed = 0
    for _ in range(300):
        name = rng.choice(small)
        b = bytearray(files[name])
        for _ in range(rng.randrange(1, 4)):
            b[rng.randrange(len(b))] = rng.getrandbits(8)
        try:
            r = vo.PartReader(dict(files, **{name: bytes(b)}))
            opened += 1
            for i in range(r.nblocks):
                try:
                    r.block(i)
                except RuntimeError:
                    pass
        except RuntimeError:
            failed += 1
    assert failed > 50 and opened + failed == 300
