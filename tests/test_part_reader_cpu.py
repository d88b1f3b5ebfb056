"""The product's part directory reader (victorialogs_b200/csrc/vl_part.h through vlscan_part_*) on the CPU: part directories written by the
oracle's restatement of the reference writer (oracle/vlo_part.h) are opened with the metadata inflated by libzstd through the
`inflate` callback (the default - the device decoder - is covered by tests/test_gpu_zz_part.py), and every descriptor the reader hands
out is compared with the block the writer was given."""
import ctypes as C
import json
import os
import random
import shutil

import pytest

from oracle import vloracle as vo
from victorialogs_b200 import scan as vs
from test_oracle_part import make_block, sorted_ts, uncompressed_size


def libzstd_inflate(frame, size):
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_decompress.restype = C.c_size_t
    out = C.create_string_buffer(max(size, 1))
    got = z.ZSTD_decompress(out, C.c_size_t(size), frame, C.c_size_t(len(frame)))
    assert got == size, got
    return out.raw[:size]


def write_part(tmp_path, name, max_shards=3, max_index_block=150, seed=11, streams=4, blocks_per_stream=3):
    rng = random.Random(seed)
    w = vo.PartWriter(max_index_block=max_index_block, max_shards=max_shards)
    originals = []
    base = 1_700_000_000_000_000_000
    for si in range(streams):
        sid = (si // 2, si % 2, 1, 5 + si)
        t = base + si * 10_000
        for k in range(blocks_per_stream):
            rows = rng.choice([1, 2, 17, 64, 257])
            extra = [("only_in_%d" % si, ["u%d" % rng.randrange(50) for _ in range(rows)])] if k == 1 else None
            b, cols = make_block(rng, rows, with_msg=not (si == 2 and k == 0), extra=extra)
            ts = sorted_ts(rng, rows, t)
            t = ts[-1] + 1
            b.set_timestamps(ts)
            w.add_block(sid, b, uncompressed_size(cols, rows))
            originals.append((sid, b, ts))
    files = w.finalize()
    path = str(tmp_path / name)
    vo.save_part(files, path)
    return path, files, originals, w.header


def all_fields(originals):
    names = []
    for _, b, _ in originals:
        for c in b.columns:
            if c.name not in names:
                names.append(c.name)
        for n, _ in b.consts:
            if n not in names:
                names.append(n)
    return names


def check_block(hb, i, fields, b):
    want_cols = {c.name: c for c in b.columns}
    want_consts = dict(b.consts)
    for f in fields:
        got = hb.column(i, f)
        if f in want_cols:
            c = want_cols[f]
            assert got is not None and got["kind"] == "values", f
            assert got["value_type"] == c.value_type and got["min_value"] == c.min_value and got["max_value"] == c.max_value, f
            assert got["values_block"] == c.values_block and got["dict"] == c.dict, f
            assert got["bloom"] == (b"" if c.value_type == 2 else c.bloom), f
        elif f in want_consts:
            assert got == dict(kind="const", value=want_consts[f]), f
        else:
            assert got is None, f
    assert hb.rows[i] == b.rows


def test_reader_hands_out_what_the_writer_was_given(tmp_path):
    path, files, originals, header = write_part(tmp_path, "part")
    p = vs.Part(path, inflate=libzstd_inflate)
    assert p.header == header and p.nblocks == len(originals)
    ref = vo.PartReader(files)
    assert p.column_names == ref.column_names and b"" in p.column_names
    for i in range(p.nblocks):
        assert p.block_header(i) == ref.block_header(i)
    fields = all_fields(originals) + [b"no_such_field"]
    hb = p.blocks(fields)
    assert hb.nblocks == len(originals) and hb.source == list(range(len(originals)))
    for i, (sid, b, ts) in enumerate(originals):
        check_block(hb, i, fields, b)
        # the timestamps block: the bytes the writer stored, decodable with the header fields alone
        bh = p.block_header(i)
        data = p.timestamps(i)
        assert (data, bh["ts_marshal_type"], bh["min_timestamp"], bh["max_timestamp"]) == b.timestamps_block()
        assert list(vo.unmarshal_timestamps(data, bh["ts_marshal_type"], bh["min_timestamp"], bh["rows_count"])) == ts
    with pytest.raises(vs.VlscanError, match="outside the part"):
        p.timestamps(p.nblocks)
    # a field list in another order, "" as the name of the message field, a sub-range of blocks
    sub = [b"status", b"", b"host"]
    hb2 = p.blocks(sub, lo=2, hi=7)
    assert hb2.source == [2, 3, 4, 5, 6] and hb2.field_names == [b"status", b"_msg", b"host"]
    for j, i in enumerate(hb2.source):
        check_block(hb2, j, [b"status", b"_msg", b"host"], originals[i][1])
    # no fields at all: just the row counts
    hb3 = p.blocks([])
    assert hb3.rows == [o[1].rows for o in originals]
    with pytest.raises(vs.VlscanError, match="duplicate field"):
        p.blocks([b"_msg", b""])
    with pytest.raises(vs.VlscanError, match="block range"):
        p.blocks([b"_msg"], lo=3, hi=p.nblocks + 1)
    with pytest.raises(vs.VlscanError, match="outside the part"):
        p.block_header(p.nblocks)


def test_time_range_selects_overlapping_blocks(tmp_path):
    path, files, originals, header = write_part(tmp_path, "part", seed=5)
    p = vs.Part(path, inflate=libzstd_inflate)
    rng = random.Random(1)
    stamps = sorted(t for _, _, ts in originals for t in (ts[0], ts[-1]))
    for _ in range(40):
        a, b = sorted((rng.choice(stamps) + rng.choice([-1, 0, 1]), rng.choice(stamps) + rng.choice([-1, 0, 1])))
        want = [i for i, (_, _, ts) in enumerate(originals) if not (ts[-1] < a or ts[0] > b)]
        assert p.blocks([b"_msg"], min_timestamp=a, max_timestamp=b).source == want
    assert p.blocks([b"_msg"], min_timestamp=header["MaxTimestamp"] + 1).source == []
    assert p.blocks([b"_msg"], max_timestamp=header["MinTimestamp"] - 1).source == []


def test_older_formats_pick_the_shard_by_name_hash(tmp_path):
    # format v2 (part.go:204-212): shard = xxhash64(name) % BloomValuesShardsCount, no column_idxs.bin.  Built from a v3 part that gave every
    # column a shard of its own by moving the shard files to where the hash points.
    path, files, originals, header = write_part(tmp_path, "v3", max_shards=64, streams=2)
    all_names = vo.PartReader(files).column_names

    def varuints(data):
        out, v, sh = [], 0, 0
        for byte in data:
            v |= (byte & 0x7F) << sh
            sh += 7
            if byte < 0x80:
                out.append(v)
                v = sh = 0
        return out

    nums = varuints(files["column_idxs.bin"])       # count, then (columnID, shardIdx) pairs (column_names.go:34-40)
    idxs = {all_names[nums[1 + 2 * k]]: nums[2 + 2 * k] for k in range(nums[0])}
    names = list(idxs)
    assert len(set(idxs.values())) == len(names) > 8 and b"" not in idxs       # every column got a shard of its own; the message field has none
    p3 = vs.Part(path, inflate=libzstd_inflate)
    nshards = next(n for n in range(len(names), 4000) if len({vo.xxh64(x) % n for x in names}) == len(names))
    v2 = str(tmp_path / "v2")
    os.makedirs(v2)
    for f in os.listdir(path):
        if not (f.startswith("bloom.bin") or f.startswith("values.bin")) and f != "column_idxs.bin":
            shutil.copy(os.path.join(path, f), os.path.join(v2, f))
    for i in range(nshards):
        for kind in ("bloom.bin", "values.bin"):
            open(os.path.join(v2, "%s%d" % (kind, i)), "wb").close()
    for name, k in idxs.items():
        for kind in ("bloom.bin", "values.bin"):
            shutil.copy(os.path.join(path, "%s%d" % (kind, k)), os.path.join(v2, "%s%d" % (kind, vo.xxh64(name) % nshards)))
    meta = json.load(open(os.path.join(path, "metadata.json")))
    json.dump(dict(meta, FormatVersion=2, BloomValuesShardsCount=nshards), open(os.path.join(v2, "metadata.json"), "w"))
    p2 = vs.Part(v2, inflate=libzstd_inflate)
    assert p2.header["FormatVersion"] == 2 and p2.header["BloomValuesShardsCount"] == nshards
    fields = all_fields(originals)
    hb2, hb3 = p2.blocks(fields), p3.blocks(fields)
    for i, (sid, b, ts) in enumerate(originals):
        check_block(hb2, i, fields, b)
        check_block(hb3, i, fields, b)
    # format v1: always 8 shards, BloomValuesShardsCount absent from metadata.json (part_header.go:66-73); only the header rule is checked here
    # (more than 8 columns cannot be laid out over 8 hash shards by moving whole files)
    v1 = str(tmp_path / "v1")
    shutil.copytree(v2, v1)
    json.dump({k: v for k, v in dict(meta, FormatVersion=1).items() if k != "BloomValuesShardsCount"}, open(os.path.join(v1, "metadata.json"), "w"))
    p1 = vs.Part(v1, inflate=libzstd_inflate)
    assert p1.header["BloomValuesShardsCount"] == 8
    # format v0 is refused
    v0 = str(tmp_path / "v0")
    shutil.copytree(v2, v0)
    json.dump({k: v for k, v in dict(meta, FormatVersion=0).items() if k != "BloomValuesShardsCount"}, open(os.path.join(v0, "metadata.json"), "w"))
    with pytest.raises(vs.VlscanError, match="version 0 is not supported"):
        vs.Part(v0, inflate=libzstd_inflate)


def test_damaged_parts_are_rejected(tmp_path):
    path, files, originals, header = write_part(tmp_path, "part", seed=9)
    fields = all_fields(originals)

    def variant(name, **changed):
        d = str(tmp_path / name)
        os.makedirs(d)
        for f, data in files.items():
            data = changed.get(f, data)
            if data is not None:
                open(os.path.join(d, f), "wb").write(data)
        return d

    with pytest.raises(vs.VlscanError, match="needs a ctx"):
        vs.Part(path)
    with pytest.raises(vs.VlscanError, match="inflate callback failed"):
        vs.Part(path, inflate=lambda frame, size: b"x")
    for k, f in enumerate(sorted(files)):
        with pytest.raises(vs.VlscanError):
            vs.Part(variant("missing%d" % k, **{f: None}), inflate=libzstd_inflate)
    for k, f in enumerate(("metadata.json", "column_names.bin", "column_idxs.bin", "metaindex.bin", "index.bin")):
        with pytest.raises(vs.VlscanError):
            vs.Part(variant("short%d" % k, **{f: files[f][:-1]}), inflate=libzstd_inflate)
    meta = json.loads(files["metadata.json"])
    for k, (key, delta) in enumerate((("BlocksCount", 1), ("RowsCount", 1), ("BloomValuesShardsCount", -1), ("FormatVersion", 1))):
        with pytest.raises(vs.VlscanError):
            vs.Part(variant("meta%d" % k, **{"metadata.json": json.dumps(dict(meta, **{key: meta[key] + delta})).encode()}), inflate=libzstd_inflate)
    # data files are only looked at when blocks are described
    p = vs.Part(variant("data0", **{"values.bin0": files["values.bin0"][:-1], "columns_header.bin": files["columns_header.bin"][:-1]}), inflate=libzstd_inflate)
    with pytest.raises(vs.VlscanError, match="outside the file"):
        p.blocks(fields)
    # random damage of the per-block headers: an error or a consistent answer, never a crash
    rng = random.Random(4)
    outcomes = {"ok": 0, "open": 0, "describe": 0}
    for k in range(150):
        f = rng.choice(["columns_header_index.bin", "columns_header.bin", "column_idxs.bin", "metadata.json"])
        b = bytearray(files[f])
        for _ in range(rng.randrange(1, 4)):
            b[rng.randrange(len(b))] = rng.getrandbits(8)
        d = variant("rnd%d" % k, **{f: bytes(b)})
        try:
            p = vs.Part(d, inflate=libzstd_inflate)
        except vs.VlscanError:
            outcomes["open"] += 1
            shutil.rmtree(d)
            continue
        try:
            hb = p.blocks(fields)
            for i in range(hb.nblocks):
                for fld in fields:
                    hb.column(i, fld)
            outcomes["ok"] += 1
        except vs.VlscanError:
            outcomes["describe"] += 1
        del p
        shutil.rmtree(d)
    assert outcomes["describe"] > 20 and sum(outcomes.values()) == 150
