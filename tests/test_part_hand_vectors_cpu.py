"""A part directory assembled BY HAND - every byte of every file spelled out from the reference's marshal functions, none of it produced by the
oracle's writer - opened by the product's reader (vlscan_part_*).  The oracle writer and the product reader are both this repository's
code (tests/test_part_reader_cpu.py checks one against the other); this file pins the reader, and through the comparison at the end the
oracle's writer as well, to the format as the reference's Go code defines it:

  metadata.json             partHeader as JSON                                   lib/logstorage/part_header.go:14-40,95-102
  column_names.bin          ZSTD(varuint n, n x (varuint len, bytes))              lib/logstorage/column_names.go:98-104
  column_idxs.bin           varuint n, n x (varuint columnID, varuint shardIdx)    lib/logstorage/column_names.go:38-46
  metaindex.bin             ZSTD(indexBlockHeader ...), 56 bytes each              lib/logstorage/index_block_header.go:77-87,112-120
  index.bin                 ZSTD(blockHeader ...) per index block                  lib/logstorage/index_block_header.go:43-57, block_header.go:69-80
  columns_header_index.bin  varuint n, n x (varuint nameID, varuint offset), varuint m, m x (...)   lib/logstorage/block_header.go:296-333
  columns_header.bin        varuint n, n x columnHeader, varuint m, m x bytes(value)               lib/logstorage/block_header.go:454-483,634-730
  timestamps.bin            encoding.MarshalTimestamps blocks                      lib/logstorage/block.go:674-690
  message_bloom.bin / message_values.bin, bloom.binN / values.binN                lib/logstorage/part.go:146-170,220-226
"""
import ctypes as C
import json
import struct

import numpy as np

from victorialogs_b200 import scan as vs
from test_part_reader_cpu import libzstd_inflate


def varuint(n):
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def zstd(data, level=1):
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_compressBound.restype = C.c_size_t
    z.ZSTD_compress.restype = C.c_size_t
    cap = z.ZSTD_compressBound(C.c_size_t(len(data)))
    out = C.create_string_buffer(cap)
    n = z.ZSTD_compress(out, C.c_size_t(cap), data, C.c_size_t(len(data)), C.c_int(level))
    return out.raw[:n]


def zigzag_varint(v):
    return varuint(((v << 1) ^ (v >> 63)) & 0xFFFFFFFFFFFFFFFF)


def test_hand_assembled_part(tmp_path):
    # ---- two blocks of one stream; columns: _msg (strings, shard files message_*), level (dict), code (uint8), host (const) ----
    # block 0: 3 rows, block 1: 2 rows.  Values blocks: bytesBlock(lens items) ++ bytesBlock(data), both short enough for the plain form
    # (0x00, len, raw: encoding.go:344-349); lens items = type byte 0 (u8 per row) or 4 (one const u8) + items (encoding.go:190-243).
    msg0 = [b"GET /a", b"error x", b"ok"]
    msg0_values = bytes([0, 4, 0, 6, 7, 2]) + bytes([0, 15]) + b"".join(msg0)               # lens {type 0; 6, 7, 2}; data 15 bytes
    msg1 = [b"abc", b"xyz"]
    msg1_values = bytes([0, 2, 4, 3]) + bytes([0, 6]) + b"".join(msg1)                      # two equal lens -> const form {type 4; 3}
    lvl0_values = bytes([0, 2, 4, 1]) + bytes([0, 3]) + bytes([0, 1, 0])                    # dict ids, 1 byte each: const lens {4; 1}
    code1_values = bytes([0, 2, 4, 1]) + bytes([0, 2]) + bytes([7, 200])                    # uint8 values 7, 200
    bloom_msg0 = struct.pack(">QQ", 0x0123456789ABCDEF, 0x1)                                # two bloom words, big endian as stored
    bloom_msg1 = struct.pack(">Q", 0xFFFF0000FFFF0000)
    bloom_code1 = struct.pack(">Q", 0x8000000000000001)
    # files of the message column and of shard 0 / shard 1 (level -> shard 1, code -> shard 0, as column_idxs.bin says below)
    message_values = msg0_values + msg1_values
    message_bloom = bloom_msg0 + bloom_msg1
    values0, bloom0 = code1_values, bloom_code1
    values1, bloom1 = lvl0_values, b""
    # column names in order of first use: id 0 = "" (the message field is stored under the empty name), 1 = "level", 2 = "host", 3 = "code"
    names = [b"", b"level", b"host", b"code"]
    column_names = zstd(varuint(len(names)) + b"".join(varuint(len(n)) + n for n in names))
    column_idxs = varuint(2) + varuint(1) + varuint(1) + varuint(3) + varuint(0)             # level -> shard 1, code -> shard 0
    # ---- columnsHeader of block 0: 2 columns (_msg, level), 1 const (host) ----
    ch_msg0 = bytes([1]) + varuint(0) + varuint(len(msg0_values)) + varuint(0) + varuint(len(bloom_msg0))        # string: values off/size, bloom off/size
    ch_lvl0 = bytes([2, 2]) + varuint(4) + b"info" + varuint(5) + b"error" + varuint(0) + varuint(len(lvl0_values))   # dict: n, n x bytes; values off/size
    csh0 = varuint(2) + ch_msg0 + ch_lvl0 + varuint(1) + varuint(6) + b"host-1"
    csh0_idx = varuint(2) + varuint(0) + varuint(1) + varuint(1) + varuint(1 + len(ch_msg0)) + varuint(1) + varuint(2) + varuint(1 + len(ch_msg0) + len(ch_lvl0) + 1)
    # ---- block 1: _msg, code (uint8 with min / max), no consts ----
    ch_msg1 = bytes([1]) + varuint(len(msg0_values)) + varuint(len(msg1_values)) + varuint(len(bloom_msg0)) + varuint(len(bloom_msg1))
    ch_code1 = bytes([3, 7, 200]) + varuint(0) + varuint(len(code1_values)) + varuint(0) + varuint(len(bloom_code1))
    csh1 = varuint(2) + ch_msg1 + ch_code1 + varuint(0)
    csh1_idx = varuint(2) + varuint(0) + varuint(1) + varuint(3) + varuint(1 + len(ch_msg1)) + varuint(0)
    columns_header = csh0 + csh1
    columns_header_index = csh0_idx + csh1_idx
    # ---- timestamps: block 0 = 100, 110, 120 (delta const: one zig-zag varint, marshal type 2); block 1 = 500, 501 + 7 -> nearest delta2 plain (type 5) ----
    ts0, ts1 = [100, 110, 120], [500, 507]
    ts0_block = zigzag_varint(10)
    ts1_block = zigzag_varint(7)                                                            # two items: the first delta only (also delta const by the rules)
    timestamps = ts0_block + ts1_block
    # ---- block headers (index.bin holds one index block with both) ----
    sid = struct.pack(">IIQQ", 1, 2, 0xAABB, 0xCCDD)
    def block_header(unc, rows, ts_off, ts_size, tmin, tmax, mt, chi_off, chi_size, ch_off, ch_size):
        return sid + varuint(unc) + varuint(rows) + struct.pack(">QQqqB", ts_off, ts_size, tmin, tmax, mt) + varuint(chi_off) + varuint(chi_size) + varuint(ch_off) + varuint(ch_size)
    bh0 = block_header(300, 3, 0, len(ts0_block), 100, 120, 2, 0, len(csh0_idx), 0, len(csh0))
    bh1 = block_header(200, 2, len(ts0_block), len(ts1_block), 500, 507, 2, len(csh0_idx), len(csh1_idx), len(csh0), len(csh1))
    index_block = zstd(bh0 + bh1)
    metaindex = zstd(sid + struct.pack(">qqQQ", 100, 507, 0, len(index_block)))
    files = {"column_names.bin": column_names, "column_idxs.bin": column_idxs, "metaindex.bin": metaindex, "index.bin": index_block, "columns_header_index.bin": columns_header_index,
             "columns_header.bin": columns_header, "timestamps.bin": timestamps, "message_values.bin": message_values, "message_bloom.bin": message_bloom,
             "values.bin0": values0, "bloom.bin0": bloom0, "values.bin1": values1, "bloom.bin1": bloom1}
    meta = {"FormatVersion": 3, "CompressedSizeBytes": sum(len(v) for v in files.values()), "UncompressedSizeBytes": 500, "RowsCount": 5, "BlocksCount": 2, "MinTimestamp": 100, "MaxTimestamp": 507,
            "BloomValuesShardsCount": 2}
    d = tmp_path / "hand_part"
    d.mkdir()
    for name, data in files.items():
        (d / name).write_bytes(data)
    (d / "metadata.json").write_text(json.dumps(meta))

    p = vs.Part(str(d), inflate=libzstd_inflate)
    assert p.nblocks == 2 and p.column_names == names
    assert {k: p.header[k] for k in ("FormatVersion", "RowsCount", "BlocksCount", "MinTimestamp", "MaxTimestamp", "BloomValuesShardsCount")} == \
        {"FormatVersion": 3, "RowsCount": 5, "BlocksCount": 2, "MinTimestamp": 100, "MaxTimestamp": 507, "BloomValuesShardsCount": 2}
    h0, h1 = p.block_header(0), p.block_header(1)
    assert (h0["account_id"], h0["project_id"], h0["id_hi"], h0["id_lo"], h0["rows_count"], h0["min_timestamp"], h0["max_timestamp"], h0["ts_marshal_type"]) == (1, 2, 0xAABB, 0xCCDD, 3, 100, 120, 2)
    assert (h1["uncompressed_size_bytes"], h1["rows_count"], h1["ts_block_offset"], h1["ts_block_size"], h1["columns_header_offset"], h1["columns_header_size"]) == (200, 2, len(ts0_block), len(ts1_block), len(csh0), len(csh1))
    assert p.timestamps(0) == ts0_block and p.timestamps(1) == ts1_block
    hb = p.blocks(["_msg", "level", "host", "code"])
    assert hb.nblocks == 2 and hb.source == [0, 1]
    c = hb.column(0, b"_msg")
    assert c["kind"] == "values" and c["value_type"] == 1 and c["values_block"] == msg0_values and c["bloom"] == bloom_msg0
    c = hb.column(0, b"level")
    assert c["value_type"] == 2 and c["dict"] == [b"info", b"error"] and c["values_block"] == lvl0_values and c["bloom"] == b""
    c = hb.column(0, b"host")
    assert c["kind"] == "const" and c["value"] == b"host-1"
    assert hb.column(0, b"code") is None and hb.column(1, b"level") is None and hb.column(1, b"host") is None
    c = hb.column(1, b"_msg")
    assert c["values_block"] == msg1_values and c["bloom"] == bloom_msg1
    c = hb.column(1, b"code")
    assert c["value_type"] == 3 and (c["min_value"], c["max_value"]) == (7, 200) and c["values_block"] == code1_values and c["bloom"] == bloom_code1
    # time pruning by block header: only block 1 overlaps [200, 600]
    assert p.blocks(["_msg"], min_timestamp=200, max_timestamp=600).source == [1]
    p.close()

    # the oracle's reader understands the same hand-made bytes, and its writer produces the same structural records for the same content
    from oracle import vloracle as vo
    assert vo.unmarshal_block_headers(bh0 + bh1)[1]["columns_header_index_offset"] == len(csh0_idx)
    assert vo.unmarshal_columns_header_index(csh0_idx) == ([(0, 1), (1, 1 + len(ch_msg0))], [(2, 1 + len(ch_msg0) + len(ch_lvl0) + 1)])
    got, used = vo.unmarshal_column_header(ch_lvl0)
    assert used == len(ch_lvl0) and got["dict"] == [b"info", b"error"] and got["values_size"] == len(lvl0_values)
    assert vo.marshal_column_header(dict_values=[b"info", b"error"], value_type=2, values_offset=0, values_size=len(lvl0_values)) == ch_lvl0
    assert vo.marshal_column_header(value_type=3, min_value=7, max_value=200, values_offset=0, values_size=len(code1_values), bloom_filter_offset=0, bloom_filter_size=8) == ch_code1
    assert vo.marshal_block_header(account_id=1, project_id=2, id_hi=0xAABB, id_lo=0xCCDD, uncompressed_size_bytes=300, rows_count=3, ts_block_offset=0, ts_block_size=len(ts0_block), min_timestamp=100,
                                   max_timestamp=120, ts_marshal_type=2, columns_header_index_offset=0, columns_header_index_size=len(csh0_idx), columns_header_offset=0, columns_header_size=len(csh0)) == bh0
    assert list(vo.unmarshal_timestamps(ts0_block, 2, 100, 3)) == ts0 and list(vo.unmarshal_timestamps(ts1_block, 2, 500, 2)) == ts1
    assert vo.marshal_timestamps(ts0)[:2] == (ts0_block, 2) and vo.marshal_timestamps(ts1)[:2] == (ts1_block, 2)
    # the values blocks spelled out above are what the reference's writer makes of those rows
    assert vo.marshal_strings_block(msg0) == msg0_values and vo.marshal_strings_block(msg1) == msg1_values
