"""Data-parallel formulation of the timestamps decode, as a numpy model checked against the oracle's sequential
UnmarshalTimestamps (oracle/vlo_timestamps.h) - the shape the device kernel will have, proven on the CPU before it costs GPU time:

  1. varint boundaries: byte i ends a varint iff its continuation bit is clear (a warp ballot on the device);
  2. every varint is assembled from the <= 10 bytes between two boundaries and zig-zag decoded, independently of the others;
  3. NearestDelta:  values = first + inclusive_scan(deltas)                                   (one prefix sum, mod 2^64)
     NearestDelta2: deltas = inclusive_scan([d1, dd2, dd3, ...]); values = first + inclusive_scan(deltas)   (two prefix sums)
     DeltaConst:    values = first + i * delta;   Const: values = first
  4. the _time filter (filterTime, lib/logstorage/filter_time.go:114-137) is then a per-row compare; since a block's timestamps are sorted
     it is also two binary searches - both are checked here.

Malformed input must be rejected exactly where the sequential decoder rejects it: a varint longer than 10 bytes (or whose 10th byte
exceeds 1), fewer or more varints than rows need, a dangling continuation byte at the end."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import vloracle as vo

MT_ZSTD_ND2, MT_DELTA_CONST, MT_CONST, MT_ZSTD_ND, MT_ND2, MT_ND = 1, 2, 3, 4, 5, 6


def inflate(data):
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
    z.ZSTD_decompress.restype = C.c_size_t
    n = z.ZSTD_getFrameContentSize(data, C.c_size_t(len(data)))
    out = C.create_string_buffer(max(n, 1))
    assert z.ZSTD_decompress(out, C.c_size_t(n), data, C.c_size_t(len(data))) == n
    return out.raw[:n]


def parallel_varints(raw):
    """-> int64 array of the zig-zag decoded varints of `raw`, every one assembled independently (steps 1 and 2); ValueError when malformed"""
    b = np.frombuffer(raw, dtype=np.uint8)
    if len(b) == 0:
        return np.zeros(0, dtype=np.int64)
    ends = np.flatnonzero((b & 0x80) == 0)                     # ballot of "continuation bit clear"
    if len(ends) == 0 or ends[-1] != len(b) - 1:
        raise ValueError("dangling continuation byte")
    starts = np.concatenate(([0], ends[:-1] + 1))
    lens = ends - starts + 1
    if lens.max() > 10:
        raise ValueError("varint longer than 10 bytes")
    u = np.zeros(len(ends), dtype=np.uint64)
    for k in range(10):                                        # lane-local loop over at most 10 bytes
        take = lens > k
        byte = b[np.minimum(starts + k, len(b) - 1)].astype(np.uint64)
        if k == 9 and np.any(take & (byte > 1)):
            raise ValueError("varint overflows 64 bits")
        u |= np.where(take, (byte & np.uint64(0x7F)) << np.uint64(7 * k), np.uint64(0))
    return ((u >> np.uint64(1)) ^ (np.uint64(0) - (u & np.uint64(1)))).astype(np.int64)     # zig-zag


def parallel_decode(data, mt, first, items):
    with np.errstate(over="ignore"):
        first_u = np.array([first], dtype=np.int64).astype(np.uint64)
        if mt in (MT_ZSTD_ND, MT_ZSTD_ND2):
            data, mt = inflate(data), (MT_ND if mt == MT_ZSTD_ND else MT_ND2)
        if mt == MT_CONST:
            if data:
                raise ValueError("unexpected data in const encoding")
            return np.full(items, first, dtype=np.int64)
        v = parallel_varints(data).astype(np.uint64)
        if mt == MT_DELTA_CONST:
            if len(v) != 1:
                raise ValueError("delta const needs exactly one varint")
            return (first_u + np.arange(items, dtype=np.uint64) * v[0]).astype(np.int64)
        if mt == MT_ND:
            if items < 1 or len(v) != items - 1:
                raise ValueError("wrong number of deltas")
            return np.concatenate((first_u, first_u + np.cumsum(v, dtype=np.uint64))).astype(np.int64)
        if mt == MT_ND2:
            if items < 2 or len(v) != items - 1:
                raise ValueError("wrong number of deltas")
            deltas = np.cumsum(v, dtype=np.uint64)             # d1, d1+dd2, d1+dd2+dd3, ...
            return np.concatenate((first_u, first_u + np.cumsum(deltas, dtype=np.uint64))).astype(np.int64)
        raise ValueError("unknown marshal type")


def series(rng, kind, n):
    base = rng.choice([0, 1_700_000_000_000_000_000, -5_000_000_000, (1 << 62)])
    if kind == "const":
        return [base] * n
    if kind == "step":
        d = rng.choice([1, 1000, 123456789])
        return [base + i * d for i in range(n)]
    if kind == "jitter":                                       # a log stream: roughly regular with noise -> NearestDelta2
        t, out = base, []
        for _ in range(n):
            t += max(0, int(rng.gauss(1_000_000, 200_000)))
            out.append(t)
        return out
    if kind == "bursty":                                       # many equal neighbours and a few big gaps
        t, out = base, []
        for _ in range(n):
            t += rng.choice([0, 0, 0, 1, 7, 10 ** rng.randrange(3, 10)])
            out.append(t)
        return out
    if kind == "gauge":                                        # not sorted: exercises the NearestDelta (gauge) branch of the codec
        return [rng.randrange(-1000, 1000) for _ in range(n)]
    raise AssertionError(kind)


def test_parallel_formulation_equals_the_sequential_decoder():
    rng = random.Random(8)
    seen = set()
    for trial in range(400):
        kind = rng.choice(["const", "step", "jitter", "bursty", "gauge"])
        n = rng.choice([1, 2, 3, 17, 64, 300, 3000])
        ts = series(rng, kind, n)
        data, mt, first = vo.marshal_timestamps(ts)
        seen.add(mt)
        want = vo.unmarshal_timestamps(data, mt, first, n)
        assert list(want) == ts
        got = parallel_decode(data, mt, first, n)
        assert np.array_equal(got, want), (kind, n, mt)
    assert seen == {MT_ZSTD_ND2, MT_DELTA_CONST, MT_CONST, MT_ZSTD_ND, MT_ND2, MT_ND}


def test_wraparound_arithmetic():
    # deltas and sums are taken mod 2^64 (Go int64 arithmetic wraps silently)
    for ts in ([-(1 << 63), (1 << 63) - 1, -(1 << 63), 5], [(1 << 63) - 1, -(1 << 63), (1 << 63) - 1], [0, -(1 << 63), 0, (1 << 62), -(1 << 62)] * 30):
        data, mt, first = vo.marshal_timestamps(ts)
        assert list(parallel_decode(data, mt, first, len(ts))) == list(vo.unmarshal_timestamps(data, mt, first, len(ts))) == ts


def test_malformed_input_is_rejected_where_the_sequential_decoder_rejects_it():
    rng = random.Random(3)
    checked = 0
    for trial in range(300):
        n = rng.choice([2, 5, 40, 400])
        ts = series(rng, rng.choice(["jitter", "bursty", "gauge"]), n)
        data, mt, first = vo.marshal_timestamps(ts)
        if mt in (MT_ZSTD_ND, MT_ZSTD_ND2):
            data, mt = inflate(data), (MT_ND if mt == MT_ZSTD_ND else MT_ND2)
        if mt not in (MT_ND, MT_ND2) or not data:
            continue
        b = bytearray(data)
        k = rng.randrange(6)
        if k == 0:
            del b[-1]
        elif k == 1:
            b.append(rng.choice([0x00, 0x80, 0x7F]))
        elif k == 2:
            b[rng.randrange(len(b))] |= 0x80
        elif k == 3:
            b[rng.randrange(len(b))] &= 0x7F
        elif k == 4:
            pos = rng.randrange(len(b) + 1)
            b[pos:pos] = bytes([0xFF] * rng.choice([9, 10, 11])) + bytes([rng.choice([0x01, 0x02, 0x7F])])
        else:
            n += rng.choice([-1, 1])
        try:
            want = list(vo.unmarshal_timestamps(bytes(b), mt, first, n))
        except RuntimeError:
            want = None
        try:
            got = list(parallel_decode(bytes(b), mt, first, n))
        except ValueError:
            got = None
        assert got == want, (k, mt, n, bytes(b)[:40])
        checked += want is None
    assert checked > 60


def test_time_filter_on_sorted_timestamps():
    rng = random.Random(5)
    for trial in range(200):
        n = rng.choice([1, 2, 64, 65, 500])
        ts = np.array(series(rng, rng.choice(["const", "step", "jitter", "bursty"]), n), dtype=np.int64)
        lo, hi = sorted((int(rng.choice(ts)) + rng.randrange(-2, 3), int(rng.choice(ts)) + rng.randrange(-2, 3)))
        blk = vo.Block.from_columns([("x", [b"v%d" % i for i in range(n)])]).set_timestamps(ts)
        want = vo.bitmap_rows(blk.search(vo.Filter.time(lo, hi)), n)
        rows = np.flatnonzero((ts >= lo) & (ts <= hi))                       # per-row compare
        a, b = np.searchsorted(ts, lo, side="left"), np.searchsorted(ts, hi, side="right")   # two binary searches: rows [a, b)
        assert list(rows) == want == list(range(a, b))
