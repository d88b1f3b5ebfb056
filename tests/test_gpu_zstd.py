"""Device ZSTD decoder (victorialogs_b200/csrc/vl_zstd.cuh) against libzstd: the reference regenerates every values block with
libzstd (unmarshalBytesBlock, lib/logstorage/encoding.go:372-426 -> lib/encoding/compress.go:24-32); here the frames are decoded in
HBM.  Frames are produced by libzstd.so.1 (the library the reference links) at several levels and with the frame options that change
the format features in play: raw / RLE / compressed blocks, raw / RLE / Huffman / treeless literals (1 and 4 streams, direct and
FSE-compressed weights), predefined / RLE / FSE / repeat sequence tables, repeat offsets, overlapping matches, multi-block frames,
window descriptors, content checksums.  Bar: byte-exact."""
import ctypes as C
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

Z = None


def zlib_():
    global Z
    if Z is None:
        Z = C.CDLL("libzstd.so.1")
        Z.ZSTD_compressBound.restype = C.c_size_t
        Z.ZSTD_compressBound.argtypes = [C.c_size_t]
        Z.ZSTD_compress.restype = C.c_size_t
        Z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        Z.ZSTD_decompress.restype = C.c_size_t
        Z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        Z.ZSTD_createCCtx.restype = C.c_void_p
        Z.ZSTD_freeCCtx.argtypes = [C.c_void_p]
        Z.ZSTD_CCtx_setParameter.restype = C.c_size_t
        Z.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
        Z.ZSTD_compress2.restype = C.c_size_t
        Z.ZSTD_compress2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        Z.ZSTD_isError.restype = C.c_uint
        Z.ZSTD_isError.argtypes = [C.c_size_t]
    return Z


def compress(data, level=3, **params):
    z = zlib_()
    cap = z.ZSTD_compressBound(len(data))
    dst = C.create_string_buffer(max(cap, 64))
    if not params:
        n = z.ZSTD_compress(dst, cap, data, len(data), level)
    else:
        ids = dict(level=100, window_log=101, checksum=201, content_size=200, min_match=105, strategy=107, target_length=106)
        cctx = z.ZSTD_createCCtx()
        assert not z.ZSTD_isError(z.ZSTD_CCtx_setParameter(cctx, 100, level))
        for k, v in params.items():
            assert not z.ZSTD_isError(z.ZSTD_CCtx_setParameter(cctx, ids[k], v)), k
        n = z.ZSTD_compress2(cctx, dst, cap, data, len(data))
        z.ZSTD_freeCCtx(cctx)
    assert not z.ZSTD_isError(n)
    return dst.raw[:n]


def cpu_decompress(frame, size):
    z = zlib_()
    dst = C.create_string_buffer(max(size, 1))
    n = z.ZSTD_decompress(dst, size, frame, len(frame))
    assert not z.ZSTD_isError(n) and n == size
    return dst.raw[:size]


def log_text(rng, n):
    words = [b"error", b"timeout", b"GET /api/v1/items", b"conn 10.0.0.7 refused", b"message for the stream", b"worker", b"uuid=", b"ip=", b"u64=", b"INFO", b"warn"]
    out = bytearray()
    while len(out) < n:
        out += rng.choice(words) + b" " + (b"%d" % rng.getrandbits(rng.choice([8, 16, 32, 64]))) + rng.choice([b" ", b"; ", b"\n"])
        if rng.random() < 0.1:
            out += bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 40)))
    return bytes(out[:n])


def corpus():
    rng = random.Random(20250922)
    rnd = lambda n: bytes(rng.getrandbits(8) for _ in range(n))
    c = {}
    c["empty"] = b""
    c["one"] = b"x"
    c["tiny"] = b"hello hello hello"
    c["short_text"] = log_text(rng, 700)
    c["zeros_1m"] = bytes(1 << 20)                                  # RLE blocks
    c["random_300k"] = rnd(300_000)                                 # raw blocks
    c["text_1m"] = log_text(rng, 1_000_000)                         # multi block: treeless literals, repeat modes
    c["text_130k"] = log_text(rng, 130_000)
    c["alphabet4"] = bytes(rng.choice(b"acgt") for _ in range(200_000))          # Huffman dominated
    c["alphabet40"] = bytes(rng.choice(range(40, 80)) for _ in range(150_000))
    c["skewed"] = bytes(min(255, int(rng.expovariate(0.08))) for _ in range(180_000))   # long codes: FSE-compressed weights
    c["period1"] = b"a" * 70_000 + b"b" + b"a" * 1000
    c["period3"] = b"abc" * 50_000
    c["period7"] = b"0123456" * 30_000 + rnd(100) + b"0123456" * 1000
    blob = rnd(200_000)
    c["far_copy"] = blob + blob                                     # offsets of 200 KB across blocks
    c["mixed"] = rnd(5000) + bytes(5000) + log_text(rng, 50_000) + rnd(140_000) + bytes(200_000) + log_text(rng, 100_000)
    c["numbers"] = b"".join(b"%d," % (i * 7919 % 100003) for i in range(60_000))
    c["be_u64"] = b"".join(int(1_700_000_000_000_000_000 + i * 1_000_003).to_bytes(8, "big") for i in range(40_000))   # iso8601 column shape
    c["lens_u8"] = bytes(100 + (i * 37) % 29 for i in range(3000))  # lens block shape
    return c


@pytest.fixture(scope="module")
def ctx():
    from victorialogs_b200 import scan as vs
    return vs, vs.Ctx(0)


def test_corpus_every_level_one_call(ctx):
    vs, cx = ctx
    frames, datas, names = [], [], []
    for name, data in corpus().items():
        for level in (-5, 1, 2, 3, 6, 12, 19):
            if level >= 12 and len(data) > 400_000:
                continue
            frames.append(compress(data, level)); datas.append(data); names.append((name, level))
    assert len(frames) > 100
    got = cx.zstd_decompress(frames, [len(d) for d in datas])
    for g, d, nm in zip(got, datas, names):
        assert g == d, nm


def test_frame_options(ctx):
    vs, cx = ctx
    rng = random.Random(3)
    text = log_text(rng, 600_000)
    cases = [
        (text, dict(checksum=1)),                       # Content_Checksum present (skipped)
        (text, dict(window_log=10)),                    # Window_Descriptor, 1 KiB blocks
        (text, dict(window_log=14)),
        (text, dict(window_log=17, checksum=1)),
        (text[:5000], dict(window_log=10)),
        (text, dict(min_match=3, strategy=1)),
        (text, dict(min_match=7, strategy=2, target_length=9)),
        (bytes(300_000), dict(window_log=12)),
    ]
    frames = [compress(d, 3, **p) for d, p in cases]
    got = cx.zstd_decompress(frames, [len(d) for d, _ in cases])
    for g, (d, p) in zip(got, cases):
        assert g == d, p
        assert cpu_decompress(frames[cases.index((d, p))], len(d)) == d


def test_many_small_frames_one_call(ctx):
    """Values blocks of short columns: thousands of frames of a few hundred bytes .. a few KB in one launch group."""
    vs, cx = ctx
    rng = random.Random(11)
    datas = []
    for i in range(3000):
        kind = i % 5
        n = rng.randint(128, 6000)
        if kind == 0: d = log_text(rng, n)
        elif kind == 1: d = bytes(rng.choice(b"0123456789") for _ in range(n))
        elif kind == 2: d = bytes([rng.randint(0, 7)]) * n
        elif kind == 3: d = bytes(rng.getrandbits(8) for _ in range(n))
        else: d = b"".join(b"%d" % rng.choice([200, 404, 500, 502, 503]) for _ in range(n // 3))
        datas.append(d)
    frames = [compress(d, 1 if len(d) <= 512 else 2 if len(d) <= 4096 else 3) for d in datas]   # getCompressLevel, encoding.go:362-370
    got = cx.zstd_decompress(frames, [len(d) for d in datas])
    assert got == datas


def test_malformed_frames_are_rejected_or_harmless(ctx):
    vs, cx = ctx
    rng = random.Random(5)
    data = log_text(rng, 50_000)
    good = compress(data, 3)
    with pytest.raises(vs.VlscanError):
        cx.zstd_decompress([good[:-1]], [len(data)])
    with pytest.raises(vs.VlscanError):
        cx.zstd_decompress([b"\x00" * 20], [10])
    with pytest.raises(vs.VlscanError):
        cx.zstd_decompress([good], [len(data) + 1])
    with pytest.raises(vs.VlscanError):
        cx.zstd_decompress([good + b"\x00"], [len(data)])
    # flipped bytes inside the payload: either detected, or decoded to bytes of the declared size; never a crash or a hang
    detected = 0
    for k in range(60):
        bad = bytearray(good)
        pos = rng.randrange(8, len(bad))
        bad[pos] ^= 1 << rng.randrange(8)
        try:
            out = cx.zstd_decompress([bytes(bad)], [len(data)])
            assert len(out[0]) == len(data)
        except vs.VlscanError as e:
            assert "cannot decompress block" in str(e)
            detected += 1
    assert detected > 0
    # the context stays usable
    assert cx.zstd_decompress([good], [len(data)]) == [data]


def test_ondisk_stage_equals_decoded_stage(ctx):
    """The same blocks handed over in on-disk form (device ZSTD decode) and in decoded form give identical bitmaps and stats."""
    vs, cx = ctx
    cfg = vs.GenConfig(seed=20250718, total_rows=64 * 3000, rows_per_block=3000, hot_block_permille=500, hit_row_permille=60, columns_mask=0xF)
    batch = cx.generate(cfg, 0, 64)
    host = vs.DownloadedBlocks(cx, batch)
    disk = host.compress()
    col = disk.column(3, "_msg")
    assert "values_block" in col and len(col["values_block"]) < 200_000
    for flt in (vs.Filter.and_([vs.Filter.phrase("_msg", "timeout"), vs.Filter.phrase("level", "error")]),
                vs.Filter.regexp("_msg", "conn.*refused"),
                vs.Filter.and_([vs.Filter.phrase("_msg", "GET"), vs.Filter.prefix("path", "api"), vs.Filter.in_("status", ["500", "502", "503"])])):
        prog = vs.Program(flt)
        w1, c1, s1 = cx.scan_batch(prog, host)
        w2, c2, s2 = cx.scan_batch(prog, disk)
        assert np.array_equal(w1, w2) and np.array_equal(c1, c2)
        assert s1.rows_matched == s2.rows_matched and s1.values_bytes == s2.values_bytes
        assert s2.h2d_bytes < s1.h2d_bytes / 2


def test_ondisk_upload_is_independent_of_host_threads(ctx, monkeypatch):
    """The header walk of an upload runs on $VLSCAN_HOST_THREADS threads (0 = block by block on the caller's thread): same tables, same bitmaps."""
    vs, cx = ctx
    nb = 700                                                 # 2800 values blocks: 11 shards at the default 16 threads
    cfg = vs.GenConfig(seed=77, total_rows=nb * 3000, rows_per_block=3000, hot_block_permille=500, hit_row_permille=60, columns_mask=0xF)
    batch = cx.generate(cfg, 0, nb)
    host = vs.DownloadedBlocks(cx, batch)
    disk = host.compress()
    prog = vs.Program(vs.Filter.and_([vs.Filter.phrase("_msg", "timeout"), vs.Filter.phrase("level", "error")]))
    w0, c0, s0 = cx.scan_batch(prog, host)
    digests = set()
    for threads in ("0", "1", "3", "16", None):
        if threads is None:
            monkeypatch.delenv("VLSCAN_HOST_THREADS", raising=False)
        else:
            monkeypatch.setenv("VLSCAN_HOST_THREADS", threads)
            digests.add(vs.zstd_walk_digest(disk, int(threads))["digest"])
        w, c, s = cx.scan_batch(prog, disk)
        assert np.array_equal(w, w0) and np.array_equal(c, c0) and s.rows_matched == s0.rows_matched, threads
    assert len(digests) == 1
    assert int(c0.sum()) > 0
