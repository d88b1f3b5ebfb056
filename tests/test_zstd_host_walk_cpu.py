"""Host half of the device ZSTD decoder (ZstdJob in victorialogs_b200/csrc/vl_zstd.cu) through vlscan_zstd_inspect, on the CPU:
the walker reads untrusted bytes (container, frame header, block headers, literals / sequences section headers), so it is checked
against libzstd's view of the same frames and fuzzed in a child process."""
import ctypes as C
import os
import random
import subprocess
import sys
import textwrap

import pytest

from victorialogs_b200 import scan as vs
from test_gpu_zstd import compress, corpus, zlib_

HERE = os.path.dirname(os.path.abspath(__file__))


def varuint(n):
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def container(frame):
    return b"\x01" + varuint(len(frame)) + frame          # marshalBytesTypeZSTD, encoding.go:351-359


def test_walk_agrees_with_libzstd():
    z = zlib_()
    z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
    z.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
    n = 0
    for name, data in corpus().items():
        for level in (1, 3, 19):
            if level == 19 and len(data) > 400_000:
                continue
            frame = compress(data, level)
            info = vs.zstd_inspect(container(frame) + b"tail")
            assert info["consumed"] == 1 + len(varuint(len(frame))) + len(frame), name
            assert info["regenerated"] == z.ZSTD_getFrameContentSize(frame, len(frame)) == len(data), name
            assert info["blocks"] >= max(1, (len(data) + (128 << 10) - 1) // (128 << 10)), name     # Block_Maximum_Size = 128 KiB
            assert info["compressed_blocks"] <= info["blocks"]
            if info["compressed_blocks"] == 0:
                assert info["sequences"] == 0
            n += 1
    assert n > 50
    # plain container (< 128 bytes are stored raw, encoding.go:344-349)
    assert vs.zstd_inspect(b"\x00\x03abcXYZ") == dict(consumed=5, regenerated=3, blocks=1, compressed_blocks=0, sequences=0)
    # frame options: window descriptor, checksum
    text = corpus()["text_130k"]
    for params in (dict(window_log=10), dict(checksum=1), dict(window_log=14, checksum=1)):
        frame = compress(text, 3, **params)
        assert vs.zstd_inspect(container(frame))["regenerated"] == len(text)


def test_walk_rejects_malformed_containers():
    frame = compress(corpus()["short_text"], 3)
    good = container(frame)
    assert vs.zstd_inspect(good)["regenerated"] == 700
    bad = [b"", b"\x02abc", b"\x00", b"\x00\x05ab", b"\x01", b"\x01\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff", good[:-1], b"\x01" + varuint(len(frame) + 1) + frame,
           container(frame[:4] + b"\x00" + frame[5:])[:12], container(b"\x28\xb5\x2f\xfd\x00"), container(frame[:-3]), container(frame + b"\x00"),
           container(b"\x00" * 16), container(frame[:4] + bytes([frame[4] | 0x08]) + frame[5:]), container(frame[:4] + bytes([frame[4] | 0x01]) + b"\x07" + frame[5:])]
    for b in bad:
        with pytest.raises(vs.VlscanError):
            vs.zstd_inspect(b)


CHILD = textwrap.dedent('''
    import sys, random
    sys.path.insert(0, %r); sys.path.insert(0, %r)
    from victorialogs_b200 import scan as vs
    from test_gpu_zstd import compress, corpus
    from test_zstd_host_walk_cpu import container
    rng = random.Random(7)
    c = corpus()
    seeds = [container(compress(c[k], lvl)) for k in ("tiny", "short_text", "text_130k", "alphabet4", "period3", "mixed", "zeros_1m") for lvl in (1, 3)]
    ok = bad = 0
    for i in range(8000):
        b = bytearray(rng.choice(seeds))
        for _ in range(rng.randrange(1, 4)):
            k = rng.randrange(4)
            pos = rng.randrange(min(len(b), 64)) if rng.random() < 0.7 else rng.randrange(len(b))   # headers sit in front: bias the damage there
            if k == 0: b[pos] = rng.getrandbits(8)
            elif k == 1: del b[pos:]
            elif k == 2: b[pos] ^= 1 << rng.randrange(8)
            else: b.insert(pos, rng.getrandbits(8))
            if not b: b = bytearray(b"\\x01")
        try:
            info = vs.zstd_inspect(bytes(b)); ok += 1
            assert info["consumed"] <= len(b) and info["regenerated"] <= 1 << 30
        except vs.VlscanError:
            bad += 1
    print("accepted", ok, "rejected", bad)
    assert bad > 1000
''')


def test_mutated_frames_never_crash_the_walker():
    r = subprocess.run([sys.executable, "-c", CHILD % (HERE, os.path.dirname(HERE))], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-2000:])
    assert "accepted" in r.stdout
