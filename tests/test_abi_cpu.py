"""CPU suite for the product's host side: the C-ABI library loads and exports every symbol include/vlscan.h declares,
the host-side program compiler derives the same tokens as the reference filters, and computing calls fail loudly
without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from golden_util import load_filter_cases, build_filter, and_or_cases
from victorialogs_b200 import scan as vs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "vlscan.h")).read()
    declared = sorted(set(re.findall(r"\b(vlscan_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    L = vs.lib()
    for name in declared:
        assert hasattr(L, name), "libvlscan.so does not export %s" % name
    assert sorted(declared) == sorted(vs.EXPORTS)


def test_struct_layouts_match_header():
    # sizes asserted against the C layout rules of include/vlscan.h (x86-64 SysV)
    assert C.sizeof(vs.CColumn) == 4 + 4 + 8 + 8 + 8 * 12
    assert C.sizeof(vs.CBlock) == 24
    assert C.sizeof(vs.CStats) == 8 * 14
    assert C.sizeof(vs.GenConfig) == 32


def test_program_tokens_match_oracle(oracle):
    """getTokens() of phrase / prefix / exact / regexp leaves (tokenizeStrings, getTokensSkipLast, skipFirstLastToken)."""
    seen = 0
    for c in load_filter_cases():
        spec = c["filter"]
        if spec["kind"] in ("not", "in"):
            continue
        want = build_filter(oracle.Filter, spec).tokens()
        got = vs.Program(build_filter(vs.Filter, spec)).leaf_tokens(0)
        assert got == want, spec
        seen += 1
    assert seen > 250


def test_program_fields_and_errors():
    p = vs.Program(vs.Filter.and_([vs.Filter.phrase("", "GET"), vs.Filter.prefix("path", "api"), vs.Filter.in_("status", ["500", "502", "503"])]))
    assert p.fields() == [b"_msg", b"path", b"status"]
    for bad in ("foo(", "a**", "[z-a]", r"\pL+", "(?P<n", "x{2,1}"):
        with pytest.raises(vs.VlscanError):
            vs.Program(vs.Filter.regexp("f", bad))
    with pytest.raises(vs.VlscanError):   # malformed tree
        vs.Program(vs.Filter(bytes([vs.F_AND, 3, vs.F_NOOP]), "truncated"))
    with pytest.raises(vs.VlscanError):
        vs.Program(vs.Filter(bytes([42]), "unknown kind"))


def test_and_or_trees_compile():
    for q, cols, f, want in and_or_cases(vs.Filter):
        vs.Program(f)


def test_format_float64_matches_oracle_and_golden(oracle):
    # host build of the per-row float64 -> text routine of the scan kernels (marshalFloat64String, values_encoder.go:1397-1399)
    import json, random, struct
    here = os.path.dirname(os.path.abspath(__file__))
    table = json.load(open(os.path.join(here, "golden", "func_tables.json")))["TestMarshalFloat64String"]
    for f, want in table:   # values_encoder_test.go TestMarshalFloat64String
        bits = struct.unpack(">Q", struct.pack(">d", float(bytes.fromhex(f["hex"]) if "hex" in f else f["num"])))[0]
        assert vs.format_float64(bits) == bytes.fromhex(want["hex"])
    rng = random.Random(20240922)
    cases = [0, 1 << 63, 1, 0x7FEFFFFFFFFFFFFF, 0x7FF0000000000000, 0xFFF0000000000000, 0x7FF8000000000000, 0x0010000000000000, 0x000FFFFFFFFFFFFF]
    cases += [rng.getrandbits(64) for _ in range(20000)]
    cases += [rng.getrandbits(52) for _ in range(2000)]   # subnormals
    cases += [struct.unpack(">Q", struct.pack(">d", rng.randint(-10**9, 10**9) / 10 ** rng.randint(0, 9)))[0] for _ in range(20000)]
    cases += [struct.unpack(">Q", struct.pack(">d", float("%de%d" % (m, e))))[0] for e in range(-330, 310) for m in (1, 5, 9)]
    cases += [(e << 52) | m for e in range(0, 2047, 3) for m in (0, 1, (1 << 52) - 1)]
    for bits in cases:
        x = struct.unpack(">d", struct.pack(">Q", bits))[0]
        if x != x:
            assert vs.format_float64(bits) == b"NaN"
            continue
        assert vs.format_float64(bits) == oracle.encoded_to_string(7, struct.pack(">Q", bits)), hex(bits)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is exercised on CPU-only hosts")
    assert vs.device_count() == 0
    with pytest.raises(vs.VlscanError) as e:
        vs.Ctx(0)
    assert "no CUDA device" in str(e.value)
