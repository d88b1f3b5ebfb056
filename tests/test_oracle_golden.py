"""CPU suite: pin the oracle (oracle/) against the reference's own known-answer vectors and test tables.

Sources (all under /root/reference/lib/logstorage, transcribed by tests/golden/extract_go_fixtures.py):
bloomfilter_test.go, hash_tokenizer_test.go, tokenizer_test.go, filter_{phrase,prefix,exact,in,regexp,not,and,or}_test.go,
values_encoder_test.go, encoding_test.go.
"""
import math
import struct

import numpy as np
import pytest

from golden_util import load_filter_cases, load_tables, unhex, build_filter, and_or_cases

TABLES = load_tables()
CASES = load_filter_cases()


def test_xxh64_kat(oracle):
    # hash_tokenizer_test.go:19
    assert oracle.xxh64(b"foo") == 0x33BF00A859C4BA3F
    import xxhash   # independent implementation present in the image
    rng = np.random.default_rng(1)
    for n in list(range(0, 70)) + [100, 255, 1000]:
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.xxh64(data) == xxhash.xxh64_intdigest(data)


def test_bloom_marshal_tokens(oracle):
    for args in TABLES["TestBloomFilterMarshalTokens"]:
        tokens, expected = unhex(args[0]) or [], unhex(args[1])
        assert oracle.bloom_marshal_tokens(tokens) == expected
    assert oracle.bloom_marshal_tokens([b"foo"]).hex() == "0000008240180004"


def test_bloom_contains_and_false_positives(oracle):
    # bloomfilter_test.go:23-87 (TestBloomFilter, TestBloomFilterFalsePositive: p <= 0.0011 at 20k tokens)
    tokens = [b"token_%d" % i for i in range(20000)]
    data = oracle.bloom_marshal_tokens(tokens)
    assert len(data) == (20000 * 16 + 63) // 64 * 8
    for i in range(0, 20000, 37):
        assert oracle.bloom_contains_all(data, [tokens[i]])
    fp = sum(oracle.bloom_contains_all(data, [b"non-existing-token_%d" % i]) for i in range(20000))
    assert fp / 20000 <= 0.0011
    assert oracle.bloom_contains_all(b"", [b"anything"])   # unitialized bloom filter matches everything


def test_tokenize_strings(oracle):
    for a, exp in TABLES["TestTokenizeStrings"]:
        assert oracle.tokenize_strings(unhex(a) or []) == (unhex(exp) or [])


def test_tokenize_hashes(oracle):
    for a, exp in TABLES["TestTokenizeHashes"]:
        got = [int(x) for x in oracle.tokenize_hashes(unhex(a) or [])]
        want = [int(x, 16) for x in (unhex(exp) or [])]
        assert got == want


def test_match_phrase_table(oracle):
    for s, phrase, want in TABLES["TestMatchPhrase"]:
        assert oracle.match_phrase(unhex(s), unhex(phrase)) == want, (unhex(s), unhex(phrase))


def test_match_prefix_table(oracle):
    for s, prefix, want in TABLES["TestMatchPrefix"]:
        assert oracle.match_prefix(unhex(s), unhex(prefix)) == want, (unhex(s), unhex(prefix))


def test_match_phrase_invalid_utf8_neighbours(oracle):
    # filter_phrase.go:247-266: a RuneError neighbour counts as a token char (the occurrence is skipped)
    assert not oracle.match_phrase(b"\xfffoo", b"foo")
    assert not oracle.match_phrase(b"foo\xff", b"foo")
    assert oracle.match_phrase(b"\xff foo \xff", b"foo")
    assert not oracle.match_phrase("яfoo".encode(), b"foo")
    assert oracle.match_phrase("«foo»".encode(), b"foo")
    assert not oracle.match_phrase(b"\xef\xbf\xbdfoo", b"foo")   # a *valid* U+FFFD is also RuneError
    assert oracle.match_phrase(b"a.foo", b".foo")               # phrase starting with a non-token char: no leading check


def test_skip_first_last_token(oracle):
    for s, want in TABLES["TestSkipFirstLastToken"]:
        assert oracle.skip_first_last_token(unhex(s)) == unhex(want)


def test_try_parse_tables(oracle):
    for (s,) in TABLES["TestTryParseIPv4String_Success"]:
        v, ok = oracle.try_parse_ipv4(unhex(s))
        assert ok and oracle.encoded_to_string(8, struct.pack(">I", v)) == unhex(s)
    for (s,) in TABLES["TestTryParseIPv4_Failure"]:
        assert not oracle.try_parse_ipv4(unhex(s))[1], unhex(s)
    for (s,) in TABLES["TestTryParseTimestampISO8601String_Success"]:
        v, ok = oracle.try_parse_iso8601(unhex(s))
        assert ok and oracle.encoded_to_string(9, struct.pack(">Q", v & (2**64 - 1))) == unhex(s)
    for (s,) in TABLES["TestTryParseTimestampISO8601_Failure"]:
        assert not oracle.try_parse_iso8601(unhex(s))[1], unhex(s)
    for s, want in TABLES["TestTryParseUint64_Success"]:
        assert oracle.try_parse_uint64(unhex(s)) == (int(unhex(want)), True)
    for (s,) in TABLES["TestTryParseUint64_Failure"]:
        assert not oracle.try_parse_uint64(unhex(s))[1], unhex(s)
    for s, want in TABLES["TestTryParseInt64_Success"]:
        assert oracle.try_parse_int64(unhex(s)) == (int(unhex(want)), True)
    for (s,) in TABLES["TestTryParseInt64_Failure"]:
        assert not oracle.try_parse_int64(unhex(s))[1], unhex(s)
    for s, want in TABLES["TestTryParseFloat64Exact_Success"]:
        v, ok = oracle.try_parse_float64(unhex(s))
        w = float(unhex(want))
        assert ok and abs(v - w) * abs(max(v, w)) < 1e-15   # float64Equal, values_encoder_test.go:556-558
    for (s,) in TABLES["TestTryParseFloat64Exact_Failure"]:
        assert not oracle.try_parse_float64(unhex(s))[1], unhex(s)
    for (s,) in TABLES["TestTryParseFloat64_Failure"]:
        assert not oracle.try_parse_float64(unhex(s))[1], unhex(s)


def test_marshal_number_strings(oracle):
    for i in range(256):   # TestMarshalUint8String
        assert oracle.encoded_to_string(3, bytes([i])) == str(i).encode()
    for f, want in TABLES["TestMarshalFloat64String"]:
        assert oracle.encoded_to_string(7, struct.pack(">d", float(unhex(f)))) == unhex(want)
    assert oracle.encoded_to_string(4, struct.pack(">H", 65535)) == b"65535"
    assert oracle.encoded_to_string(6, struct.pack(">Q", 2**64 - 1)) == b"18446744073709551615"
    assert oracle.encoded_to_string(10, struct.pack(">Q", 1)) == b"-1"     # zig-zag(−1) == 1
    assert oracle.encoded_to_string(7, struct.pack(">d", 1e21)) == b"1000000000000000000000"
    assert oracle.encoded_to_string(7, struct.pack(">d", 1.5e-7)) == b"0.00000015"


def _gofmt_g(x):
    r = repr(float(x))
    return r[:-2] if r.endswith(".0") else r


def test_values_encoder(oracle):
    # values_encoder_test.go:11-98 TestValuesEncoder: type, min, max per encoding + decode round trip
    n = 9   # maxDictLen+1
    def check(values, vt, mn, mx):
        values = [v.encode() for v in values]
        b = oracle.Block.from_columns([("c", values), ("other", [b"x%d" % i for i in range(len(values))])])
        col = [c for c in b.columns if c.name == b"c"]
        if vt is None:
            assert not col
            return
        c = col[0]
        assert (c.value_type, c.min_value, c.max_value) == (vt, mn, mx), (oracle.VT_NAMES[c.value_type], values)
        enc = oracle.unmarshal_strings_block(c.values_block, len(values))
        if vt == 1:
            assert enc == values
        elif vt == 2:
            assert [c.dict[e[0]] for e in enc] == values
        else:
            assert [oracle.encoded_to_string(vt, e) for e in enc] == values
    check(["value_%d" % i for i in range(n)], 1, 0, 0)
    check(["foo", "bar"], 2, 0, 0)
    check(["1", "2foo"], 2, 0, 0)
    check([str(i + 1) for i in range(n)], 3, 1, n)
    check([str((i + 1) << 8) for i in range(n)], 4, 1 << 8, n << 8)
    check([str((i + 1) << 16) for i in range(n)], 5, 1 << 16, n << 16)
    check([str((i + 1) << 32) for i in range(n)], 6, 1 << 32, n << 32)
    check([_gofmt_g(math.sqrt(i + 1)) for i in range(n)], 7, 4607182418800017408, 4613937818241073152)
    check(["1.2.3.%d" % i for i in range(n)], 8, 16909056, 16909064)
    check(["2011-04-19T03:44:01.%03dZ" % i for i in range(n)], 9, 1303184641000000000, 1303184641008000000)
    check([str(i - 4) for i in range(n)], 10, (-4) & (2**64 - 1), 4)
    check(["same"] * n, None, 0, 0)   # const column: not a regular column


def test_bitmap_invariants(oracle):
    # bitmap_test.go:7-133 TestBitmap (sizes 0..99 there; through 200 here to cross three words) + andNot
    assert oracle.lib().vlo_bitmap_selftest(200) == 0


def test_strings_block_codec(oracle):
    # encoding_test.go:17-97 TestMarshalUnmarshalStringsBlock (round trip + exact bytes of the plain single-item case)
    assert oracle.marshal_strings_block([b"foo"]).hex() == "000200030003666f6f"
    rng = np.random.default_rng(7)
    cases = [[], [b""], [b"", b""], [b"foo", b"bar"], [b"x" * 300] * 5, [b"a" * 70000, b"b"],
             [bytes(rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8)) for _ in range(1000)],
             [b"log line %d with some text" % i for i in range(5000)]]
    for a in cases:
        data = oracle.marshal_strings_block(a)
        assert oracle.unmarshal_strings_block(data, len(a)) == a
    # const-length + const-value special cases (encoding.go:113-120)
    lens, data = oracle.decode_values_block(oracle.marshal_strings_block([b"abc"] * 10))
    assert lens == bytes([4, 3]) and data == b"abc"
    lens, data = oracle.decode_values_block(oracle.marshal_strings_block([b"abc", b"abd", b"abe"]))
    assert lens == bytes([4, 3]) and data == b"abcabdabe"
    lens, data = oracle.decode_values_block(oracle.marshal_strings_block([b"ab", b"abd"]))
    assert lens == bytes([0, 2, 3])


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_reference_filter_tables(oracle, idx):
    """Every testFilterMatchForColumns(...) call of filter_{phrase,prefix,exact,in,regexp,not}_test.go."""
    c = CASES[idx]
    b = oracle.Block.from_columns(c["columns"])
    f = build_filter(oracle.Filter, c["filter"])
    got = oracle.bitmap_rows(b.search(f), b.rows)
    assert got == c["expected"], (c["src"], c["filter"])


def test_reference_and_or_tables(oracle):
    for q, cols, f, want in and_or_cases(oracle.Filter):
        b = oracle.Block.from_columns(cols)
        assert oracle.bitmap_rows(b.search(f), b.rows) == want, q


def test_filter_and_or_field_tokens(oracle):
    # filter_and_test.go:115-150 / filter_or_test.go:113-178 (getByFieldTokens), via blocks where only the bloom decides.
    F = oracle.Filter
    assert F.phrase("", "foo bar").tokens() == [b"foo", b"bar"]
    assert F.prefix("", "bar foo ").tokens() == [b"bar", b"foo"]
    assert F.prefix("", "bar foo").tokens() == [b"bar"]
    assert F.exact("", "a foo").tokens() == [b"a", b"foo"]
    assert F.regexp("", "foo qwe bar.+").tokens() == [b"qwe"]
    assert F.regexp("", "a.+ foo bar").tokens() == [b"foo"]
    assert F.regexp("", "conn.*refused").tokens() == []


def test_regexp_dotall_and_quirks(oracle):
    rm = oracle.regex_match
    assert rm("conn.*refused", b"x conn\n refused")          # DotNL on this path
    assert not rm(".", b"\n") and rm(".", b"a\n")             # lone-dot suffix loses DotNL (regexutil.go:229)
    assert not rm("foo.", b"foo\n") and rm("foo.", b"foo!")
    assert rm("foo.+bar", b"foo\nbar")
    # substrDotPlus first-occurrence quirk (regex.go:144-148,181-185)
    assert not rm(".+bar.+", b"bar_bar_x")
    assert not rm("foo.+bar.+", b"foobar_foo_xbar_y")
    assert rm("foo.+bar.+", b"foo_bar_")
    # or-values with prefix: HasPrefix at each prefix occurrence
    assert rm("foo(bar|baz)", b"foox foobaz") and not rm("foo(bar|baz)", b"foox fooba")
    assert rm("(?i)FoO", b"xfOox") and rm("(?i)йцу", "ЙЦУ".encode())
    d = oracle.regex_describe("foo(bar|baz)")
    assert d["prefix"] == "foo" and d["orValues"] == "[bar][baz]"
    d = oracle.regex_describe(".*foo.*")
    assert d["prefix"] == "" and d["orValues"] == "[foo]"     # SimplifyRegex returns ("", "foo") -> strings.Contains
    d = oracle.regex_describe(".+foo.+")
    assert d["substrDotPlus"] == "foo"
    with pytest.raises(RuntimeError):
        rm("foo(", b"x")
    with pytest.raises(RuntimeError):
        rm("a**", b"x")


def test_regexp_repeat_limits(oracle):
    """Go's parser rejects {n,m} repeats above 1000 and nested ones that multiply to more than 1000 copies (repeatIsValid,
    regexp/syntax/parse.go); *, + and ? do not count.  The oracle and the product's compiler follow the same rule."""
    from victorialogs_b200 import scan as vs
    cases = {"x{1000}": True, "x{1001}": False, "x{2,1}": False, "(x{50}){50}": False, "(x{2}){500}": True, "(x{2}){501}": False, "((x{10}){10}){10}": True,
             "((x{10}){10}){11}": False, "(?:x{0,1000}){2}": False, "(x{0}){5000}": False, "(x{1001}){0}": False, "(?:x{2}){0}y{1000}": True, "(x*){1000}": True,
             "(x{1000})*": True, "(x{30}|y{41}){25}": False, "(x{2,}){500}": True, "(x{3,}){500}": False}
    for rx, ok in cases.items():
        if ok:
            oracle.regex_match(rx, b"xx")
            vs.Program(vs.Filter.regexp("_msg", rx))
        else:
            with pytest.raises(RuntimeError, match="invalid repeat count"):
                oracle.regex_match(rx, b"xx")
            with pytest.raises(vs.VlscanError, match="invalid repeat count"):
                vs.Program(vs.Filter.regexp("_msg", rx))


def test_regexp_vs_python_re(oracle):
    """Differential check of the oracle's Pike VM against Python's `re` on the declared syntax subset (boolean
    matching is independent of leftmost-first vs leftmost-longest). Subjects are ASCII + valid UTF-8."""
    import re
    pats = [r"a+b", r"^abc", r"abc$", r"^a.c$", r"[a-c]+\d{2,3}x", r"(ab|cd)*e", r"x?y?z", r"\bfoo\b", r"[^a-z]+", r"\w+@\w+\.com",
            r"(?i)hello", r"a{3}", r"a{2,}b", r"(a|b)(c|d)e?", r"\.\*", r"[[:digit:]]+z", r"\s+\S", r"a.*b.*c", r"(?:x+)+y", r"q[^\n]*r", r"\Bfoo", r"^$", r"a|^b|c$"]
    rng = np.random.default_rng(3)
    alphabet = list("abcdefxyz01 .@\nAHELOhelo_q") + ["é", "я"]
    subjects = ["".join(rng.choice(alphabet, int(rng.integers(0, 12)))) for _ in range(300)] + ["", "abc", "foo", "a foo b", "hello@x.com", "aaab"]
    for p in pats:
        pyp = p.replace("[[:digit:]]", "[0-9]")
        cre = re.compile(pyp, re.DOTALL | re.ASCII if "(?i)" not in p else re.DOTALL)
        d = oracle.regex_describe(p)
        for s in subjects:
            want = cre.search(s) is not None
            if p in (r"^a.c$",):   # no prefix/suffix split: plain semantics
                pass
            assert oracle.regex_match(p, s.encode()) == want, (p, s, d)
