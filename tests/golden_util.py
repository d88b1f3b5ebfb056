"""Helpers shared by the oracle (CPU) and product (GPU) parity tests: load the fixtures transcribed from the
reference's Go tests (tests/golden/*.json, made by tests/golden/extract_go_fixtures.py)."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load_filter_cases(name="filter_cases.json"):
    cases = json.load(open(os.path.join(HERE, "golden", name)))
    for c in cases:
        c["columns"] = [(bytes.fromhex(col["name"]), [bytes.fromhex(v) for v in col["values"]]) for col in c["columns"]]
        # generateRowsFromColumns (filter_test.go:248-277) also adds the stream tags as fields
        rows = len(c["columns"][0][1])
        c["columns"] = [(b"job", [b"foobar"] * rows), (b"instance", [b"host1:234"] * rows)] + c["columns"]
    return cases


def load_tables():
    return json.load(open(os.path.join(HERE, "golden", "func_tables.json")))


def unhex(v):
    if v is None:
        return None
    if isinstance(v, list):
        return [unhex(x) for x in v]
    if isinstance(v, dict) and "hex" in v:
        return bytes.fromhex(v["hex"])
    if isinstance(v, dict) and "num" in v:
        return v["num"]
    return v


def build_filter(F, spec):
    """F: a class with phrase/prefix/exact/in_/regexp/not_ static constructors (oracle or product mirror)."""
    k = spec["kind"]
    if k == "not":
        return F.not_(build_filter(F, spec["f"]))
    field = bytes.fromhex(spec["field"])
    if k == "in":
        return F.in_(field, [bytes.fromhex(v) for v in spec["values"]])
    if k == "range":
        return F.range(field, float(spec["min"]), float(spec["max"]))
    if k == "le_field":
        return F.le_field(field, bytes.fromhex(spec["arg"]), spec["exclude_equal"])
    if k in ("contains_all", "contains_any"):
        return getattr(F, k)(field, [bytes.fromhex(v) for v in spec["values"]])
    if k == "sequence":
        return F.sequence(field, [bytes.fromhex(v) for v in spec["values"]])
    if k == "len_range":
        return F.len_range(field, spec["min"], spec["max"])
    if k == "ipv4_range":
        return F.ipv4_range(field, spec["min"], spec["max"])
    if k == "string_range":
        return F.string_range(field, bytes.fromhex(spec["min"]), bytes.fromhex(spec["max"]))
    if k in ("exact_prefix", "any_case_phrase", "any_case_prefix", "value_type", "eq_field"):
        return getattr(F, k)(field, bytes.fromhex(spec["arg"]))
    arg = bytes.fromhex(spec["arg"])
    return {"phrase": F.phrase, "prefix": F.prefix, "exact": F.exact, "regexp": F.regexp}[k](field, arg)


# Hand-transcribed from lib/logstorage/filter_and_test.go:10-77 and filter_or_test.go:10-72 (these tables go through
# ParseQuery; the LogsQL -> filter tree mapping is lib/logstorage/parser.go:1494-1722: `f:w` phrase, `f:w*` prefix,
# `f:=v` exact, `f:~re` regexp, `f:""` empty phrase, `f:*` empty prefix, `!` not).
_AND_VALUES = [b"a foo", b"a foobar", b"aa abc a", b"ca afdf a,foobar baz", b"a fddf foobarbaz", b"", b"a foobar abcdef",
               b"a kjlkjf dfff", "a ТЕСТЙЦУК НГКШ ".encode(), b"a !!,23.(!1)"]
_OR_VALUES = list(_AND_VALUES)
_OR_VALUES[5] = b"a"
AND_COLUMNS = [(b"foo", _AND_VALUES)]
OR_COLUMNS = [(b"foo", _OR_VALUES)]


def and_or_cases(F):
    """-> list of (logsql, columns, filter, expected rows)"""
    ph, pre, ex, rx, AND, OR, NOT = F.phrase, F.prefix, F.exact, F.regexp, F.and_, F.or_, F.not_
    A, O = AND_COLUMNS, OR_COLUMNS
    return [
        # filter_and_test.go:40-77
        ("foo:a AND foo:abc*", A, AND([ph("foo", "a"), pre("foo", "abc")]), [2, 6]),
        ("foo:abc* AND foo:a", A, AND([pre("foo", "abc"), ph("foo", "a")]), [2, 6]),
        ("foo:bc* AND foo:a", A, AND([pre("foo", "bc"), ph("foo", "a")]), []),
        ("foo:abc AND foo:foo*", A, AND([ph("foo", "abc"), pre("foo", "foo")]), []),
        ("foo:foo AND foo:abc*", A, AND([ph("foo", "foo"), pre("foo", "abc")]), []),
        ("foo:abc* AND foo:foo", A, AND([pre("foo", "abc"), ph("foo", "foo")]), []),
        ('foo:"" AND bar:""', A, AND([ph("foo", ""), ph("bar", "")]), [5]),
        ('foo:foo* AND bar:""', A, AND([pre("foo", "foo"), ph("bar", "")]), [0, 1, 3, 4, 6]),
        ('bar:"" AND foo:foo*', A, AND([ph("bar", ""), pre("foo", "foo")]), [0, 1, 3, 4, 6]),
        ("foo:foo* AND bar:*", A, AND([pre("foo", "foo"), pre("bar", "")]), []),
        ("bar:* AND foo:foo*", A, AND([pre("bar", ""), pre("foo", "foo")]), []),
        ('foo:"a foo"* AND (foo:="a foobar" OR boo:bbbbbbb)', A, AND([pre("foo", "a foo"), OR([ex("foo", "a foobar"), ph("boo", "bbbbbbb")])]), [1]),
        ('foo:"a foo"* AND (foo:"abcd foobar" OR foo:foobar)', A, AND([pre("foo", "a foo"), OR([ph("foo", "abcd foobar"), ph("foo", "foobar")])]), [1, 6]),
        ("(foo:foo* OR bar:baz) AND (bar:x OR foo:a)", A, AND([OR([pre("foo", "foo"), ph("bar", "baz")]), OR([ph("bar", "x"), ph("foo", "a")])]), [0, 1, 3, 4, 6]),
        ("(foo:foo* OR bar:baz) AND (bar:x OR foo:xyz)", A, AND([OR([pre("foo", "foo"), ph("bar", "baz")]), OR([ph("bar", "x"), ph("foo", "xyz")])]), []),
        ("(foo:foo* OR bar:baz) AND (bar:* OR foo:xyz)", A, AND([OR([pre("foo", "foo"), ph("bar", "baz")]), OR([pre("bar", ""), ph("foo", "xyz")])]), []),
        ('(foo:foo* OR bar:baz) AND (bar:"" OR foo:xyz)', A, AND([OR([pre("foo", "foo"), ph("bar", "baz")]), OR([ph("bar", ""), ph("foo", "xyz")])]), [0, 1, 3, 4, 6]),
        ("foo:foo* AND !foo:~bar", A, AND([pre("foo", "foo"), NOT(rx("foo", "bar"))]), [0]),
        # filter_or_test.go:40-72
        ("foo:23 OR foo:abc*", O, OR([ph("foo", "23"), pre("foo", "abc")]), [2, 6, 9]),
        ("foo:abc* OR foo:23", O, OR([pre("foo", "abc"), ph("foo", "23")]), [2, 6, 9]),
        ("foo:xabc* OR foo:23", O, OR([pre("foo", "xabc"), ph("foo", "23")]), [9]),
        ("foo:23 OR foo:xabc*", O, OR([ph("foo", "23"), pre("foo", "xabc")]), [9]),
        ("foo:a OR foo:23", O, OR([ph("foo", "a"), ph("foo", "23")]), list(range(10))),
        ("foo:23 OR foo:a", O, OR([ph("foo", "23"), ph("foo", "a")]), list(range(10))),
        ("foo:x23 OR foo:xabc", O, OR([ph("foo", "x23"), ph("foo", "xabc")]), []),
        ("foo:23 OR bar:xabc*", O, OR([ph("foo", "23"), pre("bar", "xabc")]), [9]),
        ("bar:xabc* OR foo:23", O, OR([pre("bar", "xabc"), ph("foo", "23")]), [9]),
        ('(foo:23 AND bar:"") OR (foo:foo AND bar:*)', O, OR([AND([ph("foo", "23"), ph("bar", "")]), AND([ph("foo", "foo"), pre("bar", "")])]), [9]),
        ('(foo:23 AND bar:"") OR (foo:foo AND bar:"")', O, OR([AND([ph("foo", "23"), ph("bar", "")]), AND([ph("foo", "foo"), ph("bar", "")])]), [0, 9]),
        ('(foo:23 AND bar:"") OR (foo:foo AND baz:"")', O, OR([AND([ph("foo", "23"), ph("bar", "")]), AND([ph("foo", "foo"), ph("baz", "")])]), [0, 9]),
        ('(foo:23 AND bar:abc) OR (foo:foo AND bar:"")', O, OR([AND([ph("foo", "23"), ph("bar", "abc")]), AND([ph("foo", "foo"), ph("bar", "")])]), [0]),
        ("(foo:23 AND bar:abc) OR (foo:foo AND bar:*)", O, OR([AND([ph("foo", "23"), ph("bar", "abc")]), AND([ph("foo", "foo"), pre("bar", "")])]), []),
        ("foo:baz or !foo:~foo", O, OR([ph("foo", "baz"), NOT(rx("foo", "foo"))]), [2, 3, 5, 7, 8, 9]),
    ]
