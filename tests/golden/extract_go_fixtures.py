#!/usr/bin/env python3
"""Transcribe the reference's own table-driven tests into JSON fixtures (run in the build container only).

The Go toolchain is absent, so the reference tests cannot be executed; their TABLES are the specification.  This
script parses the Go test sources under /root/reference/lib/logstorage with a small Go-literal tokenizer and emits

  tests/golden/filter_cases.json   -- every testFilterMatchForColumns(...) call of filter_{phrase,prefix,exact,in,
                                      regexp,not}_test.go: columns, filter spec, expected row indexes
  tests/golden/func_tables.json    -- f(...) tables of TestMatchPhrase, TestMatchPrefix, TestSkipFirstLastToken,
                                      TestTokenizeStrings, TestTokenizeHashes, TestBloomFilterMarshalTokens,
                                      tryParse* tables of values_encoder_test.go, TestValuesEncoder

Strings are stored as hex (Go strings are byte strings).  /root/reference is not needed at test time.
"""
import json
import os
import re
import sys

REF = "/root/reference/lib/logstorage"
OUT = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------- Go tokenizer
class Tok:
    __slots__ = ("kind", "val", "pos")

    def __init__(self, kind, val, pos):
        self.kind, self.val, self.pos = kind, val, pos

    def __repr__(self):
        return "%s:%r" % (self.kind, self.val)


def go_unquote(body):
    out = bytearray()
    i = 0
    while i < len(body):
        c = body[i]
        if c != "\\":
            out += c.encode("utf-8")
            i += 1
            continue
        i += 1
        e = body[i]
        simple = {"n": 10, "t": 9, "r": 13, "\\": 92, '"': 34, "'": 39, "a": 7, "b": 8, "f": 12, "v": 11}
        if e in simple:
            out.append(simple[e])
            i += 1
        elif e == "x":
            out.append(int(body[i + 1:i + 3], 16))
            i += 3
        elif e == "u":
            out += chr(int(body[i + 1:i + 5], 16)).encode("utf-8")
            i += 5
        elif e == "U":
            out += chr(int(body[i + 1:i + 9], 16)).encode("utf-8")
            i += 9
        elif e in "01234567":
            out.append(int(body[i:i + 3], 8))
            i += 3
        else:
            raise ValueError("bad escape \\%s" % e)
    return bytes(out)


def tokenize(src):
    toks = []
    i, n = 0, len(src)
    while i < n:
        c = src[i]
        if c.isspace():
            i += 1
        elif src.startswith("//", i):
            j = src.find("\n", i)
            i = n if j < 0 else j
        elif src.startswith("/*", i):
            i = src.find("*/", i) + 2
        elif c == '"':
            j = i + 1
            while src[j] != '"':
                j += 2 if src[j] == "\\" else 1
            toks.append(Tok("str", go_unquote(src[i + 1:j]), i))
            i = j + 1
        elif c == "`":
            j = src.find("`", i + 1)
            toks.append(Tok("str", src[i + 1:j].encode("utf-8"), i))
            i = j + 1
        elif c == "'":
            j = i + 1
            while src[j] != "'":
                j += 2 if src[j] == "\\" else 1
            toks.append(Tok("char", go_unquote(src[i + 1:j]), i))
            i = j + 1
        elif c.isalpha() or c == "_":
            j = i
            while j < n and (src[j].isalnum() or src[j] == "_"):
                j += 1
            toks.append(Tok("id", src[i:j], i))
            i = j
        elif c.isdigit():
            j = i
            while j < n and (src[j].isalnum() or src[j] in "._"):
                j += 1
            # exponent sign
            if j < n and src[j] in "+-" and src[j - 1] in "eE" and not src[i:j].startswith("0x"):
                j += 1
                while j < n and src[j].isdigit():
                    j += 1
            toks.append(Tok("num", src[i:j], i))
            i = j
        else:
            for op in (":=", "==", "!=", "<=", ">=", "&&", "||", "<<", ">>", "..."):
                if src.startswith(op, i):
                    toks.append(Tok("op", op, i))
                    i += len(op)
                    break
            else:
                toks.append(Tok("op", c, i))
                i += 1
    return toks


class P:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else Tok("eof", None, -1)

    def next(self):
        t = self.peek()
        self.i += 1
        return t

    def accept(self, kind, val=None):
        t = self.peek()
        if t.kind == kind and (val is None or t.val == val):
            self.i += 1
            return t
        return None

    def expect(self, kind, val=None):
        t = self.accept(kind, val)
        if t is None:
            raise ValueError("expected %s %r, got %r at token %d" % (kind, val, self.peek(), self.i))
        return t

    def skip_balanced(self, open_="{", close="}"):
        depth = 0
        while True:
            t = self.next()
            if t.kind == "eof":
                raise ValueError("unbalanced")
            if t.kind == "op" and t.val == open_:
                depth += 1
            elif t.kind == "op" and t.val == close:
                depth -= 1
                if depth == 0:
                    return

    # value := str | num | nil | []T{...} | T{...} | &T{...} | ident(args) | {...}
    def value(self):
        t = self.peek()
        if t.kind == "str":
            self.next()
            return t.val
        if t.kind == "num":
            self.next()
            return ("num", t.val)
        if t.kind == "op" and t.val == "-":
            self.next()
            v = self.value()
            return ("num", "-" + v[1])
        if t.kind == "op" and t.val == "&":
            self.next()
            return self.value()
        if t.kind == "op" and t.val == "[":
            self.next()
            self.expect("op", "]")
            self.expect("id")
            return self.composite()
        if t.kind == "op" and t.val == "{":
            return self.composite()
        if t.kind == "id":
            self.next()
            if t.val == "nil":
                return None
            if t.val in ("true", "false"):
                return t.val == "true"
            if self.peek().kind == "op" and self.peek().val == "{":
                body = self.composite()
                return ("struct", t.val, body)
            if self.peek().kind == "op" and self.peek().val == "(":
                self.next()
                args = []
                while not self.accept("op", ")"):
                    args.append(self.value())
                    self.accept("op", ",")
                return ("call", t.val, args)
            if self.peek().kind == "op" and self.peek().val == ".":
                # qualified identifier like math.MaxUint64
                self.next()
                t2 = self.expect("id")
                return ("ident", t.val + "." + t2.val)
            return ("ident", t.val)
        raise ValueError("unexpected token %r" % t)

    def composite(self):
        self.expect("op", "{")
        items, fields = [], {}
        while not self.accept("op", "}"):
            if self.peek().kind == "id" and self.peek(1).kind == "op" and self.peek(1).val == ":":
                k = self.next().val
                self.next()
                fields[k] = self.value()
            else:
                items.append(self.value())
            self.accept("op", ",")
        return fields if fields else items


def hx(b):
    return b.hex()


# ---------------------------------------------------------------- filter tests
def filter_spec(v):
    """('struct', 'filterPhrase', {...}) -> dict spec"""
    assert v[0] == "struct", v
    kind, f = v[1], v[2]
    field = f.get("fieldName", b"")
    if kind == "filterPhrase":
        return {"kind": "phrase", "field": hx(field), "arg": hx(f.get("phrase", b""))}
    if kind == "filterPrefix":
        return {"kind": "prefix", "field": hx(field), "arg": hx(f.get("prefix", b""))}
    if kind == "filterExact":
        return {"kind": "exact", "field": hx(field), "arg": hx(f.get("value", b""))}
    if kind == "filterRegexp":
        re_ = f["re"]
        assert re_[0] == "call" and re_[1] == "mustCompileRegex"
        return {"kind": "regexp", "field": hx(field), "arg": hx(re_[2][0])}
    if kind == "filterIn":
        vals = []
        if "values" in f:
            vals = f["values"][2].get("values", [])
        return {"kind": "in", "field": hx(field), "values": [hx(x) for x in vals]}
    if kind == "filterNot":
        return {"kind": "not", "f": filter_spec(f["f"])}
    num = lambda k: int(f[k][1], 0) if k in f else 0
    if kind == "filterExactPrefix":
        return {"kind": "exact_prefix", "field": hx(field), "arg": hx(f.get("prefix", b""))}
    if kind == "filterSequence":
        return {"kind": "sequence", "field": hx(field), "values": [hx(x) for x in (f.get("phrases") or [])]}
    if kind in ("filterContainsAll", "filterContainsAny"):
        vals = f["values"][2].get("values", []) if "values" in f else []
        return {"kind": "contains_all" if kind == "filterContainsAll" else "contains_any", "field": hx(field), "values": [hx(x) for x in (vals or [])]}
    if kind == "filterAnyCasePhrase":
        return {"kind": "any_case_phrase", "field": hx(field), "arg": hx(f.get("phrase", b""))}
    if kind == "filterAnyCasePrefix":
        return {"kind": "any_case_prefix", "field": hx(field), "arg": hx(f.get("prefix", b""))}
    if kind == "filterValueType":
        return {"kind": "value_type", "field": hx(field), "arg": hx(f.get("valueType", b""))}
    def fnum(k):   # float literal, `inf`, `-inf`
        v = f.get(k, ("num", "0"))
        return v[1] if v[0] in ("num", "ident") else "0"
    if kind == "filterRange":
        return {"kind": "range", "field": hx(field), "min": fnum("minValue"), "max": fnum("maxValue")}
    if kind == "filterLeField":
        return {"kind": "le_field", "field": hx(field), "arg": hx(f.get("otherFieldName", b"")), "exclude_equal": bool(f.get("excludeEqualValues", False))}
    if kind == "filterEqField":
        return {"kind": "eq_field", "field": hx(field), "arg": hx(f.get("otherFieldName", b""))}
    if kind == "filterLenRange":
        return {"kind": "len_range", "field": hx(field), "min": num("minLen"), "max": num("maxLen")}
    if kind == "filterStringRange":
        return {"kind": "string_range", "field": hx(field), "min": hx(f.get("minValue", b"")), "max": hx(f.get("maxValue", b""))}
    if kind == "filterIPv4Range":
        return {"kind": "ipv4_range", "field": hx(field), "min": num("minValue"), "max": num("maxValue")}
    raise KeyError(kind)


SUPPORTED = ("filterPhrase", "filterPrefix", "filterExact", "filterRegexp", "filterIn", "filterNot",
             "filterExactPrefix", "filterSequence", "filterLenRange", "filterStringRange", "filterIPv4Range", "filterContainsAll", "filterContainsAny",
             "filterAnyCasePhrase", "filterAnyCasePrefix", "filterValueType", "filterEqField", "filterRange", "filterLeField")


def extract_filter_cases(path):
    src = open(path, encoding="utf-8").read()
    toks = tokenize(src)
    p = P(toks)
    cases = []
    scope_stack = []          # names of t.Run scopes
    columns = None
    fvars = {}
    subtest = [os.path.basename(path)]
    depth_marks = []
    depth = 0
    while p.peek().kind != "eof":
        t = p.peek()
        # t.Run("name", func(t *testing.T) {
        if t.kind == "id" and t.val == "t" and p.peek(1).val == "." and p.peek(2).val == "Run":
            p.i += 4
            name = p.expect("str").val
            while not (p.peek().kind == "op" and p.peek().val == "{"):
                p.next()
            p.next()
            depth += 1
            depth_marks.append((depth, name.decode()))
            continue
        if t.kind == "op" and t.val == "{":
            depth += 1
            p.next()
            continue
        if t.kind == "op" and t.val == "}":
            if depth_marks and depth_marks[-1][0] == depth:
                depth_marks.pop()
            depth -= 1
            p.next()
            continue
        if t.kind == "id" and t.val == "columns" and p.peek(1).val == ":=":
            p.i += 2
            v = p.value()
            columns = [(c["name"], c["values"]) for c in v]
            continue
        if t.kind == "id" and p.peek(1).kind == "op" and p.peek(1).val in (":=", "=") and p.peek(2).val == "&" and p.peek(3).kind == "id" and p.peek(3).val in SUPPORTED:
            name = t.val
            p.i += 2
            fvars[name] = p.value()
            continue
        # fi.values.values = []string{...}
        if t.kind == "id" and t.val in fvars and p.peek(1).val == "." and p.peek(2).val == "values" and p.peek(3).val == "." and p.peek(4).val == "values" and p.peek(5).val == "=":
            name = t.val
            p.i += 6
            vals = p.value()
            st = fvars[name]
            st[2]["values"] = ("struct", "inValues", {"values": vals})
            continue
        if t.kind == "id" and t.val == "testFilterMatchForColumns":
            p.i += 1
            p.expect("op", "(")
            p.expect("id", "t")
            p.expect("op", ",")
            p.expect("id", "columns")
            p.expect("op", ",")
            fname = p.expect("id").val
            p.expect("op", ",")
            needed = p.expect("str").val
            p.expect("op", ",")
            exp = p.value()
            p.expect("op", ")")
            if fname not in fvars:
                continue
            rows = [] if exp is None else [int(x[1]) for x in exp]
            cases.append({
                "src": "%s:%s" % (os.path.basename(path), "/".join(n for _, n in depth_marks)),
                "columns": [{"name": hx(n), "values": [hx(x) for x in vals]} for n, vals in columns],
                "filter": filter_spec(fvars[fname]),
                "needed": hx(needed),
                "expected": rows,
            })
            continue
        p.next()
    return cases


# ---------------------------------------------------------------- f(...) tables
def extract_f_calls(path, func_name, fname="f"):
    """Return list of argument lists of `f(...)` calls inside `func <func_name>(`."""
    src = open(path, encoding="utf-8").read()
    m = re.search(r"^func %s\(" % re.escape(func_name), src, re.M)
    assert m, (path, func_name)
    # function body ends at the next line starting with "}\n"
    end = src.find("\n}\n", m.start())
    body = src[m.start():end]
    toks = tokenize(body)
    p = P(toks)
    calls = []
    while p.peek().kind != "eof":
        t = p.peek()
        if t.kind == "id" and t.val == fname and p.peek(1).val == "(" and (p.i == 0 or p.t[p.i - 1].val not in (".", "func")) and p.peek(2).val != "t":
            save = p.i
            try:
                p.i += 2
                args = []
                while not p.accept("op", ")"):
                    args.append(p.value())
                    p.accept("op", ",")
                calls.append(args)
                continue
            except ValueError:
                p.i = save + 1
                continue
        p.next()
    return calls


def jsonable(v):
    if isinstance(v, bytes):
        return {"hex": v.hex()}
    if isinstance(v, tuple):
        if v[0] == "num":
            return {"num": v[1]}
        if v[0] == "ident":
            return {"ident": v[1]}
        if v[0] == "call":
            return {"call": v[1], "args": [jsonable(a) for a in v[2]]}
        if v[0] == "struct":
            return {"struct": v[1], "fields": jsonable(v[2])}
    if isinstance(v, list):
        return [jsonable(x) for x in v]
    if isinstance(v, dict):
        return {k: jsonable(x) for k, x in v.items()}
    return v


def main():
    cases = []
    for name in ("filter_phrase_test.go", "filter_prefix_test.go", "filter_exact_test.go", "filter_in_test.go", "filter_regexp_test.go", "filter_not_test.go"):
        c = extract_filter_cases(os.path.join(REF, name))
        print(name, len(c))
        cases.extend(c)
    json.dump(cases, open(os.path.join(OUT, "filter_cases.json"), "w"), indent=0)

    # filters of SURVEY §8(f) rank 3: the oracle implements them already, the product does not yet (kept in a file of their own so
    # that the GPU parity tests keep iterating over exactly the kinds libvlscan compiles)
    cases = []
    for name in ("filter_exact_prefix_test.go", "filter_sequence_test.go", "filter_len_range_test.go", "filter_string_range_test.go", "filter_ipv4_range_test.go",
                 "filter_contains_all_test.go", "filter_contains_any_test.go", "filter_any_case_phrase_test.go", "filter_any_case_prefix_test.go",
                 "filter_value_type_test.go", "filter_eq_field_test.go", "filter_range_test.go", "filter_le_field_test.go"):
        c = extract_filter_cases(os.path.join(REF, name))
        print(name, len(c))
        cases.extend(c)
    json.dump(cases, open(os.path.join(OUT, "filter_cases_next.json"), "w"), indent=0)

    tables = {}
    spec = [
        ("TestMatchPhrase", "filter_phrase_test.go", "TestMatchPhrase"),
        ("TestMatchPrefix", "filter_prefix_test.go", "TestMatchPrefix"),
        ("TestSkipFirstLastToken", "filter_regexp_test.go", "TestSkipFirstLastToken"),
        ("TestTokenizeStrings", "tokenizer_test.go", "TestTokenizeStrings"),
        ("TestTokenizeHashes", "hash_tokenizer_test.go", "TestTokenizeHashes"),
        ("TestBloomFilterMarshalTokens", "bloomfilter_test.go", "TestBloomFilterMarshalTokens"),
        ("TestTryParseIPv4String_Success", "values_encoder_test.go", "TestTryParseIPv4String_Success"),
        ("TestTryParseIPv4_Failure", "values_encoder_test.go", "TestTryParseIPv4_Failure"),
        ("TestTryParseTimestampISO8601String_Success", "values_encoder_test.go", "TestTryParseTimestampISO8601String_Success"),
        ("TestTryParseTimestampISO8601_Failure", "values_encoder_test.go", "TestTryParseTimestampISO8601_Failure"),
        ("TestTryParseFloat64_Success", "values_encoder_test.go", "TestTryParseFloat64_Success"),
        ("TestTryParseFloat64_Failure", "values_encoder_test.go", "TestTryParseFloat64_Failure"),
        ("TestTryParseFloat64Exact_Failure", "values_encoder_test.go", "TestTryParseFloat64Exact_Failure"),
        ("TestTryParseFloat64Exact_Success", "values_encoder_test.go", "TestTryParseFloat64Exact_Success"),
        ("TestTryParseUint64_Success", "values_encoder_test.go", "TestTryParseUint64_Success"),
        ("TestTryParseUint64_Failure", "values_encoder_test.go", "TestTryParseUint64_Failure"),
        ("TestTryParseInt64_Success", "values_encoder_test.go", "TestTryParseInt64_Success"),
        ("TestTryParseInt64_Failure", "values_encoder_test.go", "TestTryParseInt64_Failure"),
        ("TestMarshalUint8String", "values_encoder_test.go", "TestMarshalUint8String"),
        ("TestMarshalFloat64String", "values_encoder_test.go", "TestMarshalFloat64String"),
    ]
    for key, fn, func in spec:
        try:
            calls = extract_f_calls(os.path.join(REF, fn), func)
            tables[key] = [jsonable(a) for a in calls]
            print(key, len(calls))
        except AssertionError:
            print(key, "NOT FOUND", file=sys.stderr)
    json.dump(tables, open(os.path.join(OUT, "func_tables.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
