// ORACLE -- TEST INFRASTRUCTURE ONLY (see vlo_util.h header).
//
// Restatement of vendor/github.com/VictoriaMetrics/VictoriaMetrics/lib/regexutil/regex.go:17-212 (Regex, NewRegex,
// MatchString, GetLiterals, matchStringNoPrefix, matchStringWithPrefix) and regexutil.go:67-351 (getOrValues,
// getLiteral, SimplifyRegex, simplifyRegex, getSubstringLiteral, isDotOp) on top of an own RE2-syntax parser and a
// rune-level Pike VM (Go's regexp / regexp/syntax are stdlib and NOT under /root/reference).
//
// PARITY UNPINNED beyond the 24 end-to-end cases of lib/logstorage/filter_regexp_test.go and TestSkipFirstLastToken:
// vm/regexutil ships no tests in vendor/.  Declared scope: literals, escapes, classes (incl. Perl \d\w\s and POSIX
// [:name:]), '.', * + ? {m,n} (lazy variants accepted), alternation, groups (capturing / non-capturing / named),
// ^ $ \A \z \b \B, flags i s m U.  \p{..} classes, \C and back-references are rejected (error) rather than guessed.
// Known modelling decisions (documented in DESIGN.md):
//   * matching is evaluated on the simplified AST; Go's String()/re-Parse fixed point is modelled structurally
//     (capture removal, concat/alternate flattening, adjacent-literal merging, empty-match removal);
//   * the `strings.ReplaceAll(s, "(?s:.)", ".")` quirk of simplifyRegex (regexutil.go:229) is modelled for the only
//     shape where Go >= 1.22's flag printer emits that exact text: a suffix that is a single any-char => it is
//     re-compiled without DotNL, i.e. it does not match '\n';
//   * alternation prefix factoring done by Go's parser is not modelled (affects only or-values extraction and the
//     bloom-token literals, both of which are result-neutral).
#pragma once
#include "vlo_util.h"
#include <memory>

namespace vlo {

struct RegexError : std::runtime_error { using std::runtime_error::runtime_error; };

enum ROp { R_EMPTY, R_LIT, R_CLASS, R_ANY, R_ANYNOTNL, R_BEGIN_TEXT, R_END_TEXT, R_BEGIN_LINE, R_END_LINE, R_WORDB, R_NWORDB,
           R_CAPTURE, R_STAR, R_PLUS, R_QUEST, R_REPEAT, R_CONCAT, R_ALT };

struct RNode;
using RP = std::shared_ptr<RNode>;
struct RNode {
    ROp op;
    std::vector<int32_t> runes;                 // R_LIT: code points ; R_CLASS: sorted [lo,hi] pairs
    bool fold = false;                          // R_LIT: FoldCase flag
    int rmin = 0, rmax = -1;                    // R_REPEAT
    std::vector<RP> sub;
};
inline RP mk(ROp op) { auto n = std::make_shared<RNode>(); n->op = op; return n; }

// ---- simple folding orbits -------------------------------------------------------------------------------------
inline const std::unordered_map<int32_t, std::vector<int32_t>>& fold_orbits() {
    static std::unordered_map<int32_t, std::vector<int32_t>> m = [] {
        std::unordered_map<int32_t, std::vector<int32_t>> r;
        size_t i = 0;
        while (i < VL_FOLD_ORBITS_FLAT_LEN) {
            unsigned n = VL_FOLD_ORBITS[i++];
            std::vector<int32_t> o(VL_FOLD_ORBITS + i, VL_FOLD_ORBITS + i + n);
            for (int32_t c : o) r[c] = o;
            i += n;
        }
        return r;
    }();
    return m;
}

struct RuneSet {   // list of inclusive ranges, normalised on demand
    std::vector<std::pair<int32_t, int32_t>> r;
    void add(int32_t lo, int32_t hi) { if (lo <= hi) r.push_back({lo, hi}); }
    void add_fold(int32_t lo, int32_t hi) {
        add(lo, hi);
        auto& fo = fold_orbits();
        // iterate orbit members (cheap: orbits table is small)
        for (auto& kv : fo) if (kv.first >= lo && kv.first <= hi) for (int32_t c : kv.second) add(c, c);
    }
    void normalize() {
        std::sort(r.begin(), r.end());
        std::vector<std::pair<int32_t, int32_t>> o;
        for (auto& p : r) {
            if (!o.empty() && p.first <= o.back().second + 1) o.back().second = std::max(o.back().second, p.second);
            else o.push_back(p);
        }
        r.swap(o);
    }
    void negate() {
        normalize();
        std::vector<std::pair<int32_t, int32_t>> o;
        int32_t next = 0;
        for (auto& p : r) { if (p.first > next) o.push_back({next, p.first - 1}); next = p.second + 1; }
        if (next <= 0x10FFFF) o.push_back({next, 0x10FFFF});
        r.swap(o);
    }
};

// ---- parser ----------------------------------------------------------------------------------------------------
struct RegexParser {
    sv src; size_t pos = 0;
    bool fI = false, fS = true /* DotNL: regexutil.go:341-343 */, fM = false;
    struct SavedFlags { bool i, s, m; };

    explicit RegexParser(sv s) : src(s) {}
    bool eof() const { return pos >= src.size(); }
    int32_t peek_rune(int* sz) const { return decode_rune((const uint8_t*)src.data() + pos, src.size() - pos, sz); }
    [[noreturn]] void fail(const char* msg) { throw RegexError(std::string("error parsing regexp: ") + msg + ": `" + std::string(src) + "`"); }

    RP parse() {
        RP r = parse_alt();
        if (!eof()) fail("unexpected )");
        return r;
    }
    RP parse_alt() {
        std::vector<RP> alts;
        alts.push_back(parse_concat());
        while (!eof() && src[pos] == '|') { pos++; alts.push_back(parse_concat()); }
        if (alts.size() == 1) return alts[0];
        RP n = mk(R_ALT); n->sub = std::move(alts); return n;
    }
    RP parse_concat() {
        std::vector<RP> items;
        while (!eof() && src[pos] != '|' && src[pos] != ')') {
            RP a = parse_repeat();
            if (a) items.push_back(a);
        }
        if (items.empty()) return mk(R_EMPTY);
        if (items.size() == 1) return items[0];
        RP n = mk(R_CONCAT); n->sub = std::move(items); return n;
    }
    RP parse_repeat() {
        size_t atom_start = pos;
        RP a = parse_atom();
        if (!a) return a;   // flag-only group
        bool repeated = false;
        while (!eof()) {
            char c = src[pos];
            ROp op; int mn = 0, mx = -1;
            if (c == '*') op = R_STAR;
            else if (c == '+') op = R_PLUS;
            else if (c == '?') op = R_QUEST;
            else if (c == '{') {
                size_t save = pos;
                if (!parse_repeat_counts(&mn, &mx)) { pos = save; break; }   // literal '{'
                op = R_REPEAT;
                pos--;   // compensate the pos++ below
            } else break;
            if (repeated) fail("invalid nested repetition operator");
            (void)atom_start;
            pos++;
            if (!eof() && src[pos] == '?') pos++;   // lazy marker: irrelevant for boolean matching
            if (a->op == R_BEGIN_TEXT || a->op == R_END_TEXT || a->op == R_BEGIN_LINE || a->op == R_END_LINE || a->op == R_WORDB || a->op == R_NWORDB) {
                // Go accepts repetition of empty-width ops; keep generic handling
            }
            RP n = mk(op); n->sub.push_back(a); n->rmin = mn; n->rmax = mx;
            if (op == R_REPEAT && (mn > 1000 || mx > 1000)) fail("invalid repeat count");
            if (op == R_REPEAT && (mn >= 2 || mx >= 2) && !repeat_is_valid(n, 1000)) fail("invalid repeat count");   // parser.repeat, regexp/syntax/parse.go
            a = n;
            repeated = true;
        }
        return a;
    }
    // repeatIsValid (regexp/syntax/parse.go): nested {n,m} repeats may not multiply to more than 1000 copies of the innermost expression
    static bool repeat_is_valid(const RP& re, int n) {
        if (re->op == R_REPEAT) {
            int m = re->rmax;
            if (m == 0) return true;
            if (m < 0) m = re->rmin;
            if (m > n) return false;
            if (m > 0) n /= m;
        }
        for (auto& s : re->sub) if (!repeat_is_valid(s, n)) return false;
        return true;
    }
    bool parse_repeat_counts(int* mn, int* mx) {   // at '{'; on success pos is after '}'
        size_t p = pos + 1;
        auto num = [&](int* out) {
            if (p >= src.size() || src[p] < '0' || src[p] > '9') return false;
            long v = 0;
            while (p < src.size() && src[p] >= '0' && src[p] <= '9') { v = v * 10 + (src[p] - '0'); if (v > 100000) v = 100000; p++; }
            *out = (int)v; return true;
        };
        if (!num(mn)) return false;
        if (p < src.size() && src[p] == ',') {
            p++;
            if (p < src.size() && src[p] == '}') *mx = -1;
            else if (!num(mx)) return false;
        } else *mx = *mn;
        if (p >= src.size() || src[p] != '}') return false;
        if (*mx >= 0 && *mn > *mx) fail("invalid repeat count");
        pos = p + 1;
        return true;
    }
    RP lit(int32_t r) { RP n = mk(R_LIT); n->runes.push_back(r); n->fold = fI; /* Go keeps FoldCase on every literal parsed under (?i) */ return n; }

    RP parse_atom() {
        char c = src[pos];
        switch (c) {
        case '(': return parse_group();
        case '[': return parse_class();
        case '.': pos++; return mk(fS ? R_ANY : R_ANYNOTNL);
        case '^': pos++; return mk(fM ? R_BEGIN_LINE : R_BEGIN_TEXT);
        case '$': pos++; return mk(fM ? R_END_LINE : R_END_TEXT);
        case '*': case '+': case '?': fail("missing argument to repetition operator");
        case '\\': return parse_escape();
        default: {
            int sz; int32_t r = peek_rune(&sz);
            if (r == RuneError && sz == 1) fail("invalid UTF-8");
            pos += sz;
            return lit(r);
        }
        }
    }
    int depth = 0;
    RP parse_group() {
        if (++depth > 1000) fail("expression nests too deeply");   // ErrNestingDepth (regexp/syntax maxHeight = 1000); here every parenthesis counts
        struct Leave { int& d; ~Leave() { d--; } } leave{depth};
        pos++;   // (
        SavedFlags saved{fI, fS, fM};
        bool capture = true;
        if (pos + 1 < src.size() && src[pos] == '?') {
            // (?P<name>  (?<name>  (?flags)  (?flags:
            if (src.compare(pos, 3, "?P<") == 0 || (src.compare(pos, 2, "?<") == 0)) {
                size_t e = src.find('>', pos);
                if (e == sv::npos) fail("invalid named capture");
                pos = e + 1;
            } else {
                pos++;
                bool neg = false, any = false;
                bool ni = fI, ns = fS, nm = fM;
                for (;;) {
                    if (eof()) fail("missing closing )");
                    char f = src[pos++];
                    if (f == 'i') { ni = !neg; any = true; }
                    else if (f == 's') { ns = !neg; any = true; }
                    else if (f == 'm') { nm = !neg; any = true; }
                    else if (f == 'U') { any = true; }
                    else if (f == '-') { if (neg) fail("invalid or unsupported Perl syntax"); neg = true; any = false; }
                    else if (f == ':') { if (neg && !any) fail("invalid or unsupported Perl syntax"); fI = ni; fS = ns; fM = nm; capture = false; break; }
                    else if (f == ')') { if (neg && !any) fail("invalid or unsupported Perl syntax"); fI = ni; fS = ns; fM = nm; return nullptr; /* flags stay until group end */ }
                    else fail("invalid or unsupported Perl syntax");
                }
            }
        }
        RP inner = parse_alt();
        if (eof() || src[pos] != ')') fail("missing closing )");
        pos++;
        fI = saved.i; fS = saved.s; fM = saved.m;
        if (!capture) return inner;
        RP n = mk(R_CAPTURE); n->sub.push_back(inner); return n;
    }
    static void perl_class(char c, RuneSet& rs, bool* neg) {
        *neg = (c == 'D' || c == 'W' || c == 'S');
        switch (c) {
        case 'd': case 'D': rs.add('0', '9'); break;
        case 'w': case 'W': rs.add('0', '9'); rs.add('A', 'Z'); rs.add('_', '_'); rs.add('a', 'z'); break;
        case 's': case 'S': rs.add('\t', '\n'); rs.add('\f', '\r'); rs.add(' ', ' '); break;
        }
    }
    RP class_node(RuneSet& rs) {
        rs.normalize();
        RP n = mk(R_CLASS);
        for (auto& p : rs.r) { n->runes.push_back(p.first); n->runes.push_back(p.second); }
        return n;
    }
    int32_t parse_escape_rune() {   // after the backslash char has been consumed; handles char escapes only
        if (eof()) fail("trailing backslash at end of expression");
        char c = src[pos];
        switch (c) {
        case 'a': pos++; return 7;
        case 'f': pos++; return '\f';
        case 'n': pos++; return '\n';
        case 'r': pos++; return '\r';
        case 't': pos++; return '\t';
        case 'v': pos++; return '\v';
        case 'x': {
            pos++;
            auto hex = [&](char h) -> int { if (h >= '0' && h <= '9') return h - '0'; if (h >= 'a' && h <= 'f') return h - 'a' + 10; if (h >= 'A' && h <= 'F') return h - 'A' + 10; return -1; };
            if (eof()) fail("invalid escape sequence");
            if (src[pos] == '{') {
                pos++;
                int32_t v = 0; int nd = 0;
                while (!eof() && src[pos] != '}') { int h = hex(src[pos]); if (h < 0) fail("invalid escape sequence"); v = v * 16 + h; if (v > 0x10FFFF) fail("invalid escape sequence"); pos++; nd++; }
                if (eof() || nd == 0) fail("invalid escape sequence");
                pos++;
                return v;
            }
            if (pos + 1 >= src.size()) fail("invalid escape sequence");
            int h1 = hex(src[pos]), h2 = hex(src[pos + 1]);
            if (h1 < 0 || h2 < 0) fail("invalid escape sequence");
            pos += 2;
            return h1 * 16 + h2;
        }
        default:
            if (c >= '0' && c <= '7') {
                // octal: \0, \012 ... (Go: \1-\7 single digit are backreferences => error unless followed by more octal digits)
                if (c != '0' && !(pos + 1 < src.size() && src[pos + 1] >= '0' && src[pos + 1] <= '7')) fail("invalid escape sequence");
                int32_t v = 0; int nd = 0;
                while (!eof() && nd < 3 && src[pos] >= '0' && src[pos] <= '7') { v = v * 8 + (src[pos] - '0'); pos++; nd++; }
                return v;
            }
            if ((unsigned char)c < 0x80 && !is_token_char((uint8_t)c)) { pos++; return c; }   // punctuation escapes
            fail("invalid escape sequence");
        }
    }
    RP parse_escape() {
        pos++;   // backslash
        if (eof()) fail("trailing backslash at end of expression");
        char c = src[pos];
        switch (c) {
        case 'A': pos++; return mk(R_BEGIN_TEXT);
        case 'z': pos++; return mk(R_END_TEXT);
        case 'b': pos++; return mk(R_WORDB);
        case 'B': pos++; return mk(R_NWORDB);
        case 'd': case 'D': case 'w': case 'W': case 's': case 'S': {
            pos++;
            RuneSet rs; bool neg; perl_class(c, rs, &neg);
            if (fI) { RuneSet f; for (auto& p : rs.r) f.add_fold(p.first, p.second); rs = f; }
            if (neg) rs.negate();
            return class_node(rs);
        }
        case 'Q': {
            pos++;
            std::vector<RP> items;
            while (!eof() && src.compare(pos, 2, "\\E") != 0) { int sz; int32_t r = peek_rune(&sz); pos += sz; items.push_back(lit(r)); }
            if (!eof()) pos += 2;
            if (items.empty()) return mk(R_EMPTY);
            if (items.size() == 1) return items[0];
            RP n = mk(R_CONCAT); n->sub = items; return n;
        }
        case 'p': case 'P': case 'C': fail("unsupported escape (outside the declared oracle scope)");
        default: return lit(parse_escape_rune());
        }
    }
    RP parse_class() {
        pos++;   // [
        RuneSet rs;
        bool neg = false;
        if (!eof() && src[pos] == '^') { neg = true; pos++; }
        bool first = true;
        for (;;) {
            if (eof()) fail("missing closing ]");
            if (src[pos] == ']' && !first) { pos++; break; }
            first = false;
            if (src[pos] == '[' && pos + 1 < src.size() && src[pos + 1] == ':') {
                size_t e = src.find(":]", pos + 2);
                if (e == sv::npos) fail("invalid character class range");
                sv name = src.substr(pos + 2, e - pos - 2);
                bool n2 = false;
                if (!name.empty() && name[0] == '^') { n2 = true; name.remove_prefix(1); }
                RuneSet cs;
                if (name == "alnum") { cs.add('0', '9'); cs.add('A', 'Z'); cs.add('a', 'z'); }
                else if (name == "alpha") { cs.add('A', 'Z'); cs.add('a', 'z'); }
                else if (name == "ascii") cs.add(0, 0x7F);
                else if (name == "blank") { cs.add('\t', '\t'); cs.add(' ', ' '); }
                else if (name == "cntrl") { cs.add(0, 0x1F); cs.add(0x7F, 0x7F); }
                else if (name == "digit") cs.add('0', '9');
                else if (name == "graph") cs.add('!', '~');
                else if (name == "lower") cs.add('a', 'z');
                else if (name == "print") cs.add(' ', '~');
                else if (name == "punct") { cs.add('!', '/'); cs.add(':', '@'); cs.add('[', '`'); cs.add('{', '~'); }
                else if (name == "space") { cs.add('\t', '\r'); cs.add(' ', ' '); }
                else if (name == "upper") cs.add('A', 'Z');
                else if (name == "word") { cs.add('0', '9'); cs.add('A', 'Z'); cs.add('a', 'z'); cs.add('_', '_'); }
                else if (name == "xdigit") { cs.add('0', '9'); cs.add('A', 'F'); cs.add('a', 'f'); }
                else fail("invalid character class range");
                if (fI) { RuneSet f; for (auto& p : cs.r) f.add_fold(p.first, p.second); cs = f; }
                if (n2) cs.negate();
                for (auto& p : cs.r) rs.add(p.first, p.second);
                pos = e + 2;
                continue;
            }
            int32_t lo;
            if (src[pos] == '\\') {
                pos++;
                if (eof()) fail("trailing backslash at end of expression");
                char c = src[pos];
                if (c == 'd' || c == 'D' || c == 'w' || c == 'W' || c == 's' || c == 'S') {
                    pos++;
                    RuneSet cs; bool n2; perl_class(c, cs, &n2);
                    if (fI) { RuneSet f; for (auto& p : cs.r) f.add_fold(p.first, p.second); cs = f; }
                    if (n2) cs.negate();
                    for (auto& p : cs.r) rs.add(p.first, p.second);
                    continue;
                }
                if (c == 'p' || c == 'P') fail("unsupported escape (outside the declared oracle scope)");
                lo = parse_escape_rune();
            } else {
                int sz; lo = peek_rune(&sz);
                if (lo == RuneError && sz == 1) fail("invalid UTF-8");
                pos += sz;
            }
            int32_t hi = lo;
            if (pos + 1 < src.size() && src[pos] == '-' && src[pos + 1] != ']') {
                pos++;
                if (src[pos] == '\\') { pos++; hi = parse_escape_rune(); }
                else { int sz; hi = peek_rune(&sz); if (hi == RuneError && sz == 1) fail("invalid UTF-8"); pos += sz; }
                if (hi < lo) fail("invalid character class range");
            }
            if (fI) rs.add_fold(lo, hi); else rs.add(lo, hi);
        }
        if (neg) {
            // Go: with ClassNL unset in Perl mode a negated class DOES match \n (Perl flags include ClassNL? no: Perl = ClassNL|OneLine|PerlX|UnicodeGroups)
            rs.negate();
        }
        return class_node(rs);
    }
};

// ---- AST simplification (structural model of regexutil.go simplifyRegexp fixed point) ---------------------------
inline bool is_literal_node(const RP& n) { return n->op == R_LIT; }

inline RP simplify_ast(RP n) {
    switch (n->op) {
    case R_CAPTURE: return simplify_ast(n->sub[0]);
    case R_STAR: case R_PLUS: case R_QUEST: case R_REPEAT: {
        RP c = mk(n->op); c->rmin = n->rmin; c->rmax = n->rmax; c->sub.push_back(simplify_ast(n->sub[0]));
        if (n->op == R_REPEAT && n->rmin == 1 && n->rmax == 1) return c->sub[0];
        return c;
    }
    case R_ALT: {
        RP c = mk(R_ALT);
        for (auto& s : n->sub) {
            RP t = simplify_ast(s);
            if (t->op == R_ALT) for (auto& u : t->sub) c->sub.push_back(u); else c->sub.push_back(t);
        }
        if (c->sub.size() == 1) return c->sub[0];
        return c;
    }
    case R_CONCAT: {
        std::vector<RP> flat;
        for (auto& s : n->sub) {
            RP t = simplify_ast(s);
            if (t->op == R_EMPTY) continue;
            if (t->op == R_CONCAT) for (auto& u : t->sub) flat.push_back(u); else flat.push_back(t);
        }
        std::vector<RP> merged;
        for (auto& t : flat) {
            if (!merged.empty() && merged.back()->op == R_LIT && t->op == R_LIT && merged.back()->fold == t->fold) {
                RP m = mk(R_LIT); m->fold = t->fold; m->runes = merged.back()->runes;
                m->runes.insert(m->runes.end(), t->runes.begin(), t->runes.end());
                merged.back() = m;
            } else merged.push_back(t);
        }
        if (merged.empty()) return mk(R_EMPTY);
        if (merged.size() == 1) return merged[0];
        RP c = mk(R_CONCAT); c->sub = merged; return c;
    }
    default: return n;
    }
}

inline std::string runes_to_string(const std::vector<int32_t>& r) { std::string s; for (int32_t c : r) append_rune(s, c); return s; }

// getLiteral regexutil.go:141-149 (capture unwrapping happens on the raw tree for GetLiterals)
inline bool get_literal(const RP& n, std::string* out) {
    if (n->op == R_CAPTURE) return get_literal(n->sub[0], out);
    if (n->op == R_LIT && !n->fold) { *out = runes_to_string(n->runes); return true; }
    return false;
}
inline bool is_dot_op(const RP& n, ROp op) { return n->op == op && n->sub[0]->op == R_ANY; }   // regexutil.go:330-335

static const size_t maxOrValues = 100;
// getOrValues regexutil.go:67-139; returns false when "nil"
inline bool get_or_values(const RP& n, std::vector<std::string>& out) {
    switch (n->op) {
    case R_CAPTURE: return get_or_values(n->sub[0], out);
    case R_LIT: { std::string v; if (!get_literal(n, &v)) return false; out.push_back(v); return true; }
    case R_EMPTY: out.push_back(""); return true;
    case R_ALT: {
        std::vector<std::string> a;
        for (auto& s : n->sub) {
            std::vector<std::string> ca;
            if (!get_or_values(s, ca) || ca.empty()) return false;
            a.insert(a.end(), ca.begin(), ca.end());
            if (a.size() > maxOrValues) return false;
        }
        out = a; return true;
    }
    case R_CLASS: {
        std::vector<std::string> a;
        for (size_t i = 0; i + 1 < n->runes.size(); i += 2) {
            for (int32_t c = n->runes[i]; c <= n->runes[i + 1]; c++) {
                std::string s; append_rune(s, c); a.push_back(s);
                if (a.size() > maxOrValues) return false;
            }
        }
        if (a.empty()) return false;
        out = a; return true;
    }
    case R_CONCAT: {
        if (n->sub.empty()) { out.push_back(""); return true; }
        std::vector<std::string> prefixes;
        if (!get_or_values(n->sub[0], prefixes) || prefixes.empty()) return false;
        if (n->sub.size() == 1) { out = prefixes; return true; }
        RP rest = mk(R_CONCAT); rest->sub.assign(n->sub.begin() + 1, n->sub.end());
        std::vector<std::string> suffixes;
        if (!get_or_values(rest, suffixes) || suffixes.empty()) return false;
        if (prefixes.size() * suffixes.size() > maxOrValues) return false;
        for (auto& p : prefixes) for (auto& s : suffixes) out.push_back(p + s);
        return true;
    }
    default: return false;
    }
}

// ---- Pike VM -----------------------------------------------------------------------------------------------------
enum IOp { I_RUNE1, I_CLASS, I_ANY, I_ANYNOTNL, I_SPLIT, I_JMP, I_MATCH, I_BT, I_ET, I_BL, I_EL, I_WB, I_NWB, I_FAIL };
struct Inst { IOp op; int x = 0, y = 0; std::vector<int32_t> cls; };
struct Prog {
    std::vector<Inst> ins;
    int emit(IOp op, int x = 0, int y = 0) { ins.push_back(Inst{op, x, y, {}}); return (int)ins.size() - 1; }
    // compile n; returns entry pc; continues to `next` pc placeholder patched by caller: we use continuation-passing
    void comp(const RP& n);
};
inline void Prog::comp(const RP& n) {
    switch (n->op) {
    case R_EMPTY: break;
    case R_LIT:
        for (int32_t r : n->runes) {
            int i = emit(I_CLASS);
            if (n->fold) { RuneSet rs; rs.add_fold(r, r); rs.normalize(); for (auto& p : rs.r) { ins[i].cls.push_back(p.first); ins[i].cls.push_back(p.second); } }
            else { ins[i].cls.push_back(r); ins[i].cls.push_back(r); }
        }
        break;
    case R_CLASS: { int i = emit(I_CLASS); ins[i].cls = n->runes; if (n->runes.empty()) ins[i].op = I_FAIL; break; }
    case R_ANY: emit(I_ANY); break;
    case R_ANYNOTNL: emit(I_ANYNOTNL); break;
    case R_BEGIN_TEXT: emit(I_BT); break;
    case R_END_TEXT: emit(I_ET); break;
    case R_BEGIN_LINE: emit(I_BL); break;
    case R_END_LINE: emit(I_EL); break;
    case R_WORDB: emit(I_WB); break;
    case R_NWORDB: emit(I_NWB); break;
    case R_CAPTURE: comp(n->sub[0]); break;
    case R_CONCAT: for (auto& s : n->sub) comp(s); break;
    case R_ALT: {
        std::vector<int> jmps;
        for (size_t i = 0; i < n->sub.size(); i++) {
            if (i + 1 < n->sub.size()) {
                int sp = emit(I_SPLIT);
                ins[sp].x = sp + 1;
                comp(n->sub[i]);
                jmps.push_back(emit(I_JMP));
                ins[sp].y = (int)ins.size();
            } else comp(n->sub[i]);
        }
        for (int j : jmps) ins[j].x = (int)ins.size();
        break;
    }
    case R_STAR: { int sp = emit(I_SPLIT); ins[sp].x = sp + 1; comp(n->sub[0]); int j = emit(I_JMP); ins[j].x = sp; ins[sp].y = (int)ins.size(); break; }
    case R_PLUS: { int st = (int)ins.size(); comp(n->sub[0]); int sp = emit(I_SPLIT); ins[sp].x = st; ins[sp].y = sp + 1; break; }
    case R_QUEST: { int sp = emit(I_SPLIT); ins[sp].x = sp + 1; comp(n->sub[0]); ins[sp].y = (int)ins.size(); break; }
    case R_REPEAT: {
        for (int i = 0; i < n->rmin; i++) comp(n->sub[0]);
        if (n->rmax < 0) { RP s = mk(R_STAR); s->sub.push_back(n->sub[0]); comp(s); }
        else for (int i = n->rmin; i < n->rmax; i++) {
            // nested optional: (x(x(x)?)?)?  -- equivalent for boolean matching to x? x? x?
            RP q = mk(R_QUEST); q->sub.push_back(n->sub[0]); comp(q);
        }
        break;
    }
    }
}

inline bool is_word_rune(int32_t r) { return r >= 0 && r < 0x80 && is_token_char((uint8_t)r); }   // regexp/syntax IsWordChar (ASCII only)

struct CompiledRe {
    Prog prog;
    bool anchored_start = false;   // informational only; '^' is an instruction
    // unanchored search semantics of regexp.MatchString
    bool match(sv s) const {
        const uint8_t* p = (const uint8_t*)s.data();
        size_t n = s.size();
        size_t np = prog.ins.size();
        std::vector<int> clist, nlist;
        std::vector<uint32_t> mark(np + 1, 0);
        uint32_t gen = 0;
        size_t pos = 0;
        int32_t prev = -1;   // previous rune, -1 at start of text
        for (;;) {
            int sz = 0; int32_t cur = -1;
            if (pos < n) cur = decode_rune(p + pos, n - pos, &sz);
            // add thread at start pc for unanchored search
            gen++;
            std::vector<int> stack;
            auto add = [&](std::vector<int>& list, int pc0) {
                stack.push_back(pc0);
                while (!stack.empty()) {
                    int pc = stack.back(); stack.pop_back();
                    if (mark[pc] == gen) continue;
                    mark[pc] = gen;
                    if ((size_t)pc == np) { list.push_back(pc); continue; }
                    const Inst& in = prog.ins[pc];
                    switch (in.op) {
                    case I_JMP: stack.push_back(in.x); break;
                    case I_SPLIT: stack.push_back(in.y); stack.push_back(in.x); break;
                    case I_BT: if (pos == 0) stack.push_back(pc + 1); break;
                    case I_ET: if (pos == n) stack.push_back(pc + 1); break;
                    case I_BL: if (pos == 0 || prev == '\n') stack.push_back(pc + 1); break;
                    case I_EL: if (pos == n || cur == '\n') stack.push_back(pc + 1); break;
                    case I_WB: if (is_word_rune(prev) != is_word_rune(cur)) stack.push_back(pc + 1); break;
                    case I_NWB: if (is_word_rune(prev) == is_word_rune(cur)) stack.push_back(pc + 1); break;
                    default: list.push_back(pc); break;
                    }
                }
            };
            // carry over threads from previous step: they were added into clist with the previous gen; to evaluate
            // empty-width ops at the current position we re-expand them here.
            std::vector<int> expanded;
            for (int pc : clist) add(expanded, pc);
            add(expanded, 0);
            for (int pc : expanded) if ((size_t)pc == np) return true;
            if (pos >= n) return false;
            nlist.clear();
            for (int pc : expanded) {
                const Inst& in = prog.ins[pc];
                bool ok = false;
                switch (in.op) {
                case I_ANY: ok = true; break;
                case I_ANYNOTNL: ok = cur != '\n'; break;
                case I_CLASS:
                    for (size_t i = 0; i + 1 < in.cls.size(); i += 2) if (cur >= in.cls[i] && cur <= in.cls[i + 1]) { ok = true; break; }
                    break;
                default: break;
                }
                if (ok) nlist.push_back(pc + 1);
            }
            clist.swap(nlist);
            prev = cur;
            pos += sz;
        }
    }
};

inline CompiledRe compile_re(const RP& ast) { CompiledRe c; c.prog.comp(ast); return c; }

// ---- regexutil.Regex -------------------------------------------------------------------------------------------------
struct Regex {
    std::string exprStr, prefix;
    bool isOnlyPrefix = false, isSuffixDotStar = false, isSuffixDotPlus = false;
    std::string substrDotStar, substrDotPlus;
    std::vector<std::string> orValues;
    CompiledRe suffixRe;        // anchored at the start iff prefix != "" (regex.go:64-69)
    RP rawAst;                  // for GetLiterals

    static bool contains(sv s, sv sub) { return s.find(sub) != sv::npos; }

    // NewRegex regex.go:49-83 + SimplifyRegex regexutil.go:157-185 + simplifyRegex :199-233
    explicit Regex(sv expr) : exprStr(expr) {
        RegexParser ps(expr);
        rawAst = ps.parse();
        RP sre = simplify_ast(rawAst);
        // simplifyRegex: literal => (lit, ""); concat with leading literal => prefix + rest
        RP suffix;
        std::string lit;
        if (sre->op == R_EMPTY) suffix = sre;
        else if (get_literal(sre, &lit)) { prefix = lit; suffix = mk(R_EMPTY); }
        else if (sre->op == R_CONCAT && get_literal(sre->sub[0], &lit)) {
            prefix = lit;
            std::vector<RP> rest(sre->sub.begin() + 1, sre->sub.end());
            if (rest.size() == 1) suffix = rest[0]; else { suffix = mk(R_CONCAT); suffix->sub = rest; }
        } else suffix = sre;
        // the "(?s:.)" -> "." textual replacement quirk: a suffix that is exactly one any-char loses DotNL
        if (suffix->op == R_ANY) suffix = mk(R_ANYNOTNL);
        // SimplifyRegex: drop .* at the start (only when prefix == "") and at the end
        if (is_dot_op(suffix, R_STAR)) suffix = mk(R_EMPTY);
        else if (suffix->op == R_CONCAT) {
            std::vector<RP> subs = suffix->sub;
            if (prefix.empty()) while (!subs.empty() && is_dot_op(subs[0], R_STAR)) subs.erase(subs.begin());
            while (!subs.empty() && is_dot_op(subs.back(), R_STAR)) subs.pop_back();
            if (subs.empty()) suffix = mk(R_EMPTY);
            else if (subs.size() == 1) suffix = subs[0];   // re-parse of the printed single element
            else { suffix = mk(R_CONCAT); suffix->sub = subs; }
        }
        // NewRegex
        std::vector<std::string> ov;
        if (get_or_values(suffix, ov)) orValues = ov;
        isOnlyPrefix = orValues.size() == 1 && orValues[0].empty();
        isSuffixDotStar = is_dot_op(suffix, R_STAR);
        isSuffixDotPlus = is_dot_op(suffix, R_PLUS);
        substrDotStar = substring_literal(suffix, R_STAR);
        substrDotPlus = substring_literal(suffix, R_PLUS);
        RP anchored = suffix;
        if (!prefix.empty()) { anchored = mk(R_CONCAT); anchored->sub.push_back(mk(R_BEGIN_TEXT)); anchored->sub.push_back(suffix); }
        suffixRe = compile_re(anchored);
    }
    static std::string substring_literal(const RP& sre, ROp op) {   // regexutil.go:316-328
        if (sre->op != R_CONCAT || sre->sub.size() != 3) return "";
        if (!is_dot_op(sre->sub[0], op) || !is_dot_op(sre->sub[2], op)) return "";
        std::string v;
        if (!get_literal(sre->sub[1], &v)) return "";
        return v;
    }
    // MatchString regex.go:86-98
    bool match_string(sv s) const {
        if (isOnlyPrefix) { if (prefix.empty()) return true; return contains(s, prefix); }
        if (prefix.empty()) return match_no_prefix(s);
        return match_with_prefix(s);
    }
    bool match_no_prefix(sv s) const {   // regex.go:131-160
        if (isSuffixDotStar) return true;
        if (isSuffixDotPlus) return !s.empty();
        if (!substrDotStar.empty()) return contains(s, substrDotStar);
        if (!substrDotPlus.empty()) { size_t n = s.find(substrDotPlus); return n != sv::npos && n > 0 && n + substrDotPlus.size() < s.size(); }
        if (orValues.empty()) return suffixRe.match(s);
        for (auto& v : orValues) if (contains(s, v)) return true;
        return false;
    }
    bool match_with_prefix(sv s) const {   // regex.go:162-212
        size_t n = s.find(prefix);
        if (n == sv::npos) return false;
        sv sNext = s.substr(n + 1);
        s = s.substr(n + prefix.size());
        if (isSuffixDotStar) return true;
        if (isSuffixDotPlus) return !s.empty();
        if (!substrDotStar.empty()) return contains(s, substrDotStar);
        if (!substrDotPlus.empty()) { size_t m = s.find(substrDotPlus); return m != sv::npos && m > 0 && m + substrDotPlus.size() < s.size(); }
        for (;;) {
            if (orValues.empty()) { if (suffixRe.match(s)) return true; }
            else for (auto& v : orValues) if (s.substr(0, v.size()) == v) return true;
            s = sNext;
            n = s.find(prefix);
            if (n == sv::npos) return false;
            sNext = s.substr(n + 1);
            s = s.substr(n + prefix.size());
        }
    }
    // GetLiterals regex.go:101-124 (raw parse tree; captures at the top are unwrapped)
    std::vector<std::string> get_literals() const {
        RP sre = rawAst;
        while (sre->op == R_CAPTURE) sre = sre->sub[0];
        // model Go's parser merging of adjacent literals inside a concat (no merging across groups => conservative)
        std::string v;
        if (get_literal(sre, &v)) return {v};
        if (sre->op != R_CONCAT) return {};
        std::vector<std::string> a;
        std::string run; bool inrun = false;
        for (auto& sub : sre->sub) {
            if (sub->op == R_LIT && !sub->fold) { run += runes_to_string(sub->runes); inrun = true; continue; }
            if (inrun) { a.push_back(run); run.clear(); inrun = false; }
            if (sub->op == R_LIT) continue;   // fold-case literal: not a literal for getLiteral
            if (get_literal(sub, &v)) a.push_back(v);
        }
        if (inrun) a.push_back(run);
        return a;
    }
};

// skipFirstLastToken lib/logstorage/filter_regexp.go:53-69
inline std::string skip_first_last_token(sv s) {
    for (;;) { int sz; int32_t r = decode_rune((const uint8_t*)s.data(), s.size(), &sz); if (!is_token_rune(r)) break; s.remove_prefix(sz); }
    for (;;) { int sz; int32_t r = decode_last_rune((const uint8_t*)s.data(), s.size(), &sz); if (!is_token_rune(r)) break; s.remove_suffix(sz); }
    return std::string(s);
}

}  // namespace vlo
