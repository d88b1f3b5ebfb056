// Host-side regexp front end of the scan engine: RE2-syntax parser -> analysis equivalent to regexutil.Regex
// (vendor/github.com/VictoriaMetrics/VictoriaMetrics/lib/regexutil/regex.go:17-212, regexutil.go:67-351) -> a DFA over
// *rune classes* with delayed empty-width assertions that the CUDA kernels execute (vl_engine.cu: dfa_run()).
//
// The kernels decode UTF-8 exactly like Go (invalid byte => U+FFFD, width 1), map the rune to a class and step the DFA,
// so matching is rune-exact, not byte-approximate.
//
// Supported syntax (anything else => compile error, never a guess): literals, escapes (\n \t \xHH \x{H..} \. ...),
// classes [..] with ranges / negation / \d \w \s / [:posix:], '.', * + ? {m,n} (+ lazy forms), |, groups (capturing,
// non-capturing, named), ^ $ \A \z \b \B, flags i s m U.  Not supported: \p{..}, \C, \Q..\E is supported.
// DotNL is on by default, like regexutil's parseRegexp (regexutil.go:341-343).
#pragma once
#include <stdint.h>
#include <algorithm>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include "vl_hd.cuh"

namespace vl {

struct RxError : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- AST (flat arena) ------------------------------------------------------------------------------------------------
enum RxOp : uint8_t { RX_EMPTY, RX_LITERAL, RX_CLASS, RX_ANY, RX_ANY_NOT_NL, RX_BOT, RX_EOT, RX_BOL, RX_EOL, RX_WORD_B, RX_NOT_WORD_B,
                      RX_GROUP, RX_STAR, RX_PLUS, RX_OPT, RX_COUNTED, RX_SEQ, RX_ALTS };
struct RxNode {
    RxOp op = RX_EMPTY;
    bool fold = false;                 // literal parsed under (?i)
    int lo = 0, hi = -1;               // counted repetition
    std::vector<int32_t> cps;          // literal: code points; class: [lo,hi] pairs (sorted, merged)
    std::vector<int> kids;
};

class RxTree {
public:
    std::vector<RxNode> nodes;
    int add(RxOp op) { nodes.emplace_back(); nodes.back().op = op; return (int)nodes.size() - 1; }
    int add(const RxNode& n) { nodes.push_back(n); return (int)nodes.size() - 1; }
    RxNode& at(int i) { return nodes[i]; }
    const RxNode& at(int i) const { return nodes[i]; }
};

// ---- case folding ------------------------------------------------------------------------------------------------------
inline const std::unordered_map<int32_t, const unsigned*>& rx_orbits() {
    static const std::unordered_map<int32_t, const unsigned*> m = [] {
        std::unordered_map<int32_t, const unsigned*> r;
        for (size_t i = 0; i < VL_FOLD_ORBITS_FLAT_LEN;) {
            unsigned n = VL_FOLD_ORBITS[i];
            for (unsigned k = 0; k < n; k++) r[(int32_t)VL_FOLD_ORBITS[i + 1 + k]] = &VL_FOLD_ORBITS[i];
            i += 1 + n;
        }
        return r;
    }();
    return m;
}

class RangeSet {
public:
    std::vector<std::pair<int32_t, int32_t>> v;
    void put(int32_t a, int32_t b) { if (a <= b) v.emplace_back(a, b); }
    void put_folded(int32_t a, int32_t b) {
        put(a, b);
        for (auto& kv : rx_orbits()) if (kv.first >= a && kv.first <= b) { const unsigned* o = kv.second; for (unsigned k = 1; k <= o[0]; k++) put((int32_t)o[k], (int32_t)o[k]); }
    }
    void canon() {
        std::sort(v.begin(), v.end());
        size_t w = 0;
        for (size_t i = 0; i < v.size(); i++) {
            if (w && v[i].first <= v[w - 1].second + 1) v[w - 1].second = std::max(v[w - 1].second, v[i].second);
            else v[w++] = v[i];
        }
        v.resize(w);
    }
    void invert() {
        canon();
        std::vector<std::pair<int32_t, int32_t>> o; int32_t nx = 0;
        for (auto& p : v) { if (p.first > nx) o.emplace_back(nx, p.first - 1); nx = p.second + 1; }
        if (nx <= 0x10FFFF) o.emplace_back(nx, 0x10FFFF);
        v.swap(o);
    }
    void merge_from(const RangeSet& o) { v.insert(v.end(), o.v.begin(), o.v.end()); }
};

// ---- parser ------------------------------------------------------------------------------------------------------------
class RxParser {
    const std::string& s_; size_t i_ = 0; RxTree& t_;
    struct Mode { bool icase = false, dotnl = true, multiline = false; } m_;
    int depth_ = 0;   // open parentheses: the parser and every later pass recurse once per level

    [[noreturn]] void bad(const std::string& why) const { throw RxError("error parsing regexp: " + why + ": `" + s_ + "`"); }
    bool done() const { return i_ >= s_.size(); }
    char cur() const { return s_[i_]; }
    bool looking_at(const char* lit) const { return s_.compare(i_, strlen(lit), lit) == 0; }
    int32_t take_rune() {
        int w; int32_t r = decode_rune((const uint8_t*)s_.data() + i_, (uint32_t)(s_.size() - i_), &w);
        if (r == kRuneError && w <= 1) bad("invalid UTF-8");
        i_ += w; return r;
    }
    int literal(int32_t cp) { int n = t_.add(RX_LITERAL); t_.at(n).cps.push_back(cp); t_.at(n).fold = m_.icase; return n; }
    int klass(RangeSet& rs) { rs.canon(); int n = t_.add(RX_CLASS); for (auto& p : rs.v) { t_.at(n).cps.push_back(p.first); t_.at(n).cps.push_back(p.second); } return n; }

    static bool shorthand(char c, RangeSet& rs) {   // \d \w \s and negations; returns whether negated
        switch (c | 0x20) {
        case 'd': rs.put('0', '9'); break;
        case 'w': rs.put('0', '9'); rs.put('A', 'Z'); rs.put('_', '_'); rs.put('a', 'z'); break;
        case 's': rs.put('\t', '\n'); rs.put('\f', '\r'); rs.put(' ', ' '); break;
        }
        return c == 'D' || c == 'W' || c == 'S';
    }
    void fold_if_needed(RangeSet& rs) { if (!m_.icase) return; RangeSet f; for (auto& p : rs.v) f.put_folded(p.first, p.second); rs = f; }
    static int hexval(char h) { if (h >= '0' && h <= '9') return h - '0'; if ((h | 0x20) >= 'a' && (h | 0x20) <= 'f') return (h | 0x20) - 'a' + 10; return -1; }

    int32_t escape_cp() {   // i_ is just after the backslash; single-code-point escapes
        if (done()) bad("trailing backslash at end of expression");
        char c = s_[i_];
        switch (c) {
        case 'a': i_++; return 7; case 'f': i_++; return 12; case 'n': i_++; return 10; case 'r': i_++; return 13; case 't': i_++; return 9; case 'v': i_++; return 11;
        case 'x': {
            i_++;
            if (done()) bad("invalid escape sequence");
            if (cur() == '{') {
                i_++; int32_t v = 0; int nd = 0;
                while (!done() && cur() != '}') { int h = hexval(cur()); if (h < 0) bad("invalid escape sequence"); v = v * 16 + h; if (v > 0x10FFFF) bad("invalid escape sequence"); i_++; nd++; }
                if (done() || !nd) bad("invalid escape sequence");
                i_++; return v;
            }
            if (i_ + 1 >= s_.size() || hexval(s_[i_]) < 0 || hexval(s_[i_ + 1]) < 0) bad("invalid escape sequence");
            int32_t v = hexval(s_[i_]) * 16 + hexval(s_[i_ + 1]); i_ += 2; return v;
        }
        default:
            if (c >= '0' && c <= '7') {
                bool more = i_ + 1 < s_.size() && s_[i_ + 1] >= '0' && s_[i_ + 1] <= '7';
                if (c != '0' && !more) bad("invalid escape sequence");   // back-reference
                int32_t v = 0; for (int k = 0; k < 3 && !done() && cur() >= '0' && cur() <= '7'; k++) { v = v * 8 + (cur() - '0'); i_++; }
                return v;
            }
            if ((unsigned char)c < 0x80 && !is_token_char((uint8_t)c)) { i_++; return c; }
            bad("invalid escape sequence");
        }
    }
    int escape_atom() {
        i_++;
        if (done()) bad("trailing backslash at end of expression");
        char c = cur();
        if (c == 'A') { i_++; return t_.add(RX_BOT); }
        if (c == 'z') { i_++; return t_.add(RX_EOT); }
        if (c == 'b') { i_++; return t_.add(RX_WORD_B); }
        if (c == 'B') { i_++; return t_.add(RX_NOT_WORD_B); }
        if (strchr("dDwWsS", c)) { i_++; RangeSet rs; bool neg = shorthand(c, rs); fold_if_needed(rs); if (neg) rs.invert(); return klass(rs); }
        if (c == 'Q') {
            i_++; int seq = t_.add(RX_SEQ);
            while (!done() && !looking_at("\\E")) { int l = literal(take_rune()); t_.at(seq).kids.push_back(l); }
            if (!done()) i_ += 2;
            return seq;
        }
        if (c == 'p' || c == 'P' || c == 'C') bad("unsupported escape (\\p, \\P, \\C are outside the supported syntax)");
        return literal(escape_cp());
    }
    void posix_class(RangeSet& out) {   // at "[:"
        size_t e = s_.find(":]", i_ + 2);
        if (e == std::string::npos) bad("invalid character class range");
        std::string name = s_.substr(i_ + 2, e - i_ - 2);
        bool neg = !name.empty() && name[0] == '^';
        if (neg) name.erase(0, 1);
        static const std::map<std::string, std::vector<std::pair<int, int>>> tbl = {
            {"alnum", {{'0', '9'}, {'A', 'Z'}, {'a', 'z'}}}, {"alpha", {{'A', 'Z'}, {'a', 'z'}}}, {"ascii", {{0, 0x7F}}}, {"blank", {{'\t', '\t'}, {' ', ' '}}},
            {"cntrl", {{0, 0x1F}, {0x7F, 0x7F}}}, {"digit", {{'0', '9'}}}, {"graph", {{'!', '~'}}}, {"lower", {{'a', 'z'}}}, {"print", {{' ', '~'}}},
            {"punct", {{'!', '/'}, {':', '@'}, {'[', '`'}, {'{', '~'}}}, {"space", {{'\t', '\r'}, {' ', ' '}}}, {"upper", {{'A', 'Z'}}},
            {"word", {{'0', '9'}, {'A', 'Z'}, {'a', 'z'}, {'_', '_'}}}, {"xdigit", {{'0', '9'}, {'A', 'F'}, {'a', 'f'}}}};
        auto it = tbl.find(name);
        if (it == tbl.end()) bad("invalid character class range");
        RangeSet rs; for (auto& p : it->second) rs.put(p.first, p.second);
        fold_if_needed(rs);
        if (neg) rs.invert();
        out.merge_from(rs);
        i_ = e + 2;
    }
    int bracket() {
        i_++;
        RangeSet rs; bool neg = false;
        if (!done() && cur() == '^') { neg = true; i_++; }
        for (bool first = true;; first = false) {
            if (done()) bad("missing closing ]");
            if (cur() == ']' && !first) { i_++; break; }
            if (looking_at("[:")) { posix_class(rs); continue; }
            int32_t a;
            if (cur() == '\\') {
                i_++;
                if (done()) bad("trailing backslash at end of expression");
                if (strchr("dDwWsS", cur())) { RangeSet c; bool n2 = shorthand(cur(), c); i_++; fold_if_needed(c); if (n2) c.invert(); rs.merge_from(c); continue; }
                if (cur() == 'p' || cur() == 'P') bad("unsupported escape (\\p, \\P are outside the supported syntax)");
                a = escape_cp();
            } else a = take_rune();
            int32_t b = a;
            if (i_ + 1 < s_.size() && cur() == '-' && s_[i_ + 1] != ']') {
                i_++;
                if (cur() == '\\') { i_++; b = escape_cp(); } else b = take_rune();
                if (b < a) bad("invalid character class range");
            }
            if (m_.icase) rs.put_folded(a, b); else rs.put(a, b);
        }
        if (neg) rs.invert();   // Perl flags include ClassNL: negated classes match '\n'
        return klass(rs);
    }
    int group() {
        // Go (>= 1.19) refuses trees higher than 1000 (ErrNestingDepth, regexp/syntax/parse.go maxHeight); here every parenthesis counts, which also
        // bounds the recursion of this compiler on hostile input
        if (++depth_ > 1000) bad("expression nests too deeply");
        struct Leave { int& d; ~Leave() { d--; } } leave{depth_};
        i_++;
        Mode outer = m_;
        bool capturing = true;
        if (!done() && cur() == '?') {
            if (looking_at("?P<") || (looking_at("?<") && !looking_at("?<=") && !looking_at("?<!"))) {
                size_t e = s_.find('>', i_); if (e == std::string::npos) bad("invalid named capture");
                i_ = e + 1;
            } else {
                i_++;
                Mode nm = m_; bool minus = false, seen = false;
                for (;;) {
                    if (done()) bad("missing closing )");
                    char f = s_[i_++];
                    if (f == 'i') { nm.icase = !minus; seen = true; }
                    else if (f == 's') { nm.dotnl = !minus; seen = true; }
                    else if (f == 'm') { nm.multiline = !minus; seen = true; }
                    else if (f == 'U') seen = true;
                    else if (f == '-') { if (minus) bad("invalid or unsupported Perl syntax"); minus = true; seen = false; }
                    else if (f == ':' || f == ')') {
                        if (minus && !seen) bad("invalid or unsupported Perl syntax");
                        m_ = nm;
                        if (f == ')') return -1;   // flags apply to the rest of the enclosing group
                        capturing = false; break;
                    } else bad("invalid or unsupported Perl syntax");
                }
            }
        }
        int inner = alternation();
        if (done() || cur() != ')') bad("missing closing )");
        i_++;
        m_ = outer;
        if (!capturing) return inner;
        int g = t_.add(RX_GROUP); t_.at(g).kids.push_back(inner); return g;
    }
    int atom() {
        switch (cur()) {
        case '(': return group();
        case '[': return bracket();
        case '.': i_++; return t_.add(m_.dotnl ? RX_ANY : RX_ANY_NOT_NL);
        case '^': i_++; return t_.add(m_.multiline ? RX_BOL : RX_BOT);
        case '$': i_++; return t_.add(m_.multiline ? RX_EOL : RX_EOT);
        case '\\': return escape_atom();
        case '*': case '+': case '?': bad("missing argument to repetition operator");
        default: return literal(take_rune());
        }
    }
    bool counts(int* lo, int* hi) {   // at '{'
        size_t p = i_ + 1;
        auto num = [&](int* o) { if (p >= s_.size() || s_[p] < '0' || s_[p] > '9') return false; long v = 0; while (p < s_.size() && s_[p] >= '0' && s_[p] <= '9') { v = std::min(v * 10 + (s_[p] - '0'), 100000L); p++; } *o = (int)v; return true; };
        if (!num(lo)) return false;
        *hi = *lo;
        if (p < s_.size() && s_[p] == ',') { p++; if (p < s_.size() && s_[p] == '}') *hi = -1; else if (!num(hi)) return false; }
        if (p >= s_.size() || s_[p] != '}') return false;
        if ((*hi >= 0 && *lo > *hi) || *lo > 1000 || *hi > 1000) bad("invalid repeat count");
        i_ = p + 1;
        return true;
    }
    // repeatIsValid (regexp/syntax/parse.go): nested {n,m} repeats may not multiply to more than 1000 copies of the innermost expression
    bool repeat_is_valid(int node, int n) {
        const RxNode& re = t_.at(node);
        if (re.op == RX_COUNTED) {
            int m = re.hi;
            if (m == 0) return true;
            if (m < 0) m = re.lo;
            if (m > n) return false;
            if (m > 0) n /= m;
        }
        const std::vector<int> kids = re.kids;
        for (int k : kids) if (!repeat_is_valid(k, n)) return false;
        return true;
    }
    int repetition() {
        int a = atom();
        if (a < 0) return a;
        for (bool had = false; !done(); had = true) {
            RxOp op; int lo = 0, hi = -1;
            char c = cur();
            if (c == '*') { op = RX_STAR; i_++; }
            else if (c == '+') { op = RX_PLUS; i_++; }
            else if (c == '?') { op = RX_OPT; i_++; }
            else if (c == '{') { if (!counts(&lo, &hi)) break; op = RX_COUNTED; }
            else break;
            if (had) bad("invalid nested repetition operator");
            if (!done() && cur() == '?') i_++;   // non-greedy marker: irrelevant for boolean matching
            int n = t_.add(op); t_.at(n).kids.push_back(a); t_.at(n).lo = lo; t_.at(n).hi = hi;
            if (op == RX_COUNTED && (lo >= 2 || hi >= 2) && !repeat_is_valid(n, 1000)) bad("invalid repeat count");   // parser.repeat, regexp/syntax/parse.go
            a = n;
        }
        return a;
    }
    int sequence() {
        std::vector<int> items;
        while (!done() && cur() != '|' && cur() != ')') { int a = repetition(); if (a >= 0) items.push_back(a); }
        if (items.empty()) return t_.add(RX_EMPTY);
        if (items.size() == 1) return items[0];
        int n = t_.add(RX_SEQ); t_.at(n).kids = items; return n;
    }
    int alternation() {
        std::vector<int> alts{sequence()};
        while (!done() && cur() == '|') { i_++; alts.push_back(sequence()); }
        if (alts.size() == 1) return alts[0];
        int n = t_.add(RX_ALTS); t_.at(n).kids = alts; return n;
    }
public:
    RxParser(const std::string& s, RxTree& t) : s_(s), t_(t) {}
    int parse() { int r = alternation(); if (!done()) bad("unexpected )"); return r; }
};

// ---- normalisation (structural model of regexutil's simplify/String/re-Parse fixed point) ----------------------------------
inline int rx_normalize(RxTree& t, int n) {
    RxNode cur = t.at(n);
    switch (cur.op) {
    case RX_GROUP: return rx_normalize(t, cur.kids[0]);
    case RX_STAR: case RX_PLUS: case RX_OPT: case RX_COUNTED: {
        int k = rx_normalize(t, cur.kids[0]);
        if (cur.op == RX_COUNTED && cur.lo == 1 && cur.hi == 1) return k;
        RxNode c; c.op = cur.op; c.lo = cur.lo; c.hi = cur.hi; c.kids = {k};
        return t.add(c);
    }
    case RX_ALTS: {
        RxNode c; c.op = RX_ALTS;
        for (int k : cur.kids) { int x = rx_normalize(t, k); if (t.at(x).op == RX_ALTS) for (int y : t.at(x).kids) c.kids.push_back(y); else c.kids.push_back(x); }
        if (c.kids.size() == 1) return c.kids[0];
        return t.add(c);
    }
    case RX_SEQ: {
        std::vector<int> flat;
        for (int k : cur.kids) {
            int x = rx_normalize(t, k);
            if (t.at(x).op == RX_EMPTY) continue;
            if (t.at(x).op == RX_SEQ) for (int y : t.at(x).kids) flat.push_back(y); else flat.push_back(x);
        }
        std::vector<int> out;
        for (int x : flat) {
            if (!out.empty() && t.at(out.back()).op == RX_LITERAL && t.at(x).op == RX_LITERAL && t.at(out.back()).fold == t.at(x).fold) {
                RxNode m = t.at(out.back()); m.cps.insert(m.cps.end(), t.at(x).cps.begin(), t.at(x).cps.end());
                out.back() = t.add(m);
            } else out.push_back(x);
        }
        if (out.empty()) return t.add(RX_EMPTY);
        if (out.size() == 1) return out[0];
        RxNode c; c.op = RX_SEQ; c.kids = out; return t.add(c);
    }
    default: return n;
    }
}

inline void rx_put_utf8(std::string& d, int32_t r) {
    if (r < 0 || r > 0x10FFFF || (r >= 0xD800 && r <= 0xDFFF)) r = kRuneError;
    if (r < 0x80) d.push_back((char)r);
    else if (r < 0x800) { d.push_back((char)(0xC0 | (r >> 6))); d.push_back((char)(0x80 | (r & 0x3F))); }
    else if (r < 0x10000) { d.push_back((char)(0xE0 | (r >> 12))); d.push_back((char)(0x80 | ((r >> 6) & 0x3F))); d.push_back((char)(0x80 | (r & 0x3F))); }
    else { d.push_back((char)(0xF0 | (r >> 18))); d.push_back((char)(0x80 | ((r >> 12) & 0x3F))); d.push_back((char)(0x80 | ((r >> 6) & 0x3F))); d.push_back((char)(0x80 | (r & 0x3F))); }
}
inline bool rx_plain_literal(const RxTree& t, int n, std::string* out) {   // getLiteral regexutil.go:141-149
    const RxNode& x = t.at(n);
    if (x.op == RX_GROUP) return rx_plain_literal(t, x.kids[0], out);
    if (x.op != RX_LITERAL || x.fold) return false;
    out->clear(); for (int32_t c : x.cps) rx_put_utf8(*out, c);
    return true;
}
inline bool rx_is_dot_rep(const RxTree& t, int n, RxOp op) { return t.at(n).op == op && t.at(t.at(n).kids[0]).op == RX_ANY; }   // isDotOp

// ---- rune-class DFA with delayed assertions --------------------------------------------------------------------------------
struct RxDfa {
    std::vector<int32_t> bounds;       // class k = [bounds[k], bounds[k+1]) ; bounds[0] == 0
    uint8_t ascii_class[128];
    uint32_t nclasses = 0, nstates = 0;
    std::vector<uint16_t> trans;       // nstates x nclasses: bit15 = match detected before consuming the rune; low 15 bits next state
    std::vector<uint8_t> accept_end;   // match when the text ends in this state
    static const uint16_t DEAD = 0x7FFF;
    bool empty_language = false;

    int class_of(int32_t r) const {
        if (r < 128) return ascii_class[r];
        int lo = 0, hi = (int)nclasses - 1;
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (bounds[mid] <= r) lo = mid; else hi = mid - 1; }
        return lo;
    }
    bool run(const uint8_t* s, uint32_t n) const {   // host reference of the device loop
        uint32_t st = 0;
        for (uint32_t i = 0; i < n;) {
            int w; int32_t r = decode_rune(s + i, n - i, &w); i += w;
            uint16_t e = trans[st * nclasses + class_of(r)];
            if (e & 0x8000) return true;
            st = e & 0x7FFF;
            if (st == DEAD) return false;
        }
        return accept_end[st] != 0;
    }
};

class RxCompiler {
    enum K : uint8_t { C_SET, C_ANY, C_ANYNL, C_FORK, C_GOTO, C_BOT, C_EOT, C_BOL, C_EOL, C_WB, C_NWB, C_NEVER };
    struct I { K k; int a = 0, b = 0; std::vector<int32_t> set; };
    std::vector<I> code_;
    const RxTree& t_;
    int emit(K k) { code_.push_back(I{k}); return (int)code_.size() - 1; }
    void gen(int n) {
        const RxNode& x = t_.at(n);
        switch (x.op) {
        case RX_EMPTY: break;
        case RX_LITERAL:
            for (int32_t c : x.cps) {
                int i = emit(C_SET);
                if (x.fold) { RangeSet rs; rs.put_folded(c, c); rs.canon(); for (auto& p : rs.v) { code_[i].set.push_back(p.first); code_[i].set.push_back(p.second); } }
                else code_[i].set = {c, c};
            }
            break;
        case RX_CLASS: { int i = emit(x.cps.empty() ? C_NEVER : C_SET); code_[i].set = x.cps; break; }
        case RX_ANY: emit(C_ANY); break;
        case RX_ANY_NOT_NL: emit(C_ANYNL); break;
        case RX_BOT: emit(C_BOT); break; case RX_EOT: emit(C_EOT); break; case RX_BOL: emit(C_BOL); break; case RX_EOL: emit(C_EOL); break;
        case RX_WORD_B: emit(C_WB); break; case RX_NOT_WORD_B: emit(C_NWB); break;
        case RX_GROUP: gen(x.kids[0]); break;
        case RX_SEQ: for (int k : x.kids) gen(k); break;
        case RX_ALTS: {
            std::vector<int> exits;
            for (size_t i = 0; i < x.kids.size(); i++) {
                if (i + 1 == x.kids.size()) { gen(x.kids[i]); break; }
                int f = emit(C_FORK); code_[f].a = f + 1;
                gen(x.kids[i]);
                exits.push_back(emit(C_GOTO));
                code_[f].b = (int)code_.size();
            }
            for (int e : exits) code_[e].a = (int)code_.size();
            break;
        }
        case RX_STAR: { int f = emit(C_FORK); code_[f].a = f + 1; gen(x.kids[0]); int g = emit(C_GOTO); code_[g].a = f; code_[f].b = (int)code_.size(); break; }
        case RX_PLUS: { int s = (int)code_.size(); gen(x.kids[0]); int f = emit(C_FORK); code_[f].a = s; code_[f].b = f + 1; break; }
        case RX_OPT: { int f = emit(C_FORK); code_[f].a = f + 1; gen(x.kids[0]); code_[f].b = (int)code_.size(); break; }
        case RX_COUNTED: {
            for (int i = 0; i < x.lo; i++) gen(x.kids[0]);
            if (x.hi < 0) { int f = emit(C_FORK); code_[f].a = f + 1; gen(x.kids[0]); int g = emit(C_GOTO); code_[g].a = f; code_[f].b = (int)code_.size(); }
            else for (int i = x.lo; i < x.hi; i++) { int f = emit(C_FORK); code_[f].a = f + 1; gen(x.kids[0]); code_[f].b = (int)code_.size(); }
            break;
        }
        }
        if (code_.size() > 20000) throw RxError("regexp too large");
    }
    enum T : uint8_t { T_BEGIN = 0, T_WORD = 1, T_NL = 2, T_OTHER = 3, T_END = 4 };
    static bool wordy(uint8_t t) { return t == T_WORD; }
    // epsilon closure at one text position, resolving assertions with (prev, cur) character types
    void closure(const std::vector<int>& seeds, uint8_t prev, uint8_t cur, std::vector<int>& out, bool* matched) {
        std::vector<char> seen(code_.size() + 1, 0);
        std::vector<int> st(seeds.rbegin(), seeds.rend());
        out.clear(); *matched = false;
        while (!st.empty()) {
            int pc = st.back(); st.pop_back();
            if (seen[pc]) continue;
            seen[pc] = 1;
            if (pc == (int)code_.size()) { *matched = true; continue; }
            const I& in = code_[pc];
            bool pass;
            switch (in.k) {
            case C_GOTO: st.push_back(in.a); continue;
            case C_FORK: st.push_back(in.b); st.push_back(in.a); continue;
            case C_BOT: pass = prev == T_BEGIN; break;
            case C_EOT: pass = cur == T_END; break;
            case C_BOL: pass = prev == T_BEGIN || prev == T_NL; break;
            case C_EOL: pass = cur == T_END || cur == T_NL; break;
            case C_WB: pass = wordy(prev) != wordy(cur); break;
            case C_NWB: pass = wordy(prev) == wordy(cur); break;
            default: out.push_back(pc); continue;
            }
            if (pass) st.push_back(pc + 1);
        }
        std::sort(out.begin(), out.end());
    }
public:
    explicit RxCompiler(const RxTree& t) : t_(t) {}
    RxDfa build(int root, size_t max_states = 4000) {
        gen(root);
        RxDfa d;
        // rune classes: boundaries of every set + word chars + '\n'
        std::vector<int32_t> cuts = {0, '\n', '\n' + 1, '0', '9' + 1, 'A', 'Z' + 1, '_', '_' + 1, 'a', 'z' + 1, 0x80};
        for (auto& in : code_) for (size_t i = 0; i + 1 < in.set.size(); i += 2) { cuts.push_back(in.set[i]); if (in.set[i + 1] < 0x10FFFF) cuts.push_back(in.set[i + 1] + 1); }
        std::sort(cuts.begin(), cuts.end()); cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
        d.bounds = cuts; d.nclasses = (uint32_t)cuts.size();
        if (d.nclasses > 1024) throw RxError("regexp uses too many distinct character classes");
        for (int c = 0; c < 128; c++) { int k = (int)(std::upper_bound(cuts.begin(), cuts.end(), c) - cuts.begin()) - 1; d.ascii_class[c] = (uint8_t)k; if (k > 255) throw RxError("too many ASCII classes"); }
        std::vector<uint8_t> ctype(d.nclasses);
        for (uint32_t k = 0; k < d.nclasses; k++) { int32_t r = cuts[k]; ctype[k] = r == '\n' ? T_NL : (r < 128 && is_token_char((uint32_t)r)) ? T_WORD : T_OTHER; }
        // per-instruction class membership
        auto consumes = [&](const I& in, uint32_t k) {
            int32_t r = cuts[k];
            if (in.k == C_ANY) return true;
            if (in.k == C_ANYNL) return r != '\n';
            if (in.k != C_SET) return false;
            for (size_t i = 0; i + 1 < in.set.size(); i += 2) if (r >= in.set[i] && r <= in.set[i + 1]) return true;
            return false;
        };
        std::map<std::pair<std::vector<int>, uint8_t>, uint32_t> ids;
        std::vector<std::pair<std::vector<int>, uint8_t>> states;
        auto intern = [&](std::vector<int> pcs, uint8_t prev) -> uint32_t {
            auto key = std::make_pair(std::move(pcs), prev);
            auto it = ids.find(key);
            if (it != ids.end()) return it->second;
            if (states.size() >= max_states) throw RxError("regexp is too complex for the DFA engine (state limit exceeded)");
            uint32_t id = (uint32_t)states.size();
            ids.emplace(key, id); states.push_back(key);
            return id;
        };
        intern({}, T_BEGIN);
        std::vector<int> cl;
        size_t work = 0;   // NFA threads visited over all (state, class) pairs: bounds the compile time of hostile expressions to about a second
        for (uint32_t s = 0; s < states.size(); s++) {
            auto [pcs, prev] = states[s];
            std::vector<int> seeds = pcs; seeds.push_back(0);   // unanchored search: a new thread starts at every position
            d.trans.resize((size_t)(s + 1) * d.nclasses);
            for (uint32_t k = 0; k < d.nclasses; k++) {
                bool m; closure(seeds, prev, ctype[k], cl, &m);
                work += cl.size() + seeds.size();
                if (work > 30000000) throw RxError("regexp is too complex for the DFA engine (work limit exceeded)");
                std::vector<int> next;
                for (int pc : cl) if (consumes(code_[pc], k)) next.push_back(pc + 1);
                std::sort(next.begin(), next.end()); next.erase(std::unique(next.begin(), next.end()), next.end());
                uint32_t ns = intern(next, ctype[k]);
                d.trans[(size_t)s * d.nclasses + k] = (uint16_t)(ns | (m ? 0x8000 : 0));
            }
            bool m; closure(seeds, prev, T_END, cl, &m);
            d.accept_end.push_back(m ? 1 : 0);
        }
        d.nstates = (uint32_t)states.size();
        // prune: states from which no match is reachable become DEAD (early exit on device)
        std::vector<char> live(d.nstates, 0);
        for (bool changed = true; changed;) {
            changed = false;
            for (uint32_t s = 0; s < d.nstates; s++) {
                if (live[s]) continue;
                bool l = d.accept_end[s];
                for (uint32_t k = 0; k < d.nclasses && !l; k++) { uint16_t e = d.trans[(size_t)s * d.nclasses + k]; l = (e & 0x8000) || live[e & 0x7FFF]; }
                if (l) { live[s] = 1; changed = true; }
            }
        }
        for (auto& e : d.trans) if (!(e & 0x8000) && !live[e & 0x7FFF]) e = RxDfa::DEAD;
        d.empty_language = !live[0];
        return d;
    }
};

// ---- the regexutil.Regex equivalent ----------------------------------------------------------------------------------------
struct CompiledRegex {
    std::string expr, prefix;
    bool isOnlyPrefix = false, isSuffixDotStar = false, isSuffixDotPlus = false;
    std::string substrDotStar, substrDotPlus;
    bool hasOrValues = false;          // informational; or-values are matched through the DFA (same language)
    RxDfa suffix;                      // anchored at the start of the remainder iff prefix != ""
    std::vector<std::string> literals; // GetLiterals() (regex.go:101-124), for bloom tokens
    std::string tailLiteral;           // prefix != "" and suffix == `(?s:.*LIT)`: suffixRe.MatchString(rem) == strings.Contains(rem, LIT)

    static bool contains(const uint8_t* s, uint32_t n, const std::string& sub, uint32_t from = 0) { return find_bytes(s, n, (const uint8_t*)sub.data(), (uint32_t)sub.size(), from) >= 0; }

    // MatchString regex.go:86-212 (host reference of the device implementation; used for const / dict values)
    bool match(const uint8_t* s, uint32_t n) const {
        if (isOnlyPrefix) return prefix.empty() || contains(s, n, prefix);
        if (prefix.empty()) {
            if (isSuffixDotStar) return true;
            if (isSuffixDotPlus) return n > 0;
            if (!substrDotStar.empty()) return contains(s, n, substrDotStar);
            if (!substrDotPlus.empty()) { int k = find_bytes(s, n, (const uint8_t*)substrDotPlus.data(), (uint32_t)substrDotPlus.size(), 0); return k > 0 && (uint32_t)k + substrDotPlus.size() < n; }
            return suffix.run(s, n);
        }
        int k = find_bytes(s, n, (const uint8_t*)prefix.data(), (uint32_t)prefix.size(), 0);
        if (k < 0) return false;
        uint32_t rem = (uint32_t)k + (uint32_t)prefix.size();
        if (isSuffixDotStar) return true;
        if (isSuffixDotPlus) return n > rem;
        if (!substrDotStar.empty()) return contains(s + rem, n - rem, substrDotStar);
        if (!substrDotPlus.empty()) { int m = find_bytes(s + rem, n - rem, (const uint8_t*)substrDotPlus.data(), (uint32_t)substrDotPlus.size(), 0); return m > 0 && (uint32_t)m + substrDotPlus.size() < n - rem; }
        for (;;) {
            if (suffix.run(s + rem, n - rem)) return true;
            k = find_bytes(s, n, (const uint8_t*)prefix.data(), (uint32_t)prefix.size(), (uint32_t)k + 1);
            if (k < 0) return false;
            rem = (uint32_t)k + (uint32_t)prefix.size();
        }
    }
};

inline std::string rx_substring_literal(const RxTree& t, int n, RxOp op) {   // getSubstringLiteral regexutil.go:316-328
    const RxNode& x = t.at(n);
    if (x.op != RX_SEQ || x.kids.size() != 3) return "";
    if (!rx_is_dot_rep(t, x.kids[0], op) || !rx_is_dot_rep(t, x.kids[2], op)) return "";
    std::string v; if (!rx_plain_literal(t, x.kids[1], &v)) return "";
    return v;
}
// does getOrValues (regexutil.go:67-139) succeed? only the yes/no matters: with or-values the reference matches by
// strings.Contains / HasPrefix, which accepts exactly the language of the suffix, so the DFA is used either way.
inline bool rx_or_values(const RxTree& t, int n, std::vector<std::string>& out) {
    const RxNode& x = t.at(n);
    switch (x.op) {
    case RX_GROUP: return rx_or_values(t, x.kids[0], out);
    case RX_LITERAL: { std::string v; if (!rx_plain_literal(t, n, &v)) return false; out = {v}; return true; }
    case RX_EMPTY: out = {""}; return true;
    case RX_ALTS: { std::vector<std::string> a; for (int k : x.kids) { std::vector<std::string> c; if (!rx_or_values(t, k, c) || c.empty()) return false; a.insert(a.end(), c.begin(), c.end()); if (a.size() > 100) return false; } out = a; return true; }
    case RX_CLASS: { std::vector<std::string> a; for (size_t i = 0; i + 1 < x.cps.size(); i += 2) for (int32_t c = x.cps[i]; c <= x.cps[i + 1]; c++) { std::string s; rx_put_utf8(s, c); a.push_back(s); if (a.size() > 100) return false; } if (a.empty()) return false; out = a; return true; }
    case RX_SEQ: {
        std::vector<std::string> acc = {""};
        for (int k : x.kids) {
            std::vector<std::string> c; if (!rx_or_values(t, k, c) || c.empty()) return false;
            if (acc.size() * c.size() > 100) return false;
            std::vector<std::string> nx; for (auto& p : acc) for (auto& q : c) nx.push_back(p + q);
            acc.swap(nx);
        }
        out = acc; return true;
    }
    default: return false;
    }
}

inline CompiledRegex compile_regex(const std::string& expr) {
    CompiledRegex r; r.expr = expr;
    RxTree t;
    int raw = RxParser(expr, t).parse();
    // GetLiterals on the raw tree: top-level captures unwrapped; adjacent literal chars of a concat form one literal
    {
        int n = raw;
        while (t.at(n).op == RX_GROUP) n = t.at(n).kids[0];
        std::string v;
        if (rx_plain_literal(t, n, &v)) r.literals = {v};
        else if (t.at(n).op == RX_SEQ) {
            std::string run; bool in = false;
            for (int k : t.at(n).kids) {
                const RxNode& x = t.at(k);
                if (x.op == RX_LITERAL && !x.fold) { for (int32_t c : x.cps) rx_put_utf8(run, c); in = true; continue; }
                if (in) { r.literals.push_back(run); run.clear(); in = false; }
                if (x.op == RX_LITERAL) continue;
                if (rx_plain_literal(t, k, &v)) r.literals.push_back(v);
            }
            if (in) r.literals.push_back(run);
        }
    }
    int sre = rx_normalize(t, raw);
    int suffix;
    std::string lit;
    if (t.at(sre).op == RX_EMPTY) suffix = sre;
    else if (rx_plain_literal(t, sre, &lit)) { r.prefix = lit; suffix = t.add(RX_EMPTY); }
    else if (t.at(sre).op == RX_SEQ && rx_plain_literal(t, t.at(sre).kids[0], &lit)) {
        r.prefix = lit;
        std::vector<int> rest(t.at(sre).kids.begin() + 1, t.at(sre).kids.end());
        if (rest.size() == 1) suffix = rest[0]; else { RxNode c; c.op = RX_SEQ; c.kids = rest; suffix = t.add(c); }
    } else suffix = sre;
    if (t.at(suffix).op == RX_ANY) suffix = t.add(RX_ANY_NOT_NL);   // "(?s:.)" -> "." textual replacement (regexutil.go:229)
    if (rx_is_dot_rep(t, suffix, RX_STAR)) suffix = t.add(RX_EMPTY);
    else if (t.at(suffix).op == RX_SEQ) {
        std::vector<int> subs = t.at(suffix).kids;
        if (r.prefix.empty()) while (!subs.empty() && rx_is_dot_rep(t, subs.front(), RX_STAR)) subs.erase(subs.begin());
        while (!subs.empty() && rx_is_dot_rep(t, subs.back(), RX_STAR)) subs.pop_back();
        if (subs.empty()) suffix = t.add(RX_EMPTY);
        else if (subs.size() == 1) suffix = subs[0];
        else { RxNode c; c.op = RX_SEQ; c.kids = subs; suffix = t.add(c); }
    }
    std::vector<std::string> ov;
    r.hasOrValues = rx_or_values(t, suffix, ov) && !ov.empty();
    r.isOnlyPrefix = r.hasOrValues && ov.size() == 1 && ov[0].empty();
    r.isSuffixDotStar = rx_is_dot_rep(t, suffix, RX_STAR);
    r.isSuffixDotPlus = rx_is_dot_rep(t, suffix, RX_PLUS);
    r.substrDotStar = rx_substring_literal(t, suffix, RX_STAR);
    r.substrDotPlus = rx_substring_literal(t, suffix, RX_PLUS);
    if (!r.prefix.empty() && t.at(suffix).op == RX_SEQ && t.at(suffix).kids.size() == 2 && rx_is_dot_rep(t, t.at(suffix).kids[0], RX_STAR)) {
        std::string tl; if (rx_plain_literal(t, t.at(suffix).kids[1], &tl) && !tl.empty()) r.tailLiteral = tl;
    }
    int root = suffix;
    if (!r.prefix.empty()) { RxNode c; c.op = RX_SEQ; c.kids = {t.add(RX_BOT), suffix}; root = t.add(c); }
    r.suffix = RxCompiler(t).build(root);
    return r;
}

}  // namespace vl
