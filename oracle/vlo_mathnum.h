// ORACLE -- TEST INFRASTRUCTURE ONLY (see vlo_util.h header).
//
// parseMathNumber (lib/logstorage/pipe_math.go:1066-1080) and what it is made of: the value -> float64 conversion behind range(),
// le_field() and lt_field().  Restated from
//   lib/logstorage/block_result.go:2710-2752   tryParseNumber, isLikelyNumber
//   lib/logstorage/parser.go:3077-3097         isNumberPrefix
//   lib/logstorage/values_encoder.go:761-850   tryParseFloat64Prefix, tryParseFloat64 (isExact=false)
//   lib/logstorage/values_encoder.go:855-974   tryParseBytes, addInt64NoOverflow
//   lib/logstorage/values_encoder.go:990-1061  tryParseDuration
//   lib/logstorage/values_encoder.go:340-423   TryParseTimestampRFC3339Nano, parseTimezoneOffset, tryParseHHMM
// Parity notes.  (1) A timestamp without `Z` or a numeric offset takes the process' local zone in the reference; this restatement
// uses UTC.  (2) strconv.ParseFloat / strconv.ParseInt(s, 0, 64) live in Go's standard library (not under /root/reference): decimal
// and hexadecimal floats, "inf"/"infinity", base prefixes 0x/0o/0b/0 and `_` digit separators are restated from their documentation;
// only the syntaxes the reference's own tables exercise are pinned (filter_range_test.go, filter_le_field_test.go).
#pragma once
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include "vlo_util.h"

namespace vlo {

// tryParseFloat64 = tryParseFloat64Internal(s, isExact=false): like the exact form, minus the 2^53 bound on plain integers
inline bool try_parse_float64(sv s, double* out) {
    if (s.empty() || s.size() > strlen("-18_446_744_073_709_551_615")) return false;
    if (s.find('.') != sv::npos) return try_parse_float64_exact(s, out);   // the dotted branch does not depend on isExact
    bool minus = s[0] == '-';
    if (minus) s.remove_prefix(1);
    uint64_t v;
    if (!try_parse_uint64(s, &v)) return false;
    double f = (double)v;
    *out = minus ? -f : f;
    return true;
}
// tryParseFloat64Prefix :762-773
inline bool try_parse_float64_prefix(sv s, double* f, sv* tail) {
    size_t i = 0;
    while (i < s.size() && ((s[i] >= '0' && s[i] <= '9') || s[i] == '.' || s[i] == '_')) i++;
    if (i == 0) return false;
    if (!try_parse_float64(s.substr(0, i), f)) return false;
    *tail = s.substr(i);
    return true;
}
inline int64_t go_int64_of_float(double f) {   // int64(f) for the values that occur here (saturating where Go's result is implementation-defined)
    if (!(f == f)) return INT64_MIN;
    if (f >= 9223372036854775808.0) return INT64_MIN;   // amd64 CVTTSD2SQ: out of range -> 0x8000000000000000
    if (f < -9223372036854775808.0) return INT64_MIN;
    return (int64_t)f;
}
inline int64_t add_int64_no_overflow(int64_t n, double f) {   // :968-974
    int64_t x = go_int64_of_float(f);
    if (n < 0 || x < 0 || x > INT64_MAX - n) return INT64_MAX;
    return n + x;
}
inline bool has_prefix(sv s, const char* p) { size_t n = strlen(p); return s.size() >= n && memcmp(s.data(), p, n) == 0; }

// tryParseDuration :990-1061
inline bool try_parse_duration(sv s, int64_t* out) {
    if (s.empty()) return false;
    bool minus = s[0] == '-';
    if (minus) s.remove_prefix(1);
    static const double US = 1e3, MS = 1e6, SEC = 1e9, MIN = 60e9, HOUR = 3600e9, DAY = 24 * 3600e9, WEEK = 7 * 24 * 3600e9, YEAR = 365 * 24 * 3600e9;
    int64_t nsecs = 0;
    while (!s.empty()) {
        double f; sv tail;
        if (!try_parse_float64_prefix(s, &f, &tail)) return false;
        s = tail;
        if (s.empty()) return false;
        if (s.size() >= 3 && has_prefix(s, "\xC2\xB5s")) { nsecs = add_int64_no_overflow(nsecs, f * US); s.remove_prefix(3); continue; }
        if (s.size() >= 2) {
            if (has_prefix(s, "ms")) { nsecs = add_int64_no_overflow(nsecs, f * MS); s.remove_prefix(2); continue; }
            if (has_prefix(s, "ns")) { nsecs = add_int64_no_overflow(nsecs, f); s.remove_prefix(2); continue; }
        }
        double unit;
        switch (s[0]) {
        case 'y': unit = YEAR; break; case 'w': unit = WEEK; break; case 'd': unit = DAY; break; case 'h': unit = HOUR; break;
        case 'm': unit = MIN; break; case 's': unit = SEC; break;
        default: return false;
        }
        nsecs = add_int64_no_overflow(nsecs, f * unit); s.remove_prefix(1);
    }
    *out = minus ? -nsecs : nsecs;
    return true;
}
// tryParseBytes :855-966
inline bool try_parse_bytes(sv s, int64_t* out) {
    if (s.empty()) return false;
    bool minus = s[0] == '-';
    if (minus) s.remove_prefix(1);
    int64_t n = 0;
    struct U { const char* name; double mul; };
    static const U u3[] = {{"KiB", 1024.0}, {"MiB", 1048576.0}, {"GiB", 1073741824.0}, {"TiB", 1099511627776.0}};
    static const U u2[] = {{"Ki", 1024.0}, {"Mi", 1048576.0}, {"Gi", 1073741824.0}, {"Ti", 1099511627776.0}, {"KB", 1e3}, {"MB", 1e6}, {"GB", 1e9}, {"TB", 1e12}};
    static const U u1[] = {{"B", 1.0}, {"K", 1e3}, {"M", 1e6}, {"G", 1e9}, {"T", 1e12}};
    while (!s.empty()) {
        double f; sv tail;
        if (!try_parse_float64_prefix(s, &f, &tail)) return false;
        if (tail.empty()) { double ip; if (std::modf(f, &ip) != 0) return false; }   // no fractional numbers without a suffix
        s = tail;
        if (s.empty()) { n = add_int64_no_overflow(n, f); continue; }
        bool hit = false;
        if (s.size() >= 3) for (auto& u : u3) if (has_prefix(s, u.name)) { n = add_int64_no_overflow(n, f * u.mul); s.remove_prefix(3); hit = true; break; }
        if (!hit && s.size() >= 2) for (auto& u : u2) if (has_prefix(s, u.name)) { n = add_int64_no_overflow(n, f * u.mul); s.remove_prefix(2); hit = true; break; }
        if (!hit) for (auto& u : u1) if (has_prefix(s, u.name)) { n = add_int64_no_overflow(n, f * u.mul); s.remove_prefix(1); hit = true; break; }
        if (!hit) return false;
    }
    *out = minus ? -n : n;
    return true;
}

// isNumberPrefix parser.go:3077-3097, isLikelyNumber block_result.go:2739-2752
inline bool is_number_prefix(sv s) {
    if (s.empty()) return false;
    if (s[0] == '-' || s[0] == '+') { s.remove_prefix(1); if (s.empty()) return false; }
    if (s.size() >= 3) { bool inf = s.size() == 3; for (size_t i = 0; inf && i < 3; i++) inf = (s[i] | 0x20) == "inf"[i]; if (inf) return true; }
    if (s[0] == '.') { s.remove_prefix(1); if (s.empty()) return false; }
    return s[0] >= '0' && s[0] <= '9';
}
inline bool is_likely_number(sv s) {
    if (!is_number_prefix(s)) return false;
    if (std::count(s.begin(), s.end(), '.') > 1) return false;                                   // likely an IP address
    if (s.find(':') != sv::npos || std::count(s.begin(), s.end(), '-') > 2) return false;       // likely a timestamp
    return true;
}
// Go digit-separator rule (strconv underscoreOK): `_` only between digits or right after a base prefix
inline bool underscores_ok(sv s) {
    char saw = '^'; size_t i = 0;
    if (!s.empty() && (s[0] == '-' || s[0] == '+')) i = 1;
    bool hex = false;
    if (s.size() - i >= 2 && s[i] == '0' && ((s[i + 1] | 0x20) == 'b' || (s[i + 1] | 0x20) == 'o' || (s[i + 1] | 0x20) == 'x')) { hex = (s[i + 1] | 0x20) == 'x'; i += 2; saw = '0'; }
    for (; i < s.size(); i++) {
        char c = s[i];
        if ((c >= '0' && c <= '9') || (hex && (c | 0x20) >= 'a' && (c | 0x20) <= 'f')) { saw = '0'; continue; }
        if (c == '_') { if (saw != '0') return false; saw = '_'; continue; }
        if (saw == '_') return false;
        saw = '!';
    }
    return saw != '_';
}
// strconv.ParseFloat(s, 64) for strings that passed isLikelyNumber
inline bool go_parse_float(sv s, double* out) {
    std::string t(s);
    if (t.find('_') != std::string::npos) { if (!underscores_ok(s)) return false; t.erase(std::remove(t.begin(), t.end(), '_'), t.end()); }
    {   // inf / infinity (any case), optionally signed
        sv u(t); bool neg = false;
        if (!u.empty() && (u[0] == '+' || u[0] == '-')) { neg = u[0] == '-'; u.remove_prefix(1); }
        std::string low; for (char c : u) low.push_back((char)(c | 0x20));
        if (low == "inf" || low == "infinity") { *out = neg ? -INFINITY : INFINITY; return true; }
    }
    for (char c : t) if (!((c >= '0' && c <= '9') || c == '.' || c == '+' || c == '-' || (c | 0x20) == 'e' || (c | 0x20) == 'x' || (c | 0x20) == 'p' || ((c | 0x20) >= 'a' && (c | 0x20) <= 'f'))) return false;
    bool hex = t.find('x') != std::string::npos || t.find('X') != std::string::npos;
    if (hex && t.find('p') == std::string::npos && t.find('P') == std::string::npos) return false;   // hexadecimal mantissa requires a 'p' exponent
    if (!hex && t.find_first_of("abcdfABCDF") != std::string::npos) return false;
    errno = 0; char* end = nullptr;
    double v = strtod(t.c_str(), &end);
    if (end != t.c_str() + t.size() || t.empty()) return false;
    if (errno == ERANGE && std::isinf(v)) return false;   // ParseFloat: value out of range -> err != nil
    *out = v;
    return true;
}
// strconv.ParseInt(s, 0, 64)
inline bool go_parse_int0(sv s, int64_t* out) {
    if (s.empty()) return false;
    if (s.find('_') != sv::npos && !underscores_ok(s)) return false;
    bool neg = false;
    if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; s.remove_prefix(1); }
    if (s.empty()) return false;
    int base = 10;
    if (s[0] == '0' && s.size() >= 2) {
        char p = (char)(s[1] | 0x20);
        if (p == 'x') { base = 16; s.remove_prefix(2); } else if (p == 'b') { base = 2; s.remove_prefix(2); } else if (p == 'o') { base = 8; s.remove_prefix(2); } else { base = 8; s.remove_prefix(1); }
        if (s.empty()) return false;
    }
    unsigned __int128 v = 0; bool any = false;
    for (char c : s) {
        if (c == '_') continue;
        int d = (c >= '0' && c <= '9') ? c - '0' : ((c | 0x20) >= 'a' && (c | 0x20) <= 'z') ? (c | 0x20) - 'a' + 10 : 99;
        if (d >= base) return false;
        v = v * (unsigned)base + (unsigned)d; any = true;
        if (v > ((unsigned __int128)1 << 63)) return false;
    }
    if (!any) return false;
    if (!neg && v > (unsigned __int128)INT64_MAX) return false;
    *out = neg ? (int64_t)(0 - (uint64_t)v) : (int64_t)v;
    return true;
}
// tryParseNumber block_result.go:2710-2737
inline bool try_parse_number(sv s, double* out) {
    if (s.empty()) return false;
    if (try_parse_float64(s, out)) return true;
    int64_t n;
    if (try_parse_duration(s, &n)) { *out = (double)n; return true; }
    if (try_parse_bytes(s, &n)) { *out = (double)n; return true; }
    if (is_likely_number(s)) {
        if (go_parse_float(s, out)) return true;
        if (go_parse_int0(s, &n)) { *out = (double)n; return true; }
    }
    return false;
}
// TryParseTimestampRFC3339Nano :340-381 (local zone == UTC here)
inline bool try_parse_timestamp_rfc3339nano(sv s, int64_t* out) {
    if (s.size() < strlen("2006-01-02T15:04:05")) return false;
    int64_t secs;
    if (!try_parse_timestamp_secs(s, &secs)) return false;
    int64_t nsecs = secs * 1000000000LL;
    // parseTimezoneOffset :383-406
    if (!s.empty() && s.back() == 'Z') s.remove_suffix(1);
    else {
        size_t n = s.find_last_of("+-");
        if (n != sv::npos) {
            sv off = s.substr(n + 1);
            bool minus = s[n] == '-';
            if (off.size() != 5 || off[2] != ':') return false;   // tryParseHHMM :408-423
            uint64_t hh, mm;
            if (!try_parse_date_uint64(off.substr(0, 2), &hh) || hh > 24) return false;
            if (!try_parse_date_uint64(off.substr(3), &mm) || mm > 60) return false;
            int64_t o = (int64_t)hh * 3600000000000LL + (int64_t)mm * 60000000000LL;
            nsecs -= minus ? -o : o;
            s = s.substr(0, n);
        }
    }
    if (s.empty()) { *out = nsecs; return true; }
    if (s[0] == '.') s.remove_prefix(1);
    size_t digits = s.size();
    if (digits > 9) return false;
    uint64_t frac;
    if (!try_parse_date_uint64(s, &frac)) return false;
    for (size_t i = digits; i < 9; i++) frac *= 10;
    *out = nsecs + (int64_t)frac;
    return true;
}
// parseMathNumber pipe_math.go:1066-1080
inline double parse_math_number(sv s) {
    double f;
    if (try_parse_number(s, &f)) return f;
    int64_t ns;
    if (try_parse_timestamp_rfc3339nano(s, &ns)) return (double)ns;
    uint32_t ip;
    if (try_parse_ipv4(s, &ip)) return (double)ip;
    return NAN;
}

}  // namespace vlo
