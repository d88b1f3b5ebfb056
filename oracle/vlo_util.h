// ORACLE -- TEST INFRASTRUCTURE ONLY.
// CPU restatement of the VictoriaLogs block-scan primitives (lib/logstorage).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may use anything under oracle/.  The product
// (victorialogs_b200/) never includes, links or calls this code.
//
// Parity status: pinned by the reference's own known-answer tests transcribed into tests/ (bloom bytes, token
// hashes, tokenizer tables, matchPhrase/matchPrefix tables, values-encoder tables, filter_*_test.go tables).
// The Go toolchain is absent, so the reference itself cannot be executed here.
//
// All file:line citations are relative to /root/reference/.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <string>
#include <string_view>
#include <vector>
#include <unordered_set>
#include <unordered_map>
#include <algorithm>
#include <stdexcept>
#include <charconv>

namespace vlo {

using sv = std::string_view;

// ---------------------------------------------------------------------------------------------------------------
// UTF-8 (Go unicode/utf8 semantics: DecodeRuneInString / DecodeLastRuneInString; stdlib, not under /root/reference)
// ---------------------------------------------------------------------------------------------------------------
static const int32_t RuneError = 0xFFFD;

inline int32_t decode_rune(const uint8_t* p, size_t n, int* size) {
    if (n < 1) { *size = 0; return RuneError; }
    uint8_t p0 = p[0];
    if (p0 < 0x80) { *size = 1; return p0; }
    // first[] table of utf8.go
    int sz; uint8_t lo = 0x80, hi = 0xBF;
    if (p0 < 0xC2) { *size = 1; return RuneError; }
    else if (p0 <= 0xDF) sz = 2;
    else if (p0 == 0xE0) { sz = 3; lo = 0xA0; }
    else if (p0 <= 0xEC) sz = 3;
    else if (p0 == 0xED) { sz = 3; hi = 0x9F; }
    else if (p0 <= 0xEF) sz = 3;
    else if (p0 == 0xF0) { sz = 4; lo = 0x90; }
    else if (p0 <= 0xF3) sz = 4;
    else if (p0 == 0xF4) { sz = 4; hi = 0x8F; }
    else { *size = 1; return RuneError; }
    if ((int)n < sz) { *size = 1; return RuneError; }
    uint8_t b1 = p[1];
    if (b1 < lo || hi < b1) { *size = 1; return RuneError; }
    if (sz == 2) { *size = 2; return ((int32_t)(p0 & 0x1F) << 6) | (b1 & 0x3F); }
    uint8_t b2 = p[2];
    if (b2 < 0x80 || 0xBF < b2) { *size = 1; return RuneError; }
    if (sz == 3) { *size = 3; return ((int32_t)(p0 & 0x0F) << 12) | ((int32_t)(b1 & 0x3F) << 6) | (b2 & 0x3F); }
    uint8_t b3 = p[3];
    if (b3 < 0x80 || 0xBF < b3) { *size = 1; return RuneError; }
    *size = 4;
    return ((int32_t)(p0 & 0x07) << 18) | ((int32_t)(b1 & 0x3F) << 12) | ((int32_t)(b2 & 0x3F) << 6) | (b3 & 0x3F);
}

inline int32_t decode_last_rune(const uint8_t* p, size_t n, int* size) {
    if (n == 0) { *size = 0; return RuneError; }
    long end = (long)n;
    long start = end - 1;
    uint8_t r = p[start];
    if (r < 0x80) { *size = 1; return r; }
    long lim = end - 4;
    if (lim < 0) lim = 0;
    for (start--; start >= lim; start--) {
        if ((p[start] & 0xC0) != 0x80) break;
    }
    if (start < 0) start = 0;
    int sz;
    int32_t rr = decode_rune(p + start, (size_t)(end - start), &sz);
    if (start + sz != end) { *size = 1; return RuneError; }
    *size = sz;
    return rr;
}

inline void append_rune(std::string& dst, int32_t r) {
    // utf8.AppendRune
    if (r < 0 || r > 0x10FFFF || (r >= 0xD800 && r <= 0xDFFF)) r = RuneError;
    if (r < 0x80) dst.push_back((char)r);
    else if (r < 0x800) { dst.push_back((char)(0xC0 | (r >> 6))); dst.push_back((char)(0x80 | (r & 0x3F))); }
    else if (r < 0x10000) { dst.push_back((char)(0xE0 | (r >> 12))); dst.push_back((char)(0x80 | ((r >> 6) & 0x3F))); dst.push_back((char)(0x80 | (r & 0x3F))); }
    else { dst.push_back((char)(0xF0 | (r >> 18))); dst.push_back((char)(0x80 | ((r >> 12) & 0x3F))); dst.push_back((char)(0x80 | ((r >> 6) & 0x3F))); dst.push_back((char)(0x80 | (r & 0x3F))); }
}

#include "unicode_tables.inc"
#include "unicode_case.inc"

// unicode.ToLower / unicode.ToUpper (simple case mappings) and the string forms the i(...) filters use:
// strings.ToLower / strings.ToUpper / stringsutil.AppendLowercase (vm/lib/stringsutil/stringsutil.go:26-51).  All three walk the string
// rune by rune; an invalid byte decodes to utf8.RuneError and is written back as U+FFFD (3 bytes).
inline int32_t map_case(const unsigned int (*tab)[2], int n, int32_t r) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) { int mid = (lo + hi) / 2; if ((int32_t)tab[mid][0] == r) return (int32_t)tab[mid][1]; if ((int32_t)tab[mid][0] < r) lo = mid + 1; else hi = mid - 1; }
    return r;
}
inline int32_t to_lower_rune(int32_t r) { if (r < 0x80) return (r >= 'A' && r <= 'Z') ? r + 32 : r; return map_case(VL_TOLOWER, VL_TOLOWER_COUNT, r); }
inline int32_t to_upper_rune(int32_t r) { if (r < 0x80) return (r >= 'a' && r <= 'z') ? r - 32 : r; return map_case(VL_TOUPPER, VL_TOUPPER_COUNT, r); }
inline std::string map_string(sv s, int32_t (*f)(int32_t)) {
    std::string out; const uint8_t* p = (const uint8_t*)s.data(); size_t left = s.size();
    while (left) { int sz; int32_t r = decode_rune(p, left, &sz); append_rune(out, f(r)); p += sz; left -= (size_t)sz; }
    return out;
}
inline std::string strings_to_lower(sv s) { return map_string(s, to_lower_rune); }
inline std::string strings_to_upper(sv s) { return map_string(s, to_upper_rune); }
inline bool is_ascii_lowercase(sv s) {   // filter_any_case_phrase.go isASCIILowercase
    for (unsigned char c : s) if (c >= 0x80 || (c >= 'A' && c <= 'Z')) return false;
    return true;
}

// lib/logstorage/tokenizer.go:128-148 isTokenChar / isTokenRune
inline bool is_token_char(uint8_t c) {
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_';
}
inline bool is_token_rune(int32_t r) {
    if (r < 0x80) return r >= 0 && is_token_char((uint8_t)r);
    // unicode.IsLetter(r) || unicode.IsDigit(r)  (r == '_' handled above)
    int lo = 0, hi = VL_TOKEN_RANGES_COUNT - 1;
    while (lo <= hi) {
        int mid = (lo + hi) / 2;
        if ((uint32_t)r < VL_TOKEN_RANGES[mid][0]) hi = mid - 1;
        else if ((uint32_t)r > VL_TOKEN_RANGES[mid][1]) lo = mid + 1;
        else return true;
    }
    return false;
}
inline bool is_ascii(sv s) {
    for (unsigned char c : s) if (c >= 0x80) return false;
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// XXH64, seed 0 (github.com/cespare/xxhash/v2 v2.3.0 == reference XXH64; vendor/github.com/cespare/xxhash/v2/xxhash.go)
// ---------------------------------------------------------------------------------------------------------------
static const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P3 = 1609587929392839161ULL,
                      P4 = 9650029242287828579ULL, P5 = 2870177450012600261ULL;
inline uint64_t rol64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }   // little-endian host
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t xxh_round(uint64_t acc, uint64_t in) { acc += in * P2; acc = rol64(acc, 31); return acc * P1; }
inline uint64_t xxh_merge(uint64_t acc, uint64_t v) { v = xxh_round(0, v); acc ^= v; return acc * P1 + P4; }
inline uint64_t xxh64(const void* data, size_t n) {
    const uint8_t* p = (const uint8_t*)data;
    const uint8_t* end = p + n;
    uint64_t h;
    if (n >= 32) {
        uint64_t v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0 - P1;
        do {
            v1 = xxh_round(v1, rd64(p)); v2 = xxh_round(v2, rd64(p + 8));
            v3 = xxh_round(v3, rd64(p + 16)); v4 = xxh_round(v4, rd64(p + 24));
            p += 32;
        } while (p + 32 <= end);
        h = rol64(v1, 1) + rol64(v2, 7) + rol64(v3, 12) + rol64(v4, 18);
        h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
    } else {
        h = P5;
    }
    h += (uint64_t)n;
    while (p + 8 <= end) { h ^= xxh_round(0, rd64(p)); h = rol64(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * P1; h = rol64(h, 23) * P2 + P3; p += 4; }
    while (p < end) { h ^= (uint64_t)(*p) * P5; h = rol64(h, 11) * P1; p++; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}
inline uint64_t xxh64(sv s) { return xxh64(s.data(), s.size()); }

// ---------------------------------------------------------------------------------------------------------------
// Tokenizers: lib/logstorage/tokenizer.go:12-117 (strings), lib/logstorage/hash_tokenizer.go:15-166 (hashes)
// ---------------------------------------------------------------------------------------------------------------
template <class F>
inline void for_each_token(sv s, F&& f) {
    const uint8_t* p = (const uint8_t*)s.data();
    size_t n = s.size();
    if (is_ascii(s)) {
        // tokenizer.go:40-78
        size_t i = 0;
        while (i < n) {
            size_t start = n;
            while (i < n) { if (!is_token_char(p[i])) { i++; continue; } start = i; i++; break; }
            size_t end = n;
            while (i < n) { if (is_token_char(p[i])) { i++; continue; } end = i; i++; break; }
            if (end <= start) break;
            f(s.substr(start, end - start));
        }
        return;
    }
    // tokenizer.go:82-117: `for offset, r := range s` decodes runes Go-style (invalid byte => RuneError, width 1)
    size_t pos = 0;
    while (pos < n) {
        size_t q = pos, tokstart = n;
        while (q < n) { int sz; int32_t r = decode_rune(p + q, n - q, &sz); if (is_token_rune(r)) { tokstart = q; break; } q += sz; }
        pos = tokstart;
        size_t tokend = n;
        q = pos;
        while (q < n) { int sz; int32_t r = decode_rune(p + q, n - q, &sz); if (!is_token_rune(r)) { tokend = q; break; } q += sz; }
        if (tokend == pos) break;
        f(s.substr(pos, tokend - pos));
        pos = tokend;
    }
}

// tokenizeStrings (tokenizer.go:12-24): unique tokens in first-seen order; a string equal to its predecessor is skipped.
inline std::vector<std::string> tokenize_strings(const std::vector<sv>& a) {
    std::vector<std::string> dst;
    std::unordered_set<std::string> seen;
    for (size_t i = 0; i < a.size(); i++) {
        if (i > 0 && a[i] == a[i - 1]) continue;
        for_each_token(a[i], [&](sv tok) {
            std::string t(tok);
            if (seen.insert(t).second) dst.push_back(std::move(t));
        });
    }
    return dst;
}
inline std::vector<std::string> tokenize_string(sv s) { return tokenize_strings(std::vector<sv>{s}); }

// tokenizeHashes (hash_tokenizer.go:15-31): unique xxh64(token) in first-seen order.
struct HashTokenizer {
    std::unordered_set<uint64_t> seen;
    void tokenize(std::vector<uint64_t>& dst, sv s) {
        for_each_token(s, [&](sv tok) { uint64_t h = xxh64(tok); if (seen.insert(h).second) dst.push_back(h); });
    }
};
template <class Vec>
inline std::vector<uint64_t> tokenize_hashes(const Vec& a) {
    std::vector<uint64_t> dst;
    HashTokenizer t;
    for (size_t i = 0; i < a.size(); i++) {
        if (i > 0 && sv(a[i]) == sv(a[i - 1])) continue;
        t.tokenize(dst, sv(a[i]));
    }
    return dst;
}

// ---------------------------------------------------------------------------------------------------------------
// Bloom filter: lib/logstorage/bloomfilter.go:16-191
// ---------------------------------------------------------------------------------------------------------------
static const int bloomFilterHashesCount = 6;
static const int bloomFilterBitsPerItem = 16;

// appendTokensHashes bloomfilter.go:126-144 / appendHashesHashes :152-170
inline void append_hashes_hashes(std::vector<uint64_t>& dst, uint64_t h0) {
    uint64_t hp = h0;
    for (int i = 0; i < bloomFilterHashesCount; i++) {
        uint8_t buf[8];
        memcpy(buf, &hp, 8);   // native-endian (little-endian on the reference's amd64/arm64 targets)
        dst.push_back(xxh64(buf, 8));
        hp++;
    }
}
inline std::vector<uint64_t> tokens_hashes(const std::vector<std::string>& tokens) {
    std::vector<uint64_t> dst;
    for (auto& t : tokens) append_hashes_hashes(dst, xxh64(t));
    return dst;
}

struct BloomFilter {
    std::vector<uint64_t> bits;
    // mustInitHashes :83-89 + initBloomFilter :109-121
    void init_hashes(const std::vector<uint64_t>& hashes) {
        size_t bitsCount = hashes.size() * bloomFilterBitsPerItem;
        size_t words = (bitsCount + 63) / 64;
        bits.assign(words, 0);
        if (words == 0) return;
        std::vector<uint64_t> hh;
        for (uint64_t h : hashes) append_hashes_hashes(hh, h);
        uint64_t maxBits = (uint64_t)words * 64;
        for (uint64_t h : hh) { uint64_t idx = h % maxBits; bits[idx / 64] |= (uint64_t)1 << (idx % 64); }
    }
    // marshal :49-55: big-endian u64 words
    std::string marshal() const {
        std::string out;
        for (uint64_t w : bits) for (int i = 7; i >= 0; i--) out.push_back((char)(w >> (8 * i)));
        return out;
    }
    // unmarshal :58-71
    bool unmarshal(sv src) {
        if (src.size() % 8 != 0) return false;
        bits.assign(src.size() / 8, 0);
        for (size_t i = 0; i < bits.size(); i++) {
            uint64_t w = 0;
            for (int k = 0; k < 8; k++) w = (w << 8) | (uint8_t)src[i * 8 + k];
            bits[i] = w;
        }
        return true;
    }
    // containsAll :173-191
    bool contains_all(const std::vector<uint64_t>& hashes) const {
        if (bits.empty()) return true;
        uint64_t maxBits = (uint64_t)bits.size() * 64;
        for (uint64_t h : hashes) {
            uint64_t idx = h % maxBits;
            if ((bits[idx / 64] & ((uint64_t)1 << (idx % 64))) == 0) return false;
        }
        return true;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Big-endian ints + varuint (vendor/.../VictoriaMetrics/lib/encoding/int.go:12-85, 287-303, 368-386)
// ---------------------------------------------------------------------------------------------------------------
inline void put_be16(std::string& d, uint16_t v) { d.push_back((char)(v >> 8)); d.push_back((char)v); }
inline void put_be32(std::string& d, uint32_t v) { for (int i = 3; i >= 0; i--) d.push_back((char)(v >> (8 * i))); }
inline void put_be64(std::string& d, uint64_t v) { for (int i = 7; i >= 0; i--) d.push_back((char)(v >> (8 * i))); }
inline uint16_t get_be16(const uint8_t* p) { return (uint16_t)((p[0] << 8) | p[1]); }
inline uint32_t get_be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint64_t get_be64(const uint8_t* p) { return ((uint64_t)get_be32(p) << 32) | get_be32(p + 4); }
// MarshalInt64 = zig-zag then big-endian (int.go:69-85)
inline uint64_t zigzag(int64_t v) { return ((uint64_t)v << 1) ^ (uint64_t)(v >> 63); }
inline int64_t unzigzag(uint64_t u) { return (int64_t)(u >> 1) ^ -(int64_t)(u & 1); }
inline void put_varuint(std::string& d, uint64_t v) {
    while (v >= 0x80) { d.push_back((char)(v | 0x80)); v >>= 7; }
    d.push_back((char)v);
}
// returns bytes consumed, 0 on error
inline int get_varuint(const uint8_t* p, size_t n, uint64_t* out) {
    uint64_t v = 0; int shift = 0;
    for (size_t i = 0; i < n && i < 10; i++) {
        v |= (uint64_t)(p[i] & 0x7F) << shift;
        if (p[i] < 0x80) { *out = v; return (int)i + 1; }
        shift += 7;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// ZSTD via the system libzstd.so.1 (same frame format as the reference's vendored libzstd 1.5.7; decompression is
// format-exact, compressed bytes are never compared). Prototypes declared here because the image has no zstd.h.
// ---------------------------------------------------------------------------------------------------------------
extern "C" {
size_t ZSTD_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int level);
size_t ZSTD_decompress(void* dst, size_t dstCapacity, const void* src, size_t compressedSize);
size_t ZSTD_compressBound(size_t srcSize);
unsigned ZSTD_isError(size_t code);
unsigned long long ZSTD_getFrameContentSize(const void* src, size_t srcSize);
}

// ---------------------------------------------------------------------------------------------------------------
// Strings block codec: lib/logstorage/encoding.go:16-50 (marshal), :83-133 (unmarshal), :190-243, :246-336, :343-426
// ---------------------------------------------------------------------------------------------------------------
inline int get_compress_level(size_t n) { return n <= 512 ? 1 : (n <= 4 * 1024 ? 2 : 3); }   // encoding.go:362-370

inline void marshal_bytes_block(std::string& dst, sv src) {   // encoding.go:343-360
    if (src.size() < 128) {
        dst.push_back(0); dst.push_back((char)src.size()); dst.append(src);
        return;
    }
    dst.push_back(1);
    size_t bound = ZSTD_compressBound(src.size());
    std::string tmp(bound, '\0');
    size_t n = ZSTD_compress(tmp.data(), bound, src.data(), src.size(), get_compress_level(src.size()));
    if (ZSTD_isError(n)) throw std::runtime_error("zstd compress failed");
    put_varuint(dst, n);
    dst.append(tmp.data(), n);
}

// returns bytes consumed; throws on malformed input (the reference returns an error -> Panicf FATAL)
inline size_t unmarshal_bytes_block(std::string& dst, const uint8_t* src, size_t n) {   // encoding.go:372-426
    if (n < 1) throw std::runtime_error("cannot unmarshal block type from empty src");
    uint8_t t = src[0];
    if (t == 0) {
        if (n < 2) throw std::runtime_error("cannot unmarshal plain block size from empty src");
        size_t len = src[1];
        if (n - 2 < len) throw std::runtime_error("cannot read plain block");
        dst.append((const char*)src + 2, len);
        return 2 + len;
    } else if (t == 1) {
        uint64_t clen; int ns = get_varuint(src + 1, n - 1, &clen);
        if (ns <= 0) throw std::runtime_error("cannot unmarshal compressed block size");
        if (n - 1 - ns < clen) throw std::runtime_error("cannot read compressed block");
        const uint8_t* frame = src + 1 + ns;
        unsigned long long dlen = ZSTD_getFrameContentSize(frame, clen);
        if (dlen == (unsigned long long)-1 || dlen == (unsigned long long)-2) throw std::runtime_error("bad zstd frame");
        size_t old = dst.size();
        dst.resize(old + dlen);
        size_t got = ZSTD_decompress(dst.data() + old, dlen, frame, clen);
        if (ZSTD_isError(got) || got != dlen) throw std::runtime_error("cannot decompress block");
        return 1 + ns + clen;
    }
    throw std::runtime_error("unexpected block type");
}

enum { uintBlockType8 = 0, uintBlockType16, uintBlockType32, uintBlockType64,
       uintBlockTypeConst8, uintBlockTypeConst16, uintBlockTypeConst32, uintBlockTypeConst64 };

inline bool are_const_u64(const std::vector<uint64_t>& a) {   // encoding.go:135-146
    if (a.empty()) return false;
    for (size_t i = 1; i < a.size(); i++) if (a[i] != a[0]) return false;
    return true;
}

inline std::string marshal_uint64_items(const std::vector<uint64_t>& a) {   // encoding.go:190-243
    std::string dst;
    uint64_t nMax = 0;
    for (uint64_t v : a) nMax = std::max(nMax, v);
    bool consts = a.size() >= 2 && are_const_u64(a);
    if (nMax < (1ULL << 8)) {
        if (consts) { dst.push_back(uintBlockTypeConst8); dst.push_back((char)a[0]); }
        else { dst.push_back(uintBlockType8); for (uint64_t v : a) dst.push_back((char)v); }
    } else if (nMax < (1ULL << 16)) {
        if (consts) { dst.push_back(uintBlockTypeConst16); put_be16(dst, (uint16_t)a[0]); }
        else { dst.push_back(uintBlockType16); for (uint64_t v : a) put_be16(dst, (uint16_t)v); }
    } else if (nMax < (1ULL << 32)) {
        if (consts) { dst.push_back(uintBlockTypeConst32); put_be32(dst, (uint32_t)a[0]); }
        else { dst.push_back(uintBlockType32); for (uint64_t v : a) put_be32(dst, (uint32_t)v); }
    } else {
        if (consts) { dst.push_back(uintBlockTypeConst64); put_be64(dst, a[0]); }
        else { dst.push_back(uintBlockType64); for (uint64_t v : a) put_be64(dst, v); }
    }
    return dst;
}

inline std::vector<uint64_t> unmarshal_uint64_items(sv src, uint64_t itemsCount) {   // encoding.go:246-336
    if (src.size() < 1) throw std::runtime_error("cannot unmarshal uint64 block type from empty src");
    uint8_t bt = (uint8_t)src[0];
    const uint8_t* p = (const uint8_t*)src.data() + 1;
    size_t n = src.size() - 1;
    std::vector<uint64_t> dst(itemsCount);
    auto need = [&](uint64_t want) { if (n != want) throw std::runtime_error("unexpected uint block length"); };
    switch (bt) {
    case uintBlockType8: need(itemsCount); for (uint64_t i = 0; i < itemsCount; i++) dst[i] = p[i]; break;
    case uintBlockType16: need(2 * itemsCount); for (uint64_t i = 0; i < itemsCount; i++) dst[i] = get_be16(p + 2 * i); break;
    case uintBlockType32: need(4 * itemsCount); for (uint64_t i = 0; i < itemsCount; i++) dst[i] = get_be32(p + 4 * i); break;
    case uintBlockType64: need(8 * itemsCount); for (uint64_t i = 0; i < itemsCount; i++) dst[i] = get_be64(p + 8 * i); break;
    case uintBlockTypeConst8: need(1); std::fill(dst.begin(), dst.end(), (uint64_t)p[0]); break;
    case uintBlockTypeConst16: need(2); std::fill(dst.begin(), dst.end(), (uint64_t)get_be16(p)); break;
    case uintBlockTypeConst32: need(4); std::fill(dst.begin(), dst.end(), (uint64_t)get_be32(p)); break;
    case uintBlockTypeConst64: need(8); std::fill(dst.begin(), dst.end(), get_be64(p)); break;
    default: throw std::runtime_error("unexpected uint64 block type");
    }
    return dst;
}

template <class Vec>
inline bool are_const_values(const Vec& a) {   // block.go areConstValues semantics: len>=1 and all equal
    if (a.size() == 0) return false;
    for (size_t i = 1; i < a.size(); i++) if (sv(a[i]) != sv(a[0])) return false;
    return true;
}

// marshalStringsBlock encoding.go:16-50
template <class Vec>
inline std::string marshal_strings_block(const Vec& a) {
    std::vector<uint64_t> lens(a.size());
    size_t total = 0;
    for (size_t i = 0; i < a.size(); i++) { lens[i] = sv(a[i]).size(); total += lens[i]; }
    std::string dst;
    marshal_bytes_block(dst, marshal_uint64_items(lens));
    if (are_const_values(a)) {
        marshal_bytes_block(dst, sv(a[0]));
    } else {
        std::string b; b.reserve(total);
        for (size_t i = 0; i < a.size(); i++) b.append(sv(a[i]));
        marshal_bytes_block(dst, b);
    }
    return dst;
}

// The "post-ZSTD" stage of a values block: the two decoded bytes blocks (lens items incl. the type byte, and data).
struct DecodedStringsBlock {
    std::string lens_items;   // type byte + items  (what unmarshalUint64Items consumes)
    std::string data;         // what unmarshalBytesBlock yields for the strings bytes
};
inline DecodedStringsBlock decode_values_block_stage(sv src) {
    DecodedStringsBlock d;
    size_t c = unmarshal_bytes_block(d.lens_items, (const uint8_t*)src.data(), src.size());
    size_t c2 = unmarshal_bytes_block(d.data, (const uint8_t*)src.data() + c, src.size() - c);
    if (c + c2 != src.size()) throw std::runtime_error("unexpected non-empty tail after reading bytes block with strings");
    return d;
}

// stringsBlockUnmarshaler.unmarshal encoding.go:83-133; values are views into `storage`
inline std::vector<sv> unmarshal_strings(const DecodedStringsBlock& d, uint64_t itemsCount) {
    std::vector<uint64_t> lens = unmarshal_uint64_items(d.lens_items, itemsCount);
    std::vector<sv> out(itemsCount);
    sv data(d.data);
    if (lens.size() >= 2 && are_const_u64(lens) && data.size() == lens[0]) {
        for (auto& o : out) o = data;
        return out;
    }
    size_t off = 0;
    for (uint64_t i = 0; i < itemsCount; i++) {
        if (data.size() - off < lens[i]) throw std::runtime_error("cannot unmarshal a string: not enough data");
        out[i] = data.substr(off, lens[i]);
        off += lens[i];
    }
    return out;
}

// ---------------------------------------------------------------------------------------------------------------
// Number / IP / timestamp parsing & formatting: lib/logstorage/values_encoder.go
// ---------------------------------------------------------------------------------------------------------------
// tryParseUint64 :553-585
inline bool try_parse_uint64(sv s, uint64_t* out) {
    if (s.empty() || s.size() > strlen("18_446_744_073_709_551_615")) return false;
    if (s.size() > 1 && s[0] == '0') return false;
    uint64_t n = 0;
    for (char ch : s) {
        if (ch == '_') continue;
        if (ch < '0' || ch > '9') return false;
        if (n > UINT64_MAX / 10) return false;
        n *= 10;
        uint64_t d = (uint64_t)(ch - '0');
        uint64_t n1 = n + d;
        if (n1 < n) return false;
        n = n1;
    }
    *out = n;
    return true;
}
// tryParseDateUint64 :588-619 (note the unchecked second digit in the 2-char fast path: byte arithmetic wraps)
inline bool try_parse_date_uint64(sv s, uint64_t* out) {
    if (s.empty() || s.size() > 9) return false;
    if (s.size() == 2) {
        if (s[0] < '0' || s[0] > '9') return false;
        *out = 10 * (uint64_t)(uint8_t)(s[0] - '0') + (uint64_t)(uint8_t)((uint8_t)s[1] - (uint8_t)'0');
        return true;
    }
    uint64_t n = 0;
    for (char ch : s) {
        if (ch < '0' || ch > '9') return false;
        n = n * 10 + (uint64_t)(ch - '0');   // cannot overflow for <= 9 digits
    }
    *out = n;
    return true;
}
// tryParseInt64 :622-645
inline bool try_parse_int64(sv s, int64_t* out) {
    if (s.empty()) return false;
    bool minus = s[0] == '-';
    if (minus) s.remove_prefix(1);
    uint64_t n;
    if (!try_parse_uint64(s, &n)) return false;
    if (n >= (1ULL << 63)) {
        if (minus && n == (1ULL << 63)) { *out = INT64_MIN; return true; }
        return false;
    }
    int64_t ni = (int64_t)n;
    *out = minus ? -ni : ni;
    return true;
}
// tryParseIPv4 :675-730
inline bool try_parse_ipv4(sv s, uint32_t* out) {
    if (s.size() < strlen("1.1.1.1") || s.size() > strlen("255.255.255.255") || std::count(s.begin(), s.end(), '.') != 3) return false;
    uint8_t oct[4];
    for (int k = 0; k < 3; k++) {
        size_t n = s.find('.');
        if (n == sv::npos || n == 0 || n > 3) return false;
        uint64_t v;
        if (!try_parse_date_uint64(s.substr(0, n), &v) || v > 255) return false;
        oct[k] = (uint8_t)v;
        s.remove_prefix(n + 1);
    }
    uint64_t v;
    if (!try_parse_date_uint64(s, &v) || v > 255) return false;
    oct[3] = (uint8_t)v;
    *out = get_be32(oct);
    return true;
}
// math.Pow10 for the exponents used by tryParseFloat64Internal (n in [-27, 0]); Go's table is exact literals 1e-N.
inline double go_pow10(int n) {
    static const double neg[] = {1e0, 1e-1, 1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8, 1e-9, 1e-10, 1e-11, 1e-12, 1e-13, 1e-14,
                                 1e-15, 1e-16, 1e-17, 1e-18, 1e-19, 1e-20, 1e-21, 1e-22, 1e-23, 1e-24, 1e-25, 1e-26, 1e-27,
                                 1e-28, 1e-29, 1e-30, 1e-31};
    if (n <= 0 && n >= -31) return neg[-n];
    return std::pow(10.0, n);
}
// tryParseFloat64Internal :788-850 (isExact=true everywhere on this path)
inline bool try_parse_float64_exact(sv s, double* out) {
    if (s.empty() || s.size() > strlen("-18_446_744_073_709_551_615")) return false;
    bool minus = s[0] == '-';
    if (minus) s.remove_prefix(1);
    size_t n = s.find('.');
    if (n == sv::npos) {
        uint64_t v;
        if (!try_parse_uint64(s, &v)) return false;
        if (v >= (1ULL << 53)) return false;
        double f = (double)v;
        *out = minus ? -f : f;
        return true;
    }
    if (n == 0 || n == s.size() - 1) return false;
    sv sInt = s.substr(0, n), sFrac = s.substr(n + 1);
    uint64_t nInt;
    if (!try_parse_uint64(sInt, &nInt)) return false;
    size_t k = 0;
    while (k + 1 < sFrac.size() && sFrac[k] == '0') k++;
    uint64_t nFrac;
    if (!try_parse_uint64(sFrac.substr(k), &nFrac)) return false;
    int underscores = (int)std::count(sFrac.begin(), sFrac.end(), '_');
    double p10 = go_pow10(underscores - (int)sFrac.size());
    double f = std::fma((double)nFrac, p10, (double)nInt);
    *out = minus ? -f : f;
    return true;
}

// Go time.Date(...).Unix() with normalisation of out-of-range fields (time.go Date); UTC only.
inline int64_t go_date_unix(int64_t year, int64_t month, int64_t day, int64_t hour, int64_t min, int64_t sec) {
    auto norm = [](int64_t& hi, int64_t& lo, int64_t base) {
        if (lo < 0) { int64_t n = (-lo - 1) / base + 1; hi -= n; lo += n * base; }
        if (lo >= base) { int64_t n = lo / base; hi += n; lo -= n * base; }
    };
    int64_t m = month - 1;
    norm(year, m, 12);
    norm(min, sec, 60);
    norm(hour, min, 60);
    norm(day, hour, 24);
    // days since 1970-01-01 of year-(m+1)-1, then + day - 1 (day may be out of range; it simply adds)
    auto days_from_civil = [](int64_t y, unsigned mth, unsigned d) -> int64_t {
        y -= mth <= 2;
        const int64_t era = (y >= 0 ? y : y - 399) / 400;
        const unsigned yoe = (unsigned)(y - era * 400);
        const unsigned doy = (153 * (mth + (mth > 2 ? -3 : 9)) + 2) / 5 + d - 1;
        const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
        return era * 146097 + (int64_t)doe - 719468;
    };
    int64_t days = days_from_civil(year, (unsigned)(m + 1), 1) + (day - 1);
    return days * 86400 + hour * 3600 + min * 60 + sec;
}
// tryParseTimestampSecs :466-551
inline bool try_parse_timestamp_secs(sv& s, int64_t* secs) {
    uint64_t n;
    if (s[4] != '-') return false;
    if (!try_parse_date_uint64(s.substr(0, 4), &n) || n < 1677 || n > 2262) return false;
    int64_t year = (int64_t)n; s.remove_prefix(5);
    if (s[2] != '-') return false;
    if (!try_parse_date_uint64(s.substr(0, 2), &n)) return false;
    int64_t month = (int64_t)n; s.remove_prefix(3);
    if (s[2] != 'T' && s[2] != ' ') return false;
    if (!try_parse_date_uint64(s.substr(0, 2), &n)) return false;
    int64_t day = (int64_t)n; s.remove_prefix(3);
    if (s[2] != ':') return false;
    if (!try_parse_date_uint64(s.substr(0, 2), &n)) return false;
    int64_t hour = (int64_t)n; s.remove_prefix(3);
    if (s[2] != ':') return false;
    if (!try_parse_date_uint64(s.substr(0, 2), &n)) return false;
    int64_t minute = (int64_t)n; s.remove_prefix(3);
    if (!try_parse_date_uint64(s.substr(0, 2), &n)) return false;
    int64_t second = (int64_t)n; s.remove_prefix(2);
    int64_t v = go_date_unix(year, month, day, hour, minute, second);
    // Go: `secs < int64(-1<<63)/1e9 || secs >= int64((1<<63)-1)/1e9` -- typed int64 constant division (truncating)
    if (v < -9223372036LL || v >= 9223372036LL) return false;
    *secs = v;
    return true;
}
// tryParseTimestampISO8601 :428-464
inline bool try_parse_timestamp_iso8601(sv s, int64_t* out) {
    if (s.size() != strlen("2006-01-02T15:04:05.000Z")) return false;
    int64_t secs;
    if (!try_parse_timestamp_secs(s, &secs)) return false;
    int64_t nsecs = secs * 1000000000LL;
    if (s[0] != '.') return false;
    s.remove_prefix(1);
    if (s[3] != 'Z') return false;
    uint64_t ms;
    if (!try_parse_date_uint64(s.substr(0, 3), &ms)) return false;
    *out = nsecs + (int64_t)ms * 1000000LL;
    return true;
}

// marshalUint8String :1367-1385 and friends
inline void marshal_uint64_string(std::string& dst, uint64_t n) { char b[24]; auto r = std::to_chars(b, b + 24, n); dst.append(b, r.ptr); }
inline void marshal_int64_string(std::string& dst, int64_t n) { char b[24]; auto r = std::to_chars(b, b + 24, n); dst.append(b, r.ptr); }
inline void marshal_ipv4_string(std::string& dst, uint32_t n) {
    marshal_uint64_string(dst, n >> 24); dst.push_back('.'); marshal_uint64_string(dst, (n >> 16) & 0xFF); dst.push_back('.');
    marshal_uint64_string(dst, (n >> 8) & 0xFF); dst.push_back('.'); marshal_uint64_string(dst, n & 0xFF);
}
// strconv.AppendFloat(dst, f, 'f', -1, 64): the shortest round-trip DIGITS (strconv ftoa.go: shortest → %e digits), laid out
// by fmtF in fixed notation: integer digits beyond the shortest ones are '0' (1.7976931348623157e308 prints as
// 17976931348623157 followed by 292 zeros, not as the exact binary value, which is what to_chars(fixed) would print).
inline void marshal_float64_string(std::string& dst, double f) {
    if (std::isnan(f)) { dst.append("NaN"); return; }
    if (std::isinf(f)) { dst.append(f > 0 ? "+Inf" : "-Inf"); return; }
    if (std::signbit(f)) { dst.push_back('-'); f = -f; }
    if (f == 0) { dst.push_back('0'); return; }
    char b[64];
    auto r = std::to_chars(b, b + sizeof(b), f, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest digits
    std::string digs; int e10 = 0; char* p = b;
    for (; p < r.ptr && *p != 'e'; p++) if (*p != '.') digs.push_back(*p);
    e10 = atoi(std::string(p + 1, r.ptr).c_str());
    int point = e10 + 1;   // number of digits before the decimal point
    if (point <= 0) { dst.append("0."); dst.append((size_t)-point, '0'); dst.append(digs); return; }
    if ((size_t)point >= digs.size()) { dst.append(digs); dst.append((size_t)point - digs.size(), '0'); return; }
    dst.append(digs, 0, (size_t)point); dst.push_back('.'); dst.append(digs, (size_t)point, std::string::npos);
}
// marshalTimestampISO8601String :1414-1418: time.Unix(0,nsecs).UTC().AppendFormat("2006-01-02T15:04:05.000Z")
inline void marshal_timestamp_iso8601_string(std::string& dst, int64_t nsecs) {
    int64_t secs = nsecs / 1000000000LL, rem = nsecs % 1000000000LL;
    if (rem < 0) { rem += 1000000000LL; secs -= 1; }
    int64_t days = secs / 86400, sod = secs % 86400;
    if (sod < 0) { sod += 86400; days -= 1; }
    // civil_from_days
    int64_t z = days + 719468;
    const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
    const unsigned doe = (unsigned)(z - era * 146097);
    const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    int64_t y = (int64_t)yoe + era * 400;
    const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const unsigned mp = (5 * doy + 2) / 153;
    const unsigned d = doy - (153 * mp + 2) / 5 + 1;
    const unsigned m = mp < 10 ? mp + 3 : mp - 9;
    y += (m <= 2);
    char b[64];
    // Go's "2006" prints at least 4 digits (zero padded); years here are within 1677..2262
    snprintf(b, sizeof b, "%04lld-%02u-%02uT%02lld:%02lld:%02lld.%03lldZ", (long long)y, m, d, (long long)(sod / 3600),
             (long long)((sod / 60) % 60), (long long)(sod % 60), (long long)(rem / 1000000LL));   // .000 truncates
    dst.append(b);
}

}  // namespace vlo
