// Host+device predicates of the filter kinds that are next in line for the device engine: i(phrase), i(prefix*),
// seq(...), contains_all(...), contains_any(...).  Same shape as the kinds already wired in (vl_hd.cuh: match_phrase, range_predicate):
// one value in, one bit out, no allocation.  Until the row kernels call them they are reachable only through vlscan_eval_predicate,
// which is how tests/test_abi_cpu.py checks them against the oracle on the CPU.
//
//   matchAnyCasePhrase / matchAnyCasePrefix   lib/logstorage/filter_any_case_phrase.go:159-191, filter_any_case_prefix.go:160-182
//   stringsutil.AppendLowercase               vm/lib/stringsutil/stringsutil.go:26-51 (rune by rune; an invalid byte becomes U+FFFD)
//   matchSequence                             lib/logstorage/filter_sequence.go:201-213
//   matchAllPhrases / matchAnyPhrase          lib/logstorage/filter_contains_all.go:309-321, filter_contains_any.go
// A translation unit that only needs the host builds (vl_engine.cu today) defines VL_ANYCASE_HOST_ONLY before including this file: no
// device table, no device functions, and its device code stays byte for byte what it was.
#pragma once
#include "vl_hd.cuh"

#if defined(__CUDACC__) && !defined(VL_ANYCASE_HOST_ONLY)
#define VLA_HD __host__ __device__ __forceinline__
#define VLA_HDN static __host__ __device__
#define VLA_DEVICE_TABLE 1
#else
#define VLA_HD inline
#define VLA_HDN inline
#endif

namespace vl {

#include "unicode_case.inc"
static const unsigned int H_TOLOWER[VL_TOLOWER_COUNT][2] = { VL_TOLOWER_INIT };
#ifdef VLA_DEVICE_TABLE
static __device__ const unsigned int D_TOLOWER[VL_TOLOWER_COUNT][2] = { VL_TOLOWER_INIT };   // 11 KB, touched for non-ASCII runes only
#endif

// unicode.ToLower
VLA_HD int32_t to_lower_rune(int32_t r) {
    if (r < 0x80) return (uint32_t)(r - 'A') < 26u ? r + 32 : r;
    int lo = 0, hi = VL_TOLOWER_COUNT - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
#if defined(__CUDA_ARCH__) && defined(VLA_DEVICE_TABLE)
        const unsigned a = D_TOLOWER[mid][0], b = D_TOLOWER[mid][1];
#else
        const unsigned a = H_TOLOWER[mid][0], b = H_TOLOWER[mid][1];
#endif
        if ((uint32_t)r < a) hi = mid - 1; else if ((uint32_t)r > a) lo = mid + 1; else return (int32_t)b;
    }
    return r;
}
// unicode.ToUpper (host only: the program compiler upper-cases i(...) needles for iso8601 columns)
static const unsigned int H_TOUPPER[VL_TOUPPER_COUNT][2] = { VL_TOUPPER_INIT };
inline int32_t to_upper_rune_host(int32_t r) {
    if (r < 0x80) return (uint32_t)(r - 'a') < 26u ? r - 32 : r;
    int lo = 0, hi = VL_TOUPPER_COUNT - 1;
    while (lo <= hi) { const int mid = (lo + hi) >> 1; const unsigned a = H_TOUPPER[mid][0]; if ((uint32_t)r < a) hi = mid - 1; else if ((uint32_t)r > a) lo = mid + 1; else return (int32_t)H_TOUPPER[mid][1]; }
    return r;
}
// utf8.AppendRune into out[4]; returns the length (surrogates and values above U+10FFFF are written as U+FFFD)
VLA_HD int encode_rune(uint8_t* out, int32_t r) {
    uint32_t c = (uint32_t)r;
    if (c < 0x80) { out[0] = (uint8_t)c; return 1; }
    if (c < 0x800) { out[0] = (uint8_t)(0xC0 | (c >> 6)); out[1] = (uint8_t)(0x80 | (c & 0x3F)); return 2; }
    if (c > 0x10FFFF || (c >= 0xD800 && c <= 0xDFFF)) c = 0xFFFD;
    if (c < 0x10000) { out[0] = (uint8_t)(0xE0 | (c >> 12)); out[1] = (uint8_t)(0x80 | ((c >> 6) & 0x3F)); out[2] = (uint8_t)(0x80 | (c & 0x3F)); return 3; }
    out[0] = (uint8_t)(0xF0 | (c >> 18)); out[1] = (uint8_t)(0x80 | ((c >> 12) & 0x3F)); out[2] = (uint8_t)(0x80 | ((c >> 6) & 0x3F)); out[3] = (uint8_t)(0x80 | (c & 0x3F));
    return 4;
}
VLA_HD bool is_ascii_lowercase(const uint8_t* s, uint32_t n) {   // isASCIILowercase filter_any_case_phrase.go:183-191
    for (uint32_t i = 0; i < n; i++) if (s[i] >= 0x80 || (uint32_t)(s[i] - 'A') < 26u) return false;
    return true;
}

// matchPhrase (prefix = false) / matchPrefix (prefix = true) of the LOWERCASED value against an already lowercased needle, without ever
// materialising the lowercased value: it is walked rune by rune, each rune lowercased and re-encoded on the fly.  The lowercased value
// is valid UTF-8 and so is the needle (strings.ToLower), so occurrences start on rune boundaries; the runes in front of and behind an
// occurrence are the lowercased neighbours in the walk, which is what DecodeLastRune / DecodeRune see in the reference's buffer.
VLA_HDN bool any_case_match(const uint8_t* s, uint32_t n, const uint8_t* nd, uint32_t m, bool prefix) {
    if (m == 0) return prefix ? n > 0 : n == 0;
    if (m > n) return false;                       // byte lengths BEFORE lowercasing, like the reference
    if (is_ascii_lowercase(s, n)) return prefix ? match_prefix(s, n, nd, m) : match_phrase(s, n, nd, m);
    const bool st = needle_starts_with_token(nd, m), en = !prefix && needle_ends_with_token(nd, m);
    int32_t prev = -1;                             // lowercased rune in front of the candidate, -1 at the start of the value
    uint32_t i = 0;
    while (i < n) {
        // does the lowercased value continue with the needle from source offset i on?
        uint32_t j = i, k = 0; bool ok = true;
        while (k < m) {
            if (j >= n) { ok = false; break; }
            int w; const int32_t r = to_lower_rune(decode_rune(s + j, n - j, &w));
            uint8_t enc[4]; const int e = encode_rune(enc, r);
            if (k + (uint32_t)e > m) { ok = false; break; }
            for (int t = 0; t < e; t++) if (enc[t] != nd[k + t]) { ok = false; break; }
            if (!ok) break;
            k += (uint32_t)e; j += (uint32_t)w;
        }
        if (ok) {
            bool fine = !(st && prev >= 0 && (prev == kRuneError || is_token_rune(prev)));
            if (fine && en && j < n) { int w; const int32_t r = to_lower_rune(decode_rune(s + j, n - j, &w)); if (r == kRuneError || is_token_rune(r)) fine = false; }
            if (fine) return true;
        }
        int w; prev = to_lower_rune(decode_rune(s + i, n - i, &w)); i += (uint32_t)w;
    }
    return false;
}

// A phrase list as the row kernels will get it: count, then count x (varuint length, bytes)
struct PhraseList {
    const uint8_t* p; uint32_t n;
    VLA_HD bool next(const uint8_t** ph, uint32_t* len) {
        if (n == 0) return false;
        uint32_t v = 0; int sh = 0;
        for (;;) { if (n == 0 || sh > 28) return false; const uint8_t b = *p++; n--; v |= (uint32_t)(b & 0x7F) << sh; if (b < 0x80) break; sh += 7; }
        if (v > n) return false;
        *ph = p; *len = v; p += v; n -= v;
        return true;
    }
};
// matchSequence: every phrase must occur, each one behind the previous occurrence (empty phrases are skipped by getPhrasePos: position 0)
VLA_HDN bool match_sequence(const uint8_t* s, uint32_t n, PhraseList L) {
    const uint8_t* ph; uint32_t m;
    while (L.next(&ph, &m)) {
        if (m == 0) continue;                                              // getPhrasePos(s, "") == 0
        const bool st = needle_starts_with_token(ph, m), en = needle_ends_with_token(ph, m);
        uint32_t pos = 0; int at = -1;
        for (;;) {
            const int k = find_bytes(s, n, ph, m, pos);
            if (k < 0) break;
            if (phrase_boundaries_ok(s, n, (uint32_t)k, m, st, en)) { at = k; break; }
            pos = (uint32_t)k + 1;
        }
        if (at < 0) return false;
        s += (uint32_t)at + m; n -= (uint32_t)at + m;
    }
    return true;
}
// matchAllPhrases: every non-empty phrase matches;   matchAnyPhrase: some phrase matches ("" matches only the empty value)
VLA_HDN bool match_all_phrases(const uint8_t* s, uint32_t n, PhraseList L) {
    const uint8_t* ph; uint32_t m;
    while (L.next(&ph, &m)) { if (m == 0) continue; if (!match_phrase(s, n, ph, m)) return false; }
    return true;
}
VLA_HDN bool match_any_phrase(const uint8_t* s, uint32_t n, PhraseList L) {
    const uint8_t* ph; uint32_t m;
    while (L.next(&ph, &m)) if (match_phrase(s, n, ph, m)) return true;
    return false;
}

}  // namespace vl
