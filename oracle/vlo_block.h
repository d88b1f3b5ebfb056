// ORACLE -- TEST INFRASTRUCTURE ONLY (see vlo_util.h header).
//
// Block builder (rows -> encoded column blocks + bloom filters), block search and the filter tree, restating
//   lib/logstorage/block.go:107-175,202-320 (const-column detection, column.mustWriteTo)
//   lib/logstorage/values_encoder.go:109-154,1141-1322 (valuesEncoder.encode order: dict,uint,int,float,ipv4,iso8601,string)
//   lib/logstorage/block_search.go:207-226,232-324,411-474 (blockSearch.search / header + bloom + values access)
//   lib/logstorage/bitmap.go:28-191
//   lib/logstorage/filter_{phrase,prefix,exact,in,regexp,and,or,not,noop}.go, lib/logstorage/in_values.go
//   and, ahead of the product: filter_{exact_prefix,sequence,contains_all,contains_any,len_range,string_range,
//   ipv4_range,any_case_phrase,any_case_prefix,value_type}.go
#pragma once
#include "vlo_util.h"
#include "vlo_mathnum.h"
#include "vlo_timestamps.h"
#include "vlo_regex.h"
#include <functional>
#include <map>
#include <memory>

namespace vlo {

enum ValueType : uint8_t { VT_UNKNOWN = 0, VT_STRING = 1, VT_DICT = 2, VT_UINT8 = 3, VT_UINT16 = 4, VT_UINT32 = 5, VT_UINT64 = 6,
                           VT_FLOAT64 = 7, VT_IPV4 = 8, VT_ISO8601 = 9, VT_INT64 = 10 };
static const size_t maxDictSizeBytes = 256, maxDictLen = 8, maxConstColumnValueSize = 256;

struct Column {
    std::string name;
    uint8_t valueType = VT_STRING;
    uint64_t minValue = 0, maxValue = 0;
    std::vector<std::string> dict;
    std::string valuesBlock;   // marshalStringsBlock(encoded values) == bytes at [valuesOffset, valuesOffset+valuesSize)
    std::string bloom;         // bytes at [bloomFilterOffset, +bloomFilterSize): big-endian u64 words
    // bench only ("post-zstd" CPU baseline): the two bytes blocks of valuesBlock already decompressed, so that the timed scan
    // starts at the same input stage as the device-resident scan.  Never set by the parity tests.
    std::shared_ptr<const DecodedStringsBlock> predecoded;
};
struct ConstColumn { std::string name, value; };
struct Block {
    uint64_t rows = 0;
    std::vector<Column> columns;
    std::vector<ConstColumn> consts;
    // timestamps column (optional in this oracle): encoded block + timestampsHeader fields (block_header.go timestampsHeader)
    bool hasTimestamps = false; EncodedInt64s ts; int64_t minTimestamp = 0, maxTimestamp = 0;
    void set_timestamps(const std::vector<int64_t>& a) {   // block.go:674-690: rows are sorted by time
        if (a.size() != rows) throw std::runtime_error("timestamps count differs from rows");
        for (size_t i = 1; i < a.size(); i++) if (a[i] < a[i - 1]) throw std::runtime_error("timestamps must be sorted");
        ts = marshal_int64_array(a); hasTimestamps = true; minTimestamp = a.front(); maxTimestamp = a.back();
    }
};

inline std::string canonical(sv name) { return name.empty() ? std::string("_msg") : std::string(name); }   // getCanonicalColumnName

// ---- values encoder ---------------------------------------------------------------------------------------------
struct EncodedValues { uint8_t vt; uint64_t minv = 0, maxv = 0; std::vector<std::string> values; std::vector<std::string> dict; };

inline EncodedValues encode_values(const std::vector<sv>& values) {   // values_encoder.go:109-154
    EncodedValues e; e.vt = VT_STRING;
    if (values.empty()) return e;
    // tryDictEncoding :1224-1241 + valuesDict.getOrAdd :1269-1288
    {
        std::vector<std::string> dict; bool ok = true;
        std::vector<std::string> enc; enc.reserve(values.size());
        for (sv v : values) {
            if (v.size() > maxDictSizeBytes) { ok = false; break; }
            size_t sz = 0; int id = -1;
            for (size_t i = 0; i < dict.size(); i++) { if (dict[i] == v) { id = (int)i; break; } sz += dict[i].size(); }
            if (id < 0) {
                if (dict.size() >= maxDictLen || sz + v.size() > maxDictSizeBytes) { ok = false; break; }
                dict.emplace_back(v); id = (int)dict.size() - 1;
            }
            enc.emplace_back(1, (char)id);
        }
        if (ok) { e.vt = VT_DICT; e.values = std::move(enc); e.dict = std::move(dict); return e; }
    }
    auto all = [&](auto parse, auto& arr) { arr.resize(values.size()); for (size_t i = 0; i < values.size(); i++) if (!parse(values[i], &arr[i])) return false; return true; };
    {   // tryUintEncoding :1168-1222
        std::vector<uint64_t> a;
        if (all(try_parse_uint64, a)) {
            uint64_t mn = a[0], mx = a[0];
            for (uint64_t v : a) { mn = std::min(mn, v); mx = std::max(mx, v); }
            int bits = mx ? 64 - __builtin_clzll(mx) : 0;
            e.minv = mn; e.maxv = mx;
            for (uint64_t v : a) {
                std::string s;
                if (bits <= 8) s.push_back((char)v); else if (bits <= 16) put_be16(s, (uint16_t)v); else if (bits <= 32) put_be32(s, (uint32_t)v); else put_be64(s, v);
                e.values.push_back(std::move(s));
            }
            e.vt = bits <= 8 ? VT_UINT8 : bits <= 16 ? VT_UINT16 : bits <= 32 ? VT_UINT32 : VT_UINT64;
            return e;
        }
    }
    {   // tryIntEncoding :1141-1166
        std::vector<int64_t> a;
        if (all(try_parse_int64, a)) {
            int64_t mn = a[0], mx = a[0];
            for (int64_t v : a) { mn = std::min(mn, v); mx = std::max(mx, v); }
            e.minv = (uint64_t)mn; e.maxv = (uint64_t)mx;
            for (int64_t v : a) { std::string s; put_be64(s, zigzag(v)); e.values.push_back(std::move(s)); }
            e.vt = VT_INT64; return e;
        }
    }
    {   // tryFloat64Encoding :732-760
        std::vector<double> a;
        if (all(try_parse_float64_exact, a)) {
            double mn = a[0], mx = a[0];
            for (size_t i = 0; i < a.size(); i++) { if (i == 0 || a[i] < mn) mn = a[i]; if (i == 0 || a[i] > mx) mx = a[i]; }
            memcpy(&e.minv, &mn, 8); memcpy(&e.maxv, &mx, 8);
            for (double v : a) { uint64_t b; memcpy(&b, &v, 8); std::string s; put_be64(s, b); e.values.push_back(std::move(s)); }
            e.vt = VT_FLOAT64; return e;
        }
    }
    {   // tryIPv4Encoding :647-672
        std::vector<uint32_t> a;
        if (all(try_parse_ipv4, a)) {
            uint32_t mn = a[0], mx = a[0];
            for (uint32_t v : a) { mn = std::min(mn, v); mx = std::max(mx, v); }
            e.minv = mn; e.maxv = mx;
            for (uint32_t v : a) { std::string s; put_be32(s, v); e.values.push_back(std::move(s)); }
            e.vt = VT_IPV4; return e;
        }
    }
    {   // tryTimestampISO8601Encoding :308-333
        std::vector<int64_t> a;
        if (all(try_parse_timestamp_iso8601, a)) {
            int64_t mn = a[0], mx = a[0];
            for (int64_t v : a) { mn = std::min(mn, v); mx = std::max(mx, v); }
            e.minv = (uint64_t)mn; e.maxv = (uint64_t)mx;
            for (int64_t v : a) { std::string s; put_be64(s, (uint64_t)v); e.values.push_back(std::move(s)); }
            e.vt = VT_ISO8601; return e;
        }
    }
    e.vt = VT_STRING;
    for (sv v : values) e.values.emplace_back(v);
    return e;
}

// column.mustWriteTo block.go:134-175
inline Column build_column(sv name, const std::vector<sv>& values) {
    Column c; c.name = canonical(name);
    EncodedValues e = encode_values(values);
    c.valueType = e.vt; c.minValue = e.minv; c.maxValue = e.maxv; c.dict = e.dict;
    c.valuesBlock = marshal_strings_block(e.values);
    if (c.valueType != VT_DICT) {
        BloomFilter bf; bf.init_hashes(tokenize_hashes(values));
        c.bloom = bf.marshal();
    }
    return c;
}

// block.mustInitFromRows fast path (all rows have the same fields) block.go:232-253 + const detection :107-122.
// Columns are given column-major: names[i], values[i][row]. Empty values mean "field missing in this row".
inline Block build_block(const std::vector<std::string>& names, const std::vector<std::vector<sv>>& cols, uint64_t rows) {
    Block b; b.rows = rows;
    for (size_t i = 0; i < names.size(); i++) {
        const auto& v = cols[i];
        bool isconst = true;
        if (!v.empty()) {
            if (v[0].size() > maxConstColumnValueSize) isconst = false;
            else for (size_t j = 1; j < v.size(); j++) if (v[j] != v[0]) { isconst = false; break; }
        }
        if (isconst) {
            // a const column with an empty value is equivalent to a missing column (getConstColumnValue returns "")
            if (!v.empty() && !v[0].empty()) b.consts.push_back({canonical(names[i]), std::string(v[0])});
            continue;
        }
        b.columns.push_back(build_column(names[i], v));
    }
    return b;
}

// ---- bitmap bitmap.go:28-191 ---------------------------------------------------------------------------------------
struct Bitmap {
    std::vector<uint64_t> a; uint64_t bitsLen = 0;
    void init(uint64_t n) { bitsLen = n; a.assign((n + 63) / 64, 0); }
    void set_bits() {   // :62-72
        for (auto& w : a) w = ~0ULL;
        uint64_t tail = a.size() * 64 - bitsLen;
        if (tail > 0) a.back() &= (~0ULL) >> tail;
    }
    void reset_bits() { for (auto& w : a) w = 0; }
    bool is_zero() const { for (auto w : a) if (w) return false; return true; }
    void and_not(const Bitmap& x) { for (size_t i = 0; i < a.size(); i++) a[i] &= ~x.a[i]; }
    uint64_t ones() const { uint64_t n = 0; for (auto w : a) n += __builtin_popcountll(w); return n; }
    bool are_all_bits_set() const {   // :83-97
        for (size_t i = 0; i < a.size(); i++) {
            if (a[i] == ~0ULL) continue;
            if (i + 1 < a.size()) return false;
            uint64_t tail = bitsLen % 64;
            if (tail == 0 || a[i] != (1ULL << tail) - 1) return false;
        }
        return true;
    }
    void set_bit(uint64_t i) { a[i / 64] |= 1ULL << (i % 64); }               // :113-118
    bool is_set_bit(uint64_t i) const { return (a[i / 64] >> (i % 64)) & 1; }   // :120-125
    template <class F> void for_each_set_bit_readonly(F&& f) const {           // :156-183
        for (size_t i = 0; i < a.size(); i++) {
            uint64_t w = a[i];
            for (int j = 0; w && j < 64; j++) {
                if (!(w >> j & 1)) continue;
                uint64_t idx = i * 64 + j;
                if (idx >= bitsLen) break;
                f(idx);
            }
        }
    }
    template <class F> void for_each_set_bit(F&& f) {   // :128-153: f returns whether to keep the bit
        for (size_t i = 0; i < a.size(); i++) {
            uint64_t w = a[i]; if (!w) continue;
            uint64_t keep = w;
            for (int j = 0; j < 64; j++) {
                if (!(w >> j & 1)) continue;
                uint64_t idx = i * 64 + j;
                if (idx >= bitsLen) break;
                if (!f(idx)) keep &= ~(1ULL << j);
            }
            a[i] = keep;
        }
    }
};

// ---- predicates ----------------------------------------------------------------------------------------------------
// getPhrasePos / matchPhrase filter_phrase.go:211-270
inline int64_t get_phrase_pos(sv s, sv phrase) {   // :220-270; -1 when the phrase is not found
    if (phrase.empty()) return 0;
    if (phrase.size() > s.size()) return -1;
    const uint8_t* sp = (const uint8_t*)s.data();
    int sz;
    int32_t r = (uint8_t)phrase[0];
    if (r >= 0x80) r = decode_rune((const uint8_t*)phrase.data(), phrase.size(), &sz);
    bool startsWithToken = is_token_rune(r);
    r = (uint8_t)phrase.back();
    if (r >= 0x80) r = decode_last_rune((const uint8_t*)phrase.data(), phrase.size(), &sz);
    bool endsWithToken = is_token_rune(r);
    size_t pos = 0;
    for (;;) {
        size_t n = s.find(phrase, pos);
        if (n == sv::npos) return -1;
        pos = n;
        if (startsWithToken && pos > 0) {
            int32_t q = sp[pos - 1];
            if (q >= 0x80) q = decode_last_rune(sp, pos, &sz);
            if (q == RuneError || is_token_rune(q)) { pos++; continue; }
        }
        if (endsWithToken && pos + phrase.size() < s.size()) {
            int32_t q = sp[pos + phrase.size()];
            if (q >= 0x80) q = decode_rune(sp + pos + phrase.size(), s.size() - pos - phrase.size(), &sz);
            if (q == RuneError || is_token_rune(q)) { pos++; continue; }
        }
        return (int64_t)pos;
    }
}
inline bool match_phrase(sv s, sv phrase) {   // :211-218
    if (phrase.empty()) return s.empty();   // the empty phrase matches only the empty string
    return get_phrase_pos(s, phrase) >= 0;
}
// matchSequence filter_sequence.go:260-269
inline bool match_sequence(sv s, const std::vector<std::string>& phrases) {
    for (auto& ph : phrases) {
        int64_t n = get_phrase_pos(s, ph);
        if (n < 0) return false;
        s.remove_prefix((size_t)n + ph.size());
    }
    return true;
}
// matchExactPrefix filter_exact_prefix.go:275-277
inline bool match_exact_prefix(sv s, sv prefix) { return s.size() >= prefix.size() && s.compare(0, prefix.size(), prefix) == 0; }
// matchLenRange filter_len_range.go:333-336: utf8.RuneCountInString counts every invalid byte as one rune
inline bool match_len_range(sv s, uint64_t minLen, uint64_t maxLen) {
    uint64_t n = 0; const uint8_t* p = (const uint8_t*)s.data(); size_t left = s.size();
    while (left) { int sz; decode_rune(p, left, &sz); p += sz; left -= (size_t)sz; n++; }
    return n >= minLen && n <= maxLen;
}
// matchStringRange filter_string_range.go:226-230: plain byte-wise comparison
inline bool match_string_range(sv s, sv minValue, sv maxValue) { return s.compare(minValue) >= 0 && s.compare(maxValue) < 0; }
// matchPrefix filter_prefix.go:318-352
inline bool match_prefix(sv s, sv prefix) {
    if (prefix.empty()) return !s.empty();
    if (prefix.size() > s.size()) return false;
    const uint8_t* sp = (const uint8_t*)s.data();
    int sz;
    int32_t r = (uint8_t)prefix[0];
    if (r >= 0x80) r = decode_rune((const uint8_t*)prefix.data(), prefix.size(), &sz);
    bool startsWithToken = is_token_rune(r);
    size_t off = 0;
    for (;;) {
        size_t n = s.find(prefix, off);
        if (n == sv::npos) return false;
        off = n;
        if (startsWithToken && off > 0) {
            int32_t q = sp[off - 1];
            if (q >= 0x80) q = decode_last_rune(sp, off, &sz);
            if (q == RuneError || is_token_rune(q)) { off++; continue; }
        }
        return true;
    }
}
// getTokensSkipLast filter_prefix.go:354-363
inline std::vector<std::string> tokens_skip_last(sv s) {
    for (;;) { int sz; int32_t r = decode_last_rune((const uint8_t*)s.data(), s.size(), &sz); if (!is_token_rune(r)) break; s.remove_suffix(sz); }
    return tokenize_string(s);
}

// ---- block search ----------------------------------------------------------------------------------------------------
struct ScanStats {   // algorithmic-bytes accounting, block-granular, following the reference's short-circuit
    uint64_t blocks = 0, rows = 0, bloom_probe_bytes = 0, values_bytes = 0, bitmap_bytes = 0, blocks_values_read = 0;
};

struct BlockSearch {
    const Block* b = nullptr;
    ScanStats* st = nullptr;
    std::map<std::string, BloomFilter> bloomCache;
    struct Vals { DecodedStringsBlock dec; const DecodedStringsBlock* d = nullptr; std::vector<sv> values; };
    std::map<std::string, std::unique_ptr<Vals>> valuesCache;
    std::vector<int64_t> timestampsCache; bool timestampsCached = false;
    const std::vector<int64_t>& timestamps() {   // getTimestamps block_search.go:479-506
        if (!timestampsCached) { timestampsCache = unmarshal_int64_array(b->ts.data, b->ts.mt, b->minTimestamp, b->rows); timestampsCached = true; }
        return timestampsCache;
    }

    void reset(const Block* blk, ScanStats* s) { b = blk; st = s; bloomCache.clear(); valuesCache.clear(); timestampsCached = false; }
    sv const_value(sv name) const {   // getConstColumnValue block_search.go:232-276
        std::string n = canonical(name);
        for (auto& cc : b->consts) if (cc.name == n) return cc.value;
        return sv();
    }
    const Column* column(sv name) const {   // getColumnHeader :278-324
        std::string n = canonical(name);
        for (auto& c : b->columns) if (c.name == n) return &c;
        return nullptr;
    }
    const BloomFilter& bloom(const Column* ch) {   // getBloomFilterForColumn :411-439
        auto it = bloomCache.find(ch->name);
        if (it != bloomCache.end()) return it->second;
        BloomFilter bf;
        if (!bf.unmarshal(ch->bloom)) throw std::runtime_error("cannot unmarshal bloom filter");
        return bloomCache.emplace(ch->name, std::move(bf)).first->second;
    }
    const std::vector<sv>& values(const Column* ch) {   // getValuesForColumn :444-474
        auto it = valuesCache.find(ch->name);
        if (it != valuesCache.end()) return it->second->values;
        auto v = std::make_unique<Vals>();
        if (ch->predecoded) v->d = ch->predecoded.get();
        else { v->dec = decode_values_block_stage(ch->valuesBlock); v->d = &v->dec; }
        v->values = unmarshal_strings(*v->d, b->rows);
        if (st) { st->values_bytes += v->d->lens_items.size() + v->d->data.size(); st->blocks_values_read++; }
        auto& ref = *v;
        valuesCache.emplace(ch->name, std::move(v));
        return ref.values;
    }
    bool bloom_all(const Column* ch, const std::vector<uint64_t>& hashes) {   // matchBloomFilterAllTokens filter_phrase.go:302-308
        if (hashes.empty()) return true;
        if (st) st->bloom_probe_bytes += 8 * hashes.size();
        return bloom(ch).contains_all(hashes);
    }
    template <class F> void visit_values(const Column* ch, Bitmap& bm, F&& f) {   // visitValues :291-300
        if (bm.is_zero()) return;
        const auto& vals = values(ch);
        bm.for_each_set_bit([&](uint64_t idx) { return f(vals[idx]); });
    }
    void match_encoded_dict(const Column* ch, Bitmap& bm, const std::vector<uint8_t>& lut) {   // matchEncodedValuesDict :272-289
        bool any = false; for (uint8_t c : lut) any |= c == 1;
        if (!any) { bm.reset_bits(); return; }
        visit_values(ch, bm, [&](sv v) {
            if (v.size() != 1) throw std::runtime_error("unexpected length for dict value");
            uint8_t idx = (uint8_t)v[0];
            if (idx >= lut.size()) throw std::runtime_error("too big index for dict value");
            return lut[idx] == 1;
        });
    }
};

// value -> string for numeric columns (filter_phrase.go:310-346, filter_prefix.go:365-408)
inline std::string encoded_to_string(uint8_t vt, sv v) {
    std::string s;
    const uint8_t* p = (const uint8_t*)v.data();
    switch (vt) {
    case VT_UINT8: if (v.size() != 1) throw std::runtime_error("bad uint8 len"); marshal_uint64_string(s, p[0]); break;
    case VT_UINT16: if (v.size() != 2) throw std::runtime_error("bad uint16 len"); marshal_uint64_string(s, get_be16(p)); break;
    case VT_UINT32: if (v.size() != 4) throw std::runtime_error("bad uint32 len"); marshal_uint64_string(s, get_be32(p)); break;
    case VT_UINT64: if (v.size() != 8) throw std::runtime_error("bad uint64 len"); marshal_uint64_string(s, get_be64(p)); break;
    case VT_INT64: if (v.size() != 8) throw std::runtime_error("bad int64 len"); marshal_int64_string(s, unzigzag(get_be64(p))); break;
    case VT_FLOAT64: { if (v.size() != 8) throw std::runtime_error("bad float64 len"); uint64_t b = get_be64(p); double f; memcpy(&f, &b, 8); marshal_float64_string(s, f); break; }
    case VT_IPV4: if (v.size() != 4) throw std::runtime_error("bad ipv4 len"); marshal_ipv4_string(s, get_be32(p)); break;
    case VT_ISO8601: if (v.size() != 8) throw std::runtime_error("bad iso8601 len"); marshal_timestamp_iso8601_string(s, (int64_t)get_be64(p)); break;
    default: s = std::string(v);
    }
    return s;
}

// ---- filters ---------------------------------------------------------------------------------------------------------
struct FieldTokens { std::string field; std::vector<std::string> tokens; std::vector<uint64_t> hashes; };

enum FilterKind { F_NOOP, F_PHRASE, F_PREFIX, F_EXACT, F_IN, F_REGEXP, F_AND, F_OR, F_NOT, F_EXACT_PREFIX, F_SEQUENCE, F_LEN_RANGE, F_STRING_RANGE, F_IPV4_RANGE, F_CONTAINS_ALL, F_CONTAINS_ANY,
                  F_ANY_CASE_PHRASE, F_ANY_CASE_PREFIX, F_VALUE_TYPE, F_EQ_FIELD, F_RANGE, F_LE_FIELD, F_TIME, F_DAY_RANGE, F_WEEK_RANGE };

struct Filter {
    FilterKind kind;
    virtual ~Filter() {}
    virtual void apply(BlockSearch& bs, Bitmap& bm) = 0;   // applyToBlockSearch
    // tokens contributed to the AND/OR bloom pre-pass (filter_and.go:131-165); has_tokens=false => "default:" branch
    virtual bool leaf_tokens(std::string*, std::vector<std::string>*) { return false; }
};
using FP = std::shared_ptr<Filter>;

// binary-equality over fixed-width encodings: matchBinaryValue filter_exact.go:356-364
inline void match_binary_value(BlockSearch& bs, const Column* ch, Bitmap& bm, sv bin, const std::vector<uint64_t>& tokens) {
    if (!bs.bloom_all(ch, tokens)) { bm.reset_bits(); return; }
    bs.visit_values(ch, bm, [&](sv v) { return v == bin; });
}

// matchUintNByExactValue / matchInt64 / matchFloat64 / matchIPv4 / matchTimestampISO8601 ByExactValue filter_exact.go:237-354
inline void match_numeric_exact(BlockSearch& bs, const Column* ch, Bitmap& bm, sv value, const std::vector<uint64_t>& tokens) {
    std::string bin;
    switch (ch->valueType) {
    case VT_UINT8: case VT_UINT16: case VT_UINT32: case VT_UINT64: {
        uint64_t n;
        if (!try_parse_uint64(value, &n) || n < ch->minValue || n > ch->maxValue) { bm.reset_bits(); return; }
        if (ch->valueType == VT_UINT8) bin.push_back((char)n); else if (ch->valueType == VT_UINT16) put_be16(bin, (uint16_t)n);
        else if (ch->valueType == VT_UINT32) put_be32(bin, (uint32_t)n); else put_be64(bin, n);
        break;
    }
    case VT_INT64: {
        int64_t n;
        if (!try_parse_int64(value, &n) || n < (int64_t)ch->minValue || n > (int64_t)ch->maxValue) { bm.reset_bits(); return; }
        put_be64(bin, zigzag(n)); break;
    }
    case VT_FLOAT64: {
        double f, mn, mx; memcpy(&mn, &ch->minValue, 8); memcpy(&mx, &ch->maxValue, 8);
        if (!try_parse_float64_exact(value, &f) || f < mn || f > mx) { bm.reset_bits(); return; }
        uint64_t b; memcpy(&b, &f, 8); put_be64(bin, b); break;
    }
    case VT_IPV4: {
        uint32_t n;
        if (!try_parse_ipv4(value, &n) || (uint64_t)n < ch->minValue || (uint64_t)n > ch->maxValue) { bm.reset_bits(); return; }
        put_be32(bin, n); break;
    }
    case VT_ISO8601: {
        int64_t n;
        if (!try_parse_timestamp_iso8601(value, &n) || n < (int64_t)ch->minValue || n > (int64_t)ch->maxValue) { bm.reset_bits(); return; }
        put_be64(bin, (uint64_t)n); break;
    }
    default: throw std::runtime_error("match_numeric_exact: bad type");
    }
    match_binary_value(bs, ch, bm, bin, tokens);
}

struct FilterNoop : Filter { FilterNoop() { kind = F_NOOP; } void apply(BlockSearch&, Bitmap&) override {} };

struct FilterPhrase : Filter {   // filter_phrase.go:25-111
    std::string field, phrase; std::vector<std::string> tokens; std::vector<uint64_t> hashes;
    FilterPhrase(sv f, sv p) : field(f), phrase(p) { kind = F_PHRASE; tokens = tokenize_string(phrase); hashes = tokens_hashes(tokens); }
    bool leaf_tokens(std::string* f, std::vector<std::string>* t) override { *f = field; *t = tokens; return true; }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        sv v = bs.const_value(field);
        if (!v.empty()) { if (!match_phrase(v, phrase)) bm.reset_bits(); return; }
        const Column* ch = bs.column(field);
        if (!ch) { if (!phrase.empty()) bm.reset_bits(); return; }
        switch (ch->valueType) {
        case VT_STRING:
            if (!bs.bloom_all(ch, hashes)) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { return match_phrase(x, phrase); });
            break;
        case VT_DICT: {
            std::vector<uint8_t> lut; for (auto& d : ch->dict) lut.push_back(match_phrase(d, phrase) ? 1 : 0);
            bs.match_encoded_dict(ch, bm, lut); break;
        }
        case VT_UINT8: case VT_UINT16: case VT_UINT32: case VT_UINT64: case VT_INT64:
            match_numeric_exact(bs, ch, bm, phrase, hashes); break;
        case VT_FLOAT64: {   // matchFloat64ByPhrase :159-186
            double f;
            if (!try_parse_float64_exact(phrase, &f) && phrase != "." && phrase != "+" && phrase != "-") { bm.reset_bits(); return; }
            size_t n = phrase.find('.');
            if (n != std::string::npos && n > 0 && n < phrase.size() - 1) { match_numeric_exact(bs, ch, bm, phrase, hashes); return; }
            to_string_match(bs, ch, bm); break;
        }
        case VT_IPV4: { uint32_t ip; if (try_parse_ipv4(phrase, &ip)) { match_numeric_exact(bs, ch, bm, phrase, hashes); return; } to_string_match(bs, ch, bm); break; }
        case VT_ISO8601: { int64_t t; if (try_parse_timestamp_iso8601(phrase, &t)) { match_numeric_exact(bs, ch, bm, phrase, hashes); return; } to_string_match(bs, ch, bm); break; }
        default: throw std::runtime_error("unknown valueType");
        }
    }
    void to_string_match(BlockSearch& bs, const Column* ch, Bitmap& bm) {
        if (!bs.bloom_all(ch, hashes)) { bm.reset_bits(); return; }
        bs.visit_values(ch, bm, [&](sv x) { return match_phrase(encoded_to_string(ch->valueType, x), phrase); });
    }
};

struct FilterPrefix : Filter {   // filter_prefix.go:20-106
    std::string field, prefix; std::vector<std::string> tokens; std::vector<uint64_t> hashes;
    FilterPrefix(sv f, sv p) : field(f), prefix(p) { kind = F_PREFIX; tokens = tokens_skip_last(prefix); hashes = tokens_hashes(tokens); }
    bool leaf_tokens(std::string* f, std::vector<std::string>* t) override { *f = field; *t = tokens; return true; }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        sv v = bs.const_value(field);
        if (!v.empty()) { if (!match_prefix(v, prefix)) bm.reset_bits(); return; }
        const Column* ch = bs.column(field);
        if (!ch) { bm.reset_bits(); return; }
        switch (ch->valueType) {
        case VT_STRING:
            if (!bs.bloom_all(ch, hashes)) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { return match_prefix(x, prefix); });
            break;
        case VT_DICT: {
            std::vector<uint8_t> lut; for (auto& d : ch->dict) lut.push_back(match_prefix(d, prefix) ? 1 : 0);
            bs.match_encoded_dict(ch, bm, lut); break;
        }
        case VT_UINT8: case VT_UINT16: case VT_UINT32: case VT_UINT64: {   // :201-291
            if (prefix.empty()) return;
            uint64_t n;
            if (!try_parse_uint64(prefix, &n) || n > ch->maxValue) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { return match_prefix(encoded_to_string(ch->valueType, x), prefix); });
            break;
        }
        case VT_INT64: {   // :293-316
            if (prefix.empty()) return;
            if (prefix != "-") {
                int64_t n;
                if (!try_parse_int64(prefix, &n) || n < (int64_t)ch->minValue || n > (int64_t)ch->maxValue) { bm.reset_bits(); return; }
            }
            bs.visit_values(ch, bm, [&](sv x) { return match_prefix(encoded_to_string(ch->valueType, x), prefix); });
            break;
        }
        case VT_FLOAT64: {   // :150-176
            if (prefix.empty()) return;
            double f;
            if (!try_parse_float64_exact(prefix, &f) && prefix != "." && prefix != "+" && prefix != "-" && prefix[0] != 'e' && prefix[0] != 'E') { bm.reset_bits(); return; }
            if (!bs.bloom_all(ch, hashes)) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { return match_prefix(encoded_to_string(ch->valueType, x), prefix); });
            break;
        }
        case VT_IPV4: case VT_ISO8601: {   // :108-148
            if (prefix.empty()) return;
            if (!bs.bloom_all(ch, hashes)) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { return match_prefix(encoded_to_string(ch->valueType, x), prefix); });
            break;
        }
        default: throw std::runtime_error("unknown valueType");
        }
    }
};

// Filters of SURVEY §8(f) rank 3 (oracle side first; the product follows in round 2).  Same skeleton as above: const column, missing
// column, then the per-valueType paths of the reference with their header-level early outs.
inline void dict_lut(BlockSearch& bs, const Column* ch, Bitmap& bm, const std::function<bool(sv)>& f) {
    std::vector<uint8_t> lut; for (auto& d : ch->dict) lut.push_back(f(d) ? 1 : 0);
    bs.match_encoded_dict(ch, bm, lut);
}

struct FilterExactPrefix : Filter {   // filter_exact_prefix.go:13-277
    std::string field, prefix; std::vector<std::string> tokens; std::vector<uint64_t> hashes;
    FilterExactPrefix(sv f, sv p) : field(f), prefix(p) { kind = F_EXACT_PREFIX; tokens = tokens_skip_last(prefix); hashes = tokens_hashes(tokens); }
    bool leaf_tokens(std::string* f, std::vector<std::string>* t) override { *f = field; *t = tokens; return true; }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        sv v = bs.const_value(field);
        if (!v.empty()) { if (!match_exact_prefix(v, prefix)) bm.reset_bits(); return; }
        const Column* ch = bs.column(field);
        if (!ch) { if (!match_exact_prefix("", prefix)) bm.reset_bits(); return; }
        auto to_string_visit = [&] { bs.visit_values(ch, bm, [&](sv x) { return match_exact_prefix(encoded_to_string(ch->valueType, x), prefix); }); };
        switch (ch->valueType) {
        case VT_STRING:
            if (!bs.bloom_all(ch, hashes)) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { return match_exact_prefix(x, prefix); });
            break;
        case VT_DICT: dict_lut(bs, ch, bm, [&](sv d) { return match_exact_prefix(d, prefix); }); break;
        case VT_UINT8: case VT_UINT16: case VT_UINT32: case VT_UINT64: {   // matchMinMaxExactPrefix :258-273
            if (prefix.empty()) return;
            if (!hashes.empty()) { bm.reset_bits(); return; }
            uint64_t n;
            if (!try_parse_uint64(prefix, &n) || n > ch->maxValue) { bm.reset_bits(); return; }
            to_string_visit(); break;
        }
        case VT_INT64: {   // :235-256
            if (prefix.empty()) return;
            if (!hashes.empty()) { bm.reset_bits(); return; }
            if (prefix != "-") {
                int64_t n;
                if (!try_parse_int64(prefix, &n) || n > (int64_t)ch->maxValue || n < (int64_t)ch->minValue) { bm.reset_bits(); return; }
            }
            to_string_visit(); break;
        }
        case VT_FLOAT64:   // :141-157
            if (prefix.empty()) return;
            if (hashes.size() > 2 * 6 || !bs.bloom_all(ch, hashes)) { bm.reset_bits(); return; }
            to_string_visit(); break;
        case VT_IPV4:      // :123-139
            if (prefix.empty()) return;
            if (prefix < "0" || prefix > "9" || hashes.size() > 3 * 6 || !bs.bloom_all(ch, hashes)) { bm.reset_bits(); return; }
            to_string_visit(); break;
        case VT_ISO8601:   // :105-121
            if (prefix.empty()) return;
            if (prefix < "0" || prefix > "9" || !bs.bloom_all(ch, hashes)) { bm.reset_bits(); return; }
            to_string_visit(); break;
        default: throw std::runtime_error("unknown valueType");
        }
    }
};

struct FilterSequence : Filter {   // filter_sequence.go:12-269
    std::string field; std::vector<std::string> phrases; std::vector<std::string> tokens; std::vector<uint64_t> hashes;
    FilterSequence(sv f, const std::vector<std::string>& ph) : field(f) {
        kind = F_SEQUENCE;
        for (auto& p : ph) if (!p.empty()) phrases.push_back(p);   // getNonEmptyPhrases :58-67
        std::vector<sv> views(phrases.begin(), phrases.end());
        tokens = tokenize_strings(views); hashes = tokens_hashes(tokens);
    }
    bool leaf_tokens(std::string* f, std::vector<std::string>* t) override { *f = field; *t = tokens; return true; }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        if (phrases.empty()) return;
        sv v = bs.const_value(field);
        if (!v.empty()) { if (!match_sequence(v, phrases)) bm.reset_bits(); return; }
        const Column* ch = bs.column(field);
        if (!ch) { if (!match_sequence("", phrases)) bm.reset_bits(); return; }
        auto to_string_visit = [&] {
            if (!bs.bloom_all(ch, hashes)) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { return match_sequence(encoded_to_string(ch->valueType, x), phrases); });
        };
        switch (ch->valueType) {
        case VT_STRING:
            if (!bs.bloom_all(ch, hashes)) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { return match_sequence(x, phrases); });
            break;
        case VT_DICT: dict_lut(bs, ch, bm, [&](sv d) { return match_sequence(d, phrases); }); break;
        case VT_UINT8: case VT_UINT16: case VT_UINT32: case VT_UINT64: case VT_INT64:   // :213-258: one phrase => exact value
            if (phrases.size() > 1) { bm.reset_bits(); return; }
            match_numeric_exact(bs, ch, bm, phrases[0], hashes); break;
        case VT_FLOAT64: to_string_visit(); break;
        case VT_IPV4: case VT_ISO8601:   // :139-175: one phrase => the phrase filter's path for the type
            if (phrases.size() == 1) { single_phrase(bs, ch, bm); return; }
            to_string_visit(); break;
        default: throw std::runtime_error("unknown valueType");
        }
    }
    void single_phrase(BlockSearch& bs, const Column* ch, Bitmap& bm) {   // matchIPv4ByPhrase / matchTimestampISO8601ByPhrase with the sequence's tokens
        const std::string& phrase = phrases[0];
        bool exact;
        if (ch->valueType == VT_IPV4) { uint32_t ip; exact = try_parse_ipv4(phrase, &ip); } else { int64_t t; exact = try_parse_timestamp_iso8601(phrase, &t); }
        if (exact) { match_numeric_exact(bs, ch, bm, phrase, hashes); return; }
        if (!bs.bloom_all(ch, hashes)) { bm.reset_bits(); return; }
        bs.visit_values(ch, bm, [&](sv x) { return match_phrase(encoded_to_string(ch->valueType, x), phrase); });
    }
};

struct FilterLenRange : Filter {   // filter_len_range.go:14-348
    std::string field; uint64_t minLen, maxLen;
    FilterLenRange(sv f, uint64_t mn, uint64_t mx) : field(f), minLen(mn), maxLen(mx) { kind = F_LEN_RANGE; }
    static size_t dec_len_u(uint64_t v) { std::string s; marshal_uint64_string(s, v); return s.size(); }
    static size_t dec_len_i(int64_t v) { std::string s; marshal_int64_string(s, v); return s.size(); }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        if (minLen > maxLen) { bm.reset_bits(); return; }
        sv v = bs.const_value(field);
        if (!v.empty()) { if (!match_len_range(v, minLen, maxLen)) bm.reset_bits(); return; }
        const Column* ch = bs.column(field);
        if (!ch) { if (!match_len_range("", minLen, maxLen)) bm.reset_bits(); return; }
        auto to_string_visit = [&] { bs.visit_values(ch, bm, [&](sv x) { return match_len_range(encoded_to_string(ch->valueType, x), minLen, maxLen); }); };
        auto uint_path = [&](uint64_t max_digits) {   // matchUintNByLenRange + matchMinMaxValueLen :338-348
            if (minLen > max_digits || maxLen == 0) { bm.reset_bits(); return; }
            if (maxLen < dec_len_u(ch->minValue) || minLen > dec_len_u(ch->maxValue)) { bm.reset_bits(); return; }
            to_string_visit();
        };
        switch (ch->valueType) {
        case VT_STRING: bs.visit_values(ch, bm, [&](sv x) { return match_len_range(x, minLen, maxLen); }); break;
        case VT_DICT: dict_lut(bs, ch, bm, [&](sv d) { return match_len_range(d, minLen, maxLen); }); break;
        case VT_UINT8: uint_path(3); break;
        case VT_UINT16: uint_path(5); break;
        case VT_UINT32: uint_path(10); break;
        case VT_UINT64: uint_path(20); break;
        case VT_INT64: {   // :305-331
            if (minLen > 21 || maxLen == 0) { bm.reset_bits(); return; }
            size_t mx = std::max(dec_len_i((int64_t)ch->minValue), dec_len_i((int64_t)ch->maxValue));
            if ((uint64_t)mx < minLen) { bm.reset_bits(); return; }
            to_string_visit(); break;
        }
        case VT_FLOAT64: if (minLen > 24 || maxLen == 0) { bm.reset_bits(); return; } to_string_visit(); break;
        case VT_IPV4: if (minLen > 15 || maxLen < 7) { bm.reset_bits(); return; } to_string_visit(); break;
        case VT_ISO8601: if (minLen > 24 || maxLen < 24) bm.reset_bits(); break;   // every value is len("2006-01-02T15:04:05.000Z") long; no values are read
        default: throw std::runtime_error("unknown valueType");
        }
    }
};

struct FilterStringRange : Filter {   // filter_string_range.go:12-230
    std::string field, minValue, maxValue;
    FilterStringRange(sv f, sv mn, sv mx) : field(f), minValue(mn), maxValue(mx) { kind = F_STRING_RANGE; }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        if (minValue > maxValue) { bm.reset_bits(); return; }
        sv v = bs.const_value(field);
        if (!v.empty()) { if (!match_string_range(v, minValue, maxValue)) bm.reset_bits(); return; }
        const Column* ch = bs.column(field);
        if (!ch) { if (!match_string_range("", minValue, maxValue)) bm.reset_bits(); return; }
        auto to_string_visit = [&] { bs.visit_values(ch, bm, [&](sv x) { return match_string_range(encoded_to_string(ch->valueType, x), minValue, maxValue); }); };
        switch (ch->valueType) {
        case VT_STRING: bs.visit_values(ch, bm, [&](sv x) { return match_string_range(x, minValue, maxValue); }); break;
        case VT_DICT: dict_lut(bs, ch, bm, [&](sv d) { return match_string_range(d, minValue, maxValue); }); break;
        case VT_UINT8: case VT_UINT16: case VT_UINT32: case VT_UINT64: case VT_IPV4: case VT_ISO8601:
            if (minValue > "9" || maxValue < "0") { bm.reset_bits(); return; }
            to_string_visit(); break;
        case VT_INT64:   // :213-224
            if ((minValue != "-" && minValue > "9") || (maxValue != "-" && maxValue < "0")) { bm.reset_bits(); return; }
            to_string_visit(); break;
        case VT_FLOAT64:
            if (minValue > "9" || maxValue < "+") { bm.reset_bits(); return; }
            to_string_visit(); break;
        default: throw std::runtime_error("unknown valueType");
        }
    }
};

struct FilterIPv4Range : Filter {   // filter_ipv4_range.go:12-191
    std::string field; uint32_t minValue, maxValue;
    FilterIPv4Range(sv f, uint32_t mn, uint32_t mx) : field(f), minValue(mn), maxValue(mx) { kind = F_IPV4_RANGE; }
    bool match_str(sv s) const { uint32_t n; return try_parse_ipv4(s, &n) && n >= minValue && n <= maxValue; }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        if (minValue > maxValue) { bm.reset_bits(); return; }
        sv v = bs.const_value(field);
        if (!v.empty()) { if (!match_str(v)) bm.reset_bits(); return; }
        const Column* ch = bs.column(field);
        if (!ch) { bm.reset_bits(); return; }
        switch (ch->valueType) {
        case VT_STRING: bs.visit_values(ch, bm, [&](sv x) { return match_str(x); }); break;
        case VT_DICT: dict_lut(bs, ch, bm, [&](sv d) { return match_str(d); }); break;
        case VT_IPV4:   // matchIPv4ByRange :176-191
            if (ch->minValue > (uint64_t)maxValue || ch->maxValue < (uint64_t)minValue) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { if (x.size() != 4) throw std::runtime_error("unexpected length for binary representation of IPv4"); uint32_t n = get_be32((const uint8_t*)x.data()); return n >= minValue && n <= maxValue; });
            break;
        case VT_UINT8: case VT_UINT16: case VT_UINT32: case VT_UINT64: case VT_INT64: case VT_FLOAT64: case VT_ISO8601: bm.reset_bits(); break;
        default: throw std::runtime_error("unknown valueType");
        }
    }
};

struct FilterExact : Filter {   // filter_exact.go:17-235
    std::string field, value; std::vector<std::string> tokens; std::vector<uint64_t> hashes;
    FilterExact(sv f, sv v) : field(f), value(v) { kind = F_EXACT; tokens = tokenize_string(value); hashes = tokens_hashes(tokens); }
    bool leaf_tokens(std::string* f, std::vector<std::string>* t) override { *f = field; *t = tokens; return true; }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        sv v = bs.const_value(field);
        if (!v.empty()) { if (value != v) bm.reset_bits(); return; }
        const Column* ch = bs.column(field);
        if (!ch) { if (!value.empty()) bm.reset_bits(); return; }
        switch (ch->valueType) {
        case VT_STRING:
            if (!bs.bloom_all(ch, hashes)) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { return x == value; });
            break;
        case VT_DICT: {
            std::vector<uint8_t> lut; for (auto& d : ch->dict) lut.push_back(d == value ? 1 : 0);
            bs.match_encoded_dict(ch, bm, lut); break;
        }
        default: match_numeric_exact(bs, ch, bm, value, hashes);
        }
    }
};

struct FilterIn : Filter {   // filter_in.go:14-234 + in_values.go
    std::string field; std::vector<std::string> values;
    std::unordered_set<std::string> strset;
    std::vector<uint64_t> commonHashes; std::vector<std::vector<uint64_t>> tokenSetsHashes;
    FilterIn(sv f, const std::vector<std::string>& vals) : field(f), values(vals) {
        kind = F_IN;
        for (auto& v : values) strset.insert(v);
        // getCommonTokensAndTokenSets in_values.go:317-346
        std::vector<std::vector<std::string>> sets;
        for (auto& v : values) sets.push_back(tokenize_string(v));
        std::vector<std::string> common;
        if (!sets.empty()) {
            common = sets[0];
            for (size_t i = 1; i < sets.size(); i++) {
                if (common.empty()) break;
                std::vector<std::string> d;
                for (auto& t : common) if (std::find(sets[i].begin(), sets[i].end(), t) != sets[i].end()) d.push_back(t);
                common = d;
            }
        }
        if (!common.empty()) for (auto& s : sets) {
            std::vector<std::string> d;
            for (auto& t : s) if (std::find(common.begin(), common.end(), t) == common.end()) d.push_back(t);
            s = d;
        }
        commonHashes = tokens_hashes(common);
        for (auto& s : sets) tokenSetsHashes.push_back(tokens_hashes(s));
    }
    // typed sets in_values.go:141-315
    std::unordered_set<std::string> bin_values(uint8_t vt) const {
        std::unordered_set<std::string> m;
        for (auto& v : values) {
            std::string b;
            switch (vt) {
            case VT_UINT8: { uint64_t n; if (!try_parse_uint64(v, &n) || n >= (1ULL << 8)) continue; b.push_back((char)n); break; }
            case VT_UINT16: { uint64_t n; if (!try_parse_uint64(v, &n) || n >= (1ULL << 16)) continue; put_be16(b, (uint16_t)n); break; }
            case VT_UINT32: { uint64_t n; if (!try_parse_uint64(v, &n) || n >= (1ULL << 32)) continue; put_be32(b, (uint32_t)n); break; }
            case VT_UINT64: { uint64_t n; if (!try_parse_uint64(v, &n)) continue; put_be64(b, n); break; }
            case VT_INT64: { int64_t n; if (!try_parse_int64(v, &n)) continue; put_be64(b, zigzag(n)); break; }
            case VT_FLOAT64: { double f; if (!try_parse_float64_exact(v, &f)) continue; uint64_t u; memcpy(&u, &f, 8); put_be64(b, u); break; }
            case VT_IPV4: { uint32_t n; if (!try_parse_ipv4(v, &n)) continue; put_be32(b, n); break; }
            case VT_ISO8601: { int64_t n; if (!try_parse_timestamp_iso8601(v, &n)) continue; put_be64(b, (uint64_t)n); break; }
            default: b = v;
            }
            m.insert(b);
        }
        return m;
    }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        if (values.empty()) { bm.reset_bits(); return; }
        sv v = bs.const_value(field);
        if (!v.empty()) { if (!strset.count(std::string(v))) bm.reset_bits(); return; }
        const Column* ch = bs.column(field);
        if (!ch) { if (!strset.count("")) bm.reset_bits(); return; }
        if (ch->valueType == VT_DICT) {
            std::vector<uint8_t> lut; for (auto& d : ch->dict) lut.push_back(strset.count(d) ? 1 : 0);
            bs.match_encoded_dict(ch, bm, lut); return;
        }
        std::unordered_set<std::string> bin = ch->valueType == VT_STRING ? strset : bin_values(ch->valueType);
        // matchAnyValue :187-200
        if (bin.empty()) { bm.reset_bits(); return; }
        // matchBloomFilterAnyTokenSet :202-218
        if (!bs.bloom_all(ch, commonHashes)) { bm.reset_bits(); return; }
        if (!(tokenSetsHashes.size() > 1000 || tokenSetsHashes.size() > 10 * bs.b->rows)) {
            bool any = false;
            const BloomFilter& bf = bs.bloom(ch);
            for (auto& t : tokenSetsHashes) { if (bs.st) bs.st->bloom_probe_bytes += 8 * t.size(); if (bf.contains_all(t)) { any = true; break; } }
            if (!any) { bm.reset_bits(); return; }
        }
        bs.visit_values(ch, bm, [&](sv x) { return bin.count(std::string(x)) > 0; });
    }
};

// contains_all() / contains_any(): the values are phrases (filter_contains_all.go, filter_contains_any.go, in_values.go)
inline bool match_all_phrases(sv v, const std::vector<std::string>& phrases) {   // filter_contains_all.go:310-321
    for (auto& p : phrases) { if (p.empty()) continue; if (!match_phrase(v, p)) return false; }
    return true;
}
inline bool match_any_phrase(sv v, const std::vector<std::string>& phrases) {    // filter_contains_any.go:293-300
    for (auto& p : phrases) if (match_phrase(v, p)) return true;
    return false;
}

struct FilterContainsAll : Filter {   // filter_contains_all.go:12-321
    FilterIn iv;   // reuses the inValues restatement (string set, typed sets)
    std::vector<uint64_t> hashesAll;   // inValues.getTokensHashesAll in_values.go:94-102
    FilterContainsAll(sv f, const std::vector<std::string>& vals) : iv(f, vals) {
        kind = F_CONTAINS_ALL;
        std::vector<sv> views(iv.values.begin(), iv.values.end());
        for (uint64_t h : tokenize_hashes(views)) append_hashes_hashes(hashesAll, h);
    }
    size_t non_empty_values_len() const { return iv.strset.size() - iv.strset.count(""); }   // :80-88
    void apply(BlockSearch& bs, Bitmap& bm) override {
        const auto& phrases = iv.values;
        if (phrases.empty() || (phrases.size() == 1 && phrases[0].empty())) return;
        sv v = bs.const_value(iv.field);
        if (!v.empty()) { if (!match_all_phrases(v, phrases)) bm.reset_bits(); return; }
        const Column* ch = bs.column(iv.field);
        if (!ch) { if (!match_all_phrases("", phrases)) bm.reset_bits(); return; }
        switch (ch->valueType) {
        case VT_STRING:
            if (!bs.bloom_all(ch, hashesAll)) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { return match_all_phrases(x, phrases); });
            break;
        case VT_DICT: dict_lut(bs, ch, bm, [&](sv d) { return match_all_phrases(d, phrases); }); break;
        case VT_UINT8: case VT_UINT16: case VT_UINT32: case VT_UINT64: {   // matchAllValues :168-189
            size_t n = non_empty_values_len();
            if (n == 0) return;
            auto bin = iv.bin_values(ch->valueType);
            if (n != 1 || n != bin.size()) { bm.reset_bits(); return; }
            if (!bs.bloom_all(ch, hashesAll)) { bm.reset_bits(); return; }
            const std::string& want = *bin.begin();
            bs.visit_values(ch, bm, [&](sv x) { return x == want; });
            break;
        }
        case VT_INT64: case VT_FLOAT64: case VT_IPV4: case VT_ISO8601:
            if (!bs.bloom_all(ch, hashesAll)) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { return match_all_phrases(encoded_to_string(ch->valueType, x), phrases); });
            break;
        default: throw std::runtime_error("unknown valueType");
        }
    }
};

struct FilterContainsAny : Filter {   // filter_contains_any.go:12-300
    FilterIn iv;   // common tokens + per-value token sets: getTokensHashesAny == the in() filter's
    FilterContainsAny(sv f, const std::vector<std::string>& vals) : iv(f, vals) { kind = F_CONTAINS_ANY; }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        const auto& phrases = iv.values;
        if (phrases.empty()) { bm.reset_bits(); return; }
        if (iv.strset.count("")) return;   // hasEmptyValue: the empty phrase matches everything here
        sv v = bs.const_value(iv.field);
        if (!v.empty()) { if (!match_any_phrase(v, phrases)) bm.reset_bits(); return; }
        const Column* ch = bs.column(iv.field);
        if (!ch) { if (!match_any_phrase("", phrases)) bm.reset_bits(); return; }
        // matchValuesAnyPhrase :170-189: only the phrases whose own tokens pass the bloom filter are tried on the rows
        auto any_phrase = [&](bool to_string) {
            if (!bs.bloom_all(ch, iv.commonHashes)) { bm.reset_bits(); return; }
            const BloomFilter& bf = bs.bloom(ch);
            std::vector<std::string> alive;
            for (size_t i = 0; i < phrases.size(); i++) { if (bs.st) bs.st->bloom_probe_bytes += 8 * iv.tokenSetsHashes[i].size(); if (bf.contains_all(iv.tokenSetsHashes[i])) alive.push_back(phrases[i]); }
            if (alive.empty()) { bm.reset_bits(); return; }
            if (to_string) bs.visit_values(ch, bm, [&](sv x) { return match_any_phrase(encoded_to_string(ch->valueType, x), alive); });
            else bs.visit_values(ch, bm, [&](sv x) { return match_any_phrase(x, alive); });
        };
        switch (ch->valueType) {
        case VT_STRING: any_phrase(false); break;
        case VT_DICT: dict_lut(bs, ch, bm, [&](sv d) { return match_any_phrase(d, phrases); }); break;
        case VT_UINT8: case VT_UINT16: case VT_UINT32: case VT_UINT64: {   // matchAnyValue filter_in.go:187-218
            auto bin = iv.bin_values(ch->valueType);
            if (bin.empty()) { bm.reset_bits(); return; }
            if (!bs.bloom_all(ch, iv.commonHashes)) { bm.reset_bits(); return; }
            if (!(iv.tokenSetsHashes.size() > 1000 || iv.tokenSetsHashes.size() > 10 * bs.b->rows)) {
                bool any = false;
                const BloomFilter& bf = bs.bloom(ch);
                for (auto& t : iv.tokenSetsHashes) { if (bs.st) bs.st->bloom_probe_bytes += 8 * t.size(); if (bf.contains_all(t)) { any = true; break; } }
                if (!any) { bm.reset_bits(); return; }
            }
            bs.visit_values(ch, bm, [&](sv x) { return bin.count(std::string(x)) > 0; });
            break;
        }
        case VT_INT64: case VT_FLOAT64: case VT_IPV4: case VT_ISO8601: any_phrase(true); break;
        default: throw std::runtime_error("unknown valueType");
        }
    }
};

// i(phrase) / i(prefix*): case-insensitive forms.  The needle is lower-cased once; a value is lower-cased only when it is not plain
// lower-case ASCII already.  The byte-length check happens BEFORE the value is lower-cased (lower-casing can change the byte length).
inline bool match_any_case_phrase(sv s, sv phraseLower) {   // filter_any_case_phrase.go:161-178
    if (phraseLower.empty()) return s.empty();
    if (phraseLower.size() > s.size()) return false;
    if (is_ascii_lowercase(s)) return match_phrase(s, phraseLower);
    return match_phrase(strings_to_lower(s), phraseLower);
}
inline bool match_any_case_prefix(sv s, sv prefixLower) {   // filter_any_case_prefix.go:150-167
    if (prefixLower.empty()) return !s.empty();
    if (prefixLower.size() > s.size()) return false;
    if (is_ascii_lowercase(s)) return match_prefix(s, prefixLower);
    return match_prefix(strings_to_lower(s), prefixLower);
}
inline std::vector<uint64_t> upper_tokens_hashes(const std::vector<std::string>& tokens) {
    std::vector<std::string> up; for (auto& t : tokens) up.push_back(strings_to_upper(t));
    return tokens_hashes(up);
}

struct FilterAnyCasePhrase : Filter {   // filter_any_case_phrase.go:14-190
    std::string field, phrase;
    FilterPhrase lower, upper;   // the numeric paths are the phrase filter's, run with the lower- (iso8601: upper-) cased phrase and THIS filter's tokens
    FilterAnyCasePhrase(sv f, sv p) : field(f), phrase(p), lower(f, strings_to_lower(p)), upper(f, strings_to_upper(p)) {
        kind = F_ANY_CASE_PHRASE;
        std::vector<std::string> tokens = tokenize_string(phrase);   // initTokens :44-53: tokens of the phrase as written
        lower.hashes = tokens_hashes(tokens); upper.hashes = upper_tokens_hashes(tokens);
    }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        const std::string& pl = lower.phrase;
        sv v = bs.const_value(field);
        if (!v.empty()) { if (!match_any_case_phrase(v, pl)) bm.reset_bits(); return; }
        const Column* ch = bs.column(field);
        if (!ch) { if (!pl.empty()) bm.reset_bits(); return; }
        switch (ch->valueType) {
        case VT_STRING: bs.visit_values(ch, bm, [&](sv x) { return match_any_case_phrase(x, pl); }); break;   // no bloom probe: tokens are case sensitive
        case VT_DICT: dict_lut(bs, ch, bm, [&](sv d) { return match_any_case_phrase(d, pl); }); break;
        case VT_ISO8601: upper.apply(bs, bm); break;
        default: lower.apply(bs, bm);
        }
    }
};

struct FilterAnyCasePrefix : Filter {   // filter_any_case_prefix.go:14-182
    std::string field, prefix;
    FilterPrefix lower, upper;
    FilterAnyCasePrefix(sv f, sv p) : field(f), prefix(p), lower(f, strings_to_lower(p)), upper(f, strings_to_upper(p)) {
        kind = F_ANY_CASE_PREFIX;
        std::vector<std::string> tokens = tokens_skip_last(prefix);   // initTokens :56-65
        lower.hashes = tokens_hashes(tokens); upper.hashes = upper_tokens_hashes(tokens);
    }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        const std::string& pl = lower.prefix;
        sv v = bs.const_value(field);
        if (!v.empty()) { if (!match_any_case_prefix(v, pl)) bm.reset_bits(); return; }
        const Column* ch = bs.column(field);
        if (!ch) { bm.reset_bits(); return; }
        switch (ch->valueType) {
        case VT_STRING: bs.visit_values(ch, bm, [&](sv x) { return match_any_case_prefix(x, pl); }); break;
        case VT_DICT: dict_lut(bs, ch, bm, [&](sv d) { return match_any_case_prefix(d, pl); }); break;
        case VT_ISO8601: upper.apply(bs, bm); break;
        default: lower.apply(bs, bm);
        }
    }
};

struct FilterValueType : Filter {   // filter_value_type.go:12-67; names: valueType.String() values_encoder.go
    std::string field, typ;
    FilterValueType(sv f, sv t) : field(f), typ(t) { kind = F_VALUE_TYPE; }
    static const char* name_of(uint8_t vt) {
        switch (vt) {
        case VT_STRING: return "string"; case VT_DICT: return "dict"; case VT_UINT8: return "uint8"; case VT_UINT16: return "uint16"; case VT_UINT32: return "uint32";
        case VT_UINT64: return "uint64"; case VT_INT64: return "int64"; case VT_FLOAT64: return "float64"; case VT_IPV4: return "ipv4"; case VT_ISO8601: return "iso8601";
        }
        return "unknown";
    }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        sv v = bs.const_value(field);
        if (!v.empty()) { if (typ != "const") bm.reset_bits(); return; }
        const Column* ch = bs.column(field);
        if (!ch) { bm.reset_bits(); return; }
        if (typ != name_of(ch->valueType)) bm.reset_bits();
    }
};

// the string value of `field` in row idx, as blockResult.getValues would yield it: const value, "" for a missing field, dict entry, or
// the text form of a typed value
inline std::string row_string(BlockSearch& bs, sv field, uint64_t idx) {
    sv v = bs.const_value(field);
    if (!v.empty()) return std::string(v);
    const Column* ch = bs.column(field);
    if (!ch) return std::string();
    sv x = bs.values(ch)[idx];
    if (ch->valueType == VT_DICT) { if (x.size() != 1 || (uint8_t)x[0] >= ch->dict.size()) throw std::runtime_error("bad dict value"); return ch->dict[(uint8_t)x[0]]; }
    return encoded_to_string(ch->valueType, x);
}

struct FilterEqField : Filter {   // filter_eq_field.go:14-238  (field:eq_field(other))
    std::string field, other;
    FilterEqField(sv f, sv o) : field(canonical(f)), other(canonical(o)) { kind = F_EQ_FIELD; }
    void by_strings(BlockSearch& bs, Bitmap& bm) { bm.for_each_set_bit([&](uint64_t idx) { return row_string(bs, field, idx) == row_string(bs, other, idx); }); }   // applyFilterString
    void apply(BlockSearch& bs, Bitmap& bm) override {
        if (field == other) return;
        sv v = bs.const_value(field), vo = bs.const_value(other);
        if (!v.empty() || !vo.empty()) {
            if (!v.empty() && !vo.empty()) { if (v != vo) bm.reset_bits(); return; }
            by_strings(bs, bm); return;
        }
        const Column* ch = bs.column(field); const Column* co = bs.column(other);
        if (!ch || !co) { if (!ch && !co) return; by_strings(bs, bm); return; }
        if (ch->valueType != co->valueType || ch->valueType == VT_STRING) { by_strings(bs, bm); return; }
        if (bm.is_zero()) return;
        const auto& a = bs.values(ch); const auto& b = bs.values(co);
        if (ch->valueType == VT_DICT) bm.for_each_set_bit([&](uint64_t idx) { return ch->dict.at((uint8_t)a[idx][0]) == co->dict.at((uint8_t)b[idx][0]); });   // applyFilterDict
        else bm.for_each_set_bit([&](uint64_t idx) { return a[idx] == b[idx]; });   // applyFilterBinValue: same type, same binary form
    }
};

struct FilterRange : Filter {   // filter_range.go:14-420  (f:range[a, b], f:>a, f:<=b ...)
    std::string field; double minValue, maxValue;
    FilterRange(sv f, double mn, double mx) : field(f), minValue(mn), maxValue(mx) { kind = F_RANGE; }
    bool match_str(sv s) const { double f = parse_math_number(s); return f >= minValue && f <= maxValue; }   // matchRange :352-355
    static uint64_t u64_clamp(double f) { return f < 0 ? 0 : f > 18446744073709551615.0 ? UINT64_MAX : (uint64_t)f; }   // toUint64Clamp :362-370
    static int64_t i64_clamp(double f) { return f < -9223372036854775808.0 ? INT64_MIN : f > 9223372036854775807.0 ? INT64_MAX : (int64_t)f; }   // toInt64Clamp :377-385
    void apply(BlockSearch& bs, Bitmap& bm) override {
        if (minValue > maxValue) { bm.reset_bits(); return; }
        sv v = bs.const_value(field);
        if (!v.empty()) { if (!match_str(v)) bm.reset_bits(); return; }
        const Column* ch = bs.column(field);
        if (!ch) { bm.reset_bits(); return; }
        const double mnc = std::ceil(minValue), mxf = std::floor(maxValue);
        switch (ch->valueType) {
        case VT_STRING: bs.visit_values(ch, bm, [&](sv x) { return match_str(x); }); break;
        case VT_DICT: dict_lut(bs, ch, bm, [&](sv d) { return match_str(d); }); break;
        case VT_UINT8: case VT_UINT16: case VT_UINT32: case VT_UINT64: {   // matchUintNByRange :246-313
            uint64_t lo = u64_clamp(mnc), hi = u64_clamp(mxf);
            if (maxValue < 0 || lo > ch->maxValue || hi < ch->minValue) { bm.reset_bits(); return; }
            size_t w = ch->valueType == VT_UINT8 ? 1 : ch->valueType == VT_UINT16 ? 2 : ch->valueType == VT_UINT32 ? 4 : 8;
            bs.visit_values(ch, bm, [&](sv x) {
                if (x.size() != w) throw std::runtime_error("unexpected length for binary representation of a uint");
                const uint8_t* p = (const uint8_t*)x.data();
                uint64_t n = w == 1 ? p[0] : w == 2 ? get_be16(p) : w == 4 ? get_be32(p) : get_be64(p);
                return n >= lo && n <= hi;
            });
            break;
        }
        case VT_INT64: {   // :314-330
            int64_t lo = i64_clamp(mnc), hi = i64_clamp(mxf);
            if (lo > (int64_t)ch->maxValue || hi < (int64_t)ch->minValue) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { int64_t n = unzigzag(get_be64((const uint8_t*)x.data())); return n >= lo && n <= hi; });
            break;
        }
        case VT_FLOAT64: {   // :216-228
            double cmn, cmx; memcpy(&cmn, &ch->minValue, 8); memcpy(&cmx, &ch->maxValue, 8);
            if (minValue > cmx || maxValue < cmn) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { uint64_t b = get_be64((const uint8_t*)x.data()); double f; memcpy(&f, &b, 8); return f >= minValue && f <= maxValue; });
            break;
        }
        case VT_IPV4: {   // toUint32Range + matchIPv4ByRange
            auto clamp32 = [](double f) -> uint32_t { return f < 0 ? 0u : f > 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)f; };
            uint32_t lo = clamp32(mnc), hi = clamp32(mxf);
            if (ch->minValue > (uint64_t)hi || ch->maxValue < (uint64_t)lo) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { uint32_t n = get_be32((const uint8_t*)x.data()); return n >= lo && n <= hi; });
            break;
        }
        case VT_ISO8601: {   // :331-347
            int64_t lo = i64_clamp(mnc), hi = i64_clamp(mxf);
            if (maxValue < 0 || lo > (int64_t)ch->maxValue || hi < (int64_t)ch->minValue) { bm.reset_bits(); return; }
            bs.visit_values(ch, bm, [&](sv x) { int64_t n = (int64_t)get_be64((const uint8_t*)x.data()); return n >= lo && n <= hi; });
            break;
        }
        default: throw std::runtime_error("unknown valueType");
        }
    }
};

// le_field() / lt_field(): filter_le_field.go:14-313
inline bool le_values_string(sv a, sv b, bool excl) {   // :283-297
    double fa = parse_math_number(a);
    if (!std::isnan(fa)) { double fb = parse_math_number(b); if (!std::isnan(fb)) return excl ? fa < fb : fa <= fb; }
    return excl ? a < b : a <= b;
}
struct FilterLeField : Filter {
    std::string field, other; bool excl;
    FilterLeField(sv f, sv o, bool e) : field(canonical(f)), other(canonical(o)), excl(e) { kind = F_LE_FIELD; }
    void by_strings(BlockSearch& bs, Bitmap& bm) { bm.for_each_set_bit([&](uint64_t idx) { return le_values_string(row_string(bs, field, idx), row_string(bs, other, idx), excl); }); }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        if (field == other) { if (excl) bm.reset_bits(); return; }
        sv v = bs.const_value(field), vo = bs.const_value(other);
        if (!v.empty() || !vo.empty()) {
            if (!v.empty() && !vo.empty()) { if (!le_values_string(v, vo, excl)) bm.reset_bits(); return; }
            by_strings(bs, bm); return;
        }
        const Column* ch = bs.column(field); const Column* co = bs.column(other);
        if (!ch || !co) { if (!ch && !co) { if (excl) bm.reset_bits(); return; } by_strings(bs, bm); return; }
        if (ch->valueType != co->valueType || ch->valueType == VT_STRING) { by_strings(bs, bm); return; }
        if (bm.is_zero()) return;
        const auto& a = bs.values(ch); const auto& b = bs.values(co);
        switch (ch->valueType) {
        case VT_DICT: bm.for_each_set_bit([&](uint64_t i) { return le_values_string(ch->dict.at((uint8_t)a[i][0]), co->dict.at((uint8_t)b[i][0]), excl); }); break;
        case VT_INT64: bm.for_each_set_bit([&](uint64_t i) { int64_t x = unzigzag(get_be64((const uint8_t*)a[i].data())), y = unzigzag(get_be64((const uint8_t*)b[i].data())); return excl ? x < y : x <= y; }); break;
        case VT_FLOAT64: bm.for_each_set_bit([&](uint64_t i) { uint64_t p = get_be64((const uint8_t*)a[i].data()), q = get_be64((const uint8_t*)b[i].data()); double x, y; memcpy(&x, &p, 8); memcpy(&y, &q, 8); return excl ? x < y : x <= y; }); break;
        default:   // applyFilterUint :246-252: uint8..uint64, ipv4, iso8601 compare their big-endian encodings AS STRINGS through leValuesString
            bm.for_each_set_bit([&](uint64_t i) { return le_values_string(a[i], b[i], excl); });
        }
    }
};

struct FilterTime : Filter {   // filter_time.go:14-137  (_time:[min, max], nanoseconds, both ends inclusive)
    int64_t minTimestamp, maxTimestamp;
    FilterTime(int64_t mn, int64_t mx) : minTimestamp(mn), maxTimestamp(mx) { kind = F_TIME; }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        if (minTimestamp > maxTimestamp) { bm.reset_bits(); return; }
        if (!bs.b->hasTimestamps) throw std::runtime_error("the block has no timestamps");
        if (minTimestamp > bs.b->maxTimestamp || maxTimestamp < bs.b->minTimestamp) { bm.reset_bits(); return; }   // header-level prune
        if (minTimestamp <= bs.b->minTimestamp && maxTimestamp >= bs.b->maxTimestamp) return;                     // the whole block is inside
        if (bm.is_zero()) return;
        const auto& t = bs.timestamps();
        bm.for_each_set_bit([&](uint64_t idx) { return t[idx] >= minTimestamp && t[idx] <= maxTimestamp; });
    }
};

struct FilterDayRange : Filter {   // filter_day_range.go:13-124  (_time:day_range[hh:mm, hh:mm] offset ...): nanoseconds inside the day
    int64_t start, end, offset;
    FilterDayRange(int64_t s, int64_t e, int64_t o) : start(s), end(e), offset(o) { kind = F_DAY_RANGE; }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        if (start > end) { bm.reset_bits(); return; }
        if (start == 0 && end == 86400000000000LL - 1) return;
        if (bm.is_zero()) return;
        const auto& t = bs.timestamps();
        bm.for_each_set_bit([&](uint64_t idx) { int64_t d = (t[idx] - offset) % 86400000000000LL; return d >= start && d <= end; });   // Go's % truncates like C++'s
    }
};
struct FilterWeekRange : Filter {   // filter_week_range.go:14-126: time.Weekday (Sunday = 0) of the UTC date
    int startDay, endDay; int64_t offset;
    FilterWeekRange(int s, int e, int64_t o) : startDay(s), endDay(e), offset(o) { kind = F_WEEK_RANGE; }
    static int weekday(int64_t ts) {
        int64_t day = ts / 86400000000000LL; if (ts % 86400000000000LL < 0) day--;   // floor: time.Unix(0, ts).UTC()
        int64_t w = (day + 4) % 7; if (w < 0) w += 7;                                 // 1970-01-01 was a Thursday
        return (int)w;
    }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        if (startDay > endDay) { bm.reset_bits(); return; }
        if (startDay <= 0 && endDay >= 6) return;
        if (bm.is_zero()) return;
        const auto& t = bs.timestamps();
        bm.for_each_set_bit([&](uint64_t idx) { int d = weekday(t[idx] - offset); return d >= startDay && d <= endDay; });
    }
};

struct FilterRegexp : Filter {   // filter_regexp.go:17-254
    std::string field; Regex re; std::vector<std::string> tokens; std::vector<uint64_t> hashes;
    FilterRegexp(sv f, sv expr) : field(f), re(expr) {
        kind = F_REGEXP;
        std::vector<std::string> lits = re.get_literals();
        std::vector<std::string> stripped; for (auto& l : lits) stripped.push_back(skip_first_last_token(l));
        std::vector<sv> views(stripped.begin(), stripped.end());
        tokens = tokenize_strings(views);
        hashes = tokens_hashes(tokens);
    }
    bool leaf_tokens(std::string* f, std::vector<std::string>* t) override { *f = field; *t = tokens; return true; }
    void apply(BlockSearch& bs, Bitmap& bm) override {
        sv v = bs.const_value(field);
        if (!v.empty()) { if (!re.match_string(v)) bm.reset_bits(); return; }
        const Column* ch = bs.column(field);
        if (!ch) { if (!re.match_string("")) bm.reset_bits(); return; }
        if (ch->valueType == VT_DICT) {
            std::vector<uint8_t> lut; for (auto& d : ch->dict) lut.push_back(re.match_string(d) ? 1 : 0);
            bs.match_encoded_dict(ch, bm, lut); return;
        }
        if (!bs.bloom_all(ch, hashes)) { bm.reset_bits(); return; }
        if (ch->valueType == VT_STRING) bs.visit_values(ch, bm, [&](sv x) { return re.match_string(x); });
        else bs.visit_values(ch, bm, [&](sv x) { return re.match_string(encoded_to_string(ch->valueType, x)); });
    }
};

inline bool match_string_by_all_tokens(sv v, const std::vector<std::string>& tokens) {   // filter_and.go:189-196
    for (auto& t : tokens) if (!match_phrase(v, t)) return false;
    return true;
}
inline bool match_dict_values_by_all_tokens(const std::vector<std::string>& dict, const std::vector<std::string>& tokens) {   // :198-208
    std::string joined; for (auto& d : dict) { joined += d; joined.push_back(','); }
    return match_string_by_all_tokens(joined, tokens);
}

struct FilterOr;
struct FilterAnd : Filter {   // filter_and.go:15-187
    std::vector<FP> filters; std::vector<FieldTokens> byField; bool inited = false;
    explicit FilterAnd(std::vector<FP> f) : filters(std::move(f)) { kind = F_AND; }
    const std::vector<FieldTokens>& by_field_tokens();
    void apply(BlockSearch& bs, Bitmap& bm) override {
        if (!match_bloom(bs)) { bm.reset_bits(); return; }
        for (auto& f : filters) { f->apply(bs, bm); if (bm.is_zero()) return; }
    }
    bool match_bloom(BlockSearch& bs) {   // :76-111
        for (auto& ft : by_field_tokens()) {
            sv v = bs.const_value(ft.field);
            if (!v.empty()) { if (match_string_by_all_tokens(v, ft.tokens)) continue; return false; }
            const Column* ch = bs.column(ft.field);
            if (!ch) return false;
            if (ch->valueType == VT_DICT) { if (match_dict_values_by_all_tokens(ch->dict, ft.tokens)) continue; return false; }
            if (!bs.bloom_all(ch, ft.hashes)) return false;
        }
        return true;
    }
};
struct FilterOr : Filter {   // filter_or.go:36-193
    std::vector<FP> filters; std::vector<FieldTokens> byField; bool inited = false;
    explicit FilterOr(std::vector<FP> f) : filters(std::move(f)) { kind = F_OR; }
    const std::vector<FieldTokens>& by_field_tokens();
    void apply(BlockSearch& bs, Bitmap& bm) override {
        if (!match_bloom(bs)) { bm.reset_bits(); return; }
        Bitmap res = bm, tmp;
        for (auto& f : filters) {
            tmp = res;
            f->apply(bs, tmp);
            res.and_not(tmp);
            if (res.is_zero()) return;
        }
        bm.and_not(res);
    }
    bool match_bloom(BlockSearch& bs) {   // :80-115
        auto& bft = by_field_tokens();
        if (bft.empty()) return true;
        for (auto& ft : bft) {
            sv v = bs.const_value(ft.field);
            if (!v.empty()) { if (match_string_by_all_tokens(v, ft.tokens)) return true; continue; }
            const Column* ch = bs.column(ft.field);
            if (!ch) continue;
            if (ch->valueType == VT_DICT) { if (match_dict_values_by_all_tokens(ch->dict, ft.tokens)) return true; continue; }
            if (bs.bloom_all(ch, ft.hashes)) return true;
        }
        return false;
    }
};
struct FilterNot : Filter {   // filter_not.go:11-46
    FP f;
    explicit FilterNot(FP x) : f(std::move(x)) { kind = F_NOT; }
    void apply(BlockSearch& bs, Bitmap& bm) override { Bitmap tmp = bm; f->apply(bs, tmp); bm.and_not(tmp); }
};

inline std::vector<std::string> dedup_keep_order(const std::vector<std::string>& a) {
    std::vector<std::string> o; std::unordered_set<std::string> seen;
    for (auto& t : a) if (seen.insert(t).second) o.push_back(t);
    return o;
}
inline const std::vector<FieldTokens>& FilterAnd::by_field_tokens() {   // getCommonTokensForAndFilters :122-187
    if (inited) return byField;
    inited = true;
    std::map<std::string, std::vector<std::string>> m; std::vector<std::string> names;
    auto merge = [&](const std::string& field, const std::vector<std::string>& tokens) {
        if (tokens.empty()) return;
        std::string fn = canonical(field);
        if (!m.count(fn)) names.push_back(fn);
        auto& v = m[fn]; v.insert(v.end(), tokens.begin(), tokens.end());
    };
    for (auto& f : filters) {
        std::string field; std::vector<std::string> toks;
        if (f->leaf_tokens(&field, &toks)) merge(field, toks);
        else if (f->kind == F_OR) for (auto& bft : static_cast<FilterOr*>(f.get())->by_field_tokens()) merge(bft.field, bft.tokens);
    }
    for (auto& n : names) { FieldTokens ft; ft.field = n; ft.tokens = dedup_keep_order(m[n]); ft.hashes = tokens_hashes(ft.tokens); byField.push_back(ft); }
    return byField;
}
inline const std::vector<FieldTokens>& FilterOr::by_field_tokens() {   // getCommonTokensForOrFilters :126-193
    if (inited) return byField;
    inited = true;
    std::map<std::string, std::vector<std::vector<std::string>>> m; std::vector<std::string> names;
    auto merge = [&](const std::string& field, const std::vector<std::string>& tokens) {
        if (tokens.empty()) return;
        std::string fn = canonical(field);
        if (!m.count(fn)) names.push_back(fn);
        m[fn].push_back(tokens);
    };
    for (auto& f : filters) {
        std::string field; std::vector<std::string> toks;
        if (f->leaf_tokens(&field, &toks)) merge(field, toks);
        else if (f->kind == F_AND) for (auto& bft : static_cast<FilterAnd*>(f.get())->by_field_tokens()) merge(bft.field, bft.tokens);
        else { byField.clear(); return byField; }   // default: cannot extract common tokens
    }
    for (auto& n : names) {
        auto& tokenss = m[n];
        if (tokenss.size() != filters.size()) continue;
        std::vector<std::string> common = tokenss[0];
        for (size_t i = 1; i < tokenss.size(); i++) {
            if (common.empty()) break;
            std::vector<std::string> d;
            for (auto& t : common) if (std::find(tokenss[i].begin(), tokenss[i].end(), t) != tokenss[i].end()) d.push_back(t);
            common = d;
        }
        if (common.empty()) continue;
        FieldTokens ft; ft.field = n; ft.tokens = common; ft.hashes = tokens_hashes(common); byField.push_back(ft);
    }
    return byField;
}

// blockSearch.search block_search.go:207-226 (filter part)
inline void block_search(BlockSearch& bs, const Block& b, Filter& f, Bitmap& bm, ScanStats* st) {
    bs.reset(&b, st);
    bm.init(b.rows);
    bm.set_bits();
    f.apply(bs, bm);
    if (st) { st->blocks++; st->rows += b.rows; if (!bm.is_zero()) st->bitmap_bytes += 8 * bm.a.size(); }
}

}  // namespace vlo
