// ORACLE -- TEST INFRASTRUCTURE ONLY (see vlo_util.h header).
//
// The timestamps column of a block, restated from
//   vm/lib/encoding/encoding.go:13-30,100-385   MarshalType, marshalInt64Array / unmarshalInt64Array, isConst / isDeltaConst / isGauge,
//                                                getCompressLevel (by number of items), minCompressibleBlockSize = 128
//   vm/lib/encoding/nearest_delta2.go, nearest_delta.go   delta-of-delta and delta coding (precisionBits = 64: lossless, the only
//                                                value VictoriaLogs uses: lib/logstorage/block.go:682)
//   vm/lib/encoding/int.go:69-130                 MarshalVarInt64: zig-zag, then 7-bit groups, low group first
//   lib/logstorage/block.go:674-690               timestampsHeader: marshalType, minTimestamp = first, maxTimestamp = last (rows are sorted)
//   lib/logstorage/block_search.go:479-506        getTimestamps
//   lib/logstorage/filter_time.go:114-137         filterTime.applyToBlockSearch
// Parity: the vendored encoding package ships no tests, so the byte format is pinned only by hand-derived vectors
// (tests/test_oracle_next_filters.py) and round trips; filterTime is pinned by filter_time_test.go.
#pragma once
#include "vlo_util.h"

namespace vlo {

enum MarshalType : uint8_t { MT_ZSTD_NEAREST_DELTA2 = 1, MT_DELTA_CONST = 2, MT_CONST = 3, MT_ZSTD_NEAREST_DELTA = 4, MT_NEAREST_DELTA2 = 5, MT_NEAREST_DELTA = 6 };

inline void put_varint64(std::string& dst, int64_t v) { put_varuint(dst, ((uint64_t)v << 1) ^ (uint64_t)(v >> 63)); }
inline int get_varint64(const uint8_t* p, size_t n, int64_t* out) {
    uint64_t u; int k = get_varuint(p, n, &u);
    if (k <= 0) return k;
    *out = (int64_t)(u >> 1) ^ -(int64_t)(u & 1);
    return k;
}
inline int ts_compress_level(size_t items) { return items <= 64 ? 1 : items <= 256 ? 2 : items <= 1024 ? 3 : items <= 4096 ? 4 : 5; }   // encoding.go:371-385

inline bool ts_is_const(const std::vector<int64_t>& a) { for (int64_t v : a) if (v != a[0]) return false; return !a.empty(); }
inline bool ts_is_delta_const(const std::vector<int64_t>& a) {
    if (a.size() < 2) return false;
    int64_t d = (int64_t)((uint64_t)a[1] - (uint64_t)a[0]);
    for (size_t i = 2; i < a.size(); i++) if ((int64_t)((uint64_t)a[i] - (uint64_t)a[i - 1]) != d) return false;
    return true;
}
inline bool ts_is_gauge(const std::vector<int64_t>& a) {   // encoding.go:336-364
    if (a.size() < 2) return false;
    size_t resets = 0; int64_t prev = a[0];
    if (prev < 0) return true;
    for (size_t i = 1; i < a.size(); i++) {
        int64_t v = a[i];
        if (v < prev) { if (v < 0) return true; if (v > (prev >> 3)) return true; resets++; }
        prev = v;
    }
    if (resets <= 2) return false;
    return resets > (a.size() >> 3);
}

struct EncodedInt64s { std::string data; uint8_t mt = 0; int64_t first = 0; };

// marshalInt64Array with precisionBits = 64 (encoding.go:119-160)
inline EncodedInt64s marshal_int64_array(const std::vector<int64_t>& a) {
    if (a.empty()) throw std::runtime_error("BUG: a must contain at least one item");
    EncodedInt64s e; e.first = a[0];
    if (ts_is_const(a)) { e.mt = MT_CONST; return e; }
    if (ts_is_delta_const(a)) { e.mt = MT_DELTA_CONST; put_varint64(e.data, (int64_t)((uint64_t)a[1] - (uint64_t)a[0])); return e; }
    std::string raw;
    if (ts_is_gauge(a)) {       // marshalInt64NearestDelta: deltas
        e.mt = MT_ZSTD_NEAREST_DELTA;
        uint64_t v = (uint64_t)a[0];
        for (size_t i = 1; i < a.size(); i++) { uint64_t d = (uint64_t)a[i] - v; v += d; put_varint64(raw, (int64_t)d); }
    } else {                    // marshalInt64NearestDelta2: first delta, then deltas of deltas
        e.mt = MT_ZSTD_NEAREST_DELTA2;
        uint64_t d1 = (uint64_t)a[1] - (uint64_t)a[0];
        put_varint64(raw, (int64_t)d1);
        uint64_t v = (uint64_t)a[1];
        for (size_t i = 2; i < a.size(); i++) { uint64_t d2 = (uint64_t)a[i] - v - d1; d1 += d2; v += d1; put_varint64(raw, (int64_t)d2); }
    }
    bool plain = raw.size() < 128;
    if (!plain) {
        size_t bound = ZSTD_compressBound(raw.size());
        std::string tmp(bound, '\0');
        size_t n = ZSTD_compress(tmp.data(), bound, raw.data(), raw.size(), ts_compress_level(a.size()));
        if (ZSTD_isError(n)) throw std::runtime_error("zstd compress failed");
        if ((double)n > 0.9 * (double)raw.size()) plain = true; else e.data.assign(tmp.data(), n);
    }
    if (plain) { e.mt = e.mt == MT_ZSTD_NEAREST_DELTA2 ? MT_NEAREST_DELTA2 : MT_NEAREST_DELTA; e.data = raw; }
    return e;
}

// unmarshalInt64Array (encoding.go:162-254)
inline std::vector<int64_t> unmarshal_int64_array(sv src, uint8_t mt, int64_t first, size_t items) {
    std::vector<int64_t> out; out.reserve(items);
    std::string raw;
    if (mt == MT_ZSTD_NEAREST_DELTA || mt == MT_ZSTD_NEAREST_DELTA2) {
        unsigned long long dlen = ZSTD_getFrameContentSize(src.data(), src.size());
        if (dlen == (unsigned long long)-1 || dlen == (unsigned long long)-2 || dlen > (1ull << 30)) throw std::runtime_error("cannot decompress zstd data");
        raw.resize(dlen);
        size_t got = ZSTD_decompress(raw.data(), dlen, src.data(), src.size());
        if (ZSTD_isError(got) || got != dlen) throw std::runtime_error("cannot decompress zstd data");
        src = raw; mt = mt == MT_ZSTD_NEAREST_DELTA ? MT_NEAREST_DELTA : MT_NEAREST_DELTA2;
    }
    const uint8_t* p = (const uint8_t*)src.data(); size_t n = src.size();
    auto next = [&](int64_t* v) { int k = get_varint64(p, n, v); if (k <= 0) throw std::runtime_error("cannot unmarshal varint"); p += k; n -= (size_t)k; };
    switch (mt) {
    case MT_CONST:
        if (!src.empty()) throw std::runtime_error("unexpected data left in const encoding");
        out.assign(items, first); return out;
    case MT_DELTA_CONST: {
        int64_t d; next(&d);
        if (n) throw std::runtime_error("unexpected trailing data after delta const");
        uint64_t v = (uint64_t)first;
        for (size_t i = 0; i < items; i++) { out.push_back((int64_t)v); v += (uint64_t)d; }
        return out;
    }
    case MT_NEAREST_DELTA: {
        if (items < 1) throw std::runtime_error("BUG: itemsCount must be greater than 0");
        uint64_t v = (uint64_t)first; out.push_back(first);
        for (size_t i = 1; i < items; i++) { int64_t d; next(&d); v += (uint64_t)d; out.push_back((int64_t)v); }
        if (n) throw std::runtime_error("unexpected tail left after unmarshaling");
        return out;
    }
    case MT_NEAREST_DELTA2: {
        if (items < 2) throw std::runtime_error("BUG: itemsCount must be greater than 1");
        int64_t d; next(&d);
        uint64_t d1 = (uint64_t)d, v = (uint64_t)first;
        out.push_back(first); v += d1; out.push_back((int64_t)v);
        for (size_t i = 2; i < items; i++) { int64_t d2; next(&d2); d1 += (uint64_t)d2; v += d1; out.push_back((int64_t)v); }
        if (n) throw std::runtime_error("unexpected tail left after unmarshaling");
        return out;
    }
    }
    throw std::runtime_error("unknown MarshalType");
}

}  // namespace vlo
