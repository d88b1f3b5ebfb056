// ORACLE -- TEST INFRASTRUCTURE ONLY (see vlo_util.h header).
//
// The on-disk part directory, restated from
//   lib/logstorage/filenames.go:3-24              file names
//   lib/logstorage/part_header.go:15-96           metadata.json (partHeader) and its validation
//   lib/logstorage/column_names.go:14-160         column_names.bin (ZSTD(varuint n, n x bytes)), column_idxs.bin (columnID -> shard)
//   lib/logstorage/index_block_header.go:13-175   metaindex.bin = ZSTD(56-byte indexBlockHeader records), index.bin = ZSTD blocks of blockHeaders
//   lib/logstorage/block_header.go:14-1014        blockHeader, columnsHeaderIndex, columnsHeader, columnHeader, timestampsHeader
//   lib/logstorage/stream_id.go:67-89, tenant_id.go:54-75, u128.go:60-74   streamID = accountID u32, projectID u32, hi u64, lo u64 (big endian)
//   lib/logstorage/rows.go:35-68                  Field.marshal (const columns: value only since format v1)
//   lib/logstorage/values_encoder.go:1289-1322    valuesDict.marshal
//   lib/logstorage/block_stream_writer.go:181-470 streamWriters (column -> bloom/values shard), blockStreamWriter.mustWriteBlockInternal / Finalize
//   lib/logstorage/block.go:134-175,457-482,678-692   column.mustWriteTo, block.mustWriteTo, mustWriteTimestampsTo
//   lib/logstorage/part.go:105-226                mustOpenFilePart, getBloomValuesFileForColumnName
//   lib/logstorage/block_search.go:232-420        getConstColumnValue / getColumnHeader / columns header (index) blocks
// Only the latest format (partFormatLatestVersion = 3, consts.go:6) is written; the reader also accepts v1/v2 shard selection.
// Parity: the reference ships no part files; the record layouts are pinned by the marshaled lengths its tests assert
// (block_header_test.go:29-56,75-94,120-149,410-440,467-479,548-557, index_block_header_test.go:27-44,127-165) and by round trips.
#pragma once
#include "vlo_block.h"
#include <map>

namespace vlo {

static const unsigned partFormatLatestVersion = 3;
static const uint64_t bloomValuesMaxShardsCount = 128, maxUncompressedIndexBlockSize = 128 * 1024, maxRowsPerBlock = 8 * 1024 * 1024,
                      maxColumnsPerBlock = 2000, maxIndexBlockSize = 8 << 20, maxTimestampsBlockSize = 8 << 20, maxValuesBlockSize = 8 << 20,
                      maxBloomFilterBlockSize = 8 << 20, maxColumnsHeaderSize = 8 << 20, maxColumnsHeaderIndexSize = 8 << 20;

struct PartError : std::runtime_error { using std::runtime_error::runtime_error; };

// a cursor over untrusted bytes
struct PReader {
    const uint8_t* p; size_t n;
    explicit PReader(sv s) : p((const uint8_t*)s.data()), n(s.size()) {}
    uint64_t varuint(const char* what) { uint64_t v; int k = get_varuint(p, n, &v); if (k <= 0) throw PartError(std::string("cannot unmarshal ") + what); p += k; n -= (size_t)k; return v; }
    sv bytes(const char* what) { uint64_t l = varuint(what); if (l > n) throw PartError(std::string("cannot unmarshal ") + what); sv s((const char*)p, l); p += l; n -= l; return s; }
    const uint8_t* fixed(size_t k, const char* what) { if (n < k) throw PartError(std::string("cannot unmarshal ") + what + ": too few bytes"); const uint8_t* q = p; p += k; n -= k; return q; }
    uint64_t be64(const char* what) { return get_be64(fixed(8, what)); }
    uint32_t be32(const char* what) { return get_be32(fixed(4, what)); }
    uint16_t be16(const char* what) { return get_be16(fixed(2, what)); }
    uint8_t u8(const char* what) { return *fixed(1, what); }
};
inline void put_bytes(std::string& d, sv s) { put_varuint(d, s.size()); d.append(s); }   // encoding.MarshalBytes

inline std::string zstd_compress_level(sv src, int level) {
    size_t bound = ZSTD_compressBound(src.size());
    std::string tmp(bound, '\0');
    size_t n = ZSTD_compress(tmp.data(), bound, src.data(), src.size(), level);
    if (ZSTD_isError(n)) throw std::runtime_error("zstd compress failed");
    tmp.resize(n); return tmp;
}
inline std::string zstd_decompress_all(sv src, const char* what) {
    unsigned long long dlen = ZSTD_getFrameContentSize(src.data(), src.size());
    if (dlen == (unsigned long long)-1 || dlen == (unsigned long long)-2 || dlen > (1ull << 31)) throw PartError(std::string("cannot decompress ") + what);
    std::string out(dlen, '\0');
    size_t got = ZSTD_decompress(out.data(), dlen, src.data(), src.size());
    if (ZSTD_isError(got) || got != dlen) throw PartError(std::string("cannot decompress ") + what);
    return out;
}

// ---- streamID --------------------------------------------------------------------------------------------------------
struct StreamID {
    uint32_t accountID = 0, projectID = 0; uint64_t hi = 0, lo = 0;
    void marshal(std::string& d) const { put_be32(d, accountID); put_be32(d, projectID); put_be64(d, hi); put_be64(d, lo); }
    void unmarshal(PReader& r) { accountID = r.be32("accountID"); projectID = r.be32("projectID"); hi = r.be64("streamID.hi"); lo = r.be64("streamID.lo"); }
    bool equal(const StreamID& a) const { return accountID == a.accountID && projectID == a.projectID && hi == a.hi && lo == a.lo; }
    bool less(const StreamID& a) const {   // stream_id.go:49-57, tenant_id.go:30-39, u128.go:19-27
        if (accountID != a.accountID) return accountID < a.accountID;
        if (projectID != a.projectID) return projectID < a.projectID;
        if (hi != a.hi) return hi < a.hi;
        return lo < a.lo;
    }
};

// ---- timestampsHeader: 33 bytes (block_header.go:919-975) ---------------------------------------------------------------
struct TimestampsHeader {
    uint64_t blockOffset = 0, blockSize = 0; int64_t minTimestamp = 0, maxTimestamp = 0; uint8_t marshalType = 0;
    void marshal(std::string& d) const { put_be64(d, blockOffset); put_be64(d, blockSize); put_be64(d, (uint64_t)minTimestamp); put_be64(d, (uint64_t)maxTimestamp); d.push_back((char)marshalType); }
    void unmarshal(PReader& r) {
        if (r.n < 33) throw PartError("cannot unmarshal timestampsHeader; need at least 33 bytes");
        blockOffset = r.be64("blockOffset"); blockSize = r.be64("blockSize");
        minTimestamp = (int64_t)r.be64("minTimestamp"); maxTimestamp = (int64_t)r.be64("maxTimestamp"); marshalType = r.u8("marshalType");
    }
};

// ---- columnHeader (block_header.go:545-917) ----------------------------------------------------------------------------
struct ColumnHeader {
    std::string name; uint8_t valueType = 0; uint64_t minValue = 0, maxValue = 0; std::vector<std::string> dict;
    uint64_t valuesOffset = 0, valuesSize = 0, bloomFilterOffset = 0, bloomFilterSize = 0;

    void marshal(std::string& d) const {
        d.push_back((char)valueType);
        auto values = [&] { put_varuint(d, valuesOffset); put_varuint(d, valuesSize); };
        auto values_bloom = [&] { values(); put_varuint(d, bloomFilterOffset); put_varuint(d, bloomFilterSize); };
        switch (valueType) {
        case VT_STRING: values_bloom(); break;
        case VT_DICT:
            if (dict.size() > maxDictLen) throw std::runtime_error("BUG: valuesDict may contain max 8 items");
            d.push_back((char)dict.size()); for (auto& v : dict) put_bytes(d, v);
            values(); break;
        case VT_UINT8: d.push_back((char)minValue); d.push_back((char)maxValue); values_bloom(); break;
        case VT_UINT16: put_be16(d, (uint16_t)minValue); put_be16(d, (uint16_t)maxValue); values_bloom(); break;
        case VT_UINT32: case VT_IPV4: put_be32(d, (uint32_t)minValue); put_be32(d, (uint32_t)maxValue); values_bloom(); break;
        case VT_UINT64: case VT_FLOAT64: case VT_ISO8601: put_be64(d, minValue); put_be64(d, maxValue); values_bloom(); break;
        case VT_INT64: put_be64(d, zigzag((int64_t)minValue)); put_be64(d, zigzag((int64_t)maxValue)); values_bloom(); break;   // encoding.MarshalInt64
        default: throw std::runtime_error("BUG: unknown valueType");
        }
    }
    void unmarshal(PReader& r, unsigned formatVersion) {
        *this = ColumnHeader();
        if (formatVersion < 1) name = std::string(r.bytes("column name"));
        valueType = r.u8("valueType");
        auto values = [&] {
            valuesOffset = r.varuint("valuesOffset"); valuesSize = r.varuint("valuesSize");
            if (valuesSize > maxValuesBlockSize) throw PartError("too big valuesSize");
        };
        auto values_bloom = [&] {
            values();
            bloomFilterOffset = r.varuint("bloomFilterOffset"); bloomFilterSize = r.varuint("bloomFilterSize");
            if (bloomFilterSize > maxBloomFilterBlockSize) throw PartError("too big bloomFilterSize");
        };
        switch (valueType) {
        case VT_STRING: values_bloom(); break;
        case VT_DICT: { int n = r.u8("dict len"); for (int i = 0; i < n; i++) dict.emplace_back(r.bytes("dict value")); values(); break; }
        case VT_UINT8: minValue = r.u8("min"); maxValue = r.u8("max"); values_bloom(); break;
        case VT_UINT16: minValue = r.be16("min"); maxValue = r.be16("max"); values_bloom(); break;
        case VT_UINT32: case VT_IPV4: minValue = r.be32("min"); maxValue = r.be32("max"); values_bloom(); break;
        case VT_UINT64: case VT_FLOAT64: case VT_ISO8601: minValue = r.be64("min"); maxValue = r.be64("max"); values_bloom(); break;
        case VT_INT64: minValue = (uint64_t)unzigzag(r.be64("min")); maxValue = (uint64_t)unzigzag(r.be64("max")); values_bloom(); break;
        default: throw PartError("unexpected valueType=" + std::to_string(valueType));
        }
    }
};

// ---- columnsHeaderIndex / columnsHeader (block_header.go:222-544) ---------------------------------------------------------
struct ColumnHeaderRef { uint64_t columnNameID = 0, offset = 0; };
struct ColumnsHeaderIndex {
    std::vector<ColumnHeaderRef> columnHeadersRefs, constColumnsRefs;
    static void put_refs(std::string& d, const std::vector<ColumnHeaderRef>& v) { put_varuint(d, v.size()); for (auto& x : v) { put_varuint(d, x.columnNameID); put_varuint(d, x.offset); } }
    static void get_refs(PReader& r, std::vector<ColumnHeaderRef>& v) {
        uint64_t n = r.varuint("the number of columnHeaderRef items");
        if (n > r.n) throw PartError("too many columnHeaderRef items");
        v.clear();
        for (uint64_t i = 0; i < n; i++) { ColumnHeaderRef x; x.columnNameID = r.varuint("column name ID"); x.offset = r.varuint("offset"); v.push_back(x); }
    }
    void marshal(std::string& d) const { put_refs(d, columnHeadersRefs); put_refs(d, constColumnsRefs); }
    void unmarshal(sv src) {
        PReader r(src); get_refs(r, columnHeadersRefs); get_refs(r, constColumnsRefs);
        if (r.n) throw PartError("unexpected non-empty tail left after unmarshaling columnsHeaderIndex");
    }
};

struct ColumnNameIDGenerator {   // column_names.go:162-187
    std::map<std::string, uint64_t> ids; std::vector<std::string> names;
    uint64_t get(const std::string& name) { auto it = ids.find(name); if (it != ids.end()) return it->second; uint64_t id = names.size(); ids[name] = id; names.push_back(name); return id; }
};

struct ColumnsHeader {
    std::vector<ColumnHeader> columnHeaders; std::vector<ConstColumn> constColumns;
    // names are the raw on-disk ones ("" for the message field)
    void marshal(std::string& d, ColumnsHeaderIndex& idx, ColumnNameIDGenerator& g) const {
        size_t base = d.size();
        idx.columnHeadersRefs.clear(); idx.constColumnsRefs.clear();
        put_varuint(d, columnHeaders.size());
        for (auto& ch : columnHeaders) { idx.columnHeadersRefs.push_back({g.get(ch.name), d.size() - base}); ch.marshal(d); }
        put_varuint(d, constColumns.size());
        for (auto& cc : constColumns) { idx.constColumnsRefs.push_back({g.get(cc.name), d.size() - base}); put_bytes(d, cc.value); }
    }
    void unmarshal(sv src, unsigned formatVersion) {
        PReader r(src);
        uint64_t n = r.varuint("columnHeaders len");
        if (n > 1000000) throw PartError("too big number of columnHeaders");
        columnHeaders.assign(n, ColumnHeader());
        for (auto& ch : columnHeaders) ch.unmarshal(r, formatVersion);
        if (columnHeaders.size() > maxColumnsPerBlock) throw PartError("too many column headers");
        n = r.varuint("constColumns len");
        if (n > 1000000) throw PartError("too big number of constColumns");
        constColumns.assign(n, ConstColumn());
        for (auto& cc : constColumns) { if (formatVersion < 1) cc.name = std::string(r.bytes("field name")); cc.value = std::string(r.bytes("field value")); }
        if (constColumns.size() + columnHeaders.size() > maxColumnsPerBlock) throw PartError("too many columns");
        if (r.n) throw PartError("unexpected non-empty tail left after unmarshaling columnsHeader");
    }
    void set_column_names(const ColumnsHeaderIndex& idx, const std::vector<std::string>& names) {
        if (idx.columnHeadersRefs.size() != columnHeaders.size()) throw PartError("unexpected number of column headers");
        for (size_t i = 0; i < columnHeaders.size(); i++) { uint64_t id = idx.columnHeadersRefs[i].columnNameID; if (id >= names.size()) throw PartError("unexpected columnNameID in columnHeadersRef"); columnHeaders[i].name = names[id]; }
        if (idx.constColumnsRefs.size() != constColumns.size()) throw PartError("unexpected number of const columns");
        for (size_t i = 0; i < constColumns.size(); i++) { uint64_t id = idx.constColumnsRefs[i].columnNameID; if (id >= names.size()) throw PartError("unexpected columnNameID in constColumnsRefs"); constColumns[i].name = names[id]; }
    }
};

// ---- blockHeader (block_header.go:14-221) --------------------------------------------------------------------------------
struct BlockHeader {
    StreamID streamID; uint64_t uncompressedSizeBytes = 0, rowsCount = 0; TimestampsHeader timestampsHeader;
    uint64_t columnsHeaderIndexOffset = 0, columnsHeaderIndexSize = 0, columnsHeaderOffset = 0, columnsHeaderSize = 0;
    void marshal(std::string& d) const {
        streamID.marshal(d); put_varuint(d, uncompressedSizeBytes); put_varuint(d, rowsCount); timestampsHeader.marshal(d);
        put_varuint(d, columnsHeaderIndexOffset); put_varuint(d, columnsHeaderIndexSize); put_varuint(d, columnsHeaderOffset); put_varuint(d, columnsHeaderSize);
    }
    void unmarshal(PReader& r, unsigned formatVersion) {
        *this = BlockHeader();
        streamID.unmarshal(r);
        uncompressedSizeBytes = r.varuint("uncompressedSizeBytes");
        rowsCount = r.varuint("rowsCount");
        if (rowsCount > maxRowsPerBlock) throw PartError("too big value for rowsCount");
        timestampsHeader.unmarshal(r);
        if (formatVersion >= 1) {
            // no limit here: the reference checks columnsHeaderIndexSize when it reads the block (readColumnsHeaderIndexBlock block_search.go:381-393)
            columnsHeaderIndexOffset = r.varuint("columnsHeaderIndexOffset"); columnsHeaderIndexSize = r.varuint("columnsHeaderIndexSize");
        }
        columnsHeaderOffset = r.varuint("columnsHeaderOffset"); columnsHeaderSize = r.varuint("columnsHeaderSize");
        if (columnsHeaderSize > maxColumnsHeaderSize) throw PartError("too big value for columnsHeaderSize");
    }
};
inline std::vector<BlockHeader> unmarshal_block_headers(sv src, unsigned formatVersion) {   // :167-204
    std::vector<BlockHeader> out; PReader r(src);
    while (r.n) { out.emplace_back(); out.back().unmarshal(r, formatVersion); }
    for (size_t i = 1; i < out.size(); i++) {
        const BlockHeader &c = out[i], &p = out[i - 1];
        if (c.streamID.less(p.streamID)) throw PartError("unexpected blockHeader with smaller streamID after bigger streamID");
        if (c.streamID.equal(p.streamID) && c.timestampsHeader.minTimestamp < p.timestampsHeader.minTimestamp) throw PartError("unexpected blockHeader with smaller timestamp after bigger timestamp");
    }
    return out;
}

// ---- indexBlockHeader: 56 bytes (index_block_header.go:13-104) -----------------------------------------------------------
struct IndexBlockHeader {
    StreamID streamID; int64_t minTimestamp = 0, maxTimestamp = 0; uint64_t indexBlockOffset = 0, indexBlockSize = 0;
    void marshal(std::string& d) const { streamID.marshal(d); put_be64(d, (uint64_t)minTimestamp); put_be64(d, (uint64_t)maxTimestamp); put_be64(d, indexBlockOffset); put_be64(d, indexBlockSize); }
    void unmarshal(PReader& r) {
        streamID.unmarshal(r);
        if (r.n < 32) throw PartError("cannot unmarshal indexBlockHeader; need at least 32 bytes");
        minTimestamp = (int64_t)r.be64("minTimestamp"); maxTimestamp = (int64_t)r.be64("maxTimestamp"); indexBlockOffset = r.be64("indexBlockOffset"); indexBlockSize = r.be64("indexBlockSize");
    }
};
inline std::vector<IndexBlockHeader> unmarshal_index_block_headers(sv src) {   // :143-175
    std::vector<IndexBlockHeader> out; PReader r(src);
    while (r.n) { out.emplace_back(); out.back().unmarshal(r); }
    for (size_t i = 1; i < out.size(); i++) if (out[i].streamID.less(out[i - 1].streamID)) throw PartError("unexpected indexBlockHeader with smaller streamID after bigger streamID");
    return out;
}

// ---- column_names.bin / column_idxs.bin (column_names.go:14-160) ---------------------------------------------------------
inline std::string marshal_column_names(const std::vector<std::string>& names) {
    std::string data; put_varuint(data, names.size()); for (auto& s : names) put_bytes(data, s);
    return zstd_compress_level(data, 1);
}
inline std::vector<std::string> unmarshal_column_names(sv src) {
    std::string data = zstd_decompress_all(src, "column names"); PReader r(data);
    uint64_t n = r.varuint("the number of column names");
    if (n > r.n) throw PartError("too many distinct column names");
    std::vector<std::string> names; std::map<std::string, uint64_t> seen;
    for (uint64_t id = 0; id < n; id++) {
        std::string s(r.bytes("column name"));
        if (!seen.emplace(s, id).second) throw PartError("duplicate ids for column name " + s);
        names.push_back(std::move(s));
    }
    if (r.n) throw PartError("unexpected non-empty tail left after unmarshaling column name ids");
    return names;
}
inline std::string marshal_column_idxs(const std::map<uint64_t, uint64_t>& idxs) {
    std::string d; put_varuint(d, idxs.size()); for (auto& kv : idxs) { put_varuint(d, kv.first); put_varuint(d, kv.second); }
    return d;
}
inline std::map<std::string, uint64_t> unmarshal_column_idxs(sv src, const std::vector<std::string>& names, uint64_t shardsCount) {
    PReader r(src); uint64_t n = r.varuint("the number of entries");
    if (n > r.n) throw PartError("too many entries");
    std::map<std::string, uint64_t> out;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t id = r.varuint("columnID"), shard = r.varuint("shardIdx");
        if (shard >= shardsCount) throw PartError("too big shardIdx");
        if (id >= names.size()) throw PartError("too big columnID");
        out[names[id]] = shard;
    }
    if (r.n) throw PartError("unexpected tail left after reading column indexes");
    return out;
}

// ---- metadata.json (part_header.go:15-96) ---------------------------------------------------------------------------------
struct PartHeader {
    uint64_t FormatVersion = 0, CompressedSizeBytes = 0, UncompressedSizeBytes = 0, RowsCount = 0, BlocksCount = 0; int64_t MinTimestamp = 0, MaxTimestamp = 0; uint64_t BloomValuesShardsCount = 0;
    std::string to_json() const {   // encoding/json field order = declaration order
        return "{\"FormatVersion\":" + std::to_string(FormatVersion) + ",\"CompressedSizeBytes\":" + std::to_string(CompressedSizeBytes) + ",\"UncompressedSizeBytes\":" + std::to_string(UncompressedSizeBytes) +
               ",\"RowsCount\":" + std::to_string(RowsCount) + ",\"BlocksCount\":" + std::to_string(BlocksCount) + ",\"MinTimestamp\":" + std::to_string(MinTimestamp) + ",\"MaxTimestamp\":" + std::to_string(MaxTimestamp) +
               ",\"BloomValuesShardsCount\":" + std::to_string(BloomValuesShardsCount) + "}";
    }
    // a flat JSON object with integer members; unknown members with scalar values are skipped
    static PartHeader from_json(sv s) {
        PartHeader ph; size_t i = 0;
        auto ws = [&] { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) i++; };
        auto expect = [&](char c) { ws(); if (i >= s.size() || s[i] != c) throw PartError("cannot parse metadata.json"); i++; };
        auto str = [&] { expect('"'); std::string k; while (i < s.size() && s[i] != '"') { if (s[i] == '\\') i++; if (i < s.size()) k.push_back(s[i++]); } expect('"'); return k; };
        expect('{'); ws();
        if (i < s.size() && s[i] == '}') { i++; }
        else for (;;) {
            std::string key = str(); expect(':'); ws();
            if (i < s.size() && s[i] == '"') { str(); }
            else {
                size_t st = i; while (i < s.size() && (s[i] == '-' || s[i] == '+' || s[i] == '.' || s[i] == 'e' || s[i] == 'E' || (s[i] >= '0' && s[i] <= '9') || (s[i] >= 'a' && s[i] <= 'z'))) i++;
                sv num = s.substr(st, i - st);
                if (num.empty()) throw PartError("cannot parse metadata.json");
                bool neg = num[0] == '-'; uint64_t u = 0; bool isint = true;
                for (size_t k = neg ? 1 : 0; k < num.size(); k++) { if (num[k] < '0' || num[k] > '9') { isint = false; break; } u = u * 10 + (uint64_t)(num[k] - '0'); }
                int64_t iv = neg ? -(int64_t)u : (int64_t)u;
                auto want_int = [&] { if (!isint) throw PartError("cannot parse metadata.json: " + key + " must be an integer"); };
                auto want_uint = [&] { want_int(); if (neg) throw PartError("cannot parse metadata.json: " + key + " must be unsigned"); };
                if (key == "FormatVersion") { want_uint(); ph.FormatVersion = u; }
                else if (key == "CompressedSizeBytes") { want_uint(); ph.CompressedSizeBytes = u; }
                else if (key == "UncompressedSizeBytes") { want_uint(); ph.UncompressedSizeBytes = u; }
                else if (key == "RowsCount") { want_uint(); ph.RowsCount = u; }
                else if (key == "BlocksCount") { want_uint(); ph.BlocksCount = u; }
                else if (key == "MinTimestamp") { want_int(); ph.MinTimestamp = iv; }
                else if (key == "MaxTimestamp") { want_int(); ph.MaxTimestamp = iv; }
                else if (key == "BloomValuesShardsCount") { want_uint(); ph.BloomValuesShardsCount = u; }
            }
            ws();
            if (i < s.size() && s[i] == ',') { i++; continue; }
            expect('}'); break;
        }
        ws(); if (i != s.size()) throw PartError("cannot parse metadata.json: trailing data");
        // mustReadMetadata :62-83
        if (ph.FormatVersion <= 1) {
            if (ph.BloomValuesShardsCount != 0) throw PartError("unexpected BloomValuesShardsCount for FormatVersion<=1");
            if (ph.FormatVersion == 1) ph.BloomValuesShardsCount = 8;
        }
        if (ph.FormatVersion > partFormatLatestVersion) throw PartError("unsupported part format version");
        if (ph.MinTimestamp > ph.MaxTimestamp) throw PartError("MinTimestamp cannot exceed MaxTimestamp");
        if (ph.BlocksCount > ph.RowsCount) throw PartError("BlocksCount cannot exceed RowsCount");
        return ph;
    }
};

// ---- the files of one part ---------------------------------------------------------------------------------------------
using PartFiles = std::map<std::string, std::string>;   // file name -> contents
inline std::string raw_field_name(sv name) { return name == "_msg" ? std::string() : std::string(name); }   // getCanonicalFieldName log_rows.go:508-513
inline std::string bloom_file(uint64_t shard) { return "bloom.bin" + std::to_string(shard); }     // part.go:219-225
inline std::string values_file(uint64_t shard) { return "values.bin" + std::to_string(shard); }

// blockStreamWriter for a file part (block_stream_writer.go:213-470)
struct PartWriter {
    PartFiles files; PartHeader ph;
    ColumnNameIDGenerator gen; std::map<uint64_t, uint64_t> columnIdxs; uint64_t nextColumnIdx = 0, shards = 0, maxShards = bloomValuesMaxShardsCount;
    StreamID sidLast, sidFirst; int64_t minTimestampLast = 0, minTimestamp = 0, maxTimestamp = 0; bool hasWrittenBlocks = false;
    std::string indexBlockData, metaindexData;
    uint64_t maxIndexBlock = maxUncompressedIndexBlockSize;   // tests lower it to get several index blocks out of few rows

    PartWriter() {
        for (const char* f : {"column_names.bin", "column_idxs.bin", "metaindex.bin", "index.bin", "columns_header_index.bin", "columns_header.bin", "timestamps.bin", "message_bloom.bin", "message_values.bin"}) files[f];
    }
    // streamWriters.getBloomValuesWriterForColumnName :181-211; returns {bloom file, values file}
    std::pair<std::string*, std::string*> bloom_values_for(const std::string& rawName) {
        if (rawName.empty()) return {&files["message_bloom.bin"], &files["message_values.bin"]};
        uint64_t id = gen.get(rawName);
        auto it = columnIdxs.find(id);
        uint64_t shard;
        if (it != columnIdxs.end()) shard = it->second;
        else {
            shard = nextColumnIdx % maxShards; nextColumnIdx++;
            columnIdxs[id] = shard;
            if (shard >= shards) { if (shard > shards) throw std::runtime_error("BUG: shardIdx must equal the number of shards"); shards++; files[bloom_file(shard)]; files[values_file(shard)]; }
        }
        return {&files[bloom_file(shard)], &files[values_file(shard)]};
    }
    // blockStreamWriter.mustWriteBlockInternal :354-412 with block.mustWriteTo (block.go:457-482)
    void write_block(const StreamID& sid, const Block& b, uint64_t uncompressedSizeBytes = 0) {
        if (b.rows == 0) return;
        if (!b.hasTimestamps) throw std::runtime_error("a block written to a part needs timestamps");
        if (sid.less(sidLast)) throw std::runtime_error("BUG: the sid cannot be smaller than the previously written sid");
        bool had = hasWrittenBlocks;
        if (!had) { sidFirst = sid; hasWrittenBlocks = true; }
        bool seen = sid.equal(sidLast);
        sidLast = sid;

        BlockHeader bh; bh.streamID = sid; bh.uncompressedSizeBytes = uncompressedSizeBytes; bh.rowsCount = b.rows;
        // mustWriteTimestampsTo block.go:678-692
        TimestampsHeader& th = bh.timestampsHeader;
        if (b.ts.data.size() > maxTimestampsBlockSize) throw std::runtime_error("BUG: too big block with timestamps");
        th.marshalType = b.ts.mt; th.minTimestamp = b.ts.first; th.maxTimestamp = b.maxTimestamp;
        th.blockOffset = files["timestamps.bin"].size(); th.blockSize = b.ts.data.size();
        files["timestamps.bin"] += b.ts.data;
        // columns, sorted by raw name (block.sortColumnsByName block.go:379-395)
        ColumnsHeader csh;
        std::vector<const Column*> cols; for (auto& c : b.columns) cols.push_back(&c);
        std::stable_sort(cols.begin(), cols.end(), [](const Column* x, const Column* y) { return raw_field_name(x->name) < raw_field_name(y->name); });
        for (const Column* c : cols) {   // column.mustWriteTo block.go:134-175
            ColumnHeader ch; ch.name = raw_field_name(c->name); ch.valueType = c->valueType; ch.minValue = c->minValue; ch.maxValue = c->maxValue; ch.dict = c->dict;
            auto bv = bloom_values_for(ch.name);
            if (c->valuesBlock.size() > maxValuesBlockSize) throw std::runtime_error("BUG: too big valuesSize");
            ch.valuesOffset = bv.second->size(); ch.valuesSize = c->valuesBlock.size(); *bv.second += c->valuesBlock;
            const std::string empty; const std::string& bloom = c->valueType == VT_DICT ? empty : c->bloom;
            ch.bloomFilterOffset = bv.first->size(); ch.bloomFilterSize = bloom.size(); *bv.first += bloom;
            csh.columnHeaders.push_back(std::move(ch));
        }
        for (auto& cc : b.consts) csh.constColumns.push_back({raw_field_name(cc.name), cc.value});
        std::stable_sort(csh.constColumns.begin(), csh.constColumns.end(), [](const ConstColumn& x, const ConstColumn& y) { return x.name < y.name; });
        if (csh.columnHeaders.size() + csh.constColumns.size() > maxColumnsPerBlock) throw std::runtime_error("BUG: too big number of columns detected in the block");
        // columnsHeader.mustWriteTo block_header.go:382-412
        ColumnsHeaderIndex idx; std::string cshData, idxData;
        csh.marshal(cshData, idx, gen); idx.marshal(idxData);
        bh.columnsHeaderIndexOffset = files["columns_header_index.bin"].size(); bh.columnsHeaderIndexSize = idxData.size(); files["columns_header_index.bin"] += idxData;
        bh.columnsHeaderOffset = files["columns_header.bin"].size(); bh.columnsHeaderSize = cshData.size(); files["columns_header.bin"] += cshData;
        if (bh.columnsHeaderSize > maxColumnsHeaderSize || bh.columnsHeaderIndexSize > maxColumnsHeaderIndexSize) throw std::runtime_error("BUG: too big columns header");

        if (ph.RowsCount == 0 || th.minTimestamp < ph.MinTimestamp) ph.MinTimestamp = th.minTimestamp;
        if (ph.RowsCount == 0 || th.maxTimestamp > ph.MaxTimestamp) ph.MaxTimestamp = th.maxTimestamp;
        if (!had || th.minTimestamp < minTimestamp) minTimestamp = th.minTimestamp;
        if (!had || th.maxTimestamp > maxTimestamp) maxTimestamp = th.maxTimestamp;
        if (seen && th.minTimestamp < minTimestampLast) throw std::runtime_error("BUG: the block cannot contain timestamp smaller than the previous block of the stream");
        minTimestampLast = th.minTimestamp;
        ph.UncompressedSizeBytes += bh.uncompressedSizeBytes; ph.RowsCount += bh.rowsCount; ph.BlocksCount++;

        bh.marshal(indexBlockData);
        if (indexBlockData.size() > maxIndexBlock) { flush_index_block(); indexBlockData.clear(); }
    }
    void flush_index_block() {   // mustFlushIndexBlock :414-423 + indexBlockHeader.mustWriteIndexBlock index_block_header.go:38-51
        if (!indexBlockData.empty()) {
            IndexBlockHeader ih; ih.streamID = sidFirst; ih.minTimestamp = minTimestamp; ih.maxTimestamp = maxTimestamp;
            std::string z = zstd_compress_level(indexBlockData, 1);
            ih.indexBlockOffset = files["index.bin"].size(); ih.indexBlockSize = z.size(); files["index.bin"] += z;
            ih.marshal(metaindexData);
        }
        hasWrittenBlocks = false; minTimestamp = maxTimestamp = 0; sidFirst = StreamID();
    }
    void finalize() {   // Finalize :430-453
        ph.FormatVersion = partFormatLatestVersion; ph.BloomValuesShardsCount = shards;
        flush_index_block(); indexBlockData.clear();
        files["column_names.bin"] = marshal_column_names(gen.names);
        files["column_idxs.bin"] = marshal_column_idxs(columnIdxs);
        files["metaindex.bin"] = zstd_compress_level(metaindexData, 1);
        uint64_t total = 0; for (auto& kv : files) total += kv.second.size();
        ph.CompressedSizeBytes = total;
        files["metadata.json"] = ph.to_json();
    }
};

// part.mustOpenFilePart (part.go:105-173) + the block access of blockSearch (block_search.go:232-474)
struct PartReader {
    const PartFiles* files = nullptr; PartHeader ph;
    std::vector<std::string> columnNames; std::map<std::string, uint64_t> columnNameIDs, columnIdxs;
    std::vector<IndexBlockHeader> indexBlockHeaders;

    const std::string& file(const std::string& name) const { auto it = files->find(name); if (it == files->end()) throw PartError("cannot open " + name); return it->second; }
    static sv read_at(const std::string& f, uint64_t off, uint64_t size, const char* what) {   // fs.MustReadAt
        if (off > f.size() || size > f.size() - off) throw PartError(std::string("cannot read ") + what + ": offset/size outside the file");
        return sv(f).substr(off, size);
    }
    void open(const PartFiles& fs) {
        files = &fs;
        ph = PartHeader::from_json(file("metadata.json"));
        if (ph.FormatVersion >= 1) { columnNames = unmarshal_column_names(file("column_names.bin")); for (size_t i = 0; i < columnNames.size(); i++) columnNameIDs[columnNames[i]] = i; }
        if (ph.FormatVersion >= 3) columnIdxs = unmarshal_column_idxs(file("column_idxs.bin"), columnNames, ph.BloomValuesShardsCount);
        indexBlockHeaders = unmarshal_index_block_headers(zstd_decompress_all(file("metaindex.bin"), "indexBlockHeader entries"));
        file("index.bin"); file("columns_header.bin"); file("timestamps.bin"); file("message_bloom.bin"); file("message_values.bin");
        if (ph.FormatVersion >= 1) { file("columns_header_index.bin"); for (uint64_t i = 0; i < ph.BloomValuesShardsCount; i++) { file(bloom_file(i)); file(values_file(i)); } }
        else { file("field_bloom.bin"); file("field_values.bin"); }
    }
    std::vector<BlockHeader> read_index_block(const IndexBlockHeader& ih) const {   // indexBlockHeader.mustReadNextIndexBlock + unmarshalBlockHeaders
        if (ih.indexBlockSize > maxIndexBlockSize) throw PartError("indexBlockHeader.indexBlockSize is too big");
        std::string data = zstd_decompress_all(read_at(file("index.bin"), ih.indexBlockOffset, ih.indexBlockSize, "index block"), "indexBlock");
        return unmarshal_block_headers(data, (unsigned)ph.FormatVersion);
    }
    // part.getBloomValuesFileForColumnName part.go:194-217
    std::pair<const std::string*, const std::string*> bloom_values_for(const std::string& rawName) const {
        if (rawName.empty()) return {&file("message_bloom.bin"), &file("message_values.bin")};
        if (ph.FormatVersion < 1) return {&file("field_bloom.bin"), &file("field_values.bin")};
        uint64_t shard = 0;
        if (ph.FormatVersion < 3) { uint64_t n = ph.BloomValuesShardsCount; if (n > 1) shard = xxh64(rawName) % n; }
        else { auto it = columnIdxs.find(rawName); if (it == columnIdxs.end()) throw PartError("BUG: unknown shard index for column " + rawName); shard = it->second; }
        return {&file(bloom_file(shard)), &file(values_file(shard))};
    }
    ColumnsHeaderIndex read_columns_header_index(const BlockHeader& bh) const {   // readColumnsHeaderIndexBlock :381-393 + getColumnsHeaderIndex :336-350
        if (bh.columnsHeaderIndexSize > maxColumnsHeaderIndexSize) throw PartError("columns header index size is too big");
        ColumnsHeaderIndex idx; idx.unmarshal(read_at(file("columns_header_index.bin"), bh.columnsHeaderIndexOffset, bh.columnsHeaderIndexSize, "columns header index"));
        return idx;
    }
    ColumnsHeader read_columns_header(const BlockHeader& bh) const {   // blockSearch.getColumnsHeader :352-372
        ColumnsHeader csh;
        csh.unmarshal(read_at(file("columns_header.bin"), bh.columnsHeaderOffset, bh.columnsHeaderSize, "columns header"), (unsigned)ph.FormatVersion);
        if (ph.FormatVersion >= 1) {
            ColumnsHeaderIndex idx = read_columns_header_index(bh);
            csh.set_column_names(idx, columnNames);
        }
        return csh;
    }
    // blockSearch.getColumnHeader :283-324 (format v1+): one header through the index, without parsing the others. Returns false when the
    // block has no such (non-const) column.
    bool get_column_header(const BlockHeader& bh, sv name, ColumnHeader* out) const {
        std::string raw = raw_field_name(name);
        auto it = columnNameIDs.find(raw); if (it == columnNameIDs.end()) return false;
        ColumnsHeaderIndex idx = read_columns_header_index(bh);
        for (auto& cr : idx.columnHeadersRefs) {
            if (cr.columnNameID != it->second) continue;
            sv b = read_at(file("columns_header.bin"), bh.columnsHeaderOffset, bh.columnsHeaderSize, "columns header");
            if (cr.offset > b.size()) throw PartError("header offset for column cannot exceed the columns header size");
            PReader r(b.substr(cr.offset)); out->unmarshal(r, partFormatLatestVersion); out->name = raw;
            return true;
        }
        return false;
    }
    std::string get_const_column_value(const BlockHeader& bh, sv name) const {   // :232-281
        std::string raw = raw_field_name(name);
        auto it = columnNameIDs.find(raw); if (it == columnNameIDs.end()) return "";
        ColumnsHeaderIndex idx = read_columns_header_index(bh);
        for (auto& cr : idx.constColumnsRefs) {
            if (cr.columnNameID != it->second) continue;
            sv b = read_at(file("columns_header.bin"), bh.columnsHeaderOffset, bh.columnsHeaderSize, "columns header");
            if (cr.offset > b.size()) throw PartError("header offset for const column cannot exceed the columns header size");
            PReader r(b.substr(cr.offset)); return std::string(r.bytes("field value"));
        }
        return "";
    }
    // the whole block as the oracle's Block (what blockSearch reads lazily: bloom :414-437, values :439-474, timestamps :479-506)
    Block read_block(const BlockHeader& bh) const {
        Block b; b.rows = bh.rowsCount;
        const TimestampsHeader& th = bh.timestampsHeader;
        if (th.blockSize > maxTimestampsBlockSize) throw PartError("timestamps block size is too big");
        b.hasTimestamps = true; b.ts.mt = th.marshalType; b.ts.first = th.minTimestamp; b.minTimestamp = th.minTimestamp; b.maxTimestamp = th.maxTimestamp;
        b.ts.data = std::string(read_at(file("timestamps.bin"), th.blockOffset, th.blockSize, "timestamps block"));
        ColumnsHeader csh = read_columns_header(bh);
        for (auto& ch : csh.columnHeaders) {
            Column c; c.name = canonical(ch.name); c.valueType = ch.valueType; c.minValue = ch.minValue; c.maxValue = ch.maxValue; c.dict = ch.dict;
            auto bv = bloom_values_for(ch.name);
            c.valuesBlock = std::string(read_at(*bv.second, ch.valuesOffset, ch.valuesSize, "values block"));
            if (ch.valueType != VT_DICT) c.bloom = std::string(read_at(*bv.first, ch.bloomFilterOffset, ch.bloomFilterSize, "bloom filter block"));
            b.columns.push_back(std::move(c));
        }
        for (auto& cc : csh.constColumns) b.consts.push_back({canonical(cc.name), cc.value});
        return b;
    }
};

}  // namespace vlo
