// Reader of one VictoriaLogs part directory: the host side of Storage.search below the partition level, ending in the vlscan_block
// descriptors (on-disk stage) that vlscan_scan_batch / vlscan_batch_upload take.  Plain C++ on untrusted files - no CUDA in here; the
// few ZSTD-compressed metadata files are inflated through a callback (the engine passes its device decoder).
//
//   part.mustOpenFilePart                       lib/logstorage/part.go:105-173        files of a part, format versions 1..3
//   partHeader.mustReadMetadata                 lib/logstorage/part_header.go:52-84   metadata.json
//   unmarshalColumnNames / unmarshalColumnIdxs  lib/logstorage/column_names.go:34-160
//   mustReadIndexBlockHeaders                   lib/logstorage/index_block_header.go:121-175   metaindex.bin -> indexBlockHeaders
//   indexBlockHeader.mustReadBlockHeaders       lib/logstorage/block_search.go:507-540, block_header.go:58-204   index.bin -> blockHeaders
//   blockSearch.getColumnHeader / getConstColumnValue   lib/logstorage/block_search.go:232-324   columnsHeaderIndex -> one columnHeader
//   columnHeader.unmarshalInplace               lib/logstorage/block_header.go:700-870
//   part.getBloomValuesFileForColumnName        lib/logstorage/part.go:194-217        which bloom / values file holds a column
// Format version 0 (column names inside the columns header, field_*.bin files) is not supported.
#pragma once
#include <errno.h>
#include <fcntl.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/vlscan.h"
#include "vl_hd.cuh"
#include "vl_zstd_walk.h"

namespace vl {
namespace part {

const uint64_t kFormatLatest = 3, kMaxTimestampsBlockSize = 8u << 20, kMaxRowsPerBlock = 8u << 20, kMaxColumnsPerBlock = 2000, kMaxIndexBlockSize = 8u << 20, kMaxValuesBlockSize = 8u << 20,
               kMaxBloomFilterBlockSize = 8u << 20, kMaxColumnsHeaderSize = 8u << 20, kMaxColumnsHeaderIndexSize = 8u << 20;   // lib/logstorage/consts.go:6-59

// inflates one ZSTD frame into exactly dst_len bytes (the Frame_Content_Size the caller read from the frame header); throws BadInput
using Inflate = std::function<void(const uint8_t* frame, size_t frame_len, uint8_t* dst, size_t dst_len)>;

struct Cursor {   // bounds-checked reads, big-endian fixed ints and LEB128 varuints (vm/lib/encoding/int.go)
    const uint8_t* p; size_t n; const char* what;
    Cursor(const uint8_t* p_, size_t n_, const char* what_) : p(p_), n(n_), what(what_) {}
    [[noreturn]] void fail(const char* field) const { throw BadInput(std::string("cannot unmarshal ") + field + " of " + what); }
    const uint8_t* take(size_t k, const char* field) { if (n < k) fail(field); const uint8_t* q = p; p += k; n -= k; return q; }
    uint8_t u8(const char* f) { return *take(1, f); }
    uint16_t be16(const char* f) { const uint8_t* q = take(2, f); return (uint16_t)((q[0] << 8) | q[1]); }
    uint32_t be32(const char* f) { const uint8_t* q = take(4, f); return ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3]; }
    uint64_t be64(const char* f) { uint64_t hi = be32(f); return (hi << 32) | be32(f); }
    uint64_t varuint(const char* f) {
        uint64_t v = 0;
        for (int i = 0, sh = 0; i < 10; i++, sh += 7) {
            if (n == 0) fail(f);
            const uint8_t b = *p++; n--;
            if (i == 9 && b > 1) fail(f);
            v |= (uint64_t)(b & 0x7F) << sh;
            if (b < 0x80) return v;
        }
        fail(f);
    }
    std::pair<const uint8_t*, size_t> bytes(const char* f) { const uint64_t l = varuint(f); if (l > n) fail(f); return {take((size_t)l, f), (size_t)l}; }
};

struct StreamID {
    uint32_t accountID = 0, projectID = 0; uint64_t hi = 0, lo = 0;
    void read(Cursor& c) { accountID = c.be32("accountID"); projectID = c.be32("projectID"); hi = c.be64("streamID"); lo = c.be64("streamID"); }
    bool less(const StreamID& a) const { if (accountID != a.accountID) return accountID < a.accountID; if (projectID != a.projectID) return projectID < a.projectID; if (hi != a.hi) return hi < a.hi; return lo < a.lo; }
    bool equal(const StreamID& a) const { return accountID == a.accountID && projectID == a.projectID && hi == a.hi && lo == a.lo; }
};
struct BlockHeader {   // block_header.go:14-165
    StreamID sid; uint64_t uncompressedSizeBytes = 0, rowsCount = 0;
    uint64_t tsOffset = 0, tsSize = 0; int64_t minTimestamp = 0, maxTimestamp = 0; uint8_t tsMarshalType = 0;   // timestampsHeader :919-1014
    uint64_t chIndexOffset = 0, chIndexSize = 0, chOffset = 0, chSize = 0;
    void read(Cursor& c) {
        sid.read(c);
        uncompressedSizeBytes = c.varuint("uncompressedSizeBytes");
        rowsCount = c.varuint("rowsCount");
        if (rowsCount > kMaxRowsPerBlock) throw BadInput("too big value for rowsCount in a blockHeader");
        tsOffset = c.be64("timestampsHeader"); tsSize = c.be64("timestampsHeader"); minTimestamp = (int64_t)c.be64("timestampsHeader"); maxTimestamp = (int64_t)c.be64("timestampsHeader");
        tsMarshalType = c.u8("timestampsHeader");
        chIndexOffset = c.varuint("columnsHeaderIndexOffset"); chIndexSize = c.varuint("columnsHeaderIndexSize");
        chOffset = c.varuint("columnsHeaderOffset"); chSize = c.varuint("columnsHeaderSize");
        if (chSize > kMaxColumnsHeaderSize) throw BadInput("too big value for columnsHeaderSize in a blockHeader");
    }
};
struct IndexBlockHeader { StreamID sid; int64_t minTimestamp = 0, maxTimestamp = 0; uint64_t offset = 0, size = 0; };   // index_block_header.go:13-104
struct PartHeader { uint64_t FormatVersion = 0, CompressedSizeBytes = 0, UncompressedSizeBytes = 0, RowsCount = 0, BlocksCount = 0; int64_t MinTimestamp = 0, MaxTimestamp = 0; uint64_t BloomValuesShardsCount = 0; };

struct ColumnHeader {   // block_header.go:545-917 (format v1+: the name comes from the columnsHeaderIndex)
    uint8_t valueType = 0; uint64_t minValue = 0, maxValue = 0;
    uint32_t dictLen = 0; std::pair<const uint8_t*, size_t> dict[8];
    uint64_t valuesOffset = 0, valuesSize = 0, bloomOffset = 0, bloomSize = 0;
    void read(Cursor& c) {
        valueType = c.u8("valueType");
        auto values = [&] { valuesOffset = c.varuint("valuesOffset"); valuesSize = c.varuint("valuesSize"); if (valuesSize > kMaxValuesBlockSize) throw BadInput("too big valuesSize in a columnHeader"); };
        auto values_bloom = [&] { values(); bloomOffset = c.varuint("bloomFilterOffset"); bloomSize = c.varuint("bloomFilterSize"); if (bloomSize > kMaxBloomFilterBlockSize) throw BadInput("too big bloomFilterSize in a columnHeader"); };
        switch (valueType) {
        case VLSCAN_VT_STRING: values_bloom(); break;
        case VLSCAN_VT_DICT: {
            dictLen = c.u8("dict len");
            if (dictLen > 8) throw BadInput("valuesDict may contain max 8 items");   // values_encoder.go:1289-1293 (the writer's limit; the scan kernels rely on it)
            for (uint32_t i = 0; i < dictLen; i++) dict[i] = c.bytes("dict value");
            values(); break;
        }
        case VLSCAN_VT_UINT8: minValue = c.u8("minValue"); maxValue = c.u8("maxValue"); values_bloom(); break;
        case VLSCAN_VT_UINT16: minValue = c.be16("minValue"); maxValue = c.be16("maxValue"); values_bloom(); break;
        case VLSCAN_VT_UINT32: case VLSCAN_VT_IPV4: minValue = c.be32("minValue"); maxValue = c.be32("maxValue"); values_bloom(); break;
        case VLSCAN_VT_UINT64: case VLSCAN_VT_FLOAT64: case VLSCAN_VT_ISO8601: minValue = c.be64("minValue"); maxValue = c.be64("maxValue"); values_bloom(); break;
        case VLSCAN_VT_INT64: {   // encoding.MarshalInt64: zig-zag, then 8 bytes big endian
            auto unzz = [](uint64_t u) { return (uint64_t)((int64_t)(u >> 1) ^ -(int64_t)(u & 1)); };
            minValue = unzz(c.be64("minValue")); maxValue = unzz(c.be64("maxValue")); values_bloom(); break;
        }
        default: throw BadInput("unexpected valueType=" + std::to_string(valueType) + " in a columnHeader");
        }
    }
};

class MappedFile {
public:
    const uint8_t* p = nullptr; size_t n = 0;
    MappedFile() = default;
    MappedFile(const MappedFile&) = delete; MappedFile& operator=(const MappedFile&) = delete;
#ifdef VL_PART_HEAP_FILES   // sanitizer builds (tests/host_asan): exact-size heap copies, so that a read past the end of a file is seen
    ~MappedFile() { delete[] p; }
#else
    ~MappedFile() { if (p) munmap((void*)p, n); }
#endif
    void open(const std::string& path) {
        const int fd = ::open(path.c_str(), O_RDONLY | O_CLOEXEC);
        if (fd < 0) throw BadInput("cannot open " + path + ": " + strerror(errno));
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { ::close(fd); throw BadInput("cannot stat " + path); }
        n = (size_t)st.st_size;
#ifdef VL_PART_HEAP_FILES
        uint8_t* buf = new uint8_t[n ? n : 1];
        size_t got = 0;
        while (got < n) { const ssize_t r = ::read(fd, buf + got, n - got); if (r <= 0) break; got += (size_t)r; }
        ::close(fd);
        if (got != n) { delete[] buf; n = 0; throw BadInput("cannot read " + path); }
        p = buf;
#else
        if (n) { void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0); if (m == MAP_FAILED) { ::close(fd); n = 0; throw BadInput("cannot map " + path + ": " + strerror(errno)); } p = (const uint8_t*)m; }
        ::close(fd);
#endif
    }
    // fs.MustReadAt: the range has to lie inside the file
    const uint8_t* at(uint64_t off, uint64_t size, const char* what) const { if (off > n || size > n - off) throw BadInput(std::string("cannot read ") + what + ": the range lies outside the file"); return p + off; }
};

// what vlscan_part_blocks hands out: descriptors over the mapped files + the storage of what had to be rebuilt (dict tables)
struct Described {
    std::vector<vlscan_block> blocks; std::vector<vlscan_column> cols; std::vector<std::string> fields;
    std::vector<std::unique_ptr<std::vector<uint8_t>>> owned;
    std::vector<uint64_t> source;   // index of each described block inside the part
};

class PartReader {
public:
    PartHeader ph; std::string path;
    std::vector<std::string> columnNames; std::unordered_map<std::string, uint64_t> columnNameIDs; std::vector<uint64_t> columnShard;   // by columnNameID; UINT64_MAX = not listed
    std::vector<IndexBlockHeader> indexBlockHeaders; std::vector<BlockHeader> blockHeaders;

    void open(const std::string& dir, const Inflate& inflate) {
        path = dir;
        read_metadata(read_small(dir + "/metadata.json"));
        if (ph.FormatVersion < 1) throw BadInput(dir + ": part format version 0 is not supported");
        { MappedFile f; f.open(dir + "/column_names.bin"); parse_column_names(inflate_frame(f.p, f.n, inflate, "column names")); }
        if (ph.FormatVersion >= 3) { MappedFile f; f.open(dir + "/column_idxs.bin"); parse_column_idxs(f.p, f.n); }
        { MappedFile f; f.open(dir + "/metaindex.bin"); parse_metaindex(inflate_frame(f.p, f.n, inflate, "indexBlockHeader entries")); }
        index_.open(dir + "/index.bin"); chIndex_.open(dir + "/columns_header_index.bin"); ch_.open(dir + "/columns_header.bin"); timestamps_.open(dir + "/timestamps.bin");
        msgBloom_.open(dir + "/message_bloom.bin"); msgValues_.open(dir + "/message_values.bin");
        bloom_.clear(); values_.clear();
        for (uint64_t i = 0; i < ph.BloomValuesShardsCount; i++) {
            bloom_.emplace_back(new MappedFile); bloom_.back()->open(dir + "/bloom.bin" + std::to_string(i));
            values_.emplace_back(new MappedFile); values_.back()->open(dir + "/values.bin" + std::to_string(i));
        }
        // every index block up front: a part holds one 56-byte record per ~128 KB of block headers
        uint64_t rows = 0;
        for (const IndexBlockHeader& ih : indexBlockHeaders) {
            if (ih.size > kMaxIndexBlockSize) throw BadInput("indexBlockHeader.indexBlockSize is too big");
            const std::vector<uint8_t> raw = inflate_frame(index_.at(ih.offset, ih.size, "an index block"), (size_t)ih.size, inflate, "an index block");
            Cursor c(raw.data(), raw.size(), "a blockHeader");
            const size_t first = blockHeaders.size();
            while (c.n) { blockHeaders.emplace_back(); blockHeaders.back().read(c); rows += blockHeaders.back().rowsCount; }
            for (size_t i = first + 1; i < blockHeaders.size(); i++) {   // validateBlockHeaders block_header.go:186-204
                const BlockHeader &cur = blockHeaders[i], &prev = blockHeaders[i - 1];
                if (cur.sid.less(prev.sid)) throw BadInput("unexpected blockHeader with smaller streamID after bigger streamID");
                if (cur.sid.equal(prev.sid) && cur.minTimestamp < prev.minTimestamp) throw BadInput("unexpected blockHeader with smaller timestamp after bigger timestamp");
            }
        }
        if (blockHeaders.size() != ph.BlocksCount) throw BadInput(dir + ": the index holds " + std::to_string(blockHeaders.size()) + " block headers, metadata.json says BlocksCount=" + std::to_string(ph.BlocksCount));
        if (rows != ph.RowsCount) throw BadInput(dir + ": the block headers hold " + std::to_string(rows) + " rows, metadata.json says RowsCount=" + std::to_string(ph.RowsCount));
    }

    // Descriptors of the blocks [lo, hi) whose time range overlaps [minTs, maxTs], restricted to `fields` (canonical names: "_msg" is the message).
    void describe(const std::vector<std::string>& fields, uint64_t lo, uint64_t hi, int64_t minTs, int64_t maxTs, Described& out) const {
        if (lo > hi || hi > blockHeaders.size()) throw BadInput("block range outside the part");
        out.fields = fields;
        std::vector<uint64_t> ids(fields.size(), UINT64_MAX);
        for (size_t f = 0; f < fields.size(); f++) { auto it = columnNameIDs.find(fields[f] == "_msg" ? std::string() : fields[f]); if (it != columnNameIDs.end()) ids[f] = it->second; }
        std::vector<size_t> first_col;
        for (uint64_t b = lo; b < hi; b++) {
            const BlockHeader& bh = blockHeaders[b];
            if (bh.maxTimestamp < minTs || bh.minTimestamp > maxTs) continue;   // the part search skips such blocks before blockSearch (block_search.go:60-63 / partition search)
            if (bh.chIndexSize > kMaxColumnsHeaderIndexSize) throw BadInput("columns header index size is too big");
            Cursor ix(chIndex_.at(bh.chIndexOffset, bh.chIndexSize, "a columns header index"), (size_t)bh.chIndexSize, "a columnsHeaderIndex");
            const uint8_t* chp = ch_.at(bh.chOffset, bh.chSize, "a columns header"); const size_t chn = (size_t)bh.chSize;
            first_col.push_back(out.cols.size());
            // columnsHeaderIndex.unmarshalInplace block_header.go:262-333: refs of the columns, then refs of the const columns
            for (int pass = 0; pass < 2; pass++) {
                const uint64_t cnt = ix.varuint("the number of columnHeaderRef items");
                if (cnt > ix.n) throw BadInput("too many columnHeaderRef items");
                if (cnt > kMaxColumnsPerBlock) throw BadInput("too many columns in a block");
                for (uint64_t i = 0; i < cnt; i++) {
                    const uint64_t id = ix.varuint("columnNameID"), off = ix.varuint("column header offset");
                    if (id >= columnNames.size()) throw BadInput("unexpected columnNameID in a columnsHeaderIndex");
                    for (size_t f = 0; f < fields.size(); f++) {
                        if (ids[f] != id) continue;
                        if (off > chn) throw BadInput("header offset for a column cannot exceed the columns header size");
                        for (size_t k = first_col.back(); k < out.cols.size(); k++) if (out.cols[k].field == f) throw BadInput("a block lists one column twice");
                        Cursor c(chp + off, chn - off, pass == 0 ? "a columnHeader" : "a const column");
                        vlscan_column col; memset(&col, 0, sizeof col); col.field = (uint32_t)f;
                        if (pass == 1) { auto v = c.bytes("field value"); col.kind = VLSCAN_COL_CONST; col.const_value = v.first; col.const_len = v.second; }
                        else fill_values_column(col, c, id, out);
                        out.cols.push_back(col);
                    }
                }
            }
            if (ix.n) throw BadInput("unexpected non-empty tail left after unmarshaling columnsHeaderIndex");
            vlscan_block blk; memset(&blk, 0, sizeof blk); blk.rows = bh.rowsCount; blk.ncols = (uint32_t)(out.cols.size() - first_col.back());
            // the timestamps column travels with the block: `_time` filters that only partly cover it and vlscan_gather_timestamps need it
            if (bh.tsSize > kMaxTimestampsBlockSize) throw BadInput("timestamps block size is too big");   // getTimestamps block_search.go:490-493
            blk.ts_marshal_type = bh.tsMarshalType; blk.timestamps = timestamps_.at(bh.tsOffset, bh.tsSize, "a timestamps block"); blk.timestamps_len = bh.tsSize;
            blk.min_timestamp = bh.minTimestamp; blk.max_timestamp = bh.maxTimestamp;
            out.blocks.push_back(blk); out.source.push_back(b);
        }
        for (size_t i = 0; i < out.blocks.size(); i++) out.blocks[i].cols = out.cols.data() + first_col[i];
    }

    const MappedFile& timestamps_file() const { return timestamps_; }

private:
    MappedFile index_, chIndex_, ch_, timestamps_, msgBloom_, msgValues_;
    std::vector<std::unique_ptr<MappedFile>> bloom_, values_;

    static std::string read_small(const std::string& path) { MappedFile f; f.open(path); if (f.n > (1u << 20)) throw BadInput(path + " is too big"); return std::string((const char*)f.p, f.n); }

    static std::vector<uint8_t> inflate_frame(const uint8_t* p, size_t n, const Inflate& inflate, const char* what) {
        std::vector<zs::ZBlock> blocks; zs::ZFrame fr{};
        try { zwalk::parse_frame_into(blocks, 0, p, n, 0, fr); } catch (const BadInput& e) { throw BadInput(std::string("cannot decompress ") + what + ": " + e.msg); }
        if (fr.fcs > (256u << 20)) throw BadInput(std::string("cannot decompress ") + what + ": it claims to regenerate more than 256 MB");   // metadata: 56 B per index block, <= 128 KB per index block
        std::vector<uint8_t> out((size_t)fr.fcs);
        inflate(p, n, out.data(), out.size());
        return out;
    }

    void fill_values_column(vlscan_column& col, Cursor& c, uint64_t nameID, Described& out) const {
        ColumnHeader h; h.read(c);
        col.kind = VLSCAN_COL_VALUES; col.stage = VLSCAN_STAGE_ONDISK; col.value_type = h.valueType; col.min_value = h.minValue; col.max_value = h.maxValue;
        const MappedFile *vf, *bf;
        files_for(nameID, &bf, &vf);
        col.values = vf->at(h.valuesOffset, h.valuesSize, "a values block"); col.values_len = h.valuesSize;
        if (h.valueType == VLSCAN_VT_DICT) {
            // device layout of a dict: u32 offsets[dict_len+1] followed by the bytes
            uint32_t total = 0; for (uint32_t i = 0; i < h.dictLen; i++) total += (uint32_t)h.dict[i].second;
            auto buf = std::make_unique<std::vector<uint8_t>>(4 * (h.dictLen + 1) + total);
            uint32_t* offs = (uint32_t*)buf->data(); uint8_t* blob = buf->data() + 4 * (h.dictLen + 1);
            uint32_t o = 0;
            for (uint32_t i = 0; i < h.dictLen; i++) { offs[i] = o; if (h.dict[i].second) memcpy(blob + o, h.dict[i].first, h.dict[i].second); o += (uint32_t)h.dict[i].second; }
            offs[h.dictLen] = o;
            col.dict_len = h.dictLen; col.dict_offsets = offs; col.dict_blob = blob;
            out.owned.push_back(std::move(buf));
            col.bloom = nullptr; col.bloom_len = 0;   // no bloom filter is stored for dict columns (block.go:158-166)
        } else {
            col.bloom = bf->at(h.bloomOffset, h.bloomSize, "a bloom filter block"); col.bloom_len = h.bloomSize;
        }
    }

    void files_for(uint64_t nameID, const MappedFile** bloom, const MappedFile** values) const {   // part.go:194-217
        const std::string& name = columnNames[nameID];
        if (name.empty()) { *bloom = &msgBloom_; *values = &msgValues_; return; }
        uint64_t shard = 0;
        if (ph.FormatVersion < 3) { const uint64_t n = ph.BloomValuesShardsCount; if (n > 1) shard = xxh64((const uint8_t*)name.data(), (uint32_t)name.size()) % n; }
        else { shard = columnShard[nameID]; if (shard == UINT64_MAX) throw BadInput("BUG: unknown shard index for column " + name); }
        if (shard >= bloom_.size()) throw BadInput("shard index outside BloomValuesShardsCount for column " + name);
        *bloom = bloom_[shard].get(); *values = values_[shard].get();
    }

    // a flat JSON object with integer members (encoding/json output of partHeader); unknown scalar members are skipped
    void read_metadata(const std::string& s) {
        size_t i = 0;
        auto ws = [&] { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) i++; };
        auto bad = [&]() -> BadInput { return BadInput(path + "/metadata.json: cannot parse"); };
        auto expect = [&](char ch) { ws(); if (i >= s.size() || s[i] != ch) throw bad(); i++; };
        auto str = [&] { expect('"'); std::string k; while (i < s.size() && s[i] != '"') { if (s[i] == '\\') i++; if (i < s.size()) k.push_back(s[i++]); } expect('"'); return k; };
        expect('{'); ws();
        if (i < s.size() && s[i] == '}') i++;
        else for (;;) {
            const std::string key = str(); expect(':'); ws();
            if (i < s.size() && s[i] == '"') str();
            else {
                const size_t st = i;
                while (i < s.size() && (s[i] == '-' || s[i] == '+' || s[i] == '.' || (s[i] >= '0' && s[i] <= '9') || (s[i] >= 'a' && s[i] <= 'z') || s[i] == 'E')) i++;
                if (i == st) throw bad();
                const bool neg = s[st] == '-'; uint64_t u = 0; bool isint = true;
                for (size_t k = st + (neg ? 1 : 0); k < i; k++) { if (s[k] < '0' || s[k] > '9' || u > (UINT64_MAX - 9) / 10) { isint = false; break; } u = u * 10 + (uint64_t)(s[k] - '0'); }
                if (i - st == (neg ? 1u : 0u)) isint = false;
                auto want_u = [&](uint64_t& dst) { if (!isint || neg) throw BadInput(path + "/metadata.json: " + key + " must be an unsigned integer"); dst = u; };
                auto want_i = [&](int64_t& dst) { if (!isint || u > (uint64_t)INT64_MAX) throw BadInput(path + "/metadata.json: " + key + " must be an integer"); dst = neg ? -(int64_t)u : (int64_t)u; };
                if (key == "FormatVersion") want_u(ph.FormatVersion); else if (key == "CompressedSizeBytes") want_u(ph.CompressedSizeBytes);
                else if (key == "UncompressedSizeBytes") want_u(ph.UncompressedSizeBytes); else if (key == "RowsCount") want_u(ph.RowsCount);
                else if (key == "BlocksCount") want_u(ph.BlocksCount); else if (key == "MinTimestamp") want_i(ph.MinTimestamp);
                else if (key == "MaxTimestamp") want_i(ph.MaxTimestamp); else if (key == "BloomValuesShardsCount") want_u(ph.BloomValuesShardsCount);
            }
            ws();
            if (i < s.size() && s[i] == ',') { i++; continue; }
            expect('}'); break;
        }
        ws(); if (i != s.size()) throw bad();
        // partHeader.mustReadMetadata part_header.go:62-83
        if (ph.FormatVersion <= 1) { if (ph.BloomValuesShardsCount != 0) throw BadInput(path + ": unexpected BloomValuesShardsCount for FormatVersion<=1"); if (ph.FormatVersion == 1) ph.BloomValuesShardsCount = 8; }
        if (ph.FormatVersion > kFormatLatest) throw BadInput(path + ": unsupported part format version " + std::to_string(ph.FormatVersion));
        if (ph.MinTimestamp > ph.MaxTimestamp) throw BadInput(path + ": MinTimestamp cannot exceed MaxTimestamp");
        if (ph.BlocksCount > ph.RowsCount) throw BadInput(path + ": BlocksCount cannot exceed RowsCount");
        if (ph.BloomValuesShardsCount > 4096) throw BadInput(path + ": too many bloom / values shards");
    }
    void parse_column_names(const std::vector<uint8_t>& raw) {   // column_names.go:113-160
        Cursor c(raw.data(), raw.size(), "column names");
        const uint64_t cnt = c.varuint("the number of column names");
        if (cnt > c.n) throw BadInput("too many distinct column names");
        for (uint64_t id = 0; id < cnt; id++) {
            auto b = c.bytes("column name");
            std::string name((const char*)b.first, b.second);
            if (!columnNameIDs.emplace(name, id).second) throw BadInput("duplicate ids for column name " + name);
            columnNames.push_back(std::move(name));
        }
        if (c.n) throw BadInput("unexpected non-empty tail left after unmarshaling column name ids");
        columnShard.assign(columnNames.size(), UINT64_MAX);
    }
    void parse_column_idxs(const uint8_t* p, size_t n) {   // column_names.go:42-83
        Cursor c(p, n, "column indexes");
        const uint64_t cnt = c.varuint("the number of entries");
        if (cnt > c.n) throw BadInput("too many column index entries");
        for (uint64_t i = 0; i < cnt; i++) {
            const uint64_t id = c.varuint("columnID"), shard = c.varuint("shardIdx");
            if (shard >= ph.BloomValuesShardsCount) throw BadInput("too big shardIdx in column indexes");
            if (id >= columnNames.size()) throw BadInput("too big columnID in column indexes");
            columnShard[id] = shard;
        }
        if (c.n) throw BadInput("unexpected tail left after reading column indexes");
    }
    void parse_metaindex(const std::vector<uint8_t>& raw) {   // index_block_header.go:143-175
        Cursor c(raw.data(), raw.size(), "an indexBlockHeader");
        while (c.n) {
            IndexBlockHeader ih; ih.sid.read(c);
            ih.minTimestamp = (int64_t)c.be64("minTimestamp"); ih.maxTimestamp = (int64_t)c.be64("maxTimestamp"); ih.offset = c.be64("indexBlockOffset"); ih.size = c.be64("indexBlockSize");
            if (!indexBlockHeaders.empty() && ih.sid.less(indexBlockHeaders.back().sid)) throw BadInput("unexpected indexBlockHeader with smaller streamID after bigger streamID");
            indexBlockHeaders.push_back(ih);
        }
    }
};

}  // namespace part
}  // namespace vl
