// Host-side walk of bytes-block containers and ZSTD frames: the only parts of a values block the host ever reads (container type and
// length, frame header, block headers, literals / sequences section headers).  Plain C++ on untrusted bytes - tests build it with
// AddressSanitizer (tests/host_asan/harness.cpp) - and free of shared state, so several threads can walk into vectors of their own.
//   unmarshalBytesBlock            lib/logstorage/encoding.go:372-426 (container), RFC 8878 3.1.1 (frames)
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <algorithm>
#include <string>
#include <vector>
#include "vl_zstd_types.h"

namespace vl {

struct BadInput { std::string msg; explicit BadInput(std::string m) : msg(std::move(m)) {} };

namespace zwalk {

using namespace zs;

const uint32_t kNone = 0xFFFFFFFEu;                 // "no table seen yet in this frame"
const uint64_t kMaxFrameContent = 1ull << 30;       // a values block never regenerates more than this (consts.go: blocks are <= 2 MB uncompressed)
const uint32_t kBlockMax = 128u << 10;              // Block_Maximum_Size upper bound

inline uint32_t le16(const uint8_t* p) { return p[0] | (p[1] << 8); }
inline uint32_t le24(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16); }
inline uint32_t le32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint64_t le40(const uint8_t* p) { return (uint64_t)le32(p) | ((uint64_t)p[4] << 32); }
inline uint64_t le64(const uint8_t* p) { return (uint64_t)le32(p) | ((uint64_t)le32(p + 4) << 32); }

// Literals_Section_Header and Sequences_Section_Header of one compressed block; table slots are numbered from 0 inside the frame
inline void parse_compressed_block(const uint8_t* b, uint32_t bsize, ZBlock& B, uint32_t& frame_huf, uint32_t& frame_fse, uint32_t& prev_huf, uint32_t prev_fse[3]) {
    if (bsize < 2) throw BadInput("cannot decompress block: compressed ZSTD block is too short");
    uint32_t lt = b[0] & 3, sf = (b[0] >> 2) & 3, hl, regen, comp, streams = 0;
    if (lt < ZL_COMPRESSED) {
        if (sf == 0 || sf == 2) { hl = 1; regen = b[0] >> 3; }
        else if (sf == 1) { hl = 2; regen = le16(b) >> 4; }
        else { hl = 3; if (bsize < 3) throw BadInput("cannot decompress block: truncated literals header"); regen = le24(b) >> 4; }
        comp = lt == ZL_RAW ? regen : 1;
    } else {
        if (bsize < 5) throw BadInput("cannot decompress block: truncated literals header");
        if (sf < 2) { hl = 3; uint32_t v = le24(b); regen = (v >> 4) & 0x3FF; comp = v >> 14; streams = sf == 0 ? 1 : 4; }
        else if (sf == 2) { hl = 4; uint32_t v = le32(b); regen = (v >> 4) & 0x3FFF; comp = v >> 18; streams = 4; }
        else { hl = 5; uint64_t v = le40(b); regen = (uint32_t)((v >> 4) & 0x3FFFF); comp = (uint32_t)(v >> 22); streams = 4; }
    }
    if (regen > kBlockMax) throw BadInput("cannot decompress block: literals section exceeds the maximum block size");
    if ((uint64_t)hl + comp >= bsize) throw BadInput("cannot decompress block: literals section exceeds the block");
    uint32_t q = hl + comp, nseq = b[q];
    if (nseq < 128) q += 1;
    else if (nseq < 255) { if (q + 2 > bsize) throw BadInput("cannot decompress block: truncated sequences header"); nseq = ((nseq - 128) << 8) + b[q + 1]; q += 2; }
    else { if (q + 3 > bsize) throw BadInput("cannot decompress block: truncated sequences header"); nseq = b[q + 1] + (b[q + 2] << 8) + 0x7F00; q += 3; }
    B.lit_type = (uint8_t)lt; B.lit_streams = (uint8_t)streams; B.lit_hdr = hl; B.lit_regen = regen; B.lit_comp = comp; B.nseq = nseq; B.seq_hdr = q;
    B.huf_own = Z_PREDEF; B.huf_slot = Z_PREDEF; B.fse_own = Z_PREDEF; B.ll_slot = B.of_slot = B.ml_slot = Z_PREDEF; B.modes = 0;
    if (lt == ZL_COMPRESSED) { B.huf_own = B.huf_slot = frame_huf++; prev_huf = B.huf_own; }
    else if (lt == ZL_TREELESS) { if (prev_huf == kNone) throw BadInput("cannot decompress block: treeless literals without a previous Huffman table"); B.huf_slot = prev_huf; }
    if (nseq == 0) { if (q != bsize) throw BadInput("cannot decompress block: bytes after an empty sequences section"); return; }
    if (q >= bsize) throw BadInput("cannot decompress block: truncated sequences header");
    B.modes = b[q];
    if (B.modes & 3) throw BadInput("cannot decompress block: reserved bits set in the symbol compression modes");
    uint32_t* slot[3] = {&B.ll_slot, &B.of_slot, &B.ml_slot};
    for (int k = 0; k < 3; k++) {
        int mode = (B.modes >> (6 - 2 * k)) & 3;
        if (mode == 0) *slot[k] = Z_PREDEF;
        else if (mode == 3) { if (prev_fse[k] == kNone) throw BadInput("cannot decompress block: repeat mode without a previous table"); *slot[k] = prev_fse[k]; }
        else { if (B.fse_own == Z_PREDEF) B.fse_own = frame_fse++; *slot[k] = B.fse_own; }
        prev_fse[k] = *slot[k];
    }
}

// Walks one frame occupying exactly [f, f+n) and appends its blocks to `blocks`; fid = the index the frame has in the job, zoff = the
// offset its first byte will have in the compressed staging buffer.  out gets fcs / blk_lo / blk_hi (indices into `blocks`).
inline void parse_frame_into(std::vector<ZBlock>& blocks, uint32_t fid, const uint8_t* f, size_t n, uint64_t zoff, ZFrame& out) {
    if (n < 6 || le32(f) != 0xFD2FB528u) throw BadInput("cannot decompress block: not a ZSTD frame");
    uint32_t fhd = f[4], fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, cks = (fhd >> 2) & 1, did_flag = fhd & 3;
    if (fhd & 8) throw BadInput("cannot decompress block: reserved bit set in the frame header");
    size_t pos = 5;
    uint64_t window = 0;
    if (!single) {
        if (pos >= n) throw BadInput("cannot decompress block: truncated frame header");
        uint32_t wd = f[pos++]; uint32_t wlog = 10 + (wd >> 3);
        if (wlog > 31) throw BadInput("cannot decompress block: window too large");
        window = (1ull << wlog) + ((1ull << wlog) >> 3) * (wd & 7);
    }
    static const int did_sz[4] = {0, 1, 2, 4};
    if (pos + did_sz[did_flag] > n) throw BadInput("cannot decompress block: truncated frame header");
    uint32_t did = 0; for (int i = 0; i < did_sz[did_flag]; i++) did |= (uint32_t)f[pos + i] << (8 * i);
    pos += did_sz[did_flag];
    if (did) throw BadInput("cannot decompress block: dictionaries are not supported");
    int fcs_sz = fcs_flag == 0 ? (single ? 1 : 0) : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
    if (!fcs_sz) throw BadInput("cannot decompress block: the frame does not declare its content size");
    if (pos + fcs_sz > n) throw BadInput("cannot decompress block: truncated frame header");
    uint64_t fcs = fcs_sz == 1 ? f[pos] : fcs_sz == 2 ? le16(f + pos) + 256u : fcs_sz == 4 ? le32(f + pos) : le64(f + pos);
    pos += fcs_sz;
    if (fcs > kMaxFrameContent) throw BadInput("cannot decompress block: frame content size is too large");
    if (single) window = fcs;
    const uint32_t block_max = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(window, 1), kBlockMax);
    ZFrame fr{}; fr.fcs = fcs; fr.blk_lo = (uint32_t)blocks.size(); fr.dst = 0;
    uint32_t frame_huf = 0, frame_fse = 0, prev_huf = kNone, prev_fse[3] = {kNone, kNone, kNone};
    bool seen_sequences = false;
    for (;;) {
        if (pos + 3 > n) throw BadInput("cannot decompress block: truncated block header");
        uint32_t h = le24(f + pos); pos += 3;
        uint32_t last = h & 1, type = (h >> 1) & 3, bsize = h >> 3;
        if (type == 3) throw BadInput("cannot decompress block: reserved block type");
        if (bsize > block_max) throw BadInput("cannot decompress block: block exceeds the maximum block size");
        size_t content = type == ZB_RLE ? 1 : bsize;
        if (pos + content > n) throw BadInput("cannot decompress block: truncated block");
        ZBlock B{}; B.src = zoff + pos; B.size = bsize; B.frame = fid; B.type = (uint8_t)type;
        B.huf_own = B.huf_slot = B.fse_own = B.ll_slot = B.of_slot = B.ml_slot = Z_PREDEF;
        if (type == ZB_COMPRESSED) {
            parse_compressed_block(f + pos, bsize, B, frame_huf, frame_fse, prev_huf, prev_fse);
            B.rep_known = seen_sequences ? 0 : 1;
            if (B.nseq) seen_sequences = true;
        }
        blocks.push_back(B);
        pos += content;
        if (last) break;
    }
    if (cks) { if (pos + 4 > n) throw BadInput("cannot decompress block: truncated content checksum"); pos += 4; }   // not verified
    if (pos != n) throw BadInput("cannot decompress block: unexpected bytes after the ZSTD frame");
    fr.blk_hi = (uint32_t)blocks.size();
    out = fr;
}

// unmarshalBytesBlock (lib/logstorage/encoding.go:372-426) without the decompression: one container starting at p (n bytes available);
// returns the bytes consumed
inline size_t parse_bytes_block_into(std::vector<ZBlock>& blocks, uint32_t fid, const uint8_t* p, size_t n, uint64_t zoff, ZFrame& out) {
    if (n < 1) throw BadInput("cannot unmarshal block type from empty src");
    if (p[0] == 0) {           // marshalBytesTypePlain
        if (n < 2) throw BadInput("cannot unmarshal plain block size from empty src");
        size_t len = p[1];
        if (n - 2 < len) throw BadInput("cannot read plain block: not enough bytes");
        ZFrame fr{}; fr.fcs = len; fr.blk_lo = (uint32_t)blocks.size();
        ZBlock B{}; B.src = zoff + 2; B.size = (uint32_t)len; B.frame = fid; B.type = ZB_RAW;
        B.huf_own = B.huf_slot = B.fse_own = B.ll_slot = B.of_slot = B.ml_slot = Z_PREDEF;
        blocks.push_back(B);
        fr.blk_hi = (uint32_t)blocks.size();
        out = fr;
        return 2 + len;
    }
    if (p[0] == 1) {           // marshalBytesTypeZSTD
        uint64_t clen = 0; int sh = 0; size_t i = 1; bool done = false;
        for (; i < n && i < 11; i++) { clen |= (uint64_t)(p[i] & 0x7F) << sh; sh += 7; if (p[i] < 0x80) { done = true; i++; break; } }
        if (!done) throw BadInput("cannot unmarshal compressed block size");
        if (n - i < clen) throw BadInput("cannot read compressed block: not enough bytes");
        parse_frame_into(blocks, fid, p + i, clen, zoff + i, out);
        return i + clen;
    }
    throw BadInput("unexpected block type; supported types: 0, 1");
}

}  // namespace zwalk
}  // namespace vl
