// ORACLE -- TEST INFRASTRUCTURE ONLY (see vlo_util.h header).
//
// CPU restatement of the deterministic synthetic-log generator.  Row SHAPE follows
// app/vlogsgenerator/main.go:240-281,335-349 (_msg template, toIPv4, toUUID, dictValues); the stock generator is
// unseeded and never emits the vocabulary the BASELINE.json configs query, so the deterministic variant below adds a
// hit/decoy vocabulary with a block-clustering knob and a row-selectivity knob, plus `level`, `path`, `status` fields.
// The GPU generator (victorialogs_b200/csrc/gen.cu) must produce byte-identical decoded column payloads and bloom
// filters to what THIS generator yields after passing through the reference writer path restated in vlo_block.h
// (valuesEncoder.encode -> marshalStringsBlock -> tokenizeHashes -> bloom); tests compare the two.
#pragma once
#include "vlo_block.h"

extern "C" {
struct vlo_gen_config {
    uint64_t seed;                 // 20250718 by convention
    uint64_t total_rows;           // rows in the data set; last block may be partial
    uint32_t rows_per_block;       // R (2 MB estimated-JSON rule evaluated by the caller for the field count)
    uint32_t hot_block_permille;   // block clustering: fraction of blocks that contain vocabulary rows
    uint32_t hit_row_permille;     // selectivity: in hot blocks, probability that a row draws a vocabulary template
    uint32_t columns_mask;         // bit0 _msg, bit1 level, bit2 path, bit3 status; bits 8..11: vocabulary focus (0 = uniform, k = always entry k - 1)
};
}

namespace vlo {

enum { GEN_COL_MSG = 0, GEN_COL_LEVEL = 1, GEN_COL_PATH = 2, GEN_COL_STATUS = 3, GEN_NCOLS = 4 };
static const char* const GEN_COL_NAMES[GEN_NCOLS] = {"_msg", "level", "path", "status"};
static const char* const GEN_VOCAB[12] = {"error", "timeout", "GET /api/v1/items", "conn 10.0.0.7 refused",
                                          "errors", "timeouts", "GETS /api/v2", "connection refuse",
                                          "conn reset by peer", "terror", "error timeout", "POST /api/v1/items"};
static const char* const GEN_LEVELS[8] = {"debug", "info", "warn", "error", "fatal", "ERROR", "FATAL", "INFO"};   // main.go:288-297
static const uint32_t GEN_STATUS[9] = {200, 201, 204, 301, 400, 404, 500, 502, 503};

inline uint64_t gen_rnd(uint64_t seed, uint64_t b, uint64_t i, uint64_t k) {
    uint64_t z = seed + (b + 1) * 0x9E3779B97F4A7C15ULL + (i + 1) * 0xD1B54A32D192ED03ULL + (k + 1) * 0x8CB92BA72F3D8DD7ULL;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31;
    return z;
}
inline uint64_t gen_nblocks(const vlo_gen_config& c) { return (c.total_rows + c.rows_per_block - 1) / c.rows_per_block; }
inline uint64_t gen_block_rows(const vlo_gen_config& c, uint64_t b) {
    uint64_t lo = b * c.rows_per_block, hi = std::min<uint64_t>(c.total_rows, lo + c.rows_per_block);
    return hi > lo ? hi - lo : 0;
}
inline bool gen_block_hot(const vlo_gen_config& c, uint64_t b) { return gen_rnd(c.seed, b, 0xFFFFFFFFULL, 0) % 1000 < c.hot_block_permille; }

inline std::string gen_value(const vlo_gen_config& c, uint64_t b, uint64_t i, int col) {
    std::string s;
    switch (col) {
    case GEN_COL_MSG: {
        uint64_t r0 = gen_rnd(c.seed, b, i, 0);
        const char* tmpl = "message";
        const uint32_t focus = (c.columns_mask >> 8) & 15;   // 1..12: every vocabulary row draws entry focus - 1; 0: uniform
        if (gen_block_hot(c, b) && (r0 % 1000) < c.hit_row_permille) tmpl = GEN_VOCAB[focus ? focus - 1 : (r0 >> 32) % 12];
        uint32_t ip = (uint32_t)gen_rnd(c.seed, b, i, 1);
        uint64_t ua = gen_rnd(c.seed, b, i, 2), ub = gen_rnd(c.seed, b, i, 3), u64 = gen_rnd(c.seed, b, i, 4);
        char buf[256];
        snprintf(buf, sizeof buf, "%s for the stream %llu and worker %llu; ip=%u.%u.%u.%u; uuid=%08llx-%04llx-%04llx-%04llx-%012llx; u64=%llu", tmpl,
                 (unsigned long long)b, (unsigned long long)(b % 7), ip >> 24, (ip >> 16) & 0xff, (ip >> 8) & 0xff, ip & 0xff,
                 (unsigned long long)(ua & 0xffffffffULL), (unsigned long long)((ua >> 32) & 0xffff), (unsigned long long)(ua >> 48),
                 (unsigned long long)(ub & 0xffff), (unsigned long long)(ub >> 16), (unsigned long long)u64);
        s = buf; break;
    }
    case GEN_COL_LEVEL: s = GEN_LEVELS[gen_rnd(c.seed, b, i, 5) % 8]; break;
    case GEN_COL_PATH: {
        uint64_t r = gen_rnd(c.seed, b, i, 6);
        char buf[64];
        switch (r % 4) {
        case 0: case 1: snprintf(buf, sizeof buf, "api/v1/items/%llu", (unsigned long long)((r >> 8) % 100000)); break;
        case 2: snprintf(buf, sizeof buf, "static/js/app.%llu.js", (unsigned long long)((r >> 8) % 1000)); break;
        default: snprintf(buf, sizeof buf, "health");
        }
        s = buf; break;
    }
    case GEN_COL_STATUS: marshal_uint64_string(s, GEN_STATUS[gen_rnd(c.seed, b, i, 7) % 9]); break;
    }
    return s;
}

inline Block gen_block(const vlo_gen_config& c, uint64_t b) {
    uint64_t rows = gen_block_rows(c, b);
    std::vector<std::string> names; std::vector<std::vector<std::string>> store; std::vector<std::vector<sv>> cols;
    for (int col = 0; col < GEN_NCOLS; col++) {
        if (!(c.columns_mask >> col & 1)) continue;
        names.push_back(GEN_COL_NAMES[col]);
        store.emplace_back();
        auto& v = store.back(); v.reserve(rows);
        for (uint64_t i = 0; i < rows; i++) v.push_back(gen_value(c, b, i, col));
    }
    for (auto& v : store) cols.emplace_back(v.begin(), v.end());
    return build_block(names, cols, rows);
}

}  // namespace vlo
