// ORACLE -- TEST INFRASTRUCTURE ONLY (see vlo_util.h header).  C API consumed through ctypes by tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs.
#include "vlo_block.h"
#include "vlo_gen.h"
#include "vlo_part.h"
#include <pthread.h>
#include <sched.h>
#include <thread>
#include <atomic>
#include <chrono>

using namespace vlo;

namespace {
thread_local std::string g_err;
struct FilterHandle { FP f; };
struct BlockHandle { Block b; };

// Packed string list: blob + (n+1) u64 offsets
std::vector<sv> unpack(const uint8_t* blob, const uint64_t* offs, uint64_t n) {
    std::vector<sv> v(n);
    for (uint64_t i = 0; i < n; i++) v[i] = sv((const char*)blob + offs[i], offs[i + 1] - offs[i]);
    return v;
}
template <class F> int guard(F&& f) {
    try { f(); return 0; } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
}  // namespace

extern "C" {

const char* vlo_last_error() { return g_err.c_str(); }

uint64_t vlo_xxh64(const void* p, uint64_t n) { return xxh64(p, n); }

// tokenizeStrings: out blob gets tokens joined by '\n' (tokens never contain '\n'); returns length or -1
int64_t vlo_tokenize_strings(const uint8_t* blob, const uint64_t* offs, uint64_t n, char* out, uint64_t cap) {
    auto toks = tokenize_strings(unpack(blob, offs, n));
    std::string s;
    for (size_t i = 0; i < toks.size(); i++) { if (i) s.push_back('\n'); s += toks[i]; }
    if (s.size() > cap) return -1;
    memcpy(out, s.data(), s.size());
    return (int64_t)s.size();
}
int64_t vlo_tokenize_hashes(const uint8_t* blob, const uint64_t* offs, uint64_t n, uint64_t* out, uint64_t cap) {
    auto h = tokenize_hashes(unpack(blob, offs, n));
    if (h.size() > cap) return -1;
    memcpy(out, h.data(), h.size() * 8);
    return (int64_t)h.size();
}
// appendTokensHashes for one token: 6 hashes
void vlo_token_hashes(const void* tok, uint64_t n, uint64_t* out6) {
    std::vector<uint64_t> v; append_hashes_hashes(v, xxh64(tok, n)); memcpy(out6, v.data(), 48);
}
// bloomFilterMarshalTokens bloomfilter.go:22-28 (tokens are NOT deduplicated here, like the reference)
int64_t vlo_bloom_marshal_tokens(const uint8_t* blob, const uint64_t* offs, uint64_t n, uint8_t* out, uint64_t cap) {
    auto toks = unpack(blob, offs, n);
    std::vector<uint64_t> hashes;
    for (auto t : toks) hashes.push_back(xxh64(t));
    BloomFilter bf; bf.init_hashes(hashes);
    std::string m = bf.marshal();
    if (m.size() > cap) return -1;
    memcpy(out, m.data(), m.size());
    return (int64_t)m.size();
}
int vlo_bloom_contains_all_tokens(const uint8_t* bloom, uint64_t bloom_len, const uint8_t* blob, const uint64_t* offs, uint64_t n) {
    BloomFilter bf; if (!bf.unmarshal(sv((const char*)bloom, bloom_len))) return -1;
    std::vector<std::string> toks; for (auto t : unpack(blob, offs, n)) toks.emplace_back(t);
    return bf.contains_all(tokens_hashes(toks)) ? 1 : 0;
}

int vlo_match_phrase(const void* s, uint64_t sl, const void* p, uint64_t pl) { return match_phrase(sv((const char*)s, sl), sv((const char*)p, pl)); }
int vlo_match_prefix(const void* s, uint64_t sl, const void* p, uint64_t pl) { return match_prefix(sv((const char*)s, sl), sv((const char*)p, pl)); }
// single-value predicates of exact_prefix (9) / len_range (10) / string_range (11) / ipv4_range (12); kind numbers as in include/vlscan.h
int vlo_eval_predicate(int kind, const void* s, uint64_t sl, const void* a, uint64_t al, const void* b, uint64_t bl, uint64_t aux0, uint64_t aux1) {
    sv v((const char*)s, sl), x((const char*)a, al), y((const char*)b, bl);
    switch (kind) {
    case 9: return match_exact_prefix(v, x);
    case 10: return match_len_range(v, aux0, aux1);
    case 11: return match_string_range(v, x, y);
    case 12: { uint32_t n; return try_parse_ipv4(v, &n) && n >= aux0 && n <= aux1; }
    case 14: return match_any_case_phrase(v, x);     // x = strings.ToLower(phrase)
    case 15: return match_any_case_prefix(v, x);
    case 16: case 17: case 18: {                     // x = phrase list: each phrase as varuint length + bytes
        std::vector<std::string> phrases;
        const uint8_t* p = (const uint8_t*)x.data(); size_t n = x.size();
        while (n) { uint64_t l; int k = get_varuint(p, n, &l); if (k <= 0 || l > n - (size_t)k) return -1; phrases.emplace_back((const char*)p + k, l); p += k + l; n -= (size_t)k + l; }
        return kind == 16 ? match_sequence(v, phrases) : kind == 17 ? match_all_phrases(v, phrases) : match_any_phrase(v, phrases);
    }
    }
    return -1;
}
// strings.ToLower
int64_t vlo_strings_to_lower(const void* s, uint64_t sl, char* out, uint64_t cap) {
    std::string r = strings_to_lower(sv((const char*)s, sl));
    if (r.size() > cap) return -1;
    memcpy(out, r.data(), r.size());
    return (int64_t)r.size();
}
int64_t vlo_skip_first_last_token(const void* s, uint64_t sl, char* out, uint64_t cap) {
    std::string r = skip_first_last_token(sv((const char*)s, sl));
    if (r.size() > cap) return -1;
    memcpy(out, r.data(), r.size()); return (int64_t)r.size();
}
// returns 1/0, or -1 on regex compile error
int vlo_regex_match(const void* expr, uint64_t el, const void* s, uint64_t sl) {
    int r = -1;
    if (guard([&] { Regex re(sv((const char*)expr, el)); r = re.match_string(sv((const char*)s, sl)) ? 1 : 0; })) return -1;
    return r;
}
// Describe regexutil.Regex fields as text (for tests / debugging)
int64_t vlo_regex_describe(const void* expr, uint64_t el, char* out, uint64_t cap) {
    std::string d;
    if (guard([&] {
            Regex re(sv((const char*)expr, el));
            d = "prefix=" + re.prefix + "\nisOnlyPrefix=" + std::to_string(re.isOnlyPrefix) + "\nisSuffixDotStar=" + std::to_string(re.isSuffixDotStar) +
                "\nisSuffixDotPlus=" + std::to_string(re.isSuffixDotPlus) + "\nsubstrDotStar=" + re.substrDotStar + "\nsubstrDotPlus=" + re.substrDotPlus + "\norValues=";
            for (auto& v : re.orValues) d += "[" + v + "]";
            d += "\nliterals=";
            for (auto& v : re.get_literals()) d += "[" + v + "]";
        })) return -1;
    if (d.size() > cap) return -1;
    memcpy(out, d.data(), d.size()); return (int64_t)d.size();
}

// parse helpers: return 1 ok / 0 fail
int vlo_try_parse_uint64(const void* s, uint64_t n, uint64_t* out) { return try_parse_uint64(sv((const char*)s, n), out); }
int vlo_try_parse_int64(const void* s, uint64_t n, int64_t* out) { return try_parse_int64(sv((const char*)s, n), out); }
int vlo_try_parse_float64(const void* s, uint64_t n, double* out) { return try_parse_float64_exact(sv((const char*)s, n), out); }
int vlo_try_parse_ipv4(const void* s, uint64_t n, uint32_t* out) { return try_parse_ipv4(sv((const char*)s, n), out); }
int vlo_try_parse_iso8601(const void* s, uint64_t n, int64_t* out) { return try_parse_timestamp_iso8601(sv((const char*)s, n), out); }
int64_t vlo_encoded_to_string(int vt, const void* v, uint64_t n, char* out, uint64_t cap) {
    std::string r;
    if (guard([&] { r = encoded_to_string((uint8_t)vt, sv((const char*)v, n)); })) return -1;
    if (r.size() > cap) return -1;
    memcpy(out, r.data(), r.size()); return (int64_t)r.size();
}

// strings block codec
int64_t vlo_marshal_strings_block(const uint8_t* blob, const uint64_t* offs, uint64_t n, uint8_t* out, uint64_t cap) {
    std::string r;
    if (guard([&] { r = marshal_strings_block(unpack(blob, offs, n)); })) return -1;
    if (r.size() > cap) return -1;
    memcpy(out, r.data(), r.size()); return (int64_t)r.size();
}
// decodes to the post-zstd stage; lens_items/data buffers; returns 0 ok
int vlo_decode_values_block(const uint8_t* src, uint64_t n, uint8_t* lens_out, uint64_t* lens_len, uint8_t* data_out, uint64_t* data_len) {
    return guard([&] {
        DecodedStringsBlock d = decode_values_block_stage(sv((const char*)src, n));
        if (d.lens_items.size() > *lens_len || d.data.size() > *data_len) throw std::runtime_error("output buffer too small");
        memcpy(lens_out, d.lens_items.data(), d.lens_items.size()); *lens_len = d.lens_items.size();
        memcpy(data_out, d.data.data(), d.data.size()); *data_len = d.data.size();
    });
}
// full unmarshal into packed strings: out_offs has items+1 entries
int vlo_unmarshal_strings_block(const uint8_t* src, uint64_t n, uint64_t items, uint8_t* out, uint64_t cap, uint64_t* out_offs) {
    return guard([&] {
        DecodedStringsBlock d = decode_values_block_stage(sv((const char*)src, n));
        auto vals = unmarshal_strings(d, items);
        uint64_t off = 0;
        for (uint64_t i = 0; i < items; i++) {
            if (off + vals[i].size() > cap) throw std::runtime_error("output buffer too small");
            memcpy(out + off, vals[i].data(), vals[i].size()); out_offs[i] = off; off += vals[i].size();
        }
        out_offs[items] = off;
    });
}

// ---- blocks ------------------------------------------------------------------------------------------------------
// names: packed list of ncols names; values: packed list of ncols*rows strings, column-major
void* vlo_block_build(const uint8_t* names_blob, const uint64_t* names_offs, uint64_t ncols, const uint8_t* vals_blob, const uint64_t* vals_offs, uint64_t rows) {
    BlockHandle* h = nullptr;
    if (guard([&] {
            auto names = unpack(names_blob, names_offs, ncols);
            auto vals = unpack(vals_blob, vals_offs, ncols * rows);
            std::vector<std::string> nm; std::vector<std::vector<sv>> cols(ncols);
            for (uint64_t c = 0; c < ncols; c++) { nm.emplace_back(names[c]); cols[c].assign(vals.begin() + c * rows, vals.begin() + (c + 1) * rows); }
            h = new BlockHandle{build_block(nm, cols, rows)};
        })) return nullptr;
    return h;
}
void vlo_block_free(void* h) { delete (BlockHandle*)h; }
int vlo_block_set_timestamps(void* h, const int64_t* ts, uint64_t n) { return guard([&] { ((BlockHandle*)h)->b.set_timestamps(std::vector<int64_t>(ts, ts + n)); }); }
// encoded timestamps of a block: pointer + length, marshal type, min / max (timestampsHeader)
int vlo_block_timestamps(void* h, const uint8_t** data, uint64_t* len, int* marshal_type, int64_t* min_ts, int64_t* max_ts) {
    Block& b = ((BlockHandle*)h)->b;
    if (!b.hasTimestamps) return -1;
    *data = (const uint8_t*)b.ts.data.data(); *len = b.ts.data.size(); *marshal_type = b.ts.mt; *min_ts = b.minTimestamp; *max_ts = b.maxTimestamp;
    return 0;
}
// encoding.MarshalTimestamps(ts, 64) / UnmarshalTimestamps
int64_t vlo_marshal_timestamps(const int64_t* ts, uint64_t n, uint8_t* out, uint64_t cap, int* marshal_type, int64_t* first) {
    int64_t r = -1;
    guard([&] { EncodedInt64s e = marshal_int64_array(std::vector<int64_t>(ts, ts + n)); if (e.data.size() > cap) throw std::runtime_error("output buffer too small");
                memcpy(out, e.data.data(), e.data.size()); *marshal_type = e.mt; *first = e.first; r = (int64_t)e.data.size(); });
    return r;
}
int vlo_unmarshal_timestamps(const uint8_t* src, uint64_t n, int marshal_type, int64_t first, uint64_t items, int64_t* out) {
    return guard([&] { auto v = unmarshal_int64_array(sv((const char*)src, n), (uint8_t)marshal_type, first, items); memcpy(out, v.data(), v.size() * 8); });
}
void* vlo_filter_day_range(int64_t start, int64_t end, int64_t offset) { return new FilterHandle{std::make_shared<FilterDayRange>(start, end, offset)}; }
void* vlo_filter_week_range(int start_day, int end_day, int64_t offset) { return new FilterHandle{std::make_shared<FilterWeekRange>(start_day, end_day, offset)}; }
void* vlo_filter_time(int64_t mn, int64_t mx) { return new FilterHandle{std::make_shared<FilterTime>(mn, mx)}; }
uint64_t vlo_block_rows(void* h) { return ((BlockHandle*)h)->b.rows; }
uint64_t vlo_block_ncolumns(void* h) { return ((BlockHandle*)h)->b.columns.size(); }
uint64_t vlo_block_nconsts(void* h) { return ((BlockHandle*)h)->b.consts.size(); }
// column accessors: pointers stay valid while the block lives
struct vlo_column_view {
    const char* name; uint64_t name_len;
    uint32_t value_type; uint32_t dict_len;
    uint64_t min_value, max_value;
    const char* dict_ptr[8]; uint64_t dict_lens[8];
    const uint8_t* values_block; uint64_t values_block_len;
    const uint8_t* bloom; uint64_t bloom_len;
};
void vlo_block_column(void* h, uint64_t i, vlo_column_view* v) {
    const Column& c = ((BlockHandle*)h)->b.columns[i];
    memset(v, 0, sizeof *v);
    v->name = c.name.data(); v->name_len = c.name.size(); v->value_type = c.valueType; v->dict_len = (uint32_t)c.dict.size();
    v->min_value = c.minValue; v->max_value = c.maxValue;
    for (size_t k = 0; k < c.dict.size(); k++) { v->dict_ptr[k] = c.dict[k].data(); v->dict_lens[k] = c.dict[k].size(); }
    v->values_block = (const uint8_t*)c.valuesBlock.data(); v->values_block_len = c.valuesBlock.size();
    v->bloom = (const uint8_t*)c.bloom.data(); v->bloom_len = c.bloom.size();
}
void vlo_block_const(void* h, uint64_t i, const char** name, uint64_t* name_len, const char** value, uint64_t* value_len) {
    const ConstColumn& c = ((BlockHandle*)h)->b.consts[i];
    *name = c.name.data(); *name_len = c.name.size(); *value = c.value.data(); *value_len = c.value.size();
}

// ---- filters -----------------------------------------------------------------------------------------------------
void* vlo_filter_phrase(const void* f, uint64_t fl, const void* p, uint64_t pl) { return new FilterHandle{std::make_shared<FilterPhrase>(sv((const char*)f, fl), sv((const char*)p, pl))}; }
void* vlo_filter_prefix(const void* f, uint64_t fl, const void* p, uint64_t pl) { return new FilterHandle{std::make_shared<FilterPrefix>(sv((const char*)f, fl), sv((const char*)p, pl))}; }
void* vlo_filter_exact(const void* f, uint64_t fl, const void* p, uint64_t pl) { return new FilterHandle{std::make_shared<FilterExact>(sv((const char*)f, fl), sv((const char*)p, pl))}; }
void* vlo_filter_in(const void* f, uint64_t fl, const uint8_t* blob, const uint64_t* offs, uint64_t n) {
    std::vector<std::string> vals; for (auto v : unpack(blob, offs, n)) vals.emplace_back(v);
    return new FilterHandle{std::make_shared<FilterIn>(sv((const char*)f, fl), vals)};
}
void* vlo_filter_exact_prefix(const void* f, uint64_t fl, const void* p, uint64_t pl) { return new FilterHandle{std::make_shared<FilterExactPrefix>(sv((const char*)f, fl), sv((const char*)p, pl))}; }
void* vlo_filter_sequence(const void* f, uint64_t fl, const uint8_t* blob, const uint64_t* offs, uint64_t n) {
    std::vector<std::string> vals; for (auto v : unpack(blob, offs, n)) vals.emplace_back(v);
    return new FilterHandle{std::make_shared<FilterSequence>(sv((const char*)f, fl), vals)};
}
void* vlo_filter_contains_all(const void* f, uint64_t fl, const uint8_t* blob, const uint64_t* offs, uint64_t n) {
    std::vector<std::string> vals; for (auto v : unpack(blob, offs, n)) vals.emplace_back(v);
    return new FilterHandle{std::make_shared<FilterContainsAll>(sv((const char*)f, fl), vals)};
}
void* vlo_filter_contains_any(const void* f, uint64_t fl, const uint8_t* blob, const uint64_t* offs, uint64_t n) {
    std::vector<std::string> vals; for (auto v : unpack(blob, offs, n)) vals.emplace_back(v);
    return new FilterHandle{std::make_shared<FilterContainsAny>(sv((const char*)f, fl), vals)};
}
void* vlo_filter_any_case_phrase(const void* f, uint64_t fl, const void* p, uint64_t pl) { return new FilterHandle{std::make_shared<FilterAnyCasePhrase>(sv((const char*)f, fl), sv((const char*)p, pl))}; }
void* vlo_filter_any_case_prefix(const void* f, uint64_t fl, const void* p, uint64_t pl) { return new FilterHandle{std::make_shared<FilterAnyCasePrefix>(sv((const char*)f, fl), sv((const char*)p, pl))}; }
void* vlo_filter_value_type(const void* f, uint64_t fl, const void* p, uint64_t pl) { return new FilterHandle{std::make_shared<FilterValueType>(sv((const char*)f, fl), sv((const char*)p, pl))}; }
void* vlo_filter_eq_field(const void* f, uint64_t fl, const void* p, uint64_t pl) { return new FilterHandle{std::make_shared<FilterEqField>(sv((const char*)f, fl), sv((const char*)p, pl))}; }
void* vlo_filter_range(const void* f, uint64_t fl, double mn, double mx) { return new FilterHandle{std::make_shared<FilterRange>(sv((const char*)f, fl), mn, mx)}; }
void* vlo_filter_le_field(const void* f, uint64_t fl, const void* p, uint64_t pl, int exclude_equal) {
    return new FilterHandle{std::make_shared<FilterLeField>(sv((const char*)f, fl), sv((const char*)p, pl), exclude_equal != 0)};
}
double vlo_parse_math_number(const void* s, uint64_t n) { return parse_math_number(sv((const char*)s, n)); }
void* vlo_filter_len_range(const void* f, uint64_t fl, uint64_t mn, uint64_t mx) { return new FilterHandle{std::make_shared<FilterLenRange>(sv((const char*)f, fl), mn, mx)}; }
void* vlo_filter_string_range(const void* f, uint64_t fl, const void* a, uint64_t al, const void* b, uint64_t bl) {
    return new FilterHandle{std::make_shared<FilterStringRange>(sv((const char*)f, fl), sv((const char*)a, al), sv((const char*)b, bl))};
}
void* vlo_filter_ipv4_range(const void* f, uint64_t fl, uint32_t mn, uint32_t mx) { return new FilterHandle{std::make_shared<FilterIPv4Range>(sv((const char*)f, fl), mn, mx)}; }
void* vlo_filter_regexp(const void* f, uint64_t fl, const void* p, uint64_t pl) {
    FilterHandle* h = nullptr;
    if (guard([&] { h = new FilterHandle{std::make_shared<FilterRegexp>(sv((const char*)f, fl), sv((const char*)p, pl))}; })) return nullptr;
    return h;
}
void* vlo_filter_noop() { return new FilterHandle{std::make_shared<FilterNoop>()}; }
void* vlo_filter_and(void** hs, uint64_t n) { std::vector<FP> v; for (uint64_t i = 0; i < n; i++) v.push_back(((FilterHandle*)hs[i])->f); return new FilterHandle{std::make_shared<FilterAnd>(v)}; }
void* vlo_filter_or(void** hs, uint64_t n) { std::vector<FP> v; for (uint64_t i = 0; i < n; i++) v.push_back(((FilterHandle*)hs[i])->f); return new FilterHandle{std::make_shared<FilterOr>(v)}; }
void* vlo_filter_not(void* h) { return new FilterHandle{std::make_shared<FilterNot>(((FilterHandle*)h)->f)}; }
void vlo_filter_free(void* h) { delete (FilterHandle*)h; }
// tokens of a leaf filter joined by '\n' (for tests of getTokens())
int64_t vlo_filter_tokens(void* h, char* out, uint64_t cap) {
    std::string field; std::vector<std::string> toks;
    if (!((FilterHandle*)h)->f->leaf_tokens(&field, &toks)) return -1;
    std::string s; for (size_t i = 0; i < toks.size(); i++) { if (i) s.push_back('\n'); s += toks[i]; }
    if (s.size() > cap) return -1;
    memcpy(out, s.data(), s.size()); return (int64_t)s.size();
}

// The per-field tokens of the bloom pre-pass of every AND / OR node of a tree (filterAnd.byFieldTokens / filterOr.byFieldTokens), in pre-order:
// one line per node: "A" or "O", then for each field "\t" field "\x1f" token "\x1f" token ...
static void dump_prepass(const FP& f, std::string& out) {
    if (auto* a = dynamic_cast<FilterAnd*>(f.get())) {
        out += "A"; for (auto& ft : a->by_field_tokens()) { out += "\t" + ft.field; for (auto& t : ft.tokens) out += "\x1f" + t; } out += "\n";
        for (auto& k : a->filters) dump_prepass(k, out);
    } else if (auto* o = dynamic_cast<FilterOr*>(f.get())) {
        out += "O"; for (auto& ft : o->by_field_tokens()) { out += "\t" + ft.field; for (auto& t : ft.tokens) out += "\x1f" + t; } out += "\n";
        for (auto& k : o->filters) dump_prepass(k, out);
    } else if (auto* n = dynamic_cast<FilterNot*>(f.get())) dump_prepass(n->f, out);
}
// in(): the hashes of the common tokens, then of every per-value token set (in_values.go:317-371): out = [n_common, common..., n_sets, {n, hashes...}...]
int64_t vlo_filter_in_hashes(void* h, uint64_t* out, uint64_t cap) {
    auto* f = dynamic_cast<FilterIn*>(((FilterHandle*)h)->f.get());
    if (!f) return -1;
    std::vector<uint64_t> v;
    v.push_back(f->commonHashes.size()); v.insert(v.end(), f->commonHashes.begin(), f->commonHashes.end());
    v.push_back(f->tokenSetsHashes.size());
    for (auto& s : f->tokenSetsHashes) { v.push_back(s.size()); v.insert(v.end(), s.begin(), s.end()); }
    if (v.size() > cap) return -1;
    memcpy(out, v.data(), v.size() * 8);
    return (int64_t)v.size();
}
// in(): the typed value set for a column of value type vt (in_values.go:141-315) as sorted u64 (uintN / ipv4: the number; int64: zig-zag;
// float64: the bits; iso8601: nanoseconds)
int64_t vlo_filter_in_typed(void* h, int vt, uint64_t* out, uint64_t cap) {
    auto* f = dynamic_cast<FilterIn*>(((FilterHandle*)h)->f.get());
    if (!f) return -1;
    std::vector<uint64_t> v;
    for (auto& b : f->bin_values((uint8_t)vt)) { uint64_t x = 0; for (unsigned char c : b) x = (x << 8) | c; v.push_back(x); }
    std::sort(v.begin(), v.end());
    if (v.size() > cap) return -1;
    memcpy(out, v.data(), v.size() * 8);
    return (int64_t)v.size();
}
int64_t vlo_filter_prepass_tokens(void* h, char* out, uint64_t cap) {
    int64_t r = -1;
    guard([&] { std::string s; dump_prepass(((FilterHandle*)h)->f, s); if (s.size() > cap) throw std::runtime_error("output buffer too small"); memcpy(out, s.data(), s.size()); r = (int64_t)s.size(); });
    return r;
}

// blockSearch.search for one block: out_words must hold ceil(rows/64) u64. stats (6 u64, may be NULL) are ACCUMULATED.
int vlo_block_search(void* block, void* filter, uint64_t* out_words, uint64_t* stats6) {
    return guard([&] {
        BlockSearch bs; Bitmap bm; ScanStats st;
        block_search(bs, ((BlockHandle*)block)->b, *((FilterHandle*)filter)->f, bm, stats6 ? &st : nullptr);
        memcpy(out_words, bm.a.data(), bm.a.size() * 8);
        if (stats6) { stats6[0] += st.blocks; stats6[1] += st.rows; stats6[2] += st.bloom_probe_bytes; stats6[3] += st.values_bytes; stats6[4] += st.bitmap_bytes; stats6[5] += st.blocks_values_read; }
    });
}

// ---- synthetic generator (CPU restatement of victorialogs_b200/csrc/gen.cuh; see vlo_gen.h) ---------------------------
void* vlo_gen_block(const vlo_gen_config* cfg, uint64_t block_id) {
    BlockHandle* h = nullptr;
    if (guard([&] { h = new BlockHandle{gen_block(*cfg, block_id)}; })) return nullptr;
    return h;
}
// raw generated rows of one column (before encoding), packed; for tests
int vlo_gen_rows(const vlo_gen_config* cfg, uint64_t block_id, int column, uint8_t* out, uint64_t cap, uint64_t* out_offs) {
    return guard([&] {
        uint64_t rows = gen_block_rows(*cfg, block_id);
        uint64_t off = 0;
        for (uint64_t i = 0; i < rows; i++) {
            std::string s = gen_value(*cfg, block_id, i, column);
            if (off + s.size() > cap) throw std::runtime_error("output buffer too small");
            memcpy(out + off, s.data(), s.size()); out_offs[i] = off; off += s.size();
        }
        out_offs[rows] = off;
    });
}

// Multi-threaded scan over generated blocks [block_lo, block_hi): the CPU baseline ("port").
// Blocks are statically sharded across threads like Storage.search workers (storage_search.go:1040-1067).
// Blocks are built outside the timed region; the timed region is `passes` x blockSearch over the pre-built (ZSTD-compressed)
// blocks ("with-zstd" variant of BASELINE.md); *secs covers all passes.
// Returns seconds spent scanning in *secs; accumulates stats; out_counts (may be NULL) gets per-block match counts;
// out_digest gets xor of xxh64(bitmap words) keyed by block id.
// flags: bit 0 = "post-zstd" variant (values blocks decompressed before the timed region: the input stage of the device-resident scan),
// bit 1 = pin worker t to CPU t % ncpu (steadier numbers on a shared host).
int vlo_scan_generated(const vlo_gen_config* cfg, void* filter, uint64_t block_lo, uint64_t block_hi, int threads, int passes, double* secs,
                       uint64_t* stats6, uint32_t* out_counts, uint64_t* out_digest, uint64_t* total_matches, int flags) {
    return guard([&] {
        uint64_t nb = block_hi - block_lo;
        std::vector<Block> blocks(nb);
        const bool post_zstd = flags & 1, pin = flags & 2;
        {
            std::vector<std::thread> th; std::atomic<uint64_t> next{0};
            for (int t = 0; t < threads; t++) th.emplace_back([&] {
                for (;;) {
                    uint64_t i = next++; if (i >= nb) break;
                    blocks[i] = gen_block(*cfg, block_lo + i);
                    if (post_zstd) for (Column& c : blocks[i].columns) c.predecoded = std::make_shared<const DecodedStringsBlock>(decode_values_block_stage(c.valuesBlock));
                }
            });
            for (auto& t : th) t.join();
        }
        std::vector<ScanStats> sts(threads); std::vector<uint64_t> digs(threads, 0), tots(threads, 0);
        std::vector<std::string> errs(threads);
        // rebuild the filter per thread is unnecessary: filters are immutable after by_field_tokens() is initialised
        Filter& f = *((FilterHandle*)filter)->f;
        { BlockSearch bs; Bitmap bm; if (nb) block_search(bs, blocks[0], f, bm, nullptr); }   // warm lazily-initialised token caches
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        const unsigned ncpu = std::max(1u, std::thread::hardware_concurrency());
        for (int t = 0; t < threads; t++) th.emplace_back([&, t] {
            try {
                if (pin) { cpu_set_t set; CPU_ZERO(&set); CPU_SET((unsigned)t % ncpu, &set); pthread_setaffinity_np(pthread_self(), sizeof set, &set); }
                BlockSearch bs; Bitmap bm;
                uint64_t lo = nb * t / threads, hi = nb * (t + 1) / threads;
                for (int pass = 0; pass < passes; pass++)
                for (uint64_t i = lo; i < hi; i++) {
                    if (pass) { block_search(bs, blocks[i], f, bm, nullptr); continue; }   // extra passes only add time (same work)
                    block_search(bs, blocks[i], f, bm, &sts[t]);
                    uint64_t ones = bm.ones();
                    if (out_counts) out_counts[i] = (uint32_t)ones;
                    tots[t] += ones;
                    uint64_t key = block_lo + i;
                    digs[t] ^= xxh64(bm.a.data(), bm.a.size() * 8) * (2 * key + 1);
                }
            } catch (const std::exception& e) { errs[t] = e.what(); }
        });
        for (auto& t : th) t.join();
        *secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (auto& e : errs) if (!e.empty()) throw std::runtime_error(e);
        uint64_t dig = 0, tot = 0;
        for (int t = 0; t < threads; t++) {
            dig ^= digs[t]; tot += tots[t];
            if (stats6) { stats6[0] += sts[t].blocks; stats6[1] += sts[t].rows; stats6[2] += sts[t].bloom_probe_bytes; stats6[3] += sts[t].values_bytes; stats6[4] += sts[t].bitmap_bytes; stats6[5] += sts[t].blocks_values_read; }
        }
        if (out_digest) *out_digest = dig;
        if (total_matches) *total_matches = tot;
    });
}


// The scenario of TestBitmap (lib/logstorage/bitmap_test.go:7-133) on the restated bitmap; returns 0, or 1000*bits + the number of the
// check that failed.
int vlo_bitmap_selftest(int max_bits) {
    for (int i = 0; i < max_bits; i++) {
        auto fail = [&](int step) { return 1000 * i + step; };
        Bitmap bm; bm.init((uint64_t)i);
        if (bm.bitsLen != (uint64_t)i) return fail(1);
        if (!bm.is_zero()) return fail(2);
        if (i == 0 && !bm.are_all_bits_set()) return fail(3);
        if (i > 0 && bm.are_all_bits_set()) return fail(4);
        if (bm.ones() != 0) return fail(5);
        bm.set_bits();
        if (bm.ones() != (uint64_t)i) return fail(6);
        uint64_t next = 0; bool ok = true;
        bm.for_each_set_bit_readonly([&](uint64_t idx) { if (idx >= (uint64_t)i || idx != next) ok = false; next++; });
        if (!ok || next != (uint64_t)i) return fail(7);
        if (!bm.are_all_bits_set()) return fail(8);
        bm.for_each_set_bit([&](uint64_t idx) { return idx % 2 != 0; });   // clear the even bits
        if (i <= 1 && !bm.is_zero()) return fail(9);
        if (i > 1 && bm.is_zero()) return fail(10);
        if (i == 0 && !bm.are_all_bits_set()) return fail(11);
        if (i > 0 && bm.are_all_bits_set()) return fail(12);
        next = 1; ok = true;
        bm.for_each_set_bit_readonly([&](uint64_t idx) { if (idx != next) ok = false; next += 2; });
        if (!ok || next < (uint64_t)i) return fail(13);
        bm.for_each_set_bit([&](uint64_t) { return false; });   // clear all
        if (!bm.is_zero()) return fail(14);
        if (i == 0 && !bm.are_all_bits_set()) return fail(15);
        if (i > 0 && bm.are_all_bits_set()) return fail(16);
        if (bm.ones() != 0) return fail(17);
        uint64_t cnt = 0; bm.for_each_set_bit_readonly([&](uint64_t) { cnt++; });
        if (cnt) return fail(18);
        for (int k = 0; k < i; k++) {
            if (bm.ones() != (uint64_t)k) return fail(19);
            if (bm.is_set_bit((uint64_t)k)) return fail(20);
            bm.set_bit((uint64_t)k);
            if (!bm.is_set_bit((uint64_t)k)) return fail(21);
            if (bm.ones() != (uint64_t)k + 1) return fail(22);
        }
        // andNot (bitmap.go:99-111) against an independent model
        Bitmap x; x.init((uint64_t)i);
        for (int k = 0; k < i; k += 3) x.set_bit((uint64_t)k);
        bm.and_not(x);
        for (int k = 0; k < i; k++) if (bm.is_set_bit((uint64_t)k) != (k % 3 != 0)) return fail(23);
    }
    return 0;
}

// ---- part files (vlo_part.h) -----------------------------------------------------------------------------------------------
// Records travel as flat u64 arrays:
//   blockHeader[15]      = accountID, projectID, id.hi, id.lo, uncompressedSizeBytes, rowsCount, th.blockOffset, th.blockSize, th.minTimestamp,
//                          th.maxTimestamp, th.marshalType, columnsHeaderIndexOffset, columnsHeaderIndexSize, columnsHeaderOffset, columnsHeaderSize
//   indexBlockHeader[8]  = accountID, projectID, id.hi, id.lo, minTimestamp, maxTimestamp, indexBlockOffset, indexBlockSize
//   columnHeader[7]      = valueType, minValue, maxValue, valuesOffset, valuesSize, bloomFilterOffset, bloomFilterSize
//   partHeader[8]        = FormatVersion, CompressedSizeBytes, UncompressedSizeBytes, RowsCount, BlocksCount, MinTimestamp, MaxTimestamp, BloomValuesShardsCount
static void bh_from(const uint64_t* f, BlockHeader& b) {
    b.streamID = StreamID{(uint32_t)f[0], (uint32_t)f[1], f[2], f[3]}; b.uncompressedSizeBytes = f[4]; b.rowsCount = f[5];
    b.timestampsHeader.blockOffset = f[6]; b.timestampsHeader.blockSize = f[7]; b.timestampsHeader.minTimestamp = (int64_t)f[8]; b.timestampsHeader.maxTimestamp = (int64_t)f[9];
    b.timestampsHeader.marshalType = (uint8_t)f[10]; b.columnsHeaderIndexOffset = f[11]; b.columnsHeaderIndexSize = f[12]; b.columnsHeaderOffset = f[13]; b.columnsHeaderSize = f[14];
}
static void bh_to(const BlockHeader& b, uint64_t* f) {
    f[0] = b.streamID.accountID; f[1] = b.streamID.projectID; f[2] = b.streamID.hi; f[3] = b.streamID.lo; f[4] = b.uncompressedSizeBytes; f[5] = b.rowsCount;
    f[6] = b.timestampsHeader.blockOffset; f[7] = b.timestampsHeader.blockSize; f[8] = (uint64_t)b.timestampsHeader.minTimestamp; f[9] = (uint64_t)b.timestampsHeader.maxTimestamp;
    f[10] = b.timestampsHeader.marshalType; f[11] = b.columnsHeaderIndexOffset; f[12] = b.columnsHeaderIndexSize; f[13] = b.columnsHeaderOffset; f[14] = b.columnsHeaderSize;
}
static void ih_to(const IndexBlockHeader& h, uint64_t* f) {
    f[0] = h.streamID.accountID; f[1] = h.streamID.projectID; f[2] = h.streamID.hi; f[3] = h.streamID.lo; f[4] = (uint64_t)h.minTimestamp; f[5] = (uint64_t)h.maxTimestamp; f[6] = h.indexBlockOffset; f[7] = h.indexBlockSize;
}
static void ch_from(const uint64_t* f, ColumnHeader& c) { c.valueType = (uint8_t)f[0]; c.minValue = f[1]; c.maxValue = f[2]; c.valuesOffset = f[3]; c.valuesSize = f[4]; c.bloomFilterOffset = f[5]; c.bloomFilterSize = f[6]; }
static void ch_to(const ColumnHeader& c, uint64_t* f) { f[0] = c.valueType; f[1] = c.minValue; f[2] = c.maxValue; f[3] = c.valuesOffset; f[4] = c.valuesSize; f[5] = c.bloomFilterOffset; f[6] = c.bloomFilterSize; }
static int64_t emit(const std::string& s, uint8_t* out, uint64_t cap) { if (s.size() > cap) throw std::runtime_error("output buffer too small"); memcpy(out, s.data(), s.size()); return (int64_t)s.size(); }

int64_t vlo_part_marshal_block_header(const uint64_t* f15, uint8_t* out, uint64_t cap) {
    int64_t r = -1; guard([&] { BlockHeader b; bh_from(f15, b); std::string s; b.marshal(s); r = emit(s, out, cap); }); return r;
}
// returns the number of records (validated like unmarshalBlockHeaders) or -1
int64_t vlo_part_unmarshal_block_headers(const uint8_t* src, uint64_t n, unsigned format_version, uint64_t* out, uint64_t cap_records) {
    int64_t r = -1;
    guard([&] { auto v = unmarshal_block_headers(sv((const char*)src, n), format_version); if (v.size() > cap_records) throw std::runtime_error("output buffer too small");
                for (size_t i = 0; i < v.size(); i++) bh_to(v[i], out + 15 * i);
                r = (int64_t)v.size(); });
    return r;
}
int64_t vlo_part_marshal_index_block_header(const uint64_t* f8, uint8_t* out, uint64_t cap) {
    int64_t r = -1;
    guard([&] { IndexBlockHeader h; h.streamID = StreamID{(uint32_t)f8[0], (uint32_t)f8[1], f8[2], f8[3]}; h.minTimestamp = (int64_t)f8[4]; h.maxTimestamp = (int64_t)f8[5]; h.indexBlockOffset = f8[6]; h.indexBlockSize = f8[7];
                std::string s; h.marshal(s); r = emit(s, out, cap); });
    return r;
}
int64_t vlo_part_unmarshal_index_block_headers(const uint8_t* src, uint64_t n, uint64_t* out, uint64_t cap_records) {
    int64_t r = -1;
    guard([&] { auto v = unmarshal_index_block_headers(sv((const char*)src, n)); if (v.size() > cap_records) throw std::runtime_error("output buffer too small");
                for (size_t i = 0; i < v.size(); i++) ih_to(v[i], out + 8 * i);
                r = (int64_t)v.size(); });
    return r;
}
int64_t vlo_part_marshal_column_header(const uint64_t* f7, const uint8_t* dict_blob, const uint64_t* dict_offs, uint64_t ndict, uint8_t* out, uint64_t cap) {
    int64_t r = -1;
    guard([&] { ColumnHeader c; ch_from(f7, c); for (auto d : unpack(dict_blob, dict_offs, ndict)) c.dict.emplace_back(d); std::string s; c.marshal(s); r = emit(s, out, cap); });
    return r;
}
// returns the bytes consumed or -1; dict values come back packed (dict_offs has room for 257 entries)
int64_t vlo_part_unmarshal_column_header(const uint8_t* src, uint64_t n, unsigned format_version, uint64_t* f7, uint8_t* dict_out, uint64_t dict_cap, uint64_t* dict_offs, uint64_t* ndict) {
    int64_t r = -1;
    guard([&] { PReader rd(sv((const char*)src, n)); ColumnHeader c; c.unmarshal(rd, format_version); ch_to(c, f7);
                std::string cat; dict_offs[0] = 0; for (size_t i = 0; i < c.dict.size(); i++) { cat += c.dict[i]; dict_offs[i + 1] = cat.size(); }
                emit(cat, dict_out, dict_cap); *ndict = c.dict.size(); r = (int64_t)(n - rd.n); });
    return r;
}
// refs: pairs (columnNameID, offset)
int64_t vlo_part_marshal_columns_header_index(const uint64_t* refs, uint64_t nrefs, const uint64_t* crefs, uint64_t ncrefs, uint8_t* out, uint64_t cap) {
    int64_t r = -1;
    guard([&] { ColumnsHeaderIndex x; for (uint64_t i = 0; i < nrefs; i++) x.columnHeadersRefs.push_back({refs[2 * i], refs[2 * i + 1]}); for (uint64_t i = 0; i < ncrefs; i++) x.constColumnsRefs.push_back({crefs[2 * i], crefs[2 * i + 1]});
                std::string s; x.marshal(s); r = emit(s, out, cap); });
    return r;
}
int vlo_part_unmarshal_columns_header_index(const uint8_t* src, uint64_t n, uint64_t* refs, uint64_t* nrefs, uint64_t* crefs, uint64_t* ncrefs, uint64_t cap_pairs) {
    return guard([&] { ColumnsHeaderIndex x; x.unmarshal(sv((const char*)src, n));
                       if (x.columnHeadersRefs.size() > cap_pairs || x.constColumnsRefs.size() > cap_pairs) throw std::runtime_error("output buffer too small");
                       for (size_t i = 0; i < x.columnHeadersRefs.size(); i++) { refs[2 * i] = x.columnHeadersRefs[i].columnNameID; refs[2 * i + 1] = x.columnHeadersRefs[i].offset; }
                       for (size_t i = 0; i < x.constColumnsRefs.size(); i++) { crefs[2 * i] = x.constColumnsRefs[i].columnNameID; crefs[2 * i + 1] = x.constColumnsRefs[i].offset; }
                       *nrefs = x.columnHeadersRefs.size(); *ncrefs = x.constColumnsRefs.size(); });
}
// columnsHeader.marshal with a fresh columnNameIDGenerator: names = ncols column names then nconst const-column names; values = nconst values.
// Dict columns are not supported by this entry point.  out gets the columnsHeader, idx_out the columnsHeaderIndex.
int64_t vlo_part_marshal_columns_header(uint64_t ncols, const uint64_t* f7s, const uint8_t* names_blob, const uint64_t* names_offs, uint64_t nconst, const uint8_t* vals_blob, const uint64_t* vals_offs,
                                        uint8_t* out, uint64_t cap, uint8_t* idx_out, uint64_t idx_cap, uint64_t* idx_len) {
    int64_t r = -1;
    guard([&] { auto names = unpack(names_blob, names_offs, ncols + nconst); auto vals = unpack(vals_blob, vals_offs, nconst);
                ColumnsHeader csh;
                for (uint64_t i = 0; i < ncols; i++) { ColumnHeader c; ch_from(f7s + 7 * i, c); c.name = std::string(names[i]); csh.columnHeaders.push_back(c); }
                for (uint64_t i = 0; i < nconst; i++) csh.constColumns.push_back({std::string(names[ncols + i]), std::string(vals[i])});
                ColumnsHeaderIndex idx; ColumnNameIDGenerator g; std::string s, si; csh.marshal(s, idx, g); idx.marshal(si);
                *idx_len = (uint64_t)emit(si, idx_out, idx_cap); r = emit(s, out, cap); });
    return r;
}
// unmarshals a columnsHeader + its index, resolves names, marshals both again and reports whether the bytes are identical (1), differ (0) or are malformed (-1).
// The names come back joined by '\0' (columns first, then const columns).
int vlo_part_columns_header_roundtrip(const uint8_t* src, uint64_t n, const uint8_t* idx_src, uint64_t idx_n, const uint8_t* names_blob, const uint64_t* names_offs, uint64_t nnames,
                                      char* names_out, uint64_t names_cap, uint64_t* names_len) {
    int same = -1;
    guard([&] { std::vector<std::string> names; for (auto s : unpack(names_blob, names_offs, nnames)) names.emplace_back(s);
                ColumnsHeader csh; csh.unmarshal(sv((const char*)src, n), partFormatLatestVersion);
                ColumnsHeaderIndex idx; idx.unmarshal(sv((const char*)idx_src, idx_n)); csh.set_column_names(idx, names);
                ColumnNameIDGenerator g; for (auto& s : names) g.get(s);
                ColumnsHeaderIndex idx2; std::string s, si; csh.marshal(s, idx2, g); idx2.marshal(si);
                std::string cat; for (auto& c : csh.columnHeaders) { cat += c.name; cat.push_back('\0'); } for (auto& c : csh.constColumns) { cat += c.name; cat.push_back('\0'); }
                *names_len = (uint64_t)emit(cat, (uint8_t*)names_out, names_cap);
                same = s == std::string((const char*)src, n) && si == std::string((const char*)idx_src, idx_n); });
    return same;
}
int64_t vlo_part_header_json(const uint64_t* f8, char* out, uint64_t cap) {
    int64_t r = -1;
    guard([&] { PartHeader ph; ph.FormatVersion = f8[0]; ph.CompressedSizeBytes = f8[1]; ph.UncompressedSizeBytes = f8[2]; ph.RowsCount = f8[3]; ph.BlocksCount = f8[4]; ph.MinTimestamp = (int64_t)f8[5];
                ph.MaxTimestamp = (int64_t)f8[6]; ph.BloomValuesShardsCount = f8[7]; r = emit(ph.to_json(), (uint8_t*)out, cap); });
    return r;
}
static void ph_to(const PartHeader& ph, uint64_t* f) {
    f[0] = ph.FormatVersion; f[1] = ph.CompressedSizeBytes; f[2] = ph.UncompressedSizeBytes; f[3] = ph.RowsCount; f[4] = ph.BlocksCount; f[5] = (uint64_t)ph.MinTimestamp; f[6] = (uint64_t)ph.MaxTimestamp; f[7] = ph.BloomValuesShardsCount;
}
int vlo_part_header_parse(const char* json, uint64_t n, uint64_t* f8) { return guard([&] { ph_to(PartHeader::from_json(sv(json, n)), f8); }); }

struct PartWriterHandle { PartWriter w; std::vector<const std::string*> names; };
struct PartReaderHandle { PartFiles files; PartReader r; std::vector<BlockHeader> bhs; };

void* vlo_part_writer_new(uint64_t max_index_block, uint64_t max_shards) {
    auto* h = new PartWriterHandle; if (max_index_block) h->w.maxIndexBlock = max_index_block; if (max_shards) h->w.maxShards = max_shards; return h;
}
void vlo_part_writer_free(void* h) { delete (PartWriterHandle*)h; }
int vlo_part_writer_add_block(void* h, const uint64_t* sid4, void* block, uint64_t uncompressed_size) {
    return guard([&] { ((PartWriterHandle*)h)->w.write_block(StreamID{(uint32_t)sid4[0], (uint32_t)sid4[1], sid4[2], sid4[3]}, ((BlockHandle*)block)->b, uncompressed_size); });
}
int vlo_part_writer_finalize(void* h, uint64_t* f8) { return guard([&] { auto& w = ((PartWriterHandle*)h)->w; w.finalize(); ph_to(w.ph, f8); }); }
uint64_t vlo_part_writer_nfiles(void* h) { return ((PartWriterHandle*)h)->w.files.size(); }
void vlo_part_writer_file(void* h, uint64_t i, const char** name, uint64_t* name_len, const uint8_t** data, uint64_t* len) {
    auto it = ((PartWriterHandle*)h)->w.files.begin(); std::advance(it, (long)i);
    *name = it->first.data(); *name_len = it->first.size(); *data = (const uint8_t*)it->second.data(); *len = it->second.size();
}
// opens a part given as packed (file name, contents) lists; reads every index block up front
void* vlo_part_reader_open(const uint8_t* names_blob, const uint64_t* names_offs, const uint8_t* data_blob, const uint64_t* data_offs, uint64_t nfiles) {
    PartReaderHandle* h = new PartReaderHandle;
    if (guard([&] { auto names = unpack(names_blob, names_offs, nfiles); auto datas = unpack(data_blob, data_offs, nfiles);
                    for (uint64_t i = 0; i < nfiles; i++) h->files[std::string(names[i])] = std::string(datas[i]);
                    h->r.open(h->files);
                    for (auto& ih : h->r.indexBlockHeaders) { auto v = h->r.read_index_block(ih); h->bhs.insert(h->bhs.end(), v.begin(), v.end()); }
                    if (h->bhs.size() != h->r.ph.BlocksCount) throw PartError("the number of block headers differs from partHeader.BlocksCount");
                    uint64_t rows = 0; for (auto& b : h->bhs) rows += b.rowsCount;
                    if (rows != h->r.ph.RowsCount) throw PartError("the number of rows differs from partHeader.RowsCount"); })) { delete h; return nullptr; }
    return h;
}
void vlo_part_reader_free(void* h) { delete (PartReaderHandle*)h; }
void vlo_part_reader_header(void* h, uint64_t* f8) { ph_to(((PartReaderHandle*)h)->r.ph, f8); }
uint64_t vlo_part_reader_nindex(void* h) { return ((PartReaderHandle*)h)->r.indexBlockHeaders.size(); }
void vlo_part_reader_index_header(void* h, uint64_t i, uint64_t* f8) { ih_to(((PartReaderHandle*)h)->r.indexBlockHeaders[i], f8); }
uint64_t vlo_part_reader_nblocks(void* h) { return ((PartReaderHandle*)h)->bhs.size(); }
void vlo_part_reader_block_header(void* h, uint64_t i, uint64_t* f15) { bh_to(((PartReaderHandle*)h)->bhs[i], f15); }
void* vlo_part_reader_block(void* h, uint64_t i) {
    BlockHandle* b = nullptr; auto* p = (PartReaderHandle*)h;
    if (guard([&] { b = new BlockHandle{p->r.read_block(p->bhs.at(i))}; })) return nullptr;
    return b;
}
// column names joined by '\0', in columnNameID order
int64_t vlo_part_reader_column_names(void* h, char* out, uint64_t cap) {
    int64_t r = -1; guard([&] { std::string cat; for (auto& s : ((PartReaderHandle*)h)->r.columnNames) { cat += s; cat.push_back('\0'); } r = emit(cat, (uint8_t*)out, cap); }); return r;
}
// blockSearch.getColumnHeader: 1 found, 0 no such column in the block, -1 error
int vlo_part_reader_column_header(void* h, uint64_t i, const char* name, uint64_t name_len, uint64_t* f7) {
    int found = -1; auto* p = (PartReaderHandle*)h;
    guard([&] { ColumnHeader c; bool ok = p->r.get_column_header(p->bhs.at(i), sv(name, name_len), &c); if (ok) ch_to(c, f7); found = ok; });
    return found;
}
int64_t vlo_part_reader_const_value(void* h, uint64_t i, const char* name, uint64_t name_len, char* out, uint64_t cap) {
    int64_t r = -1; auto* p = (PartReaderHandle*)h;
    guard([&] { r = emit(p->r.get_const_column_value(p->bhs.at(i), sv(name, name_len)), (uint8_t*)out, cap); });
    return r;
}

}  // extern "C"
