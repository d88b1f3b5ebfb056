// Sanitizer build of the host-side parsers of libvlscan.so that read untrusted bytes:
//   tree  the filter-tree program compiler incl. the regexp compiler (victorialogs_b200/csrc/vl_program.h, vl_regex.h)
//   zstd  the bytes-block / ZSTD header walk                         (victorialogs_b200/csrc/vl_zstd_walk.h)
//   pred  the per-value predicates and formatters that the row kernels run on every value - host builds of the same host+device code
//         (vl_hd.cuh, vl_anycase.cuh): random values, needles and number bits in exact-size heap blocks
//   part  the part directory reader                                  (victorialogs_b200/csrc/vl_part.h; the seed directory is ONE part directory,
//         each iteration damages one of its files in a scratch copy; metadata frames are inflated with libzstd)
// usage: harness tree|zstd|part <seed directory> <iterations> <rng seed>
// Every input lives in a heap block of its exact size, so AddressSanitizer sees any read past its end; malformed input must end in
// the parser's own exception type.  Built and run by tests/test_host_asan_cpu.py with -fsanitize=address,undefined.
#include <dirent.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>
#define VL_PART_HEAP_FILES 1
#include "vl_program.h"
#include "vl_zstd_walk.h"
#include "vl_part.h"
#include "vl_anycase.cuh"

extern "C" {   // the image has libzstd.so.1 but no zstd.h
size_t ZSTD_decompress(void* dst, size_t dstCapacity, const void* src, size_t compressedSize);
unsigned ZSTD_isError(size_t code);
}

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static size_t below(size_t n) { return n ? (size_t)(rnd() % n) : 0; }

static std::vector<std::string> read_seeds(const char* dir) {
    std::vector<std::string> names, out;
    DIR* d = opendir(dir);
    if (!d) { perror(dir); exit(2); }
    while (dirent* e = readdir(d)) if (e->d_name[0] != '.') names.push_back(e->d_name);
    closedir(d);
    std::sort(names.begin(), names.end());
    for (auto& n : names) { std::ifstream f(std::string(dir) + "/" + n, std::ios::binary); out.emplace_back((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>()); }
    if (out.empty()) { fprintf(stderr, "no seeds in %s\n", dir); exit(2); }
    return out;
}

static std::string mutate(const std::vector<std::string>& seeds) {
    std::string b = seeds[below(seeds.size())];
    static const uint8_t edge[] = {0x00, 0x01, 0x7F, 0x80, 0xFF, 0xFE, 0x28, 0xB5, 0x2F, 0xFD};
    for (int k = (int)below(4); k >= 0; k--) {
        switch (below(9)) {
        case 0: if (!b.empty()) b[below(b.size())] = (char)rnd(); break;
        case 1: if (!b.empty()) b[below(b.size())] ^= (char)(1u << below(8)); break;
        case 2: if (!b.empty()) b.resize(below(b.size())); break;
        case 3: b.insert(below(b.size() + 1), 1, (char)rnd()); break;
        case 4: if (!b.empty()) b[below(std::min<size_t>(b.size(), 24))] = (char)edge[below(sizeof edge)]; break;       // headers sit in front
        case 5: { const std::string& o = seeds[below(seeds.size())]; size_t a = below(o.size() + 1), n = below(o.size() - a + 1); b.insert(below(b.size() + 1), o, a, n); break; }
        case 6: if (b.size() > 2) { size_t a = below(b.size()), n = below(std::min<size_t>(b.size() - a, 64) + 1); b.erase(a, n); } break;
        case 7: for (int i = (int)below(6); i >= 0; i--) b.push_back((char)rnd()); break;
        case 8: {   // nesting bomb: a long run of one structural byte (NOT / AND / OR node kinds, parentheses, repetition)
            static const char bomb[] = {8, 6, 7, '(', ')', '[', '{', '*', '|', '\\'};
            const size_t len = below(4) ? below(300) : below(200000);
            b.insert(below(b.size() + 1), len, bomb[below(sizeof bomb)]);
            break;
        }
        }
    }
    return b;
}

// exact-size heap copy of a byte string (AddressSanitizer then sees any read past its end, and before its start)
struct Exact {
    std::unique_ptr<uint8_t[]> p; uint32_t n;
    explicit Exact(const std::string& s) : p(new uint8_t[s.size() ? s.size() : 1]), n((uint32_t)s.size()) { if (n) memcpy(p.get(), s.data(), n); }
};
static std::string random_value() {
    static const char* pieces[] = {"error", "GET", " ", "/api/v1", "10.0.0.1", ".", "-", "_", "0", "255", "2024-05-06T07:08:09.123Z", "\xc3\xa9", "\xd0\x99", "\xe6\x97\xa5", "\xf0\x9f\x99\x82", "\xc4\xb0", "\xc8\xba",
                                   "\xff", "\xe2\x82", "\xf0\x9f", "A", "Z", "timeout", ":", "1e5", "-17", "18446744073709551615", "0.5"};
    std::string v;
    switch (below(4)) {
    case 0: for (int i = (int)below(12); i > 0; i--) v += pieces[below(sizeof pieces / sizeof *pieces)]; break;
    case 1: for (int i = (int)below(24); i > 0; i--) v.push_back((char)rnd()); break;
    case 2: for (int i = (int)below(40); i > 0; i--) v.push_back("ab AB01._-:/"[below(12)]); break;
    default: for (int i = (int)below(6); i > 0; i--) { v += pieces[below(sizeof pieces / sizeof *pieces)]; if (below(3) == 0) v.push_back((char)rnd()); } break;
    }
    return v;
}
static std::string pack_list(const std::vector<std::string>& v) {
    std::string out;
    for (auto& ph : v) { size_t n = ph.size(); while (n >= 0x80) { out.push_back((char)((n & 0x7F) | 0x80)); n >>= 7; } out.push_back((char)n); out += ph; }
    return out;
}
static int fuzz_predicates(long iters) {
    uint64_t acc = 0;
    for (long it = 0; it < iters; it++) {
        const std::string sv = random_value();
        std::string nv = below(3) ? random_value() : sv.substr(below(sv.size() + 1), below(8));
        const std::string bv = random_value();
        Exact s(sv), a(nv), b(bv);
        acc += vl::match_phrase(s.p.get(), s.n, a.p.get(), a.n) + 2 * vl::match_prefix(s.p.get(), s.n, a.p.get(), a.n) + 4 * vl::bytes_equal(s.p.get(), s.n, a.p.get(), a.n);
        for (int kind = 9; kind <= 12; kind++) acc += vl::range_predicate(kind, s.p.get(), s.n, a.p.get(), a.n, b.p.get(), b.n, rnd() % 12, rnd() % 0x100000000ull);
        acc += vl::any_case_match(s.p.get(), s.n, a.p.get(), a.n, false) + vl::any_case_match(s.p.get(), s.n, a.p.get(), a.n, true);
        std::vector<std::string> phrases; for (int i = (int)below(4); i > 0; i--) phrases.push_back(below(2) ? sv.substr(below(sv.size() + 1), below(6)) : random_value().substr(0, below(5)));
        std::string packed = pack_list(phrases);
        if (below(8) == 0 && !packed.empty()) packed[below(packed.size())] = (char)rnd();      // a damaged list must not be read past its end either
        Exact L(packed);
        acc += vl::match_sequence(s.p.get(), s.n, vl::PhraseList{L.p.get(), L.n}) + vl::match_all_phrases(s.p.get(), s.n, vl::PhraseList{L.p.get(), L.n}) + vl::match_any_phrase(s.p.get(), s.n, vl::PhraseList{L.p.get(), L.n});
        acc += vl::rune_count(s.p.get(), s.n) + (uint64_t)vl::bytes_cmp(s.p.get(), s.n, a.p.get(), a.n) + vl::xxh64(s.p.get(), s.n);
        uint64_t u; uint32_t ip; acc += vl::parse_date_u64_hd(s.p.get(), s.n, &u) + vl::parse_ipv4_hd(s.p.get(), s.n, &ip);
        int w; acc += (uint64_t)vl::decode_rune(s.p.get(), s.n, &w) + (uint64_t)vl::decode_last_rune(s.p.get(), s.n, &w);
        // number -> text formatters into buffers of exactly the documented size
        {
            std::unique_ptr<uint8_t[]> f(new uint8_t[VL_FMT_F64_MAX]);
            uint64_t bits = rnd(); if (below(4) == 0) bits &= 0x800FFFFFFFFFFFFFull; if (below(4) == 0) bits |= 0x7FE0000000000000ull;
            const int n = vl::fmt_f64(f.get(), bits); if (n <= 0 || n > VL_FMT_F64_MAX) { fprintf(stderr, "fmt_f64 length %d\n", n); return 3; }
            std::unique_ptr<uint8_t[]> g(new uint8_t[32]);
            acc += (uint64_t)vl::fmt_u64(g.get(), rnd()) + (uint64_t)vl::fmt_i64(g.get(), (int64_t)rnd()) + (uint64_t)vl::fmt_ipv4(g.get(), (uint32_t)rnd());
            const int64_t ns = below(2) ? (int64_t)rnd() : (int64_t)(rnd() % 4102444800ull) * 1000000000ll + (int64_t)(rnd() % 1000000000ull);   // any int64 / years 1970..2100
            acc += (uint64_t)vl::fmt_iso8601(g.get(), ns);
        }
    }
    printf("accepted %ld rejected %ld\n", iters, (long)(acc & 1));
    return 0;
}

static std::string mutate_one(const std::string& in) { std::vector<std::string> one{in}; return mutate(one); }

// part mode: `dir` is copied file by file into `scratch`; every iteration one file of the copy is replaced by a damaged version
static int fuzz_part(const char* dir, long iters) {
    std::vector<std::string> names; std::map<std::string, std::string> files;
    DIR* d = opendir(dir);
    if (!d) { perror(dir); return 2; }
    while (dirent* e = readdir(d)) if (e->d_name[0] != '.') names.push_back(e->d_name);
    closedir(d);
    std::sort(names.begin(), names.end());
    for (auto& n : names) { std::ifstream f(std::string(dir) + "/" + n, std::ios::binary); files[n] = std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>()); }
    char tmpl[] = "/tmp/vl_part_fuzz_XXXXXX";
    const char* scratch = mkdtemp(tmpl);
    if (!scratch) { perror("mkdtemp"); return 2; }
    auto put = [&](const std::string& n, const std::string& data) { std::ofstream f(std::string(scratch) + "/" + n, std::ios::binary | std::ios::trunc); f.write(data.data(), (std::streamsize)data.size()); };
    for (auto& kv : files) put(kv.first, kv.second);
    const vl::part::Inflate inflate = [](const uint8_t* f, size_t n, uint8_t* dst, size_t dn) {
        const size_t got = ZSTD_decompress(dst, dn, f, n);
        if (ZSTD_isError(got) || got != dn) throw vl::BadInput("cannot decompress a metadata frame");
    };
    std::vector<std::string> fields{"_msg", "level", "status", "bytes", "delta", "ratio", "ip", "ts", "host", "sparse", "only_in_1", "nope"};
    long ok = 0, bad = 0;
    for (long it = -1; it < iters; it++) {
        std::string victim;
        if (it >= 0) {
            do victim = names[below(names.size())]; while (below(4) && (victim.rfind("values.bin", 0) == 0 || victim.rfind("bloom.bin", 0) == 0 || victim.rfind("message_", 0) == 0 || victim == "timestamps.bin"));
            put(victim, below(16) ? mutate_one(files[victim]) : std::string());
        }
        try {
            vl::part::PartReader r; r.open(scratch, inflate);
            vl::part::Described dsc; r.describe(fields, 0, r.blockHeaders.size(), INT64_MIN, INT64_MAX, dsc);
            uint64_t sum = 0;   // touch every byte the descriptors point at
            for (const vlscan_column& c : dsc.cols) {
                for (uint64_t i = 0; i < c.const_len; i++) sum += c.const_value[i];
                for (uint64_t i = 0; i < c.values_len; i++) sum += c.values[i];
                for (uint64_t i = 0; i < c.bloom_len; i++) sum += c.bloom[i];
                if (c.dict_len) for (uint32_t i = 0; i < c.dict_offsets[c.dict_len]; i++) sum += c.dict_blob[i];
            }
            if (sum == 0x1234567812345678ull) printf("!");
            ok++;
        } catch (const vl::BadInput&) { bad++; }
        if (it < 0 && bad) { fprintf(stderr, "the undamaged part was rejected\n"); return 4; }
        if (it >= 0) put(victim, files[victim]);
    }
    for (auto& n : names) unlink((std::string(scratch) + "/" + n).c_str());
    rmdir(scratch);
    printf("accepted %ld rejected %ld\n", ok, bad);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: harness tree|zstd|part <seed dir> <iterations> <rng seed>\n"); return 2; }
    const std::string mode = argv[1];
    if (mode == "pred") { rng_state ^= (uint64_t)atoll(argv[4]) * 0x9E3779B97F4A7C15ull; return fuzz_predicates(atol(argv[3])); }
    if (mode == "part") { rng_state ^= (uint64_t)atoll(argv[4]) * 0x9E3779B97F4A7C15ull; return fuzz_part(argv[2], atol(argv[3])); }
    const std::vector<std::string> seeds = read_seeds(argv[2]);
    const long iters = atol(argv[3]);
    rng_state ^= (uint64_t)atoll(argv[4]) * 0x9E3779B97F4A7C15ull;
    long ok = 0, bad = 0;
    for (long it = -(long)seeds.size(); it < iters; it++) {
        const std::string in = it < 0 ? seeds[(size_t)(it + (long)seeds.size())] : mutate(seeds);    // the untouched seeds first
        std::unique_ptr<uint8_t[]> buf(new uint8_t[in.size() ? in.size() : 1]);
        if (!in.empty()) memcpy(buf.get(), in.data(), in.size());
        try {
            if (mode == "tree") {
                vl::Program P; vl::ProgramBuilder B(buf.get(), in.size(), P); B.build();
                size_t toks = 0; for (auto& t : P.leaf_tokens) toks += t.size();
                if (P.root < 0 || P.leaves.size() != P.leaf_tokens.size() || toks > (1u << 26)) { fprintf(stderr, "inconsistent program\n"); return 3; }
            } else {
                // a values block: bytesBlock(lens) ++ bytesBlock(data), nothing behind them (stringsBlockUnmarshaler.unmarshal encoding.go:83-108)
                std::vector<vl::zs::ZBlock> blocks; vl::zs::ZFrame f1{}, f2{};
                const size_t c1 = vl::zwalk::parse_bytes_block_into(blocks, 0, buf.get(), in.size(), 512, f1);
                const size_t c2 = vl::zwalk::parse_bytes_block_into(blocks, 1, buf.get() + c1, in.size() - c1, 512 + c1, f2);
                if (c1 + c2 != in.size()) throw vl::BadInput("unexpected non-empty tail after reading bytes block with strings");
                // what the walk promises the device: every block lies inside the input, frames own consecutive block ranges
                if (f1.blk_lo != 0 || f1.blk_hi != f2.blk_lo || f2.blk_hi != blocks.size() || f1.blk_hi == f1.blk_lo || f2.blk_hi == f2.blk_lo) { fprintf(stderr, "bad block ranges\n"); return 3; }
                for (const vl::zs::ZBlock& b : blocks) {
                    const uint64_t content = b.type == vl::zs::ZB_RLE ? 1 : b.size;
                    if (b.src < 512 || b.src - 512 + content > in.size()) { fprintf(stderr, "block outside the input\n"); return 3; }
                    if (b.type == vl::zs::ZB_COMPRESSED && ((uint64_t)b.lit_hdr + b.lit_comp >= b.size || b.seq_hdr > b.size || b.lit_regen > (128u << 10))) { fprintf(stderr, "section outside its block\n"); return 3; }
                }
            }
            ok++;
        } catch (const vl::ProgError&) { bad++; }
        catch (const vl::BadInput&) { bad++; }
        if (it < 0 && bad) { fprintf(stderr, "seed %ld was rejected\n", it + (long)seeds.size()); return 4; }
    }
    printf("accepted %ld rejected %ld\n", ok, bad);
    return 0;
}
