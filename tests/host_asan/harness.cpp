// Sanitizer build of the host-side parsers of libvlscan.so that read untrusted bytes:
//   tree  the filter-tree program compiler incl. the regexp compiler (victorialogs_b200/csrc/vl_program.h, vl_regex.h)
//   zstd  the bytes-block / ZSTD header walk                         (victorialogs_b200/csrc/vl_zstd_walk.h)
// usage: harness tree|zstd <seed directory> <iterations> <rng seed>
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
#include "vl_program.h"
#include "vl_zstd_walk.h"

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
        switch (below(8)) {
        case 0: if (!b.empty()) b[below(b.size())] = (char)rnd(); break;
        case 1: if (!b.empty()) b[below(b.size())] ^= (char)(1u << below(8)); break;
        case 2: if (!b.empty()) b.resize(below(b.size())); break;
        case 3: b.insert(below(b.size() + 1), 1, (char)rnd()); break;
        case 4: if (!b.empty()) b[below(std::min<size_t>(b.size(), 24))] = (char)edge[below(sizeof edge)]; break;       // headers sit in front
        case 5: { const std::string& o = seeds[below(seeds.size())]; size_t a = below(o.size() + 1), n = below(o.size() - a + 1); b.insert(below(b.size() + 1), o, a, n); break; }
        case 6: if (b.size() > 2) { size_t a = below(b.size()), n = below(std::min<size_t>(b.size() - a, 64) + 1); b.erase(a, n); } break;
        case 7: for (int i = (int)below(6); i >= 0; i--) b.push_back((char)rnd()); break;
        }
    }
    return b;
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: harness tree|zstd <seed dir> <iterations> <rng seed>\n"); return 2; }
    const std::string mode = argv[1];
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
