// ThreadSanitizer build of the multi-threaded header walk of an upload (victorialogs_b200/csrc/vl_zstd_job.h): the tables must come out the
// same for every thread count, and no two threads may touch the same bytes without ordering.
// usage: walk_tsan <directory of values blocks> <replication>      (built and run by tests/test_host_asan_cpu.py with -fsanitize=thread)
#include <dirent.h>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>
#include "vl_zstd_job.h"

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: walk_tsan <seed dir> <replication>\n"); return 2; }
    std::vector<std::string> names, seeds;
    DIR* d = opendir(argv[1]);
    if (!d) { perror(argv[1]); return 2; }
    while (dirent* e = readdir(d)) if (e->d_name[0] != '.') names.push_back(e->d_name);
    closedir(d);
    std::sort(names.begin(), names.end());
    for (auto& n : names) { std::ifstream f(std::string(argv[1]) + "/" + n, std::ios::binary); seeds.emplace_back((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>()); }
    if (seeds.empty()) { fprintf(stderr, "no seeds\n"); return 2; }
    const int rep = atoi(argv[2]);
    std::vector<vl::ZValuesBlock> v; uint64_t zoff = 512;
    for (int r = 0; r < rep; r++) for (size_t i = 0; i < seeds.size(); i++) { const std::string& s = seeds[(i * 7 + (size_t)r) % seeds.size()]; v.push_back({(const uint8_t*)s.data(), s.size(), zoff}); zoff += s.size(); }
    uint64_t ref[4] = {0, 0, 0, 0}; size_t ref_groups = 0;
    for (int T : {1, 2, 7, 16, 33}) {
        vl::ZstdJobImpl J; std::vector<vl::ZValuesInfo> info(v.size()); size_t bad = SIZE_MAX; std::string msg;
        J.walk_values_blocks(v.data(), v.size(), T, info.data(), &bad, &msg);
        if (bad != SIZE_MAX) { fprintf(stderr, "threads=%d: block %zu rejected: %s\n", T, bad, msg.c_str()); return 3; }
        J.prepare();
        uint64_t dg[4]; J.digest(dg);
        std::vector<uint8_t> copy(J.blocks.size() * sizeof(vl::zs::ZBlock));
        J.spread_copy(copy.data(), J.blocks.data(), copy.size());
        if (memcmp(copy.data(), J.blocks.data(), copy.size()) != 0) { fprintf(stderr, "threads=%d: spread_copy differs\n", T); return 3; }
        if (T == 1) { memcpy(ref, dg, sizeof ref); ref_groups = J.groups.size(); }
        else if (memcmp(ref, dg, sizeof ref) != 0) { fprintf(stderr, "threads=%d: digest differs from the single-threaded walk\n", T); return 3; }
        printf("threads=%d frames=%zu blocks=%zu groups=%zu\n", T, J.frames.size(), J.blocks.size(), J.groups.size());
    }
    // a damaged block: every thread count reports the same (first) one
    std::string broken = seeds[0]; broken.resize(broken.size() - 1);
    size_t where[3] = {v.size() / 3, v.size() / 2, v.size() - 2};
    std::vector<vl::ZValuesBlock> w = v;
    for (size_t k : where) { w[k].p = (const uint8_t*)broken.data(); w[k].n = broken.size(); }
    std::string first_msg;
    for (int T : {1, 5, 16}) {
        vl::ZstdJobImpl J; std::vector<vl::ZValuesInfo> info(w.size()); size_t bad = SIZE_MAX; std::string msg;
        J.walk_values_blocks(w.data(), w.size(), T, info.data(), &bad, &msg);
        if (bad != where[0]) { fprintf(stderr, "threads=%d: reported block %zu, want %zu\n", T, bad, where[0]); return 3; }
        if (T == 1) first_msg = msg; else if (msg != first_msg) { fprintf(stderr, "threads=%d: another message\n", T); return 3; }
    }
    printf("ok groups=%zu\n", ref_groups);
    return 0;
}
