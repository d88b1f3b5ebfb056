// ThreadSanitizer harness for csrc/vl_hostpool.h: jobs of varying width back to back (packing a pinned chunk is one job), every index
// of every job runs exactly once, the writes of a job are visible to the caller when run() returns, destruction joins the workers.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "vl_hostpool.h"

int main(int argc, char** argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 2000;
    unsigned long long total = 0;
    for (int rep = 0; rep < 3; rep++) {
        vl::HostPool pool;
        std::vector<int> cells(64, 0);            // plain ints: a missing happens-before edge is a reported race
        unsigned state = 12345u + rep;
        for (int r = 0; r < rounds; r++) {
            state = state * 1664525u + 1013904223u;
            const int n = 1 + (int)((state >> 16) % 33);
            std::atomic<int> ran{0};
            pool.run(n, [&](int t) { cells[t] += t + 1; ran.fetch_add(1, std::memory_order_relaxed); });
            if (ran.load() != n) { printf("bad: job of %d ran %d indices\n", n, ran.load()); return 1; }
            for (int t = 0; t < n; t++) total += (unsigned long long)cells[t];
        }
    }
    printf("ok total=%llu\n", total);
    return 0;
}
