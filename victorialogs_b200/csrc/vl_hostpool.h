// Persistent host threads of a ctx (CUDA-free; tests/host_asan/pool_tsan.cpp runs it under ThreadSanitizer).
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace vl {

// Persistent host threads of a ctx for packing pageable memory into the pinned staging ring.  Spawning threads per 64 MB chunk put ~16 clone()
// calls of a process with a CUDA-sized address space in front of every chunk; the workers here are started once and woken per job.
struct HostPool {
    std::vector<std::thread> workers;
    std::mutex mu; std::condition_variable cv, cv_done;
    const std::function<void(int)>* job = nullptr;
    int active = 0, pending = 0; uint64_t gen = 0; bool stop = false;
    void worker(int idx) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* f;
            {
                std::unique_lock<std::mutex> g(mu);
                cv.wait(g, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                if (idx >= active) continue;
                f = job;
            }
            (*f)(idx);
            std::lock_guard<std::mutex> g(mu);
            if (--pending == 0) cv_done.notify_one();
        }
    }
    // f(0) .. f(n - 1); f(0) runs on the calling thread.  One job at a time (a ctx belongs to one search worker).
    void run(int n, const std::function<void(int)>& f) {
        if (n <= 1) { f(0); return; }
        while ((int)workers.size() < n - 1) { const int idx = (int)workers.size() + 1; workers.emplace_back([this, idx] { worker(idx); }); }
        { std::lock_guard<std::mutex> g(mu); job = &f; active = n; pending = n - 1; gen++; }
        cv.notify_all();
        f(0);
        std::unique_lock<std::mutex> g(mu);
        cv_done.wait(g, [&] { return pending == 0; });
        job = nullptr;
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> g(mu); stop = true; }
        cv.notify_all();
        for (auto& t : workers) t.join();
    }
};

}  // namespace vl
