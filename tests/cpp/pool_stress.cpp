// Stress test of the engine's WorkerPool (meilisearch_b200/csrc/engine.h): every index of every parallel-for must run exactly once,
// also when jobs follow each other faster than late-waking workers notice (the announce / drain / publish protocol).
#include "../../meilisearch_b200/csrc/engine.h"

#include <cstdio>
using namespace b200;

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    WorkerPool pool(7);
    std::vector<std::atomic<int>> hits(4096);
    size_t bad = 0;
    for (int it = 0; it < iters; it++) {
        size_t n = 4 + (size_t)(it * 7919) % 3000;
        for (size_t i = 0; i < n; i++) hits[i].store(0, std::memory_order_relaxed);
        pool.run(n, [&](size_t i) { hits[i].fetch_add(1, std::memory_order_relaxed); });
        for (size_t i = 0; i < n; i++)
            if (hits[i].load() != 1) bad++;
        if ((it % 5000) == 4999) std::this_thread::sleep_for(std::chrono::milliseconds(2));  // let the workers fall asleep
    }
    printf("bad=%zu\n", bad);
    return bad != 0;
}
