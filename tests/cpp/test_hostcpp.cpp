// Exercises centrolign_b200/hostcpp/po_poa_b200.hpp with a minimal graph type that models the
// reference's Graph concept, on the golden cases of the reference's own unit test
// (src/test/test_alignment.cpp:684-773).  Needs a GPU to run; `--link-only` just proves that it
// builds and links against libcentrolign_b200.so.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "chain_batcher.hpp"
#include "po_poa_b200.hpp"

struct MiniGraph {  // node_size / label / previous, like centrolign::BaseGraph (graph.hpp:94-149)
    std::string labels;
    std::vector<std::vector<uint64_t>> prev;
    uint64_t add_node(char c) { labels.push_back(c); prev.emplace_back(); return labels.size() - 1; }
    void add_edge(uint64_t a, uint64_t b) { prev[b].push_back(a); }
    size_t node_size() const { return labels.size(); }
    char label(uint64_t v) const { return labels[v]; }
    const std::vector<uint64_t>& previous(uint64_t v) const { return prev[v]; }
};

static MiniGraph bubbles(const char* seq) {
    MiniGraph g;
    for (const char* c = seq; *c; ++c) g.add_node(*c);
    const int e[8][2] = {{0, 1}, {0, 2}, {1, 3}, {2, 3}, {3, 4}, {3, 5}, {4, 6}, {5, 6}};
    for (auto& x : e) g.add_edge(x[0], x[1]);
    return g;
}

static int check(const centrolign_b200::Alignment& got, const std::vector<centrolign_b200::AlignedPair>& want, const char* what) {
    if (got == want) return 0;
    std::fprintf(stderr, "FAILED %s: got", what);
    for (auto& p : got) std::fprintf(stderr, " (%lld,%lld)", (long long)p.node_id1, (long long)p.node_id2);
    std::fprintf(stderr, "\n");
    return 1;
}

// ---- ChainBatcher without a device: jobs are integers in disguise, the stand-in runner squares them ----
namespace batcher_test {
std::mutex mu;
std::vector<size_t> batch_sizes;
std::vector<int> runs_of;   // how often job k was run
std::vector<long> result;   // k * k once job k has run
int fail_on = -1;           // job whose batch reports an error
int fake_run(int, int64_t n, clb_chain_job* const* jobs) {
    std::lock_guard<std::mutex> lk(mu);
    batch_sizes.push_back((size_t)n);
    bool fail = false;
    for (int64_t i = 0; i < n; ++i) {
        const long k = (long)(reinterpret_cast<uintptr_t>(jobs[i]) - 1);
        runs_of[(size_t)k] += 1;
        result[(size_t)k] = k * k;
        fail |= k == fail_on;
    }
    return fail ? CLB_ECUDA : CLB_OK;
}
int run(int n_items, int pool, bool with_failure) {
    batch_sizes.clear();
    runs_of.assign((size_t)n_items, 0);
    result.assign((size_t)n_items, -1);
    fail_on = with_failure ? n_items / 2 : -1;
    std::vector<long> seen((size_t)n_items, -2);
    bool threw = false;
    try {
        centrolign_b200::batched_parallel_for((size_t)n_items, 0, [&](size_t k) {
            if (k % 7 == 3) return;  // a gap without matches: no chaining call at all
            if (k % 5 == 0) std::this_thread::sleep_for(std::chrono::microseconds(200 + 37 * (k % 11)));
            clb_chain_job* job = reinterpret_cast<clb_chain_job*>(uintptr_t(k + 1));
            if (centrolign_b200::ChainBatcher* b = centrolign_b200::ChainBatcher::current()) b->solve(job);
            else if (fake_run(0, 1, &job) != CLB_OK) throw std::runtime_error("serial call failed");  // the serial loop: one call per gap
            seen[k] = result[k];  // the job has run when solve returns
        }, fake_run, pool);
    } catch (const std::runtime_error&) {
        threw = true;
    }
    int bad = 0;
    if (threw != with_failure) { std::fprintf(stderr, "FAILED batcher: exception %d, expected %d\n", (int)threw, (int)with_failure); ++bad; }
    size_t total = 0, largest = 0;
    for (size_t b : batch_sizes) { total += b; largest = std::max(largest, b); }
    for (int k = 0; k < n_items && !with_failure; ++k) {
        const bool skipped = k % 7 == 3;
        if (runs_of[(size_t)k] != (skipped ? 0 : 1) || seen[(size_t)k] != (skipped ? -2 : (long)k * k)) {
            std::fprintf(stderr, "FAILED batcher: job %d ran %d times, worker saw %ld\n", k, runs_of[(size_t)k], seen[(size_t)k]);
            ++bad;
            break;
        }
    }
    if (!with_failure && pool > 1 && largest < 2) { std::fprintf(stderr, "FAILED batcher: no launch was shared (%zu launches)\n", batch_sizes.size()); ++bad; }
    if (largest > (size_t)pool) { std::fprintf(stderr, "FAILED batcher: a batch of %zu from a pool of %d\n", largest, pool); ++bad; }
    std::printf("batcher: %d items, pool %d%s: %zu launches, largest %zu, %zu jobs\n", n_items, pool, with_failure ? ", failing job" : "", batch_sizes.size(),
                largest, total);
    return bad;
}
}  // namespace batcher_test

int main(int argc, char** argv) {
    if (argc > 1 && !std::strcmp(argv[1], "--batcher")) {  // host-only: the rendezvous of the fill-in pool (chain_batcher.hpp)
        int bad = 0;
        bad += batcher_test::run(2000, 32, false);
        bad += batcher_test::run(50, 64, false);   // more workers than items
        bad += batcher_test::run(300, 1, false);   // serial loop: no batcher at all
        bad += batcher_test::run(400, 16, true);   // an error in one launch reaches the caller as an exception
        std::printf(bad ? "FAILED\n" : "batcher passed all tests!\n");
        return bad ? 1 : 0;
    }
    if (argc > 1 && !std::strcmp(argv[1], "--link-only")) {
        std::printf("linked, %d CUDA device(s)\n", clb_device_count());
        return 0;
    }
    using namespace centrolign_b200;
    const uint64_t G = AlignedPair::gap;
    AlignmentParameters<1> p{1, 1, {1}, {1}};
    int bad = 0;
    MiniGraph g1 = bubbles("ACGTGCA"), g2 = bubbles("AGTTTGA");
    bad += check(po_poa<1>(g1, g2, {0}, {0}, {6}, {6}, p), {{0, 0}, {2, 1}, {3, 3}, {4, 5}, {6, 6}}, "double bubble");
    g1.add_node('T'); g1.add_edge(7, 0);
    g2.add_node('T'); g2.add_edge(6, 7);
    int64_t score = 0;
    bad += check(po_poa<1>(g1, g2, {7}, {0}, {6}, {7}, p, &score),
                 {{7, G}, {0, 0}, {2, 1}, {3, 3}, {4, 5}, {6, 6}, {G, 7}}, "lead/trail gaps");
    bad += check(po_poa<1>(g2, g1, {0}, {7}, {7}, {6}, p),
                 {{G, 7}, {0, 0}, {1, 2}, {3, 3}, {5, 4}, {6, 6}, {7, G}}, "flipped");
    // batched form: three windows in one call, production parameters (src/stitcher.cpp:13-22)
    AlignmentParameters<3> prod{20, 80, {60, 800, 2500}, {30, 5, 1}};
    PoPoaBatch batch;
    batch.add(g1, g2, {7}, {0}, {6}, {7});
    batch.add(g2, g1, {0}, {7}, {7}, {6});
    batch.add(g1, g1, {7}, {7}, {6}, {6});
    std::vector<Alignment> alns;
    std::vector<int64_t> scores;
    batch.align<3>(prod, alns, &scores);
    if (alns.size() != 3 || scores[2] != 6 * 20) { std::fprintf(stderr, "FAILED batch: self alignment score %lld\n", (long long)scores[2]); ++bad; }
    // several GPUs of the box: the same windows dealt to every visible device in cell-balanced bins, results in window order
    {
        std::vector<int> devs;
        for (int d = 0; d < clb_device_count(); ++d) devs.push_back(d);
        PoPoaBatch multi(devs);
        for (int rep = 0; rep < 5; ++rep) {
            multi.add(g1, g2, {7}, {0}, {6}, {7});
            multi.add(g2, g1, {0}, {7}, {7}, {6});
            multi.add(g1, g1, {7}, {7}, {6}, {6});
        }
        std::vector<Alignment> malns;
        std::vector<int64_t> mscores;
        multi.align<3>(prod, malns, &mscores);
        for (size_t w = 0; w < malns.size(); ++w)
            if (!(malns[w] == alns[w % 3]) || mscores[w] != scores[w % 3]) { std::fprintf(stderr, "FAILED multi-device batch, window %zu on %zu device(s)\n", w, devs.size()); ++bad; break; }
        std::printf("multi-device batch: %zu windows over %zu device(s)\n", malns.size(), devs.size());
        // a device listed twice is an argument error
        bool threw2 = false;
        if (!devs.empty()) {
            PoPoaBatch twice(std::vector<int>(2, devs[0]));
            twice.add(g1, g2, {7}, {0}, {6}, {7});
            twice.add(g2, g1, {0}, {7}, {7}, {6});
            try { std::vector<Alignment> x; twice.align<3>(prod, x); } catch (const std::runtime_error&) { threw2 = true; }
            if (!threw2) { std::fprintf(stderr, "FAILED: duplicate device did not raise\n"); ++bad; }
        }
    }
    // cyclic input is an error, not a crash (the reference asserts, topological_order.hpp:56)
    MiniGraph cyc = bubbles("ACGTGCA");
    cyc.add_edge(6, 0);
    bool threw = false;
    try { po_poa<1>(cyc, g2, {0}, {0}, {6}, {6}, p); } catch (const std::runtime_error&) { threw = true; }
    if (!threw) { std::fprintf(stderr, "FAILED: cyclic graph did not raise\n"); ++bad; }
    if (!bad) std::printf("passed all tests!\n");
    return bad ? 1 : 0;
}
