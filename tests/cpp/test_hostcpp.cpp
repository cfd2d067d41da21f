// Exercises centrolign_b200/hostcpp/po_poa_b200.hpp with a minimal graph type that models the
// reference's Graph concept, on the golden cases of the reference's own unit test
// (src/test/test_alignment.cpp:684-773).  Needs a GPU to run; `--link-only` just proves that it
// builds and links against libcentrolign_b200.so.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

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

int main(int argc, char** argv) {
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
