// Drop-in check with the reference's OWN types: centrolign::BaseGraph, centrolign::AlignmentParameters
// and centrolign::Alignment go straight into centrolign_b200::po_poa, and the result must equal
// centrolign::po_poa (include/centrolign/alignment.hpp:78-85) pair for pair.  Built in the container
// where /root/reference exists (oracle/Makefile -> oracle/_ref/test_dropin), run on the GPU box.
#include <cstdio>
#include <random>
#include <string>
#include <vector>

#include "centrolign/alignment.hpp"
#include "centrolign/graph.hpp"
#include "po_poa_b200.hpp"

using namespace centrolign;

static BaseGraph random_bubbly(std::mt19937& gen, int len, std::vector<uint64_t>& src, std::vector<uint64_t>& snk) {
    BaseGraph g;
    std::uniform_int_distribution<int> base(0, 3);
    std::uniform_real_distribution<double> u(0, 1);
    const char* alpha = "ACGT";
    for (int i = 0; i < len; ++i) g.add_node(alpha[base(gen) & 1]);
    for (int i = 1; i < len; ++i) g.add_edge(i - 1, i);
    for (int p = 1; p + 1 < len; ++p) {
        if (u(gen) < 0.15) {  // SNP bubble
            uint64_t a = g.add_node(alpha[base(gen)]);
            g.add_edge(p - 1, a);
            g.add_edge(a, p + 1);
        }
        if (u(gen) < 0.04 && p + 3 < len) g.add_edge(p, p + 2 + base(gen) % 2);  // deletion edge
    }
    src.clear(); snk.clear();
    for (uint64_t v = 0; v < g.node_size(); ++v) {
        if (g.previous_size(v) == 0) src.push_back(v);
        if (g.next_size(v) == 0) snk.push_back(v);
    }
    return g;
}

template <int P>
static int trial(std::mt19937& gen, const AlignmentParameters<P>& params, int len1, int len2) {
    std::vector<uint64_t> s1, k1, s2, k2;
    BaseGraph g1 = random_bubbly(gen, len1, s1, k1), g2 = random_bubbly(gen, len2, s2, k2);
    int64_t sr = 0, sg = 0;
    Alignment ref = po_poa(g1, g2, s1, s2, k1, k2, params, &sr);
    Alignment got = centrolign_b200::po_poa<P, BaseGraph, AlignmentParameters<P>, Alignment>(g1, g2, s1, s2, k1, k2, params, &sg);
    if (sr != sg || !(ref == got)) {
        std::fprintf(stderr, "MISMATCH P=%d len %d x %d: score %lld vs %lld, %zu vs %zu pairs\n", P, len1, len2,
                     (long long)sr, (long long)sg, ref.size(), got.size());
        return 1;
    }
    return 0;
}

// the wavefront variant: centrolign_b200::pwfa_po_poa against centrolign::pwfa_po_poa (alignment.hpp:117-125)
template <int P>
static int trial_pwfa(std::mt19937& gen, const AlignmentParameters<P>& params, int len1, int len2, int64_t prune_limit) {
    std::vector<uint64_t> s1, k1, s2, k2;
    BaseGraph g1 = random_bubbly(gen, len1, s1, k1), g2 = random_bubbly(gen, len2, s2, k2);
    int64_t sr = 0, sg = 0;
    Alignment ref = pwfa_po_poa(g1, g2, s1, s2, k1, k2, params, prune_limit, &sr);
    Alignment got = centrolign_b200::pwfa_po_poa<P, BaseGraph, AlignmentParameters<P>, Alignment>(g1, g2, s1, s2, k1, k2, params,
                                                                                                 prune_limit, &sg);
    if (sr != sg || !(ref == got)) {
        std::fprintf(stderr, "PWFA MISMATCH P=%d len %d x %d prune %lld: score %lld vs %lld, %zu vs %zu pairs\n", P, len1, len2,
                     (long long)prune_limit, (long long)sr, (long long)sg, ref.size(), got.size());
        return 1;
    }
    return 0;
}

int main() {
    std::mt19937 gen(20261017);
    AlignmentParameters<3> prod;  // src/stitcher.cpp:13-22
    prod.match = 20; prod.mismatch = 80;
    prod.gap_open = {60, 800, 2500};
    prod.gap_extend = {30, 5, 1};
    int bad = 0;
    std::uniform_int_distribution<int> len(1, 400);
    for (int t = 0; t < 60; ++t) {
        bad += trial<3>(gen, prod, len(gen), len(gen));
        bad += trial<2>(gen, truncate_parameters<3, 2>(prod), len(gen) % 90 + 1, len(gen) % 90 + 1);
        bad += trial<1>(gen, truncate_parameters<3, 1>(prod), len(gen) % 40 + 1, len(gen) % 40 + 1);
    }
    for (int t = 0; t < 40; ++t) {
        const int l = len(gen);
        bad += trial_pwfa<3>(gen, prod, l, l + t % 7, 50);  // Stitcher: 2 * wfa_pruning_dist (stitcher.hpp:339)
        bad += trial_pwfa<2>(gen, truncate_parameters<3, 2>(prod), l % 90 + 1, l % 90 + 2, 6);
        bad += trial_pwfa<1>(gen, truncate_parameters<3, 1>(prod), l % 40 + 1, l % 40 + 1, 1000000);
    }
    if (!bad) std::printf("passed all tests!\n");
    return bad ? 1 : 0;
}
