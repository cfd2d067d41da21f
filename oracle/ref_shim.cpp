/*
 * oracle/ref_shim.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * extern "C" doorway onto the UNMODIFIED reference implementation, compiled from the
 * reference's own sources where they lie (see oracle/Makefile; nothing is copied into
 * this repository).  It exposes the same flat-array signature as the C oracle
 * (oracle/po_poa_oracle.c) so the two can be compared call for call, and so that
 * bench.py can time the reference's CPU path (cpu_baseline.kind = "reference").
 *
 * Wrapped reference entry points:
 *   centrolign::po_poa<NumPW, BaseGraph>      include/centrolign/alignment.hpp:78-85
 *   centrolign::pwfa_po_poa<NumPW, BaseGraph> include/centrolign/alignment.hpp:117-125
 */
#include "centrolign/alignment.hpp"
#include "centrolign/graph.hpp"

#include <cstdint>
#include <vector>

using namespace centrolign;

namespace {

BaseGraph build(uint32_t n, const uint8_t* label, const uint32_t* pred_off, const uint32_t* pred) {
    BaseGraph g;
    for (uint32_t v = 0; v < n; ++v) g.add_node((char)label[v]);
    // adding edges grouped by head node, in list order, reproduces previous() order
    for (uint32_t v = 0; v < n; ++v)
        for (uint32_t k = pred_off[v]; k < pred_off[v + 1]; ++k) g.add_edge(pred[k], v);
    return g;
}

std::vector<uint64_t> widen(uint32_t n, const uint32_t* v) { return std::vector<uint64_t>(v, v + n); }

template <int P>
AlignmentParameters<P> unpack(const uint32_t* params) {
    AlignmentParameters<P> p;
    p.match = params[0];
    p.mismatch = params[1];
    for (int k = 0; k < P; ++k) {
        p.gap_open[k] = params[2 + k];
        p.gap_extend[k] = params[5 + k];
    }
    return p;
}

void emit(const Alignment& aln, int32_t* aln_out, uint32_t* aln_len) {
    uint32_t len = 0;
    for (const auto& ap : aln) {
        aln_out[2 * len] = ap.node_id1 == AlignedPair::gap ? -1 : (int32_t)ap.node_id1;
        aln_out[2 * len + 1] = ap.node_id2 == AlignedPair::gap ? -1 : (int32_t)ap.node_id2;
        ++len;
    }
    *aln_len = len;
}

}  // namespace

extern "C" int clref_po_poa(int P, const uint32_t* params,
                            uint32_t n1, const uint8_t* label1, const uint32_t* pred_off1, const uint32_t* pred1,
                            uint32_t nsrc1, const uint32_t* src1, uint32_t nsnk1, const uint32_t* snk1,
                            uint32_t n2, const uint8_t* label2, const uint32_t* pred_off2, const uint32_t* pred2,
                            uint32_t nsrc2, const uint32_t* src2, uint32_t nsnk2, const uint32_t* snk2,
                            int64_t* score_out, int32_t* aln_out, uint32_t* aln_len) {
    BaseGraph g1 = build(n1, label1, pred_off1, pred1);
    BaseGraph g2 = build(n2, label2, pred_off2, pred2);
    auto s1 = widen(nsrc1, src1), s2 = widen(nsrc2, src2), k1 = widen(nsnk1, snk1), k2 = widen(nsnk2, snk2);
    Alignment aln;
    int64_t score = 0;
    switch (P) {
        case 1: aln = po_poa(g1, g2, s1, s2, k1, k2, unpack<1>(params), &score); break;
        case 2: aln = po_poa(g1, g2, s1, s2, k1, k2, unpack<2>(params), &score); break;
        case 3: aln = po_poa(g1, g2, s1, s2, k1, k2, unpack<3>(params), &score); break;
        default: return -3;
    }
    if (score_out) *score_out = score;
    emit(aln, aln_out, aln_len);
    return 0;
}

extern "C" int clref_pwfa_po_poa(int P, const uint32_t* params, int64_t prune_limit,
                                 uint32_t n1, const uint8_t* label1, const uint32_t* pred_off1, const uint32_t* pred1,
                                 uint32_t nsrc1, const uint32_t* src1, uint32_t nsnk1, const uint32_t* snk1,
                                 uint32_t n2, const uint8_t* label2, const uint32_t* pred_off2, const uint32_t* pred2,
                                 uint32_t nsrc2, const uint32_t* src2, uint32_t nsnk2, const uint32_t* snk2,
                                 int64_t* score_out, int32_t* aln_out, uint32_t* aln_len) {
    BaseGraph g1 = build(n1, label1, pred_off1, pred1);
    BaseGraph g2 = build(n2, label2, pred_off2, pred2);
    auto s1 = widen(nsrc1, src1), s2 = widen(nsrc2, src2), k1 = widen(nsnk1, snk1), k2 = widen(nsnk2, snk2);
    Alignment aln;
    int64_t score = 0;
    switch (P) {
        case 1: aln = pwfa_po_poa(g1, g2, s1, s2, k1, k2, unpack<1>(params), prune_limit, &score); break;
        case 2: aln = pwfa_po_poa(g1, g2, s1, s2, k1, k2, unpack<2>(params), prune_limit, &score); break;
        case 3: aln = pwfa_po_poa(g1, g2, s1, s2, k1, k2, unpack<3>(params), prune_limit, &score); break;
        default: return -3;
    }
    if (score_out) *score_out = score;
    emit(aln, aln_out, aln_len);
    return 0;
}

// Same call with SUCCESSOR lists in next() order: edges are added grouped by tail node, in list order, which
// reproduces exactly that next() order (pwfa's enumeration order, alignment.hpp:1788-1826, depends on it).
extern "C" int clref_pwfa_po_poa_succ(int P, const uint32_t* params, int64_t prune_limit,
                                      uint32_t n1, const uint8_t* label1, const uint32_t* next_off1, const uint32_t* next1,
                                      uint32_t nsrc1, const uint32_t* src1, uint32_t nsnk1, const uint32_t* snk1,
                                      uint32_t n2, const uint8_t* label2, const uint32_t* next_off2, const uint32_t* next2,
                                      uint32_t nsrc2, const uint32_t* src2, uint32_t nsnk2, const uint32_t* snk2,
                                      int64_t* score_out, int32_t* aln_out, uint32_t* aln_len, int64_t* /*stats_out*/) {
    BaseGraph g1, g2;
    for (uint32_t v = 0; v < n1; ++v) g1.add_node((char)label1[v]);
    for (uint32_t v = 0; v < n2; ++v) g2.add_node((char)label2[v]);
    for (uint32_t v = 0; v < n1; ++v)
        for (uint32_t k = next_off1[v]; k < next_off1[v + 1]; ++k) g1.add_edge(v, next1[k]);
    for (uint32_t v = 0; v < n2; ++v)
        for (uint32_t k = next_off2[v]; k < next_off2[v + 1]; ++k) g2.add_edge(v, next2[k]);
    auto s1 = widen(nsrc1, src1), s2 = widen(nsrc2, src2), k1 = widen(nsnk1, snk1), k2 = widen(nsnk2, snk2);
    Alignment aln;
    int64_t score = 0;
    switch (P) {
        case 1: aln = pwfa_po_poa(g1, g2, s1, s2, k1, k2, unpack<1>(params), prune_limit, &score); break;
        case 2: aln = pwfa_po_poa(g1, g2, s1, s2, k1, k2, unpack<2>(params), prune_limit, &score); break;
        case 3: aln = pwfa_po_poa(g1, g2, s1, s2, k1, k2, unpack<3>(params), prune_limit, &score); break;
        default: return -3;
    }
    if (score_out) *score_out = score;
    emit(aln, aln_out, aln_len);
    return 0;
}
