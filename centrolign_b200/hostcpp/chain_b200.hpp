// chain_b200.hpp -- C++ host layer of the sparse anchor-chaining DP (clb_chain_dp, include/centrolign_b200.h).
//
// The reference runs its chaining DP inside two member templates of Anchorer,
//
//     sparse_affine_chain_dp<UIntSet, UIntMatch, UIntDist, IntShift, UIntAnchor, ScoreFloat, ..., MBank, FwdEdges>
//         (match_sets, graph1, graph2, xmerge1, xmerge2, gap_open, gap_extend, local_scale, num_match_sets,
//          suppress_verbose_logging, sources1, sources2, sinks1, sinks2, masked_matches)
//                                                    reference: include/centrolign/anchorer.hpp:1812-2471
//     sparse_chain_dp<UIntDist, UIntSet, UIntMatch, UIntAnchor, ScoreFloat, ..., MBank, FwdEdges>
//         (match_sets, graph1, chain_merge1, chain_merge2, num_match_sets, ...)
//                                                    reference: include/centrolign/anchorer.hpp:1511-1750
//
// both called from Anchorer::anchor_chain (anchorer.hpp:1213-1307).  The functions below take the objects
// those templates build before their main loop -- the MatchBank, the ForwardEdges, the two
// PostSwitchDistances, the XMerge reachability structures, the topological order of graph 1 -- read them
// through their public interfaces, and write the flat clb_chain_problem the CUDA library consumes: per
// match the tree keys of its end point for every path pair and the query shifts / offsets of its start
// point for every path, and per graph-1 node the matches that end there and the (match, chain) queries of
// its forward edges, all in the reference's own iteration order.  The gap-measuring lambdas of the
// reference (anchorer.hpp:1875-2000) are restated here because the lead / final gap terms and the anchor
// annotations need them.  No DP happens on the host; errors surface as std::runtime_error.
#ifndef CENTROLIGN_B200_CHAIN_HPP
#define CENTROLIGN_B200_CHAIN_HPP

#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

#include "centrolign_b200.h"

namespace centrolign_b200 {

// Owns the arrays a clb_chain_problem points into.
struct ChainProblem {
    int num_pw = 0;
    double gap_open[CLB_MAX_PW] = {0, 0, 0}, gap_extend[CLB_MAX_PW] = {0, 0, 0}, scale = 1.0;
    int32_t n_chain1 = 0, n_chain2 = 0;
    float min_score = 0.0f;
    std::vector<float> weight, dp_init, final_term;
    std::vector<int64_t> end_off{0}, qry_off{0}, ins_off{0};
    std::vector<uint32_t> end_match, qry_match, qry_chain1, ins_p1, ins_p2, ins_offset, qoff;
    std::vector<int32_t> ins_shift, qa1, qa2;
    std::vector<uint8_t> ins_active;  // empty = all active
    std::vector<std::tuple<size_t, size_t, size_t>> ids;  // rank -> (match set, walk1 index, walk2 index)

    size_t n_match() const { return weight.size(); }

    clb_chain_problem view() const {
        clb_chain_problem p;
        p.num_pw = num_pw;
        for (int k = 0; k < CLB_MAX_PW; ++k) {
            p.gap_open[k] = gap_open[k];
            p.gap_extend[k] = gap_extend[k];
        }
        p.scale = scale;
        p.n_chain1 = n_chain1;
        p.n_chain2 = n_chain2;
        p.n_match = (int64_t)weight.size();
        p.weight = weight.data();
        p.dp_init = dp_init.data();
        p.final_term = final_term.data();
        p.min_score = min_score;
        p.n_step = (int64_t)end_off.size() - 1;
        p.end_off = end_off.data();
        p.end_match = end_match.data();
        p.qry_off = qry_off.data();
        p.qry_match = qry_match.data();
        p.qry_chain1 = qry_chain1.data();
        p.ins_off = ins_off.data();
        p.ins_p1 = ins_p1.data();
        p.ins_p2 = ins_p2.data();
        p.ins_shift = ins_shift.data();
        p.ins_offset = ins_offset.data();
        p.ins_active = ins_active.empty() ? nullptr : ins_active.data();
        p.qa1 = qa1.data();
        p.qa2 = qa2.data();
        p.qoff = qoff.data();
        return p;
    }

    // Runs the DP + traceback on the GPU; returns the chain as match ranks in forward order.
    std::vector<int64_t> solve(int device = 0, float* opt_score = nullptr, clb_chain_stats* stats = nullptr,
                               std::vector<float>* dp_out = nullptr, std::vector<int64_t>* backptr_out = nullptr) const {
        const clb_chain_problem p = view();
        std::vector<int64_t> chain(weight.size() + 1);
        int64_t len = 0;
        if (dp_out) dp_out->assign(weight.size(), 0.0f);
        if (backptr_out) backptr_out->assign(weight.size(), -1);
        const int rc = clb_chain_dp(device, &p, dp_out ? dp_out->data() : nullptr, backptr_out ? backptr_out->data() : nullptr,
                                    chain.data(), &len, opt_score, stats);
        if (rc != CLB_OK) throw std::runtime_error(std::string("centrolign_b200: ") + clb_last_error());
        chain.resize((size_t)len);
        return chain;
    }
};

namespace detail {

// rank of every (set, i1, i2) in MatchBank iteration order (match_bank.hpp:188-267), UINT32_MAX if masked
template <class MBank, class MatchSets>
struct RankTable {
    std::vector<size_t> base;     // per set
    std::vector<uint32_t> rank;   // base[set] + i1 * walks2.size() + i2
    const MatchSets* sets;
    RankTable(const MBank& bank, const MatchSets& match_sets, size_t num_match_sets, ChainProblem& P) : sets(&match_sets) {
        base.assign(num_match_sets + 1, 0);
        for (size_t s = 0; s < num_match_sets; ++s)
            base[s + 1] = base[s] + match_sets[s].walks1.size() * match_sets[s].walks2.size();
        rank.assign(base[num_match_sets], 0xffffffffu);
        uint32_t r = 0;
        for (auto it = bank.begin(), end = bank.end(); it != end; ++it) {
            const auto idx = bank.get_match_indexes(*it);
            rank[at(std::get<0>(idx), std::get<1>(idx), std::get<2>(idx))] = r++;
            P.ids.emplace_back(std::get<0>(idx), std::get<1>(idx), std::get<2>(idx));
        }
    }
    size_t at(size_t s, size_t i1, size_t i2) const { return base[s] + i1 * (*sets)[s].walks2.size() + i2; }
    template <class MatchId>
    uint32_t operator()(const MBank& bank, const MatchId& id) const {
        const auto idx = bank.get_match_indexes(id);
        return rank[at(std::get<0>(idx), std::get<1>(idx), std::get<2>(idx))];
    }
};

// events of the main loop, in the reference's order (anchorer.hpp:2290-2416 / :1640-1727)
template <class MBank, class FwdEdges, class BGraph, class Ranks, class TopoOrder>
void collect_events(ChainProblem& P, const MBank& bank, const FwdEdges& forward_edges, const BGraph&, const Ranks& ranks,
                    const TopoOrder& order) {
    for (uint64_t node_id : order) {
        const size_t e0 = P.end_match.size(), q0 = P.qry_match.size();
        for (const auto& match_id : bank.ends_on(node_id)) P.end_match.push_back(ranks(bank, match_id));
        for (auto edge : forward_edges.edges(node_id)) {
            const uint64_t fwd_id = edge.first;
            const uint64_t chain1 = edge.second;
            for (const auto& match_id : bank.starts_on(fwd_id)) {
                P.qry_match.push_back(ranks(bank, match_id));
                P.qry_chain1.push_back((uint32_t)chain1);
            }
        }
        if (P.end_match.size() != e0 || P.qry_match.size() != q0) {
            P.end_off.push_back((int64_t)P.end_match.size());
            P.qry_off.push_back((int64_t)P.qry_match.size());
        }
    }
}

}  // namespace detail

// The gap measures of sparse_affine_chain_dp (anchorer.hpp:1875-2000), same integer and float types.
template <typename IntShift, typename ScoreFloat, class XMerge, class SwitchDists, size_t NumPW>
struct GapMeasure {
    const XMerge& xmerge1;
    const XMerge& xmerge2;
    const SwitchDists& switch_dists1;
    const SwitchDists& switch_dists2;
    const std::array<double, NumPW>& gap_open;
    const std::array<double, NumPW>& gap_extend;
    double local_scale;

    IntShift basic_source_shift(uint64_t src_id1, uint64_t src_id2, uint64_t path1, uint64_t path2) const {  // :1875-1877
        return xmerge1.index_on(src_id1, path1) - xmerge2.index_on(src_id2, path2);
    }
    IntShift basic_query_shift(uint64_t query_id1, uint64_t query_id2, uint64_t path1, uint64_t path2) const {  // :1886-1889
        return (xmerge1.predecessor_index(query_id1, path1) - xmerge2.predecessor_index(query_id2, path2) +
                switch_dists1.distance(query_id1, path1) - switch_dists2.distance(query_id2, path2));
    }
    ScoreFloat score_gap(IntShift gap) const {  // :1906-1918
        ScoreFloat score = std::numeric_limits<ScoreFloat>::lowest();
        if (gap == 0) {
            score = 0.0;
        } else if (gap != std::numeric_limits<IntShift>::max()) {
            for (size_t pw = 0; pw < NumPW; ++pw)
                score = std::max<ScoreFloat>(score, -local_scale * (gap_open[pw] + gap_extend[pw] * std::abs(gap)));
        }
        return score;
    }
    IntShift measure_gap(uint64_t prev_id1, uint64_t prev_id2, uint64_t curr_id1, uint64_t curr_id2) const {  // :1919-1936
        IntShift gap = std::numeric_limits<IntShift>::max();
        if ((prev_id1 == curr_id1 || xmerge1.reachable(prev_id1, curr_id1)) &&
            (prev_id2 == curr_id2 || xmerge2.reachable(prev_id2, curr_id2))) {
            for (auto p1 : xmerge1.chains_on(prev_id1)) {
                for (auto p2 : xmerge2.chains_on(prev_id2)) {
                    IntShift gap_here = basic_source_shift(prev_id1, prev_id2, p1, p2) - basic_query_shift(curr_id1, curr_id2, p1, p2);
                    if (std::abs(gap_here) < std::abs(gap)) gap = gap_here;
                }
            }
        }
        return gap;
    }
    std::pair<IntShift, ScoreFloat> measure_gap_nn(uint64_t p1, uint64_t p2, uint64_t c1, uint64_t c2) const {  // :1937-1943
        std::pair<IntShift, ScoreFloat> r;
        r.first = measure_gap(p1, p2, c1, c2);
        r.second = score_gap(r.first);
        return r;
    }
    // the set variants compare |gap_here| with the signed running value, as the reference does (:1954, :1971, :1991)
    std::pair<IntShift, ScoreFloat> measure_gap_sn(const std::vector<uint64_t>& prev1, const std::vector<uint64_t>& prev2,
                                                   uint64_t curr_id1, uint64_t curr_id2) const {  // :1946-1961
        std::pair<IntShift, ScoreFloat> r(std::numeric_limits<IntShift>::max(), std::numeric_limits<ScoreFloat>::lowest());
        for (uint64_t prev_id1 : prev1)
            for (uint64_t prev_id2 : prev2) {
                IntShift gap_here = measure_gap(prev_id1, prev_id2, curr_id1, curr_id2);
                if (std::abs(gap_here) < r.first) r.first = gap_here;
            }
        r.second = score_gap(r.first);
        return r;
    }
    std::pair<IntShift, ScoreFloat> measure_gap_ns(uint64_t prev_id1, uint64_t prev_id2, const std::vector<uint64_t>& curr1,
                                                   const std::vector<uint64_t>& curr2) const {  // :1963-1978
        std::pair<IntShift, ScoreFloat> r(std::numeric_limits<IntShift>::max(), std::numeric_limits<ScoreFloat>::lowest());
        for (uint64_t curr_id1 : curr1)
            for (uint64_t curr_id2 : curr2) {
                IntShift gap_here = measure_gap(prev_id1, prev_id2, curr_id1, curr_id2);
                if (std::abs(gap_here) < r.first) r.first = gap_here;
            }
        r.second = score_gap(r.first);
        return r;
    }
    std::pair<IntShift, ScoreFloat> measure_gap_ss(const std::vector<uint64_t>& prev1, const std::vector<uint64_t>& prev2,
                                                   const std::vector<uint64_t>& curr1, const std::vector<uint64_t>& curr2) const {  // :1980-2000
        std::pair<IntShift, ScoreFloat> r(std::numeric_limits<IntShift>::max(), std::numeric_limits<ScoreFloat>::lowest());
        for (uint64_t curr_id1 : curr1)
            for (uint64_t curr_id2 : curr2)
                for (uint64_t prev_id1 : prev1)
                    for (uint64_t prev_id2 : prev2) {
                        IntShift gap_here = measure_gap(prev_id1, prev_id2, curr_id1, curr_id2);
                        if (std::abs(gap_here) < r.first) r.first = gap_here;
                    }
        r.second = score_gap(r.first);
        return r;
    }
};

// Flat problem of sparse_affine_chain_dp.  `weight_of(match_set)` is the reference's
// score_function->anchor_weight(count1, count2, walks1.front().size(), full_length) (anchorer.hpp:2023-2024).
template <typename IntShift, class MBank, class FwdEdges, class SwitchDists, class BGraph, class XMerge, class MatchSets,
          class TopoOrder, class WeightFn, size_t NumPW>
ChainProblem build_affine_chain_problem(const MBank& bank, const FwdEdges& forward_edges, const SwitchDists& switch_dists1,
                                        const SwitchDists& switch_dists2, const BGraph& graph1, const TopoOrder& order1,
                                        const XMerge& xmerge1, const XMerge& xmerge2, const MatchSets& match_sets,
                                        size_t num_match_sets, const std::array<double, NumPW>& gap_open,
                                        const std::array<double, NumPW>& gap_extend, double local_scale,
                                        const std::vector<uint64_t>* sources1, const std::vector<uint64_t>* sources2,
                                        const std::vector<uint64_t>* sinks1, const std::vector<uint64_t>* sinks2,
                                        const WeightFn& weight_of) {
    static_assert(NumPW >= 1 && NumPW <= CLB_MAX_PW, "1..3 gap pieces");
    typedef float ScoreFloat;  // the reference instantiates ScoreFloat = float (anchorer.hpp:1217)
    const ScoreFloat mininf = std::numeric_limits<ScoreFloat>::lowest();
    GapMeasure<IntShift, ScoreFloat, XMerge, SwitchDists, NumPW> gaps{xmerge1, xmerge2, switch_dists1, switch_dists2,
                                                                     gap_open, gap_extend, local_scale};
    ChainProblem P;
    P.num_pw = (int)NumPW;
    for (size_t k = 0; k < NumPW; ++k) {
        P.gap_open[k] = gap_open[k];
        P.gap_extend[k] = gap_extend[k];
    }
    P.scale = local_scale;
    P.n_chain1 = (int32_t)xmerge1.chain_size();
    P.n_chain2 = (int32_t)xmerge2.chain_size();
    detail::RankTable<MBank, MatchSets> ranks(bank, match_sets, num_match_sets, P);
    const size_t C1 = xmerge1.chain_size(), C2 = xmerge2.chain_size();

    for (auto it = bank.begin(), end = bank.end(); it != end; ++it) {
        const auto& match_set = bank.match_set(*it);
        const uint64_t start1 = bank.walk1(*it).front(), start2 = bank.walk2(*it).front();
        const uint64_t end1 = bank.walk1(*it).back(), end2 = bank.walk2(*it).back();
        ScoreFloat weight = weight_of(match_set);
        P.weight.push_back(weight);
        if (sources1) {  // anchorer.hpp:2026-2039
            ScoreFloat lead_indel_score = gaps.measure_gap_sn(*sources1, *sources2, start1, start2).second;
            if (lead_indel_score == mininf) weight = mininf;
            else weight += lead_indel_score;
        }
        P.dp_init.push_back(weight);
        P.final_term.push_back(sinks1 ? gaps.measure_gap_ns(end1, end2, *sinks1, *sinks2).second : ScoreFloat(0.0));  // :2431-2438
        for (auto p1 : xmerge1.chains_on(end1))  // anchorer.hpp:2043-2048, 2309-2318
            for (auto p2 : xmerge2.chains_on(end2)) {
                P.ins_p1.push_back((uint32_t)p1);
                P.ins_p2.push_back((uint32_t)p2);
                P.ins_shift.push_back((int32_t)gaps.basic_source_shift(end1, end2, p1, p2));
                P.ins_offset.push_back((uint32_t)xmerge2.index_on(end2, p2));
            }
        P.ins_off.push_back((int64_t)P.ins_p1.size());
        // query shift (anchorer.hpp:1886-1892) is separable in wrapping arithmetic: (pred1 + D1) - (pred2 + D2)
        for (size_t c1 = 0; c1 < C1; ++c1)
            P.qa1.push_back((int32_t)(uint32_t)(uint64_t)((uint64_t)xmerge1.predecessor_index(start1, c1) +
                                                          (uint64_t)switch_dists1.distance(start1, c1)));
        for (size_t c2 = 0; c2 < C2; ++c2) {
            P.qa2.push_back((int32_t)(uint32_t)(uint64_t)((uint64_t)xmerge2.predecessor_index(start2, c2) +
                                                          (uint64_t)switch_dists2.distance(start2, c2)));
            // "we rely on -1's overflowing to 0" (anchorer.hpp:1898-1901); the sentinel has the XMerge's own integer width
            auto pred = xmerge2.predecessor_index(start2, c2);
            P.qoff.push_back((uint32_t)(decltype(pred))(pred + 1));
        }
    }
    P.min_score = 0.0f;
    if (sources1 && sinks1) P.min_score = gaps.measure_gap_ss(*sources1, *sources2, *sinks1, *sinks2).second;  // :2419-2424
    detail::collect_events(P, bank, forward_edges, graph1, ranks, order1);
    return P;
}

// Flat problem of the gap-free sparse_chain_dp (anchorer.hpp:1511-1750): one search tree per (chain1, chain2)
// keyed (index on chain2, match) -- the same machinery with a single diagonal per chain pair and no gap pieces.
template <class MBank, class FwdEdges, class BGraph, class XMerge, class MatchSets, class TopoOrder, class WeightFn>
ChainProblem build_gapfree_chain_problem(const MBank& bank, const FwdEdges& forward_edges, const BGraph& graph1,
                                         const TopoOrder& order1, const XMerge& chain_merge1, const XMerge& chain_merge2,
                                         const MatchSets& match_sets, size_t num_match_sets,
                                         const std::vector<uint64_t>* sources1, const std::vector<uint64_t>* sources2,
                                         const std::vector<uint64_t>* sinks1, const std::vector<uint64_t>* sinks2,
                                         const WeightFn& weight_of) {
    typedef float ScoreFloat;
    const ScoreFloat mininf = std::numeric_limits<ScoreFloat>::lowest();
    ChainProblem P;
    P.num_pw = 0;
    P.n_chain1 = (int32_t)chain_merge1.chain_size();
    P.n_chain2 = (int32_t)chain_merge2.chain_size();
    detail::RankTable<MBank, MatchSets> ranks(bank, match_sets, num_match_sets, P);
    const size_t C1 = chain_merge1.chain_size(), C2 = chain_merge2.chain_size();
    for (auto it = bank.begin(), end = bank.end(); it != end; ++it) {
        const auto& match_set = bank.match_set(*it);
        const uint64_t start1 = bank.walk1(*it).front(), start2 = bank.walk2(*it).front();
        const uint64_t end1 = bank.walk1(*it).back(), end2 = bank.walk2(*it).back();
        ScoreFloat weight = weight_of(match_set);
        P.weight.push_back(weight);
        if (sources1) {  // anchorer.hpp:1559-1577
            bool found1 = false, found2 = false;
            for (auto src_id1 : *sources1)
                if (src_id1 == start1 || chain_merge1.reachable(src_id1, start1)) { found1 = true; break; }
            for (auto src_id2 : *sources2)
                if (src_id2 == start2 || chain_merge2.reachable(src_id2, start2)) { found2 = true; break; }
            if (!found1 || !found2) weight = mininf;
        }
        P.dp_init.push_back(weight);
        ScoreFloat fin = 0.0;  // anchorer.hpp:1730-1747
        if (sinks1) {
            fin = mininf;
            for (auto snk_id1 : *sinks1) {
                for (auto snk_id2 : *sinks2)
                    if ((snk_id1 == end1 || chain_merge1.reachable(end1, snk_id1)) &&
                        (snk_id2 == end2 || chain_merge2.reachable(end2, snk_id2))) { fin = 0.0; break; }
                if (fin == ScoreFloat(0.0)) break;
            }
        }
        P.final_term.push_back(fin);
        // every chain of graph 1 has a tree over all matches ending on a given chain of graph 2, key = index on that
        // chain (:1546-1552, 1595-1603); the match is entered only into the tree of its own graph-1 chain (:1657-1666)
        for (size_t c1 = 0; c1 < C1; ++c1) {
            P.ins_p1.push_back((uint32_t)c1);
            P.ins_p2.push_back((uint32_t)chain_merge2.chain(end2).first);
            P.ins_shift.push_back(0);
            P.ins_offset.push_back((uint32_t)chain_merge2.chain(end2).second);
            P.ins_active.push_back(c1 == chain_merge1.chain(end1).first ? 1 : 0);
        }
        P.ins_off.push_back((int64_t)P.ins_p1.size());
        for (size_t c1 = 0; c1 < C1; ++c1) P.qa1.push_back(0);
        for (size_t c2 = 0; c2 < C2; ++c2) {
            P.qa2.push_back(0);
            auto pred = chain_merge2.predecessor_index(start2, c2);  // :1699-1712: -1 = no reachable node, else [0, pred + 1)
            P.qoff.push_back((uint32_t)(decltype(pred))(pred + 1));
        }
    }
    P.min_score = 0.0f;  // traceback_sparse_dp(..., final_term, 0.0, ...), anchorer.hpp:1749
    detail::collect_events(P, bank, forward_edges, graph1, ranks, order1);
    return P;
}

}  // namespace centrolign_b200

#endif  // CENTROLIGN_B200_CHAIN_HPP
