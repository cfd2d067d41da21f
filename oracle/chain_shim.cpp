/*
 * oracle/chain_shim.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Fixture generator and CPU baseline for the sparse anchor-chaining DP, built on the UNMODIFIED reference
 * (its objects are compiled from /root/reference by integration/Makefile; nothing is copied).  For a FASTA of
 * 2..4 sequences it builds the chaining problem the reference's Core would see --
 *   "pair": sequence 0 against sequence 1 (two single-path graphs),
 *   "msa":  the merged graph of the first ceil(n/2) sequences (made by the reference's own Core) against the
 *           merged graph of the rest (multi-path graphs: several chains per node, real forward-edge fan-in),
 * finds matches with the reference's PathMatchFinder, keeps the smallest match sets up to a pair budget, and runs
 *   Anchorer::sparse_chain_dp          (include/centrolign/anchorer.hpp:1511-1750)
 *   Anchorer::sparse_affine_chain_dp   (include/centrolign/anchorer.hpp:1812-2471)
 * in the production instantiation (anchorer.hpp:1271-1278, 1286-1291) with global sources / sinks, exactly as
 * Anchorer::anchor_chain does (anchorer.hpp:1069-1075).  It writes, into one binary file, the flat problems
 * produced by the product wrapper (centrolign_b200/hostcpp/chain_b200.hpp) and the chains the reference found,
 * as match ranks.  tests/golden/make_chain_golden.py turns such files into the committed fixtures.
 *
 * usage: chain_fixture <fasta> <out.bin> <pair|msa> <max_pairs> [scale]
 */
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "centrolign/anchorer.hpp"
#include "centrolign/core.hpp"
#include "centrolign/forward_edges.hpp"
#include "centrolign/match_bank.hpp"
#include "centrolign/match_finder.hpp"
#include "centrolign/modify_graph.hpp"
#include "centrolign/parameters.hpp"
#include "centrolign/path_merge.hpp"
#include "centrolign/post_switch_distances.hpp"
#include "centrolign/topological_order.hpp"
#include "centrolign/tree.hpp"
#include "centrolign/utility.hpp"

#include "chain_b200.hpp"

using namespace centrolign;

namespace {

class OpenAnchorer : public Anchorer {
public:
    OpenAnchorer(const Anchorer& other) : Anchorer(other) {}
    using Anchorer::gap_extend;
    using Anchorer::gap_open;
    using Anchorer::generate_forward_edge_masks;
    using Anchorer::score_function;
    using Anchorer::sparse_affine_chain_dp;
    using Anchorer::sparse_chain_dp;
};

typedef PathMerge<uint32_t, uint8_t> XMerge;  // core.hpp:336-338
typedef MatchBank<uint32_t, uint16_t, float> Bank;
typedef ForwardEdges<XMerge::node_id_t, XMerge::chain_id_t> FwdEdges;
typedef std::vector<std::pair<int32_t, Bank::match_id_t>> ShiftMatchVec;
typedef std::vector<std::pair<uint32_t, Bank::match_id_t>> DistMatchVec;

FILE* g_out = nullptr;

void put_bytes(const std::string& name, uint32_t code, const void* data, uint64_t n, size_t elem) {
    const uint32_t len = (uint32_t)name.size();
    fwrite(&len, 4, 1, g_out);
    fwrite(name.data(), 1, len, g_out);
    fwrite(&code, 4, 1, g_out);
    fwrite(&n, 8, 1, g_out);
    if (n) fwrite(data, elem, n, g_out);
}
// dtype codes: 0 f32, 1 i32, 2 u32, 3 i64, 4 f64
void put(const std::string& n, const std::vector<float>& v) { put_bytes(n, 0, v.data(), v.size(), 4); }
void put(const std::string& n, const std::vector<int32_t>& v) { put_bytes(n, 1, v.data(), v.size(), 4); }
void put(const std::string& n, const std::vector<uint32_t>& v) { put_bytes(n, 2, v.data(), v.size(), 4); }
void put(const std::string& n, const std::vector<int64_t>& v) { put_bytes(n, 3, v.data(), v.size(), 8); }
void put(const std::string& n, const std::vector<double>& v) { put_bytes(n, 4, v.data(), v.size(), 8); }

void dump(const std::string& pre, const centrolign_b200::ChainProblem& P, const std::vector<anchor_t>& chain, double ref_ms) {
    put(pre + "params", std::vector<double>{(double)P.num_pw, P.gap_open[0], P.gap_open[1], P.gap_open[2], P.gap_extend[0],
                                            P.gap_extend[1], P.gap_extend[2], P.scale, (double)P.n_chain1, (double)P.n_chain2,
                                            ref_ms});
    put(pre + "min_score", std::vector<float>{P.min_score});
    put(pre + "weight", P.weight);
    put(pre + "dp_init", P.dp_init);
    put(pre + "final_term", P.final_term);
    put(pre + "end_off", P.end_off);
    put(pre + "end_match", P.end_match);
    put(pre + "qry_off", P.qry_off);
    put(pre + "qry_match", P.qry_match);
    put(pre + "qry_chain1", P.qry_chain1);
    put(pre + "ins_off", P.ins_off);
    put(pre + "ins_p1", P.ins_p1);
    put(pre + "ins_p2", P.ins_p2);
    put(pre + "ins_shift", P.ins_shift);
    put(pre + "ins_offset", P.ins_offset);
    put(pre + "ins_active", P.ins_active.empty() ? std::vector<uint32_t>(P.ins_offset.size(), 1u)
                                                 : std::vector<uint32_t>(P.ins_active.begin(), P.ins_active.end()));
    put(pre + "qa1", P.qa1);
    put(pre + "qa2", P.qa2);
    put(pre + "qoff", P.qoff);
    std::map<std::tuple<size_t, size_t, size_t>, int64_t> rank;
    for (size_t r = 0; r < P.ids.size(); ++r) rank[P.ids[r]] = (int64_t)r;
    std::vector<int64_t> expect;
    for (const auto& a : chain) expect.push_back(rank.at(std::make_tuple(a.match_set, a.idx1, a.idx2)));
    put(pre + "expect_chain", expect);
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

Subproblem merged(std::vector<std::pair<std::string, std::string>> seqs) {
    if (seqs.size() == 1) {
        Subproblem sp;
        sp.graph = make_base_graph(seqs[0].first, seqs[0].second);
        sp.tableau = add_sentinels(sp.graph, 5, 6);  // src/execution.cpp:75-76
        return sp;
    }
    std::string newick = "(";
    for (size_t i = 0; i < seqs.size(); ++i) newick += (i ? "," : "") + seqs[i].first;
    newick += ");";
    Tree tree(newick);
    Core core(std::move(seqs), std::move(tree));
    Parameters().apply(core);
    core.execute();
    return core.root_subproblem();
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 5) {
        std::cerr << "usage: chain_fixture <fasta> <out.bin> <pair|msa> <max_pairs> [scale]\n";
        return 2;
    }
    logging::level = logging::Silent;
    std::ifstream fin(argv[1]);
    auto seqs = parse_fasta(fin);
    const std::string mode = argv[3];
    const size_t max_pairs = std::stoull(argv[4]);
    const double scale = argc > 5 ? std::stod(argv[5]) : 1.0;
    if (seqs.size() < 2) return 2;
    const size_t half = mode == "pair" ? 1 : (seqs.size() + 1) / 2;
    std::vector<std::pair<std::string, std::string>> left(seqs.begin(), seqs.begin() + half);
    std::vector<std::pair<std::string, std::string>> right(seqs.begin() + half, mode == "pair" ? seqs.begin() + 2 : seqs.end());
    Subproblem sp1 = merged(left), sp2 = merged(right);
    reassign_sentinels(sp1.graph, sp1.tableau, 5, 6);  // core.hpp:287-288
    reassign_sentinels(sp2.graph, sp2.tableau, 7, 8);

    // modules configured the way the CLI configures them
    std::string cfg_newick = "(";
    for (size_t i = 0; i < seqs.size(); ++i) cfg_newick += (i ? "," : "") + seqs[i].first;
    Core cfg(std::vector<std::pair<std::string, std::string>>(seqs), Tree(cfg_newick + ");"));  // wires the score function
    Parameters().apply(cfg);
    auto matches = cfg.path_match_finder.find_matches(sp1.graph, sp2.graph, sp1.tableau, sp2.tableau);
    // keep the smallest match sets within the pair budget (any subset is a valid input of the DP functions)
    std::stable_sort(matches.begin(), matches.end(), [](const match_set_t& a, const match_set_t& b) {
        return a.walks1.size() * a.walks2.size() < b.walks1.size() * b.walks2.size();
    });
    size_t kept = 0, pairs = 0;
    while (kept < matches.size() && pairs + matches[kept].walks1.size() * matches[kept].walks2.size() <= max_pairs) {
        pairs += matches[kept].walks1.size() * matches[kept].walks2.size();
        ++kept;
    }
    matches.resize(kept);

    XMerge xmerge1(sp1.graph, sp1.tableau), xmerge2(sp2.graph, sp2.tableau);
    OpenAnchorer anchorer(cfg.anchorer);
    const std::vector<uint64_t>& sources1 = sp1.graph.next(sp1.tableau.src_id);  // anchorer.hpp:1071-1072
    const std::vector<uint64_t>& sources2 = sp2.graph.next(sp2.tableau.src_id);
    const std::vector<uint64_t>& sinks1 = sp1.graph.previous(sp1.tableau.snk_id);
    const std::vector<uint64_t>& sinks2 = sp2.graph.previous(sp2.tableau.snk_id);
    auto weight_of = [&](const match_set_t& ms) -> float {
        return anchorer.score_function->anchor_weight(ms.count1, ms.count2, ms.walks1.front().size(), ms.full_length);
    };

    g_out = fopen(argv[2], "wb");
    if (!g_out) return 2;
    put("meta", std::vector<double>{(double)sp1.graph.node_size(), (double)sp2.graph.node_size(), (double)xmerge1.chain_size(),
                                    (double)xmerge2.chain_size(), (double)matches.size(), (double)pairs});

    // ---- the reference, production instantiations ----
    double t0 = now_ms();
    auto chain_gf = anchorer.sparse_chain_dp<uint32_t, uint32_t, uint16_t, uint32_t, float, DistMatchVec, std::vector<uint32_t>, Bank, FwdEdges>(
        matches, sp1.graph, xmerge1, xmerge2, matches.size(), true, &sources1, &sources2, &sinks1, &sinks2, nullptr);
    const double ms_gf = now_ms() - t0;
    t0 = now_ms();
    auto chain_af = anchorer.sparse_affine_chain_dp<uint32_t, uint16_t, uint32_t, int32_t, uint32_t, float, ShiftMatchVec, DistMatchVec,
                                                    std::vector<uint32_t>, std::vector<uint32_t>, Bank, FwdEdges>(
        matches, sp1.graph, sp2.graph, xmerge1, xmerge2, anchorer.gap_open, anchorer.gap_extend, scale, matches.size(), true,
        &sources1, &sources2, &sinks1, &sinks2, nullptr);
    const double ms_af = now_ms() - t0;
    t0 = now_ms();
    auto chain_local = anchorer.sparse_affine_chain_dp<uint32_t, uint16_t, uint32_t, int32_t, uint32_t, float, ShiftMatchVec, DistMatchVec,
                                                       std::vector<uint32_t>, std::vector<uint32_t>, Bank, FwdEdges>(
        matches, sp1.graph, sp2.graph, xmerge1, xmerge2, anchorer.gap_open, anchorer.gap_extend, scale, matches.size(), true);
    const double ms_local = now_ms() - t0;

    // ---- the flat problems, through the product wrapper ----
    Bank bank(sp1.graph, matches, matches.size(), true, nullptr);
    std::vector<bool> mask_to, mask_from;
    std::tie(mask_to, mask_from) = anchorer.generate_forward_edge_masks(sp1.graph, matches, matches.size());
    FwdEdges forward_edges(xmerge1, &mask_to, &mask_from);
    PostSwitchDistances<std::vector<uint32_t>> sd1(sp1.graph, xmerge1), sd2(sp2.graph, xmerge2);
    auto order1 = topological_order(sp1.graph);
    auto P_gf = centrolign_b200::build_gapfree_chain_problem(bank, forward_edges, sp1.graph, order1, xmerge1, xmerge2, matches,
                                                             matches.size(), &sources1, &sources2, &sinks1, &sinks2, weight_of);
    auto P_af = centrolign_b200::build_affine_chain_problem<int32_t>(bank, forward_edges, sd1, sd2, sp1.graph, order1, xmerge1, xmerge2,
                                                                     matches, matches.size(), anchorer.gap_open, anchorer.gap_extend,
                                                                     scale, &sources1, &sources2, &sinks1, &sinks2, weight_of);
    auto P_local = centrolign_b200::build_affine_chain_problem<int32_t>(
        bank, forward_edges, sd1, sd2, sp1.graph, order1, xmerge1, xmerge2, matches, matches.size(), anchorer.gap_open,
        anchorer.gap_extend, scale, (const std::vector<uint64_t>*)nullptr, (const std::vector<uint64_t>*)nullptr,
        (const std::vector<uint64_t>*)nullptr, (const std::vector<uint64_t>*)nullptr, weight_of);
    dump("gapfree.", P_gf, chain_gf, ms_gf);
    dump("affine.", P_af, chain_af, ms_af);
    dump("local.", P_local, chain_local, ms_local);
    fclose(g_out);
    std::cout << "nodes " << sp1.graph.node_size() << " x " << sp2.graph.node_size() << ", chains " << xmerge1.chain_size() << " x "
              << xmerge2.chain_size() << ", match sets " << matches.size() << ", pairs " << pairs << "; reference ms: gap-free "
              << ms_gf << ", affine " << ms_af << ", affine local " << ms_local << "; chain lengths " << chain_gf.size() << " / "
              << chain_af.size() << " / " << chain_local.size() << "\n";
    return 0;
}
