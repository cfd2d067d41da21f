// chain_device.cuh -- device-side layout of the sparse anchor-chaining DP (clb_chain_dp).
//
// The reference keeps, per pair of paths (p1, p2) of the two graphs,
//   * one MaxSearchTree per diagonal ("gap-free" trees, keyed (offset on p2, match));
//     reference: include/centrolign/anchorer.hpp:2136-2241, include/centrolign/max_search_tree.hpp
//   * 2*NumPW OrthogonalMaxSearchTrees keyed ((shift, match), offset on p2), all with the same keys;
//     reference: anchorer.hpp:2084-2111, include/centrolign/orthogonal_max_search_tree.hpp
// Both are static in SHAPE (implicit heap layout, filled in key order, max_search_tree.hpp:108-150) and only
// their values move.  Here the same shapes are laid out once on the host (chain_host.cu) and the value side
// is held in atomically-max-able words so that all updates of one DP step can run in parallel:
//   gap-free tree node x : best[x] = max over the subtree of x of pack(value, ~insertion sequence number)
//                          -- exactly the reference's subtree_max pointer, which moves only on a strictly
//                          greater value and therefore names the EARLIEST inserted maximum
//                          (max_search_tree.hpp:312-358);
//   orthogonal tree node r: a Fenwick tree over the elements of r's subtree sorted by offset, holding
//                          pack(value, outer node index): the reference's cross tree compares
//                          (value, outer index) pairs (orthogonal_max_search_tree.hpp:74), so its range
//                          maxima are unique, and its queries are always prefixes [0, offset)
//                          (anchorer.hpp:2396-2407), which a Fenwick tree over max answers.
// The queries walk the outer trees exactly as the reference's range_max does and take the FIRST block, in
// that walk's order, that attains the maximum (strict '>' replaces, max_search_tree.hpp:361-444,
// orthogonal_max_search_tree.hpp:340-470).
#pragma once
#include <stdint.h>

namespace clb {

constexpr int kChainMaxTrees = 6;  // 2 * NumPW orthogonal value sets

struct ChainArgs {
    // ---- problem (see clb_chain_problem) ----
    int num_pw;
    int n_chain1, n_chain2;
    double scale_ext[3];   // local_scale * gap_extend[k]
    double gap_open[3], gap_extend[3], scale;
    int64_t n_match;
    const float* weight;
    float* dp;             // [n_match] starts as dp_init
    uint32_t* backptr;     // [n_match] 0xffffffff = none
    int64_t n_step;
    const int64_t* sins_off;      // [n_step+1] insert entries of the step, reference order (= sequence numbers)
    const uint32_t* sins_entry;   // entry id
    const uint32_t* ent_match;    // [n_entry]
    const int64_t* qry_off;       // [n_step+1]
    const uint32_t* qry_match;
    const uint32_t* qry_chain1;
    const int32_t* qa1;
    const int32_t* qa2;
    const uint32_t* qoff;
    // ---- gap-free trees ----
    const int64_t* pair_grp_off;  // [n_chain1*n_chain2+1] groups (diagonals) of the pair, ascending shift
    const int32_t* grp_shift;     // [n_grp]
    const int64_t* grp_base;      // [n_grp] first node of the group's tree
    const uint32_t* grp_n;        // [n_grp]
    const uint32_t* gf_key;       // [n_entry] offset key, heap layout per group
    const uint32_t* gf_match;     // [n_entry]
    float* gf_val;                // [n_entry]
    unsigned long long* gf_best;  // [n_entry]
    const uint32_t* ent_gf_grp;   // [n_entry] group of the entry
    const uint32_t* ent_gf_node;  // [n_entry] heap index inside the group
    // ---- orthogonal trees ----
    const int64_t* pair_base;     // [npair+1] first outer node of the pair
    const int32_t* or_shift;      // [n_entry] heap layout per pair
    const uint32_t* or_off;       // [n_entry]
    const uint32_t* or_match;     // [n_entry]
    float* or_val;                // [2*num_pw][n_entry]
    const int64_t* in_base;       // [n_entry] per outer node: first slot of its inner list, -1 = none (outer spine)
    const uint32_t* in_n;         // [n_entry]
    const uint32_t* in_off;       // [n_inner] offsets of the inner list, ascending
    unsigned long long* bit;      // [2*num_pw][n_inner] Fenwick trees over max
    int64_t n_inner;
    int64_t n_entry;
    const uint32_t* ent_pair;     // [n_entry]
    const uint32_t* ent_or_node;  // [n_entry] outer heap index inside the pair
    const int32_t* ent_shift;     // [n_entry]
    const int64_t* ent_rank_off;  // [n_entry+1] ranks of the entry in the inner lists of its non-spine ancestors, self first
    const uint32_t* ent_rank;
    // ---- per-step candidate exchange ----
    unsigned long long* cand_best;  // [n_match] pack(value, ~(query index in step * n_chain2 + chain2)), 0 = none
    uint32_t* cand_bp;              // [max queries per step * n_chain2]
    // ---- stats ----
    unsigned long long* counters;   // [0] tree queries answered
};

}  // namespace clb
