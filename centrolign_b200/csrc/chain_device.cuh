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
//
// Values are stored as order-preserving 32-bit images of the reference's floats ("ord", 0 = never entered), so
// that zero-filled memory is the initial state and comparisons are integer comparisons.  Everything about a query
// that does not depend on DP values -- its diagonal's tree, the shape of its three tree walks -- is computed for
// all queries at once by a preparation kernel (QueryRec), so that a DP step only has to look at values.
#pragma once
#include <stdint.h>

namespace clb {

constexpr int kChainMaxTrees = 6;            // 2 * NumPW orthogonal value sets
constexpr uint32_t kChainNone = 0xffffffffu;
constexpr int kChainSmallArena = 216 * 1024;   // problems whose whole arena fits run from shared memory

// One tree insertion (a match end on one path pair), in the reference's insertion order; the index of the
// record is the insertion sequence number that breaks ties inside the gap-free trees.
struct InsRec {
    uint32_t match;
    uint32_t gf_base;   // first node of the gap-free tree of its diagonal
    uint32_t gf_node;   // its heap index in that tree
    uint32_t or_base;   // first outer node of the path pair's orthogonal trees
    uint32_t or_node;   // its outer heap index
    int32_t shift;
    uint32_t rank_off;  // ranks in the inner lists of its non-spine ancestors, self first
    uint32_t n_rank;
};

// One (query, path of graph 2) pair.  A tree walk is stored as its split node S and the in-range decisions along
// the one conditional side of the walk (the other side takes every node), LSB first.
struct QueryRec {
    uint32_t match;
    float weight;
    uint32_t offset;    // 0 = nothing on this path reaches the match: no candidates
    int32_t q;          // query shift
    uint32_t gf_base, gf_n;   // gap-free tree of diagonal q, gf_n == 0 if there is none
    uint32_t gf_S, gf_bits;   // prefix walk over keys < offset
    uint32_t or_base, or_n;
    uint32_t ev_S, ev_bits;   // even pieces: shift > q (suffix walk)
    uint32_t or_base2, or_n2; // or_base, or_n again: every kind of work item reads two 16-byte quarters of the record
    uint32_t od_S, od_bits;   // odd pieces:  shift < q (prefix walk)
};

struct ChainArgs {
    int num_pw;
    int n_chain1, n_chain2;
    double scale_ext[3];   // local_scale * gap_extend[k]
    double gap_open[3], gap_extend[3], scale;
    int64_t n_match;
    float* dp;             // [n_match] starts as dp_init
    uint32_t* backptr;     // [n_match] kChainNone = none
    int64_t n_step;
    const int64_t* sins_off;   // [n_step+1] insertions of the step
    const InsRec* ins;
    int64_t n_ins;
    const int64_t* qry_off;    // [n_step+1]
    const uint32_t* qry_match;
    QueryRec* qrec;            // [n_qry * n_chain2]
    int64_t n_qry;
    // raw query data, read by the preparation kernel only
    const float* weight;
    const uint32_t* qry_chain1;
    const int32_t* qa1;
    const int32_t* qa2;
    const uint32_t* qoff;
    const int64_t* pair_grp_off;  // [npair+1] diagonals of the pair, ascending shift
    const int32_t* grp_shift;
    const uint32_t* grp_base;
    const uint32_t* grp_n;
    const uint32_t* pair_base;    // [npair+1] first outer node of the pair
    // gap-free trees
    const uint32_t* gf_key;       // [n_entry] offset key, heap layout per tree
    const uint32_t* gf_match;     // [n_entry]
    uint32_t* gf_ord;             // [n_entry] value of the node
    unsigned long long* gf_best;  // [n_entry] subtree maximum
    // orthogonal trees
    const int32_t* or_shift;      // [n_entry] heap layout per pair
    const uint32_t* or_off;
    const uint32_t* or_match;
    uint32_t* or_ord;             // [2*num_pw][n_entry]
    const uint32_t* in_base;      // [n_entry] per outer node: first slot of its inner list, kChainNone on the outer spines
    const uint32_t* in_n;
    const uint32_t* in_off;       // [n_inner] offsets of the inner lists, ascending
    unsigned long long* bit;      // [2*num_pw][n_inner] Fenwick trees over max
    const uint32_t* ent_rank;     // [n_inner]
    int64_t n_inner;
    int64_t n_entry;
    // per-step candidate exchange
    unsigned long long* cand_best;  // [2][n_match] by step parity: pack(value, ~order), 0 = none
    uint32_t* cand_bp;              // [2][cand_bp_stride] by step parity, cand_bp_stride = max queries per step * n_chain2 * (2*num_pw+1)
    int64_t cand_bp_stride;
    unsigned long long* counters;   // [0] tree queries answered
    // ranks of the subtree blocks of every orthogonal walk, computed by the preparation kernel: per (query, chain2,
    // parity) rank_stride words, one per subtree block in walk order; nullptr = computed on the fly
    uint32_t* rank_pool;
    int rank_stride;
    const char* arena_base;  // the arena all pointers above point into, and its size
    int64_t arena_bytes;
    int64_t copy_bytes;      // shared-memory kernels: this many bytes are copied from arena_base, the rest of the arena starts as zero
    float* out_dp;           // batched launch: compact result arrays of the problem (nullptr: results go back into dp / backptr)
    uint32_t* out_backptr;
    int sync_mode;     // how the phases of a step are separated: 0 one CTA (__syncthreads), 1 one thread-block cluster
                       // (barrier.cluster), 2 cooperative grid (grid.sync)
};

}  // namespace clb
