// chain_kernels.cu -- sm_100a kernel of the sparse anchor-chaining DP (clb_chain_dp).
//
// Reference: Anchorer::sparse_affine_chain_dp main loop (include/centrolign/anchorer.hpp:2290-2417) and
// Anchorer::sparse_chain_dp main loop (:1640-1728); search structures max_search_tree.hpp:312-444 and
// orthogonal_max_search_tree.hpp:293-470.  Data layout and the equivalence argument: chain_device.cuh.
//
// A preparation kernel first computes, for all (query, path of graph 2) pairs at once, everything that does not
// depend on DP values: the gap-free tree of the query's diagonal and the shape of its three tree walks.
// Then one persistent cooperative kernel walks the graph-1 nodes ("steps") in topological order.  Per step:
//   A  every match ending here enters its DP value into the gap-free tree of its diagonal and into the
//      2*NumPW orthogonal value sets, for every path pair of its end point -- all entries of the step in
//      parallel, one warp per entry, lanes over the ancestors of the entry's node (atomicMax on packed words);
//   B  every (match starting behind a forward edge, path of graph 2) pair is one warp: it answers the
//      gap-free query and the 2*NumPW orthogonal queries by the reference's own tree walks, lanes over the
//      blocks of a walk, keeps the first strictly greater candidate in the reference's order (gap-free, then
//      pieces 0..2P-1) and posts it with atomicMax keyed by (value, earlier query first);
//   C  the winning candidate of each match is applied if it beats the match's current value (update_dp,
//      match_bank.hpp:171-184).
// Phases are separated by a grid-wide barrier (cooperative groups), or __syncthreads() when one CTA runs.
// Scores are float, gap terms double, every operation with an explicit rounding intrinsic so that no FMA
// contraction can change a bit relative to the reference's scalar code.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "chain_device.cuh"

namespace cg = cooperative_groups;

namespace clb {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxBlocks = 132;  // entries of one tree walk: 1 + 4 * depth, depth <= 32

__device__ __forceinline__ float mininf() { return -3.402823466e+38f; }  // numeric_limits<float>::lowest()

__device__ __forceinline__ uint32_t ford(float f) {  // order-preserving map float -> uint32 (-0 folded into +0)
    const uint32_t b = __float_as_uint(__fadd_rn(f, 0.0f));
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float funord(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ unsigned long long pack(float v, uint32_t low) {
    return ((unsigned long long)ford(v) << 32) | low;
}
constexpr uint32_t kOrdLowest = 0x00800000u;  // ford(lowest()); 0 = never entered; both fail "ord > kOrdLowest"

// Loads.  SM = the whole problem was copied into shared memory (small problems: the Anchorer's fill-in pass makes
// thousands of chaining calls with a few dozen matches each, where a step is nothing but memory latency).
template <bool SM, class T>
__device__ __forceinline__ T ld_const(const T* p) {  // never written while a kernel runs
    if (SM) return *p;
    return __ldg(p);
}
template <bool SM, class T>
__device__ __forceinline__ T ld_live(const T* p) {  // written by other warps between barriers
    if (SM) return *reinterpret_cast<const volatile T*>(p);
    return __ldcg(p);
}
template <bool SM>
__device__ __forceinline__ uint4 ld_const4(const uint4* p) {
    if (SM) return *p;
    return __ldg(p);
}

struct WalkEntry {
    uint32_t node;  // heap index
    uint32_t kind;  // 0: the node itself, 1: the whole subtree of `node`
};

struct WarpScratch {
    WalkEntry blk[kMaxBlocks];
};

// The walk shared by MaxSearchTree::range_max (max_search_tree.hpp:361-444) and the outer level of
// OrthogonalMaxSearchTree::range_max (orthogonal_max_search_tree.hpp:340-470), for one-sided key ranges:
//   prefix: in_range(x) <=> key[x] <  bound  (nothing lies below the range: its left walk takes every node)
//   suffix: in_range(x) <=> key[x] >  bound  (nothing lies above the range: its right walk takes every node)
// walk_shape finds the split node S and the in-range decisions along the conditional side (LSB first).  The walk
// is uniform over the warp; to avoid one memory round trip per tree level, the range flags of a node and of its
// descendants four levels down (31 nodes) are fetched by the 31 lanes at once and the next moves are resolved
// from the resulting bit mask.
template <class InRange>
__device__ void walk_shape(uint32_t n, bool prefix, const InRange& in_range, int lane, uint32_t& S_out, uint32_t& bits_out) {
    uint32_t wbase = kChainNone;
    unsigned wbits = 0;
    auto flag = [&](uint32_t y) -> bool {
        const uint32_t rel = y + 1;
        int d = -1;
        if (wbase != kChainNone) {
            const uint32_t b = wbase + 1;
            d = (31 - __clz(rel)) - (31 - __clz(b));
            if (d < 0 || d > 4 || (rel >> d) != b) d = -1;
        }
        if (d < 0) {  // fetch the window rooted at y
            wbase = y;
            bool f = false;
            if (lane < 31) {
                uint32_t idx = y;
                if (lane < 30) {
                    const int dj = 31 - __clz(lane + 2);
                    idx = (rel << dj) - 1 + (uint32_t)(lane + 2 - (1 << dj));
                    if ((rel << dj) < rel) idx = kChainNone;  // overflow: no such node
                }
                if (idx < n) f = in_range(idx);
            }
            wbits = __ballot_sync(kFull, f);
            d = 0;
        }
        if (d == 0) return (wbits >> 30) & 1u;
        const uint32_t k = rel - ((wbase + 1) << d);
        return (wbits >> ((1u << d) - 2 + k)) & 1u;
    };
    uint32_t cursor = 0;
    while (cursor < n && !flag(cursor)) cursor = prefix ? 2 * cursor + 1 : 2 * cursor + 2;
    S_out = cursor < n ? cursor : kChainNone;
    uint32_t bits = 0;
    if (cursor < n) {
        int nbit = 0;
        uint32_t c = prefix ? 2 * cursor + 2 : 2 * cursor + 1;  // the conditional side: right walk of a prefix, left walk of a suffix
        while (c < n) {
            const bool f = flag(c);
            bits |= (uint32_t)f << nbit;
            ++nbit;
            c = (f == prefix) ? 2 * c + 2 : 2 * c + 1;  // prefix: in range -> right, else left; suffix: in range -> left, else right
        }
    }
    bits_out = bits;
}

// Expands a stored walk into the blocks the reference tests, in its order: S, the left walk, the right walk
// (each taken node followed by the subtree hanging off the walk).  Pure arithmetic; uniform, lane 0 writes.
__device__ int walk_blocks(uint32_t n, bool prefix, uint32_t S, uint32_t bits, WalkEntry* blk, int lane) {
    int nb = 0;
    auto push = [&](uint32_t node, uint32_t kind) {
        if (lane == 0 && nb < kMaxBlocks) blk[nb] = WalkEntry{node, kind};
        ++nb;
    };
    push(S, 0);
    uint32_t lc = 2 * S + 1, rc = 2 * S + 2;
    int i = 0;
    while (lc < n) {  // leftward: right subtrees hang entirely inside the range
        if (prefix || ((bits >> i++) & 1u)) {
            push(lc, 0);
            if (2 * lc + 2 < n) push(2 * lc + 2, 1);
            lc = 2 * lc + 1;
        } else {
            lc = 2 * lc + 2;
        }
    }
    while (rc < n) {  // rightward: left subtrees hang entirely inside
        if (!prefix || ((bits >> i++) & 1u)) {
            push(rc, 0);
            if (2 * rc + 1 < n) push(2 * rc + 1, 1);
            rc = 2 * rc + 2;
        } else {
            rc = 2 * rc + 1;
        }
    }
    return nb;
}

// first index in [lo, hi) whose value is >= q, by 32-ary search over the warp (uniform result)
template <bool SM>
__device__ int64_t warp_lower_bound(const int32_t* arr, int64_t lo, int64_t hi, int q, int lane) {
    while (hi - lo > 32) {
        const int64_t step = (hi - lo + 31) / 32;
        const int64_t at = lo + (int64_t)lane * step;
        const bool less = at < hi && ld_const<SM>(&arr[at]) < q;
        const int cnt = __popc(__ballot_sync(kFull, less));  // pivots below q form a prefix
        if (cnt == 0) return lo;
        const int64_t nlo = lo + (int64_t)(cnt - 1) * step + 1;
        hi = min(hi, lo + (int64_t)cnt * step);
        lo = nlo;
    }
    const bool less = lo + lane < hi && ld_const<SM>(&arr[lo + lane]) < q;
    return lo + __popc(__ballot_sync(kFull, less));
}

// number of entries of the ascending list `a[0, n)` that are < key, 8-ary search with independent probes
template <bool SM>
__device__ __forceinline__ uint32_t count_less(const uint32_t* a, uint32_t n, uint32_t key) {
    uint32_t lo = 0, hi = n;  // answer in [lo, hi]
    while (hi - lo > 7) {
        const uint32_t step = (hi - lo) >> 3;
        uint32_t v[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) v[i] = ld_const<SM>(&a[lo + (i + 1) * step - 1]);
        int c = 0;
#pragma unroll
        for (int i = 0; i < 7; ++i) c += v[i] < key;
        const uint32_t nlo = lo + c * step;
        hi = c == 7 ? hi : lo + (c + 1) * step - 1;
        lo = nlo;
    }
    uint32_t v[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) v[i] = lo + i < hi ? ld_const<SM>(&a[lo + i]) : kChainNone;
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) c += (lo + i < hi) && v[i] < key;
    return lo + c;
}

}  // namespace

// Everything about a (query, path of graph 2) pair that does not depend on DP values (anchorer.hpp:2374-2381).
template <bool SM>
__device__ void prepare_queries(const ChainArgs& A, WalkEntry* blk, int lane, int64_t gwarp, int64_t nwarp) {
    const int C1 = A.n_chain1, C2 = A.n_chain2;
    for (int64_t qc = gwarp; qc < A.n_qry * C2; qc += nwarp) {
        const int64_t k = qc / C2;
        const int c2 = (int)(qc - k * C2);
        const uint32_t m = A.qry_match[k], c1 = A.qry_chain1[k];
        QueryRec r;
        r.match = m;
        r.weight = A.weight[m];
        r.offset = A.qoff[(int64_t)m * C2 + c2];
        r.q = (int)((uint32_t)A.qa1[(int64_t)m * C1 + c1] - (uint32_t)A.qa2[(int64_t)m * C2 + c2]);
        r.gf_base = r.gf_n = r.or_base = r.or_n = 0;
        r.gf_S = r.ev_S = r.od_S = kChainNone;
        r.gf_bits = r.ev_bits = r.od_bits = 0;
        r.pad[0] = r.pad[1] = 0;
        if (r.offset != 0) {
            const int64_t pair = (int64_t)c1 * C2 + c2;
            const int64_t g1 = A.pair_grp_off[pair + 1];
            const int64_t g = warp_lower_bound<SM>(A.grp_shift, A.pair_grp_off[pair], g1, r.q, lane);
            if (g < g1 && ld_const<SM>(&A.grp_shift[g]) == r.q) {  // the diagonal exists (anchorer.hpp:2379-2382)
                r.gf_base = A.grp_base[g];
                r.gf_n = A.grp_n[g];
                const uint32_t* key = A.gf_key + r.gf_base;
                const uint32_t offset = r.offset;
                walk_shape(r.gf_n, true, [&](uint32_t x) { return ld_const<SM>(&key[x]) < offset; }, lane, r.gf_S, r.gf_bits);
            }
            if (A.num_pw > 0) {
                r.or_base = A.pair_base[pair];
                r.or_n = A.pair_base[pair + 1] - r.or_base;
                const int32_t* shift = A.or_shift + r.or_base;
                const int q = r.q;
                if (r.or_n) {
                    walk_shape(r.or_n, false, [&](uint32_t x) { return ld_const<SM>(&shift[x]) > q; }, lane, r.ev_S, r.ev_bits);
                    walk_shape(r.or_n, true, [&](uint32_t x) { return ld_const<SM>(&shift[x]) < q; }, lane, r.od_S, r.od_bits);
                    if (A.rank_pool) {  // how many elements of every subtree block lie below the query offset
                        for (int par = 0; par < 2; ++par) {
                            const uint32_t S = par ? r.od_S : r.ev_S;
                            if (S == kChainNone) continue;
                            __syncwarp();
                            const int nb = walk_blocks(r.or_n, par == 1, S, par ? r.od_bits : r.ev_bits, blk, lane);
                            __syncwarp();
                            uint32_t* pool = A.rank_pool + ((int64_t)qc * 2 + par) * A.rank_stride;
                            int seen = 0;  // subtree blocks before this group of 32
                            for (int b0 = 0; b0 < nb && b0 < kMaxBlocks; b0 += 32) {
                                const int b = b0 + lane;
                                const bool sub = b < nb && b < kMaxBlocks && blk[b].kind == 1;
                                const unsigned subm = __ballot_sync(kFull, sub);
                                if (sub) {
                                    const int j = seen + __popc(subm & ((1u << lane) - 1));
                                    const uint32_t node = blk[b].node;
                                    if (j < A.rank_stride)
                                        pool[j] = count_less<SM>(A.in_off + ld_const<SM>(&A.in_base[r.or_base + node]), ld_const<SM>(&A.in_n[r.or_base + node]), r.offset);
                                }
                                seen += __popc(subm);
                            }
                        }
                    }
                }
            }
        }
        if (lane == 0) A.qrec[qc] = r;
    }
}

__global__ void __launch_bounds__(256) chain_prepare_kernel(const ChainArgs A) {
    __shared__ WarpScratch prep_scratch[8];
    prepare_queries<false>(A, prep_scratch[threadIdx.x >> 5].blk, threadIdx.x & 31,
                           ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, ((int64_t)gridDim.x * blockDim.x) >> 5);
}

template <bool SM>
__device__ __forceinline__ void chain_steps(const ChainArgs& A, WarpScratch* scratch) {
    cg::grid_group grid = cg::this_grid();
    const bool multi = !SM && gridDim.x > 1;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // the shared-memory flavour is one CTA per problem whatever the grid is (a batch launch has one CTA per problem)
    const int64_t cta = SM ? 0 : (int64_t)blockIdx.x, nctas = SM ? 1 : (int64_t)gridDim.x;
    const int64_t gwarp = cta * kWarps + wib, nwarp = nctas * kWarps;
    const int64_t gthread = cta * kThreads + threadIdx.x, nthread = nctas * kThreads;
    const int C2 = A.n_chain2, P = A.num_pw, T = 2 * A.num_pw;
    const int n_type = P > 0 ? 3 : 1;        // work items per (query, chain2): gap-free tree, even pieces, odd pieces
    const uint32_t slots = (uint32_t)(T + 1);  // candidate slots per (query, chain2), in the reference's order
    WalkEntry* blk = scratch[wib].blk;
    unsigned long long n_tree_queries = 0;

    auto barrier = [&]() {
        if (multi) grid.sync();
        else __syncthreads();
    };
    unsigned long long* post_to = A.cand_best;
    auto post = [&](uint32_t m, int64_t qc, uint32_t slot, float cand, uint32_t bp) {  // lane 0 only
        const uint32_t order = (uint32_t)qc * slots + slot;
        A.cand_bp[order] = bp;
        atomicMax(&post_to[m], pack(cand, ~order));
    };

    // update_dp of the winners of one step (match_bank.hpp:171-184).  Runs in the same phase as the insertions of
    // the next step; a match may be queried in one step and end in the next, so an insertion takes the maximum of
    // the stored value and the pending winner (effective_dp).  The candidate words are double-buffered by step
    // parity and cleared one phase later, so nothing a concurrent reader looks at is ever reset under it.
    auto apply_winners = [&](int64_t q0, int64_t q1, const unsigned long long* cand) {
        for (int64_t qi = gthread; qi < q1 - q0; qi += nthread) {
            const uint32_t m = A.qry_match[q0 + qi];
            const unsigned long long pk = ld_live<SM>(&cand[m]);
            if (!pk) continue;
            const uint32_t order = ~(uint32_t)pk;
            if ((int64_t)(order / ((uint32_t)C2 * slots)) != qi) continue;  // the winner was posted by another query of this match
            const float v = funord((uint32_t)(pk >> 32));
            if (v > ld_live<SM>(&A.dp[m])) {
                A.dp[m] = v;
                A.backptr[m] = ld_live<SM>(&A.cand_bp[order]);
            }
        }
    };
    auto effective_dp = [&](uint32_t m, const unsigned long long* cand) -> float {
        const unsigned long long pk = ld_live<SM>(&cand[m]);
        float dpv = ld_live<SM>(&A.dp[m]);
        if (pk) {
            const float v = funord((uint32_t)(pk >> 32));
            if (v > dpv) dpv = v;
        }
        return dpv;
    };

    int64_t pq0 = 0, pq1 = 0;  // queries of the previous step, whose winners are still to be applied
    for (int64_t s = 0; s < A.n_step; ++s) {
        // ---------------- update_dp of the previous step + A: inserts (anchorer.hpp:2301-2345) ----------------
        const int64_t i0 = A.sins_off[s], i1 = A.sins_off[s + 1];
        const int64_t q0 = A.qry_off[s], q1 = A.qry_off[s + 1];
        unsigned long long* cand_prev = A.cand_best + ((s + 1) & 1) * A.n_match;  // posted in step s-1
        unsigned long long* cand_cur = A.cand_best + (s & 1) * A.n_match;         // posted in this step
        if (pq0 != pq1) {
            apply_winners(pq0, pq1, cand_prev);
            if (A.split_phases) barrier();
        }
        for (int64_t i = i0 + gwarp; i < i1; i += nwarp) {
            const uint4 ra = ld_const4<SM>(reinterpret_cast<const uint4*>(&A.ins[i]));
            const uint4 rb = ld_const4<SM>(reinterpret_cast<const uint4*>(&A.ins[i]) + 1);
            const uint32_t m = ra.x, gf_base = ra.y, gf_node = ra.z, or_base = ra.w, or_node = rb.x, rank_off = rb.z;
            const int shift = (int)rb.y, nr = (int)rb.w;
            int64_t ib = 0;
            uint32_t cn = 0, rank = 0;
            if (P > 0 && lane < nr) {  // lane = level: the node itself and its ancestors below the outer spines
                const uint32_t a = ((or_node + 1) >> lane) - 1;
                ib = ld_const<SM>(&A.in_base[or_base + a]);
                cn = ld_const<SM>(&A.in_n[or_base + a]);
                rank = ld_const<SM>(&A.ent_rank[rank_off + lane]);
            }
            const float dpv = effective_dp(m, cand_prev);
            if (!(dpv > mininf())) continue;  // entering lowest() changes nothing in the reference's trees
            {   // gap-free tree of the entry's diagonal: lanes = the node and its ancestors
                if (lane == 0) A.gf_ord[gf_base + gf_node] = ford(dpv);
                const uint32_t anc = (gf_node + 1) >> lane;
                if (anc) atomicMax(&A.gf_best[gf_base + anc - 1], pack(dpv, ~(uint32_t)i));
            }
            for (int t = 0; t < T; ++t) {
                // anchorer.hpp:2328-2335: odd pieces add, even pieces subtract local_scale * gap_extend * shift
                const double gap = __dmul_rn(A.scale_ext[t >> 1], (double)shift);
                const float v = __double2float_rn((t & 1) ? __dadd_rn((double)dpv, gap) : __dsub_rn((double)dpv, gap));
                if (!(v > mininf())) continue;  // anchorer.hpp:2338
                if (lane == 0) A.or_ord[(int64_t)t * A.n_entry + or_base + or_node] = ford(v);
                if (lane < nr) {
                    const unsigned long long pk = pack(v, or_node);
                    unsigned long long* bit = A.bit + (int64_t)t * A.n_inner + ib;
                    for (uint32_t k = rank; k < cn; k |= k + 1) atomicMax(&bit[k], pk);
                }
            }
        }
        if (i0 != i1 || pq0 != pq1) barrier();
        // the previous step's candidate words are spent now: clear them while this step's queries run
        const bool cleared = pq0 != pq1;
        for (int64_t qi = gthread; qi < pq1 - pq0; qi += nthread) cand_prev[A.qry_match[pq0 + qi]] = 0;
        pq0 = q0;
        pq1 = q1;
        if (q0 == q1) {
            if (cleared) barrier();  // the clears must land before that buffer is posted to again
            continue;
        }

        // ------------------------------ B: queries (anchorer.hpp:2352-2416) ------------------------------
        post_to = cand_cur;
        const int64_t n_items = (q1 - q0) * C2 * n_type;
        for (int64_t item = gwarp; item < n_items; item += nwarp) {
            const int64_t qc = item / n_type;  // (query, chain2) pair inside the step
            const int type = (int)(item - qc * n_type);
            const uint4* rp = reinterpret_cast<const uint4*>(&A.qrec[q0 * C2 + qc]);
            const uint4 r0 = ld_const4<SM>(rp);  // match, weight, offset, q
            const uint32_t offset = r0.z;
            if (offset == 0) continue;  // nothing on this path reaches the match: every range [0, 0) is empty
            const uint32_t m = r0.x;
            const float w = __uint_as_float(r0.y);
            const int q = (int)r0.w;
            __syncwarp();
            if (type == 0) {
                // same diagonal (anchorer.hpp:2379-2389): MaxSearchTree::range_max((0, min), (offset, min))
                const uint4 r1 = ld_const4<SM>(rp + 1);  // gf_base, gf_n, gf_S, gf_bits
                if (r1.y == 0 || r1.z == kChainNone) continue;
                ++n_tree_queries;
                const uint32_t base = r1.x;
                const int nb = walk_blocks(r1.y, true, r1.z, r1.w, blk, lane);
                __syncwarp();
                unsigned long long lbest = 0;  // pack(ord, ~block index)
                uint32_t lsel = 0;             // kind << 31 | node, resolved to a match for the winner only
                for (int b = lane; b < nb && b < kMaxBlocks; b += 32) {
                    const WalkEntry we = blk[b];
                    uint32_t ord, sel;
                    if (we.kind == 0) {
                        ord = ld_live<SM>(&A.gf_ord[base + we.node]);
                        sel = we.node;
                    } else {
                        const unsigned long long pk = ld_live<SM>(&A.gf_best[base + we.node]);
                        ord = (uint32_t)(pk >> 32);
                        sel = 0x80000000u | (~(uint32_t)pk & 0x7fffffffu);  // the insertion sequence number (< 2^31)
                    }
                    if (ord > kOrdLowest && ord > (uint32_t)(lbest >> 32)) {
                        lbest = ((unsigned long long)ord << 32) | (~(uint32_t)b);
                        lsel = sel;
                    }
                }
                unsigned long long r = lbest;
#pragma unroll
                for (int d = 16; d; d >>= 1) {
                    const unsigned long long o = __shfl_xor_sync(kFull, r, d);
                    r = o > r ? o : r;
                }
                if (!r) continue;
                const unsigned src = __ffs(__ballot_sync(kFull, lbest == r)) - 1;
                const uint32_t sel = __shfl_sync(kFull, lsel, src);
                if (lane == 0) {
                    const uint32_t bm = (sel & 0x80000000u) ? ld_const<SM>(&A.ins[sel & 0x7fffffffu].match) : ld_const<SM>(&A.gf_match[base + sel]);
                    post(m, qc, 0, __fadd_rn(funord((uint32_t)(r >> 32)), w), bm);
                }
                continue;
            }
            // orthogonal trees of one parity (anchorer.hpp:2390-2413): par 0 = even pieces (shift > q), par 1 = odd (shift < q)
            const int par = type - 1;
            const uint4 r2 = ld_const4<SM>(rp + 2);  // or_base, or_n, ev_S, ev_bits
            const uint32_t ob = r2.x, n = r2.y;
            uint32_t S = r2.z, bits = r2.w;
            if (par) {
                const uint4 r3 = ld_const4<SM>(rp + 3);  // od_S, od_bits
                S = r3.x;
                bits = r3.y;
            }
            if (n == 0 || S == kChainNone) continue;
            n_tree_queries += P;
            const int nb = walk_blocks(n, par == 1, S, bits, blk, lane);
            __syncwarp();
            const uint32_t* ranks = A.rank_pool ? A.rank_pool + ((q0 * C2 + qc) * 2 + par) * A.rank_stride : nullptr;
            unsigned long long lbest[3] = {0, 0, 0};  // per piece: pack(ord, ~block index) of this lane's best block
            uint32_t lnode[3] = {0, 0, 0};
            int seen = 0;  // subtree blocks before the current group of 32
            for (int b0 = 0; b0 < nb && b0 < kMaxBlocks; b0 += 32) {
                const int b = b0 + lane;
                const bool mine = b < nb && b < kMaxBlocks;
                const WalkEntry we = mine ? blk[b] : WalkEntry{0, 0};
                const unsigned subm = __ballot_sync(kFull, mine && we.kind == 1);
                const int j = seen + __popc(subm & ((1u << lane) - 1));
                seen += __popc(subm);
                if (!mine) continue;
                if (we.kind == 0) {
                    const uint32_t off = ld_const<SM>(&A.or_off[ob + we.node]);
                    uint32_t v[3];
                    for (int k = 0; k < P; ++k) v[k] = ld_live<SM>(&A.or_ord[(int64_t)(2 * k + par) * A.n_entry + ob + we.node]);
                    if (off < offset) {
                        for (int k = 0; k < P; ++k)
                            if (v[k] > kOrdLowest && v[k] > (uint32_t)(lbest[k] >> 32)) {
                                lbest[k] = ((unsigned long long)v[k] << 32) | (~(uint32_t)b);
                                lnode[k] = we.node;
                            }
                    }
                } else {
                    const uint32_t ib = ld_const<SM>(&A.in_base[ob + we.node]);
                    const uint32_t cnt = (ranks && j < A.rank_stride) ? ld_const<SM>(&ranks[j])
                                                                     : count_less<SM>(A.in_off + ib, ld_const<SM>(&A.in_n[ob + we.node]), offset);
                    if (cnt) {  // Fenwick prefix maximum over the first cnt entries
                        unsigned long long r[3] = {0, 0, 0};
                        uint32_t c = cnt;
                        while (c) {  // the probe addresses only depend on cnt: issue six levels of probes before using any
                            uint32_t at[6];
#pragma unroll
                            for (int jj = 0; jj < 6; ++jj) {
                                at[jj] = c ? c - 1 : kChainNone;
                                c &= c - 1;  // 0 stays 0
                            }
                            unsigned long long x[6][3];
#pragma unroll
                            for (int jj = 0; jj < 6; ++jj)
#pragma unroll
                                for (int k = 0; k < 3; ++k)
                                    x[jj][k] = (k < P && at[jj] != kChainNone) ? ld_live<SM>(&A.bit[(int64_t)(2 * k + par) * A.n_inner + ib + at[jj]]) : 0ull;
#pragma unroll
                            for (int jj = 0; jj < 6; ++jj)
#pragma unroll
                                for (int k = 0; k < 3; ++k) r[k] = x[jj][k] > r[k] ? x[jj][k] : r[k];
                        }
                        for (int k = 0; k < P; ++k)
                            if (r[k] && (r[k] >> 32) > (lbest[k] >> 32)) {
                                lbest[k] = (r[k] & 0xffffffff00000000ull) | (~(uint32_t)b);
                                lnode[k] = (uint32_t)r[k];
                            }
                    }
                }
            }
            for (int k = 0; k < P; ++k) {  // first block, in walk order, that attains the maximum value
                unsigned long long r = lbest[k];
#pragma unroll
                for (int d = 16; d; d >>= 1) {
                    const unsigned long long o = __shfl_xor_sync(kFull, r, d);
                    r = o > r ? o : r;
                }
                if (!r) continue;
                const unsigned src = __ffs(__ballot_sync(kFull, lbest[k] == r)) - 1;
                const uint32_t node = __shfl_sync(kFull, lnode[k], src);
                if (lane == 0) {
                    const int t = 2 * k + par;
                    const double eq = __dmul_rn(A.gap_extend[k], (double)q);
                    const double pen = __dmul_rn(A.scale, par ? __dadd_rn(A.gap_open[k], eq) : __dsub_rn(A.gap_open[k], eq));
                    const float cand = __double2float_rn(__dsub_rn((double)__fadd_rn(funord((uint32_t)(r >> 32)), w), pen));
                    post(m, qc, (uint32_t)(1 + t), cand, ld_const<SM>(&A.or_match[ob + node]));
                }
            }
        }
        barrier();
    }
    if (pq0 != pq1) apply_winners(pq0, pq1, A.cand_best + ((A.n_step + 1) & 1) * A.n_match);  // the last step's winners
    if (lane == 0 && n_tree_queries) atomicAdd(A.counters, n_tree_queries);
}

__global__ void __launch_bounds__(kThreads, 1) chain_kernel(const ChainArgs A) {
    __shared__ WarpScratch scratch[kWarps];
    chain_steps<false>(A, scratch);
}

// Small problems: one CTA copies the whole arena into shared memory, rebases the pointers, prepares the queries,
// runs the same step loop on shared memory and copies the DP values and back-pointers out again.
__device__ __forceinline__ void chain_small_body(const ChainArgs& G, WarpScratch* scratch, uint4* arena_smem) {
    const char* gbase = G.arena_base;
    const int64_t ncopy = (G.copy_bytes + 15) / 16, nall = (G.arena_bytes + 15) / 16;
    for (int64_t i = threadIdx.x; i < ncopy; i += kThreads)
        arena_smem[i] = __ldg(reinterpret_cast<const uint4*>(gbase) + i);
    for (int64_t i = ncopy + threadIdx.x; i < nall; i += kThreads) arena_smem[i] = make_uint4(0u, 0u, 0u, 0u);
    ChainArgs A = G;
    char* sbase = reinterpret_cast<char*>(arena_smem);
#define CLB_REBASE(f) A.f = reinterpret_cast<decltype(A.f)>(sbase + (reinterpret_cast<const char*>(G.f) - gbase))
    CLB_REBASE(dp); CLB_REBASE(backptr); CLB_REBASE(sins_off); CLB_REBASE(ins); CLB_REBASE(qry_off); CLB_REBASE(qry_match);
    CLB_REBASE(qrec); CLB_REBASE(weight); CLB_REBASE(qry_chain1); CLB_REBASE(qa1); CLB_REBASE(qa2); CLB_REBASE(qoff);
    CLB_REBASE(pair_grp_off); CLB_REBASE(grp_shift); CLB_REBASE(grp_base); CLB_REBASE(grp_n); CLB_REBASE(pair_base);
    CLB_REBASE(gf_key); CLB_REBASE(gf_match); CLB_REBASE(gf_ord); CLB_REBASE(gf_best); CLB_REBASE(or_shift); CLB_REBASE(or_off);
    CLB_REBASE(or_match); CLB_REBASE(or_ord); CLB_REBASE(in_base); CLB_REBASE(in_n); CLB_REBASE(in_off); CLB_REBASE(bit);
    CLB_REBASE(ent_rank); CLB_REBASE(cand_best); CLB_REBASE(cand_bp); CLB_REBASE(counters);
    if (G.rank_pool) CLB_REBASE(rank_pool);
#undef CLB_REBASE
    __syncthreads();
    prepare_queries<true>(A, scratch[threadIdx.x >> 5].blk, threadIdx.x & 31, threadIdx.x >> 5, kWarps);
    __syncthreads();
    chain_steps<true>(A, scratch);
    __syncthreads();
    float* const odp = G.out_dp ? G.out_dp : G.dp;
    uint32_t* const obp = G.out_backptr ? G.out_backptr : G.backptr;
    for (int64_t m = threadIdx.x; m < G.n_match; m += kThreads) {
        odp[m] = A.dp[m];
        obp[m] = A.backptr[m];
    }
}

// Small problems: one CTA copies the whole arena into shared memory, rebases the pointers, prepares the queries,
// runs the same step loop on shared memory and copies the DP values and back-pointers out again.
__global__ void __launch_bounds__(kThreads, 1) chain_small_kernel(const ChainArgs G) {
    __shared__ WarpScratch scratch[kWarps];
    extern __shared__ uint4 arena_smem[];
    chain_small_body(G, scratch, arena_smem);
}

// The same for a batch of independent problems: CTA b solves problem b (clb_chain_dp_batch).
__global__ void __launch_bounds__(kThreads, 1) chain_small_batch_kernel(const ChainArgs* __restrict__ problems) {
    __shared__ WarpScratch scratch[kWarps];
    extern __shared__ uint4 arena_smem[];
    __shared__ ChainArgs G;
    if (threadIdx.x == 0) G = problems[blockIdx.x];
    __syncthreads();
    chain_small_body(G, scratch, arena_smem);
}

cudaError_t launch_chain_small_batch(const ChainArgs* d_args, int n, int smem_bytes, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e0 = cudaFuncSetAttribute(chain_small_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmallArena);
        if (e0 != cudaSuccess) return e0;
        attr_set = true;
    }
    chain_small_batch_kernel<<<n, kThreads, (size_t)smem_bytes, stream>>>(d_args);
    return cudaGetLastError();
}

cudaError_t launch_chain(const ChainArgs& args, int grid, int prepare_grid, cudaStream_t stream, cudaEvent_t after_prepare) {
    if (grid == 0) {  // the whole problem fits into shared memory
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e0 = cudaFuncSetAttribute(chain_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmallArena);
            if (e0 != cudaSuccess) return e0;
            attr_set = true;
        }
        if (after_prepare) cudaEventRecord(after_prepare, stream);
        chain_small_kernel<<<1, kThreads, (size_t)((args.arena_bytes + 15) / 16 * 16), stream>>>(args);
        return cudaGetLastError();
    }
    if (args.n_qry > 0) chain_prepare_kernel<<<prepare_grid, 256, 0, stream>>>(args);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (after_prepare) cudaEventRecord(after_prepare, stream);
    void* params[] = {(void*)&args};
    if (grid > 1)
        return cudaLaunchCooperativeKernel((const void*)chain_kernel, dim3(grid), dim3(kThreads), params, 0, stream);
    chain_kernel<<<1, kThreads, 0, stream>>>(args);
    return cudaGetLastError();
}

int chain_max_grid(int device) {
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chain_kernel, kThreads, 0) != cudaSuccess) return 1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 1;
    return per_sm > 0 ? sms : 1;
}

}  // namespace clb
