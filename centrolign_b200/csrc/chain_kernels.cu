// chain_kernels.cu -- sm_100a kernel of the sparse anchor-chaining DP (clb_chain_dp).
//
// Reference: Anchorer::sparse_affine_chain_dp main loop (include/centrolign/anchorer.hpp:2290-2417) and
// Anchorer::sparse_chain_dp main loop (:1640-1728); search structures max_search_tree.hpp:312-444 and
// orthogonal_max_search_tree.hpp:293-470.  Data layout and the equivalence argument: chain_device.cuh.
//
// One persistent cooperative kernel walks the graph-1 nodes ("steps") in topological order.  Per step:
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

struct WalkEntry {
    uint32_t node;  // outer heap index
    uint32_t kind;  // 0: the node itself, 1: the whole subtree of `node` (its inner list)
};

struct WarpScratch {
    WalkEntry blk[kMaxBlocks];
};

// MaxSearchTree::range_max(lo = (0, min), hi = (offset, min)) on one gap-free tree (max_search_tree.hpp:361-444).
// Keys (offset, match) < hi  <=>  offset key < offset; nothing is below lo.  Executed uniformly by the warp.
__device__ bool gapfree_range_max(const ChainArgs& A, int64_t base, uint32_t n, uint32_t offset, float& val_out, uint32_t& match_out) {
    uint32_t cursor = 0;
    while (cursor < n && __ldg(&A.gf_key[base + cursor]) >= offset) cursor = 2 * cursor + 1;
    if (cursor >= n) return false;
    float best = __ldcg(&A.gf_val[base + cursor]);
    uint32_t best_match = __ldg(&A.gf_match[base + cursor]);
    uint32_t lc = 2 * cursor + 1, rc = 2 * cursor + 2;
    while (lc < n) {  // every key on this side is >= lo: take the node and its whole right subtree
        const float v = __ldcg(&A.gf_val[base + lc]);
        if (v > best) {
            best = v;
            best_match = __ldg(&A.gf_match[base + lc]);
        }
        const uint32_t r = 2 * lc + 2;
        if (r < n) {
            const unsigned long long pk = __ldcg(&A.gf_best[base + r]);
            if (pk) {
                const float sv = funord((uint32_t)(pk >> 32));
                if (sv > best) {
                    best = sv;
                    best_match = __ldg(&A.ent_match[__ldg(&A.sins_entry[~(uint32_t)pk])]);
                }
            }
        }
        lc = 2 * lc + 1;
    }
    while (rc < n) {
        if (__ldg(&A.gf_key[base + rc]) < offset) {
            const float v = __ldcg(&A.gf_val[base + rc]);
            if (v > best) {
                best = v;
                best_match = __ldg(&A.gf_match[base + rc]);
            }
            const uint32_t l = 2 * rc + 1;
            if (l < n) {
                const unsigned long long pk = __ldcg(&A.gf_best[base + l]);
                if (pk) {
                    const float sv = funord((uint32_t)(pk >> 32));
                    if (sv > best) {
                        best = sv;
                        best_match = __ldg(&A.ent_match[__ldg(&A.sins_entry[~(uint32_t)pk])]);
                    }
                }
            }
            rc = 2 * rc + 2;
        } else {
            rc = 2 * rc + 1;
        }
    }
    val_out = best;
    match_out = best_match;
    return true;
}

// The outer walk of OrthogonalMaxSearchTree::range_max (orthogonal_max_search_tree.hpp:340-470) for
//   prefix: key1 in [(-inf, min), (q, min))  <=> shift <  q   (odd pieces,  anchorer.hpp:2396-2398)
//   suffix: key1 in [(q+1, min), (+inf, max)) <=> shift >  q   (even pieces, anchorer.hpp:2405-2407)
// Records the blocks in the order the reference tests them.  Uniform over the warp; lane 0 writes.
__device__ int ortho_walk(const ChainArgs& A, int64_t ob, uint32_t n, int q, bool prefix, WalkEntry* blk, int lane) {
    auto in_range = [&](uint32_t x) {
        const int s = __ldg(&A.or_shift[ob + x]);
        return prefix ? (s < q) : (s > q);
    };
    uint32_t cursor = 0;
    while (cursor < n && !in_range(cursor)) cursor = prefix ? 2 * cursor + 1 : 2 * cursor + 2;
    if (cursor >= n) return 0;
    int nb = 0;
    auto push = [&](uint32_t node, uint32_t kind) {
        if (lane == 0 && nb < kMaxBlocks) blk[nb] = WalkEntry{node, kind};
        ++nb;
    };
    push(cursor, 0);
    uint32_t lc = 2 * cursor + 1, rc = 2 * cursor + 2;
    while (lc < n) {  // leftward: right subtrees hang entirely inside the key-1 range
        if (prefix || in_range(lc)) {
            push(lc, 0);
            if (2 * lc + 2 < n) push(2 * lc + 2, 1);
            lc = 2 * lc + 1;
        } else {
            lc = 2 * lc + 2;
        }
    }
    while (rc < n) {  // rightward: left subtrees hang entirely inside
        if (!prefix || in_range(rc)) {
            push(rc, 0);
            if (2 * rc + 1 < n) push(2 * rc + 1, 1);
            rc = 2 * rc + 2;
        } else {
            rc = 2 * rc + 1;
        }
    }
    return nb;
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1) chain_kernel(const ChainArgs A) {
    __shared__ WarpScratch scratch[kWarps];
    cg::grid_group grid = cg::this_grid();
    const bool multi = gridDim.x > 1;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t gwarp = (int64_t)blockIdx.x * kWarps + wib, nwarp = (int64_t)gridDim.x * kWarps;
    const int64_t gthread = (int64_t)blockIdx.x * kThreads + threadIdx.x, nthread = (int64_t)gridDim.x * kThreads;
    const int C1 = A.n_chain1, C2 = A.n_chain2, P = A.num_pw, T = 2 * A.num_pw;
    WalkEntry* blk = scratch[wib].blk;
    unsigned long long n_tree_queries = 0;

    auto barrier = [&]() {
        if (multi) grid.sync();
        else __syncthreads();
    };

    for (int64_t s = 0; s < A.n_step; ++s) {
        // ------------------------------ A: inserts (anchorer.hpp:2301-2345) ------------------------------
        const int64_t i0 = A.sins_off[s], i1 = A.sins_off[s + 1];
        for (int64_t i = i0 + gwarp; i < i1; i += nwarp) {
            const uint32_t e = A.sins_entry[i];
            const uint32_t m = A.ent_match[e];
            const float dpv = __ldcg(&A.dp[m]);
            if (!(dpv > mininf())) continue;  // entering lowest() changes nothing in the reference's trees
            {   // gap-free tree of the entry's diagonal: lanes = the node and its ancestors
                const uint32_t g = A.ent_gf_grp[e], x = A.ent_gf_node[e];
                const int64_t base = A.grp_base[g];
                if (lane == 0) A.gf_val[base + x] = dpv;
                const uint32_t anc = (x + 1) >> lane;
                if (anc) atomicMax(&A.gf_best[base + anc - 1], pack(dpv, ~(uint32_t)i));
            }
            if (P > 0) {
                const uint32_t pr = A.ent_pair[e], oh = A.ent_or_node[e];
                const int64_t ob = A.pair_base[pr];
                const int shift = A.ent_shift[e];
                const int64_t r0 = A.ent_rank_off[e];
                const int nr = (int)(A.ent_rank_off[e + 1] - r0);
                for (int t = 0; t < T; ++t) {
                    // anchorer.hpp:2328-2335: odd pieces add, even pieces subtract local_scale * gap_extend * shift
                    const double gap = __dmul_rn(A.scale_ext[t >> 1], (double)shift);
                    const float v = __double2float_rn((t & 1) ? __dadd_rn((double)dpv, gap) : __dsub_rn((double)dpv, gap));
                    if (!(v > mininf())) continue;  // anchorer.hpp:2338
                    if (lane == 0) A.or_val[(int64_t)t * A.n_entry + ob + oh] = v;
                    if (lane < nr) {
                        const uint32_t a = ((oh + 1) >> lane) - 1;
                        const int64_t ib = A.in_base[ob + a];
                        const uint32_t n = A.in_n[ob + a];
                        const unsigned long long pk = pack(v, oh);
                        unsigned long long* bit = A.bit + (int64_t)t * A.n_inner + ib;
                        for (uint32_t k = A.ent_rank[r0 + lane]; k < n; k |= k + 1) atomicMax(&bit[k], pk);
                    }
                }
            }
        }
        const int64_t q0 = A.qry_off[s], q1 = A.qry_off[s + 1];
        if (q0 == q1) {
            if (i0 != i1) barrier();
            continue;
        }
        if (i0 != i1) barrier();

        // ------------------------------ B: queries (anchorer.hpp:2352-2416) ------------------------------
        const int64_t n_items = (q1 - q0) * C2;
        for (int64_t item = gwarp; item < n_items; item += nwarp) {
            const int64_t qi = item / C2;
            const int c2 = (int)(item - qi * C2);
            const uint32_t m = A.qry_match[q0 + qi];
            const uint32_t c1 = A.qry_chain1[q0 + qi];
            const uint32_t offset = A.qoff[(int64_t)m * C2 + c2];
            if (offset == 0) continue;  // nothing on this path reaches the match: every range [0, 0) is empty
            const int q = (int)((uint32_t)A.qa1[(int64_t)m * C1 + c1] - (uint32_t)A.qa2[(int64_t)m * C2 + c2]);
            const int64_t pair = (int64_t)c1 * C2 + c2;
            const float w = A.weight[m];
            float best = mininf();
            uint32_t best_bp = 0xffffffffu;
            {   // same diagonal (anchorer.hpp:2379-2389): binary search the pair's diagonals for shift == q
                int64_t lo = A.pair_grp_off[pair], hi = A.pair_grp_off[pair + 1];
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (__ldg(&A.grp_shift[mid]) < q) lo = mid + 1;
                    else hi = mid;
                }
                if (lo < A.pair_grp_off[pair + 1] && __ldg(&A.grp_shift[lo]) == q) {
                    float v;
                    uint32_t bm;
                    ++n_tree_queries;
                    if (gapfree_range_max(A, A.grp_base[lo], A.grp_n[lo], offset, v, bm) && v > mininf()) {
                        const float cand = __fadd_rn(v, w);
                        if (cand > best) {
                            best = cand;
                            best_bp = bm;
                        }
                    }
                }
            }
            if (P > 0) {
                const int64_t ob = A.pair_base[pair];
                const uint32_t n = (uint32_t)(A.pair_base[pair + 1] - ob);
                float tv[kChainMaxTrees];
                uint32_t tn[kChainMaxTrees];
                for (int t = 0; t < kChainMaxTrees; ++t) {
                    tv[t] = mininf();
                    tn[t] = 0;
                }
                for (int par = 0; par < 2 && n > 0; ++par) {  // par 0: even pieces (suffix), par 1: odd pieces (prefix)
                    __syncwarp();
                    const int nb = ortho_walk(A, ob, n, q, par == 1, blk, lane);
                    __syncwarp();
                    unsigned long long lbest[3] = {0, 0, 0};  // per piece: pack(value, ~block index) of this lane's best block
                    uint32_t lnode[3] = {0, 0, 0};
                    for (int b = lane; b < nb && b < kMaxBlocks; b += 32) {
                        const WalkEntry we = blk[b];
                        if (we.kind == 0) {
                            if (__ldg(&A.or_off[ob + we.node]) < offset) {
                                for (int k = 0; k < P; ++k) {
                                    const float v = __ldcg(&A.or_val[(int64_t)(2 * k + par) * A.n_entry + ob + we.node]);
                                    const unsigned long long pk = pack(v, ~(uint32_t)b);
                                    if (v > mininf() && (pk >> 32) > (lbest[k] >> 32)) {
                                        lbest[k] = pk;
                                        lnode[k] = we.node;
                                    }
                                }
                            }
                        } else {
                            const int64_t ib = A.in_base[ob + we.node];
                            const uint32_t cn = A.in_n[ob + we.node];
                            uint32_t lo = 0, hi = cn;  // elements of the subtree with offset < `offset`
                            while (lo < hi) {
                                const uint32_t mid = (lo + hi) >> 1;
                                if (__ldg(&A.in_off[ib + mid]) < offset) lo = mid + 1;
                                else hi = mid;
                            }
                            if (lo) {
                                for (int k = 0; k < P; ++k) {
                                    const unsigned long long* bit = A.bit + (int64_t)(2 * k + par) * A.n_inner + ib;
                                    unsigned long long r = 0;
                                    for (uint32_t c = lo; c > 0; c &= c - 1) {
                                        const unsigned long long x = __ldcg(&bit[c - 1]);
                                        r = x > r ? x : r;
                                    }
                                    if (r && (r >> 32) > (lbest[k] >> 32)) {
                                        lbest[k] = (r & 0xffffffff00000000ull) | (~(uint32_t)b);
                                        lnode[k] = (uint32_t)r;
                                    }
                                }
                            }
                        }
                    }
                    n_tree_queries += P;
                    for (int k = 0; k < P; ++k) {  // first block, in walk order, that attains the maximum value
                        unsigned long long r = lbest[k];
#pragma unroll
                        for (int d = 16; d; d >>= 1) {
                            const unsigned long long o = __shfl_xor_sync(kFull, r, d);
                            r = o > r ? o : r;
                        }
                        if (r) {
                            const unsigned src = __ffs(__ballot_sync(kFull, lbest[k] == r)) - 1;
                            tv[2 * k + par] = funord((uint32_t)(r >> 32));
                            tn[2 * k + par] = __shfl_sync(kFull, lnode[k], src);
                        } else {
                            __ballot_sync(kFull, false);
                        }
                    }
                }
                for (int t = 0; t < T; ++t) {  // pieces in the reference's order (anchorer.hpp:2390-2413)
                    if (!(tv[t] > mininf())) continue;
                    const int k = t >> 1;
                    const double eq = __dmul_rn(A.gap_extend[k], (double)q);
                    const double pen = __dmul_rn(A.scale, (t & 1) ? __dadd_rn(A.gap_open[k], eq) : __dsub_rn(A.gap_open[k], eq));
                    const float cand = __double2float_rn(__dsub_rn((double)__fadd_rn(tv[t], w), pen));
                    if (cand > best) {
                        best = cand;
                        best_bp = __ldg(&A.or_match[ob + tn[t]]);
                    }
                }
            }
            if (lane == 0 && best_bp != 0xffffffffu) {
                const uint32_t order = (uint32_t)(qi * C2 + c2);
                A.cand_bp[order] = best_bp;
                atomicMax(&A.cand_best[m], pack(best, ~order));
            }
        }
        barrier();

        // ------------------------------ C: update_dp (match_bank.hpp:171-184) ------------------------------
        for (int64_t qi = gthread; qi < q1 - q0; qi += nthread) {
            const uint32_t m = A.qry_match[q0 + qi];
            const unsigned long long pk = __ldcg(&A.cand_best[m]);
            if (!pk) continue;
            const uint32_t order = ~(uint32_t)pk;
            if ((int64_t)(order / (uint32_t)C2) != qi) continue;  // the winner is posted by another query of this match
            const float v = funord((uint32_t)(pk >> 32));
            if (v > __ldcg(&A.dp[m])) {
                A.dp[m] = v;
                A.backptr[m] = __ldcg(&A.cand_bp[order]);
            }
            A.cand_best[m] = 0;
        }
        barrier();
    }
    if (lane == 0 && n_tree_queries) atomicAdd(A.counters, n_tree_queries);
}

cudaError_t launch_chain(const ChainArgs& args, int grid, cudaStream_t stream) {
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
