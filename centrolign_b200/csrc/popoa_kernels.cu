// popoa_kernels.cu -- hand-written sm_100a kernels for the piecewise-affine PO-to-PO graph DP.
//
// What is computed (pull form of the reference's push loop, include/centrolign/alignment.hpp:898-938):
//   s(i,j)   = +match if label1(i)==label2(j) else -mismatch
//   I_k(i,j) = max_{p in pred1(i)} max( M(p,j) - (o_k+e_k), I_k(p,j) - e_k )
//   D_k(i,j) = max_{q in pred2(j)} max( M(i,q) - (o_k+e_k), D_k(i,q) - e_k )
//   M(i,j)   = max( max_{p,q} M(p,q) + s(i,j), I_0..I_{P-1}, D_0..D_{P-1} )
// Integer max distributes over "+const", so for a node with several predecessors the kernel
// first folds the predecessor rows (columns) into ONE effective row (column) by an
// element-wise max and then applies the single-predecessor update: identical values, and the
// common case (one predecessor = previous index) never leaves registers.
//
// Execution model.  One CTA per window at a time (persistent CTAs pull windows, largest first, from
// an atomic queue): 11 fill warps + 1 traceback warp.  The matrix is cut into strips of 128 columns
// (32 for narrow windows); a warp sweeps a strip top to bottom as a skewed wavefront (lane t is on
// row r-t, 4 columns per lane), the row above lives in registers, neighbour columns arrive by warp
// shuffles, and the warps pipeline over consecutive strips -- the hand-off being the last column of
// a strip, published through the window workspace with a fence + progress word in shared memory and
// prefetched by the consumer with cp.async.  DPX instructions (__viaddmax_s32, __vimax3_s32) carry
// the three-piece affine recurrences.  No tensor cores: this is integer max-plus, not a GEMM.
// Wide windows with many rows are filled as (row panel, strip) tiles from a per-window queue, so that a
// slow strip does not hold back the strips behind it (see "Tiled windows" below), and strips without rare
// predecessor columns run a lean copy of the step (fill_strip_wide).
//
// Traceback does not store the matrix.  The fill keeps only "persisted" rows {M,I_k} and
// columns {M,D_k}: every 64th row / 32nd column plus any row / column that has a far successor.
// The traceback warp re-computes the 64x32 tile the path is in (tile_strip, all cells kept in
// shared memory) and applies the reference's own tests (alignment.hpp:1036-1138) to the real cell
// values, so tie-breaking is the reference's by construction.  It runs concurrently with the fill
// of the next window (two workspace slots per CTA).
#include <cuda_runtime.h>
#include <limits.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "popoa_device.cuh"

namespace clb {

#ifndef CLB_WARPS
#define CLB_WARPS 12
#endif
constexpr int kWarps = CLB_WARPS;  // fill warps + 1 traceback warp per CTA (one CTA per SM)
constexpr int kThreads = kWarps * 32;
constexpr unsigned kFull = 0xffffffffu;

// -DCLB_PROFILE builds: cycles the warps spend in their waits, by kind (printed by the host when CLB_WAIT_PROFILE is set)
#ifdef CLB_PROFILE
__device__ unsigned long long g_wait_cycles[8];  // 0 left-strip rows, 1 panel above, 2 next window, 3 traceback waits for the fill, 4 fill total, 5 traceback busy
#define CLB_WAIT_BEGIN const long long _wt0 = clock64()
#define CLB_WAIT_END(kind) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_wait_cycles[kind], (unsigned long long)(clock64() - _wt0)); } while (0)
#else
#define CLB_WAIT_BEGIN do {} while (0)
#define CLB_WAIT_END(kind) do {} while (0)
#endif

// window view, resolved once per window into shared memory
struct Win {
    int n1, n2;
    int nsnk1, nsnk2;
    const uint32_t *info1, *info2;
    const int32_t *slot1, *slot2;
    const uint32_t *depth1, *depth2;
    const uint32_t *poff1, *poff2;
    const uint32_t *pidx1, *pidx2;
    const uint32_t *snk1, *snk2;
    int4 *rowbuf, *colbuf, *brow, *bcol;
    int* coleff;  // per persisted column and row r: max over pred1(r) of M(p, column) -- the diagonal input of its right neighbours
    int64_t out;
    int id;
    int cw;  // columns per lane of the fill (1: 32-column strips, 4: 128-column strips)
};

__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }

__device__ __forceinline__ void max4(int& m, int (&v)[3], const int4 a) {
    m = imax(m, a.x);
    v[0] = imax(v[0], a.y);
    v[1] = imax(v[1], a.z);
    v[2] = imax(v[2], a.w);
}

// What a cell with gap state s (I_k or D_k) and value m offers its successors in piece k: max(s - e_k, m - oe_k).
__device__ __forceinline__ int offer(const Params& prm, int s, int m, int k) { return __viaddmax_s32(s, -prm.e[k], m - prm.oe[k]); }
template <int P>
__device__ __forceinline__ int4 to_offered(const Params& prm, const int4 v) {  // {M, I_k} -> {M, H_k} (or {M, D_k} -> {M, G_k})
    int4 o = make_int4(v.x, kMinInf, kMinInf, kMinInf);
    o.y = offer(prm, v.y, v.x, 0);
    if (P > 1) o.z = offer(prm, v.z, v.x, 1);
    if (P > 2) o.w = offer(prm, v.w, v.x, 2);
    return o;
}

// ------------------------------------------------------------------------------------------
// Asynchronous global->shared copies (LDGSTS) for the left-column prefetch.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ int4 lds_128(uint32_t saddr) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ int lds_32(uint32_t saddr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts_128(uint32_t saddr, const int4 v) {
    asm volatile("st.shared.v4.s32 [%0], {%1,%2,%3,%4};\n" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// per-fill-warp shared memory
// (the start lag between neighbouring strips is a launch parameter, LaunchArgs::start_lag; 64 rows measured best)
constexpr unsigned kPollNs = 200;  // poll interval of the strip hand-off waits (800 ns measured the same)
constexpr int kProgMask = 511;  // progress words per CTA (one per strip in flight; a tiled window keeps all its strips live)
constexpr int kFillRing = 4;    // ring rows: a row is read at most kNear steps after it was written
struct __align__(16) FillSmem {
    int4 ringA[kFillRing * 32];  // {M, I_k} of the last rows, own column
    int4 ringB[kFillRing * 32];  // {diagonal input, D_k} of the current row, for the columns to the right
    int4 leftv[3][2][32];        // prefetched {M, D_k} of the up-to-3 columns just left of the strip, 32-row blocks
    int lefte[3][2][32];         // prefetched diagonal input of those columns
};

template <int P>
__device__ __forceinline__ int4 boundary_cell(uint32_t depth, const Params& prm);

// ------------------------------------------------------------------------------------------
// DP fill of one strip, rows 1..n1 -- the hot loop.  Same recurrence as process_strip, tuned:
//  * predecessors at distance 1..kNear (SNP-sized bubbles, the common irregular case) are named
//    by bits of the node's info word and served branch-light from registers (distance 1) or the
//    shared-memory ring; only genuinely far predecessors (long bubbles, deletion edges, the
//    boundary) walk the predecessor list;
//  * the diagonal input of a column's right neighbours is kept explicitly (ringB.x / coleff);
//  * the columns just left of the strip (written by another warp) are prefetched 32 rows at a
//    time with cp.async, so no global-memory latency sits on the per-row critical path and the
//    producer's progress word is polled once per 32 rows;
//  * 32-bit element offsets into the window workspace, workspace slot packed in the info word.
// `g` is the CTA-wide running strip number (progress tag), `cs` the strip index in the window.
// ------------------------------------------------------------------------------------------
template <int P>
__device__ __forceinline__ void fill_strip(const Win& Wsh, const Params& prm, const int cs, const int g, FillSmem& sm,
                                           volatile unsigned long long* progress, const int lane, const int start_lag) {
    constexpr int H = kFillRing;
    const int C0 = 1 + kStrip * cs;
    const int n1 = Wsh.n1, n2 = Wsh.n2;
    const uint32_t* __restrict__ info1 = Wsh.info1;
    const int32_t* __restrict__ slot1 = Wsh.slot1;
    const uint32_t* __restrict__ poff1 = Wsh.poff1;
    const uint32_t* __restrict__ pidx1 = Wsh.pidx1;
    const int32_t* __restrict__ slot2 = Wsh.slot2;
    const uint32_t* __restrict__ pidx2 = Wsh.pidx2;
    int4* const rowbuf = Wsh.rowbuf;
    int4* const colbuf = Wsh.colbuf;
    int* const coleff = Wsh.coleff;
    const uint32_t rstride = row_stride((uint32_t)n2), cstride = (uint32_t)n1 + 1u;  // host guarantees 32-bit offsets

    const int j = C0 + lane;
    const bool jvalid = j <= n2;
    const uint32_t cinfo = jvalid ? Wsh.info2[j] : (kInfoRegular | (1u << kInfoNearShift) | 0xffu);
    const int clabel = (int)(cinfo & kInfoLabelMask);
    const bool creg = (cinfo & kInfoRegular) != 0;
    const uint32_t cmask = (cinfo >> kInfoNearShift) & 7u;
    const bool cfar = (cinfo & kInfoFar) != 0;
    const bool colpath = !creg || lane == 0;  // lanes whose left state does not simply come from the shuffle
    const uint32_t cp0 = (jvalid && cfar) ? Wsh.poff2[j] : 0u, cp1 = (jvalid && cfar) ? Wsh.poff2[j + 1] : 0u;
    const bool hascol = jvalid && (cinfo & kInfoPersist);
    const uint32_t myoff = hascol ? (uint32_t)slot2[j] * cstride : 0u;
    const uint32_t persist_mask = jvalid ? kInfoPersist : 0u;  // lanes beyond n2 never store rows

    // ---- boundary data owned by this strip (alignment.hpp:814-894, see boundary_cell) ----
    int upM = kMinInf, upI[3] = {kMinInf, kMinInf, kMinInf};
    if (jvalid) {
        upM = boundary_cell<P>(Wsh.depth2[j], prm).x;  // M(0,j); I_k(0,j) = -inf
        rowbuf[j] = to_offered<P>(prm, make_int4(upM, kMinInf, kMinInf, kMinInf));  // persisted rows hold {M, H_k}
    }
    if (cs == 0) {  // boundary column as a predecessor column: {M(i,0), -inf} and its diagonal input
        for (int i = lane; i <= n1; i += 32) {
            const int m = i == 0 ? 0 : boundary_cell<P>(Wsh.depth1[i], prm).x;  // M(0,0)=0 seeds the sources
            colbuf[i] = make_int4(m, kMinInf, kMinInf, kMinInf);
            int e = kMinInf;
            const uint32_t a1 = poff1[i + 1];
            for (uint32_t a = poff1[i]; a < a1; ++a) {
                const uint32_t p = pidx1[a];
                e = imax(e, p == 0 ? 0 : boundary_cell<P>(Wsh.depth1[p], prm).x);
            }
            coleff[i] = e;
        }
    }
    // ---- which of the three columns left of the strip are needed (warp-uniform) ----
    // lane t with near bit d' (> t) reads left column d = d' - t, i.e. matrix column C0 - d
    unsigned needbits = 0;
    if (jvalid) {
#pragma unroll
        for (int dd = 1; dd <= 3; ++dd)
            if (dd > lane && ((cmask >> (dd - 1)) & 1u)) needbits |= 1u << (dd - lane - 1);
        if (lane == 0 && creg) needbits |= 1u;  // column 1's only predecessor is the boundary column
    }
    needbits = __reduce_or_sync(kFull, needbits);
    uint32_t xoffw[3] = {0, 0, 0};
#pragma unroll
    for (int d = 1; d <= 3; ++d)
        if (needbits & (1u << (d - 1))) xoffw[d - 1] = (uint32_t)slot2[C0 - d] * cstride;
    __syncwarp();

    int avail = cs == 0 ? INT_MAX : 0;
    auto wait_rows = [&](int want) {
        if (avail < want) {
            for (;;) {
                const unsigned long long v = progress[(g - 1) & kProgMask];
                avail = ((int)(v >> 32) == g) ? (int)(v & 0xffffffffu) : 0;
                if (avail >= want) break;
                __nanosleep(kPollNs);
            }
            __threadfence_block();
        }
    };
    auto prefetch_block = [&](int b) {  // rows 32b+1 .. 32b+32 of the needed left columns -> buffer b&1
        const int row = 32 * b + 1 + lane;
        if (row <= n1) {
#pragma unroll
            for (int d = 0; d < 3; ++d)
                if (needbits & (1u << d)) {
                    cp_async_16(&sm.leftv[d][b & 1][lane], colbuf + (xoffw[d] + (uint32_t)row));
                    cp_async_4(&sm.lefte[d][b & 1][lane], coleff + (xoffw[d] + (uint32_t)row));
                }
        }
        cp_async_commit();
    };
    // Start well behind the producer strip: the lag set here persists (all strips advance at the same
    // rate), and the slack absorbs scheduling jitter so the per-block waits below rarely spin.
    wait_rows(min(max(start_lag, 32), n1));
    prefetch_block(0);
    cp_async_wait_all();
    __syncwarp();

    int outM = kMinInf, outD[3] = {kMinInf, kMinInf, kMinInf}, outEff = kMinInf;
    uint32_t rinfo_next = (lane == 0) ? info1[1] : 0u;  // software-pipelined row info
    int4* const myA = sm.ringA + lane;
    int4* const myB = sm.ringB + lane;

    // one wavefront step; GUARD = some lanes may be outside rows 1..n1 (pipeline fill / drain)
    auto step = [&](const int s, auto guard_tag) {
        constexpr bool GUARD = decltype(guard_tag)::value;
        const int r = 1 + s - lane;
        bool act = true;
        if (GUARD) act = r >= 1 && r <= n1;
        int lM = __shfl_up_sync(kFull, outM, 1);
        int lD[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) lD[k] = (k < P) ? __shfl_up_sync(kFull, outD[k], 1) : kMinInf;
        int lEff = __shfl_up_sync(kFull, outEff, 1);
        const uint32_t rinfo = rinfo_next;
        if (GUARD) rinfo_next = (r >= 0 && r < n1) ? info1[r + 1] : 0u;
        else rinfo_next = info1[r + 1];  // r+1 <= n1+1: the info array is padded by one entry
        if (act) {
            const int rs = (r & (H - 1)) * 32;
            // ---- effective predecessor row ----
            int eM = upM, eI[3] = {upI[0], upI[1], upI[2]};
            int fH[3] = {kMinInf, kMinInf, kMinInf};  // offered by far predecessor rows (the workspace holds {M, H_k})
            if (!(rinfo & kInfoRegular)) {
                if (!(rinfo & (1u << kInfoNearShift))) { eM = kMinInf; eI[0] = eI[1] = eI[2] = kMinInf; }
                if (rinfo & (2u << kInfoNearShift)) max4(eM, eI, myA[rs ^ 64]);  // row r-2
                if (rinfo & ((4u << kInfoNearShift) | kInfoFar)) {
                    if (rinfo & (4u << kInfoNearShift)) max4(eM, eI, myA[((r - 3) & (H - 1)) * 32]);
                    if (rinfo & kInfoFar) {
                        const uint32_t rp1 = poff1[r + 1];
                        for (uint32_t a = poff1[r]; a < rp1; ++a) {
                            const int p = (int)pidx1[a];
                            if (p >= 1 && r - p <= kNear) continue;
                            max4(eM, fH, rowbuf[(uint32_t)slot1[p] * rstride + (uint32_t)j]);
                        }
                    }
                }
            }
            // ---- effective predecessor column + diagonal input ----
            if (colpath) {
                const int lb = ((r - 1) >> 5) & 1, li = (r - 1) & 31;  // prefetch buffer / index of row r
                if (!(cmask & 1u) && !creg) { lM = kMinInf; lD[0] = lD[1] = lD[2] = kMinInf; lEff = kMinInf; }
                else if (lane == 0) {
                    const int4 b = sm.leftv[0][lb][li];
                    lM = b.x; lD[0] = b.y; lD[1] = b.z; lD[2] = b.w;
                    lEff = sm.lefte[0][lb][li];
                }
                if (cmask & 6u) {
#pragma unroll
                    for (int d = 2; d <= 3; ++d) {
                        if (cmask & (1u << (d - 1))) {
                            int4 v;
                            int m;
                            if (lane >= d) {
                                v = sm.ringB[rs + lane - d];           // {diag input, D_k}
                                m = sm.ringA[rs + lane - d].x;          // M
                            } else {
                                const int4 t = sm.leftv[d - lane - 1][lb][li];  // {M, D_k}
                                m = t.x;
                                v = make_int4(sm.lefte[d - lane - 1][lb][li], t.y, t.z, t.w);
                            }
                            lM = imax(lM, m);
                            max4(lEff, lD, v);
                        }
                    }
                }
                if (cfar) {
                    for (uint32_t b = cp0; b < cp1; ++b) {
                        const int q = (int)pidx2[b];
                        if (q >= 1 && j - q <= kNear) continue;
                        const uint32_t o = (uint32_t)slot2[q] * cstride + (uint32_t)r;
                        max4(lM, lD, colbuf[o]);
                        lEff = imax(lEff, coleff[o]);
                    }
                }
            }
            // ---- the cell ----
            const int sub = ((int)(rinfo & kInfoLabelMask) == clabel) ? prm.match : -prm.mismatch;
            int I[3] = {kMinInf, kMinInf, kMinInf}, D[3] = {kMinInf, kMinInf, kMinInf};
            int M = __viaddmax_s32(lEff, sub, kMinInf);
#pragma unroll
            for (int k = 0; k < P; ++k) {
                I[k] = imax(__viaddmax_s32(eI[k], -prm.e[k], eM - prm.oe[k]), fH[k]);
                D[k] = __viaddmax_s32(lD[k], -prm.e[k], lM - prm.oe[k]);
                M = __vimax3_s32(M, I[k], D[k]);
            }
            const int4 cellA = make_int4(M, I[0], I[1], I[2]);
            myA[rs] = cellA;
            myB[rs] = make_int4(eM, D[0], D[1], D[2]);
            if (rinfo & persist_mask) {
                uint32_t slot = rinfo >> kInfoSlotShift;
                if (slot == kInfoSlotEscape) slot = (uint32_t)slot1[r];
                rowbuf[slot * rstride + (uint32_t)j] = to_offered<P>(prm, cellA);
            }
            if (hascol) {
                colbuf[myoff + (uint32_t)r] = make_int4(M, D[0], D[1], D[2]);
                coleff[myoff + (uint32_t)r] = eM;
            }
            upM = M; upI[0] = I[0]; upI[1] = I[1]; upI[2] = I[2];
            outM = M; outD[0] = D[0]; outD[1] = D[1]; outD[2] = D[2];
            outEff = eM;
        }
        __syncwarp();
    };
    auto publish = [&](int r31) {  // rows 1..r31 of this strip are complete (called by lane 31)
        __threadfence_block();
        progress[g & kProgMask] = ((unsigned long long)(g + 1) << 32) | (unsigned)r31;
    };

    const int nsteps = n1 + 31;
    for (int s0 = 0; s0 < nsteps; s0 += 32) {  // 32-step blocks; lane 0 is on rows s0+1 .. s0+32
        const int s1 = min(s0 + 32, nsteps);
        const bool inner = s0 >= 31 && s1 <= n1;  // every lane is inside rows 1..n1 for the whole block
#pragma unroll 1
        for (int q8 = s0; q8 < s1; q8 += 8) {
            const int e8 = min(q8 + 8, s1);
            if (inner) {
#pragma unroll 1
                for (int s = q8; s < e8; ++s) step(s, std::false_type{});
            } else {
#pragma unroll 1
                for (int s = q8; s < e8; ++s) step(s, std::true_type{});
            }
            if (q8 == s0) {  // lanes 1,2 have left block b-1 by now: refill its buffer with block b+1
                const int b = (s0 >> 5) + 1;
                if (32 * b + 1 <= n1) {
                    wait_rows(min(32 * b + 32, n1));
                    prefetch_block(b);
                }
            }
            if (lane == 31) {
                const int r31 = e8 - 31;  // last row lane 31 finished
                if (r31 >= 1) publish(min(r31, n1));
            }
        }
        cp_async_wait_all();  // next block's left columns have landed
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// DP fill, wide variant: strips of 128 columns, FOUR consecutive columns per lane, PUSH form.
//
// What a cell keeps for its successors is not {M, I_k, D_k} but what it OFFERS them:
//     T_k = M - oe_k,   H_k = max(I_k - e_k, T_k)  (to the rows below),   G_k = max(D_k - e_k, T_k)  (to the columns right)
// so that I_k(i,j) = max over predecessor rows of H_k and D_k(i,j) = max over predecessor columns of G_k: for the common
// single-predecessor cell the gap states cost nothing, the three subtractions are shared by both directions, and all
// seven values of the recurrence are still exactly the reference's (integer max/add; alignment.hpp:898-938 is itself a
// push loop).  Per lane the row above lives in registers {M, H_k} x 4 columns and is updated IN PLACE; the values a
// column offers to its right neighbours stay in registers inside the lane and cross lanes by shfl.up (lane-1 finished the
// same row one step earlier).  Which neighbour columns feed a column (distance 1 / 2: SNP-sized bubbles) is lane-constant,
// so it is not branched or predicated on but multiplied in: v*m + b with (m,b) = (1,0) or (0,-inf) -- IMADs, which run on
// the FMA pipe next to the DPX instructions on the ALU pipe (measured: profiles/r02_pipe_rates.txt; each pipe issues 2
// warp instructions per clock and SM, 3 together).  Rows with other predecessors than the row above fold the ring rows
// into the registers under one divergent branch; persisted rows store {I_k} before and {M} after the cell update.
// ------------------------------------------------------------------------------------------
constexpr int kWideCols = 4;
struct __align__(16) FillSmemWide {
    int4 ringA[kFillRing * kWideCols * 32];  // [row & 3][c][lane]: {M, H_k} of the last rows, own columns
    int4 leftv[3][2][16];                    // the 3 columns left of the strip, 16-row blocks: lands as {M, D_k}, converted to {G_k, E}
    int lefte[3][2][16];                     // landing zone of their diagonal input E
};

template <int P>
__device__ __forceinline__ void fill_strip_wide(const Win& Wsh, const Params& prm, const int cs, const int g,
                                                FillSmemWide& sm, volatile unsigned long long* progress, const int lane_in,
                                                const int start_lag, const int R0, const int R1, const int dbg) {
    // Rows R0+1 .. R1 of the strip (a panel; R0 = 0, R1 = n1 is the whole strip).  A panel that does not start at
    // the top takes over from whoever filled the panel above through the window workspace: rows R0-2 .. R0 are
    // persisted (host: popoa_host.cu) and the strip's own progress word says when they are there.
    constexpr int H = kFillRing, C = kWideCols, W = 32 * C, PB = 16;  // PB = rows per left-column prefetch block
    int lane = lane_in;
    asm volatile("mov.u32 %0, %0;" : "+r"(lane));  // a register, not a re-read of %tid in the row loop
    const int C0 = 1 + W * cs;
    const int n1 = Wsh.n1, n2 = Wsh.n2;
    const uint32_t* __restrict__ info1 = Wsh.info1;
    const uint32_t* __restrict__ poff1 = Wsh.poff1;
    const uint32_t* __restrict__ pidx1 = Wsh.pidx1;
    const int32_t* __restrict__ slot2 = Wsh.slot2;
    int4* const rowbuf = Wsh.rowbuf;
    int4* const colbuf = Wsh.colbuf;
    int* const coleff = Wsh.coleff;
    const uint32_t rstride = row_stride((uint32_t)n2), cstride = (uint32_t)n1 + 1u;
    const int j0 = C0 + C * lane;

    // ---- per-column constants ----
    uint32_t cbits = 0;   // per column c, bits 8c..8c+4: near1|regular, near2, near3, far, persisted
    int cl[C];
    int4 St[C];           // the row above: {M, H_k}, one aligned register quad per column (LDS.128 / STS.128 without moves)
    uint32_t nvalid = 0;  // number of this lane's columns that exist
    uint32_t coff3 = 0xffffffffu;  // workspace offset of the lane's last column if it is persisted (the regular persisted columns are)
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int j = j0 + c;
        const bool jv = j <= n2;
        nvalid += jv ? 1u : 0u;
        const uint32_t ci = jv ? Wsh.info2[j] : (kInfoRegular | (1u << kInfoNearShift) | 0xffu);
        cl[c] = (int)(ci & kInfoLabelMask);
        // a regular column always takes its distance-1 neighbour (column 1's is the boundary column)
        const bool c1 = (ci & ((1u << kInfoNearShift) | kInfoRegular)) != 0, c2 = (ci & (2u << kInfoNearShift)) != 0;
        const bool pers = jv && (ci & kInfoPersist);
        cbits |= ((c1 ? 1u : 0u) | (c2 ? 2u : 0u) | ((ci & (4u << kInfoNearShift)) ? 4u : 0u) | ((ci & kInfoFar) ? 8u : 0u) | (pers ? 16u : 0u)) << (8 * c);
        if (c == C - 1 && pers) coff3 = (uint32_t)slot2[j] * cstride;
        St[c] = make_int4(kMinInf, kMinInf, kMinInf, kMinInf);
        if (jv && R0 == 0) {  // boundary row (alignment.hpp:814-894): M(0,j), I_k(0,j) = -inf
            St[c] = to_offered<P>(prm, make_int4(boundary_cell<P>(Wsh.depth2[j], prm).x, kMinInf, kMinInf, kMinInf));
            rowbuf[j] = St[c];
        }
    }
    if (cs == 0 && R0 == 0) {  // boundary column as a predecessor column, and its diagonal input
        for (int i = lane; i <= n1; i += 32) {
            const int m = i == 0 ? 0 : boundary_cell<P>(Wsh.depth1[i], prm).x;
            colbuf[i] = make_int4(m, kMinInf, kMinInf, kMinInf);
            int e = kMinInf;
            const uint32_t a1 = poff1[i + 1];
#pragma unroll 1
            for (uint32_t a = poff1[i]; a < a1; ++a) {
                const uint32_t p = pidx1[a];
                e = imax(e, p == 0 ? 0 : boundary_cell<P>(Wsh.depth1[p], prm).x);
            }
            coleff[i] = e;
        }
    }
    // ---- which neighbour columns are needed ----
    // column c with near bit d (distance d) and c-d < 0 needs column 4+c-d (= 3,2,1) of the previous lane;
    // for lane 0 that is the column C0-(d-c) left of the strip.
    unsigned nb = 1u;  // bit0: previous lane's column 3 (always: column 0's distance-1 predecessor, if any)
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int d = 1; d <= 3; ++d)
            if (d > c && ((cbits >> (8 * c + d - 1)) & 1u)) nb |= 1u << (d - c - 1);
    const unsigned needS = __reduce_or_sync(kFull, nb);           // shuffles needed by any lane
    const unsigned needL = __shfl_sync(kFull, nb, 0);             // left columns needed by lane 0
    uint32_t xoffw[3] = {0, 0, 0};
#pragma unroll
    for (int d = 1; d <= 3; ++d)
        if ((needL & (1u << (d - 1))) && C0 - d >= 0 && slot2[C0 - d] >= 0) xoffw[d - 1] = (uint32_t)slot2[C0 - d] * cstride;
    __syncwarp();

    int avail = cs == 0 ? INT_MAX : 0;
    auto wait_rows = [&](int want) {
        if (avail < want) {
            CLB_WAIT_BEGIN;
            for (;;) {
                const unsigned long long v = progress[(g - 1) & kProgMask];
                avail = ((int)(v >> 32) == g) ? (int)(v & 0xffffffffu) : 0;
                if (avail >= want) break;
                __nanosleep(kPollNs);
            }
            __threadfence_block();
            CLB_WAIT_END(0);
        }
    };
    auto prefetch_block = [&](int b) {  // rows PB*b+1 .. PB*b+PB of the needed left columns -> buffer b&1
        const int row = PB * b + 1 + lane;
        if (lane < PB && row <= R1) {
#pragma unroll
            for (int d = 0; d < 3; ++d)
                if (needL & (1u << d)) {
                    cp_async_16(&sm.leftv[d][b & 1][lane], colbuf + (xoffw[d] + (uint32_t)row));
                    cp_async_4(&sm.lefte[d][b & 1][lane], coleff + (xoffw[d] + (uint32_t)row));
                }
        }
        cp_async_commit();
    };
    auto convert_block = [&](int b) {  // landed {M, D_k} + E -> {G_k, E}, in place (one lane per row)
        if (lane < PB) {
#pragma unroll
            for (int d = 0; d < 3; ++d)
                if (needL & (1u << d)) {
                    const int4 p = to_offered<P>(prm, sm.leftv[d][b & 1][lane]);
                    sm.leftv[d][b & 1][lane] = make_int4(p.y, p.z, p.w, sm.lefte[d][b & 1][lane]);
                }
        }
        __syncwarp();
    };
    // never wait for rows below the panel: the tile that fills them may be queued behind this one
    wait_rows(min(R0 + max(start_lag, PB), R1));
    if (R0 > 0) {  // the panel above must be complete before anything of this one is read (strip 0 has no other wait)
        CLB_WAIT_BEGIN;
        for (;;) {
            const unsigned long long v = progress[g & kProgMask];
            if ((int)(v >> 32) == g + 1 && (int)(v & 0xffffffffu) >= R0) break;
            __nanosleep(200);
        }
        __threadfence_block();
        CLB_WAIT_END(1);
    }
    prefetch_block(R0 / PB);
    if (R0 > 0) {  // take over rows R0-2 .. R0 from the panel above (the workspace holds {M, H_k})
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int row = R0 - t;
            const uint32_t o = (uint32_t)Wsh.slot1[row] * rstride + (uint32_t)j0;
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm.ringA) + (uint32_t)lane * 16u + (uint32_t)(row & (H - 1)) * (C * 32 * 16);
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (c < (int)nvalid) {
                    const int4 v = rowbuf[o + c];
                    sts_128(sa + c * (32 * 16), v);
                    if (t == 0) St[c] = v;
                }
        }
    }
    cp_async_wait_all();
    __syncwarp();
    convert_block(R0 / PB);

    // offered by this lane's columns 3, 2, 1 for the row just computed: {G_k} and the diagonal input E
    int O3g[3] = {kMinInf, kMinInf, kMinInf}, O2g[3] = {kMinInf, kMinInf, kMinInf}, O1g[3] = {kMinInf, kMinInf, kMinInf};
    int O3e = kMinInf, O2e = kMinInf, O1e = kMinInf;
    // explicit 32-bit shared-window addresses: keeps ptxas from re-deriving them every step
    uint32_t saA = (uint32_t)__cvta_generic_to_shared(sm.ringA) + (uint32_t)lane * 16u;
    const uint32_t saLv = (uint32_t)__cvta_generic_to_shared(&sm.leftv[0][0][0]);
    asm volatile("mov.u32 %0, %0;" : "+r"(saA));

    // Far predecessor columns (long bubbles, the boundary column of a source).  The common shape -- one such
    // column in the lane, with ONE far predecessor that lies in an earlier strip or at least two lanes back --
    // takes a fast path: the predecessor's persisted entry {M, D_k} / diagonal input for row r+1 is requested
    // while row r is being computed (it was written a step earlier at the latest), so the next step folds it in
    // without a list walk and without a load on the column chain.  Anything else keeps the generic walk.
    int farc = -1;
    const int4* farv = colbuf;
    const int* fare = coleff;
    uint32_t slowbits = 0;  // bit c: column c walks its predecessor list
#pragma unroll
    for (int c = 0; c < C; ++c) {
        if ((cbits >> (8 * c)) & 8u) {  // never set for a column beyond n2
            const int j = j0 + c;
            int nf = 0, qsel = 0;
            const uint32_t b1e = Wsh.poff2[j + 1];
#pragma unroll 1
            for (uint32_t b = Wsh.poff2[j]; b < b1e; ++b) {
                const int q = (int)Wsh.pidx2[b];
                if (q >= 1 && j - q <= kNear) continue;
                ++nf;
                qsel = q;
            }
            if (nf == 1 && farc < 0 && slot2[qsel] >= 0 && (qsel < C0 || (qsel - C0) / C <= lane - 2)) {
                farc = c;
                farv = colbuf + (size_t)((uint32_t)slot2[qsel] * cstride);
                fare = coleff + (size_t)((uint32_t)slot2[qsel] * cstride);
            } else {
                slowbits |= 1u << c;
            }
        }
    }
    int4 Fv = make_int4(kMinInf, kMinInf, kMinInf, kMinInf);
    int Fe = kMinInf;

    // A plain strip -- no column with a distance-3 / far predecessor -- runs the lean step: none of the rare code,
    // the four columns of a lane are one basic block.
    const bool lean = !__any_sync(kFull, (cbits & 0x0c0c0c0cu) != 0) && !(needS & 4u) && !(needL & 4u) && !(dbg & 2);
    // warp-uniform: does any lane of this strip persist one of its first three columns (a column with a far successor,
    // or whose distance-2 successor lies in the next 32-column block)?  Their offsets; the stores sit behind one branch.
    const bool strip_pers012 = __any_sync(kFull, (cbits & 0x00101010u) != 0);
    uint32_t coff012[C - 1];  // their workspace offsets
#pragma unroll
    for (int c = 0; c < C - 1; ++c) coff012[c] = ((cbits >> (8 * c)) & 16u) ? (uint32_t)slot2[j0 + c] * cstride : 0xffffffffu;
    const uint32_t rpersist = nvalid ? kInfoPersist : 0u;  // a lane entirely beyond n2 stores no rows
    const bool slot_escape = n1 >= (int)kInfoSlotEscape;    // warp-uniform: row slots may exceed the info word's field
    // lane-constant column shapes of the lean step and the labels, in registers of their own (re-deriving them from the
    // packed words costs a dozen instructions per step)
    int cB[C], cJ[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        cB[c] = ((cbits >> (8 * c)) & 1u) ? 0 : 1;        // no distance-1 predecessor: a second allele
        cJ[c] = (((cbits >> (8 * c)) & 3u) == 3u) ? 1 : 0; // distance 1 and 2: the node after a bubble
        asm volatile("mov.u32 %0, %0;" : "+r"(cB[c]));
        asm volatile("mov.u32 %0, %0;" : "+r"(cJ[c]));
        asm volatile("mov.u32 %0, %0;" : "+r"(cl[c]));
    }
    const bool lane0L1 = lane == 0 && (needL & 1u);        // lane 0 takes its left neighbour from the left-column buffer

    // 7-way max of the recurrence (2P+1 values), DPX three-input maxima; the left-dependent values come last
    auto cell_max = [&](int x, const int4& I, const int (&D)[3]) {
        if (P == 1) return __vimax3_s32(x, I.y, D[0]);
        if (P == 2) return __vimax3_s32(__vimax3_s32(x, I.y, I.z), D[0], D[1]);
        return __vimax3_s32(__vimax3_s32(__vimax3_s32(x, I.y, I.z), D[0], D[1]), D[2], I.w);
    };
    auto fold4 = [&](int4& a, const int4 b) {
        a.x = imax(a.x, b.x); a.y = imax(a.y, b.y); a.z = imax(a.z, b.z); a.w = imax(a.w, b.w);
    };

    // lane 0's entry of the left-column buffers for the row it is on: rows PB*b+1.. live in buffer b&1 at index (row-1) % PB,
    // column d at leftv[d]; the address advances by one entry per step and flips buffers every PB steps
    uint32_t laddr = saLv + (uint32_t)((R0 / PB) & 1) * (PB * 16u);
    // Row info words travel down the lanes with the wavefront (lane t is on the row lane t-1 was on a step earlier):
    // one shuffle per step, and lane 0 takes the word of its new row, which every lane requested (one broadcast
    // address) three steps earlier -- no step waits for a load.
    uint32_t rinfo_cur = (lane == 0) ? info1[R0 + 1] : 0u;                 // row of this lane in the coming step
    uint32_t rq1 = info1[min(R0 + 2, R1 + 1)], rq2 = info1[min(R0 + 3, R1 + 2)];  // lane 0's rows one and two steps ahead

    auto step = [&](const int s, auto guard_tag, auto lean_tag) {
        (void)guard_tag;
        constexpr bool LEAN = decltype(lean_tag)::value;
        const int r = 1 + s - lane;
        const bool act = r > R0 && r <= R1;  // lanes outside the panel (pipeline fill / drain) idle
        // requests first: lane 0's left-column entry (all lanes read it: a broadcast) and the info word two rows down
        const int4 tl = lds_128(laddr);
        laddr += 16u;
        const uint32_t rq3 = info1[(uint32_t)min(s + 4, R1 + 2)];  // lane 0's row three steps ahead (the info array is padded by two entries)
        // previous lane's columns for this row (it finished the row one step ago); in lane 0, the columns left of the
        // strip.  Columns that no lane needs are not shuffled: whatever value stands in for them is never selected.
        int P3g[3] = {kMinInf, kMinInf, kMinInf}, P2g[3] = {O2g[0], O2g[1], O2g[2]}, P1g[3] = {O1g[0], O1g[1], O1g[2]};
        int P3e, P2e = O2e, P1e = O1e;
#pragma unroll
        for (int k = 0; k < P; ++k) P3g[k] = __shfl_up_sync(kFull, O3g[k], 1);
        P3e = __shfl_up_sync(kFull, O3e, 1);
        if (needS & 2u) {
#pragma unroll
            for (int k = 0; k < P; ++k) P2g[k] = __shfl_up_sync(kFull, O2g[k], 1);
            P2e = __shfl_up_sync(kFull, O2e, 1);
        }
        if (!LEAN && (needS & 4u)) {
#pragma unroll
            for (int k = 0; k < P; ++k) P1g[k] = __shfl_up_sync(kFull, O1g[k], 1);
            P1e = __shfl_up_sync(kFull, O1e, 1);
        }
        if (lane0L1) { P3g[0] = tl.x; P3g[1] = tl.y; P3g[2] = tl.z; P3e = tl.w; }
        if (needL & 6u) {
            if (needL & 2u) {
                const int4 t = lds_128(laddr + (2 * PB - 1) * 16);
                if (lane == 0) { P2g[0] = t.x; P2g[1] = t.y; P2g[2] = t.z; P2e = t.w; }
            }
            if (!LEAN && (needL & 4u)) {
                const int4 t = lds_128(laddr + (4 * PB - 1) * 16);
                if (lane == 0) { P1g[0] = t.x; P1g[1] = t.y; P1g[2] = t.z; P1e = t.w; }
            }
        }
        const uint32_t rinfo = rinfo_cur;
        if (act) {
            const uint32_t rsA = saA + (uint32_t)(r & (H - 1)) * (C * 32 * 16);
            const int rlabel = (int)(rinfo & kInfoLabelMask);
            // ---- rows that are not "one predecessor, the row above" ----
            // The two SNP-bubble shapes are folded in WITHOUT a branch (with 32 lanes on 32 different rows some lane is
            // on such a row in nine steps out of ten, so a branch would be taken by the whole warp anyway): the second
            // allele (predecessor r-2 only) reloads its state from the ring, the node after the bubble (r-1 and r-2)
            // takes the element-wise maximum; predicated loads and maxima, one basic block with the cells.
            {
                const uint32_t rs2 = saA + (uint32_t)((r - 2) & (H - 1)) * (C * 32 * 16);  // row r-2
                const uint32_t nb12 = rinfo & (3u << kInfoNearShift);
                const uint32_t isB = (nb12 == (2u << kInfoNearShift)) ? 1u : 0u, isJ = (nb12 == (3u << kInfoNearShift)) ? 1u : 0u;
                asm volatile(
                    "{\n .reg .pred pb, pj;\n .reg .s32 t<16>;\n"
                    " setp.ne.u32 pb, %17, 0;\n setp.ne.u32 pj, %18, 0;\n"
                    " @pb ld.shared.v4.s32 {%0,%1,%2,%3}, [%16];\n"
                    " @pb ld.shared.v4.s32 {%4,%5,%6,%7}, [%16+512];\n"
                    " @pb ld.shared.v4.s32 {%8,%9,%10,%11}, [%16+1024];\n"
                    " @pb ld.shared.v4.s32 {%12,%13,%14,%15}, [%16+1536];\n"
                    " @pj ld.shared.v4.s32 {t0,t1,t2,t3}, [%16];\n"
                    " @pj ld.shared.v4.s32 {t4,t5,t6,t7}, [%16+512];\n"
                    " @pj ld.shared.v4.s32 {t8,t9,t10,t11}, [%16+1024];\n"
                    " @pj ld.shared.v4.s32 {t12,t13,t14,t15}, [%16+1536];\n"
                    " @pj max.s32 %0, %0, t0;\n @pj max.s32 %1, %1, t1;\n @pj max.s32 %2, %2, t2;\n @pj max.s32 %3, %3, t3;\n"
                    " @pj max.s32 %4, %4, t4;\n @pj max.s32 %5, %5, t5;\n @pj max.s32 %6, %6, t6;\n @pj max.s32 %7, %7, t7;\n"
                    " @pj max.s32 %8, %8, t8;\n @pj max.s32 %9, %9, t9;\n @pj max.s32 %10, %10, t10;\n @pj max.s32 %11, %11, t11;\n"
                    " @pj max.s32 %12, %12, t12;\n @pj max.s32 %13, %13, t13;\n @pj max.s32 %14, %14, t14;\n @pj max.s32 %15, %15, t15;\n"
                    "}\n"
                    : "+r"(St[0].x), "+r"(St[0].y), "+r"(St[0].z), "+r"(St[0].w), "+r"(St[1].x), "+r"(St[1].y), "+r"(St[1].z), "+r"(St[1].w),
                      "+r"(St[2].x), "+r"(St[2].y), "+r"(St[2].z), "+r"(St[2].w), "+r"(St[3].x), "+r"(St[3].y), "+r"(St[3].z), "+r"(St[3].w)
                    : "r"(rs2), "r"(isB), "r"(isJ)
                    : "memory");
            }
            // everything else -- a distance-3 or far predecessor, or no predecessor within two rows -- is uncommon
            if ((rinfo & ((4u << kInfoNearShift) | kInfoFar)) || !(rinfo & (kInfoRegular | (3u << kInfoNearShift)))) {
                if (!(rinfo & (kInfoRegular | (3u << kInfoNearShift)))) {
#pragma unroll
                    for (int c = 0; c < C; ++c) St[c] = make_int4(kMinInf, kMinInf, kMinInf, kMinInf);
                }
                if (rinfo & (4u << kInfoNearShift)) {
                    const uint32_t rs3 = saA + (uint32_t)((r - 3) & (H - 1)) * (C * 32 * 16);
#pragma unroll
                    for (int c = 0; c < C; ++c) fold4(St[c], lds_128(rs3 + c * (32 * 16)));
                }
                if (rinfo & kInfoFar) {
                    const uint32_t rp1 = Wsh.poff1[r + 1];
#pragma unroll 1
                    for (uint32_t a = Wsh.poff1[r]; a < rp1; ++a) {
                        const int p = (int)Wsh.pidx1[a];
                        if (p >= 1 && r - p <= kNear) continue;
                        const uint32_t o = (uint32_t)Wsh.slot1[p] * rstride + (uint32_t)j0;
#pragma unroll
                        for (int c = 0; c < C; ++c)
                            if (c < (int)nvalid) fold4(St[c], rowbuf[o + c]);  // the workspace holds {M, H_k}
                    }
                }
            }
            int G[C][3], eM[C];
            auto column = [&](auto ctag) {
                constexpr int c = decltype(ctag)::value;
                eM[c] = St[c].x;  // effective M of the row above = diagonal input of the columns to the right
                // ---- effective predecessor column: the distance-1 neighbour, the distance-2 neighbour, or both ----
                int d1g[3], d2g[3], d1e, d2e;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if constexpr (c == 0) { d1g[k] = P3g[k]; d2g[k] = P2g[k]; }
                    else if constexpr (c == 1) { d1g[k] = G[0][k]; d2g[k] = P3g[k]; }
                    else { d1g[k] = G[c - 1][k]; d2g[k] = G[c - 2][k]; }
                }
                if constexpr (c == 0) { d1e = P3e; d2e = P2e; }
                else if constexpr (c == 1) { d1e = eM[0]; d2e = P3e; }
                else { d1e = eM[c - 1]; d2e = eM[c - 2]; }
                const bool c1 = (cbits >> (8 * c)) & 1u, c2 = (cbits >> (8 * c)) & 2u;
                int D[3] = {kMinInf, kMinInf, kMinInf};
                int E;
                if (LEAN) {  // a lean strip knows three column shapes: (c1, !c2) regular, (!c1, c2) second allele, (c1, c2) after a bubble
                    const bool pB = cB[c] != 0, pJ = cJ[c] != 0;
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        D[k] = pB ? d2g[k] : d1g[k];
                        if (pJ) D[k] = imax(D[k], d2g[k]);
                    }
                    E = pB ? d2e : d1e;
                    if (pJ) E = imax(E, d2e);
                } else {
#pragma unroll
                    for (int k = 0; k < P; ++k) D[k] = imax(c1 ? d1g[k] : kMinInf, c2 ? d2g[k] : kMinInf);
                    E = imax(c1 ? d1e : kMinInf, c2 ? d2e : kMinInf);
                }
                if (!LEAN && ((cbits >> (8 * c)) & 12u)) {
                    if (farc == c) {  // fast far column
                        if (r == R0 + 1) { Fv = farv[r]; Fe = fare[r]; }
                        const int4 fp = to_offered<P>(prm, Fv);
                        D[0] = imax(D[0], fp.y); D[1] = imax(D[1], fp.z); D[2] = imax(D[2], fp.w);
                        E = imax(E, Fe);
                        const int rn = min(r + 1, R1);  // not below the panel: those rows may not exist yet
                        Fv = farv[rn]; Fe = fare[rn];
                    }
                    if ((cbits >> (8 * c)) & 4u) {  // distance 3
                        int d3e;
#pragma unroll
                        for (int k = 0; k < P; ++k) {
                            int d3g;
                            if constexpr (c >= 3) d3g = G[c - 3][k];
                            else if constexpr (c == 2) d3g = P3g[k];
                            else if constexpr (c == 1) d3g = P2g[k];
                            else d3g = P1g[k];
                            D[k] = imax(D[k], d3g);
                        }
                        if constexpr (c >= 3) d3e = eM[c - 3];
                        else if constexpr (c == 2) d3e = P3e;
                        else if constexpr (c == 1) d3e = P2e;
                        else d3e = P1e;
                        E = imax(E, d3e);
                    }
                    if (slowbits & (1u << c)) {
                        const int j = j0 + c;
                        const uint32_t b1e = Wsh.poff2[j + 1];
#pragma unroll 1
                        for (uint32_t b = Wsh.poff2[j]; b < b1e; ++b) {
                            const int q = (int)Wsh.pidx2[b];
                            if (q >= 1 && j - q <= kNear) continue;
                            const uint32_t o = (uint32_t)Wsh.slot2[q] * cstride + (uint32_t)r;
                            const int4 t = to_offered<P>(prm, colbuf[o]);
                            D[0] = imax(D[0], t.y); D[1] = imax(D[1], t.z); D[2] = imax(D[2], t.w);
                            E = imax(E, coleff[o]);
                        }
                    }
                }
                // ---- the cell ----
                const int x = (rlabel == cl[c]) ? E + prm.match : E - prm.mismatch;
                const int Mn = cell_max(x, St[c], D);
                {
                    const int T = Mn - prm.oe[0];
                    G[c][0] = __viaddmax_s32(D[0], -prm.e[0], T);
                    St[c].y = __viaddmax_s32(St[c].y, -prm.e[0], T);
                }
                G[c][1] = G[c][2] = kMinInf;
                if (P > 1) {
                    const int T = Mn - prm.oe[1];
                    G[c][1] = __viaddmax_s32(D[1], -prm.e[1], T);
                    St[c].z = __viaddmax_s32(St[c].z, -prm.e[1], T);
                }
                if (P > 2) {
                    const int T = Mn - prm.oe[2];
                    G[c][2] = __viaddmax_s32(D[2], -prm.e[2], T);
                    St[c].w = __viaddmax_s32(St[c].w, -prm.e[2], T);
                }
                St[c].x = Mn;
                sts_128(rsA + c * (32 * 16), St[c]);
                if constexpr (c == C - 1) {
                    if (coff3 != 0xffffffffu) {  // the regular persisted columns (every 32nd) are a lane's last
                        colbuf[coff3 + (uint32_t)r] = make_int4(Mn, D[0], D[1], D[2]);
                        coleff[coff3 + (uint32_t)r] = eM[c];
                    }
                } else {
                    if (strip_pers012) {  // warp-uniform and uncommon
                        if (coff012[c] != 0xffffffffu) {
                            const uint32_t o = coff012[c] + (uint32_t)r;
                            colbuf[o] = make_int4(Mn, D[0], D[1], D[2]);
                            coleff[o] = eM[c];
                        }
                    }
                }
            };
            column(std::integral_constant<int, 0>{});
            column(std::integral_constant<int, 1>{});
            column(std::integral_constant<int, 2>{});
            column(std::integral_constant<int, 3>{});
            {   // persisted row: {M, H_k} of the lane's four columns (rows are padded: columns beyond n2 land in the padding)
                uint32_t rslot = rinfo >> kInfoSlotShift;
                if (slot_escape) {
                    if ((rinfo & rpersist) && rslot == kInfoSlotEscape) rslot = (uint32_t)Wsh.slot1[r];
                }
                int4* const rp = rowbuf + (size_t)(rslot * rstride + (uint32_t)j0);
                if (rinfo & rpersist) {
#pragma unroll
                    for (int c = 0; c < C; ++c) rp[c] = St[c];
                }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) { O3g[k] = G[3][k]; O2g[k] = G[2][k]; if (!LEAN) O1g[k] = G[1][k]; }
            O3e = eM[3]; O2e = eM[2];
            if (!LEAN) O1e = eM[1];
        }
        {   // next step: every lane moves one row down
            const uint32_t up = __shfl_up_sync(kFull, rinfo, 1);
            rinfo_cur = (lane == 0) ? rq1 : up;
            rq1 = rq2; rq2 = rq3;
        }
        __syncwarp();
    };
    auto publish = [&](int r31) {
        __threadfence_block();
        progress[g & kProgMask] = ((unsigned long long)(g + 1) << 32) | (unsigned)r31;
    };

    // PB-step blocks; lane 0 is on rows s0+1 .. s0+PB (R0 is a multiple of PB).  Lean and generic strips have loops of
    // their own, so that what only the generic step keeps across steps is not live in the lean loop.
    const int nsteps = R1 + 31;
    auto block_head = [&](int s0) {
        const int bnext = s0 / PB + 1;
        const bool more = PB * bnext + 1 <= R1;
        if (more) {  // only lane 0 reads the prefetch buffers, so the next block can be requested right away
            wait_rows(min(PB * bnext + PB, R1));
            prefetch_block(bnext);
        }
        laddr = saLv + (uint32_t)((s0 / PB) & 1) * (PB * 16u);
        return more;
    };
    auto block_tail = [&](int s0, bool more) {
        cp_async_wait_all();
        __syncwarp();
        if (more) convert_block(s0 / PB + 1);
    };
    auto publish_after = [&](int e8) {
        if (lane == 31) {
            const int r31 = e8 - 31;
            if (r31 > R0) publish(min(r31, R1));
        }
    };
    if (lean) {
        for (int s0 = R0; s0 < nsteps; s0 += PB) {
            const int s1 = min(s0 + PB, nsteps);
            const bool more = block_head(s0);
#pragma unroll 1
            for (int q8 = s0; q8 < s1; q8 += 8) {
                const int e8 = min(q8 + 8, s1);
#pragma unroll 1
                for (int s = q8; s < e8; ++s) step(s, std::false_type{}, std::true_type{});
                publish_after(e8);
            }
            block_tail(s0, more);
        }
    } else {
        for (int s0 = R0; s0 < nsteps; s0 += PB) {
            const int s1 = min(s0 + PB, nsteps);
            const bool more = block_head(s0);
#pragma unroll 1
            for (int q8 = s0; q8 < s1; q8 += 8) {
                const int e8 = min(q8 + 8, s1);
#pragma unroll 1
                for (int s = q8; s < e8; ++s) step(s, std::true_type{}, std::false_type{});
                publish_after(e8);
            }
            block_tail(s0, more);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Boundary row / column (alignment.hpp:814-894).  In the boundary column only the lead
// insertion is extended, so I_k(i,0) = -(o_k + e_k * depth(i)) with depth = fewest nodes on a
// path from a source; the host supplies depth, which makes this pass embarrassingly parallel.
// ------------------------------------------------------------------------------------------
template <int P>
__device__ __forceinline__ int4 boundary_cell(uint32_t depth, const Params& prm) {
    int v[3] = {kMinInf, kMinInf, kMinInf};
    int m = kMinInf;
    if (depth) {
#pragma unroll
        for (int k = 0; k < P; ++k) {
            v[k] = imax(kMinInf, (int)(0u - (uint32_t)prm.oe[k] - (uint32_t)prm.e[k] * (depth - 1)));
            m = imax(m, v[k]);
        }
    }
    return make_int4(m, v[0], v[1], v[2]);
}

// Boundary data that only the traceback reads: the full boundary row {M,D_k} / column {M,I_k}
// and the boundary entries of persisted rows / columns.  Run by the traceback warp.
template <int P>
__device__ void tb_boundary(const Win& W, const Params& prm, int lane) {
    const int64_t rstride = (int64_t)row_stride((uint32_t)W.n2), cstride = (int64_t)W.n1 + 1;
    const int4 corner = make_int4(0, kMinInf, kMinInf, kMinInf);
    for (int j = lane; j <= W.n2; j += 32) {
        const int4 b = j == 0 ? corner : boundary_cell<P>(W.depth2[j], prm);
        W.brow[j] = b;
        if (j == 0 || W.n1 == 0) W.rowbuf[j] = make_int4(b.x, kMinInf, kMinInf, kMinInf);
        if (W.info2[j] & kInfoPersist) W.colbuf[(int64_t)W.slot2[j] * cstride] = b;
    }
    for (int i = lane; i <= W.n1; i += 32) {
        const int4 b = i == 0 ? corner : boundary_cell<P>(W.depth1[i], prm);
        W.bcol[i] = b;
        if (W.n2 == 0) W.colbuf[i] = make_int4(b.x, kMinInf, kMinInf, kMinInf);
        if (W.info1[i] & kInfoPersist) W.rowbuf[(int64_t)W.slot1[i] * rstride] = b;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Traceback (the CTA's traceback warp).  Mirrors alignment.hpp:979-1138 on recomputed cell values.
// ------------------------------------------------------------------------------------------
struct TileView {
    int R0, R1, C0;  // rows R0..R1 of the tile are in shared memory, columns C0..C0+31; R0 = 0 means "no tile"
    int B0;          // first row of the recomputation (the 64-row block): base of the cached row info words
    int W0;          // the walk may stand on rows W0..R1: its near predecessors (up to kNear rows up) must be in shared memory too
};

// shared memory of the traceback warp.  A tile is recomputed from the top of its 64-row block down to the row the
// walk is on, but only its last kTileKeep rows are kept (a ring): the walk re-enters the recomputation when it
// climbs above them.  Half the shared memory of keeping all 64 rows, which the fill warps' L1 gets.
constexpr int kTileKeep = 32;
struct __align__(16) TileSmem {
    int4 A[kTileKeep * 32];       // {M, I_k}       [row & 31][lane]: the LAST kTileKeep rows of the recomputed tile
    int4 B[kTileKeep * 32];       // {diag in, D_k} [row & 31][lane]
    int4 leftv[3][kRowBlock];     // {M, D_k} of the 3 columns left of the tile, rows R0..R1
    int lefte[3][kRowBlock];      // their diagonal input
    uint32_t rinfo[kRowBlock];    // info words of the tile's rows
    uint32_t cinfo[kStrip];       // info words of the tile's columns
};

// Recompute one tile: rows R0..R1 (same 64-row block) x 32 columns from C0, every cell kept in
// shared memory.  Same recurrence and same near / far predecessor handling as fill_strip; what
// lies outside the tile comes from the persisted rows / columns of the window workspace.
template <int P>
__device__ __forceinline__ void tile_strip(const Win& Wsh, const Params& prm, const int C0, const int R0, const int R1,
                                           TileSmem& sm, const int lane) {
    constexpr int H = kTileKeep;
    const int n1 = Wsh.n1, n2 = Wsh.n2;
    const uint32_t* __restrict__ info1 = Wsh.info1;
    const int32_t* __restrict__ slot1 = Wsh.slot1;
    const uint32_t* __restrict__ poff1 = Wsh.poff1;
    const uint32_t* __restrict__ pidx1 = Wsh.pidx1;
    const int32_t* __restrict__ slot2 = Wsh.slot2;
    const uint32_t* __restrict__ pidx2 = Wsh.pidx2;
    const int4* const rowbuf = Wsh.rowbuf;
    const int4* const colbuf = Wsh.colbuf;
    const int* const coleff = Wsh.coleff;
    const uint32_t rstride = row_stride((uint32_t)n2), cstride = (uint32_t)n1 + 1u;
    const int nrows = R1 - R0 + 1;

    const int j = C0 + lane;
    const bool jvalid = j <= n2;
    const uint32_t cinfo = jvalid ? Wsh.info2[j] : (kInfoRegular | (1u << kInfoNearShift) | 0xffu);
    sm.cinfo[lane] = cinfo;
    const int clabel = (int)(cinfo & kInfoLabelMask);
    const bool creg = (cinfo & kInfoRegular) != 0;
    const uint32_t cmask = (cinfo >> kInfoNearShift) & 7u;
    const bool cfar = (cinfo & kInfoFar) != 0;
    const bool colpath = !creg || lane == 0;
    const uint32_t cp0 = (jvalid && cfar) ? Wsh.poff2[j] : 0u, cp1 = (jvalid && cfar) ? Wsh.poff2[j + 1] : 0u;

    unsigned needbits = 0;
    if (jvalid) {
#pragma unroll
        for (int dd = 1; dd <= 3; ++dd)
            if (dd > lane && ((cmask >> (dd - 1)) & 1u)) needbits |= 1u << (dd - lane - 1);
        if (lane == 0 && creg) needbits |= 1u;
    }
    needbits = __reduce_or_sync(kFull, needbits);
    // cache the left columns and the row info words of the tile
    for (int idx = lane; idx < nrows; idx += 32) {
        const int r = R0 + idx;
        sm.rinfo[idx] = info1[r];
#pragma unroll
        for (int d = 0; d < 3; ++d)
            if (needbits & (1u << d)) {
                const uint32_t o = (uint32_t)slot2[C0 - 1 - d] * cstride + (uint32_t)r;
                sm.leftv[d][idx] = colbuf[o];
                sm.lefte[d][idx] = coleff[o];
            }
    }
    // the row above as {M, H_k}: what it offers this row (rows outside the tile come from the workspace in that form)
    int upM = kMinInf, upH[3] = {kMinInf, kMinInf, kMinInf};
    if (jvalid) {
        const int s0 = slot1[R0 - 1];
        if (s0 >= 0) {
            const int4 a = rowbuf[(uint32_t)s0 * rstride + (uint32_t)j];
            upM = a.x; upH[0] = a.y; upH[1] = a.z; upH[2] = a.w;
        }
    }
    __syncwarp();

    int outM = kMinInf, outD[3] = {kMinInf, kMinInf, kMinInf}, outEff = kMinInf;
    int4* const myA = sm.A + lane;
    int4* const myB = sm.B + lane;
    const int nsteps = nrows + 31;
#pragma unroll 1
    for (int s = 0; s < nsteps; ++s) {
        const int r = R0 + s - lane;
        const bool act = jvalid && r >= R0 && r <= R1;
        int lM = __shfl_up_sync(kFull, outM, 1);
        int lD[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) lD[k] = (k < P) ? __shfl_up_sync(kFull, outD[k], 1) : kMinInf;
        int lEff = __shfl_up_sync(kFull, outEff, 1);
        if (act) {
            const int li = r - R0;
            const uint32_t rinfo = sm.rinfo[li];
            const int rs = (r & (H - 1)) * 32;
            // ---- effective predecessor row: I_k(r,j) = max over predecessor rows of H_k, diagonal input eM likewise ----
            int eM = upM, eH[3] = {upH[0], upH[1], upH[2]};
            if (!(rinfo & kInfoRegular)) {
                if (!(rinfo & (1u << kInfoNearShift))) { eM = kMinInf; eH[0] = eH[1] = eH[2] = kMinInf; }
#pragma unroll
                for (int d = 2; d <= 3; ++d) {
                    if (rinfo & ((1u << (d - 1)) << kInfoNearShift)) {
                        const int p = r - d;  // inside the tile: shared memory holds {M, I_k}; above it: the workspace, {M, H_k}
                        const int4 v = p >= R0 ? to_offered<P>(prm, myA[(p & (H - 1)) * 32]) : rowbuf[(uint32_t)slot1[p] * rstride + (uint32_t)j];
                        max4(eM, eH, v);
                    }
                }
                if (rinfo & kInfoFar) {
                    const uint32_t rp1 = poff1[r + 1];
#pragma unroll 1
                    for (uint32_t a = poff1[r]; a < rp1; ++a) {
                        const int p = (int)pidx1[a];
                        if (p >= 1 && r - p <= kNear) continue;
                        max4(eM, eH, rowbuf[(uint32_t)slot1[p] * rstride + (uint32_t)j]);
                    }
                }
            }
            // ---- effective predecessor column + diagonal input ----
            if (colpath) {
                if (!(cmask & 1u) && !creg) { lM = kMinInf; lD[0] = lD[1] = lD[2] = kMinInf; lEff = kMinInf; }
                else if (lane == 0) {
                    const int4 b = sm.leftv[0][li];
                    lM = b.x; lD[0] = b.y; lD[1] = b.z; lD[2] = b.w;
                    lEff = sm.lefte[0][li];
                }
#pragma unroll
                for (int d = 2; d <= 3; ++d) {
                    if (cmask & (1u << (d - 1))) {
                        int4 v;
                        int m;
                        if (lane >= d) {
                            v = sm.B[rs + lane - d];
                            m = sm.A[rs + lane - d].x;
                        } else {
                            const int4 t = sm.leftv[d - lane - 1][li];
                            m = t.x;
                            v = make_int4(sm.lefte[d - lane - 1][li], t.y, t.z, t.w);
                        }
                        lM = imax(lM, m);
                        max4(lEff, lD, v);
                    }
                }
                if (cfar) {
#pragma unroll 1
                    for (uint32_t b = cp0; b < cp1; ++b) {
                        const int q = (int)pidx2[b];
                        if (q >= 1 && j - q <= kNear) continue;
                        const uint32_t o = (uint32_t)slot2[q] * cstride + (uint32_t)r;
                        max4(lM, lD, colbuf[o]);
                        lEff = imax(lEff, coleff[o]);
                    }
                }
            }
            // ---- the cell ----
            const int sub = ((int)(rinfo & kInfoLabelMask) == clabel) ? prm.match : -prm.mismatch;
            int I[3] = {kMinInf, kMinInf, kMinInf}, D[3] = {kMinInf, kMinInf, kMinInf};
            int M = __viaddmax_s32(lEff, sub, kMinInf);
#pragma unroll
            for (int k = 0; k < P; ++k) {
                I[k] = eH[k];
                D[k] = __viaddmax_s32(lD[k], -prm.e[k], lM - prm.oe[k]);
                M = __vimax3_s32(M, I[k], D[k]);
            }
            myA[rs] = make_int4(M, I[0], I[1], I[2]);
            myB[rs] = make_int4(eM, D[0], D[1], D[2]);
            upM = M;
#pragma unroll
            for (int k = 0; k < P; ++k) upH[k] = offer(prm, I[k], M, k);
            outM = M; outD[0] = D[0]; outD[1] = D[1]; outD[2] = D[2];
            outEff = eM;
        }
        __syncwarp();
    }
}

template <int P>
struct Walker {
    const Win& W;
    const Params& prm;
    const TileSmem& sm;
    TileView tv;
    uint32_t rstride, cstride;

    __device__ bool in_tile(int i, int j) const {
        return tv.R0 > 0 && i >= tv.R0 && i <= tv.R1 && j >= tv.C0 && j < tv.C0 + kStrip && j <= W.n2;
    }
    __device__ bool walkable(int i, int j) const { return i >= tv.W0 && in_tile(i, j); }
    // M(i,j); the corner reads as -inf in every traceback test (it is never a match)
    __device__ int cM(int i, int j) const {
        if (i == 0) return j == 0 ? kMinInf : W.brow[j].x;
        if (j == 0) return W.bcol[i].x;
        if (in_tile(i, j)) return sm.A[(i & (kTileKeep - 1)) * 32 + (j - tv.C0)].x;
        const int s = W.slot1[i];
        if (s >= 0) return W.rowbuf[(uint32_t)s * rstride + (uint32_t)j].x;
        return W.colbuf[(uint32_t)W.slot2[j] * cstride + (uint32_t)i].x;
    }
    // I_k(i,j) of the CURRENT cell: always on the boundary or inside the recomputed tile
    __device__ int cI(int i, int j, int k) const {
        if (i == 0) return kMinInf;
        const int4 v = (j == 0) ? W.bcol[i] : sm.A[(i & (kTileKeep - 1)) * 32 + (j - tv.C0)];
        return k == 0 ? v.y : (k == 1 ? v.z : v.w);
    }
    // "I_k(p,j) - e_k" of a PREDECESSOR row p, for the extension test of alignment.hpp:1101-1116.  A row outside the tile
    // is in the workspace as H_k = max(I_k - e_k, M - oe_k); the reference tests the opening (cur == M(p,j) - oe_k)
    // before the extension for the same predecessor, so when this value is compared the opening has already failed,
    // and then cur == I_k - e_k  <=>  cur == H_k  (cur >= H_k >= M - oe_k != cur rules the other branch of the max out).
    __device__ int cIext(int p, int j, int k) const {
        if (p == 0) return kMinInf;
        if (j == 0 || in_tile(p, j)) return cI(p, j, k) - prm.e[k];
        const int4 v = W.rowbuf[(uint32_t)W.slot1[p] * rstride + (uint32_t)j];
        return k == 0 ? v.y : (k == 1 ? v.z : v.w);
    }
    __device__ int cD(int i, int j, int k) const {
        if (j == 0) return kMinInf;
        int4 v;
        if (i == 0) v = W.brow[j];
        else if (in_tile(i, j)) v = sm.B[(i & (kTileKeep - 1)) * 32 + (j - tv.C0)];
        else v = W.colbuf[(uint32_t)W.slot2[j] * cstride + (uint32_t)i];
        return k == 0 ? v.y : (k == 1 ? v.z : v.w);
    }
    __device__ uint32_t rinfo(int i) const { return (tv.R0 > 0 && i >= tv.B0 && i <= tv.R1) ? sm.rinfo[i - tv.B0] : W.info1[i]; }
    __device__ uint32_t cinfo(int j) const { return (tv.R0 > 0 && j >= tv.C0 && j < tv.C0 + kStrip && j <= W.n2) ? sm.cinfo[j - tv.C0] : W.info2[j]; }
};

// predecessor list of a node in previous() order; regular nodes (one predecessor = index-1) need no memory access
struct PredList {
    const uint32_t* ptr;
    int n, single;
    __device__ int at(int k) const { return ptr ? (int)ptr[k] : single; }
};
__device__ __forceinline__ PredList pred_list(uint32_t info, int idx, const uint32_t* poff, const uint32_t* pidx) {
    PredList pl;
    if (info & kInfoRegular) {
        pl.ptr = nullptr; pl.n = 1; pl.single = idx - 1;
    } else {
        const uint32_t a = poff[idx];
        pl.ptr = pidx + a; pl.n = (int)(poff[idx + 1] - a); pl.single = 0;
    }
    return pl;
}

template <int P>
__device__ void traceback(const Win& W, const Params& prm, TileSmem& sm, int lane, int64_t* score_out, int32_t* aln,
                          uint32_t* len_out) {
    const int n1 = W.n1, n2 = W.n2;
    const uint32_t rstride = row_stride((uint32_t)n2);
    // ---- best sink pair: first maximum in caller order, strict '>' (alignment.hpp:979-1008) ----
    long long npairs;
    if (n1 != 0 && n2 != 0) npairs = (long long)W.nsnk1 * W.nsnk2;
    else if (n1 != 0) npairs = W.nsnk1;
    else if (n2 != 0) npairs = W.nsnk2;
    else npairs = 0;
    int best = INT_MIN;
    long long besti = LLONG_MAX;
    for (long long x = lane; x < npairs; x += 32) {
        int v;
        if (n1 != 0 && n2 != 0) {
            const int i = (int)W.snk1[x / W.nsnk2], j = (int)W.snk2[x % W.nsnk2];
            v = W.rowbuf[(uint32_t)W.slot1[i] * rstride + (uint32_t)j].x;
        } else if (n1 != 0) {
            v = W.bcol[W.snk1[x]].x;
        } else {
            v = W.brow[W.snk2[x]].x;
        }
        if (v > best) { best = v; besti = x; }  // x ascending per lane: keeps the first maximum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ob = __shfl_xor_sync(kFull, best, o);
        const long long oi = __shfl_xor_sync(kFull, besti, o);
        if (oi != LLONG_MAX && (besti == LLONG_MAX || ob > best || (ob == best && oi < besti))) { best = ob; besti = oi; }
    }
    int ci = -1, cj = -1;
    if (besti != LLONG_MAX) {
        if (n1 != 0 && n2 != 0) { ci = (int)W.snk1[besti / W.nsnk2]; cj = (int)W.snk2[besti % W.nsnk2]; }
        else if (n1 != 0) { ci = (int)W.snk1[besti]; cj = 0; }
        else { ci = 0; cj = (int)W.snk2[besti]; }
    }
    if (lane == 0) *score_out = (ci >= 0) ? (long long)best : 0;

    Walker<P> wk{W, prm, sm, TileView{0, 0, 0, 0, 0}, rstride, (uint32_t)n1 + 1u};
    const int cap = n1 + n2;
    int len = 0, comp = 0;
    while (ci >= 0) {  // warp-uniform: ci/cj are broadcast from lane 0 below
        if (ci >= 1 && cj >= 1 && !wk.walkable(ci, cj)) {
            const int R0 = 1 + ((ci - 1) / kRowBlock) * kRowBlock;
            const int C0 = 1 + ((cj - 1) / kStrip) * kStrip;
            __syncwarp();
            tile_strip<P>(W, prm, C0, R0, ci, sm, lane);
            __syncwarp();
            const int keep0 = max(R0, ci - (kTileKeep - 1));  // rows above the block are persisted, rows above keep0 inside it are not
            wk.tv = TileView{keep0, ci, C0, R0, keep0 == R0 ? R0 : keep0 + kNear};
        }
        if (lane == 0) {
            // walk while the current cell is on the boundary or inside the recomputed tile
            while (ci >= 0 && (ci == 0 || cj == 0 || wk.walkable(ci, cj)) && len < cap) {
                const int M = wk.cM(ci, cj);
                if (comp == 0) {
                    for (int k = 0; k < P; ++k) {
                        if (M == wk.cI(ci, cj, k)) { comp = k + 1; break; }
                        if (M == wk.cD(ci, cj, k)) { comp = -k - 1; break; }
                    }
                }
                int ni = -1, nj = -1;
                int32_t* o = aln + 2 * (int64_t)(cap - 1 - len);
                if (comp == 0) {
                    o[0] = ci - 1; o[1] = cj - 1;
                    const uint32_t ri = wk.rinfo(ci), cf = wk.cinfo(cj);
                    const PredList p1 = pred_list(ri, ci, W.poff1, W.pidx1), p2 = pred_list(cf, cj, W.poff2, W.pidx2);
                    const int sub = ((ri & kInfoLabelMask) == (cf & kInfoLabelMask)) ? prm.match : -prm.mismatch;
                    for (int a = 0; a < p1.n; ++a) {  // last prev1 with a match wins, with its first prev2
                        const int p = p1.at(a);
                        for (int b = 0; b < p2.n; ++b) {
                            const int q = p2.at(b);
                            if (wk.cM(p, q) + sub == M) { ni = p; nj = q; break; }
                        }
                    }
                } else if (comp > 0) {
                    o[0] = ci - 1; o[1] = -1;
                    const int k = comp - 1;
                    const int cur = wk.cI(ci, cj, k);
                    const PredList p1 = pred_list(wk.rinfo(ci), ci, W.poff1, W.pidx1);
                    for (int a = 0; a < p1.n; ++a) {
                        const int p = p1.at(a);
                        if (cur == wk.cM(p, cj) - prm.oe[k]) { comp = 0; ni = p; nj = cj; break; }
                        if (cur == wk.cIext(p, cj, k)) { ni = p; nj = cj; break; }
                    }
                } else {
                    o[0] = -1; o[1] = cj - 1;
                    const int k = -comp - 1;
                    const int cur = wk.cD(ci, cj, k);
                    const PredList p2 = pred_list(wk.cinfo(cj), cj, W.poff2, W.pidx2);
                    for (int b = 0; b < p2.n; ++b) {
                        const int q = p2.at(b);
                        if (cur == wk.cM(ci, q) - prm.oe[k]) { comp = 0; ni = ci; nj = q; break; }
                        if (cur == wk.cD(ci, q, k) - prm.e[k]) { ni = ci; nj = q; break; }
                    }
                }
                ++len;
                ci = ni; cj = nj;
                if (ci == 0 && cj == 0) ci = -1;  // unreachable on valid input; the corner ends every path
            }
            if (len >= cap) ci = -1;
        }
        ci = __shfl_sync(kFull, ci, 0);
        cj = __shfl_sync(kFull, cj, 0);
    }
    if (lane == 0) *len_out = (uint32_t)len;
}

// ------------------------------------------------------------------------------------------
// Persistent kernel.  Each CTA works through its own sequence of windows (pulled from the
// global queue, largest first) with kFillWarps fill warps and one traceback warp:
//   * fill warps take strips round-robin by a CTA-wide running strip number, so they flow
//     from one window into the next without any CTA-wide barrier;
//   * the traceback warp fetches windows (which also hands out the workspace slot: two slots
//     per CTA, window k uses slot k&1), waits until all strips of a window are filled, then
//     traces it back while the fill warps are already on the next window.
// ------------------------------------------------------------------------------------------
constexpr int kFillWarps = kWarps - 1;
constexpr int kWideMinCols = 384;  // windows at least this wide use 128-column strips
union FillSmemAny {
    FillSmem narrow;
    FillSmemWide wide;
};
constexpr int kTileInt4 = (int)(sizeof(TileSmem) / sizeof(int4));

struct CtaState {
    Win win[2];
    unsigned long long progress[kProgMask + 1];
    int seq_win[8];      // window id of the CTA's k-th window (-1 = queue exhausted)
    int seq_nstrips[8];
    int seq_tag[8];      // k+1 once entry k is published
    int strips_done[2];  // per workspace slot: finished strips (or tiles, for a tiled window)
    int tile_next[2];    // per workspace slot: next tile of a tiled window
    int warps_left[2];   // per workspace slot: fill warps that are done with the slot's window (strips or not); the slot and
                         // its window view are handed to the next window only when all of them have left
};

// Tiled windows.  A strip that is slower than its neighbours (far predecessor columns, the generic step) holds back
// every strip behind it for its whole length when strips are filled top to bottom in one piece.  Wide windows with
// more than panel_rows rows are therefore cut into panels: a tile = (panel p, strip cs), handed out from a per-window
// queue in the order of d = p * kTileSkew + cs (p ascending inside d).  Both predecessors of a tile, (p, cs-1) and
// (p-1, cs), have a smaller d, so they were taken earlier and no wait can deadlock; with about one tile per panel in
// flight, a tile's left neighbour is usually finished, and nothing waits for a slow strip.
constexpr int kTileSkew = kWarps - 1;
constexpr int kMaxTiledStrips = kProgMask - 64;
__device__ __forceinline__ bool window_tiled(const Win& W, int nstrips, int panel_cfg) {
    const int panel_rows = panel_rows_for(W.n1, panel_cfg);
    return W.cw == kWideCols && panel_rows > 0 && W.n1 > panel_rows + panel_rows / 4 && nstrips <= kMaxTiledStrips;
}

__device__ __forceinline__ unsigned smid() {
    unsigned v;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
    return v;
}
__global__ void nsmid_kernel(unsigned* out) {
    unsigned v;
    asm volatile("mov.u32 %0, %%nsmid;" : "=r"(v));
    *out = v;
}
__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

template <int P>
__global__ void __launch_bounds__(kThreads, 1) popoa_kernel(const __grid_constant__ LaunchArgs A) {
    extern __shared__ int4 smem[];
    __shared__ CtaState S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Params& prm = A.prm;  // __grid_constant__: the parameters are constant-bank operands, not registers
    for (int i = tid; i <= kProgMask; i += kThreads) S.progress[i] = 0ull;
    if (tid < 8) S.seq_tag[tid] = 0;
    if (tid < 2) S.warps_left[tid] = 0;
    __syncthreads();

    if (warp < kFillWarps) {
        // ================= fill warps =================
        CLB_WAIT_BEGIN;
        FillSmemAny& sm = *reinterpret_cast<FillSmemAny*>(smem + kTileInt4 + warp * (int)(sizeof(FillSmemAny) / sizeof(int4)));
        int G = 0;  // running strip number at the start of window k
        for (int k = 0;; ++k) {
            {
                CLB_WAIT_BEGIN;
                while (ld_volatile(&S.seq_tag[k & 7]) != k + 1) __nanosleep(200);
                CLB_WAIT_END(2);
            }
            __threadfence_block();
            const int w = ld_volatile(&S.seq_win[k & 7]);
            if (w < 0) break;
            const int nstrips = ld_volatile(&S.seq_nstrips[k & 7]);
            const Win& W = S.win[k & 1];
            // One work loop for both schedules (a single call site per fill routine: the step code exists once).
            // Tiled windows hand out (panel, strip) tiles from the per-window queue; the others give every warp the
            // strips cs = first, first + kFillWarps, ... top to bottom.
            const bool tiled = window_tiled(W, nstrips, A.panel_rows);
            const int H = tiled ? panel_rows_for(W.n1, A.panel_rows) : max(W.n1, 1);
            const int T = tiled ? (W.n1 + H - 1) / H : 1, ntiles = T * nstrips;
            int d = 0, p = 0, qi = 0;  // enumeration cursor of the tile queue: tile number qi is (p, d - p * kTileSkew)
            int cs = (warp - G % kFillWarps + kFillWarps) % kFillWarps;  // my first strip of an untiled window
            for (;;) {
                if (tiled) {
                    int q = 0;
                    if (lane == 0) q = atomicAdd(&S.tile_next[k & 1], 1);
                    q = __shfl_sync(kFull, q, 0);
                    if (q >= ntiles) break;
                    while (qi < q) {
                        do {
                            if (++p >= T) { p = 0; ++d; }
                        } while (d - p * kTileSkew < 0 || d - p * kTileSkew >= nstrips);
                        ++qi;
                    }
                    cs = d - p * kTileSkew;
                } else if (cs >= nstrips) {
                    break;
                }
                if (W.cw == kWideCols)
                    fill_strip_wide<P>(W, prm, cs, G + cs, *reinterpret_cast<FillSmemWide*>(&sm), S.progress, lane, A.start_lag,
                                       p * H, min((p + 1) * H, W.n1), A.debug_flags);
                else
                    fill_strip<P>(W, prm, cs, G + cs, *reinterpret_cast<FillSmem*>(&sm), S.progress, lane, A.start_lag);
                if (lane == 0) {
                    __threadfence_block();
                    atomicAdd(&S.strips_done[k & 1], 1);
                }
                __syncwarp();
                if (!tiled) cs += kFillWarps;
            }
            __syncwarp();
            if (lane == 0) atomicAdd(&S.warps_left[k & 1], 1);  // this warp will not look at S.win[k & 1] / the k-th queue entry again
            G += nstrips;
        }
        CLB_WAIT_END(4);
    } else {
        // ================= traceback warp =================
        int fetched = 0;
        bool ended = false;
        auto fetch = [&]() {
            const int k = fetched++;
            int w = -1;
            if (lane == 0) {
                if (k >= 2) {  // every fill warp must have left window k-2 before its view and workspace slot are reused
                    while (ld_volatile(&S.warps_left[k & 1]) < kFillWarps) __nanosleep(200);
                }
                S.warps_left[k & 1] = 0;
                const int qi = atomicAdd(A.queue, 1);
                w = qi < A.n_windows ? A.order[qi] : -1;
                int nstrips = 0;
                if (w >= 0) {
                    const WindowMeta m = A.meta[w];
                    Win& W = S.win[k & 1];
                    W.n1 = (int)m.n1; W.n2 = (int)m.n2; W.nsnk1 = (int)m.nsnk1; W.nsnk2 = (int)m.nsnk2;
                    W.info1 = A.s1.info + m.node1; W.info2 = A.s2.info + m.node2;
                    W.slot1 = A.s1.slot + m.node1; W.slot2 = A.s2.slot + m.node2;
                    W.depth1 = A.s1.depth + m.node1; W.depth2 = A.s2.depth + m.node2;
                    W.poff1 = A.s1.poff + m.poff1; W.poff2 = A.s2.poff + m.poff2;
                    W.pidx1 = A.s1.pidx + m.pidx1; W.pidx2 = A.s2.pidx + m.pidx2;
                    W.snk1 = A.s1.sinks + m.snk1; W.snk2 = A.s2.sinks + m.snk2;
                    // one CTA per SM at a time (217 KB of shared memory each), so the SM id names a free slot pair
                    const int64_t pair = A.slot_by_smid ? (int64_t)smid() : (int64_t)blockIdx.x;
                    int4* ws = reinterpret_cast<int4*>(A.workspace + (pair * 2 + (k & 1)) * A.slot_bytes);
                    W.rowbuf = ws;
                    W.colbuf = W.rowbuf + (int64_t)m.nrslot * row_stride(m.n2);
                    W.brow = W.colbuf + (int64_t)m.ncslot * (m.n1 + 1);
                    W.bcol = W.brow + (m.n2 + 1);
                    W.coleff = reinterpret_cast<int*>(W.bcol + (m.n1 + 1));
                    W.out = m.out;
                    W.id = w;
                    W.cw = (int)m.n2 >= kWideMinCols ? kWideCols : 1;
                    const int sw = kStrip * W.cw;
                    nstrips = m.n1 >= 1 ? (int)((m.n2 + sw - 1) / sw) : 0;
                }
                S.strips_done[k & 1] = 0;
                S.tile_next[k & 1] = 0;
                S.seq_nstrips[k & 7] = nstrips;
                S.seq_win[k & 7] = w;
                __threadfence_block();
                *reinterpret_cast<volatile int*>(&S.seq_tag[k & 7]) = k + 1;
            }
            w = __shfl_sync(kFull, w, 0);
            if (w < 0) ended = true;
        };
        fetch();
        if (!ended) fetch();
        for (int t = 0;; ++t) {
            const int w = ld_volatile(&S.seq_win[t & 7]);  // written by this warp
            if (w < 0) break;
            const int nstrips = ld_volatile(&S.seq_nstrips[t & 7]);
            const Win& W = S.win[t & 1];
            int nwait = nstrips;
            if (window_tiled(W, nstrips, A.panel_rows)) {
                const int H = panel_rows_for(W.n1, A.panel_rows);
                nwait *= (W.n1 + H - 1) / H;
            }
            {
                CLB_WAIT_BEGIN;
                while (ld_volatile(&S.strips_done[t & 1]) < nwait) __nanosleep(500);
                CLB_WAIT_END(3);
            }
            __threadfence_block();
            CLB_WAIT_BEGIN;
            tb_boundary<P>(W, prm, lane);
#ifdef CLB_PROFILE
            if (!(A.debug_flags & 1))
#endif
                traceback<P>(W, prm, *reinterpret_cast<TileSmem*>(smem), lane, A.score + w, A.aln + 2 * W.out, A.aln_len + w);
            __syncwarp();
            CLB_WAIT_END(5);
            if (!ended) fetch();  // hands slot t&1 to window t+2
        }
    }
}

int popoa_nsmid() {  // size of the %smid id space (>= number of SMs)
    unsigned* d = nullptr;
    unsigned h = 0;
    if (cudaMalloc(&d, 4) != cudaSuccess) return -1;
    nsmid_kernel<<<1, 1>>>(d);
    const cudaError_t e = cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? (int)h : -1;
}
#ifdef CLB_PROFILE
void popoa_wait_profile(bool reset) {
    unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (!reset) {
        cudaMemcpyFromSymbol(h, g_wait_cycles, sizeof(h));
        const double fill = (double)h[4] > 0 ? (double)h[4] : 1.0;
        fprintf(stderr, "[clb] fill-warp cycles %.3g: waiting for the left strip %.1f %%, for the panel above %.1f %%, for the next window %.1f %%; "
                        "traceback warp: waiting for the fill %.3g cycles, busy %.3g cycles (%.1f %% of a fill warp's time)\n",
                fill, 100.0 * h[0] / fill, 100.0 * h[1] / fill, 100.0 * h[2] / fill, (double)h[3], (double)h[5], 100.0 * h[5] / (fill / kFillWarps));
    }
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_wait_cycles, z, sizeof(z));
}
#else
void popoa_wait_profile(bool) {}
#endif

int popoa_smem_bytes() { return (kTileInt4 + kFillWarps * (int)(sizeof(FillSmemAny) / sizeof(int4))) * (int)sizeof(int4); }
int popoa_threads() { return kThreads; }

cudaError_t launch_popoa(int num_pw, const LaunchArgs& args, int grid, cudaStream_t stream) {
    int smem = popoa_smem_bytes();
#ifdef CLB_PROFILE
    if (const char* e = getenv("CLB_SMEM_PAD_KB")) smem = std::min(smem + 1024 * atoi(e), 227 * 1024 - 6 * 1024);  // experiment: shrink the L1 share
#endif
    cudaError_t err;
    switch (num_pw) {
        case 1:
            err = cudaFuncSetAttribute(popoa_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (err != cudaSuccess) return err;
            popoa_kernel<1><<<grid, kThreads, smem, stream>>>(args);
            break;
        case 2:
            err = cudaFuncSetAttribute(popoa_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (err != cudaSuccess) return err;
            popoa_kernel<2><<<grid, kThreads, smem, stream>>>(args);
            break;
        case 3:
            err = cudaFuncSetAttribute(popoa_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (err != cudaSuccess) return err;
            popoa_kernel<3><<<grid, kThreads, smem, stream>>>(args);
            break;
        default:
            return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// INT32 issue-rate probe: independent add/max chains, no memory traffic.  use_dpx=1 issues the
// fused DPX form (VIADDMNMX), use_dpx=0 plain IADD + IMNMX pairs.  One "op" = one add or max.
// ------------------------------------------------------------------------------------------
template <bool DPX>
__global__ void __launch_bounds__(256) int32_probe_kernel(int* out, int iters, int seed) {
    int a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + k * 7 + threadIdx.x;
    const int d = seed | 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (DPX) a[k] = __viaddmax_s32(a[k], d, a[(k + 1) & 7]);
                else a[k] = max(a[k] + d, a[(k + 1) & 7] - it);
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= a[k];
    if (s == 0x7fffffff) out[0] = s;
}

double int32_probe(int use_dpx, int sm_count) {
    int* d = nullptr;
    if (cudaMalloc(&d, 4) != cudaSuccess) return -1.0;
    const int iters = 4096, grid = sm_count * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        if (use_dpx) int32_probe_kernel<true><<<grid, threads>>>(d, iters, 12345);
        else int32_probe_kernel<false><<<grid, threads>>>(d, iters, 12345);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.f; break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (best <= 0) return -1.0;
    // per thread per iteration: 4*8 updates, each 1 add + 1 max (DPX fuses them) -> 2 ops; the plain
    // variant has one more subtraction per update that is not counted.
    const double ops = (double)grid * threads * iters * 4.0 * 8.0 * 2.0;
    return ops / (best * 1e-3) / 1e12;
}

}  // namespace clb
