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

#include <type_traits>

#include "popoa_device.cuh"

namespace clb {

constexpr int kWarps = 12;
constexpr int kThreads = kWarps * 32;
constexpr unsigned kFull = 0xffffffffu;

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
    const uint32_t rstride = (uint32_t)n2 + 1u, cstride = (uint32_t)n1 + 1u;  // host guarantees 32-bit offsets

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
        rowbuf[j] = make_int4(upM, kMinInf, kMinInf, kMinInf);
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
                            max4(eM, eI, rowbuf[(uint32_t)slot1[p] * rstride + (uint32_t)j]);
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
                I[k] = __viaddmax_s32(eI[k], -prm.e[k], eM - prm.oe[k]);
                D[k] = __viaddmax_s32(lD[k], -prm.e[k], lM - prm.oe[k]);
                M = __vimax3_s32(M, I[k], D[k]);
            }
            const int4 cellA = make_int4(M, I[0], I[1], I[2]);
            myA[rs] = cellA;
            myB[rs] = make_int4(eM, D[0], D[1], D[2]);
            if (rinfo & persist_mask) {
                uint32_t slot = rinfo >> kInfoSlotShift;
                if (slot == kInfoSlotEscape) slot = (uint32_t)slot1[r];
                rowbuf[slot * rstride + (uint32_t)j] = cellA;
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
// DP fill, wide variant: strips of 128 columns, FOUR consecutive columns per lane.
// The per-step overhead (shuffles, row info, loop, waits) is shared by 128 cells instead of 32,
// a column's near predecessors inside the lane are plain registers, and the ones that fall into
// the previous lane come over warp shuffles (lane-1 finished the same row one step earlier), so
// no shared-memory traffic is needed for columns at all; rows still use a small ring.
// ------------------------------------------------------------------------------------------
constexpr int kWideCols = 4;
struct __align__(16) FillSmemWide {
    int4 ringA[kFillRing * kWideCols * 32];  // [row & 3][c][lane]: {M, I_k} of the last rows, own columns
    int4 leftv[3][2][16];                    // prefetched {M, D_k} of the 3 columns left of the strip, 16-row blocks
    int lefte[3][2][16];
};

struct ColState {  // what a column offers to the columns right of it, for the current row
    int M, D[3], E;  // E = diagonal input = max over predecessor rows of M
};

template <int P>
__device__ __forceinline__ void fill_strip_wide(const Win& Wsh, const Params& prm, const int cs, const int g,
                                                FillSmemWide& sm, volatile unsigned long long* progress, const int lane,
                                                const int start_lag, const int R0, const int R1, const int dbg) {
    // Rows R0+1 .. R1 of the strip (a panel; R0 = 0, R1 = n1 is the whole strip).  A panel that does not start at
    // the top takes over from whoever filled the panel above through the window workspace: rows R0-2 .. R0 are
    // persisted (host: popoa_host.cu) and the strip's own progress word says when they are there.
    constexpr int H = kFillRing, C = kWideCols, W = 32 * C, PB = 16;  // PB = rows per left-column prefetch block
    const int C0 = 1 + W * cs;
    const int n1 = Wsh.n1, n2 = Wsh.n2;
    const uint32_t* __restrict__ info1 = Wsh.info1;
    const uint32_t* __restrict__ poff1 = Wsh.poff1;
    const uint32_t* __restrict__ pidx1 = Wsh.pidx1;
    const int32_t* __restrict__ slot2 = Wsh.slot2;
    int4* const rowbuf = Wsh.rowbuf;
    int4* const colbuf = Wsh.colbuf;
    int* const coleff = Wsh.coleff;
    const uint32_t rstride = (uint32_t)n2 + 1u, cstride = (uint32_t)n1 + 1u;
    const int j0 = C0 + C * lane;

    // ---- per-column constants ----
    uint32_t cinfo[C], coff[C];
    int upM[C], upI[C][3];
    uint32_t nvalid = 0;  // number of this lane's columns that exist
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int j = j0 + c;
        const bool jv = j <= n2;
        nvalid += jv ? 1u : 0u;
        cinfo[c] = jv ? Wsh.info2[j] : (kInfoRegular | (1u << kInfoNearShift) | 0xffu);
        coff[c] = (jv && (cinfo[c] & kInfoPersist)) ? (uint32_t)slot2[j] * cstride : 0xffffffffu;
        upM[c] = kMinInf; upI[c][0] = upI[c][1] = upI[c][2] = kMinInf;
        if (jv && R0 == 0) {  // boundary row (alignment.hpp:814-894): M(0,j), I_k(0,j) = -inf
            upM[c] = boundary_cell<P>(Wsh.depth2[j], prm).x;
            rowbuf[j] = make_int4(upM[c], kMinInf, kMinInf, kMinInf);
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
            if (d > c && ((cinfo[c] >> (kInfoNearShift + d - 1)) & 1u)) nb |= 1u << (d - c - 1);
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
            for (;;) {
                const unsigned long long v = progress[(g - 1) & kProgMask];
                avail = ((int)(v >> 32) == g) ? (int)(v & 0xffffffffu) : 0;
                if (avail >= want) break;
                __nanosleep(kPollNs);
            }
            __threadfence_block();
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
    // never wait for rows below the panel: the tile that fills them may be queued behind this one
    wait_rows(min(R0 + max(start_lag, PB), R1));
    if (R0 > 0) {  // the panel above must be complete before anything of this one is read (strip 0 has no other wait)
        for (;;) {
            const unsigned long long v = progress[g & kProgMask];
            if ((int)(v >> 32) == g + 1 && (int)(v & 0xffffffffu) >= R0) break;
            __nanosleep(200);
        }
        __threadfence_block();
    }
    prefetch_block(R0 / PB);
    if (R0 > 0) {  // take over rows R0-2 .. R0 from the panel above
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
                    if (t == 0) { upM[c] = v.x; upI[c][0] = v.y; upI[c][1] = v.z; upI[c][2] = v.w; }
                }
        }
    }
    cp_async_wait_all();
    __syncwarp();

    ColState out[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { out[c].M = kMinInf; out[c].D[0] = out[c].D[1] = out[c].D[2] = kMinInf; out[c].E = kMinInf; }
    uint32_t rinfo_next = (lane == 0) ? info1[R0 + 1] : 0u;
    // explicit 32-bit shared-window addresses: keeps ptxas from re-deriving them every step
    uint32_t saA = (uint32_t)__cvta_generic_to_shared(sm.ringA) + (uint32_t)lane * 16u;
    uint32_t saLv = (uint32_t)__cvta_generic_to_shared(&sm.leftv[0][0][0]);
    uint32_t saLe = (uint32_t)__cvta_generic_to_shared(&sm.lefte[0][0][0]);
    // opaque copies: stops ptxas from re-deriving the addresses (S2R + LEA + IMAD) inside the row loop
    asm volatile("mov.u32 %0, %0;" : "+r"(saA));
    asm volatile("mov.u32 %0, %0;" : "+r"(saLv));
    asm volatile("mov.u32 %0, %0;" : "+r"(saLe));
    // per-column constant predicates
    bool cb1[C], cb2[C], crare[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        cb1[c] = (cinfo[c] & ((1u << kInfoNearShift) | kInfoRegular)) != 0 &&
                 ((cinfo[c] & (1u << kInfoNearShift)) != 0 || (cinfo[c] & kInfoRegular) != 0);
        // a regular column always takes its distance-1 neighbour (column 1's is the boundary column)
        cb2[c] = (cinfo[c] & (2u << kInfoNearShift)) != 0;
        crare[c] = (cinfo[c] & ((4u << kInfoNearShift) | kInfoFar)) != 0;
    }

    // Far predecessor columns (long bubbles, the boundary column of a source).  The common shape -- one such
    // column in the lane, with ONE far predecessor that lies in an earlier strip or at least two lanes back --
    // takes a fast path: the predecessor's persisted entry {M, D_k} / diagonal input for row r+1 is requested
    // while row r is being computed (it was written a step earlier at the latest), so the next step folds it in
    // without a list walk and without a load on the column chain.  Anything else keeps the generic walk.
    int farc = -1;
    const int4* farv = colbuf;
    const int* fare = coleff;
    bool slowfar[C], cx[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        slowfar[c] = false;
        if (cinfo[c] & kInfoFar) {  // never set for a column beyond n2
            const int j = j0 + c;
            int nf = 0, qsel = 0;
            const uint32_t b1 = Wsh.poff2[j + 1];
#pragma unroll 1
            for (uint32_t b = Wsh.poff2[j]; b < b1; ++b) {
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
                slowfar[c] = true;
            }
        }
        crare[c] = (cinfo[c] & (4u << kInfoNearShift)) != 0 || slowfar[c];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) cx[c] = crare[c] || farc == c;
    int4 Fv = make_int4(kMinInf, kMinInf, kMinInf, kMinInf);
    int Fe = kMinInf;

    // warp-uniform: does any lane of this strip persist one of its first three columns?
    bool lane_pers012 = false;
#pragma unroll
    for (int c = 0; c < C - 1; ++c) lane_pers012 |= coff[c] != 0xffffffffu;
    const bool strip_pers012 = __any_sync(kFull, lane_pers012);
    // A plain strip -- no column with a distance-3 / far predecessor -- runs a lean step without any of that
    // code: the four columns of a lane become one basic block (measured on windows without bubbles: 297 -> 368 GCUPS).
    bool lane_cx = false;
#pragma unroll
    for (int c = 0; c < C; ++c) lane_cx |= cx[c];
    const bool lean = !__any_sync(kFull, lane_cx) && !(needS & 4u) && !(needL & 4u) && !(dbg & 2);

    auto shfl_col = [&](const ColState& v) {
        ColState o;
        o.M = __shfl_up_sync(kFull, v.M, 1);
#pragma unroll
        for (int k = 0; k < 3; ++k) o.D[k] = (k < P) ? __shfl_up_sync(kFull, v.D[k], 1) : kMinInf;
        o.E = __shfl_up_sync(kFull, v.E, 1);
        return o;
    };
    auto fold = [&](ColState& a, const ColState& b) {
        a.M = imax(a.M, b.M);
#pragma unroll
        for (int k = 0; k < P; ++k) a.D[k] = imax(a.D[k], b.D[k]);
        a.E = imax(a.E, b.E);
    };
    // all lanes read lane 0's left-column entry (a broadcast, no divergence); lane 0 keeps it
    auto take_left = [&](ColState& S, int d, int s) {
        const uint32_t slot = (uint32_t)(d * 2 + ((s / PB) & 1)) * PB + (uint32_t)(s % PB);
        const int4 t = lds_128(saLv + slot * 16u);
        const int e = lds_32(saLe + slot * 4u);
        if (lane == 0) { S.M = t.x; S.D[0] = t.y; S.D[1] = t.z; S.D[2] = t.w; S.E = e; }
    };

    auto step = [&](const int s, auto guard_tag, auto lean_tag) {
        constexpr bool GUARD = decltype(guard_tag)::value;
        constexpr bool LEAN = decltype(lean_tag)::value;
        const int r = 1 + s - lane;
        bool act = true;
        if (GUARD) act = r > R0 && r <= R1;
        // previous lane's columns for this row (it finished the row one step ago)
        ColState S0, S1, S2;  // its column 3, 2, 1
        S0 = shfl_col(out[3]);
        if (needS & 2u) S1 = shfl_col(out[2]);
        if (!LEAN && (needS & 4u)) S2 = shfl_col(out[1]);
        // lane 0 is on row s+1: rows PB*b+1.. live in buffer b&1 at index (row-1) % PB
        if (needL & 1u) take_left(S0, 0, s);
        if (needL & 6u) {
            if (needL & 2u) take_left(S1, 1, s);
            if (!LEAN && (needL & 4u)) take_left(S2, 2, s);
        }
        const uint32_t rinfo = rinfo_next;
        if (GUARD) rinfo_next = (r >= R0 && r < R1) ? info1[(uint32_t)(r + 1)] : 0u;
        else rinfo_next = info1[(uint32_t)(r + 1)];  // the info array is padded by one entry
        if (act) {
            const uint32_t rsA = saA + (uint32_t)(r & (H - 1)) * (C * 32 * 16);
            const int rlabel = (int)(rinfo & kInfoLabelMask);
            // ---- effective predecessor row, folded in place into the up registers ----
            if (!(rinfo & kInfoRegular)) {
                const bool rb1 = (rinfo & (1u << kInfoNearShift)) != 0;
                const uint32_t rs2 = saA + (uint32_t)((r - 2) & (H - 1)) * (C * 32 * 16);  // row r-2
                if (rinfo & (2u << kInfoNearShift)) {
                    if (!rb1) {  // second allele of a SNP bubble: row r-2 replaces row r-1
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const int4 t = lds_128(rs2 + c * (32 * 16));
                            upM[c] = t.x; upI[c][0] = t.y; upI[c][1] = t.z; upI[c][2] = t.w;
                        }
                    } else {     // node after the bubble: both
#pragma unroll
                        for (int c = 0; c < C; ++c) max4(upM[c], upI[c], lds_128(rs2 + c * (32 * 16)));
                    }
                } else if (!rb1) {
#pragma unroll
                    for (int c = 0; c < C; ++c) { upM[c] = kMinInf; upI[c][0] = upI[c][1] = upI[c][2] = kMinInf; }
                }
                if (rinfo & ((4u << kInfoNearShift) | kInfoFar)) {
                    if (rinfo & (4u << kInfoNearShift)) {
                        const uint32_t rs3 = saA + (uint32_t)((r - 3) & (H - 1)) * (C * 32 * 16);
#pragma unroll
                        for (int c = 0; c < C; ++c) max4(upM[c], upI[c], lds_128(rs3 + c * (32 * 16)));
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
                                if (c < (int)nvalid) max4(upM[c], upI[c], rowbuf[o + c]);
                        }
                    }
                }
            }
            ColState cur[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                // ---- effective predecessor column: distance 1 and 2 branch-free, the rest rarely ----
                const ColState& d1 = (c == 0) ? S0 : cur[c > 0 ? c - 1 : 0];
                const ColState& d2 = (c >= 2) ? cur[c >= 2 ? c - 2 : 0] : (c == 1 ? S0 : S1);
                ColState L;
                L.M = cb1[c] ? d1.M : kMinInf;
#pragma unroll
                for (int k = 0; k < 3; ++k) L.D[k] = (k < P && cb1[c]) ? d1.D[k] : kMinInf;
                L.E = cb1[c] ? d1.E : kMinInf;
                if (cb2[c]) fold(L, d2);
                if (!LEAN && cx[c]) {
                  if (farc == c) {  // fast far column
                    if (r == R0 + 1) { Fv = farv[r]; Fe = fare[r]; }
                    L.M = imax(L.M, Fv.x); L.D[0] = imax(L.D[0], Fv.y); L.D[1] = imax(L.D[1], Fv.z); L.D[2] = imax(L.D[2], Fv.w);
                    L.E = imax(L.E, Fe);
                    const int rn = min(r + 1, R1);  // not below the panel: those rows may not exist yet
                    Fv = farv[rn]; Fe = fare[rn];
                  }
                  if (crare[c]) {
                    const uint32_t ci = cinfo[c];
                    if (ci & (4u << kInfoNearShift)) {
                        const ColState& d3 = (c >= 3) ? cur[0] : (c == 2 ? S0 : (c == 1 ? S1 : S2));
                        fold(L, d3);
                    }
                    if (slowfar[c]) {
                        const int j = j0 + c;
                        const uint32_t b1 = Wsh.poff2[j + 1];
#pragma unroll 1
                        for (uint32_t b = Wsh.poff2[j]; b < b1; ++b) {
                            const int q = (int)Wsh.pidx2[b];
                            if (q >= 1 && j - q <= kNear) continue;
                            const uint32_t o = (uint32_t)Wsh.slot2[q] * cstride + (uint32_t)r;
                            const int4 t = colbuf[o];
                            L.M = imax(L.M, t.x); L.D[0] = imax(L.D[0], t.y); L.D[1] = imax(L.D[1], t.z); L.D[2] = imax(L.D[2], t.w);
                            L.E = imax(L.E, coleff[o]);
                        }
                    }
                  }
                }
                // ---- the cell ----
                const int sub = (rlabel == (int)(cinfo[c] & kInfoLabelMask)) ? prm.match : -prm.mismatch;
                const int eM = upM[c];
                int I[3] = {kMinInf, kMinInf, kMinInf};
                cur[c].D[0] = cur[c].D[1] = cur[c].D[2] = kMinInf;
                int M = __viaddmax_s32(L.E, sub, kMinInf);
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    I[k] = __viaddmax_s32(upI[c][k], -prm.e[k], eM - prm.oe[k]);
                    cur[c].D[k] = __viaddmax_s32(L.D[k], -prm.e[k], L.M - prm.oe[k]);
                    M = __vimax3_s32(M, I[k], cur[c].D[k]);
                }
                cur[c].M = M;
                cur[c].E = eM;
                const int4 cellA = make_int4(M, I[0], I[1], I[2]);
                sts_128(rsA + c * (32 * 16), cellA);
                if (c == C - 1 && coff[c] != 0xffffffffu) {  // the regular persisted columns (every 32nd) are a lane's last
                    colbuf[coff[c] + (uint32_t)r] = make_int4(M, cur[c].D[0], cur[c].D[1], cur[c].D[2]);
                    coleff[coff[c] + (uint32_t)r] = eM;
                }
                upM[c] = M; upI[c][0] = I[0]; upI[c][1] = I[1]; upI[c][2] = I[2];
            }
            if (strip_pers012) {  // warp-uniform and rare: a column with a far successor among a lane's first three
#pragma unroll
                for (int c = 0; c < C - 1; ++c)
                    if (coff[c] != 0xffffffffu) {
                        colbuf[coff[c] + (uint32_t)r] = make_int4(cur[c].M, cur[c].D[0], cur[c].D[1], cur[c].D[2]);
                        coleff[coff[c] + (uint32_t)r] = cur[c].E;
                    }
            }
            if (rinfo & kInfoPersist) {  // persisted row: the new row is in the up registers
                uint32_t rslot = rinfo >> kInfoSlotShift;
                if (rslot == kInfoSlotEscape) rslot = (uint32_t)Wsh.slot1[r];
                int4* const rp = rowbuf + (size_t)(rslot * rstride + (uint32_t)j0);  // one address, immediate offsets
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (c < (int)nvalid) rp[c] = make_int4(upM[c], upI[c][0], upI[c][1], upI[c][2]);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) out[c] = cur[c];
        }
        __syncwarp();
    };
    auto publish = [&](int r31) {
        __threadfence_block();
        progress[g & kProgMask] = ((unsigned long long)(g + 1) << 32) | (unsigned)r31;
    };

    const int nsteps = R1 + 31;
    for (int s0 = R0; s0 < nsteps; s0 += PB) {  // PB-step blocks; lane 0 is on rows s0+1 .. s0+PB (R0 is a multiple of PB)
        const int s1 = min(s0 + PB, nsteps);
        {  // only lane 0 reads the prefetch buffers, so the next block can be requested right away
            const int b = s0 / PB + 1;
            if (PB * b + 1 <= R1) {
                wait_rows(min(PB * b + PB, R1));
                prefetch_block(b);
            }
        }
        const bool inner = s0 >= R0 + 31 && s1 <= R1;
#pragma unroll 1
        for (int q8 = s0; q8 < s1; q8 += 8) {
            const int e8 = min(q8 + 8, s1);
            if (inner && lean) {  // exactly 8 steps: unrolled in pairs so that the row registers ping-pong instead of being copied
#pragma unroll 1
                for (int s = q8; s < q8 + 8; ++s) step(s, std::false_type{}, std::true_type{});
            } else {  // the guarded step doubles as the generic one (one copy of it: instruction cache)
#pragma unroll 1
                for (int s = q8; s < e8; ++s) step(s, std::true_type{}, std::false_type{});
            }
            if (lane == 31) {
                const int r31 = e8 - 31;
                if (r31 > R0) publish(min(r31, R1));
            }
        }
        cp_async_wait_all();
        __syncwarp();
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
    const int64_t rstride = (int64_t)W.n2 + 1, cstride = (int64_t)W.n1 + 1;
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
    int R0, R1, C0;  // tile rows R0..R1, columns C0..C0+31; R0 = 0 means "no tile"
};

// shared memory of the traceback warp
struct __align__(16) TileSmem {
    int4 A[kRowBlock * 32];       // {M, I_k}       [row & 63][lane]
    int4 B[kRowBlock * 32];       // {diag in, D_k} [row & 63][lane]
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
    constexpr int H = kRowBlock;
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
    const uint32_t rstride = (uint32_t)n2 + 1u, cstride = (uint32_t)n1 + 1u;
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
    int upM = kMinInf, upI[3] = {kMinInf, kMinInf, kMinInf};
    if (jvalid) {
        const int s0 = slot1[R0 - 1];
        if (s0 >= 0) {
            const int4 a = rowbuf[(uint32_t)s0 * rstride + (uint32_t)j];
            upM = a.x; upI[0] = a.y; upI[1] = a.z; upI[2] = a.w;
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
            // ---- effective predecessor row ----
            int eM = upM, eI[3] = {upI[0], upI[1], upI[2]};
            if (!(rinfo & kInfoRegular)) {
                if (!(rinfo & (1u << kInfoNearShift))) { eM = kMinInf; eI[0] = eI[1] = eI[2] = kMinInf; }
#pragma unroll
                for (int d = 2; d <= 3; ++d) {
                    if (rinfo & ((1u << (d - 1)) << kInfoNearShift)) {
                        const int p = r - d;
                        const int4 v = p >= R0 ? myA[(p & (H - 1)) * 32] : rowbuf[(uint32_t)slot1[p] * rstride + (uint32_t)j];
                        max4(eM, eI, v);
                    }
                }
                if (rinfo & kInfoFar) {
                    const uint32_t rp1 = poff1[r + 1];
#pragma unroll 1
                    for (uint32_t a = poff1[r]; a < rp1; ++a) {
                        const int p = (int)pidx1[a];
                        if (p >= 1 && r - p <= kNear) continue;
                        max4(eM, eI, rowbuf[(uint32_t)slot1[p] * rstride + (uint32_t)j]);
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
                I[k] = __viaddmax_s32(eI[k], -prm.e[k], eM - prm.oe[k]);
                D[k] = __viaddmax_s32(lD[k], -prm.e[k], lM - prm.oe[k]);
                M = __vimax3_s32(M, I[k], D[k]);
            }
            myA[rs] = make_int4(M, I[0], I[1], I[2]);
            myB[rs] = make_int4(eM, D[0], D[1], D[2]);
            upM = M; upI[0] = I[0]; upI[1] = I[1]; upI[2] = I[2];
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
    // M(i,j); the corner reads as -inf in every traceback test (it is never a match)
    __device__ int cM(int i, int j) const {
        if (i == 0) return j == 0 ? kMinInf : W.brow[j].x;
        if (j == 0) return W.bcol[i].x;
        if (in_tile(i, j)) return sm.A[(i & (kRowBlock - 1)) * 32 + (j - tv.C0)].x;
        const int s = W.slot1[i];
        if (s >= 0) return W.rowbuf[(uint32_t)s * rstride + (uint32_t)j].x;
        return W.colbuf[(uint32_t)W.slot2[j] * cstride + (uint32_t)i].x;
    }
    __device__ int cI(int i, int j, int k) const {
        if (i == 0) return kMinInf;
        int4 v;
        if (j == 0) v = W.bcol[i];
        else if (in_tile(i, j)) v = sm.A[(i & (kRowBlock - 1)) * 32 + (j - tv.C0)];
        else v = W.rowbuf[(uint32_t)W.slot1[i] * rstride + (uint32_t)j];
        return k == 0 ? v.y : (k == 1 ? v.z : v.w);
    }
    __device__ int cD(int i, int j, int k) const {
        if (j == 0) return kMinInf;
        int4 v;
        if (i == 0) v = W.brow[j];
        else if (in_tile(i, j)) v = sm.B[(i & (kRowBlock - 1)) * 32 + (j - tv.C0)];
        else v = W.colbuf[(uint32_t)W.slot2[j] * cstride + (uint32_t)i];
        return k == 0 ? v.y : (k == 1 ? v.z : v.w);
    }
    __device__ uint32_t rinfo(int i) const { return (tv.R0 > 0 && i >= tv.R0 && i <= tv.R1) ? sm.rinfo[i - tv.R0] : W.info1[i]; }
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
    const uint32_t rstride = (uint32_t)n2 + 1u;
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

    Walker<P> wk{W, prm, sm, TileView{0, 0, 0}, rstride, (uint32_t)n1 + 1u};
    const int cap = n1 + n2;
    int len = 0, comp = 0;
    while (ci >= 0) {  // warp-uniform: ci/cj are broadcast from lane 0 below
        if (ci >= 1 && cj >= 1 && !wk.in_tile(ci, cj)) {
            const int R0 = 1 + ((ci - 1) / kRowBlock) * kRowBlock;
            const int C0 = 1 + ((cj - 1) / kStrip) * kStrip;
            __syncwarp();
            tile_strip<P>(W, prm, C0, R0, ci, sm, lane);
            __syncwarp();
            wk.tv = TileView{R0, ci, C0};
        }
        if (lane == 0) {
            // walk while the current cell is on the boundary or inside the recomputed tile
            while (ci >= 0 && (ci == 0 || cj == 0 || wk.in_tile(ci, cj)) && len < cap) {
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
                        if (cur == wk.cI(p, cj, k) - prm.e[k]) { ni = p; nj = cj; break; }
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
};

// Tiled windows.  A strip that is slower than its neighbours (far predecessor columns, the generic step) holds back
// every strip behind it for its whole length when strips are filled top to bottom in one piece.  Wide windows with
// more than panel_rows rows are therefore cut into panels: a tile = (panel p, strip cs), handed out from a per-window
// queue in the order of d = p * kTileSkew + cs (p ascending inside d).  Both predecessors of a tile, (p, cs-1) and
// (p-1, cs), have a smaller d, so they were taken earlier and no wait can deadlock; with about one tile per panel in
// flight, a tile's left neighbour is usually finished, and nothing waits for a slow strip.
constexpr int kTileSkew = 11;
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
__global__ void __launch_bounds__(kThreads, 1) popoa_kernel(const LaunchArgs A) {
    extern __shared__ int4 smem[];
    __shared__ CtaState S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Params prm = A.prm;
    for (int i = tid; i <= kProgMask; i += kThreads) S.progress[i] = 0ull;
    if (tid < 8) S.seq_tag[tid] = 0;
    __syncthreads();

    if (warp < kFillWarps) {
        // ================= fill warps =================
        FillSmemAny& sm = *reinterpret_cast<FillSmemAny*>(smem + kTileInt4 + warp * (int)(sizeof(FillSmemAny) / sizeof(int4)));
        int G = 0;  // running strip number at the start of window k
        for (int k = 0;; ++k) {
            while (ld_volatile(&S.seq_tag[k & 7]) != k + 1) __nanosleep(200);
            __threadfence_block();
            const int w = ld_volatile(&S.seq_win[k & 7]);
            if (w < 0) break;
            const int nstrips = ld_volatile(&S.seq_nstrips[k & 7]);
            const Win& W = S.win[k & 1];
            if (window_tiled(W, nstrips, A.panel_rows)) {
                const int H = panel_rows_for(W.n1, A.panel_rows), T = (W.n1 + H - 1) / H, ntiles = T * nstrips;
                int d = 0, p = 0, qi = 0;  // enumeration cursor: tile number qi is (p, d - p * kTileSkew)
                for (;;) {
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
                    const int cs = d - p * kTileSkew;
                    fill_strip_wide<P>(W, prm, cs, G + cs, *reinterpret_cast<FillSmemWide*>(&sm), S.progress, lane, A.start_lag,
                                       p * H, min((p + 1) * H, W.n1), A.debug_flags);
                    if (lane == 0) {
                        __threadfence_block();
                        atomicAdd(&S.strips_done[k & 1], 1);
                    }
                    __syncwarp();
                }
                G += nstrips;
                continue;
            }
            int first = (warp - G % kFillWarps + kFillWarps) % kFillWarps;  // my first strip of this window
            for (int cs = first; cs < nstrips; cs += kFillWarps) {
                if (W.cw == kWideCols) fill_strip_wide<P>(W, prm, cs, G + cs, *reinterpret_cast<FillSmemWide*>(&sm), S.progress, lane, A.start_lag, 0, W.n1, A.debug_flags);
                else fill_strip<P>(W, prm, cs, G + cs, *reinterpret_cast<FillSmem*>(&sm), S.progress, lane, A.start_lag);
                if (lane == 0) {
                    __threadfence_block();
                    atomicAdd(&S.strips_done[k & 1], 1);
                }
                __syncwarp();
            }
            G += nstrips;
        }
    } else {
        // ================= traceback warp =================
        int fetched = 0;
        bool ended = false;
        auto fetch = [&]() {
            const int k = fetched++;
            int w = -1;
            if (lane == 0) {
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
                    W.colbuf = W.rowbuf + (int64_t)m.nrslot * (m.n2 + 1);
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
            while (ld_volatile(&S.strips_done[t & 1]) < nwait) __nanosleep(500);
            __threadfence_block();
            tb_boundary<P>(W, prm, lane);
            if (!(A.debug_flags & 1))
                traceback<P>(W, prm, *reinterpret_cast<TileSmem*>(smem), lane, A.score + w, A.aln + 2 * W.out, A.aln_len + w);
            __syncwarp();
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
int popoa_smem_bytes() { return (kTileInt4 + kFillWarps * (int)(sizeof(FillSmemAny) / sizeof(int4))) * (int)sizeof(int4); }
int popoa_threads() { return kThreads; }

cudaError_t launch_popoa(int num_pw, const LaunchArgs& args, int grid, cudaStream_t stream) {
    const int smem = popoa_smem_bytes();
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
