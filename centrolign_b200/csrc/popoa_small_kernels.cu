// popoa_small_kernels.cu -- the gap-fill DP for SMALL windows: one warp per window, the whole matrix in shared memory.
//
// The windows the reference's Stitcher really sends to po_poa are tiny (2 x 93 kbp HOR arrays: median 9 cells, p90 49,
// p99 169).  The strip / tile machinery of popoa_kernels.cu (a CTA of twelve warps per window,
// persisted rows and columns, tile recomputation in the traceback) is built for matrices of 10^6..10^8 cells and
// spends microseconds of fixed cost on a window of a dozen cells.  Here a window of at most kSmallCells matrix cells is
// given to ONE WARP:
//   * all seven values of every cell live in shared memory ({M, I_k} and {D_k} per cell), 12 KB per warp;
//   * the fill runs over anti-diagonals i + j = d: predecessors have smaller topological ranks, so the cells of a
//     diagonal are independent and a lane takes one cell, in pull form over the CSR predecessor lists -- the reference's
//     recurrence (include/centrolign/alignment.hpp:898-938) with no assumption about the graph shape;
//   * the traceback walks the stored values directly (alignment.hpp:979-1138: same tests, same order).
// Same inputs (the flattened batch of popoa_host.cu) and outputs as popoa_kernel; the host routes windows by size.
#include <cuda_runtime.h>
#include <limits.h>

#include "popoa_device.cuh"

namespace clb {

namespace {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ int smax(int a, int b) { return a > b ? a : b; }

template <int P>
__device__ __forceinline__ int4 small_boundary_cell(uint32_t depth, const Params& prm) {  // alignment.hpp:814-894, see popoa_kernels.cu
    int v[3] = {kMinInf, kMinInf, kMinInf};
    int m = kMinInf;
    if (depth) {
#pragma unroll
        for (int k = 0; k < P; ++k) {
            v[k] = smax(kMinInf, (int)(0u - (uint32_t)prm.oe[k] - (uint32_t)prm.e[k] * (depth - 1)));
            m = smax(m, v[k]);
        }
    }
    return make_int4(m, v[0], v[1], v[2]);
}

struct SmallWin {
    int n1, n2, W;  // W = n2 + 1: row stride of the matrix
    const uint32_t *info1, *info2, *poff1, *poff2, *pidx1, *pidx2, *snk1, *snk2;
    int nsnk1, nsnk2;
};

// predecessor list of a node in previous() order; regular nodes (one predecessor = index-1) need no memory access
struct SmallPreds {
    const uint32_t* ptr;
    int n, single;
    __device__ __forceinline__ int at(int k) const { return ptr ? (int)ptr[k] : single; }
};
__device__ __forceinline__ SmallPreds small_preds(uint32_t info, int idx, const uint32_t* poff, const uint32_t* pidx) {
    SmallPreds pl;
    if (info & kInfoRegular) {
        pl.ptr = nullptr; pl.n = 1; pl.single = idx - 1;
    } else {
        const uint32_t a = poff[idx];
        pl.ptr = pidx + a; pl.n = (int)(poff[idx + 1] - a); pl.single = 0;
    }
    return pl;
}

}  // namespace

template <int P>
__global__ void __launch_bounds__(kSmallWarps * 32) popoa_small_kernel(const __grid_constant__ LaunchArgs A, const int first, const int count) {
    extern __shared__ int4 small_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int4* const mA = small_smem + (size_t)warp * 2 * kSmallCells;  // {M, I_k}
    int4* const mB = mA + kSmallCells;                             // {-, D_k}
    const Params& prm = A.prm;
    const int nwarps = gridDim.x * kSmallWarps;
    for (int idx = first + blockIdx.x * kSmallWarps + warp; idx < first + count; idx += nwarps) {
        const int w = A.order[idx];
        const WindowMeta m = A.meta[w];
        SmallWin S;
        S.n1 = (int)m.n1; S.n2 = (int)m.n2; S.W = S.n2 + 1;
        S.info1 = A.s1.info + m.node1; S.info2 = A.s2.info + m.node2;
        S.poff1 = A.s1.poff + m.poff1; S.poff2 = A.s2.poff + m.poff2;
        S.pidx1 = A.s1.pidx + m.pidx1; S.pidx2 = A.s2.pidx + m.pidx2;
        S.snk1 = A.s1.sinks + m.snk1; S.snk2 = A.s2.sinks + m.snk2;
        S.nsnk1 = (int)m.nsnk1; S.nsnk2 = (int)m.nsnk2;
        const uint32_t* depth1 = A.s1.depth + m.node1;
        const uint32_t* depth2 = A.s2.depth + m.node2;
        const int n1 = S.n1, n2 = S.n2, W = S.W;
        const int4 none = make_int4(kMinInf, kMinInf, kMinInf, kMinInf);
        // ---- boundary row and column (alignment.hpp:814-894): M(0,0) = 0 seeds the sources ----
        for (int j = lane; j <= n2; j += 32) {
            const int4 b = j == 0 ? make_int4(0, kMinInf, kMinInf, kMinInf) : small_boundary_cell<P>(depth2[j], prm);  // {M, D_k}
            mA[j] = make_int4(b.x, kMinInf, kMinInf, kMinInf);
            mB[j] = make_int4(0, b.y, b.z, b.w);
        }
        for (int i = 1 + lane; i <= n1; i += 32) {
            mA[i * W] = small_boundary_cell<P>(depth1[i], prm);  // {M, I_k}
            mB[i * W] = none;
        }
        __syncwarp();
        // ---- fill, one anti-diagonal at a time ----
        for (int d = 2; d <= n1 + n2; ++d) {
            const int ilo = smax(1, d - n2), ihi = min(n1, d - 1);
            for (int i = ilo + lane; i <= ihi; i += 32) {
                const int j = d - i;
                const uint32_t ri = S.info1[i], cj = S.info2[j];
                const SmallPreds p1 = small_preds(ri, i, S.poff1, S.pidx1), p2 = small_preds(cj, j, S.poff2, S.pidx2);
                const int sub = ((ri & kInfoLabelMask) == (cj & kInfoLabelMask)) ? prm.match : -prm.mismatch;
                int diag = kMinInf, I[3] = {kMinInf, kMinInf, kMinInf}, D[3] = {kMinInf, kMinInf, kMinInf};
                for (int a = 0; a < p1.n; ++a) {
                    const int p = p1.at(a);
                    const int4 up = mA[p * W + j];
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        const int s = k == 0 ? up.y : (k == 1 ? up.z : up.w);
                        I[k] = smax(I[k], __viaddmax_s32(s, -prm.e[k], up.x - prm.oe[k]));
                    }
                    for (int b = 0; b < p2.n; ++b) diag = smax(diag, mA[p * W + p2.at(b)].x);
                }
                for (int b = 0; b < p2.n; ++b) {
                    const int q = p2.at(b);
                    const int lm = mA[i * W + q].x;
                    const int4 ld = mB[i * W + q];
#pragma unroll
                    for (int k = 0; k < P; ++k) {
                        const int s = k == 0 ? ld.y : (k == 1 ? ld.z : ld.w);
                        D[k] = smax(D[k], __viaddmax_s32(s, -prm.e[k], lm - prm.oe[k]));
                    }
                }
                int M = __viaddmax_s32(diag, sub, kMinInf);
#pragma unroll
                for (int k = 0; k < P; ++k) M = __vimax3_s32(M, I[k], D[k]);
                mA[i * W + j] = make_int4(M, I[0], I[1], I[2]);
                mB[i * W + j] = make_int4(0, D[0], D[1], D[2]);
            }
            __syncwarp();
        }
        // ---- best sink pair: first maximum in caller order, strict '>' (alignment.hpp:979-1008) ----
        long long npairs;
        if (n1 != 0 && n2 != 0) npairs = (long long)S.nsnk1 * S.nsnk2;
        else if (n1 != 0) npairs = S.nsnk1;
        else if (n2 != 0) npairs = S.nsnk2;
        else npairs = 0;
        int best = INT_MIN;
        long long besti = LLONG_MAX;
        for (long long x = lane; x < npairs; x += 32) {
            int v;
            if (n1 != 0 && n2 != 0) v = mA[(int)S.snk1[x / S.nsnk2] * W + (int)S.snk2[x % S.nsnk2]].x;
            else if (n1 != 0) v = mA[(int)S.snk1[x] * W].x;
            else v = mA[(int)S.snk2[x]].x;
            if (v > best) { best = v; besti = x; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int ob = __shfl_xor_sync(kFullMask, best, o);
            const long long oi = __shfl_xor_sync(kFullMask, besti, o);
            if (oi != LLONG_MAX && (besti == LLONG_MAX || ob > best || (ob == best && oi < besti))) { best = ob; besti = oi; }
        }
        if (lane == 0) {
            int ci = -1, cj = -1;
            if (besti != LLONG_MAX) {
                if (n1 != 0 && n2 != 0) { ci = (int)S.snk1[besti / S.nsnk2]; cj = (int)S.snk2[besti % S.nsnk2]; }
                else if (n1 != 0) { ci = (int)S.snk1[besti]; cj = 0; }
                else { ci = 0; cj = (int)S.snk2[besti]; }
            }
            A.score[w] = (ci >= 0) ? (long long)best : 0;
            // ---- traceback on the stored values (alignment.hpp:1036-1138): gap-close before diagonal, I before D, lower
            // piece first; last prev1 with its first prev2; open tested before extend per predecessor ----
            auto cM = [&](int i, int j) { return (i == 0 && j == 0) ? kMinInf : mA[i * W + j].x; };  // the corner is never a match
            auto cI = [&](int i, int j, int k) {
                if (i == 0) return kMinInf;
                const int4 v = mA[i * W + j];
                return k == 0 ? v.y : (k == 1 ? v.z : v.w);
            };
            auto cD = [&](int i, int j, int k) {
                if (j == 0) return kMinInf;
                const int4 v = mB[i * W + j];
                return k == 0 ? v.y : (k == 1 ? v.z : v.w);
            };
            int32_t* const aln = A.aln + 2 * m.out;
            const int cap = n1 + n2;
            int len = 0, comp = 0;
            while (ci >= 0 && len < cap) {
                const int M = cM(ci, cj);
                if (comp == 0) {
                    for (int k = 0; k < P; ++k) {
                        if (M == cI(ci, cj, k)) { comp = k + 1; break; }
                        if (M == cD(ci, cj, k)) { comp = -k - 1; break; }
                    }
                }
                int ni = -1, nj = -1;
                int32_t* o = aln + 2 * (int64_t)(cap - 1 - len);
                if (comp == 0) {
                    o[0] = ci - 1; o[1] = cj - 1;
                    const uint32_t ri = S.info1[ci], cf = S.info2[cj];
                    const SmallPreds p1 = small_preds(ri, ci, S.poff1, S.pidx1), p2 = small_preds(cf, cj, S.poff2, S.pidx2);
                    const int sub = ((ri & kInfoLabelMask) == (cf & kInfoLabelMask)) ? prm.match : -prm.mismatch;
                    for (int a = 0; a < p1.n; ++a) {
                        const int p = p1.at(a);
                        for (int b = 0; b < p2.n; ++b) {
                            const int q = p2.at(b);
                            if (cM(p, q) + sub == M) { ni = p; nj = q; break; }
                        }
                    }
                } else if (comp > 0) {
                    o[0] = ci - 1; o[1] = -1;
                    const int k = comp - 1;
                    const int cur = cI(ci, cj, k);
                    const SmallPreds p1 = small_preds(S.info1[ci], ci, S.poff1, S.pidx1);
                    for (int a = 0; a < p1.n; ++a) {
                        const int p = p1.at(a);
                        if (cur == cM(p, cj) - prm.oe[k]) { comp = 0; ni = p; nj = cj; break; }
                        if (cur == cI(p, cj, k) - prm.e[k]) { ni = p; nj = cj; break; }
                    }
                } else {
                    o[0] = -1; o[1] = cj - 1;
                    const int k = -comp - 1;
                    const int cur = cD(ci, cj, k);
                    const SmallPreds p2 = small_preds(S.info2[cj], cj, S.poff2, S.pidx2);
                    for (int b = 0; b < p2.n; ++b) {
                        const int q = p2.at(b);
                        if (cur == cM(ci, q) - prm.oe[k]) { comp = 0; ni = ci; nj = q; break; }
                        if (cur == cD(ci, q, k) - prm.e[k]) { ni = ci; nj = q; break; }
                    }
                }
                ++len;
                ci = ni; cj = nj;
                if (ci == 0 && cj == 0) ci = -1;  // the corner ends every path
            }
            A.aln_len[w] = (uint32_t)len;
        }
        __syncwarp();
    }
}

int popoa_small_smem_bytes() { return kSmallWarps * 2 * kSmallCells * (int)sizeof(int4); }

cudaError_t launch_popoa_small(int num_pw, const LaunchArgs& args, int first, int count, int grid, cudaStream_t stream) {
    const int smem = popoa_small_smem_bytes();
    cudaError_t err;
    switch (num_pw) {
        case 1:
            err = cudaFuncSetAttribute(popoa_small_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (err != cudaSuccess) return err;
            popoa_small_kernel<1><<<grid, kSmallWarps * 32, smem, stream>>>(args, first, count);
            break;
        case 2:
            err = cudaFuncSetAttribute(popoa_small_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (err != cudaSuccess) return err;
            popoa_small_kernel<2><<<grid, kSmallWarps * 32, smem, stream>>>(args, first, count);
            break;
        case 3:
            err = cudaFuncSetAttribute(popoa_small_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (err != cudaSuccess) return err;
            popoa_small_kernel<3><<<grid, kSmallWarps * 32, smem, stream>>>(args, first, count);
            break;
        default:
            return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace clb
