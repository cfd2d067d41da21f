// scan_fill.cu -- prototype of ONE structural alternative to the skewed gap-fill step (DESIGN.md
// section 4 "what would come next"): rows are NOT skewed over the lanes.  A warp owns a strip of 128 columns (4 per lane)
// and computes one matrix row per step; the horizontal gap chain of every piece, which forces the skew in
// popoa_kernels.cu, is resolved inside the row by a max-plus prefix scan over the lanes:
//     H_k(j) = max over j' < j of ( M'(j') - oe_k - (j-1-j') e_k ),   M' = the cell's maximum without its horizontal gaps
// (opening a gap right after closing one never wins, oe_k >= e_k), i.e. per lane an aggregate of its four columns,
// 5 shuffle+max rounds per piece over (aggregate + lane * 4 e_k), and a three-step chain inside the lane.
// This file measures the bubble-free CEILING of that step: two LINEAR sequences per window (no graph features, no
// persisted rows, no traceback), strips of a window pipelined over the warps of a CTA through a shared-memory ring, three
// gap pieces with the production parameters.  Scores are checked against a plain CPU DP on small windows.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o scan_fill scan_fill.cu
// run:   ./scan_fill [n (multiple of 128, default 4096)] [windows (default 592)] [warps per CTA: 8, 12, 16 or 24 (default 12)]
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace {

constexpr int kMinInf = INT32_MIN / 2;
constexpr int kRing = 256;  // rows of a panel = rows of boundary column kept per strip slot
constexpr unsigned kFull = 0xffffffffu;

struct Params {
    int match, mismatch;
    int oe[3], e[3];
};

__device__ __forceinline__ int gap_row0(const Params& p, int j) {  // M(0, j) = M(j, 0): one gap of length j
    if (j == 0) return 0;
    // (written with the explicit three-input minimum: nvcc 12.9 turned max(max(-a, -b), -c) into min(max(a, b), c) here)
    return -__vimin3_s32(p.oe[0] + p.e[0] * (j - 1), p.oe[1] + p.e[1] * (j - 1), p.oe[2] + p.e[2] * (j - 1));
}

// One CTA per window.  The matrix is cut into tiles of kRing rows x 128 columns; warp w takes the strips w, w + W, ... of a
// panel of kRing rows, then the next panel (so a strip never runs more than one panel ahead of its right neighbour and the
// ring of a strip slot is exactly one panel deep).  Between two panels a strip parks its 16 state words per lane in `park`.
template <int W>
__global__ void __launch_bounds__(W * 32, 1) scan_fill_kernel(const unsigned char* __restrict__ seq_a, const unsigned char* __restrict__ seq_b,
                                                              int n1, int n2, const Params prm, int* __restrict__ park_all,
                                                              int* __restrict__ score_out) {
    extern __shared__ unsigned char smem_raw[];
    int4* ring = reinterpret_cast<int4*>(smem_raw);                          // [W][kRing] boundary column of a tile: {M, C_1, C_2, C_3}
    volatile int* prod = reinterpret_cast<volatile int*>(ring + W * kRing);  // [W] rows published by warp w, counted over all its tiles
    volatile int* cons = prod + W;                                           // [W] rows of warp w's output its right neighbour has read
    unsigned char* sa = reinterpret_cast<unsigned char*>(const_cast<int*>(cons + W));  // [n1 + 1] sequence of the rows
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned char* a = seq_a + (size_t)blockIdx.x * n1;
    const unsigned char* b = seq_b + (size_t)blockIdx.x * n2;
    const int nstrips = n2 / 128;
    const int cntw = w < nstrips ? (nstrips - w + W - 1) / W : 0;  // tiles per panel of this warp
    int* park = park_all + (size_t)blockIdx.x * nstrips * 17 * 32;
    for (int i = threadIdx.x; i < n1; i += W * 32) sa[i + 1] = a[i];
    if (threadIdx.x < W) { prod[threadIdx.x] = 0; cons[threadIdx.x] = 0; }
    __syncthreads();
    int up[3], down[3];  // lane * 4 e_k and -(lane - 1) * 4 e_k
#pragma unroll
    for (int k = 0; k < 3; ++k) { up[k] = lane * 4 * prm.e[k]; down[k] = -(lane - 1) * 4 * prm.e[k]; }

#ifdef DEBUG_ROW
    if (blockIdx.x == 0 && threadIdx.x == 0)
        printf("prm %d %d oe %d %d %d e %d %d %d row0(1) %d row0(5) %d n1 %d n2 %d\n", prm.match, prm.mismatch, prm.oe[0], prm.oe[1], prm.oe[2], prm.e[0], prm.e[1],
               prm.e[2], gap_row0(prm, 1), gap_row0(prm, 5), n1, n2);
#endif
    for (int p0 = 0; p0 < n1; p0 += kRing) {  // panel: rows p0+1 .. p0+kRing
        const int pidx = p0 / kRing, prow = min(kRing, n1 - p0);
        for (int s = w; s < nstrips; s += W) {
            const int slot_out = w, slot_in = (s + W - 1) % W;  // a warp always publishes into its own slot
            const int cnt_in = (nstrips - slot_in + W - 1) / W;
            const int base_in = s > 0 ? (pidx * cnt_in + (s - 1) / W) * kRing : 0;  // rows the left neighbour published before the tile read here
            const int base_out = (pidx * cntw + s / W) * kRing;
            // the tile that used this slot last must have been read: it is this warp's previous tile, whose reader may lag
            const int sprev = s >= W ? s - W : (pidx > 0 ? w + (cntw - 1) * W : -1);
            const bool wait_reader = sprev >= 0 && sprev + 1 < nstrips;
            const int j0 = s * 128 + lane * 4;  // this lane's columns are j0+1 .. j0+4
            unsigned char bc[4];
            int Mp[4], G[4][3], left_prev;
            int* pk = park + (size_t)s * 17 * 32 + lane;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                bc[c] = b[j0 + c];
                if (p0 == 0) {
                    Mp[c] = gap_row0(prm, j0 + c + 1);
#pragma unroll
                    for (int k = 0; k < 3; ++k) G[c][k] = Mp[c] - prm.oe[k];
                } else {
                    Mp[c] = pk[(c * 4) * 32];
#pragma unroll
                    for (int k = 0; k < 3; ++k) G[c][k] = pk[(c * 4 + 1 + k) * 32];
                }
            }
            left_prev = s == 0 ? gap_row0(prm, p0) : (p0 == 0 ? gap_row0(prm, s * 128) : pk[16 * 32]);  // M(p0, column left of the strip)
            for (int r0 = 0; r0 < prow; r0 += 32) {
                const int r1 = min(r0 + 32, prow);
                // all lanes poll the same word (one broadcast load): a single polling lane leaves the warp split in two for
                // the whole row loop (measured: 15.7 active threads per instruction, half of all instructions in the poll)
                if (s > 0) while (prod[slot_in] < base_in + r1) __nanosleep(100);
                if (wait_reader) while (cons[slot_out] < base_out - kRing + r1) __nanosleep(100);
                __syncwarp();
                for (int r = r0; r < r1; ++r) {
                    const int i = p0 + r + 1;
                    int4 L;  // {M(i, left), offers of the left column to this strip's first column}
                    if (s > 0) {
                        L = ring[slot_in * kRing + r];
                    } else {
                        const int m = gap_row0(prm, i);
                        L = make_int4(m, m - prm.oe[0], m - prm.oe[1], m - prm.oe[2]);
                    }
                    const int ai = sa[i];
                    // diagonal inputs: the previous row's M one column to the left
                    int dg = __shfl_up_sync(kFull, Mp[3], 1);
                    if (lane == 0) dg = left_prev;
                    left_prev = L.x;
                    int Mq[4], o[4][3];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int d = c == 0 ? dg : Mp[c - 1];
                        const int sc = ai == bc[c] ? prm.match : -prm.mismatch;
                        Mq[c] = __vimax3_s32(__viaddmax_s32(d, sc, G[c][0]), G[c][1], G[c][2]);  // M': diagonal or a vertical gap
#pragma unroll
                        for (int k = 0; k < 3; ++k) o[c][k] = Mq[c] - prm.oe[k];               // what the cell offers the next column
                    }
                    int hh[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        // what this lane's columns offer the column right of the lane, then a prefix maximum with 4 e_k decay per lane
                        int agg = __viaddmax_s32(o[2][k], -prm.e[k], o[3][k]);
                        agg = __viaddmax_s32(o[1][k], -2 * prm.e[k], agg);
                        agg = __viaddmax_s32(o[0][k], -3 * prm.e[k], agg);
                        int t = agg + up[k];
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) t = max(t, __shfl_up_sync(kFull, t, d));  // lanes below d get their own value back
                        int x = __shfl_up_sync(kFull, t, 1);
                        if (lane == 0) x = kMinInf;
                        const int from_left = (k == 0 ? L.y : k == 1 ? L.z : L.w) - up[k];  // the left strip's offer, lane * 4 columns further
                        hh[k] = __viaddmax_s32(x, down[k], from_left);
                    }
#ifdef DEBUG_ROW
                    if (blockIdx.x == 0 && i <= 2 && s == 0 && (lane < 3 || lane == 31))
                        printf("row %d lane %d ai %d bc %d%d%d%d dg %d Mq %d %d %d %d Hin %d %d %d Lyzw %d %d %d G0 %d %d %d\n", i, lane, ai, bc[0], bc[1], bc[2], bc[3], dg, Mq[0], Mq[1], Mq[2],
                               Mq[3], hh[0], hh[1], hh[2], L.y, L.z, L.w, G[0][0], G[0][1], G[0][2]);
#endif
                    int Mn[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        Mn[c] = max(__vimax3_s32(Mq[c], hh[0], hh[1]), hh[2]);
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            hh[k] = __viaddmax_s32(hh[k], -prm.e[k], o[c][k]);                  // horizontal gap into the next column
                            G[c][k] = __viaddmax_s32(G[c][k], -prm.e[k], Mn[c] - prm.oe[k]);    // vertical offer to the row below
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) Mp[c] = Mn[c];
                    if (lane == 31 && s + 1 < nstrips) ring[slot_out * kRing + r] = make_int4(Mn[3], hh[0], hh[1], hh[2]);
                }
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    if (s + 1 < nstrips) prod[slot_out] = base_out + r1;
                    if (s > 0) cons[slot_in] = base_in + r1;
                }
            }
            if (p0 + kRing < n1) {  // park the state for this strip's tile of the next panel
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    pk[(c * 4) * 32] = Mp[c];
#pragma unroll
                    for (int k = 0; k < 3; ++k) pk[(c * 4 + 1 + k) * 32] = G[c][k];
                }
                pk[16 * 32] = left_prev;
            } else if (s == nstrips - 1 && lane == 31) {
                score_out[blockIdx.x] = Mp[3];
            }
        }
    }
}

int cpu_score(const unsigned char* a, const unsigned char* b, int n1, int n2, const Params& p) {
    std::vector<int> M(n2 + 1), V((size_t)3 * (n2 + 1), kMinInf), Mprev(n2 + 1);
    auto row0 = [&](int j) { return j == 0 ? 0 : std::max(std::max(-(p.oe[0] + p.e[0] * (j - 1)), -(p.oe[1] + p.e[1] * (j - 1))), -(p.oe[2] + p.e[2] * (j - 1))); };
    for (int j = 0; j <= n2; ++j) Mprev[j] = row0(j);
    for (int i = 1; i <= n1; ++i) {
        int H[3] = {kMinInf, kMinInf, kMinInf};
        M[0] = row0(i);
        for (int j = 1; j <= n2; ++j) {
            int best = Mprev[j - 1] + (a[i - 1] == b[j - 1] ? p.match : -p.mismatch);
            for (int k = 0; k < 3; ++k) {
                int& v = V[(size_t)k * (n2 + 1) + j];
                v = std::max(v - p.e[k], Mprev[j] - p.oe[k]);
                H[k] = std::max(H[k] - p.e[k], M[j - 1] - p.oe[k]);
                best = std::max(best, std::max(v, H[k]));
            }
            M[j] = best;
        }
        std::swap(M, Mprev);
    }
    return Mprev[n2];
}

template <int W>
float run(const unsigned char* da, const unsigned char* db, int n, int windows, const Params& prm, int* dpark, int* dscore, int reps) {
    const size_t smem = (size_t)W * kRing * 16 + 2 * W * 4 + n + 16;
    cudaFuncSetAttribute(scan_fill_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    scan_fill_kernel<W><<<windows, W * 32, smem>>>(da, db, n, n, prm, dpark, dscore);  // warm-up
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) scan_fill_kernel<W><<<windows, W * 32, smem>>>(da, db, n, n, prm, dpark, dscore);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (cudaGetLastError() != cudaSuccess) { printf("CUDA error\n"); exit(1); }
    return ms / reps;
}

}  // namespace

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 4096;
    const int windows = argc > 2 ? atoi(argv[2]) : 592;
    const int W = argc > 3 ? atoi(argv[3]) : 12;
    if (n % 128 || n < 128) { printf("n must be a multiple of 128\n"); return 1; }
    const Params prm = {20, 80, {60 + 30, 800 + 5, 2500 + 1}, {30, 5, 1}};
    std::vector<unsigned char> a((size_t)windows * n), b((size_t)windows * n);
    unsigned long long x = 88172645463325252ull;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    for (int wdw = 0; wdw < windows; ++wdw) {  // b = a with 5 % substitutions and a few indels
        size_t pa = 0;
        for (int j = 0; j < n; ++j) a[(size_t)wdw * n + j] = (unsigned char)(rnd() & 3);
        for (int j = 0; j < n; ++j) {
            const unsigned r = (unsigned)(rnd() % 1000);
            if (r < 5 && pa + 1 < (size_t)n) ++pa;  // deletion
            unsigned char c = a[(size_t)wdw * n + std::min<size_t>(pa, n - 1)];
            if (r >= 5 && r < 55) c = (unsigned char)((c + 1 + rnd() % 3) & 3);
            b[(size_t)wdw * n + j] = c;
            if (!(r >= 55 && r < 60)) ++pa;  // insertion keeps pa
        }
    }
    unsigned char *da, *db;
    int* dscore;
    cudaMalloc(&da, a.size());
    cudaMalloc(&db, b.size());
    cudaMalloc(&dscore, windows * sizeof(int));
    int* dpark;
    cudaMalloc(&dpark, (size_t)windows * (n / 128) * 17 * 32 * sizeof(int));
    cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice);
    const int reps = 3;
    float ms = W == 8 ? run<8>(da, db, n, windows, prm, dpark, dscore, reps) : W == 16 ? run<16>(da, db, n, windows, prm, dpark, dscore, reps)
             : W == 24 ? run<24>(da, db, n, windows, prm, dpark, dscore, reps) : run<12>(da, db, n, windows, prm, dpark, dscore, reps);
    std::vector<int> score(windows);
    cudaMemcpy(score.data(), dscore, windows * sizeof(int), cudaMemcpyDeviceToHost);
    int checked = 0, bad = 0;
    for (int wdw = 0; wdw < windows && (size_t)checked * n * n < (size_t)400 << 20; wdw += std::max(1, windows / 4), ++checked)
    {
        const int ref = cpu_score(&a[(size_t)wdw * n], &b[(size_t)wdw * n], n, n, prm);
        if (ref != score[wdw]) { ++bad; printf("  window %d: GPU %d, CPU %d\n", wdw, score[wdw], ref); }
    }
    const double cells = (double)windows * (n + 1.0) * (n + 1.0);
    printf("scan_fill: %d windows of %d x %d, %d warps per CTA: %.3f ms, %.1f GCUPS; %d windows checked against the CPU DP, %d differ\n", windows, n, n, W,
           ms, cells / (ms * 1e-3) * 1e-9, checked, bad);
    return bad ? 2 : 0;
}
