// scan_fill_snp.cu -- the un-skewed gap-fill step of scan_fill.cu on GRAPHS WITH SNP BUBBLES in both dimensions: the
// "lean strip" shapes of configs[1] (a node has the predecessor index-1, or index-2 only -- the second allele --, or both --
// the node after a bubble).  What changes against the linear prototype:
//   * rows: all lanes of a warp are on the same row, so the row's shape is a warp-uniform branch; the state of the last
//     two rows is kept (registers + a two-row ring in shared memory) and a second-allele / join row takes row i-2 / the
//     element-wise maximum as its "row above";
//   * columns: the diagonal input of a column comes from column c-1, c-2 or the maximum of both (lane-constant flags);
//     the horizontal gap chain stays a prefix scan because an SNP bubble keeps a well-defined column POSITION (both
//     alleles have the same one): H_k(c) = max over columns c' that reach c of (M'(c') + pos(c') e_k) - oe_k - (pos(c)-1) e_k,
//     and the only columns before c that do not reach it are first alleles seen from their second allele, which takes the
//     prefix from one column earlier (one extra shuffle per piece for a bubble that straddles two lanes).
// Still score only (no persisted rows / columns, no traceback, no alternate paths of other lengths); bubbles do not
// straddle a strip boundary (128 columns) or start a panel (256 rows) -- the generator below keeps them away.  Scores are
// checked against a plain CPU DP over the two graphs.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o scan_fill_snp scan_fill_snp.cu
// run:   ./scan_fill_snp [n nodes per graph, multiple of 256 (default 4096)] [windows (592)] [warps per CTA 8|12|16 (8)] [SNP rate per 1000 (50)] [columns per lane 4|8 (4)]
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
enum : unsigned char { TR = 0, TB = 1, TJ = 2 };  // predecessor index-1 / index-2 / both

struct Params {
    int match, mismatch;
    int oe[3], e[3];
};

template <int W, int C>  // C columns per lane: a strip is 32 * C columns wide
__global__ void __launch_bounds__(W * 32, 1)
snp_kernel(const unsigned char* __restrict__ seq_a, const unsigned char* __restrict__ typ_a, const unsigned char* __restrict__ seq_b,
           const unsigned char* __restrict__ typ_b, const int* __restrict__ pos_b, const int* __restrict__ row0, const int* __restrict__ col0, int n1, int n2,
           const Params prm, int* __restrict__ park_all, int* __restrict__ score_out) {
    extern __shared__ int4 smem4[];
    int4* ring = smem4;                                                         // [W][kRing] boundary column of a tile: {M, P'_1, P'_2, P'_3}
    int4* rows = ring + W * kRing;                                              // [W][2][C][32] state of the last two rows of the tile in work
    volatile int* prod = reinterpret_cast<volatile int*>(rows + W * 2 * C * 32);  // [W] rows published by warp w, counted over all its tiles
    volatile int* cons = prod + W;                                              // [W] rows of warp w's output its right neighbour has read
    unsigned char* sa = reinterpret_cast<unsigned char*>(const_cast<int*>(cons + W));  // [n1 + 1] labels of the rows
    unsigned char* ta = sa + n1 + 1;                                            // [n1 + 1] shapes of the rows
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t wa = (size_t)blockIdx.x * (n1 + 1), wb = (size_t)blockIdx.x * (n2 + 1);
    const int nstrips = n2 / (32 * C);
    const int cntw = w < nstrips ? (nstrips - w + W - 1) / W : 0;  // tiles per panel of this warp
    int* park = park_all + (size_t)blockIdx.x * nstrips * (4 * C + 1) * 32;
    for (int i = threadIdx.x; i <= n1; i += W * 32) { sa[i] = seq_a[wa + i]; ta[i] = typ_a[wa + i]; }
    if (threadIdx.x < W) { prod[threadIdx.x] = 0; cons[threadIdx.x] = 0; }
    __syncthreads();
    int4* myrows = rows + w * (2 * C * 32) + lane;  // slot r: myrows[(r * C + c) * 32]
    const int eo[3] = {prm.e[0] - prm.oe[0], prm.e[1] - prm.oe[1], prm.e[2] - prm.oe[2]};

    for (int p0 = 0; p0 < n1; p0 += kRing) {  // panel: rows p0+1 .. p0+kRing
        const int pidx = p0 / kRing, prow = min(kRing, n1 - p0);
        for (int s = w; s < nstrips; s += W) {
            const int slot_out = w, slot_in = (s + W - 1) % W;
            const int cnt_in = (nstrips - slot_in + W - 1) / W;
            const int base_in = s > 0 ? (pidx * cnt_in + (s - 1) / W) * kRing : 0;
            const int base_out = (pidx * cntw + s / W) * kRing;
            const int sprev = s >= W ? s - W : (pidx > 0 ? w + (cntw - 1) * W : -1);
            const bool wait_reader = sprev >= 0 && sprev + 1 < nstrips;
            const int j0 = (s * 32 + lane) * C;  // this lane's columns are j0+1 .. j0+C
            unsigned char bc[C];
            bool isB[C], isJ[C];
            int pe[C][3], hq[C][3], Mp[C], G[C][3], lp1, lp2 = kMinInf;  // pos(c) * e_k and (e_k - oe_k) - pos(c) * e_k
            int* pk = park + (size_t)s * (4 * C + 1) * 32 + lane;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                bc[c] = seq_b[wb + j0 + c + 1];
                const unsigned char t = typ_b[wb + j0 + c + 1];
                isB[c] = t == TB;
                isJ[c] = t == TJ;
                const int pos = pos_b[wb + j0 + c + 1];
#pragma unroll
                for (int k = 0; k < 3; ++k) { pe[c][k] = pos * prm.e[k]; hq[c][k] = eo[k] - pe[c][k]; }
                if (p0 == 0) {
                    Mp[c] = row0[wb + j0 + c + 1];
#pragma unroll
                    for (int k = 0; k < 3; ++k) G[c][k] = Mp[c] - prm.oe[k];
                } else {
                    Mp[c] = pk[(c * 4) * 32];
#pragma unroll
                    for (int k = 0; k < 3; ++k) G[c][k] = pk[(c * 4 + 1 + k) * 32];
                }
                myrows[((p0 & 1) * C + c) * 32] = make_int4(Mp[c], G[c][0], G[c][1], G[c][2]);  // state of row p0
            }
            lp1 = s == 0 ? col0[wa + p0] : (p0 == 0 ? row0[wb + s * 32 * C] : pk[4 * C * 32]);  // M(p0, column left of the strip)
            for (int r0 = 0; r0 < prow; r0 += 32) {
                const int r1 = min(r0 + 32, prow);
                if (s > 0) while (prod[slot_in] < base_in + r1) __nanosleep(100);
                if (wait_reader) while (cons[slot_out] < base_out - kRing + r1) __nanosleep(100);
                __syncwarp();
                for (int r = r0; r < r1; ++r) {
                    const int i = p0 + r + 1;
                    int4 L;  // {M(i, left), prefix maxima P'_k over everything left of the strip}
                    if (s > 0) {
                        L = ring[slot_in * kRing + r];
                    } else {
                        const int m = col0[wa + i];
                        L = make_int4(m, m, m, m);  // column 0 has position 0: M' + 0 * e_k
                    }
                    const int ai = sa[i];
                    const int rt = ta[i];  // warp-uniform
                    int Mn[C], Gn[C][3], Pk[3];
                    // one row, given the effective "row above" (EM, EG) and the effective M left of the strip in that row
                    auto body = [&](const int (&EM)[C], const int (&EG)[C][3], const int lpe) {
                        int em1 = __shfl_up_sync(kFull, EM[C - 1], 1), em2 = __shfl_up_sync(kFull, EM[C - 2], 1);
                        if (lane == 0) { em1 = lpe; em2 = kMinInf; }
                        int Mq[C], T[3] = {kMinInf, kMinInf, kMinInf}, Tw[3], tq[C][3];
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const int d1 = c == 0 ? em1 : EM[c - 1];
                            const int d2 = c == 0 ? em2 : (c == 1 ? em1 : EM[c - 2]);
                            int d = isB[c] ? d2 : d1;
                            if (isJ[c]) d = max(d, d2);
                            const int sc = ai == bc[c] ? prm.match : -prm.mismatch;
                            Mq[c] = __vimax3_s32(__viaddmax_s32(d, sc, EG[c][0]), EG[c][1], EG[c][2]);  // M': diagonal or a vertical gap
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                tq[c][k] = Mq[c] + pe[c][k];
                                if (c == C - 1) Tw[k] = T[k];
                                T[k] = max(T[k], tq[c][k]);
                            }
                        }
                        int P[3], Pp[3];
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            int t = T[k];
#pragma unroll
                            for (int dd = 1; dd < 32; dd <<= 1) t = max(t, __shfl_up_sync(kFull, t, dd));  // lanes below dd get their own value back
                            int x = __shfl_up_sync(kFull, t, 1);
                            const int carry = k == 0 ? L.y : k == 1 ? L.z : L.w;
                            x = lane == 0 ? carry : max(x, carry);           // everything before this lane
                            const int y = max(x, Tw[k]);                     // ... and this lane's first three columns
                            int yp = __shfl_up_sync(kFull, y, 1);            // everything before the previous lane's last column
                            if (lane == 0) yp = x;
                            P[k] = x;
                            Pp[k] = yp;
                        }
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            int h[3];
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                const int pin = isB[c] ? Pp[k] : P[k];  // a second allele is not reached from the first
                                h[k] = pin + hq[c][k];
                                Pp[k] = P[k];
                                P[k] = max(P[k], tq[c][k]);
                            }
                            Mn[c] = max(__vimax3_s32(Mq[c], h[0], h[1]), h[2]);
#pragma unroll
                            for (int k = 0; k < 3; ++k) Gn[c][k] = __viaddmax_s32(EG[c][k], -prm.e[k], Mn[c] - prm.oe[k]);  // vertical offer to the rows below
                        }
#pragma unroll
                        for (int k = 0; k < 3; ++k) Pk[k] = P[k];
                    };
                    if (rt == TR) {
                        body(Mp, G, lp1);
                    } else {
                        int EM[C], EG[C][3];
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const int4 v = myrows[((i & 1) * C + c) * 32];  // state of row i-2
                            if (rt == TB) {
                                EM[c] = v.x; EG[c][0] = v.y; EG[c][1] = v.z; EG[c][2] = v.w;
                            } else {
                                EM[c] = max(Mp[c], v.x); EG[c][0] = max(G[c][0], v.y); EG[c][1] = max(G[c][1], v.z); EG[c][2] = max(G[c][2], v.w);
                            }
                        }
                        body(EM, EG, rt == TB ? lp2 : max(lp1, lp2));
                    }
                    const bool needed_later = r + 2 < prow && ta[i + 2] != TR;  // warp-uniform: row i+2 takes this row as a "row above"
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        Mp[c] = Mn[c];
#pragma unroll
                        for (int k = 0; k < 3; ++k) G[c][k] = Gn[c][k];
                        if (needed_later) myrows[((i & 1) * C + c) * 32] = make_int4(Mn[c], Gn[c][0], Gn[c][1], Gn[c][2]);
                    }
                    lp2 = lp1;
                    lp1 = L.x;
                    if (lane == 31 && s + 1 < nstrips) ring[slot_out * kRing + r] = make_int4(Mn[C - 1], Pk[0], Pk[1], Pk[2]);
                }
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    if (s + 1 < nstrips) prod[slot_out] = base_out + r1;
                    if (s > 0) cons[slot_in] = base_in + r1;
                }
            }
            if (p0 + kRing < n1) {  // park the state for this strip's tile of the next panel (its first row is a plain row)
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    pk[(c * 4) * 32] = Mp[c];
#pragma unroll
                    for (int k = 0; k < 3; ++k) pk[(c * 4 + 1 + k) * 32] = G[c][k];
                }
                pk[4 * C * 32] = lp1;
            } else if (s == nstrips - 1 && lane == 31) {
                score_out[blockIdx.x] = Mp[C - 1];
            }
        }
    }
}

// ------------------------------------------------------------------ host ------------------------------------------------------------------
struct Graph {
    std::vector<unsigned char> lab, typ;  // index 0 = the boundary
    std::vector<int> pos;
};

int gap_of(const Params& p, int len) {  // score of one gap over `len` nodes
    if (len <= 0) return 0;
    return -std::min(std::min(p.oe[0] + p.e[0] * (len - 1), p.oe[1] + p.e[1] * (len - 1)), p.oe[2] + p.e[2] * (len - 1));
}

int cpu_score(const Graph& a, const Graph& b, int n1, int n2, const Params& p) {
    // plain DP over the two graphs: every node takes the maximum over its predecessors (1 or 2 per dimension)
    std::vector<std::vector<int>> M(n1 + 1, std::vector<int>(n2 + 1)), V[3], H[3];
    for (int k = 0; k < 3; ++k) { V[k].assign(n1 + 1, std::vector<int>(n2 + 1, kMinInf)); H[k] = V[k]; }
    auto preds = [](const Graph& g, int i, int* out) {
        int n = 0;
        if (g.typ[i] != TB) out[n++] = i - 1;
        if (g.typ[i] != TR) out[n++] = i - 2;
        return n;
    };
    for (int j = 0; j <= n2; ++j) M[0][j] = gap_of(p, b.pos[j]);
    for (int i = 1; i <= n1; ++i) {
        M[i][0] = gap_of(p, a.pos[i]);
        int pi[2], pj[2];
        const int ni = preds(a, i, pi);
        for (int j = 1; j <= n2; ++j) {
            const int nj = preds(b, j, pj);
            int best = kMinInf;
            for (int x = 0; x < ni; ++x)
                for (int y = 0; y < nj; ++y) best = std::max(best, M[pi[x]][pj[y]]);
            best += a.lab[i] == b.lab[j] ? p.match : -p.mismatch;
            for (int k = 0; k < 3; ++k) {
                int v = kMinInf, h = kMinInf;
                for (int x = 0; x < ni; ++x) v = std::max(v, std::max(V[k][pi[x]][j] - p.e[k], M[pi[x]][j] - p.oe[k]));
                for (int y = 0; y < nj; ++y) h = std::max(h, std::max(H[k][i][pj[y]] - p.e[k], M[i][pj[y]] - p.oe[k]));
                V[k][i][j] = v;
                H[k][i][j] = h;
                best = std::max(best, std::max(v, h));
            }
            M[i][j] = best;
        }
    }
    return M[n1][n2];
}

template <int W, int C>
float run(const unsigned char* da, const unsigned char* dta, const unsigned char* db, const unsigned char* dtb, const int* dpos, const int* drow0,
          const int* dcol0, int n, int windows, const Params& prm, int* dpark, int* dscore, int reps) {
    const size_t smem = (size_t)W * kRing * 16 + (size_t)W * 2 * C * 32 * 16 + 2 * W * 4 + 2 * (n + 1) + 16;
    cudaFuncSetAttribute(snp_kernel<W, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    snp_kernel<W, C><<<windows, W * 32, smem>>>(da, dta, db, dtb, dpos, drow0, dcol0, n, n, prm, dpark, dscore);  // warm-up
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) snp_kernel<W, C><<<windows, W * 32, smem>>>(da, dta, db, dtb, dpos, drow0, dcol0, n, n, prm, dpark, dscore);
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
    const int W = argc > 3 ? atoi(argv[3]) : 8;
    const unsigned snp = argc > 4 ? (unsigned)atoi(argv[4]) : 50;
    const int C = argc > 5 ? atoi(argv[5]) : 4;
    if (n % 256 || n < 256) { printf("n must be a multiple of 256\n"); return 1; }
    const Params prm = {20, 80, {60 + 30, 800 + 5, 2500 + 1}, {30, 5, 1}};
    unsigned long long x = 88172645463325252ull;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    // a graph of n nodes from a base sequence: plain nodes and SNP bubbles (first allele, second allele, join); bubbles keep
    // away from the first column of a strip / first row of a panel (`block`)
    auto make_graph = [&](const std::vector<unsigned char>& base, int block) {
        Graph g;
        g.lab.assign(n + 1, 0); g.typ.assign(n + 1, TR); g.pos.assign(n + 1, 0);
        size_t q = 0;
        auto next_base = [&]() { return base[q++ % base.size()]; };
        auto blocked = [&](int idx) { return (idx - 1) % block == 0; };
        for (int i = 1; i <= n;) {
            if (i >= 2 && i + 2 <= n && rnd() % 1000 < snp && !blocked(i + 1) && !blocked(i + 2)) {
                g.lab[i] = next_base(); g.typ[i] = TR;
                g.lab[i + 1] = (unsigned char)((g.lab[i] + 1 + rnd() % 3) & 3); g.typ[i + 1] = TB;
                g.lab[i + 2] = next_base(); g.typ[i + 2] = TJ;
                i += 3;
            } else {
                g.lab[i] = next_base(); g.typ[i] = TR;
                i += 1;
            }
        }
        for (int i = 1; i <= n; ++i) g.pos[i] = g.typ[i] == TB ? g.pos[i - 1] : g.pos[i - 1] + 1;
        return g;
    };
    const size_t stride = (size_t)n + 1;
    std::vector<unsigned char> la(windows * stride), ta(windows * stride), lb(windows * stride), tb(windows * stride);
    std::vector<int> posb(windows * stride), row0(windows * stride), col0(windows * stride);
    std::vector<Graph> keepa, keepb;
    for (int wdw = 0; wdw < windows; ++wdw) {
        std::vector<unsigned char> base(n), other;
        for (auto& c : base) c = (unsigned char)(rnd() & 3);
        for (int j = 0; j < n; ++j) {  // the second sequence: 3 % substitutions, 1 % indels
            const unsigned r = (unsigned)(rnd() % 1000);
            if (r < 5) continue;
            other.push_back(r < 35 ? (unsigned char)((base[j] + 1 + rnd() % 3) & 3) : base[j]);
            if (r >= 35 && r < 40) other.push_back((unsigned char)(rnd() & 3));
        }
        Graph ga = make_graph(base, 256), gb = make_graph(other, 32 * C);
        for (int i = 0; i <= n; ++i) {
            la[wdw * stride + i] = ga.lab[i]; ta[wdw * stride + i] = ga.typ[i];
            lb[wdw * stride + i] = gb.lab[i]; tb[wdw * stride + i] = gb.typ[i];
            posb[wdw * stride + i] = gb.pos[i];
            row0[wdw * stride + i] = gap_of(prm, gb.pos[i]);
            col0[wdw * stride + i] = gap_of(prm, ga.pos[i]);
        }
        if (wdw % std::max(1, windows / 4) == 0) { keepa.push_back(ga); keepb.push_back(gb); }
    }
    unsigned char *da, *dta, *db, *dtb;
    int *dpos, *drow0, *dcol0, *dscore, *dpark;
    cudaMalloc(&da, la.size()); cudaMalloc(&dta, ta.size()); cudaMalloc(&db, lb.size()); cudaMalloc(&dtb, tb.size());
    cudaMalloc(&dpos, posb.size() * 4); cudaMalloc(&drow0, row0.size() * 4); cudaMalloc(&dcol0, col0.size() * 4);
    cudaMalloc(&dscore, windows * sizeof(int));
    cudaMalloc(&dpark, (size_t)windows * (n / (32 * C)) * (4 * C + 1) * 32 * sizeof(int));
    cudaMemcpy(da, la.data(), la.size(), cudaMemcpyHostToDevice); cudaMemcpy(dta, ta.data(), ta.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, lb.data(), lb.size(), cudaMemcpyHostToDevice); cudaMemcpy(dtb, tb.data(), tb.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dpos, posb.data(), posb.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(drow0, row0.data(), row0.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dcol0, col0.data(), col0.size() * 4, cudaMemcpyHostToDevice);
    const int reps = 3;
#define RUN(WW, CC) run<WW, CC>(da, dta, db, dtb, dpos, drow0, dcol0, n, windows, prm, dpark, dscore, reps)
    const float ms = C == 8 ? (W == 12 ? RUN(12, 8) : W == 16 ? RUN(16, 8) : RUN(8, 8)) : (W == 12 ? RUN(12, 4) : W == 16 ? RUN(16, 4) : RUN(8, 4));
#undef RUN
    std::vector<int> score(windows);
    cudaMemcpy(score.data(), dscore, windows * sizeof(int), cudaMemcpyDeviceToHost);
    int checked = 0, bad = 0;
    for (size_t q = 0; q < keepa.size() && (size_t)checked * n * n < (size_t)64 << 20; ++q, ++checked) {
        const int wdw = (int)q * std::max(1, windows / 4);
        const int ref = cpu_score(keepa[q], keepb[q], n, n, prm);
        if (ref != score[wdw]) { ++bad; printf("  window %d: GPU %d, CPU %d\n", wdw, score[wdw], ref); }
    }
    size_t nb = 0;
    for (unsigned char t : tb) nb += t == TB;
    const double cells = (double)windows * (n + 1.0) * (n + 1.0);
    printf("scan_fill_snp: %d windows of %d x %d nodes, %.1f %% of the nodes in SNP bubbles, %d warps per CTA, %d columns per lane: %.3f ms, %.1f GCUPS; %d windows "
           "checked against the CPU DP, %d differ\n", windows, n, n, 300.0 * nb / tb.size(), W, C, ms, cells / (ms * 1e-3) * 1e-9, checked, bad);
    return bad ? 2 : 0;
}
