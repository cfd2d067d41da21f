// pipes.cu -- issue-rate probes for the INT32 instruction classes the gap-fill step is made of (sm_100a).
// Each kernel runs NCH independent dependency chains per thread; result = warp-instructions per clock per SM.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu ; run: ./pipes
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define NCH 8
#define UNROLL 8

enum Kind { K_VIADDMNMX_RRR, K_VIADDMNMX_RCR, K_VIMNMX3, K_VIMNMX2, K_IADD, K_IMAD, K_SEL, K_ISETP_SEL, K_MIX_DPX_IMAD, K_MIX_DPX_IADD,
            K_MIX_DPX_MNMX2, K_SHFL, K_LDS128, K_MIX_DPX_IMAD2, K_CELL, K_PREDOFF, K_LOP3, K_MIX_MNMX2_IMAD, K_MIX_DPXC_IMADC, K_MIX_DPX_IADDC, K_MIX_DPXC_IADDC, K_MIX_2DPXC_IADDC, K_MIX_MNMX2_IADDC, K_NKINDS };
static const char* kNames[] = {"VIADDMNMX r,r,r", "VIADDMNMX r,c,r", "VIMNMX3", "VIMNMX (2-in)", "IADD3", "IMAD (x*y+z)", "SEL", "ISETP+SEL",
                               "VIADDMNMX + IMAD 1:1", "VIADDMNMX + IADD3 1:1", "VIADDMNMX + VIMNMX2 1:1", "SHFL.UP", "LDS.128",
                               "VIADDMNMX + 2 IMAD", "bare cell (11 ALU + 8 IMAD)", "predicated-off VIMNMX", "LOP3", "VIMNMX2 + IMAD 1:1", "VIADDMNMX r,c,r + IMAD r,c,c", "VIADDMNMX r,r,r + IADD r,c", "VIADDMNMX r,c,r + IADD r,c", "2 VIADDMNMX r,c,r + IADD r,c", "VIMNMX2 + IADD r,c"};
// instructions counted per chain-update
static const int kInstr[] = {1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 1, 1, 3, 19, 1, 1, 2, 2, 2, 2, 3, 2};

__constant__ int cst[8];

template <int K>
__global__ void __launch_bounds__(1024) probe(int* out, int iters, int seed, int one) {
    __shared__ int4 sm[1024];
    int a[NCH], b[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) { a[k] = seed + k * 7 + threadIdx.x; b[k] = seed * 3 + k + threadIdx.x; }
    sm[threadIdx.x] = make_int4(a[0], a[1], a[2], a[3]);
    __syncthreads();
    const int d = seed | 1;
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int o = b[(k + 1) % NCH];
                if (K == K_VIADDMNMX_RRR) a[k] = __viaddmax_s32(a[k], d, o);
                if (K == K_VIADDMNMX_RCR) a[k] = __viaddmax_s32(a[k], cst[1], o);
                if (K == K_VIMNMX3) a[k] = __vimax3_s32(a[k], d, o);
                if (K == K_VIMNMX2) { if (u & 1) asm volatile("max.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(o)); else asm volatile("min.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(o)); }
                if (K == K_IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(o));
                if (K == K_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(one), "r"(o));
                if (K == K_SEL) asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.s32 %0, %0, %1, p;}" : "+r"(a[k]) : "r"(o), "r"(one));
                if (K == K_ISETP_SEL) asm volatile("{.reg .pred p; setp.gt.s32 p, %0, %1; selp.s32 %0, %1, %2, p;}" : "+r"(a[k]) : "r"(o), "r"(d));
                if (K == K_MIX_DPX_IMAD) { a[k] = __viaddmax_s32(a[k], d, o); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(b[k]) : "r"(one), "r"(d)); }
                if (K == K_MIX_DPX_IADD) { a[k] = __viaddmax_s32(a[k], d, o); asm volatile("add.s32 %0, %0, %1;" : "+r"(b[k]) : "r"(d)); }
                if (K == K_MIX_DPX_MNMX2) { a[k] = __viaddmax_s32(a[k], d, o); asm volatile("max.s32 %0, %0, %1;" : "+r"(b[k]) : "r"(d + k + it)); }
                if (K == K_SHFL) a[k] = __shfl_up_sync(0xffffffffu, a[k], 1);
                if (K == K_LDS128) { const int4 v = sm[(a[k] + threadIdx.x) & 1023]; a[k] = v.x + v.w; }
                if (K == K_MIX_DPX_IMAD2) {
                    a[k] = __viaddmax_s32(a[k], d, o);
                    asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(b[k]) : "r"(one), "r"(d));
                    asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(b[k]) : "r"(one), "r"(d));
                }
                if (K == K_PREDOFF) asm volatile("{.reg .pred p; setp.eq.s32 p, %2, 12345; @p max.s32 %0, %0, %1;}" : "+r"(a[k]) : "r"(o), "r"(one));
                if (K == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(o), "r"(d));
                if (K == K_MIX_DPXC_IMADC) { a[k] = __viaddmax_s32(a[k], cst[1], o); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(b[k]) : "r"(cst[7]), "r"(cst[2])); }
                if (K == K_MIX_DPX_IADDC) { a[k] = __viaddmax_s32(a[k], d, o); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(b[k]) : "r"(cst[7]), "r"(cst[2])); }
                if (K == K_MIX_DPXC_IADDC) { a[k] = __viaddmax_s32(a[k], cst[1], o); asm volatile("mad.lo.s32 %0, %0, 1, %1;" : "+r"(b[k]) : "r"(cst[2])); }
                if (K == K_MIX_2DPXC_IADDC) { a[k] = __viaddmax_s32(a[k], cst[1], o); b[k] = __viaddmax_s32(b[k], cst[3], a[(k+2)%NCH]); asm volatile("mad.lo.s32 %0, %0, 1, %1;" : "+r"(b[(k+3)%NCH]) : "r"(cst[2])); }
                if (K == K_MIX_MNMX2_IADDC) { asm volatile("max.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(o)); asm volatile("mad.lo.s32 %0, %0, 1, %1;" : "+r"(b[k]) : "r"(cst[2])); }
                if (K == K_MIX_MNMX2_IMAD) { asm volatile("max.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(o)); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(b[k]) : "r"(one), "r"(d)); }
            }
            if (K == K_CELL) {
                // one DP cell per chain pair: state M=a[0], I=a[1..3], D=a[4..6], E=a[7]; constants from the constant bank
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    int& M = a[0]; int& E = a[7];
                    const bool p = (b[c] & 3) == (lane & 3);
                    int x = p ? E + cst[0] : E - cst[1];
                    int t0 = M - cst[2], t1 = M - cst[3], t2 = M - cst[4];
                    a[1] = __viaddmax_s32(a[1], -cst[5], t0);
                    a[2] = __viaddmax_s32(a[2], -cst[6], t1);
                    a[3] = __viaddmax_s32(a[3], -cst[7], t2);
                    x = __vimax3_s32(x, a[1], a[2]);
                    x = max(x, a[3]);
                    int u0 = b[2] - cst[2], u1 = b[2] - cst[3], u2 = b[2] - cst[4];
                    a[4] = __viaddmax_s32(a[4], -cst[5], u0);
                    a[5] = __viaddmax_s32(a[5], -cst[6], u1);
                    a[6] = __viaddmax_s32(a[6], -cst[7], u2);
                    x = __vimax3_s32(x, a[4], a[5]);
                    E = M;
                    M = max(x, a[6]);
                    b[2] = M; b[c] += x;
                }
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < NCH; ++k) s ^= a[k] ^ b[k];
    if (s == 0x7fffffff) out[0] = s;
}

template <int K>
double run(int warps_per_sm, int sms, double clk_ghz) {
    int* d;
    cudaMalloc(&d, 4);
    const int iters = 2048;
    const int threads = 32 * (warps_per_sm > 32 ? 32 : warps_per_sm);
    const int grid = sms * (warps_per_sm > 32 ? warps_per_sm / 32 : 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        probe<K><<<grid, threads>>>(d, iters, 12345, 1);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    cudaFree(d);
    const double per_update = (K == K_CELL) ? 2.0 * kInstr[K] / NCH : kInstr[K];  // K_CELL: 2 cells per u-iteration, not per chain
    const double winstr = (double)grid * (threads / 32) * iters * UNROLL * NCH * per_update;
    return winstr / (best * 1e-3 * clk_ghz * 1e9) / sms;  // warp-instr per clock per SM
}

int main() {
    int h[8] = {20, 40, 61, 361, 2051, 30, 5, 1};
    cudaMemcpyToSymbol(cst, h, sizeof(h));
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz * 1e-6;
    printf("device %s, %d SMs, clock %.3f GHz (results assume this clock)\n", p.name, p.multiProcessorCount, ghz);
    const int wps[] = {4, 8, 12, 16, 32};
    printf("%-30s", "warp-instr/clk/SM at warps/SM:");
    for (int w : wps) printf(" %7d", w);
    printf("\n");
#define ROW(K) { printf("%-30s", kNames[K]); for (int w : wps) printf(" %7.3f", run<K>(w, p.multiProcessorCount, ghz)); printf("\n"); }
    ROW(K_VIADDMNMX_RRR) ROW(K_VIADDMNMX_RCR) ROW(K_VIMNMX3) ROW(K_VIMNMX2) ROW(K_IADD) ROW(K_IMAD) ROW(K_SEL) ROW(K_ISETP_SEL) ROW(K_LOP3)
    ROW(K_MIX_DPX_IMAD) ROW(K_MIX_DPX_IADD) ROW(K_MIX_DPX_MNMX2) ROW(K_MIX_MNMX2_IMAD) ROW(K_MIX_DPX_IMAD2) ROW(K_MIX_DPXC_IMADC) ROW(K_MIX_DPX_IADDC) ROW(K_MIX_DPXC_IADDC) ROW(K_MIX_2DPXC_IADDC) ROW(K_MIX_MNMX2_IADDC) ROW(K_SHFL) ROW(K_LDS128) ROW(K_PREDOFF) ROW(K_CELL)
    return 0;
}
