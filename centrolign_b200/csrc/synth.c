/*
 * synth.c -- deterministic synthetic PO-to-PO gap-fill windows (bench / test input).
 *
 * Implements the "batched inter-anchor windows" workload of BASELINE.json configs[1] as
 * specified in SURVEY.md section 8(d): both graphs of a window descend from one random
 * base sequence (3 % substitutions, one 7-bp indel per kbp each), and each is turned into
 * a partial-order graph with SNP bubbles at 5 % of positions and one 171-node alternate
 * path per 2 kbp.  Node ids are NOT topologically ordered (bubble nodes are appended after
 * the backbone) and predecessor lists are shuffled, the way the reference's callers hand
 * graphs to po_poa (include/centrolign/alignment.hpp:78-85).  Sources / sinks are the
 * in-degree-0 / out-degree-0 nodes.  Plain C, no CUDA: the CPU tests use it too.
 *
 * Output layout == the C-ABI batch layout of include/centrolign_b200.h.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t s;
} rng_t;

static inline uint64_t rng_next(rng_t* r) { /* splitmix64 */
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double rng_unit(rng_t* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t rng_below(rng_t* r, uint32_t n) { return (uint32_t)(rng_unit(r) * n); }

static const char ALPHA[4] = {'A', 'C', 'G', 'T'};

typedef struct {
    int64_t n_windows;
    /* per side s in {0,1}; sizes: node_off[nw+1], label[N], edge_off[nw+1],
       pred_off[N+nw], pred[E], src_off[nw+1], src[], snk_off[nw+1], snk[] */
    int64_t* node_off[2];
    uint8_t* label[2];
    int64_t* edge_off[2];
    uint32_t* pred_off[2];
    uint32_t* pred[2];
    int64_t* src_off[2];
    uint32_t* src[2];
    int64_t* snk_off[2];
    uint32_t* snk[2];
} clsynth_batch;

typedef struct {
    uint8_t* label;
    uint32_t* efrom;
    uint32_t* eto;
    size_t n, e, ncap, ecap;
} gbuild;

static void gb_node(gbuild* g, uint8_t c) {
    if (g->n == g->ncap) {
        g->ncap = g->ncap ? g->ncap * 2 : 1024;
        g->label = (uint8_t*)realloc(g->label, g->ncap);
    }
    g->label[g->n++] = c;
}
static void gb_edge(gbuild* g, uint32_t a, uint32_t b) {
    if (g->e == g->ecap) {
        g->ecap = g->ecap ? g->ecap * 2 : 2048;
        g->efrom = (uint32_t*)realloc(g->efrom, g->ecap * sizeof(uint32_t));
        g->eto = (uint32_t*)realloc(g->eto, g->ecap * sizeof(uint32_t));
    }
    g->efrom[g->e] = a;
    g->eto[g->e++] = b;
}

/* mutated copy of base: substitutions + 7-bp indels */
static size_t mutate(rng_t* r, const uint8_t* base, size_t len, uint8_t* out, double sub, double indel) {
    size_t m = 0;
    for (size_t i = 0; i < len; ++i) {
        double u = rng_unit(r);
        if (u < indel) {
            if (rng_next(r) & 1) { i += 6; continue; } /* delete 7 */
            for (int k = 0; k < 7; ++k) out[m++] = ALPHA[rng_below(r, 4)];
        }
        uint8_t c = base[i];
        if (rng_unit(r) < sub) {
            uint8_t d;
            do { d = ALPHA[rng_below(r, 4)]; } while (d == c);
            c = d;
        }
        out[m++] = c;
    }
    if (m == 0) out[m++] = base[0];
    return m;
}

static void build_graph(rng_t* r, const uint8_t* seq, size_t len, gbuild* g, double snp_rate,
                        size_t alt_len, size_t alt_period) {
    g->n = g->e = 0;
    for (size_t i = 0; i < len; ++i) gb_node(g, seq[i]);
    for (size_t i = 1; i < len; ++i) gb_edge(g, (uint32_t)(i - 1), (uint32_t)i);
    for (size_t p = 1; p + 1 < len; ++p) {
        if (rng_unit(r) < snp_rate) {
            uint8_t d;
            do { d = ALPHA[rng_below(r, 4)]; } while (d == seq[p]);
            uint32_t a = (uint32_t)g->n;
            gb_node(g, d);
            gb_edge(g, (uint32_t)(p - 1), a);
            gb_edge(g, a, (uint32_t)(p + 1));
        }
    }
    if (alt_period && len > alt_len + 3) {
        size_t n_alt = (len + alt_period / 2) / alt_period;
        for (size_t b = 0; b < n_alt; ++b) {
            size_t p = 1 + rng_below(r, (uint32_t)(len - alt_len - 2)); /* p .. p+alt_len+1 < len */
            uint32_t prev = (uint32_t)(p - 1);
            for (size_t k = 0; k < alt_len; ++k) {
                uint8_t c = seq[p + k];
                if (rng_unit(r) < 0.10) c = ALPHA[rng_below(r, 4)];
                uint32_t a = (uint32_t)g->n;
                gb_node(g, c);
                gb_edge(g, prev, a);
                prev = a;
            }
            gb_edge(g, prev, (uint32_t)(p + alt_len));
        }
    }
}

typedef struct {
    size_t cap_n, cap_e, cap_s, cap_k;
} caps_t;

#define GROW(ptr, type, need, cap)                                   \
    do {                                                             \
        if ((size_t)(need) > (cap)) {                                \
            while ((size_t)(need) > (cap)) (cap) = (cap) ? (cap) * 2 : 1 << 16; \
            (ptr) = (type*)realloc((ptr), (cap) * sizeof(type));     \
        }                                                            \
    } while (0)

/*
 * Generate windows first_index .. first_index+n_windows-1 of the stream defined by `seed`.
 * Backbone length of a window is log-uniform in [len_min, len_max]; node counts come out
 * ~13.5 % larger (bubbles).  Returns a heap-allocated batch; release with clsynth_free.
 */
static clsynth_batch* generate_impl(int64_t n_windows, int64_t first_index, const int64_t* index_list, uint64_t seed, double len_min,
                                    double len_max, double snp_rate, int64_t alt_len, int64_t alt_period);

clsynth_batch* clsynth_generate(int64_t n_windows, int64_t first_index, uint64_t seed, double len_min,
                                double len_max, double snp_rate, int64_t alt_len, int64_t alt_period) {
    return generate_impl(n_windows, first_index, NULL, seed, len_min, len_max, snp_rate, alt_len, alt_period);
}

/* The windows with the stream indices index_list[0..n_windows-1] (any order): a shard of one batch. */
clsynth_batch* clsynth_generate_list(int64_t n_windows, const int64_t* index_list, uint64_t seed, double len_min,
                                     double len_max, double snp_rate, int64_t alt_len, int64_t alt_period) {
    return generate_impl(n_windows, 0, index_list, seed, len_min, len_max, snp_rate, alt_len, alt_period);
}

/* Backbone length of the windows first_index .. first_index+n-1 without generating them (cell-balanced sharding
   needs the sizes of the whole batch on every rank; node counts are ~13.5 % above the backbone length). */
void clsynth_backbone_lengths(int64_t n_windows, int64_t first_index, uint64_t seed, double len_min, double len_max, int64_t* out) {
    for (int64_t w = 0; w < n_windows; ++w) {
        rng_t r = {seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(first_index + w + 1))};
        rng_next(&r);
        double L = exp(log(len_min) + rng_unit(&r) * (log(len_max) - log(len_min)));
        int64_t len0 = (int64_t)(L + 0.5);
        out[w] = len0 < 2 ? 2 : len0;
    }
}

static clsynth_batch* generate_impl(int64_t n_windows, int64_t first_index, const int64_t* index_list, uint64_t seed, double len_min,
                                    double len_max, double snp_rate, int64_t alt_len, int64_t alt_period) {
    clsynth_batch* B = (clsynth_batch*)calloc(1, sizeof(clsynth_batch));
    B->n_windows = n_windows;
    caps_t caps[2] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    size_t capo[2] = {0, 0};
    for (int s = 0; s < 2; ++s) {
        B->node_off[s] = (int64_t*)calloc((size_t)n_windows + 1, sizeof(int64_t));
        B->edge_off[s] = (int64_t*)calloc((size_t)n_windows + 1, sizeof(int64_t));
        B->src_off[s] = (int64_t*)calloc((size_t)n_windows + 1, sizeof(int64_t));
        B->snk_off[s] = (int64_t*)calloc((size_t)n_windows + 1, sizeof(int64_t));
    }
    gbuild g = {0, 0, 0, 0, 0, 0, 0};
    uint8_t* base = NULL;
    uint8_t* seq = NULL;
    size_t cap_seq = 0;
    uint32_t* cnt = NULL;
    uint32_t* outdeg = NULL;
    size_t cap_cnt = 0;
    for (int64_t w = 0; w < n_windows; ++w) {
        const int64_t stream_index = index_list ? index_list[w] : first_index + w;
        rng_t r = {seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(stream_index + 1))};
        rng_next(&r);
        double L = exp(log(len_min) + rng_unit(&r) * (log(len_max) - log(len_min)));
        size_t len0 = (size_t)(L + 0.5);
        if (len0 < 2) len0 = 2;
        if (2 * len0 + 64 > cap_seq) {
            cap_seq = 2 * len0 + 64;
            base = (uint8_t*)realloc(base, cap_seq);
            seq = (uint8_t*)realloc(seq, cap_seq);
        }
        for (size_t i = 0; i < len0; ++i) base[i] = ALPHA[rng_below(&r, 4)];
        for (int s = 0; s < 2; ++s) {
            size_t len = mutate(&r, base, len0, seq, 0.03, 0.001);
            build_graph(&r, seq, len, &g, snp_rate, (size_t)alt_len, (size_t)alt_period);
            const int64_t n0 = B->node_off[s][w], e0 = B->edge_off[s][w];
            GROW(B->label[s], uint8_t, n0 + g.n, caps[s].cap_n);
            GROW(B->pred_off[s], uint32_t, n0 + w + g.n + 1, capo[s]);
            GROW(B->pred[s], uint32_t, e0 + g.e, caps[s].cap_e);
            memcpy(B->label[s] + n0, g.label, g.n);
            if (g.n + 1 > cap_cnt) {
                cap_cnt = 2 * (g.n + 1);
                cnt = (uint32_t*)realloc(cnt, cap_cnt * sizeof(uint32_t));
                outdeg = (uint32_t*)realloc(outdeg, cap_cnt * sizeof(uint32_t));
            }
            memset(cnt, 0, (g.n + 1) * sizeof(uint32_t));
            memset(outdeg, 0, (g.n + 1) * sizeof(uint32_t));
            for (size_t k = 0; k < g.e; ++k) { cnt[g.eto[k] + 1]++; outdeg[g.efrom[k]]++; }
            uint32_t* po = B->pred_off[s] + n0 + w;
            po[0] = 0;
            for (size_t v = 0; v < g.n; ++v) po[v + 1] = po[v] + cnt[v + 1];
            memset(cnt, 0, (g.n + 1) * sizeof(uint32_t));
            uint32_t* pr = B->pred[s] + e0;
            for (size_t k = 0; k < g.e; ++k) { uint32_t v = g.eto[k]; pr[po[v] + cnt[v]++] = g.efrom[k]; }
            for (size_t v = 0; v < g.n; ++v) { /* shuffle predecessor order */
                uint32_t d = po[v + 1] - po[v];
                for (uint32_t a = d; a > 1; --a) {
                    uint32_t b = rng_below(&r, a);
                    uint32_t t = pr[po[v] + a - 1]; pr[po[v] + a - 1] = pr[po[v] + b]; pr[po[v] + b] = t;
                }
            }
            int64_t ns = B->src_off[s][w], nk = B->snk_off[s][w];
            for (size_t v = 0; v < g.n; ++v) {
                if (po[v + 1] == po[v]) { GROW(B->src[s], uint32_t, ns + 1, caps[s].cap_s); B->src[s][ns++] = (uint32_t)v; }
                if (outdeg[v] == 0) { GROW(B->snk[s], uint32_t, nk + 1, caps[s].cap_k); B->snk[s][nk++] = (uint32_t)v; }
            }
            B->node_off[s][w + 1] = n0 + (int64_t)g.n;
            B->edge_off[s][w + 1] = e0 + (int64_t)g.e;
            B->src_off[s][w + 1] = ns;
            B->snk_off[s][w + 1] = nk;
        }
    }
    free(g.label); free(g.efrom); free(g.eto); free(base); free(seq); free(cnt); free(outdeg);
    return B;
}

void clsynth_free(clsynth_batch* B) {
    if (!B) return;
    for (int s = 0; s < 2; ++s) {
        free(B->node_off[s]); free(B->label[s]); free(B->edge_off[s]); free(B->pred_off[s]);
        free(B->pred[s]); free(B->src_off[s]); free(B->src[s]); free(B->snk_off[s]); free(B->snk[s]);
    }
    free(B);
}
