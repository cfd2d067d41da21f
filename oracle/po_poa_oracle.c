/*
 * oracle/po_poa_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's piecewise-affine PO-to-PO graph DP,
 * `po_poa<NumPW,Graph>` -> `po_poa_internal<true,...>`
 * (reference: include/centrolign/alignment.hpp:753-1163).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call this; the CUDA product path never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (a) the golden alignments of the reference's own unit test
 *       (src/test/test_alignment.cpp:684-773), transcribed into tests/golden/, and
 *   (b) outputs of the unmodified reference compiled here from its own sources
 *       (oracle/Makefile -> oracle/_ref/libclref.so), committed as fixtures in
 *       tests/golden/ by tests/golden/make_golden.py, and compared live whenever
 *       oracle/_ref/libclref.so is present.
 *
 * The restatement keeps the reference's evaluation order where it is observable:
 *   - 32-bit wrapping arithmetic on scores, "minus infinity" = INT32_MIN/2
 *     (alignment.hpp:736-751);
 *   - boundary row/column passes before the interior (alignment.hpp:814-894);
 *   - first-best sink pair in caller order, strict '>' (alignment.hpp:979-990);
 *   - traceback preference rules (alignment.hpp:1036-1138).
 * It is written in pull-free "push" form like the reference, on one flat table.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CLO_MAXPW 3
#define CLO_MININF (INT32_MIN / 2)

typedef struct {
    int32_t M;
    int32_t I[CLO_MAXPW];
    int32_t D[CLO_MAXPW];
} clo_cell;

typedef struct {
    uint32_t n;
    const uint8_t* label;
    const uint32_t* pred_off; /* n+1 */
    const uint32_t* pred;     /* predecessor ids, in the graph's previous() order */
    uint32_t n_src;
    const uint32_t* src;
    uint32_t n_snk;
    const uint32_t* snk;
    /* derived */
    uint32_t* succ_off;
    uint32_t* succ;
    uint32_t* order;
    uint8_t* is_src;
} clo_graph;

static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wsub(int32_t a, uint32_t b) { return (int32_t)((uint32_t)a - b); }
static inline int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }

/* successor lists + a Kahn topological order (topological_order.hpp:11-60; the DP
 * values do not depend on which topological order is used) */
static int derive(clo_graph* g) {
    uint32_t n = g->n, e = g->pred_off[n];
    g->succ_off = (uint32_t*)calloc((size_t)n + 2, sizeof(uint32_t));
    g->succ = (uint32_t*)malloc(((size_t)e + 1) * sizeof(uint32_t));
    g->order = (uint32_t*)malloc(((size_t)n + 1) * sizeof(uint32_t));
    g->is_src = (uint8_t*)calloc((size_t)n + 1, 1);
    uint32_t* indeg = (uint32_t*)malloc(((size_t)n + 1) * sizeof(uint32_t));
    uint32_t* stack = (uint32_t*)malloc(((size_t)n + 1) * sizeof(uint32_t));
    if (!g->succ_off || !g->succ || !g->order || !g->is_src || !indeg || !stack) return -1;
    for (uint32_t v = 0; v < n; ++v)
        for (uint32_t k = g->pred_off[v]; k < g->pred_off[v + 1]; ++k) g->succ_off[g->pred[k] + 1]++;
    for (uint32_t v = 0; v < n; ++v) g->succ_off[v + 1] += g->succ_off[v];
    uint32_t* fill = (uint32_t*)calloc((size_t)n + 1, sizeof(uint32_t));
    for (uint32_t v = 0; v < n; ++v)
        for (uint32_t k = g->pred_off[v]; k < g->pred_off[v + 1]; ++k) {
            uint32_t u = g->pred[k];
            g->succ[g->succ_off[u] + fill[u]++] = v;
        }
    free(fill);
    uint32_t sp = 0, cnt = 0;
    for (uint32_t v = 0; v < n; ++v) {
        indeg[v] = g->pred_off[v + 1] - g->pred_off[v];
        if (indeg[v] == 0) stack[sp++] = v;
    }
    while (sp) {
        uint32_t v = stack[--sp];
        g->order[cnt++] = v;
        for (uint32_t k = g->succ_off[v]; k < g->succ_off[v + 1]; ++k)
            if (--indeg[g->succ[k]] == 0) stack[sp++] = g->succ[k];
    }
    free(indeg);
    free(stack);
    for (uint32_t k = 0; k < g->n_src; ++k) g->is_src[g->src[k]] = 1;
    return cnt == n ? 0 : -2; /* -2: cyclic input */
}

static void release(clo_graph* g) {
    free(g->succ_off); free(g->succ); free(g->order); free(g->is_src);
}

/*
 * params[8] = match, mismatch, open[0..2], extend[0..2]   (alignment.hpp:56-65)
 * aln_out   = 2*(n1+n2) int32, pairs (id1,id2), -1 = gap (AlignedPair::gap)
 * returns 0, or <0 on allocation failure / cyclic input.
 */
int clo_po_poa(int P, const uint32_t* params,
               uint32_t n1, const uint8_t* label1, const uint32_t* pred_off1, const uint32_t* pred1,
               uint32_t nsrc1, const uint32_t* src1, uint32_t nsnk1, const uint32_t* snk1,
               uint32_t n2, const uint8_t* label2, const uint32_t* pred_off2, const uint32_t* pred2,
               uint32_t nsrc2, const uint32_t* src2, uint32_t nsnk2, const uint32_t* snk2,
               int64_t* score_out, int32_t* aln_out, uint32_t* aln_len) {
    if (P < 1 || P > CLO_MAXPW) return -3;
    const uint32_t match = params[0], mismatch = params[1];
    const uint32_t* open = params + 2;
    const uint32_t* ext = params + 5;
    clo_graph g1 = {n1, label1, pred_off1, pred1, nsrc1, src1, nsnk1, snk1, 0, 0, 0, 0};
    clo_graph g2 = {n2, label2, pred_off2, pred2, nsrc2, src2, nsnk2, snk2, 0, 0, 0, 0};
    int rc = derive(&g1);
    if (rc == 0) rc = derive(&g2);
    const size_t W = (size_t)n2 + 1;
    clo_cell* dp = rc == 0 ? (clo_cell*)malloc(((size_t)n1 + 1) * W * sizeof(clo_cell)) : NULL;
    if (!dp) { release(&g1); release(&g2); return rc ? rc : -1; }
    for (size_t c = 0; c < ((size_t)n1 + 1) * W; ++c) {
        dp[c].M = CLO_MININF;
        for (int k = 0; k < CLO_MAXPW; ++k) dp[c].I[k] = dp[c].D[k] = CLO_MININF;
    }
#define CELL(i, j) dp[(size_t)(i) * W + (j)]
#define SUB(i, j) (label1[i] == label2[j] ? (int32_t)match : (int32_t)(0u - mismatch))
    /* boundary initialisation (alignment.hpp:814-829) */
    for (uint32_t a = 0; a < nsrc1; ++a) {
        uint32_t i = src1[a];
        for (uint32_t b = 0; b < nsrc2; ++b) CELL(i, src2[b]).M = SUB(i, src2[b]);
        for (int k = 0; k < P; ++k) CELL(i, n2).I[k] = (int32_t)(0u - open[k] - ext[k]);
    }
    for (uint32_t b = 0; b < nsrc2; ++b)
        for (int k = 0; k < P; ++k) CELL(n1, src2[b]).D[k] = (int32_t)(0u - open[k] - ext[k]);
    /* lead-insertion column (alignment.hpp:833-862) */
    for (uint32_t oi = 0; oi < n1; ++oi) {
        uint32_t i = g1.order[oi];
        clo_cell* c = &CELL(i, n2);
        for (int k = 0; k < P; ++k) c->M = imax(c->M, c->I[k]);
        for (uint32_t s = g1.succ_off[i]; s < g1.succ_off[i + 1]; ++s) {
            clo_cell* nc = &CELL(g1.succ[s], n2);
            for (int k = 0; k < P; ++k) nc->I[k] = imax(nc->I[k], wsub(c->I[k], ext[k]));
        }
        for (uint32_t b = 0; b < nsrc2; ++b) {
            clo_cell* nc = &CELL(i, src2[b]);
            for (int k = 0; k < P; ++k) nc->D[k] = imax(nc->D[k], wsub(c->M, open[k] + ext[k]));
        }
        for (uint32_t s = g1.succ_off[i]; s < g1.succ_off[i + 1]; ++s)
            for (uint32_t b = 0; b < nsrc2; ++b) {
                uint32_t ni = g1.succ[s], nj = src2[b];
                clo_cell* nc = &CELL(ni, nj);
                nc->M = imax(nc->M, wadd(c->M, SUB(ni, nj)));
            }
    }
    /* lead-deletion row (alignment.hpp:865-894) */
    for (uint32_t oj = 0; oj < n2; ++oj) {
        uint32_t j = g2.order[oj];
        clo_cell* c = &CELL(n1, j);
        for (int k = 0; k < P; ++k) c->M = imax(c->M, c->D[k]);
        for (uint32_t s = g2.succ_off[j]; s < g2.succ_off[j + 1]; ++s) {
            clo_cell* nc = &CELL(n1, g2.succ[s]);
            for (int k = 0; k < P; ++k) nc->D[k] = imax(nc->D[k], wsub(c->D[k], ext[k]));
        }
        for (uint32_t a = 0; a < nsrc1; ++a) {
            clo_cell* nc = &CELL(src1[a], j);
            for (int k = 0; k < P; ++k) nc->I[k] = imax(nc->I[k], wsub(c->M, open[k] + ext[k]));
        }
        for (uint32_t s = g2.succ_off[j]; s < g2.succ_off[j + 1]; ++s)
            for (uint32_t a = 0; a < nsrc1; ++a) {
                uint32_t ni = src1[a], nj = g2.succ[s];
                clo_cell* nc = &CELL(ni, nj);
                nc->M = imax(nc->M, wadd(c->M, SUB(ni, nj)));
            }
    }
    /* interior, push form (alignment.hpp:898-938) */
    for (uint32_t oi = 0; oi < n1; ++oi) {
        uint32_t i = g1.order[oi];
        for (uint32_t oj = 0; oj < n2; ++oj) {
            uint32_t j = g2.order[oj];
            clo_cell* c = &CELL(i, j);
            for (int k = 0; k < P; ++k) c->M = imax(c->M, imax(c->I[k], c->D[k]));
            for (uint32_t s = g1.succ_off[i]; s < g1.succ_off[i + 1]; ++s) {
                clo_cell* nc = &CELL(g1.succ[s], j);
                for (int k = 0; k < P; ++k)
                    nc->I[k] = imax(nc->I[k], imax(wsub(c->M, open[k] + ext[k]), wsub(c->I[k], ext[k])));
            }
            for (uint32_t t = g2.succ_off[j]; t < g2.succ_off[j + 1]; ++t) {
                clo_cell* nc = &CELL(i, g2.succ[t]);
                for (int k = 0; k < P; ++k)
                    nc->D[k] = imax(nc->D[k], imax(wsub(c->M, open[k] + ext[k]), wsub(c->D[k], ext[k])));
            }
            for (uint32_t s = g1.succ_off[i]; s < g1.succ_off[i + 1]; ++s)
                for (uint32_t t = g2.succ_off[j]; t < g2.succ_off[j + 1]; ++t) {
                    uint32_t ni = g1.succ[s], nj = g2.succ[t];
                    clo_cell* nc = &CELL(ni, nj);
                    nc->M = imax(nc->M, wadd(c->M, SUB(ni, nj)));
                }
        }
    }
    /* best sink pair: first maximum in caller order (alignment.hpp:979-1008) */
    const uint32_t NONE = UINT32_MAX;
    uint32_t t1 = NONE, t2 = NONE;
    if (n1 != 0 && n2 != 0) {
        for (uint32_t a = 0; a < nsnk1; ++a)
            for (uint32_t b = 0; b < nsnk2; ++b)
                if (t1 == NONE || CELL(snk1[a], snk2[b]).M > CELL(t1, t2).M) { t1 = snk1[a]; t2 = snk2[b]; }
    } else if (n1 != 0) {
        for (uint32_t a = 0; a < nsnk1; ++a)
            if (t1 == NONE || CELL(snk1[a], 0).M > CELL(t1, 0).M) { t1 = snk1[a]; t2 = 0; }
    } else if (n2 != 0) {
        for (uint32_t b = 0; b < nsnk2; ++b)
            if (t2 == NONE || CELL(0, snk2[b]).M > CELL(0, t2).M) { t1 = 0; t2 = snk2[b]; }
    }
    if (score_out) *score_out = (t1 != NONE) ? (int64_t)CELL(t1, t2).M : 0;
    /* traceback (alignment.hpp:1036-1138) */
    uint32_t len = 0;
    int comp = 0; /* 0 = M, +k = I[k-1], -k = D[k-1] */
    const uint32_t cap = n1 + n2;
    while (t1 != NONE && t2 != NONE) {
        const uint32_t h1 = t1, h2 = t2;
        t1 = t2 = NONE;
        const clo_cell* c = &CELL(h1, h2);
        if (comp == 0) {
            for (int k = 0; k < P; ++k) {
                if (c->M == c->I[k]) { comp = k + 1; break; }
                if (c->M == c->D[k]) { comp = -k - 1; break; }
            }
        }
        /* predecessor lists: graph order, then the boundary index for sources */
        const uint32_t np1 = h1 < n1 ? pred_off1[h1 + 1] - pred_off1[h1] : 0;
        const uint32_t np2 = h2 < n2 ? pred_off2[h2 + 1] - pred_off2[h2] : 0;
        const uint32_t x1 = (h1 < n1 && g1.is_src[h1]) ? 1 : 0;
        const uint32_t x2 = (h2 < n2 && g2.is_src[h2]) ? 1 : 0;
#define P1(a) ((a) < np1 ? pred1[pred_off1[h1] + (a)] : n1)
#define P2(b) ((b) < np2 ? pred2[pred_off2[h2] + (b)] : n2)
        if (len >= cap) { rc = -4; break; } /* cannot happen on valid inputs */
        if (comp == 0) {
            aln_out[2 * len] = (int32_t)h1; aln_out[2 * len + 1] = (int32_t)h2; ++len;
            const int32_t s = SUB(h1, h2);
            for (uint32_t a = 0; a < np1 + x1; ++a)
                for (uint32_t b = 0; b < np2 + x2; ++b)
                    if (wadd(CELL(P1(a), P2(b)).M, s) == c->M) { t1 = P1(a); t2 = P2(b); break; }
        } else if (comp > 0) {
            aln_out[2 * len] = (int32_t)h1; aln_out[2 * len + 1] = -1; ++len;
            const int k = comp - 1;
            for (uint32_t a = 0; a < np1 + x1; ++a) {
                const clo_cell* pc = &CELL(P1(a), h2);
                if (c->I[k] == wsub(pc->M, open[k] + ext[k])) { comp = 0; t1 = P1(a); t2 = h2; break; }
                if (c->I[k] == wsub(pc->I[k], ext[k])) { t1 = P1(a); t2 = h2; break; }
            }
        } else {
            aln_out[2 * len] = -1; aln_out[2 * len + 1] = (int32_t)h2; ++len;
            const int k = -comp - 1;
            for (uint32_t b = 0; b < np2 + x2; ++b) {
                const clo_cell* pc = &CELL(h1, P2(b));
                if (c->D[k] == wsub(pc->M, open[k] + ext[k])) { comp = 0; t1 = h1; t2 = P2(b); break; }
                if (c->D[k] == wsub(pc->D[k], ext[k])) { t1 = h1; t2 = P2(b); break; }
            }
        }
    }
    /* reverse into forward order (alignment.hpp:1141) */
    for (uint32_t a = 0, b = len ? len - 1 : 0; a < b; ++a, --b) {
        int32_t x = aln_out[2 * a], y = aln_out[2 * a + 1];
        aln_out[2 * a] = aln_out[2 * b]; aln_out[2 * a + 1] = aln_out[2 * b + 1];
        aln_out[2 * b] = x; aln_out[2 * b + 1] = y;
    }
    *aln_len = len;
    free(dp);
    release(&g1);
    release(&g2);
    return rc;
}
