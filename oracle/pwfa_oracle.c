/*
 * oracle/pwfa_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's "wavefront variant" of the PO-to-PO gap fill,
 * `pwfa_po_poa<NumPW,Graph,BackingMap>` (reference: include/centrolign/alignment.hpp:2299-2338)
 *   -> `pwfa_po_poa_internal` (:1959-2033) -> `wfa_iteration<true>` (:1712-1827),
 * with `to_wfa_params` (:1613-1654), `wfa_traceback` (:1892-1923), `convert_wfa_score`
 * (:1877-1890), `minmax_distance` (minmax_distance.hpp:15-73) and `target_reachability`
 * (target_reachability.hpp:15-33).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call this; the CUDA product path never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (a) the golden alignment of the reference's own unit test for wfa/pwfa
 *       (src/test/test_alignment.cpp:684-773, same expected pairs as po_poa), and
 *   (b) fixtures produced by the unmodified reference (oracle/_ref/libclref.so via
 *       clref_pwfa_po_poa_succ in oracle/ref_shim.cpp; tests/golden/make_pwfa_golden.py),
 *       and live against that library whenever it is present.
 *
 * What the algorithm is (and is not): Dijkstra over states (node1, node2, component) with a
 * bucket queue, one FIFO per integer score; component 0 = match state, +k = insertion piece k-1,
 * -k = deletion piece k-1.  The FIRST dequeue of a state fixes its back-pointer; the search stops
 * at the first dequeued (sink1, sink2, 0).  Everything observable therefore hangs on the FIFO
 * order, which this file keeps literally: one queue per score, entries appended in the reference's
 * enumeration order (alignment.hpp:1776-1826).
 *
 * Graph input: SUCCESSOR lists in the graph's next() order (graph.hpp next()), because that order
 * is the enumeration order above.  Node id n (== node_size) is the virtual start whose successor
 * list is `sources` in caller order (alignment.hpp:1992-1997).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CLW_MAXPW 3

typedef struct {
    uint32_t n;
    const uint8_t* label;
    const uint32_t* next_off; /* n+1 */
    const uint32_t* next;
    uint32_t n_src;
    const uint32_t* src;
    uint32_t n_snk;
    const uint32_t* snk;
    /* derived */
    int64_t* mind; /* minmax_distance(...).first  */
    int64_t* maxd; /* minmax_distance(...).second */
    uint8_t* reach; /* target_reachability */
    uint8_t* is_sink;
} clw_graph;

/* Kahn order; any topological order gives the same distances / reachability */
static int clw_prepare(clw_graph* g) {
    const uint32_t n = g->n;
    uint32_t* indeg = (uint32_t*)calloc(n + 1, sizeof(uint32_t));
    uint32_t* order = (uint32_t*)malloc((n + 1) * sizeof(uint32_t));
    g->mind = (int64_t*)malloc((n + 1) * sizeof(int64_t));
    g->maxd = (int64_t*)malloc((n + 1) * sizeof(int64_t));
    g->reach = (uint8_t*)calloc(n + 1, 1);
    g->is_sink = (uint8_t*)calloc(n + 1, 1);
    if (!indeg || !order || !g->mind || !g->maxd || !g->reach || !g->is_sink) return -1;
    for (uint32_t e = 0; e < g->next_off[n]; ++e) {
        if (g->next[e] >= n) return -2;
        ++indeg[g->next[e]];
    }
    uint32_t cnt = 0;
    for (uint32_t v = 0; v < n; ++v)
        if (!indeg[v]) order[cnt++] = v;
    for (uint32_t k = 0; k < cnt; ++k) {
        const uint32_t v = order[k];
        for (uint32_t e = g->next_off[v]; e < g->next_off[v + 1]; ++e)
            if (--indeg[g->next[e]] == 0) order[cnt++] = g->next[e];
    }
    if (cnt != n) return -2; /* cycle */
    /* minmax_distance.hpp:22-60: (INT64_MAX, -1) = not reached from the sources */
    for (uint32_t v = 0; v < n; ++v) {
        g->mind[v] = INT64_MAX;
        g->maxd[v] = -1;
    }
    for (uint32_t k = 0; k < g->n_src; ++k) {
        if (g->src[k] >= n) return -2;
        g->mind[g->src[k]] = 0;
        g->maxd[g->src[k]] = 0;
    }
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t v = order[k];
        if (g->mind[v] == INT64_MAX) continue;
        for (uint32_t e = g->next_off[v]; e < g->next_off[v + 1]; ++e) {
            const uint32_t u = g->next[e];
            if (g->mind[v] + 1 < g->mind[u]) g->mind[u] = g->mind[v] + 1;
            if (g->maxd[v] + 1 > g->maxd[u]) g->maxd[u] = g->maxd[v] + 1;
        }
    }
    /* target_reachability.hpp:18-30 */
    for (uint32_t k = 0; k < g->n_snk; ++k) {
        if (g->snk[k] >= n) return -2;
        g->reach[g->snk[k]] = 1;
        g->is_sink[g->snk[k]] = 1;
    }
    for (uint32_t k = n; k-- > 0;) {
        const uint32_t v = order[k];
        for (uint32_t e = g->next_off[v]; e < g->next_off[v + 1]; ++e)
            if (g->reach[g->next[e]]) g->reach[v] = 1;
    }
    free(indeg);
    free(order);
    return 0;
}

static void clw_release(clw_graph* g) {
    free(g->mind);
    free(g->maxd);
    free(g->reach);
    free(g->is_sink);
}

/* successors of `v`, the virtual start (v == n) standing for the sources (alignment.hpp:1992-1997) */
static inline uint32_t clw_deg(const clw_graph* g, uint32_t v) {
    return v == g->n ? g->n_src : g->next_off[v + 1] - g->next_off[v];
}
static inline const uint32_t* clw_next(const clw_graph* g, uint32_t v) {
    return v == g->n ? g->src : g->next + g->next_off[v];
}

/* ---- queue entry, bucket queue ("std::deque<std::queue<tuple<from..., to...>>>", :1986) ---- */
typedef struct {
    uint32_t f1, f2, t1, t2;
    int8_t fc, tc;
} clw_entry;

typedef struct {
    clw_entry* e;
    size_t head, size, cap;
} clw_fifo;

typedef struct {
    clw_fifo* b; /* b[k] = bucket of score min_score + k, stored at absolute index */
    size_t nb, cap;
    int64_t min_score; /* absolute index of the front bucket */
} clw_queue;

static int clw_push(clw_queue* q, uint64_t penalty, clw_entry en) {
    const size_t idx = (size_t)q->min_score + (size_t)penalty;
    if (idx >= q->cap) {
        size_t nc = q->cap ? q->cap : 1024;
        while (nc <= idx) nc *= 2;
        clw_fifo* nb = (clw_fifo*)realloc(q->b, nc * sizeof(clw_fifo));
        if (!nb) return -1;
        memset(nb + q->cap, 0, (nc - q->cap) * sizeof(clw_fifo));
        q->b = nb;
        q->cap = nc;
    }
    if (idx >= q->nb) q->nb = idx + 1;
    clw_fifo* f = &q->b[idx];
    if (f->size == f->cap) {
        const size_t nc = f->cap ? 2 * f->cap : 8;
        clw_entry* ne = (clw_entry*)realloc(f->e, nc * sizeof(clw_entry));
        if (!ne) return -1;
        f->e = ne;
        f->cap = nc;
    }
    f->e[f->size++] = en;
    return 0;
}

/* ---- back-pointer map (HashBackedMap, :1689-1706): open addressing on (id1, id2, comp) ---- */
typedef struct {
    uint64_t key; /* UINT64_MAX = empty */
    uint32_t f1, f2;
    int32_t fc;
} clw_slot;

typedef struct {
    clw_slot* s;
    size_t cap, used;
} clw_map;

static inline uint64_t clw_key(uint32_t a, uint32_t b, int c) {
    return ((uint64_t)a << 34) | ((uint64_t)b << 4) | (uint64_t)(c + CLW_MAXPW);
}
static inline size_t clw_hash(uint64_t k, size_t cap) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return (size_t)k & (cap - 1);
}
static clw_slot* clw_find(const clw_map* m, uint64_t key) {
    size_t h = clw_hash(key, m->cap);
    while (m->s[h].key != UINT64_MAX && m->s[h].key != key) h = (h + 1) & (m->cap - 1);
    return &m->s[h];
}
static int clw_grow(clw_map* m) {
    clw_map o = *m;
    m->cap = o.cap ? 2 * o.cap : (1u << 16);
    m->s = (clw_slot*)malloc(m->cap * sizeof(clw_slot));
    if (!m->s) return -1;
    memset(m->s, 0xff, m->cap * sizeof(clw_slot));
    for (size_t i = 0; i < o.cap; ++i)
        if (o.s[i].key != UINT64_MAX) *clw_find(m, o.s[i].key) = o.s[i];
    free(o.s);
    return 0;
}

/* :1617-1628; the reference divides by zero for a zero gap_open, so such parameters are outside its domain */
static uint32_t clw_gcd(uint32_t a, uint32_t b) {
    while (b) {
        const uint32_t r = a % b;
        a = b;
        b = r;
    }
    return a;
}

/*
 * params = {match, mismatch, open[3], extend[3]} (same packing as clo_po_poa).
 * Returns 0, or <0: -1 out of memory, -2 malformed graph, -3 bad P, -4 the queue ran dry (the reference
 * would dereference an empty deque: no source reaches a sink within the pruning rule).
 * stats_out (optional): [0] states settled, [1] queue entries dequeued, [2] final WFA score.
 */
int clo_pwfa_po_poa(int P, const uint32_t* params, int64_t prune_limit,
                    uint32_t n1, const uint8_t* label1, const uint32_t* next_off1, const uint32_t* next1,
                    uint32_t nsrc1, const uint32_t* src1, uint32_t nsnk1, const uint32_t* snk1,
                    uint32_t n2, const uint8_t* label2, const uint32_t* next_off2, const uint32_t* next2,
                    uint32_t nsrc2, const uint32_t* src2, uint32_t nsnk2, const uint32_t* snk2,
                    int64_t* score_out, int32_t* aln_out, uint32_t* aln_len, int64_t* stats_out) {
    if (P < 1 || P > CLW_MAXPW) return -3;
    clw_graph g1 = {n1, label1, next_off1, next1, nsrc1, src1, nsnk1, snk1, 0, 0, 0, 0};
    clw_graph g2 = {n2, label2, next_off2, next2, nsrc2, src2, nsnk2, snk2, 0, 0, 0, 0};
    int rc = clw_prepare(&g1);
    if (rc == 0) rc = clw_prepare(&g2);
    if (rc) {
        clw_release(&g1);
        clw_release(&g2);
        return rc;
    }

    /* to_wfa_params, :1630-1651 */
    const uint32_t match = params[0];
    uint32_t w_mismatch = 2 * (params[0] + params[1]), w_open[CLW_MAXPW], w_ext[CLW_MAXPW];
    uint32_t factor = w_mismatch;
    for (int k = 0; k < P; ++k) {
        w_open[k] = 2 * params[2 + k];
        w_ext[k] = 2 * params[5 + k] + match;
        factor = clw_gcd(factor, w_open[k]);
        factor = clw_gcd(factor, w_ext[k]);
    }
    if (factor > 1) {
        w_mismatch /= factor;
        for (int k = 0; k < P; ++k) {
            w_open[k] /= factor;
            w_ext[k] /= factor;
        }
    }

    clw_queue q = {0, 0, 0, 0};
    clw_map bp = {0, 0, 0};
    int64_t furthest = INT64_MIN + prune_limit; /* :2314 */
    int64_t n_settled = 0, n_dequeued = 0;
    uint32_t end1 = UINT32_MAX, end2 = UINT32_MAX;
    rc = clw_grow(&bp);
    {
        clw_entry first = {UINT32_MAX, UINT32_MAX, n1, n2, 0, 0}; /* :1988 */
        if (rc == 0) rc = clw_push(&q, 0, first);
    }

#define PUSH(T1, T2, TC, PEN)                                   \
    do {                                                        \
        clw_entry en_ = {h1, h2, (T1), (T2), (int8_t)hc, (int8_t)(TC)}; \
        if (clw_push(&q, (PEN), en_)) rc = -1;                  \
    } while (0)

    while (rc == 0) {
        /* :1738-1744 advance to the next non-empty bucket */
        while ((size_t)q.min_score < q.nb && q.b[q.min_score].head == q.b[q.min_score].size) {
            free(q.b[q.min_score].e);
            q.b[q.min_score].e = 0;
            ++q.min_score;
        }
        if ((size_t)q.min_score >= q.nb) {
            rc = -4;
            break;
        }
        const clw_entry en = q.b[q.min_score].e[q.b[q.min_score].head++];
        ++n_dequeued;
        const uint32_t h1 = en.t1, h2 = en.t2;
        const int hc = en.tc;
        /* prune (:2317-2325) or already reached (:1752) */
        if ((h1 < n1 && !g1.reach[h1]) || (h2 < n2 && !g2.reach[h2])) continue;
        {
            const int64_t d1 = h1 != n1 ? g1.maxd[h1] : -1, d2 = h2 != n2 ? g2.maxd[h2] : -1;
            if (d1 + d2 < furthest - prune_limit) continue;
        }
        if (bp.used * 2 >= bp.cap && clw_grow(&bp)) {
            rc = -1;
            break;
        }
        clw_slot* sl = clw_find(&bp, clw_key(h1, h2, hc));
        if (sl->key != UINT64_MAX) continue;
        /* update (:2327-2334); the reachability test there is already implied by the prune above */
        {
            const int64_t d1 = h1 != n1 ? g1.mind[h1] : -1, d2 = h2 != n2 ? g2.mind[h2] : -1;
            if (d1 + d2 > furthest) furthest = d1 + d2;
        }
        sl->key = clw_key(h1, h2, hc); /* :1766 */
        sl->f1 = en.f1;
        sl->f2 = en.f2;
        sl->fc = en.fc;
        ++bp.used;
        ++n_settled;
        /* stop (:2003-2006); an empty sink list would accept any node, but then nothing is reachable */
        if (hc == 0 && (nsnk1 == 0 || (h1 < n1 && g1.is_sink[h1])) && (nsnk2 == 0 || (h2 < n2 && g2.is_sink[h2]))) {
            end1 = h1;
            end2 = h2;
            break;
        }
        const uint32_t deg1 = clw_deg(&g1, h1), deg2 = clw_deg(&g2, h2);
        const uint32_t *nx1 = clw_next(&g1, h1), *nx2 = clw_next(&g2, h2);
        if (hc == 0) {
            /* greedy (:2008-2013): unique successors on both sides, neither node a sink, labels equal */
            if (deg1 == 1 && deg2 == 1 && !(h1 < n1 && g1.is_sink[h1]) && !(h2 < n2 && g2.is_sink[h2]) &&
                label1[nx1[0]] == label2[nx2[0]]) {
                PUSH(nx1[0], nx2[0], 0, 0); /* :1785 */
            } else {
                for (uint32_t a = 0; a < deg1; ++a) { /* :1788-1798 */
                    for (uint32_t b = 0; b < deg2; ++b)
                        PUSH(nx1[a], nx2[b], 0, label1[nx1[a]] == label2[nx2[b]] ? 0 : w_mismatch);
                    for (int k = 0; k < P; ++k) PUSH(nx1[a], h2, k + 1, w_open[k] + w_ext[k]);
                }
                for (uint32_t b = 0; b < deg2; ++b) /* :1799-1805 */
                    for (int k = 0; k < P; ++k) PUSH(h1, nx2[b], -k - 1, w_open[k] + w_ext[k]);
            }
        } else {
            PUSH(h1, h2, 0, 0); /* gap close, :1810 */
            if (hc > 0)
                for (uint32_t a = 0; a < deg1; ++a) PUSH(nx1[a], h2, hc, w_ext[hc - 1]); /* :1814-1817 */
            else
                for (uint32_t b = 0; b < deg2; ++b) PUSH(h1, nx2[b], hc, w_ext[-hc - 1]); /* :1821-1824 */
        }
    }
#undef PUSH

    if (rc == 0) {
        /* wfa_traceback, :1892-1923 */
        uint32_t t1 = end1, t2 = end2, len = 0;
        int tc = 0;
        int64_t total_len = 0;
        while (t1 != n1 || t2 != n2) {
            const clw_slot* sl = clw_find(&bp, clw_key(t1, t2, tc));
            if (sl->f1 != t1 && sl->f2 != t2) {
                aln_out[2 * len] = (int32_t)t1;
                aln_out[2 * len + 1] = (int32_t)t2;
                ++len;
                total_len += 2;
            } else if (sl->f1 != t1) {
                aln_out[2 * len] = (int32_t)t1;
                aln_out[2 * len + 1] = -1;
                ++len;
                ++total_len;
            } else if (sl->f2 != t2) {
                aln_out[2 * len] = -1;
                aln_out[2 * len + 1] = (int32_t)t2;
                ++len;
                ++total_len;
            }
            t1 = sl->f1;
            t2 = sl->f2;
            tc = sl->fc;
        }
        for (uint32_t a = 0, b = len; a + 1 < b; ++a, --b) { /* :1920 reverse */
            const int32_t x = aln_out[2 * a], y = aln_out[2 * a + 1];
            aln_out[2 * a] = aln_out[2 * (b - 1)];
            aln_out[2 * a + 1] = aln_out[2 * (b - 1) + 1];
            aln_out[2 * (b - 1)] = x;
            aln_out[2 * (b - 1) + 1] = y;
        }
        *aln_len = len;
        /* convert_wfa_score, :1889 */
        if (score_out) *score_out = ((int64_t)match * total_len - q.min_score * (int64_t)factor) / 2;
        if (stats_out) {
            stats_out[0] = n_settled;
            stats_out[1] = n_dequeued;
            stats_out[2] = q.min_score;
        }
    }
    for (size_t i = 0; i < q.nb; ++i) free(q.b[i].e);
    free(q.b);
    free(bp.s);
    clw_release(&g1);
    clw_release(&g2);
    return rc;
}
