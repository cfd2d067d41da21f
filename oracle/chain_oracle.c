/*
 * oracle/chain_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's sparse anchor-chaining DP on the flat problem layout of
 * include/centrolign_b200.h (clb_chain_problem):
 *   Anchorer::sparse_affine_chain_dp  main loop   include/centrolign/anchorer.hpp:2290-2417
 *   Anchorer::sparse_chain_dp         main loop   include/centrolign/anchorer.hpp:1640-1728
 *   Anchorer::traceback_sparse_dp                 include/centrolign/anchorer.hpp:2473-2547
 *   MaxSearchTree                                 include/centrolign/max_search_tree.hpp:93-444
 *   OrthogonalMaxSearchTree                       include/centrolign/orthogonal_max_search_tree.hpp:105-520
 *   MatchBank::update_dp                          include/centrolign/match_bank.hpp:171-184
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call this;
 * the CUDA product path never does.
 *
 * The search trees are restated LITERALLY -- implicit heap filled in key order, subtree_max pointers that
 * move only on a strictly greater value, the reference's own update / reidentify / range_max walks, cross
 * trees holding (value, outer index) pairs compared lexicographically -- because which of several equal
 * maxima a query returns is decided by exactly these details, and the chain is only reproducible with them.
 * (The CUDA kernels use a different value representation, see centrolign_b200/csrc/chain_device.cuh; this file
 * is the independent check that the two agree.)
 *
 * Parity status: PINNED.  tests/test_oracle.py runs this file on tests/golden/chain_golden.npz -- flat problems
 * written by the product host layer from the reference's own objects, with the chains the unmodified reference
 * returned for them (oracle/chain_shim.cpp -> oracle/_ref/chain_fixture; tests/golden/make_chain_golden.py) --
 * and requires the identical chain for all of them.
 */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CLC_LOWEST (-FLT_MAX) /* std::numeric_limits<float>::lowest() */

/* ------------------------------------------------------------------------------------------------------------
 * MaxSearchTree<K, V>: V is (float, uint32) compared like std::pair; the plain-float trees keep idx == 0.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    float v;
    uint32_t idx;
} clc_val;

static inline int clc_gt(clc_val a, clc_val b) { return a.v > b.v || (!(b.v > a.v) && a.idx > b.idx); }

typedef struct {
    size_t n;
    int64_t* key;
    clc_val* val;
    uint32_t* smax; /* subtree_max */
} clc_mst;

#define LEFT(x) (2 * (x) + 1)
#define RIGHT(x) (2 * (x) + 2)
#define PARENT(x) (((x)-1) / 2)

/* keys must already be in the tree's order (the reference stable-sorts by key, max_search_tree.hpp:100-106) */
static int clc_mst_build(clc_mst* t, size_t n, const int64_t* keys, const clc_val* vals) {
    t->n = n;
    t->key = (int64_t*)malloc((n ? n : 1) * sizeof(int64_t));
    t->val = (clc_val*)malloc((n ? n : 1) * sizeof(clc_val));
    t->smax = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    size_t* stack = (size_t*)malloc(66 * sizeof(size_t));
    if (!t->key || !t->val || !t->smax || !stack) return -1;
    size_t sp = 0, vec_idx = 0, cur = 0;
    while (cur < n || sp) { /* in-order walk, :119-150 */
        while (cur < n) {
            stack[sp++] = cur;
            cur = LEFT(cur);
        }
        const size_t x = stack[--sp];
        t->key[x] = keys[vec_idx];
        t->val[x] = vals[vec_idx];
        t->smax[x] = (uint32_t)x;
        ++vec_idx;
        cur = RIGHT(x);
    }
    free(stack);
    for (size_t i = n; i-- > 1;) { /* :159-164 */
        const size_t par = PARENT(i);
        if (clc_gt(t->val[t->smax[i]], t->val[t->smax[par]])) t->smax[par] = t->smax[i];
    }
    return 0;
}

static void clc_mst_free(clc_mst* t) {
    free(t->key);
    free(t->val);
    free(t->smax);
    t->n = 0;
}

static void clc_mst_reidentify(clc_mst* t, size_t x) { /* :300-310 */
    size_t new_max = x;
    const size_t l = LEFT(x), r = RIGHT(x);
    if (l < t->n && clc_gt(t->val[t->smax[l]], t->val[new_max])) new_max = t->smax[l];
    if (r < t->n && clc_gt(t->val[t->smax[r]], t->val[new_max])) new_max = t->smax[r];
    t->smax[x] = (uint32_t)new_max;
}

static void clc_mst_update(clc_mst* t, size_t i, clc_val nv) { /* :312-358 */
    if (clc_gt(nv, t->val[t->smax[i]])) {
        t->smax[i] = (uint32_t)i;
        size_t here = i;
        while (here != 0) {
            here = PARENT(here);
            if (clc_gt(nv, t->val[t->smax[here]])) t->smax[here] = (uint32_t)i;
            else break;
        }
        t->val[i] = nv;
    } else {
        t->val[i] = nv;
        if (t->smax[i] == i) {
            clc_mst_reidentify(t, i);
            size_t here = i;
            while (here != 0) {
                here = PARENT(here);
                if (t->smax[here] != i) break;
                clc_mst_reidentify(t, here);
            }
        }
    }
}

static size_t clc_mst_find(const clc_mst* t, int64_t key) { /* :218-231 */
    size_t cursor = 0;
    while (cursor < t->n) {
        if (t->key[cursor] == key) return cursor;
        cursor = t->key[cursor] > key ? LEFT(cursor) : RIGHT(cursor);
    }
    return t->n;
}

/* first node of equal_range(key) (:234-296) and the iterator's ++ (in-order successor) */
static size_t clc_mst_lower(const clc_mst* t, int64_t key) {
    size_t lower = t->n, cursor = 0;
    while (cursor < t->n) {
        if (t->key[cursor] == key) {
            lower = cursor;
            cursor = LEFT(cursor);
        } else if (t->key[cursor] > key) {
            cursor = LEFT(cursor);
        } else {
            cursor = RIGHT(cursor);
        }
    }
    return lower;
}
static size_t clc_mst_next(const clc_mst* t, size_t i) {
    if (RIGHT(i) < t->n) {
        i = RIGHT(i);
        while (LEFT(i) < t->n) i = LEFT(i);
        return i;
    }
    while (i != 0 && i == RIGHT(PARENT(i))) i = PARENT(i);
    return i == 0 ? t->n : PARENT(i);
}

static size_t clc_mst_range_max(const clc_mst* t, int64_t lo, int64_t hi) { /* :361-444 */
    size_t cursor = 0;
    while (cursor < t->n && (t->key[cursor] < lo || t->key[cursor] >= hi)) cursor = t->key[cursor] >= lo ? LEFT(cursor) : RIGHT(cursor);
    if (cursor >= t->n) return t->n;
    size_t max_idx = cursor, right_cursor = RIGHT(cursor), left_cursor = LEFT(cursor);
    while (left_cursor < t->n) {
        if (t->key[left_cursor] >= lo) {
            if (clc_gt(t->val[left_cursor], t->val[max_idx])) max_idx = left_cursor;
            const size_t r = RIGHT(left_cursor);
            if (r < t->n && clc_gt(t->val[t->smax[r]], t->val[max_idx])) max_idx = t->smax[r];
            left_cursor = LEFT(left_cursor);
        } else {
            left_cursor = RIGHT(left_cursor);
        }
    }
    while (right_cursor < t->n) {
        if (t->key[right_cursor] < hi) {
            if (clc_gt(t->val[right_cursor], t->val[max_idx])) max_idx = right_cursor;
            const size_t l = LEFT(right_cursor);
            if (l < t->n && clc_gt(t->val[t->smax[l]], t->val[max_idx])) max_idx = t->smax[l];
            right_cursor = RIGHT(right_cursor);
        } else {
            right_cursor = LEFT(right_cursor);
        }
    }
    return max_idx;
}

/* ------------------------------------------------------------------------------------------------------------
 * OrthogonalMaxSearchTree<K1 = (shift, match), K2 = offset, V = float>
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    size_t n;
    int64_t* key1; /* shift << 32 | match: orders like the pair (shift, match) */
    uint32_t* key2;
    float* val;
    clc_mst* cross;
} clc_omst;

typedef struct {
    int64_t key1;
    uint32_t key2;
    uint32_t pos; /* position in the outer key order */
} clc_elem;

static int clc_cmp_key2(const void* a, const void* b) { /* stable sort on key 2 only: ties keep the outer order */
    const clc_elem *x = (const clc_elem*)a, *y = (const clc_elem*)b;
    if (x->key2 != y->key2) return x->key2 < y->key2 ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos);
}

/* data sorted by (key1, key2); all values start at lowest() (anchorer.hpp:2046) */
static int clc_omst_build(clc_omst* t, size_t n, const clc_elem* data) {
    memset(t, 0, sizeof(*t));
    t->n = n;
    if (!n) return 0;
    t->key1 = (int64_t*)malloc(n * sizeof(int64_t));
    t->key2 = (uint32_t*)malloc(n * sizeof(uint32_t));
    t->val = (float*)malloc(n * sizeof(float));
    t->cross = (clc_mst*)calloc(n, sizeof(clc_mst));
    uint32_t* indexes = (uint32_t*)malloc(n * sizeof(uint32_t)); /* node of the element at each ordinal position */
    uint32_t* pos_of = (uint32_t*)malloc(n * sizeof(uint32_t));
    uint8_t* make_cross = (uint8_t*)malloc(n);
    size_t* stack = (size_t*)malloc(66 * sizeof(size_t));
    if (!t->key1 || !t->key2 || !t->val || !t->cross || !indexes || !pos_of || !make_cross || !stack) return -1;
    size_t sp = 0, vec_idx = 0, cur = 0;
    while (cur < n || sp) { /* :144-171 */
        while (cur < n) {
            stack[sp++] = cur;
            cur = LEFT(cur);
        }
        const size_t x = stack[--sp];
        indexes[vec_idx] = (uint32_t)x;
        pos_of[x] = (uint32_t)vec_idx;
        t->key1[x] = data[vec_idx].key1;
        t->key2[x] = data[vec_idx].key2;
        t->val[x] = CLC_LOWEST;
        ++vec_idx;
        cur = RIGHT(x);
    }
    memset(make_cross, 1, n); /* :174-182: the outermost nodes' cross trees are never queried */
    for (size_t c = 0; c < n; c = LEFT(c)) make_cross[c] = 0;
    for (size_t c = RIGHT(0); c < n; c = RIGHT(c)) make_cross[c] = 0;
    /* :187-239: each node's cross tree holds the node's whole subtree.  The keys were laid out in order, so a
     * subtree is a contiguous range of ordinal positions: [first position of its leftmost node, last of its rightmost] */
    int rc = 0;
    clc_elem* tmp = (clc_elem*)malloc(n * sizeof(clc_elem));
    int64_t* ckeys = (int64_t*)malloc(n * sizeof(int64_t));
    clc_val* cvals = (clc_val*)malloc(n * sizeof(clc_val));
    if (!tmp || !ckeys || !cvals) rc = -1;
    for (size_t x = 0; x < n && rc == 0; ++x) {
        if (!make_cross[x]) continue;
        size_t lo = x, hi = x;
        while (LEFT(lo) < n) lo = LEFT(lo);
        while (RIGHT(hi) < n) hi = RIGHT(hi);
        size_t m = 0;
        for (uint32_t p = pos_of[lo]; p <= pos_of[hi]; ++p) { /* the ordinal positions between its leftmost and rightmost node */
            tmp[m].key1 = data[p].key1;
            tmp[m].key2 = data[p].key2;
            tmp[m].pos = p;
            ++m;
        }
        qsort(tmp, m, sizeof(clc_elem), clc_cmp_key2);
        for (size_t k = 0; k < m; ++k) {
            ckeys[k] = (int64_t)tmp[k].key2;
            cvals[k].v = CLC_LOWEST;
            cvals[k].idx = indexes[tmp[k].pos];
        }
        rc = clc_mst_build(&t->cross[x], m, ckeys, cvals);
    }
    free(tmp);
    free(ckeys);
    free(cvals);
    free(indexes);
    free(pos_of);
    free(make_cross);
    free(stack);
    return rc;
}

static void clc_omst_free(clc_omst* t) {
    if (t->cross)
        for (size_t x = 0; x < t->n; ++x)
            if (t->cross[x].n) clc_mst_free(&t->cross[x]);
    free(t->cross);
    free(t->key1);
    free(t->key2);
    free(t->val);
}

static size_t clc_omst_find(const clc_omst* t, int64_t k1, uint32_t k2) { /* :276-291 */
    size_t cursor = 0;
    while (cursor < t->n) {
        if (t->key1[cursor] == k1 && t->key2[cursor] == k2) return cursor;
        if (t->key1[cursor] > k1 || (t->key1[cursor] == k1 && t->key2[cursor] > k2)) cursor = LEFT(cursor);
        else cursor = RIGHT(cursor);
    }
    return t->n;
}

static void clc_omst_update(clc_omst* t, size_t i, float nv) { /* :294-338 */
    t->val[i] = nv;
    for (size_t cursor = i; cursor < t->n; cursor = cursor ? PARENT(cursor) : t->n) {
        clc_mst* cross = &t->cross[cursor];
        if (cross->n == 0) break;
        size_t it = clc_mst_lower(cross, (int64_t)t->key2[i]);
        while (cross->val[it].idx != i) it = clc_mst_next(cross, it);
        clc_val v = {nv, (uint32_t)i};
        clc_mst_update(cross, it, v);
    }
}

/* range_max over [lo1, hi1) x [lo2, hi2) (:340-520); returns the outer node or n */
static size_t clc_omst_range_max(const clc_omst* t, int64_t lo1, int64_t hi1, uint32_t lo2, uint32_t hi2) {
    size_t cursor = 0;
    while (cursor < t->n && (t->key1[cursor] < lo1 || t->key1[cursor] >= hi1)) cursor = t->key1[cursor] >= hi1 ? LEFT(cursor) : RIGHT(cursor);
    if (cursor >= t->n) return t->n;
    int have = 0;
    float best = 0.0f;
    size_t best_idx = t->n;
#define IS_OPT(V) (!have || (V) > best)
    if (t->key2[cursor] >= lo2 && t->key2[cursor] < hi2) {
        have = 1;
        best = t->val[cursor];
        best_idx = cursor;
    }
    size_t right_cursor = RIGHT(cursor), left_cursor = LEFT(cursor);
    while (left_cursor < t->n) {
        if (t->key1[left_cursor] >= lo1) {
            if (t->key2[left_cursor] >= lo2 && t->key2[left_cursor] < hi2 && IS_OPT(t->val[left_cursor])) {
                have = 1;
                best = t->val[left_cursor];
                best_idx = left_cursor;
            }
            const size_t r = RIGHT(left_cursor);
            if (r < t->n) {
                const size_t it = clc_mst_range_max(&t->cross[r], (int64_t)lo2, (int64_t)hi2);
                if (it != t->cross[r].n && IS_OPT(t->cross[r].val[it].v)) {
                    have = 1;
                    best = t->cross[r].val[it].v;
                    best_idx = t->cross[r].val[it].idx;
                }
            }
            left_cursor = LEFT(left_cursor);
        } else {
            left_cursor = RIGHT(left_cursor);
        }
    }
    while (right_cursor < t->n) {
        if (t->key1[right_cursor] < hi1) {
            if (t->key2[right_cursor] >= lo2 && t->key2[right_cursor] < hi2 && IS_OPT(t->val[right_cursor])) {
                have = 1;
                best = t->val[right_cursor];
                best_idx = right_cursor;
            }
            const size_t l = LEFT(right_cursor);
            if (l < t->n) {
                const size_t it = clc_mst_range_max(&t->cross[l], (int64_t)lo2, (int64_t)hi2);
                if (it != t->cross[l].n && IS_OPT(t->cross[l].val[it].v)) {
                    have = 1;
                    best = t->cross[l].val[it].v;
                    best_idx = t->cross[l].val[it].idx;
                }
            }
            right_cursor = RIGHT(right_cursor);
        } else {
            right_cursor = LEFT(right_cursor);
        }
    }
#undef IS_OPT
    return have ? best_idx : t->n;
}

/* ------------------------------------------------------------------------------------------------------------
 * The DP
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t pair;
    int32_t shift;
    uint32_t offset, match, entry;
} clc_entry;

static int clc_cmp_gf(const void* a, const void* b) { /* (pair, shift, offset, match) */
    const clc_entry *x = (const clc_entry*)a, *y = (const clc_entry*)b;
    if (x->pair != y->pair) return x->pair < y->pair ? -1 : 1;
    if (x->shift != y->shift) return x->shift < y->shift ? -1 : 1;
    if (x->offset != y->offset) return x->offset < y->offset ? -1 : 1;
    return x->match < y->match ? -1 : (x->match > y->match);
}
static int clc_cmp_or(const void* a, const void* b) { /* (pair, shift, match) */
    const clc_entry *x = (const clc_entry*)a, *y = (const clc_entry*)b;
    if (x->pair != y->pair) return x->pair < y->pair ? -1 : 1;
    if (x->shift != y->shift) return x->shift < y->shift ? -1 : 1;
    return x->match < y->match ? -1 : (x->match > y->match);
}

static inline void clc_update_dp(float* dp, int64_t* bp, uint32_t m, float value, int64_t from) { /* match_bank.hpp:171-184 */
    if (value > dp[m]) {
        dp[m] = value;
        bp[m] = from;
    }
}

/*
 * Arguments mirror clb_chain_problem field by field.  Returns 0, -1 out of memory, -2 malformed input.
 */
int clo_chain_dp(int num_pw, const double* gap_open, const double* gap_extend, double scale, int n_chain1, int n_chain2,
                 int64_t n_match, const float* weight, const float* dp_init, const float* final_term, float min_score,
                 int64_t n_step, const int64_t* end_off, const uint32_t* end_match, const int64_t* qry_off,
                 const uint32_t* qry_match, const uint32_t* qry_chain1, const int64_t* ins_off, const uint32_t* ins_p1,
                 const uint32_t* ins_p2, const int32_t* ins_shift, const uint32_t* ins_offset, const uint8_t* ins_active,
                 const int32_t* qa1, const int32_t* qa2, const uint32_t* qoff, float* dp_out, int64_t* backptr_out,
                 int64_t* chain_out, int64_t* chain_len, float* opt_score) {
    const int64_t M = n_match, E = M ? ins_off[M] : 0;
    const int C1 = n_chain1, C2 = n_chain2, P = num_pw;
    const size_t npair = (size_t)C1 * C2;
    *chain_len = 0;
    if (opt_score) *opt_score = CLC_LOWEST;
    if (M == 0) return 0;
    int rc = 0;
    float* dp = (float*)malloc(M * sizeof(float));
    int64_t* bp = (int64_t*)malloc(M * sizeof(int64_t));
    clc_entry* ent = (clc_entry*)malloc((E ? E : 1) * sizeof(clc_entry));
    clc_entry* srt = (clc_entry*)malloc((E ? E : 1) * sizeof(clc_entry));
    /* gap-free trees: one per (pair, diagonal) */
    size_t n_grp = 0;
    clc_mst* gf = NULL;
    int64_t *grp_first = NULL, *pair_grp = (int64_t*)calloc(npair + 2, sizeof(int64_t));
    int32_t* grp_shift = NULL;
    uint32_t* ent_grp = (uint32_t*)malloc((E ? E : 1) * sizeof(uint32_t));
    uint32_t* ent_node = (uint32_t*)malloc((E ? E : 1) * sizeof(uint32_t));
    clc_omst* ortho = (clc_omst*)calloc((size_t)(P ? 2 * P : 1) * npair, sizeof(clc_omst));
    uint32_t* ent_onode = (uint32_t*)malloc((E ? E : 1) * sizeof(uint32_t));
    if (!dp || !bp || !ent || !srt || !pair_grp || !ent_grp || !ent_node || !ortho || !ent_onode) {
        rc = -1;
        goto done;
    }
    for (int64_t m = 0; m < M; ++m) {
        dp[m] = dp_init[m]; /* update_dp(*it, weight, max()) on a fresh entry (anchorer.hpp:2041, 1579) */
        bp[m] = -1;
        for (int64_t e = ins_off[m]; e < ins_off[m + 1]; ++e) {
            if (ins_p1[e] >= (uint32_t)C1 || ins_p2[e] >= (uint32_t)C2) {
                rc = -2;
                goto done;
            }
            ent[e].pair = ins_p1[e] * (uint32_t)C2 + ins_p2[e];
            ent[e].shift = ins_shift[e];
            ent[e].offset = ins_offset[e];
            ent[e].match = (uint32_t)m;
            ent[e].entry = (uint32_t)e;
        }
    }
    /* ---- gap-free search trees (anchorer.hpp:2136-2241; sparse_chain_dp :1538-1604) ---- */
    memcpy(srt, ent, E * sizeof(clc_entry));
    qsort(srt, E, sizeof(clc_entry), clc_cmp_gf);
    for (int64_t i = 0; i < E; ++i)
        if (i == 0 || srt[i].pair != srt[i - 1].pair || srt[i].shift != srt[i - 1].shift) ++n_grp;
    gf = (clc_mst*)calloc(n_grp ? n_grp : 1, sizeof(clc_mst));
    grp_first = (int64_t*)malloc((n_grp + 1) * sizeof(int64_t));
    grp_shift = (int32_t*)malloc((n_grp ? n_grp : 1) * sizeof(int32_t));
    if (!gf || !grp_first || !grp_shift) {
        rc = -1;
        goto done;
    }
    {
        size_t g = 0;
        int64_t* keys = (int64_t*)malloc((E ? E : 1) * sizeof(int64_t));
        clc_val* vals = (clc_val*)malloc((E ? E : 1) * sizeof(clc_val));
        if (!keys || !vals) rc = -1;
        for (int64_t i = 0; i < E && rc == 0;) {
            int64_t j = i;
            while (j < E && srt[j].pair == srt[i].pair && srt[j].shift == srt[i].shift) ++j;
            for (int64_t k = i; k < j; ++k) {
                keys[k - i] = ((int64_t)srt[k].offset << 32) | srt[k].match; /* gf_key_t (offset, match) */
                vals[k - i].v = CLC_LOWEST;
                vals[k - i].idx = 0;
            }
            rc = clc_mst_build(&gf[g], (size_t)(j - i), keys, vals);
            grp_first[g] = i;
            grp_shift[g] = srt[i].shift;
            pair_grp[srt[i].pair + 1] += 1;
            for (int64_t k = i; k < j && rc == 0; ++k) {
                ent_grp[srt[k].entry] = (uint32_t)g;
                ent_node[srt[k].entry] = (uint32_t)clc_mst_find(&gf[g], keys[k - i]);
            }
            ++g;
            i = j;
        }
        free(keys);
        free(vals);
        for (size_t pr = 0; pr < npair; ++pr) pair_grp[pr + 1] += pair_grp[pr];
        if (rc) goto done;
    }
    /* ---- orthogonal search trees, 2 * NumPW per pair over the same keys (anchorer.hpp:2084-2111) ---- */
    if (P > 0) {
        memcpy(srt, ent, E * sizeof(clc_entry));
        qsort(srt, E, sizeof(clc_entry), clc_cmp_or);
        clc_elem* data = (clc_elem*)malloc((E ? E : 1) * sizeof(clc_elem));
        if (!data) {
            rc = -1;
            goto done;
        }
        for (int64_t i = 0; i < E && rc == 0;) {
            int64_t j = i;
            while (j < E && srt[j].pair == srt[i].pair) ++j;
            for (int64_t k = i; k < j; ++k) {
                data[k - i].key1 = ((int64_t)srt[k].shift << 32) | srt[k].match; /* key_t (shift, match) */
                data[k - i].key2 = srt[k].offset;
                data[k - i].pos = (uint32_t)(k - i);
            }
            for (int pw = 0; pw < 2 * P && rc == 0; ++pw) rc = clc_omst_build(&ortho[(size_t)pw * npair + srt[i].pair], (size_t)(j - i), data);
            for (int64_t k = i; k < j && rc == 0; ++k)
                ent_onode[srt[k].entry] = (uint32_t)clc_omst_find(&ortho[srt[i].pair], data[k - i].key1, data[k - i].key2);
            i = j;
        }
        free(data);
        if (rc) goto done;
    }

    /* ---- main loop over the nodes of graph 1 in topological order ---- */
    for (int64_t s = 0; s < n_step; ++s) {
        for (int64_t k = end_off[s]; k < end_off[s + 1]; ++k) { /* anchorer.hpp:2301-2345 / 1653-1667 */
            const uint32_t m = end_match[k];
            const float dp_val = dp[m];
            for (int64_t e = ins_off[m]; e < ins_off[m + 1]; ++e) {
                if (ins_active && !ins_active[e]) continue;
                {
                    clc_mst* tree = &gf[ent_grp[e]];
                    clc_val v = {dp_val, 0};
                    if (P > 0 || tree->val[ent_node[e]].v < dp_val) clc_mst_update(tree, ent_node[e], v); /* :2323 unconditional, :1664 guarded */
                }
                for (int pw = 0; pw < 2 * P; ++pw) {
                    float value;
                    if (pw % 2 == 1) value = (float)(dp_val + scale * gap_extend[pw / 2] * ins_shift[e]); /* :2328-2335 */
                    else value = (float)(dp_val - scale * gap_extend[pw / 2] * ins_shift[e]);
                    clc_omst* tree = &ortho[(size_t)pw * npair + ent[e].pair];
                    if (value > tree->val[ent_onode[e]]) clc_omst_update(tree, ent_onode[e], value); /* :2338-2341 */
                }
            }
        }
        for (int64_t k = qry_off[s]; k < qry_off[s + 1]; ++k) { /* anchorer.hpp:2352-2416 / 1676-1727 */
            const uint32_t m = qry_match[k], chain1 = qry_chain1[k];
            const float w = weight[m];
            for (int chain2 = 0; chain2 < C2; ++chain2) {
                const int32_t query = (int32_t)((uint32_t)qa1[(int64_t)m * C1 + chain1] - (uint32_t)qa2[(int64_t)m * C2 + chain2]);
                const uint32_t offset = qoff[(int64_t)m * C2 + chain2];
                const size_t pair = (size_t)chain1 * C2 + chain2;
                {   /* same diagonal, :2379-2389 (sparse_chain_dp: its single tree per chain pair, :1709-1724) */
                    int64_t g = -1, lo_g = pair_grp[pair], hi_g = pair_grp[pair + 1]; /* the diagonal with shift == query, if any */
                    while (lo_g < hi_g) {
                        const int64_t mid = (lo_g + hi_g) / 2;
                        if (grp_shift[mid] < query) lo_g = mid + 1;
                        else hi_g = mid;
                    }
                    if (lo_g < pair_grp[pair + 1] && grp_shift[lo_g] == query) g = lo_g;
                    if (g >= 0) {
                        const clc_mst* tree = &gf[g];
                        const size_t it = clc_mst_range_max(tree, 0, (int64_t)offset << 32);
                        if (it != tree->n) {
                            const float value = tree->val[it].v + w;
                            clc_update_dp(dp, bp, m, value, (int64_t)(tree->key[it] & 0xffffffff));
                        }
                    }
                }
                for (int pw = 0; pw < 2 * P; ++pw) { /* :2390-2413 */
                    const clc_omst* tree = &ortho[(size_t)pw * npair + pair];
                    if (pw % 2 == 1) {
                        const size_t it = clc_omst_range_max(tree, INT64_MIN, (int64_t)query << 32, 0, offset);
                        if (it != tree->n) {
                            const float value = (float)((tree->val[it] + w) - scale * (gap_open[pw / 2] + gap_extend[pw / 2] * query));
                            clc_update_dp(dp, bp, m, value, (int64_t)(tree->key1[it] & 0xffffffff));
                        }
                    } else {
                        const size_t it = clc_omst_range_max(tree, ((int64_t)query + 1) << 32, ((int64_t)INT32_MAX << 32) | 0xffffffffll, 0, offset);
                        if (it != tree->n) {
                            const float value = (float)((tree->val[it] + w) - scale * (gap_open[pw / 2] - gap_extend[pw / 2] * query));
                            clc_update_dp(dp, bp, m, value, (int64_t)(tree->key1[it] & 0xffffffff));
                        }
                    }
                }
            }
        }
    }

    /* ---- traceback_sparse_dp (anchorer.hpp:2483-2534) ---- */
    {
        float opt_value = CLC_LOWEST;
        int64_t opt = -1, len = 0;
        for (int64_t m = 0; m < M; ++m) {
            float dp_val = dp[m];
            if (final_term[m] == CLC_LOWEST) dp_val = final_term[m];
            else dp_val += final_term[m];
            if (dp_val > opt_value && dp_val > min_score) {
                opt_value = dp_val;
                opt = m;
            }
        }
        for (int64_t here = opt; here >= 0; here = bp[here]) chain_out[len++] = here;
        for (int64_t a = 0, b = len - 1; a < b; ++a, --b) {
            const int64_t x = chain_out[a];
            chain_out[a] = chain_out[b];
            chain_out[b] = x;
        }
        *chain_len = len;
        if (opt_score) *opt_score = opt_value;
        if (dp_out) memcpy(dp_out, dp, M * sizeof(float));
        if (backptr_out) memcpy(backptr_out, bp, M * sizeof(int64_t));
    }
done:
    if (gf)
        for (size_t g = 0; g < n_grp; ++g) clc_mst_free(&gf[g]);
    if (ortho)
        for (size_t k = 0; k < (size_t)(P ? 2 * P : 1) * npair; ++k) clc_omst_free(&ortho[k]);
    free(gf);
    free(ortho);
    free(grp_first);
    free(grp_shift);
    free(pair_grp);
    free(ent_grp);
    free(ent_node);
    free(ent_onode);
    free(ent);
    free(srt);
    free(dp);
    free(bp);
    return rc;
}
