/*
 * centrolign_b200.h -- C ABI of the B200-native (sm_100a) gap-fill path.
 *
 * This is the drop-in boundary for the Stitcher's inter-anchor gap fill of
 * jeizenga/centrolign.  The reference has no FFI layer: its boundary is the C++ template
 *
 *     Alignment po_poa<NumPW, Graph>(graph1, graph2, sources1, sources2, sinks1, sinks2,
 *                                    const AlignmentParameters<NumPW>&, int64_t* score_out)
 *         -- reference: include/centrolign/alignment.hpp:78-85 (body :753-1163)
 *
 * called once per inter-anchor window from Stitcher::do_alignment
 * (reference: include/centrolign/stitcher.hpp:295-303) inside the serial loop of
 * Stitcher::stitch (stitcher.hpp:157-203).  The windows of one stitch() call are
 * independent, so this ABI takes a whole BATCH of windows; the C++ wrapper that keeps the
 * reference signature is centrolign_b200/hostcpp/po_poa_b200.hpp, and INTEGRATION.md
 * shows the lines a maintainer adds to stitcher.hpp.
 *
 * Conventions
 *   - plain pointers + sizes, caller-owned host memory, nothing retained after return;
 *   - node ids are window-relative and in the CALLER'S order (no topological order
 *     required; the library renumbers internally);
 *   - predecessor lists are in the graph's previous() order (graph.hpp:111) and source /
 *     sink lists in the caller's order: the reference's traceback tie-breaking
 *     (alignment.hpp:979-990, 1048-1137) depends on both and is reproduced exactly;
 *   - results: per window the optimal score (== *score_out of the reference) and the
 *     alignment as (node_id1, node_id2) int32 pairs, CLB_GAP (-1) = AlignedPair::gap
 *     (alignment.hpp:34-51), in forward order;
 *   - return value 0 = success, otherwise a CLB_E* code; clb_last_error() gives text.
 *     There is no CPU fallback: without a usable CUDA device every compute entry point
 *     fails with CLB_ECUDA.
 */
#ifndef CENTROLIGN_B200_H
#define CENTROLIGN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLB_GAP (-1)
#define CLB_MAX_PW 3

enum {
    CLB_OK = 0,
    CLB_EINVAL = 1,  /* malformed arguments (bad sizes, ids out of range, NumPW not in 1..3) */
    CLB_ECYCLE = 2,  /* a graph is not acyclic (the reference asserts, topological_order.hpp:56) */
    CLB_ECUDA = 3,   /* CUDA runtime / launch failure, or no device */
    CLB_ENOMEM = 4,  /* host or device allocation failed (a window's workspace does not fit) */
    CLB_ESTATE = 5   /* call sequence error on a clb_batch handle */
};

/* AlignmentParameters<NumPW> (alignment.hpp:56-65); penalties are stored positive. */
typedef struct clb_params {
    int32_t num_pw; /* 1..3 */
    uint32_t match;
    uint32_t mismatch;
    uint32_t gap_open[CLB_MAX_PW];
    uint32_t gap_extend[CLB_MAX_PW];
} clb_params;

/* One side (graph1 or graph2) of every window of a batch, concatenated. */
typedef struct clb_graph_batch {
    const int64_t* node_off;  /* [n_windows+1] first node of window w in label[] */
    const uint8_t* label;     /* [node_off[n_windows]] node labels (graph.label(id)), compared for equality */
    const int64_t* edge_off;  /* [n_windows+1] first edge of window w in pred[] */
    const uint32_t* pred_off; /* [node_off[n_windows]+n_windows]; window w owns n_w+1 window-relative
                                 CSR offsets starting at pred_off[node_off[w]+w] */
    const uint32_t* pred;     /* [edge_off[n_windows]] predecessor ids, previous() order */
    const int64_t* src_off;   /* [n_windows+1] */
    const uint32_t* src;      /* sources, caller order */
    const int64_t* snk_off;   /* [n_windows+1] */
    const uint32_t* snk;      /* sinks, caller order */
} clb_graph_batch;

/*
 * One-shot: align every window of the batch on `device`.
 *   score_out [n_windows]                          optimal scores
 *   aln_off   [n_windows+1]  (input)               window w's pairs go to aln_pairs[2*aln_off[w] ...];
 *                                                  capacity aln_off[w+1]-aln_off[w] must be >= n1_w+n2_w
 *   aln_pairs [2*aln_off[n_windows]]               (id1,id2) pairs, forward order
 *   aln_len   [n_windows]                          number of pairs written per window
 * Replaces n_windows calls of po_poa (alignment.hpp:78-85).
 */
int clb_popoa_batch(int device, int32_t n_windows, const clb_graph_batch* g1, const clb_graph_batch* g2,
                    const clb_params* params, int64_t* score_out, const int64_t* aln_off, int32_t* aln_pairs,
                    uint32_t* aln_len);

/*
 * Staged form of the same call, for callers that keep a batch resident in HBM
 * (bench.py times clb_batch_run alone for the device-resident figure):
 *   create   : validate + renumber each graph topologically into pinned host staging
 *   upload   : host -> device copies, workspace allocation
 *   run      : all kernels (fill + traceback), synchronous; results stay on the device
 *   download : device -> host copy of scores / alignments, translated back to caller ids
 */
typedef struct clb_batch clb_batch;

int clb_batch_create(int device, int32_t n_windows, const clb_graph_batch* g1, const clb_graph_batch* g2,
                     const clb_params* params, clb_batch** out);
int clb_batch_upload(clb_batch* b);
int clb_batch_run(clb_batch* b);
int clb_batch_download(clb_batch* b, int64_t* score_out, const int64_t* aln_off, int32_t* aln_pairs,
                       uint32_t* aln_len);
void clb_batch_destroy(clb_batch* b);

/* Introspection for benchmarks (valid after clb_batch_run). */
typedef struct clb_batch_stats {
    double cells;            /* sum over windows of (n1+1)*(n2+1)  (stitcher.hpp:241) */
    double kernel_ms;        /* CUDA-event time of the last run on the batch's stream */
    double fill_ms;          /* of which: time inside the DP-fill kernel launches (0 if fused) */
    int64_t kernel_launches; /* kernels launched by the last run */
    int64_t h2d_bytes;       /* bytes clb_batch_upload copies */
    int64_t d2h_bytes;       /* bytes clb_batch_download copies */
    int64_t workspace_bytes; /* device workspace held by the batch */
    int64_t int_ops;         /* algorithmic INT32 add/max count of the fill (SURVEY.md 8d formula) */
    int64_t persist_bytes;   /* bytes of persisted rows/columns the fill writes to HBM (its algorithmic traffic) */
} clb_batch_stats;
int clb_batch_get_stats(const clb_batch* b, clb_batch_stats* out);

/*
 * Wavefront variant of the gap fill.  Replaces n_windows calls of
 *
 *     Alignment pwfa_po_poa<NumPW, Graph>(graph1, graph2, sources1, sources2, sinks1, sinks2,
 *                                         const AlignmentParameters<NumPW>&, int64_t prune_limit, int64_t* score_out)
 *         -- reference: include/centrolign/alignment.hpp:117-125 (body :2299-2338, search :1959-2033, :1712-1827),
 *            called from Stitcher::do_alignment (include/centrolign/stitcher.hpp:336-339) with
 *            prune_limit = 2 * wfa_pruning_dist.
 *
 * This algorithm walks the graphs FORWARD, so the two sides are given as SUCCESSOR lists: same struct
 * layout as clb_graph_batch, but pred_off / pred hold, per node, its successors in the graph's next()
 * order (include/centrolign/graph.hpp next()); that order and the caller's order of the source lists
 * decide ties (alignment.hpp:1788-1826) and are reproduced exactly: identical alignment, identical score.
 * Outputs as for clb_popoa_batch.  Preconditions are the reference's: acyclic graphs, some source reaches
 * some sink on both sides (else CLB_EINVAL where the reference would dereference an empty queue),
 * non-zero gap_open (the reference's gcd divides by it), prune_limit >= 0.
 */
typedef clb_graph_batch clb_succ_graph_batch;

typedef struct clb_pwfa_stats {
    double kernel_ms;        /* CUDA-event time of the search+traceback kernel(s) */
    int64_t kernel_launches; /* 1 + re-runs of windows whose tables had to be enlarged */
    int64_t retries;         /* enlargement rounds */
    int64_t states;          /* search states settled (first dequeues), all windows */
    int64_t dequeued;        /* queue entries dequeued */
    int64_t steps;           /* warp steps (each resolves up to 32 queue entries in order) */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    int64_t workspace_bytes; /* back-pointer tables + FIFOs held during the run */
} clb_pwfa_stats;

int clb_pwfa_batch(int device, int32_t n_windows, const clb_succ_graph_batch* g1, const clb_succ_graph_batch* g2,
                   const clb_params* params, int64_t prune_limit, int64_t* score_out, const int64_t* aln_off,
                   int32_t* aln_pairs, uint32_t* aln_len, clb_pwfa_stats* stats /* may be NULL */);

/* Measured INT32 issue-rate probe (a dependent-free add/max loop on every SM):
 * returns achieved 10^12 INT32 lane-ops per second on `device`, <0 on error. */
double clb_int32_peak_tops(int device, int use_dpx);

/* Staging buffers (pinned host, device) are cached across calls; this frees the cache. */
void clb_release_cached_memory(void);

const char* clb_last_error(void);
int clb_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CENTROLIGN_B200_H */
