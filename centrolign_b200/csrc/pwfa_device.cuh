// pwfa_device.cuh -- data layout of the wavefront-variant gap fill (pwfa_po_poa) shared by
// pwfa_host.cu and pwfa_kernels.cu.
//
// Reference: include/centrolign/alignment.hpp:2299-2338 (pwfa_po_poa), :1959-2033
// (pwfa_po_poa_internal), :1712-1827 (wfa_iteration<true>).  The search is Dijkstra over states
// (node1, node2, component) on a bucket queue with one FIFO per integer score; node id n (node_size)
// is the virtual start whose successors are the sources.  Node ids stay in the CALLER'S numbering
// (no renumbering is needed: the search only follows next() lists).
#pragma once
#include <stdint.h>

namespace clb {

constexpr int kPwfaMaxClass = 8;            // distinct penalties: mismatch, 3 x (open+extend), 3 x extend, and 0
constexpr uint32_t kPwfaNoParent = 0xffffffffu;
constexpr int kPwfaNegInf = -(1 << 30);     // "nothing reached yet" for the furthest-distance bound

// per node (index n = virtual start), one int4:
//   x = minmax_distance(...).first   (fewest nodes from a source;  -1 for the virtual start)
//   y = minmax_distance(...).second  (most nodes from a source;    -1 for the virtual start)
//   z = offset of its successor list in next[] / nlab[] (window-relative)
//   w = label (8 bits) | can-reach-a-sink << 8 | is-sink << 9 | out-degree << 10
constexpr uint32_t kPwfaReach = 1u << 8;
constexpr uint32_t kPwfaSink = 1u << 9;
constexpr uint32_t kPwfaDegShift = 10;

struct PwfaWindow {
    int64_t info1, info2;  // first int4 of this window's node records per side (n+1 records each)
    int64_t next1, next2;  // first entry of this window's successor arrays (real edges, then the sources)
    int64_t out;           // first (id1,id2) pair slot of this window in the output
    uint32_t n1, n2;
    uint32_t hash_log2;    // back-pointer table slots (16 B each)
    uint32_t fifo_log2;    // entries per penalty-class FIFO (16 B each)
};

struct PwfaParams {
    int num_pw;
    int n_class;                    // distinct penalties; class index ascending = penalty descending; last class = 0
    uint32_t pen[kPwfaMaxClass];
    int cls_mismatch;
    int cls_open[3];                // class of open_k + extend_k
    int cls_ext[3];
    uint32_t match;                 // original match score
    uint32_t factor;                // gcd the WFA parameters were divided by (to_wfa_params, alignment.hpp:1644-1651)
    int prune_limit;                // clamped to [0, 2^30]
};

struct PwfaArgs {
    const int4* info1;
    const int4* info2;
    const uint32_t* next1;
    const uint32_t* next2;
    const uint8_t* nlab1;  // label of next1[k]
    const uint8_t* nlab2;
    const PwfaWindow* win;
    const int32_t* order;  // window ids to run, most work first
    int32_t n_run;
    int32_t* queue;        // atomic work counter
    char* workspace;       // gridDim.x slots
    int64_t slot_bytes;
    int64_t* score;        // [n_windows]
    int32_t* status;       // [n_windows] 0 ok, else kPwfa* below
    uint32_t* aln_len;     // [n_windows]
    int32_t* aln;          // pairs, written backwards from the end of each window's region
    int64_t* wstats;       // [n_windows][4]: states settled, entries dequeued, final WFA score, chunks
    PwfaParams prm;
};

enum { kPwfaOk = 0, kPwfaHashFull = 1, kPwfaFifoFull = 2, kPwfaQueueDry = 3, kPwfaInternal = 4 };

}  // namespace clb
