// popoa_device.cuh -- data layout shared by the host flattening code and the sm_100a kernels.
//
// Matrix coordinates.  A window's DP matrix has rows 0..n1 and columns 0..n2.  Row / column 0
// is the reference's "boundary" row / column (the extra final row / column of its table,
// include/centrolign/alignment.hpp:788-790, moved to the front); row i >= 1 is the graph-1 node
// with topological rank i-1, column j >= 1 likewise for graph 2.  Every predecessor therefore
// has a smaller index, sources carry the extra predecessor 0 appended LAST in their list
// (alignment.hpp:1078-1084), and (n1+1)*(n2+1) is the cell count of the GCUPS metric.
#pragma once
#include <stdint.h>

namespace clb {

constexpr int kMinInf = INT32_MIN / 2;  // cell_t::mininf, alignment.hpp:740
constexpr int kStrip = 32;              // columns per strip = lanes per warp
constexpr int kRowBlock = 64;           // rows per traceback tile
constexpr int kNear = 3;                // predecessors at most this far back are served from the shared-memory ring
constexpr int kRingRows = 8;            // fill-kernel ring depth (>= 2*kNear+2)
constexpr int kPanelRowsMax = 2048;     // panel height of tiled windows (default; a multiple of kRowBlock); in "auto" mode about
constexpr int kPanelRowsMin = 512;      // n1/8 between these bounds (host and kernel call panel_rows_for); CLB_PANEL_ROWS, 0 = off

// Windows whose whole matrix has at most kSmallCells cells go to popoa_small_kernel (one warp per window, the matrix in
// shared memory: 2 x 16 B per cell, kSmallWarps warps per CTA = 192 KB); the reference's own Stitcher windows are
// almost all of this kind (median 9 cells, p99 169).
constexpr int kSmallCells = 384;
constexpr int kSmallWarps = 16;

// per-node info word
constexpr uint32_t kInfoLabelMask = 0xffu;
constexpr uint32_t kInfoRegular = 1u << 8;  // exactly one predecessor and it is index-1
constexpr uint32_t kInfoPersist = 1u << 9;  // row / column is kept in the window workspace
constexpr uint32_t kInfoNearShift = 10;     // bits 10..12: predecessor index-d present, d = 1..kNear (index-d >= 1)
constexpr uint32_t kInfoNearMask = 7u << kInfoNearShift;
constexpr uint32_t kInfoFar = 1u << 13;     // has a predecessor not covered by the near bits (far back, or the boundary 0)
constexpr uint32_t kInfoSlotShift = 14;     // bits 14..31: workspace slot of a persisted row / column ...
constexpr uint32_t kInfoSlotEscape = 0x3ffffu;  // ... or this value: look the slot up in the slot array

// One side (graph 1 = rows, graph 2 = columns) of all windows, concatenated on the device.
struct SideArrays {
    const uint32_t* info;   // per window n+1 entries (index 0 = boundary): label | flags
    const int32_t* slot;    // per window n+1 entries: workspace slot of a persisted row/column, else -1
    const uint32_t* depth;  // per window n+1 entries: fewest nodes on a path from a source (0 = unreachable)
    const uint32_t* poff;   // per window n+2 entries: predecessor list offsets (window-relative)
    const uint32_t* pidx;   // predecessor matrix indices, previous() order, 0 appended last for sources
    const uint32_t* sinks;  // sink matrix indices, caller order
};

struct WindowMeta {
    uint32_t n1, n2;
    uint32_t nsnk1, nsnk2;
    uint32_t nrslot, ncslot;  // persisted rows / columns
    int64_t node1, node2;     // base of info/slot/depth entries
    int64_t poff1, poff2;     // base of poff entries
    int64_t pidx1, pidx2;     // base of pidx entries
    int64_t snk1, snk2;       // base of sinks entries
    int64_t out;              // first (id1,id2) pair slot of this window in the output
};

struct Params {
    int match, mismatch;  // mismatch stored positive
    int oe[3];            // open + extend
    int e[3];
};

struct LaunchArgs {
    SideArrays s1, s2;
    const WindowMeta* meta;
    const int32_t* order;   // window ids, largest first
    int32_t n_windows;
    int32_t* queue;         // atomic work counter
    char* workspace;        // gridDim.x slots
    int64_t slot_bytes;
    int64_t* score;         // [n_windows]
    int32_t* aln;           // pairs, written backwards from the end of each window's region
    uint32_t* aln_len;      // [n_windows]
    Params prm;
    int debug_flags;        // bit1: generic fill step everywhere; bit0 (-DCLB_PROFILE builds only): skip the traceback walk, results invalid
    int start_lag;          // rows a strip stays behind its left neighbour when it starts
    int slot_by_smid;       // 1: workspace slot pair chosen by %smid (kernels of several chunks share one workspace)
    int panel_rows;         // panel_rows_for(n1, this): wide windows with more rows are filled as (panel, strip) tiles; rows p*H-2..p*H are persisted
};

// Panel height of a window with n1 rows; cfg < 0: automatic, 0: no tiling, > 0: fixed.
#ifdef __CUDACC__
__host__ __device__
#endif
inline int panel_rows_for(int n1, int cfg) {
    if (cfg >= 0) return cfg;
    int h = (n1 / 8 + kRowBlock - 1) / kRowBlock * kRowBlock;
    return h < kPanelRowsMin ? kPanelRowsMin : (h > kPanelRowsMax ? kPanelRowsMax : h);
}

// Workspace of one window: rowbuf {M,H_k} + colbuf {M,D_k} + boundary row + boundary column (16 B per
// entry), then coleff (4 B per persisted-column entry: the column's effective diagonal input per row).
// A persisted ROW holds, per column, M and what the cell offers the rows below it: H_k = max(I_k - e_k, M - oe_k)
// (the wide fill stores it with one 16-byte store per column; everything the traceback needs from a row outside
// the recomputed tile follows from it, see popoa_kernels.cu "Traceback").  Rows are padded by kRowPad entries so that
// a lane of the wide fill may store all of its four columns when the last ones lie beyond n2.
constexpr uint32_t kRowPad = 3;
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t row_stride(uint32_t n2) { return n2 + 1u + kRowPad; }
inline int64_t workspace_int4(uint32_t n1, uint32_t n2, uint32_t nrslot, uint32_t ncslot) {
    return (int64_t)nrslot * row_stride(n2) + (int64_t)ncslot * (n1 + 1) + (n2 + 1) + (n1 + 1);
}
inline int64_t workspace_bytes(uint32_t n1, uint32_t n2, uint32_t nrslot, uint32_t ncslot) {
    return 16 * workspace_int4(n1, n2, nrslot, ncslot) + ((4 * (int64_t)ncslot * (n1 + 1) + 15) & ~int64_t(15));
}

}  // namespace clb
