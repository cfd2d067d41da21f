// integration/shadow/centrolign/stitcher.hpp -- zero-edit drop-in of the B200 gap fill into the reference.
//
// Put `integration/shadow` BEFORE the reference's include directory on the compiler's include path.
// Every `#include "centrolign/stitcher.hpp"` then lands here first; this file pulls in the untouched
// reference header with #include_next, but while that header is being read the identifier `po_poa`
// is redirected to `po_poa_b200` below.  The only use of `po_poa` in that header is the gap-fill call
// in Stitcher::do_alignment (reference: include/centrolign/stitcher.hpp:297-299), so that one call --
// and nothing else in the reference -- goes to the GPU.  alignment.hpp is included beforehand, so the
// reference's own po_poa definition is not renamed.
#ifndef CENTROLIGN_B200_SHADOW_STITCHER_HPP
#define CENTROLIGN_B200_SHADOW_STITCHER_HPP

#include "centrolign/alignment.hpp"
#include "centrolign/graph.hpp"
#include "po_poa_b200.hpp"

namespace centrolign {

// same argument list as centrolign::po_poa (include/centrolign/alignment.hpp:78-85)
template <int NumPW, class Graph>
Alignment po_poa_b200(const Graph& graph1, const Graph& graph2, const std::vector<uint64_t>& sources1,
                      const std::vector<uint64_t>& sources2, const std::vector<uint64_t>& sinks1,
                      const std::vector<uint64_t>& sinks2, const AlignmentParameters<NumPW>& params,
                      int64_t* score_out = nullptr) {
    return centrolign_b200::po_poa<NumPW, Graph, AlignmentParameters<NumPW>, Alignment>(
        graph1, graph2, sources1, sources2, sinks1, sinks2, params, score_out);
}

}  // namespace centrolign

// headers the reference's stitcher.hpp includes, read now so the redirection below cannot touch them
#include "centrolign/chain_merge.hpp"
#include "centrolign/topological_order.hpp"
#include "centrolign/step_index.hpp"
#include "centrolign/anchorer.hpp"
#include "centrolign/subgraph_extraction.hpp"
#include "centrolign/logging.hpp"
#include "centrolign/partition_client.hpp"

#define po_poa po_poa_b200
#include_next "centrolign/stitcher.hpp"
#undef po_poa

#endif
