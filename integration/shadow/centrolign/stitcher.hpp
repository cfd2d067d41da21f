// integration/shadow/centrolign/stitcher.hpp -- zero-edit drop-in of the B200 gap fill into the reference.
//
// Put `integration/shadow` BEFORE the reference's include directory on the include path of the four
// translation units that see stitcher.hpp (stitcher.cpp, core.cpp, parameters.cpp, main.cpp).  Every
// `#include "centrolign/stitcher.hpp"` then lands here first.  This file reads the untouched reference
// header with #include_next, with two identifiers redirected while it is being read:
//   * `Stitcher` -> `StitcherReference`: the reference class, whole and unchanged, under another name;
//   * `po_poa`   -> `po_poa_b200`: the one use of po_poa in that header is the gap-fill call in
//     Stitcher::do_alignment (reference: include/centrolign/stitcher.hpp:297-299);
//   * `pwfa_po_poa` -> `pwfa_po_poa_b200`: likewise the wavefront route (stitcher.hpp:336-339).
// `centrolign::Stitcher` is then a thin subclass whose stitch() / internal_stitch() run the reference's
// own loop in recording mode and finish with one batched GPU call (centrolign_b200/hostcpp/
// stitch_recorder.hpp).  In stitcher.cpp itself (compiled with -DCLB_SHADOW_STITCHER_TU) the member
// definitions must land on the reference class and its translate() call is hooked, so this header ends by
// redirecting `Stitcher` and `translate` for the rest of that one translation unit.
#ifndef CENTROLIGN_B200_SHADOW_STITCHER_HPP
#define CENTROLIGN_B200_SHADOW_STITCHER_HPP

// everything the reference's stitcher.hpp includes, read first so the redirections cannot touch it
#include "centrolign/chain_merge.hpp"
#include "centrolign/alignment.hpp"
#include "centrolign/graph.hpp"
#include "centrolign/topological_order.hpp"
#include "centrolign/step_index.hpp"
#include "centrolign/anchorer.hpp"
#include "centrolign/subgraph_extraction.hpp"
#include "centrolign/logging.hpp"
#include "centrolign/partition_client.hpp"

#include <algorithm>

#include "stitch_recorder.hpp"

namespace centrolign {

// same argument list as centrolign::po_poa (include/centrolign/alignment.hpp:78-85)
template <int NumPW, class Graph>
Alignment po_poa_b200(const Graph& graph1, const Graph& graph2, const std::vector<uint64_t>& sources1,
                      const std::vector<uint64_t>& sources2, const std::vector<uint64_t>& sinks1,
                      const std::vector<uint64_t>& sinks2, const AlignmentParameters<NumPW>& params,
                      int64_t* score_out = nullptr) {
    auto& rec = centrolign_b200::StitchRecorder<Alignment>::instance();
    if (rec.active() && !score_out)
        return rec.record<NumPW>(graph1, graph2, sources1, sources2, sinks1, sinks2, params);
    return centrolign_b200::po_poa<NumPW, Graph, AlignmentParameters<NumPW>, Alignment>(
        graph1, graph2, sources1, sources2, sinks1, sinks2, params, score_out);
}

// same argument list as centrolign::pwfa_po_poa (include/centrolign/alignment.hpp:117-125)
template <int NumPW, class Graph>
Alignment pwfa_po_poa_b200(const Graph& graph1, const Graph& graph2, const std::vector<uint64_t>& sources1,
                           const std::vector<uint64_t>& sources2, const std::vector<uint64_t>& sinks1,
                           const std::vector<uint64_t>& sinks2, const AlignmentParameters<NumPW>& params,
                           int64_t prune_limit, int64_t* score_out = nullptr) {
    // The device search keeps WFA scores in 29 bits (pwfa_host.cu); a window whose worst-case score does not fit -- more than
    // ~200 k nodes with the production parameters, i.e. only with a user-raised max_wfa_size -- keeps the reference's 64-bit CPU
    // code instead of failing the whole batch.  Conservative bound: penalties before the division by their gcd (alignment.hpp:1959-2033).
    {
        uint64_t max_pen = 2ull * ((uint64_t)params.match + params.mismatch);
        for (int k = 0; k < NumPW; ++k)
            max_pen = std::max<uint64_t>(max_pen, 2ull * params.gap_open[k] + 2ull * params.gap_extend[k] + params.match);
        if ((uint64_t)(graph1.node_size() + graph2.node_size() + 2) * (max_pen + 1) >= (uint64_t(1) << 29))
            return pwfa_po_poa<NumPW>(graph1, graph2, sources1, sources2, sinks1, sinks2, params, prune_limit, score_out);
    }
    auto& rec = centrolign_b200::StitchRecorder<Alignment>::instance();
    if (rec.active() && !score_out)
        return rec.record_pwfa<NumPW>(graph1, graph2, sources1, sources2, sinks1, sinks2, params, prune_limit);
    return centrolign_b200::pwfa_po_poa<NumPW, Graph, AlignmentParameters<NumPW>, Alignment>(
        graph1, graph2, sources1, sources2, sinks1, sinks2, params, prune_limit, score_out);
}

// hook for the translate() call in Stitcher::subalign (src/stitcher.cpp:66)
inline void translate_b200(Alignment& alignment, const std::vector<uint64_t>& back_translation1,
                           const std::vector<uint64_t>& back_translation2) {
    if (centrolign_b200::StitchRecorder<Alignment>::instance().note_translation(alignment, back_translation1, back_translation2))
        return;  // a marker: translated when the batch result is spliced in
    translate(alignment, back_translation1, back_translation2);
}

}  // namespace centrolign

#define Stitcher StitcherReference
#define po_poa po_poa_b200
#define pwfa_po_poa pwfa_po_poa_b200
#include_next "centrolign/stitcher.hpp"
#undef pwfa_po_poa
#undef po_poa
#undef Stitcher

namespace centrolign {

class Stitcher : public StitcherReference {
public:
    // reference signature, include/centrolign/stitcher.hpp:34-38
    template <class BGraph1, class BGraph2, class XMerge1, class XMerge2>
    Alignment stitch(const std::vector<std::vector<anchor_t>>& anchor_segments, const BGraph1& graph1,
                     const BGraph2& graph2, const SentinelTableau& tableau1, const SentinelTableau& tableau2,
                     XMerge1& chain_merge1, XMerge2& chain_merge2) const {
        auto& rec = centrolign_b200::StitchRecorder<Alignment>::instance();
        centrolign_b200::StitchRecorder<Alignment>::Scope recording(rec);  // switched off again if anything below throws
        Alignment with_markers = StitcherReference::stitch(anchor_segments, graph1, graph2, tableau1, tableau2,
                                                           chain_merge1, chain_merge2);
        return rec.finish(std::move(with_markers));
    }
    // reference signature, include/centrolign/stitcher.hpp:41-43
    template <class BGraph, class XMerge>
    Alignment internal_stitch(const std::vector<anchor_t>& anchors, const BGraph& graph, const XMerge& xmerge) const {
        auto& rec = centrolign_b200::StitchRecorder<Alignment>::instance();
        centrolign_b200::StitchRecorder<Alignment>::Scope recording(rec);
        Alignment with_markers = StitcherReference::internal_stitch(anchors, graph, xmerge);
        return rec.finish(std::move(with_markers));
    }
};

}  // namespace centrolign

#ifdef CLB_SHADOW_STITCHER_TU
// once per program (this one translation unit): start creating the CUDA context while the CLI reads its input
namespace {
const int clb_context_started_at_load = (clb_warm_up(centrolign_b200::chain_device()), 0);
}
// the rest of this translation unit is src/stitcher.cpp: its member definitions belong to the reference class
#define Stitcher StitcherReference
#define translate translate_b200
#endif

#endif
