// ref_glue_mvs_ada.cpp — the reference's MultiViewStereo with AdaptiveWeight as its support-weight
// functor (BASELINE configs[1]: "adaptive-weight aggregation").  The reference selects the functor
// with a file-scope typedef (`typedef GeodesicWeight WeightFunc;`, multiviewstereo.cpp:107-109; the
// author switched it by editing that line).  Here the same switch is made from outside: both
// functor headers are included under their own names first, then the NAME GeodesicWeight is mapped
// to AdaptiveWeight for the rest of this translation unit, so the reference's typedef reads
// `typedef AdaptiveWeight WeightFunc;`.  Everything the unit defines is renamed so that it can be
// linked next to the GeodesicWeight build of the same file (ref_glue_mvs.cpp).  TEST INFRASTRUCTURE;
// no reference code here.
#include "stereo/adaptiveweight.hpp"
#include "stereo/geodesicweight.hpp"
#define GeodesicWeight AdaptiveWeight
#define MultiViewStereo MultiViewStereoAdaptive
#define CostFunction CostFunctionAdaptive
#define outputPLYFile outputPLYFileAdaptive
#define cost_ncc cost_ncc_adaptive
#define REF_MVS_ADAPTIVE_TU 1
#define ref_mvs ref_mvsa
#define ref_mvs_create ref_mvsa_create
#define ref_mvs_destroy ref_mvsa_destroy
#define ref_mvs_num_views ref_mvsa_num_views
#define ref_mvs_run ref_mvsa_run
#define ref_mvs_initial_estimate ref_mvsa_initial_estimate
#define ref_mvs_set_neighbours ref_mvsa_set_neighbours
#define ref_mvs_mask_rows ref_mvsa_mask_rows
#define ref_mvs_num_threads ref_mvsa_num_threads
#define ref_mvs_cost_ncc ref_mvsa_cost_ncc
#define ref_mvs_curve ref_mvsa_curve
#include "ref_glue_mvs.cpp"
