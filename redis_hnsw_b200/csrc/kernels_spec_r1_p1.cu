// Instantiates the SPEC builder's K1 (spec.cuh) for rows of 32 * 1 floats, one group of list classes; see spec_launch.cuh.
#include "spec_launch.cuh"
HNSW_DEFINE_SPEC_KIND_PART(r1, 1, 1)
