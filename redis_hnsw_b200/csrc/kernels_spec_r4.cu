// Instantiates the SPEC builder's K1 (spec.cuh) for rows of 32 * 4 floats; see spec_launch.cuh.
#include "spec_launch.cuh"
HNSW_DEFINE_SPEC_KIND(r4, 4)
