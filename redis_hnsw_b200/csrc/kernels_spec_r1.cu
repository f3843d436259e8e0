// Instantiates the SPEC builder's K1 (spec.cuh) for rows of 32 * 1 floats; see spec_launch.cuh.
#include "spec_launch.cuh"
HNSW_DEFINE_SPEC_KIND(r1, 1)
