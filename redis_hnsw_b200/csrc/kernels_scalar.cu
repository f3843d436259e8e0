// Instantiates the search kernels for one distance mode (DistScalar); see search.cuh / launch.cuh.
#include "launch.cuh"
HNSW_DEFINE_KIND(scalar, DistScalar)
