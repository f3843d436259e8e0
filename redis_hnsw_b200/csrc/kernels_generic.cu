// Instantiates the search kernels for one distance mode (DistGeneric); see search.cuh / launch.cuh.
#include "launch.cuh"
HNSW_DEFINE_KIND(generic, DistGeneric)
