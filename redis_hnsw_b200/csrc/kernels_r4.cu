// Instantiates the search kernels for one distance mode (DistReg<4>); see search.cuh / launch.cuh.
#include "launch.cuh"
HNSW_DEFINE_KIND(r4, DistReg<4>)
