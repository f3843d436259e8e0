// Instantiates the search kernels for one distance mode (DistReg<1>); see search.cuh / launch.cuh.
#include "launch.cuh"
HNSW_DEFINE_KIND(r1, DistReg<1>)
