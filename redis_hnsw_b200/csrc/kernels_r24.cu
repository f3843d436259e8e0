// Instantiates the search kernels for one distance mode (DistReg<24>); see search.cuh / launch.cuh.
#include "launch.cuh"
HNSW_DEFINE_KIND(r24, DistReg<24>)
