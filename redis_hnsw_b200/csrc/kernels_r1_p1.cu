// Instantiates the kernels of one distance mode (DistReg<1>) for one group of list classes; see launch.cuh (run_kind).
#include "launch.cuh"
HNSW_DEFINE_KIND_PART(r1, DistReg<1>, 1)
