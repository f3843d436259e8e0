// Entry point of one distance mode (DistReg<4>): its list classes are compiled in kernels_r4_p{1,2,3}.cu.
#include "launch.cuh"
HNSW_DECLARE_KIND_PARTS(r4)
