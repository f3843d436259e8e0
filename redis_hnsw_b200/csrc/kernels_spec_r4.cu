// Entry point of the SPEC builder's K1 for rows of 32 * 4 floats: classes compiled in kernels_spec_r4_p{1,2}.cu.
#include "spec_launch.cuh"
HNSW_DECLARE_SPEC_KIND_PARTS(r4)
