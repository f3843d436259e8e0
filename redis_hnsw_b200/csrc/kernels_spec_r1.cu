// Entry point of the SPEC builder's K1 for rows of 32 * 1 floats: classes compiled in kernels_spec_r1_p{1,2}.cu.
#include "spec_launch.cuh"
HNSW_DECLARE_SPEC_KIND_PARTS(r1)
