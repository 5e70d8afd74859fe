// grid_subsample.cu -- voxel-grid barycentre subsampling (placeholder).
#include "common.cuh"
extern "C" int ffb6d_grid_subsample_host(const float *, size_t, const float *, size_t, const int *,
                                         size_t, float, float *, float *, int *, size_t *)
{
    ffb6d::set_error("grid_subsample: not implemented in this build");
    return FFB6D_ERR_INVALID;
}
