// knn_grid.cu -- uniform-grid exact KNN (placeholder until the grid search lands:
// it declines every problem so the dispatcher uses the tiled scan).
#include "common.cuh"
#include "knn_common.cuh"
namespace ffb6d {
size_t knn_grid_workspace_bytes(int64_t, int64_t, int64_t, int) { return 0; }
int knn_grid_launch(const float *, const float *, int64_t, int64_t, int64_t, int, void *, int, void *,
                    size_t, cudaStream_t)
{
    set_error("knn grid search not available in this build");
    return FFB6D_ERR_INVALID;
}
}  // namespace ffb6d
