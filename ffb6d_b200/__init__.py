"""ffb6d_b200 -- B200 (sm_100a) implementation of FFB6D's bidirectional-fusion hot path.

Public surface (same names and argument meaning as the reference, ethnhe/FFB6D):

* :func:`knn_search` -- ``DataProcessing.knn_search`` (models/RandLA/helper_tool.py:160-170)
* :func:`random_sample`, :func:`nearest_interpolation` -- ``FFB6D.random_sample`` /
  ``FFB6D.nearest_interpolation`` (models/ffb6d.py:159-194) and the RandLA twins
  (models/RandLA/RandLANet.py:87-117)
* :func:`gather_neighbour`, :func:`relative_pos_encoding` -- ``Building_block`` helpers
  (models/RandLA/RandLANet.py:216-234)
* :func:`grid_sub_sampling` -- ``DataProcessing.grid_sub_sampling`` (helper_tool.py:199-219)
* :func:`build_ffb6d_indices` -- the 22-call KNN schedule of the datasets
  (datasets/ycb/ycb_dataset.py:269-309) run on the GPU in one go.

Everything runs through libffb6d_b200.so (hand-written CUDA behind a C ABI, see
include/ffb6d_b200.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises if the CUDA library is not built)
from .ops import (knn_search, random_sample, nearest_interpolation, gather_neighbour,  # noqa: F401
                  relative_pos_encoding, choose_gather, grid_sub_sampling, KnnGrid, backproject, fusion_mlp, fusion_mlp_pack, PackedWeight, fold_batchnorm,
                  att_pool)
from . import randla  # noqa: F401
from .schedule import (build_ffb6d_indices, build_ffb6d_indices_from_depth, build_ffb6d_indices_native, knn_schedule,  # noqa: F401
                       gather_schedule)
from .helper_tool import DataProcessing  # noqa: F401

__all__ = [
    "knn_search", "random_sample", "nearest_interpolation", "gather_neighbour",
    "relative_pos_encoding", "choose_gather", "grid_sub_sampling", "KnnGrid", "backproject", "fusion_mlp", "fusion_mlp_pack", "PackedWeight", "fold_batchnorm", "att_pool", "randla", "build_ffb6d_indices", "build_ffb6d_indices_from_depth", "build_ffb6d_indices_native",
    "knn_schedule", "gather_schedule", "DataProcessing",
]
