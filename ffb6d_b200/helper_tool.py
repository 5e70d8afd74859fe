"""``DataProcessing`` with the reference's method names (models/RandLA/helper_tool.py:
160-170, 199-219), so ``from helper_tool import DataProcessing as DP`` call sites
(datasets/ycb/ycb_dataset.py:22, 275-308; RandLA/test_model.py:55-58) keep working with
``from ffb6d_b200.helper_tool import DataProcessing as DP``.  Only the two methods on the
fusion hot path exist here; the S3DIS / SemanticKITTI helpers are out of scope.
"""
from . import ops


class DataProcessing:
    @staticmethod
    def knn_search(support_pts, query_pts, k):
        """:param support_pts: points you have, B*N1*3
        :param query_pts: points you want to know the neighbour index, B*N2*3
        :param k: Number of neighbours in knn search
        :return: neighbor_idx: neighboring points indexes, B*N2*k (int32)"""
        return ops.knn_search(support_pts, query_pts, k)

    @staticmethod
    def grid_sub_sampling(points, features=None, labels=None, grid_size=0.1, verbose=0):
        """:param points: (N, 3) matrix of input points
        :param features: optional (N, d) matrix of features (floating number)
        :param labels: optional (N,) matrix of integer labels
        :param grid_size: parameter defining the size of grid voxels
        :param verbose: 1 to display
        :return: sub_sampled points, with features and/or labels depending of the input"""
        return ops.grid_sub_sampling(points, features=features, labels=labels,
                                     grid_size=grid_size, verbose=verbose)
