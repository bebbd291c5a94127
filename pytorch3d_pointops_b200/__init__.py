"""pytorch3d_pointops_b200 -- Blackwell-native (sm_100a) batched nearest-neighbour ops behind the
unchanged python API of mikel-zhobro/pytorch3d_pointops.

    from pytorch3d_pointops_b200.functions import knn_points, ball_query, sample_farthest_points
    from pytorch3d_pointops_b200.functions.chamfer import chamfer_distance
    from pytorch3d_pointops_b200.structures import Pointclouds

`import pytorch3d_pointops` resolves to the same modules through the alias package at the
repository root, so existing user code keeps working unchanged.
"""
__version__ = "0.7.8"  # the reference package version this API mirrors (pytorch3d_pointops/__init__.py:7)
