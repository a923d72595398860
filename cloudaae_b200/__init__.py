"""cloudaae_b200 — B200-native (sm_100a) hot path of GeeeG/CloudAAE behind the reference's operator API.

Operator API (same names and argument order as the reference's tf_ops):
    farthest_point_sample(npoint, inp), gather_point(inp, idx), prob_sample(inp, inpr)
    nn_distance(xyz1, xyz2) -> (dist1, idx1, dist2, idx2)
"""
from ._capi import CloudAAENativeError, InvalidArgumentError  # noqa: F401
from .tf_ops.nn_distance.tf_nndistance import nn_distance, nn_distance_grad  # noqa: F401
from .tf_ops.sampling.tf_sampling import (farthest_point_sample, farthest_point_sample_gather,  # noqa: F401
                                          gather_point, gather_point_grad, prob_sample)

__version__ = "0.1.0"
