"""Host-side operator API: same names / argument order / error behaviour as the reference's
tf_sampling.py and tf_nndistance.py; no CPU kernels."""
import pytest
import torch

from cloudaae_b200 import InvalidArgumentError
from cloudaae_b200.tf_ops.nn_distance import tf_nndistance
from cloudaae_b200.tf_ops.sampling import tf_sampling


def test_public_names():
    for name in ("nn_distance",):
        assert callable(getattr(tf_nndistance, name))
    for name in ("prob_sample", "gather_point", "farthest_point_sample"):
        assert callable(getattr(tf_sampling, name))
    import cloudaae_b200
    assert cloudaae_b200.nn_distance is tf_nndistance.nn_distance
    assert cloudaae_b200.farthest_point_sample is tf_sampling.farthest_point_sample


def test_nn_distance_shape_errors_use_reference_messages():
    a = torch.zeros(2, 4, 3)
    with pytest.raises(InvalidArgumentError, match=r"NnDistance requires xyz1 be of shape \(batch,#points,3\)"):
        tf_nndistance.nn_distance(a[0], a)
    with pytest.raises(InvalidArgumentError, match="NnDistance only accepts 3d point set xyz1"):
        tf_nndistance.nn_distance(a[..., :2], a)
    with pytest.raises(InvalidArgumentError, match="NnDistance only accepts 3d point set xyz2"):
        tf_nndistance.nn_distance(a, a[..., :2])
    with pytest.raises(InvalidArgumentError, match="same batch size"):
        tf_nndistance.nn_distance(a, a[:1])
    with pytest.raises(InvalidArgumentError, match="float32"):
        tf_nndistance.nn_distance(a.double(), a.double())
    with pytest.raises(InvalidArgumentError, match="NnDistanceGrad requires idx1 be of shape"):
        tf_nndistance.nn_distance_grad(a, a, torch.zeros(2, 4), torch.zeros(2, 3, dtype=torch.int32),
                                       torch.zeros(2, 4), torch.zeros(2, 4, dtype=torch.int32))


def test_sampling_shape_errors():
    a = torch.zeros(2, 8, 3)
    with pytest.raises(InvalidArgumentError, match="positive npoint"):
        tf_sampling.farthest_point_sample(0, a)
    with pytest.raises(InvalidArgumentError, match=r"\(batch_size,num_points,3\) inp shape"):
        tf_sampling.farthest_point_sample(4, a[..., :2])
    with pytest.raises(InvalidArgumentError, match=r"GatherPoint expects \(batch_size,num_result\) idx shape"):
        tf_sampling.gather_point(a, torch.zeros(3, 4, dtype=torch.int32))
    with pytest.raises(InvalidArgumentError, match="int32 idx"):
        tf_sampling.gather_point(a, torch.zeros(2, 4, dtype=torch.int64))
    with pytest.raises(InvalidArgumentError, match="out_g shape"):
        tf_sampling.gather_point_grad(a, torch.zeros(2, 4, dtype=torch.int32), torch.zeros(2, 5, 3))


def test_no_cpu_kernels():
    a = torch.zeros(2, 8, 3)
    with pytest.raises(NotImplementedError, match="no CPU kernel"):
        tf_nndistance.nn_distance(a, a)
    with pytest.raises(NotImplementedError, match="no CPU kernel"):
        tf_sampling.farthest_point_sample(4, a)
    with pytest.raises(NotImplementedError, match="no CPU kernel"):
        tf_sampling.gather_point(a, torch.zeros(2, 4, dtype=torch.int32))
