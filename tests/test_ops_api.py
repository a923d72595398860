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


def test_evaluation_front_end_and_icp_argument_errors():
    """cloudaae_b200/evaluate_cloudAAE_ycbv.py: shape / dtype violations first, then 'no CPU kernel'."""
    from cloudaae_b200 import evaluate_cloudAAE_ycbv as EV
    for name in ("get_pointcloud", "get_outlier_idx", "FPS_random", "icp_refine", "SegmentFrontEnd", "pose_to_matrix"):
        assert hasattr(EV, name)
    depth = torch.zeros(2, 4, 6, dtype=torch.int16)
    label = torch.zeros(2, 4, 6, dtype=torch.uint8)
    intr = torch.ones(2, 5)
    thr = torch.full((21,), 0.2)
    with pytest.raises(InvalidArgumentError, match="frames, height, width"):
        EV.SegmentFrontEnd(depth[0], label[0], intr, thr)
    with pytest.raises(InvalidArgumentError, match="16-bit depth"):
        EV.SegmentFrontEnd(depth.float(), label, intr, thr)
    with pytest.raises(InvalidArgumentError, match=r"\(frames, 5\)"):
        EV.SegmentFrontEnd(depth, label, intr[:, :4], thr)
    with pytest.raises(NotImplementedError, match="CUDA"):
        EV.SegmentFrontEnd(depth, label, intr, thr)
    with pytest.raises(InvalidArgumentError, match="height, width"):
        EV.get_pointcloud(depth, 1.0, 1.0, 0.0, 0.0, 1.0)
    with pytest.raises(NotImplementedError):
        EV.get_outlier_idx(torch.zeros(10, 3), 100, 0.02, 0.5)
    with pytest.raises(ValueError, match="empty segment"):
        EV.FPS_random(torch.zeros(0, 3), 4)
    with pytest.raises(NotImplementedError):
        EV.FPS_random(torch.zeros(5, 3), 4, first_idx=0)
    src, tgt, init = torch.zeros(2, 16, 3), torch.zeros(2, 8, 3), torch.zeros(2, 4, 4, dtype=torch.float64)
    with pytest.raises(InvalidArgumentError, match="float64 init"):
        EV.icp_refine(src, tgt, init.float())
    with pytest.raises(InvalidArgumentError, match="source_of_seg"):
        EV.icp_refine(src[:1], tgt, init)
    with pytest.raises(NotImplementedError):
        EV.icp_refine(src, tgt, init)
