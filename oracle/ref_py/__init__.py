"""Executes the reference's OWN Python model / loss code in place — TEST INFRASTRUCTURE ONLY.

TensorFlow 1.12 cannot be installed here, but /root/reference/models/pointnet_ycb_23_decoder_4.py, utils/tf_util.py and
losses/*.py are plain Python over a small TF API surface.  `load()` puts an eager TensorFlow stand-in over torch
(oracle/ref_py/shim/tensorflow) and a brute-force tf_nndistance on sys.path and imports those files FROM WHERE THEY LIE
(nothing is copied); tests/test_ref_py_pins_model_oracle.py then requires oracle/model_ref.py to reproduce their
outputs and gradients, and commits the resulting vectors under tests/golden/ for the GPU box (where /root/reference
does not exist).
"""
import contextlib
import importlib
import io
import os
import sys

REF_ROOT = "/root/reference"
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "pointnet_ycb_23_decoder_4.py"))


def load():
    """Returns (tf stand-in, reference model module, tf_util, {chamfer_loss, trans_distance, angular_distance_taylor})."""
    if not available():
        raise FileNotFoundError(f"{REF_ROOT} is not present")
    saved_path, saved_tf = list(sys.path), sys.modules.get("tensorflow")
    for name in ("tensorflow", "tf_util", "tf_nndistance", "pointnet_ycb_23_decoder_4", "chamfer_loss", "trans_distance",
                 "angular_distance_taylor"):
        sys.modules.pop(name, None)
    sys.path[:0] = [_SHIM, os.path.join(REF_ROOT, "utils"), os.path.join(REF_ROOT, "models"), os.path.join(REF_ROOT, "losses")]
    try:
        tf = importlib.import_module("tensorflow")
        tf_util = importlib.import_module("tf_util")
        model = importlib.import_module("pointnet_ycb_23_decoder_4")
        losses = {n: importlib.import_module(n) for n in ("chamfer_loss", "trans_distance", "angular_distance_taylor")}
    finally:
        sys.path[:] = saved_path
        if saved_tf is not None:
            sys.modules["tensorflow"] = saved_tf
    return tf, model, tf_util, losses


@contextlib.contextmanager
def quiet():
    """The reference's model functions print every layer."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
