"""ctypes binding of libcloudaae_b200.so (the C ABI declared in include/cloudaae_b200.h).

There is no CPU fallback: if the library is missing or an entry point reports an error the call
raises.  Tensors cross the boundary as raw device pointers plus the current CUDA stream handle.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcloudaae_b200.so")

ABI_VERSION = 4   # 3: split-precision forward GEMMs (caae_gemm_tf32x3, caae_split_tf32, caae_edge_apply out_lo); 4: caae_edge_apply* record (pos_cnt, pos_sum), caae_edge_bwd_stats

_int = ctypes.c_int
_ptr = ctypes.c_void_p


class InvalidArgumentError(ValueError):
    """Shape/dtype violation — the role of tf.errors.InvalidArgumentError raised by OP_REQUIRES."""


class CloudAAENativeError(RuntimeError):
    """A C-ABI entry point returned a non-zero status."""


# name -> argument codes (i = int, l = long, f = float, d = double, p = pointer / stream handle).
# Every entry point returns an int status unless listed in _SPECIAL.
_CODES = {"i": _int, "l": ctypes.c_long, "f": ctypes.c_float, "d": ctypes.c_double, "p": _ptr,
          "Q": ctypes.c_ulonglong}
_SIGNATURES = {
    # tf_ops drop-ins
    "caae_fps": "iiipppp",
    "caae_fps_gather": "iiippppp",
    "caae_gather": "iiipppp",
    "caae_gather_grad": "iiipppp",
    "caae_prob_sample": "iiippppp",
    "caae_nn_distance": "iipippppp" "p",
    "caae_nn_distance_grad": "iipippppppp" "p",
    # model building blocks
    "caae_gemm_f32": "iiiiipipipipi" "p",
    "caae_gemm_tf32": "iiiiipipipipi" "p",
    "caae_gemm_tf32_stats": "iiipipipipp" "p",
    "caae_gemm_tf32x3": "iiiiippippipipip" "p",
    "caae_split_tf32": "lipipi" "p",
    "caae_gemm_tf32_pool": "iiippippipppiipp" "p",
    "caae_knn": "iiiipip" "p",
    "caae_knn_ffma": "iiiipip" "p",
    "caae_knn_classify": "iiipip" "p",
    "caae_debug_knn_shortlist": "iiiipipp" "p",
    "caae_knn_part": "ipiiiipip" "p",
    "caae_edge_stats": "iiiipipp" "p",
    "caae_edge_apply": "iiiipippppip" "p",
    "caae_edge_apply_fused": "iiiipippid" "pppppppppp" "ip" "ppi" "p",
    "caae_edge_bwd_stats": "iiiiipippipp" "p",
    "caae_edge_bwd_apply_fused": "iiiipip" "ppppp" "id" "pppp" "pipi" "p",
    "caae_edge_bwd_reduce": "iiiipippppppip" "p",
    "caae_edge_bwd_apply": "iiiipipppppppipi" "p",
    "caae_col_stats": "iipip" "p",
    "caae_bn_finalize": "ipidppppppppp" "p",
    "caae_bn_eval_coeffs": "ippppp" "pp",
    "caae_bn_bwd_finalize": "ipidppppp" "p",
    "caae_bn_act": "iipippipi" "p",
    "caae_bn_act_pool": "iiipippipppp" "p",
    "caae_bn_pool_bwd_finalize": "iiipifpppppppp" "p",
    "caae_bn_act_bwd_reduce": "iipipppppiifipp" "p",
    "caae_bn_act_bwd_apply": "iipippppppiifippi" "p",
    "caae_colsum": "iipip" "p",
    "caae_fc_bn_fwd": "iipipppppppppipip" "p",
    "caae_fc_bn_bwd": "iipipppppipipippp" "p",
    "caae_add3": "lpppp" "p",
    "caae_edge_fold_weights": "iippppi" "p",
    "caae_edge_unfold_wgrad": "iipip" "p",
    # losses / optimiser / step state
    "caae_pose_losses": "ipppppffppppp" "p",
    "caae_loss_reduce": "lppippp" "p",
    "caae_add_cloud_vec": "iippp" "p",
    "caae_prepare_input": "iiipppipp" "p",
    "caae_step_begin": "pi" "p",
    "caae_adam_tf": "lpppppfffff" "p",
    "caae_fill_f32": "lpf" "p",
    # on-line synthesis
    "caae_philox_fill": "lpQipi" "p",
    "caae_synth_points": "iiippppppffffppp" "p",
    "caae_hpr_select": "iippiipppp" "p",
    "caae_hpr_select_pair": "iipipppipippppi" "p",
    # real-segment front end of evaluation, ICP refinement
    "caae_segment_extract": "iiiippppppipppppp" "p",
    "caae_radius_outlier": "iippidippp" "p",
    "caae_fps_seeded_f64": "iiipppppp" "p",
    "caae_icp_refine": "iiippippddiiddpppp" "p",
    "caae_pose_transform_models": "iiipppppp" "p",
    "caae_add_reduce": "iippppp" "p",
}
_SIGNATURES = {k: [_CODES[c] for c in v] for k, v in _SIGNATURES.items()}
_SPECIAL = {
    "caae_abi_version": ([], _int),
    "caae_status_string": ([_int], ctypes.c_char_p),
    "caae_crc32c": ([ctypes.c_uint, _ptr, ctypes.c_ulonglong], ctypes.c_uint),
    "caae_fps_scratch_bytes": ([_int, _int], ctypes.c_size_t),
    "caae_edge_parts": ([_int, _int, _int, _int, _int], _int),
    "caae_col_parts": ([_int], _int),
    "caae_gemm_tf32_stats_parts": ([_int, _int, _int, _int], _int),
    "caae_debug_hpr_timing": ([_ptr], _int),
    "caae_gemm_tf32_supported": ([_int, _int, _int, _int, _int, _ptr, _int, _ptr, _int], _int),
}

EXPORTED_SYMBOLS = tuple(sorted(list(_SIGNATURES) + list(_SPECIAL)))

_lib = None


def lib() -> ctypes.CDLL:
    """Load the native library once; raise loudly when it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the sm_100a kernels are not built. Run `python -c 'import "
            f"__graft_entry__ as g; g.build()'` or `make -C cloudaae_b200/csrc`. There is no CPU fallback.")
    handle = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(handle, name)
        fn.argtypes = argtypes
        fn.restype = _int
    for name, (argtypes, restype) in _SPECIAL.items():
        fn = getattr(handle, name)
        fn.argtypes = argtypes
        fn.restype = restype
    got = handle.caae_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"{LIB_PATH} has ABI version {got}, expected {ABI_VERSION}: rebuild it")
    _lib = handle
    return _lib


COUNTER = [0]  # C-ABI kernel-launching calls issued by this process (each launches >= 1 kernel)
CALLS: dict = {}  # the same, per entry point (tests assert which kernel family a configuration really took)


def check(status: int, what: str) -> None:
    COUNTER[0] += 1
    CALLS[what] = CALLS.get(what, 0) + 1
    if status != 0:
        msg = lib().caae_status_string(status).decode()
        raise CloudAAENativeError(f"{what} failed with status {status}: {msg}")


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def stream_of(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise NotImplementedError(
            f"{name}: cloudaae_b200 ops run on CUDA (sm_100a) tensors only; there is no CPU kernel")
