"""Generate golden vectors for the custom ops (run in the build container).

    python tests/golden/make_golden_ops.py

Outputs tests/golden/ops_golden.npz.  Inputs are regenerated from seeds by the tests
(tests/cases.py); only expected OUTPUTS are stored:
  nnd_smoke_*  the reference's own smoke-test arrays (tf_nndistance.py:42-49: np.random.seed(100),
               randn(32,16384,3) / randn(32,1024,3)), first 2 clouds.
               *_cpu = produced by the REFERENCE'S OWN CPU OpKernel compiled from /root/reference
               (oracle/_ref), *_gpu = oracle 'gpu' arithmetic mode (fma order of the CUDA kernel).
  fps_ycb      FPS 2048->256 of the 21 YCB models, each posed by record 0 of its class; oracle.
  fps_ties     FPS 512->64 on a 2048-point cloud whose second half duplicates the first
               (exercises the (k mod 512) tie rule).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ops as O  # noqa: E402
import cases  # noqa: E402


def main():
    out = {}
    x1, x2 = cases.nnd_smoke_inputs()
    d1, i1, d2, i2 = O.ref_cpu_nn_distance(x1, x2)          # the reference's own code
    od1, oi1, od2, oi2 = O.nn_distance(x1, x2, "cpu")
    assert (d1 == od1).all() and (i1 == oi1).all() and (d2 == od2).all() and (i2 == oi2).all()
    out.update(nnd_smoke_cpu_dist1=d1, nnd_smoke_cpu_idx1=i1.astype(np.int16), nnd_smoke_cpu_dist2=d2,
               nnd_smoke_cpu_idx2=i2.astype(np.int16))
    g = O.nn_distance(x1, x2, "gpu")
    out.update(nnd_smoke_gpu_dist1=g[0], nnd_smoke_gpu_idx1=g[1].astype(np.int16), nnd_smoke_gpu_dist2=g[2],
               nnd_smoke_gpu_idx2=g[3].astype(np.int16))
    gd1, gd2 = cases.nnd_smoke_grads()
    r1, r2 = O.ref_cpu_nn_distance_grad(x1, x2, gd1, i1, gd2, i2)
    out.update(nnd_smoke_cpu_gxyz1=r1, nnd_smoke_cpu_gxyz2=r2)

    out["fps_ycb"] = O.fps(cases.fps_ycb_inputs(), 256, threads=8).astype(np.int16)
    out["fps_ties"] = O.fps(cases.fps_ties_inputs(), 64).astype(np.int16)
    np.savez_compressed(os.path.join(HERE, "ops_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
