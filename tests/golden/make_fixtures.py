"""Extract the reference's input fixtures into small committed arrays.

Run in the build container (needs /root/reference):
    python tests/golden/make_fixtures.py

Writes
  tests/golden/ycb_models_xyz.npy   float32 [21,2048,3]  (xyz columns of obj_models.tfrecords)
  tests/golden/ycb_poses.npz        translation f32[21*P,3], axisangle f32[21*P,3], class_id i64[21*P]
                                    = the first P=256 records of every <cls>_syn.tfrecords
  tests/golden/ref_model.ckpt.index the TensorBundle index of the reference's shipped network snapshot
                                    (trained_network/20200908-204328/model.ckpt.index, 6.6 KB; names, shapes and
                                    offsets only — the reference does not ship the tensor data)
The GPU box has no /root/reference, so tests and bench.py read these instead.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cloudaae_b200.data.tfrecord import read_object_models, read_pose_records  # noqa: E402

REF = os.environ.get("CLOUDAAE_REFERENCE", "/root/reference")
P = 256


def main():
    models = read_object_models(os.path.join(REF, "object_model_tfrecord/obj_models.tfrecords"))
    assert models.shape == (21, 2048, 6), models.shape
    np.save(os.path.join(HERE, "ycb_models_xyz.npy"), np.ascontiguousarray(models[:, :, :3]))
    ts, axs, cs = [], [], []
    total = 0
    for cls in range(21):
        t, a, c = read_pose_records(
            os.path.join(REF, f"ycb_video_data_tfRecords/train_syn/{cls}_syn.tfrecords"))
        assert (c == cls).all()
        total += len(c)
        ts.append(t[:P]); axs.append(a[:P]); cs.append(c[:P])
    np.savez(os.path.join(HERE, "ycb_poses.npz"), translation=np.concatenate(ts),
             axisangle=np.concatenate(axs), class_id=np.concatenate(cs), total_records=np.int64(total))
    print("models", models.shape, "poses kept", sum(map(len, cs)), "of", total)
    import shutil
    shutil.copyfile(os.path.join(REF, "trained_network/20200908-204328/model.ckpt.index"),
                    os.path.join(HERE, "ref_model.ckpt.index"))


if __name__ == "__main__":
    main()
