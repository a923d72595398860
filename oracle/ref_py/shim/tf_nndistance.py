"""Stand-in for the reference's tf_ops/nn_distance module (its real one needs TensorFlow): brute-force squared
nearest-neighbour distances with first-argmin indices, chunked over the batch.  TEST INFRASTRUCTURE ONLY."""
import torch


def nn_distance(xyz1, xyz2, chunk=8):
    d1, i1, d2, i2 = [], [], [], []
    for s in range(0, xyz1.shape[0], chunk):
        d = ((xyz1[s:s + chunk].unsqueeze(2) - xyz2[s:s + chunk].unsqueeze(1)) ** 2).sum(-1)
        a, b = d.min(dim=2), d.min(dim=1)
        d1.append(a.values); i1.append(a.indices); d2.append(b.values); i2.append(b.indices)
    return torch.cat(d1), torch.cat(i1), torch.cat(d2), torch.cat(i2)
