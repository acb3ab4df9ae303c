"""CPU test of the product-side training bookkeeping (pcdb200/train.py) with the GPU calls replaced by the oracle:
the resulting codebook must equal the oracle's own ImplicitShapeModel::train restatement."""
import numpy as np
import pytest

from pcdb200 import synth, train
from pcdb200.structs import DIST_CHISQUARED, DIST_EUCLIDEAN


class OracleBackedCtx:
    """Stands in for pcdb200.api.Context on a machine without a GPU (test infrastructure only)."""

    def __init__(self, orc, prm):
        self.orc, self.prm, self.m = orc, prm, None

    def set_params(self, prm):
        self.prm = prm
        if self.m is not None:
            self.m.set_params(prm)

    def set_codebook(self, cb, row_base=0):
        self.m = self.orc.Model(self.prm, cb)

    def knn(self, q, k=None, dist_type=None, mode=0):
        return self.m.knn(q, k=k, dist_type=dist_type)

    def distance_pairs(self, a, b, dist_type):
        return self.orc.distance(a, b, dist_type)


@pytest.mark.parametrize("dist,k", [(DIST_EUCLIDEAN, 1), (DIST_CHISQUARED, 1), (DIST_EUCLIDEAN, 3)])
def test_train_codebook_matches_oracle(orc, dist, k):
    prm = synth.workload_params("c2", distance_type=dist, knn_k=k)
    tr_cls = [0, 0, 1, 1, 1, 3]
    xyz, nrm, rgb, off = synth.make_clouds(tr_cls, [70 + i for i in range(len(tr_cls))], 1200)
    fx, fl, fd, foff = orc.compute_features(prm, xyz, nrm, rgb, off)
    bb = np.stack([train.aabb(xyz[off[i]:off[i + 1]]) for i in range(len(tr_cls))])
    assert np.array_equal(bb, np.stack([orc.aabb(xyz[off[i]:off[i + 1]]) for i in range(len(tr_cls))]))
    inst = [5, 6, 7, 8, 9, 10]
    ref = orc.train(prm, fx, fl, fd, foff, tr_cls, inst, bb, 4)
    got = train.train_codebook(OracleBackedCtx(orc, prm), prm, fx, fl, fd, foff, tr_cls, inst, bb, 4)
    assert got.N == ref.N and got.N > 0
    assert np.array_equal(got.codeword_ids, ref.codeword_ids)
    assert np.array_equal(got.words, ref.words)
    assert np.array_equal(got.vote_off, ref.vote_off)
    assert np.array_equal(got.vote_class, ref.vote_class) and np.array_equal(got.vote_instance, ref.vote_instance)
    assert np.array_equal(got.vote_xyz.view(np.uint32), ref.vote_xyz.view(np.uint32))  # same float op order
    assert np.array_equal(got.vote_bbox.view(np.uint32), ref.vote_bbox.view(np.uint32))
    assert np.array_equal(got.kp_train, ref.kp_train)
    assert np.allclose(got.vote_weight, ref.vote_weight, rtol=1e-6)
    assert np.allclose(got.sigma2, ref.sigma2, rtol=1e-6), (got.sigma2, ref.sigma2)


def test_lrf_quat_negative_trace_branch(orc):
    # rotations by ~180 degrees exercise the "trace <= 0" branch of matrix2Quat
    rng = np.random.default_rng(0)
    R = []
    for _ in range(50):
        a = rng.normal(size=3)
        a /= np.linalg.norm(a)
        R.append(2 * np.outer(a, a) - np.eye(3) + 1e-3 * rng.normal(size=(3, 3)))
    rf = np.array(R, np.float32).reshape(-1, 9)
    q = train.lrf_quat(rf)
    v = rng.normal(size=(50, 3)).astype(np.float32)
    back = train.quat_rotate_inv(q, train.quat_rotate(q, v))
    n2 = (q.astype(np.float64) ** 2).sum(1, keepdims=True)
    assert np.allclose(back, v * n2 * n2, atol=1e-4)
