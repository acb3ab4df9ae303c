import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "point-cloud-donkey_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no GPU in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The oracle and the CUDA library are prebuilt by __graft_entry__.build(); build on demand if missing."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_build", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    if not os.path.exists(os.path.join(PKG, "csrc", "libpcdb200.so")):
        subprocess.check_call(["make", "-C", os.path.join(PKG, "csrc"), "-j8", "-s"])


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle_py
    return oracle_py


@pytest.fixture(scope="session")
def small_world(orc):
    """A tiny trained world shared by several tests: params, training clouds, codebook (trained by the oracle from
    oracle features), test clouds."""
    import numpy as np
    from pcdb200 import synth
    prm = synth.workload_params("c2")
    n_cls, n_train, P = 4, 3, 1536
    tr_cls = [c for c in range(n_cls) for _ in range(n_train)]
    xyz, nrm, rgb, off = synth.make_clouds(tr_cls, [1000 + i for i in range(len(tr_cls))], P)
    fx, fl, fd, foff = orc.compute_features(prm, xyz, nrm, rgb, off)
    bb = np.stack([orc.aabb(xyz[off[i]:off[i + 1]]) for i in range(len(tr_cls))])
    cb = orc.train(prm, fx, fl, fd, foff, tr_cls, list(range(len(tr_cls))), bb, n_cls)
    te_cls = [c for c in range(n_cls) for _ in range(2)]
    xt, nt, rt, ot = synth.make_clouds(te_cls, [5000 + i for i in range(len(te_cls))], P)
    return dict(prm=prm, n_cls=n_cls, train=(xyz, nrm, rgb, off, tr_cls), feats=(fx, fl, fd, foff), cb=cb,
                test=(xt, nt, rt, ot, te_cls))
