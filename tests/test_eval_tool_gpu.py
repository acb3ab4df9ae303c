"""GPU test of the C++ drop-in front-end: eval_tool train + classify on synthetic PCD files (the reference's quick-start
flow, README / data/qs_*_list.txt) must give the labels the Python path and the oracle give, and the .ism/.ismd pair it
writes must load back."""
import os
import re
import subprocess

import numpy as np
import pytest

from pcdb200 import pcd, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "point-cloud-donkey_b200", "host")


def test_eval_tool_train_and_classify(tmp_path, orc):
    from pcdb200 import api, train
    tool = os.path.join(HOST, "eval_tool")
    if not os.path.exists(tool):
        subprocess.check_call(["make", "-C", HOST, "-s"])
    names = ["cat", "horse", "wolf", "lion"]
    n_train, n_test, P = 2, 2, 1536
    prm = synth.workload_params("c2")
    tr_cls = [c for c in range(len(names)) for _ in range(n_train)]
    te_cls = [c for c in range(len(names)) for _ in range(n_test)]
    xyz, nrm, rgb, off = synth.make_clouds(tr_cls, [100 + i for i in range(len(tr_cls))], P)
    xt, nt, rt, ot = synth.make_clouds(te_cls, [900 + i for i in range(len(te_cls))], P)
    with open(tmp_path / "train.txt", "w") as f:
        f.write("# train\n")
        for i, c in enumerate(tr_cls):
            p = str(tmp_path / ("train_%d.pcd" % i))
            pcd.write_pcd(p, xyz[off[i]:off[i + 1]], nrm[off[i]:off[i + 1]], rgb[off[i]:off[i + 1]], ascii=(i == 0))
            f.write("%s %s\n" % (p, names[c]))
    with open(tmp_path / "test.txt", "w") as f:
        f.write("# test\n")
        for i, c in enumerate(te_cls):
            p = str(tmp_path / ("test_%d.pcd" % i))
            pcd.write_pcd(p, xt[ot[i]:ot[i + 1]], nt[ot[i]:ot[i + 1]], rt[ot[i]:ot[i + 1]])
            f.write("%s %s\n" % (p, names[c]))
    cfg = os.path.join(ROOT, "config", "c2_synthetic.ism")
    model = str(tmp_path / "model.ism")
    r = subprocess.run([tool, "-t", cfg, "-f", str(tmp_path / "train.txt"), "-o", model], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert os.path.exists(model) and os.path.exists(model + "d")
    outdir = str(tmp_path / "out")
    r = subprocess.run([tool, "-d", model, "-f", str(tmp_path / "test.txt"), "-o", outdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    summary = open(os.path.join(outdir, "summary.txt")).read()
    got = [int(m) for m in re.findall(r"classified class: (-?\d+)", summary)]
    assert len(got) == len(te_cls)
    # same pipeline through the Python binding (GPU) and the oracle (CPU)
    ctx = api.Context(prm)
    fx, fl, fd, foff = ctx.compute_features(xyz, nrm, rgb, off)
    bb = np.stack([train.aabb(xyz[off[i]:off[i + 1]]) for i in range(len(tr_cls))])
    cb = train.train_codebook(ctx, prm, fx, fl, fd, foff, tr_cls, tr_cls, bb, len(names))
    ctx.set_codebook(cb)
    labels, _, _ = ctx.classify_batch(xt, nt, rt, ot)
    ref, _, _ = orc.Model(prm, cb).classify_batch(xt, nt, rt, ot)
    assert got == labels.tolist() == ref.tolist()
    assert "Accuracy:" in summary and "class id to class name mapping" in summary and "0: cat" in summary
    ctx.close()


def test_eval_tool_on_clouds_without_normals(tmp_path):
    """Raw x y z rgb PCD files: the front-end estimates the normals (NormalRadius / ConsistentNormalsMethod of the
    config) in training and in classification, like the reference's hasNormals == false path; the labels must be the
    ones the Python binding gives with normals=None."""
    import json
    from pcdb200 import api, train
    tool = os.path.join(HOST, "eval_tool")
    names = ["cat", "horse", "wolf"]
    P = 1536
    tr_cls = [c for c in range(3) for _ in range(2)]
    te_cls = [0, 1, 2]
    xyz, _, rgb, off = synth.make_clouds(tr_cls, [300 + i for i in range(len(tr_cls))], P)
    xt, _, rt, ot = synth.make_clouds(te_cls, [800 + i for i in range(len(te_cls))], P)
    cfg = json.load(open(os.path.join(ROOT, "config", "c2_synthetic.ism")))
    cfg["ObjectConfig"]["Parameters"]["NormalRadius"] = 0.08
    cfg["ObjectConfig"]["Parameters"]["ConsistentNormalsMethod"] = 2
    cfg_path = str(tmp_path / "cfg.ism")
    json.dump(cfg, open(cfg_path, "w"), indent=1)
    for name, (x, r, o, cls) in {"train": (xyz, rgb, off, tr_cls), "test": (xt, rt, ot, te_cls)}.items():
        with open(tmp_path / (name + ".txt"), "w") as f:
            f.write("# %s\n" % name)  # the reference's list files open with a header line (data/qs_train_list.txt)
            for i, c in enumerate(cls):
                p = str(tmp_path / ("%s_%d.pcd" % (name, i)))
                pcd.write_pcd(p, x[o[i]:o[i + 1]], None, r[o[i]:o[i + 1]], ascii=(i == 1))
                f.write("%s %s\n" % (p, names[c]))
    model = str(tmp_path / "model.ism")
    r = subprocess.run([tool, "-t", cfg_path, "-f", str(tmp_path / "train.txt"), "-o", model], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    outdir = str(tmp_path / "out")
    r = subprocess.run([tool, "-d", model, "-f", str(tmp_path / "test.txt"), "-o", outdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    summary = open(os.path.join(outdir, "summary.txt")).read()
    got = [int(m) for m in re.findall(r"classified class: (-?\d+)", summary)]
    prm = synth.workload_params("c2", normal_radius=0.08, consistent_normals_method=2)
    ctx = api.Context(prm)
    fx, fl, fd, foff = ctx.compute_features(xyz, None, rgb, off)
    bb = np.stack([train.aabb(xyz[off[i]:off[i + 1]]) for i in range(len(tr_cls))])
    cb = train.train_codebook(ctx, prm, fx, fl, fd, foff, tr_cls, tr_cls, bb, len(names))
    ctx.set_codebook(cb)
    labels, _, _ = ctx.classify_batch(xt, None, rt, ot)
    assert got == labels.tolist()
    ctx.close()


def test_eval_tool_multi_gpu(tmp_path):
    """--gpus 2: the test list is sharded over two devices (one context and host thread each, no communication);
    the summary must be identical to the single-GPU run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    tool = os.path.join(HOST, "eval_tool")
    names = ["cat", "horse", "wolf"]
    P = 1536
    tr_cls = [c for c in range(3) for _ in range(2)]
    te_cls = [i % 3 for i in range(7)]
    sets = {"train": (synth.make_clouds(tr_cls, [400 + i for i in range(len(tr_cls))], P), tr_cls),
            "test": (synth.make_clouds(te_cls, [600 + i for i in range(len(te_cls))], P), te_cls)}
    for name, ((x, n, r, o), cls) in sets.items():
        with open(tmp_path / (name + ".txt"), "w") as f:
            f.write("# %s\n" % name)
            for i, c in enumerate(cls):
                p = str(tmp_path / ("%s_%d.pcd" % (name, i)))
                pcd.write_pcd(p, x[o[i]:o[i + 1]], n[o[i]:o[i + 1]], r[o[i]:o[i + 1]])
                f.write("%s %s\n" % (p, names[c]))
    cfg = os.path.join(ROOT, "config", "c2_synthetic.ism")
    model = str(tmp_path / "model.ism")
    r = subprocess.run([tool, "-t", cfg, "-f", str(tmp_path / "train.txt"), "-o", model], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    outs = []
    for extra, tag in (([], "one"), (["--gpus", "2"], "two")):
        outdir = str(tmp_path / tag)
        r = subprocess.run([tool, "-d", model, "-f", str(tmp_path / "test.txt"), "-o", outdir] + extra,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        summary = open(os.path.join(outdir, "summary.txt")).read()
        outs.append(re.findall(r"file: (\S+), ground truth class: (\d+), classified class: (-?\d+)", summary))
    assert outs[0] == outs[1] and len(outs[0]) == len(te_cls)
