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


def _params_from_config(cfg):
    """The oracle-side parameter struct of a (restated) shipped configuration."""
    from pcdb200.structs import DIST_CHISQUARED, DIST_EUCLIDEAN, FEATURE_CSHOT, FEATURE_SHOT, default_params
    oc = cfg["ObjectConfig"]
    ch, top = oc["Children"], oc["Parameters"]
    v, f = ch["Voting"]["Parameters"], ch["Features"]
    return default_params(
        feature_type=FEATURE_CSHOT if f["Type"] == "CSHOT" else FEATURE_SHOT, feature_radius=f["Parameters"]["Radius"],
        lrf_radius=f["Parameters"]["ReferenceFrameRadius"], leaf_size=ch["Keypoints"]["Parameters"]["LeafSize"],
        distance_type=DIST_CHISQUARED if top["DistanceType"] == "ChiSquared" else DIST_EUCLIDEAN, knn_k=1,
        bandwidth=v["Bandwidth"], ms_threshold=v["Threshold"], ms_max_iter=v["MaxIter"],
        average_rotation=int(v["AverageRotation"]), single_object_mode=int(v["SingleObjectMode"]),
        normal_radius=top["NormalRadius"], consistent_normals_method=top["ConsistentNormalsMethod"])


@pytest.mark.parametrize("which", ["qs_input_config", "default_config_kinect"])
def test_eval_tool_with_the_reference_shipped_configs(tmp_path, orc, which):
    """BASELINE.json config 1 / config 4 as shipped: eval_tool -t <the reference's own configuration> on synthetic PCD
    files of the scale the file is written for (qs: millimetre-like units, x y z only -> the normals stage runs;
    kinect: 0.2 m objects with colour and normals), ChiSquared activation, single-object mode.  The labels must equal
    the oracle's for the same parameters.  (The configuration is tests/ref_configs.py's restatement of the shipped
    file — pinned to the real file by tests/test_host_formats.py where the reference is mounted.)"""
    import ref_configs
    tool = os.path.join(HOST, "eval_tool")
    qs = which == "qs_input_config"
    cfg = ref_configs.QS_INPUT_CONFIG if qs else ref_configs.DEFAULT_CONFIG_KINECT
    cfg_path = ref_configs.write(cfg, str(tmp_path / (which + ".ism")))
    names = ["bottle", "mug", "plant", "chair", "lamp"]
    scale, P = (250.0, 5000) if qs else (0.2, 6000)
    tr_cls, te_cls = list(range(5)), list(range(5))
    xyz, nrm, rgb, off = synth.make_clouds(tr_cls, [100 + c for c in tr_cls], P, scale=scale)
    xt, nt, rt, ot = synth.make_clouds(te_cls, [200 + c for c in te_cls], P, scale=scale)
    if qs:
        nrm = nt = None  # quick-start clouds carry no normals
    for name, (x, n, r, o, cls) in {"train": (xyz, nrm, rgb, off, tr_cls), "test": (xt, nt, rt, ot, te_cls)}.items():
        with open(tmp_path / (name + ".txt"), "w") as f:
            f.write("# %s\n" % name)
            for i, c in enumerate(cls):
                p = str(tmp_path / ("%s_%d.pcd" % (name, i)))
                pcd.write_pcd(p, x[o[i]:o[i + 1]], None if n is None else n[o[i]:o[i + 1]], r[o[i]:o[i + 1]])
                f.write("%s %s\n" % (p, names[c]))
    model = str(tmp_path / "model.ism")
    r = subprocess.run([tool, "-t", cfg_path, "-f", str(tmp_path / "train.txt"), "-o", model], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if qs:  # BoundingBoxType MVBB: trained with AABB, said so, and recorded in the written model
        assert "MVBB" in r.stderr and "AABB" in r.stderr
        import json
        assert json.load(open(model))["ObjectConfig"]["Parameters"]["BoundingBoxType"] == "AABB"
    outdir = str(tmp_path / "out")
    r = subprocess.run([tool, "-d", model, "-f", str(tmp_path / "test.txt"), "-o", outdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = [int(m) for m in re.findall(r"classified class: (-?\d+)", open(os.path.join(outdir, "summary.txt")).read())]
    # the oracle, same parameters, same clouds
    prm = _params_from_config(cfg)
    fx, fl, fd, foff = orc.compute_features(prm, xyz, nrm, rgb, off)
    bb = np.stack([orc.aabb(xyz[off[i]:off[i + 1]]) for i in range(len(tr_cls))])
    cb = orc.train(prm, fx, fl, fd, foff, tr_cls, tr_cls, bb, len(names))
    ref, _, _ = orc.Model(prm, cb).classify_batch(xt, nt, rt, ot)
    assert got == ref.tolist(), (got, ref.tolist())
    assert sum(int(a == b) for a, b in zip(got, te_cls)) >= 4  # and they are the right classes


def test_eval_tool_detection_on_a_scene(tmp_path, orc):
    """eval_tool_detection (src/eval_tool/eval_detection.cpp): detect on a cluttered scene, match the maxima with the
    ground-truth annotation, write summary.txt with precision / recall / AP."""
    import json
    tool, dtool = os.path.join(HOST, "eval_tool"), os.path.join(HOST, "eval_tool_detection")
    names = ["cat", "horse", "wolf", "lion"]
    P = 2048
    tr_cls = [c for c in range(4) for _ in range(3)]
    xyz, nrm, rgb, off = synth.make_clouds(tr_cls, [100 + i for i in range(len(tr_cls))], P)
    with open(tmp_path / "train.txt", "w") as f:
        f.write("# train\n")
        for i, c in enumerate(tr_cls):
            p = str(tmp_path / ("train_%d.pcd" % i))
            pcd.write_pcd(p, xyz[off[i]:off[i + 1]], nrm[off[i]:off[i + 1]], rgb[off[i]:off[i + 1]])
            f.write("%s %s\n" % (p, names[c]))
    cfg = json.load(open(os.path.join(ROOT, "config", "c2_synthetic.ism")))
    v = cfg["ObjectConfig"]["Children"]["Voting"]["Parameters"]
    v["SingleObjectMode"], v["MaxFilterType"], v["MinVotesThreshold"], v["MinThreshold"] = False, "Simple", 5, 0.05
    cfg["ObjectConfig"]["Parameters"]["DistanceThresholdDetection"] = 0.25
    cfg_path = str(tmp_path / "det.ism")
    json.dump(cfg, open(cfg_path, "w"), indent=1)
    model = str(tmp_path / "model.ism")
    r = subprocess.run([tool, "-t", cfg_path, "-f", str(tmp_path / "train.txt"), "-o", model], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with open(tmp_path / "test.txt", "w") as f:
        f.write("# test detection\n")
        for s in range(2):
            classes = [(s + j) % 4 for j in range(3)]
            sx, sn, sc, truth = synth.make_scene(classes, 40 + s, P, plane_points=6000, clutter_points=1500)
            p = str(tmp_path / ("scene_%d.pcd" % s))
            pcd.write_pcd(p, sx, sn, sc)
            a = str(tmp_path / ("scene_%d.txt" % s))
            with open(a, "w") as g:
                for cls, centre in truth:
                    g.write("%s (0.0) %.6f %.6f %.6f\n" % (names[cls], centre[0], centre[1], centre[2]))
            f.write("%s %s\n" % (p, a))
    outdir = str(tmp_path / "det_out")
    r = subprocess.run([dtool, "-d", model, "-f", str(tmp_path / "test.txt"), "-o", outdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    summary = open(os.path.join(outdir, "summary.txt")).read()
    m = re.search(r"ground-truth objects: (\d+), detections: (\d+), tp: (\d+), fp: (\d+)", summary)
    assert m and int(m.group(1)) == 6 and int(m.group(3)) >= 4, summary
    assert float(re.search(r"mAP: ([0-9.]+)", summary).group(1)) > 0.5, summary
