"""CPU tests of the host-side formats (C++ shim, no GPU call): boost-archive framing of the .ismd data, .ism JSON
parsing incl. the reference's shipped configs when mounted, PCD ascii/binary reading."""
import os
import subprocess

import numpy as np
import pytest

from pcdb200 import pcd, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "point-cloud-donkey_b200", "host")
TOOL = os.path.join(HOST, "ismd_roundtrip")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(TOOL):
        subprocess.check_call(["make", "-C", HOST, "-s", "ismd_roundtrip"])


def test_boost_archive_framing():
    out = subprocess.check_output([TOOL, "archive"]).decode()
    assert "archive ok" in out


def test_config_parse_and_roundtrip(tmp_path):
    cfg = os.path.join(ROOT, "config", "c2_synthetic.ism")
    out = subprocess.check_output([TOOL, "json", cfg, str(tmp_path / "o.ism")]).decode()
    assert "DistanceType=Euclidean Features=SHOT Radius=0.4 K=1 Bandwidth=0.3" in out


@pytest.mark.parametrize("name,expect", [
    ("qs_input_config.ism", "DistanceType=ChiSquared Features=SHOT Radius=60 K=1 Bandwidth=50"),
    ("default.ism", "DistanceType=Euclidean Features=CSHOT Radius=0.4 K=1 Bandwidth=0.6"),
    ("default_config_kinect.ism", "DistanceType=ChiSquared Features=CSHOT Radius=0.05 K=1 Bandwidth=0.045"),
])
def test_reference_shipped_configs_parse(name, expect):
    path = os.path.join("/root/reference/config", name)
    if not os.path.exists(path):
        pytest.skip("reference not mounted")
    assert expect in subprocess.check_output([TOOL, "json", path]).decode()


@pytest.mark.parametrize("ascii_mode", [False, True])
def test_pcd_reader(tmp_path, ascii_mode):
    xyz, nrm, rgb, off = synth.make_clouds([1], [5], 300)
    p = str(tmp_path / "c.pcd")
    pcd.write_pcd(p, xyz, nrm, rgb, ascii=ascii_mode)
    out = subprocess.check_output([TOOL, "pcd", p]).decode()
    f = dict(kv.split("=") for kv in out.split())
    assert int(f["points"]) == 300 and f["normals"] == "1" and f["rgb"] == "1"
    assert abs(float(f["sum_xyz"]) - float(xyz.astype(np.float64).sum())) < 1e-3
    assert abs(float(f["sum_n"]) - float(nrm.astype(np.float64).sum())) < 1e-3
    assert int(f["sum_rgb"]) == int(rgb.astype(np.uint64).sum())


def _summary(path):
    out = subprocess.check_output([TOOL, "pcd", path]).decode()
    return dict(kv.split("=") for kv in out.split())


def test_pcd_binary_compressed_reader(tmp_path):
    """DATA binary_compressed (LZF, structure-of-arrays payload): literal runs and back references, incl. the long
    overlapping ones a constant column produces."""
    xyz, nrm, rgb, off = synth.make_clouds([2], [6], 700)
    rgb[:300] = rgb[0]  # long runs -> extended-length back references
    p = str(tmp_path / "c.pcd")
    pcd.write_pcd_compressed(p, xyz, nrm, rgb)
    assert os.path.getsize(p) < 700 * 32  # it did compress
    f = _summary(p)
    assert int(f["points"]) == 700 and f["normals"] == "1" and f["rgb"] == "1"
    assert abs(float(f["sum_xyz"]) - float(xyz.astype(np.float64).sum())) < 1e-3
    assert abs(float(f["sum_n"]) - float(nrm.astype(np.float64).sum())) < 1e-3
    assert int(f["sum_rgb"]) == int(rgb.astype(np.uint64).sum())
    # a corrupt stream is an error, not garbage
    raw = bytearray(open(p, "rb").read())
    raw[-40:] = b"\xff" * 40
    open(p, "wb").write(raw)
    assert subprocess.run([TOOL, "pcd", p], capture_output=True).returncode != 0


def test_lzf_roundtrip_matches_python_reference():
    rng = np.random.default_rng(0)
    data = bytes(rng.integers(0, 4, 5000, dtype=np.uint8)) + b"\x00" * 1000 + bytes(rng.integers(0, 256, 300, dtype=np.uint8))
    comp = pcd.lzf_compress(data)
    # independent decoder of the liblzf stream format
    out = bytearray()
    i = 0
    while i < len(comp):
        c = comp[i]
        i += 1
        if c < 32:
            out += comp[i:i + c + 1]
            i += c + 1
        else:
            ln = c >> 5
            if ln == 7:
                ln += comp[i]
                i += 1
            dist = ((c & 31) << 8) + comp[i] + 1
            i += 1
            for _ in range(ln + 2):
                out.append(out[-dist])
    assert bytes(out) == data and len(comp) < len(data)


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
@pytest.mark.parametrize("with_normals", [True, False])
def test_ply_reader(tmp_path, fmt, with_normals):
    xyz, nrm, rgb, off = synth.make_clouds([3], [7], 250)
    p = str(tmp_path / "c.ply")
    pcd.write_ply(p, xyz, nrm if with_normals else None, rgb, fmt=fmt, with_faces=True)
    f = _summary(p)
    assert int(f["points"]) == 250 and f["normals"] == ("1" if with_normals else "0") and f["rgb"] == "1"
    assert abs(float(f["sum_xyz"]) - float(xyz.astype(np.float64).sum())) < 1e-3
    if with_normals:
        assert abs(float(f["sum_n"]) - float(nrm.astype(np.float64).sum())) < 1e-3
    assert int(f["sum_rgb"]) == int(rgb.astype(np.uint64).sum())


def test_unknown_extension_is_rejected(tmp_path):
    p = str(tmp_path / "c.xyz")
    open(p, "w").write("0 0 0\n")
    r = subprocess.run([TOOL, "pcd", p], capture_output=True, text=True)
    assert r.returncode != 0 and "Unknown extension" in (r.stdout + r.stderr)


def test_lzf_reader_against_the_references_own_encoder(tmp_path):
    """tests/golden/lzf_golden.npz holds a PCD payload compressed by the reference's vendored liblzf 3.6 (compiled from
    /root/reference into oracle/_ref by oracle/Makefile; generator tests/golden/make_golden.py).  The host reader must
    decode it to the original column data; when oracle/_ref/libref_lzf.so is present, the reference's decoder must
    also accept what this repository's test compressor writes."""
    import ctypes as C
    G = np.load(os.path.join(ROOT, "tests", "golden", "lzf_golden.npz"))
    raw, comp, n = G["raw"].tobytes(), G["compressed"].tobytes(), int(G["points"])
    head = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb normal_x normal_y normal_z curvature\n"
            "SIZE 4 4 4 4 4 4 4 4\nTYPE F F F U F F F F\nCOUNT 1 1 1 1 1 1 1 1\nWIDTH %d\nHEIGHT 1\n"
            "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary_compressed\n" % (n, n))
    p = str(tmp_path / "ref.pcd")
    with open(p, "wb") as f:
        f.write(head.encode() + np.array([len(comp), len(raw)], np.uint32).tobytes() + comp)
    f_ = _summary(p)
    cols = np.frombuffer(raw, np.float32).reshape(8, n)
    rgb = np.frombuffer(raw, np.uint32).reshape(8, n)[3]
    assert int(f_["points"]) == n and f_["normals"] == "1"
    assert abs(float(f_["sum_xyz"]) - float(cols[:3].astype(np.float64).sum())) < 1e-3
    assert abs(float(f_["sum_n"]) - float(cols[4:7].astype(np.float64).sum())) < 1e-3
    assert int(f_["sum_rgb"]) == int(rgb.astype(np.uint64).sum())
    so = os.path.join(ROOT, "oracle", "_ref", "libref_lzf.so")
    if os.path.exists(so):
        lz = C.CDLL(so)
        lz.lzf_decompress.restype = C.c_uint
        lz.lzf_decompress.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint]
        mine = pcd.lzf_compress(raw[:20000])
        out = C.create_string_buffer(20000)
        assert lz.lzf_decompress(mine, len(mine), out, 20000) == 20000 and out.raw == raw[:20000]


# ---- the reference's shipped configurations through the shim's own config reader (host-only: no device context) ------
CONFIG_PARAMS = os.path.join(HOST, "config_params")


def _params_of(path):
    if not os.path.exists(CONFIG_PARAMS):
        subprocess.check_call(["make", "-C", HOST, "-s", "config_params"])
    r = subprocess.run([CONFIG_PARAMS, path], capture_output=True, text=True)
    return r.returncode, r.stdout.strip(), r.stderr


@pytest.mark.parametrize("name,restated", [("qs_input_config.ism", "QS_INPUT_CONFIG"),
                                           ("default_config_kinect.ism", "DEFAULT_CONFIG_KINECT")])
def test_shipped_configs_read_like_their_restatement(tmp_path, name, restated):
    """ImplicitShapeModel::readObject on the reference's real file and on tests/ref_configs.py's restatement of it (what
    the GPU tests use, where /root/reference does not exist) must yield the same parameters."""
    import ref_configs
    rc, mine, err = _params_of(ref_configs.write(getattr(ref_configs, restated), str(tmp_path / name)))
    assert rc == 0, err
    assert "distance_type=1" in mine and "single_object_mode=1" in mine  # chi^2, single-object mode: as shipped
    real = os.path.join("/root/reference/config", name)
    if not os.path.exists(real):
        pytest.skip("reference not mounted")
    rc, theirs, err = _params_of(real)
    assert rc == 0, err
    assert mine == theirs


def test_default_ism_is_rejected_like_the_reference_rejects_it():
    """config/default.ism has no FeatureWeighting child: the reference's own loader refuses it
    (implicit_shape_model.cpp:1095-1103 "could not find necessary json entries") and so does the shim."""
    real = "/root/reference/config/default.ism"
    if not os.path.exists(real):
        pytest.skip("reference not mounted")
    rc, out, err = _params_of(real)
    assert rc == 2 and "FeatureWeighting" in err


def test_options_outside_the_path_are_named(tmp_path):
    import copy
    import ref_configs
    cfg = copy.deepcopy(ref_configs.DEFAULT_CONFIG_KINECT)
    cfg["ObjectConfig"]["Children"]["Voting"]["Parameters"]["UseGlobalFeatures"] = True
    cfg["ObjectConfig"]["Parameters"]["UseSvmTraining"] = True
    rc, out, err = _params_of(ref_configs.write(cfg, str(tmp_path / "g.ism")))
    assert rc == 2 and "UseGlobalFeatures" in err and "UseSvmTraining" in err
    cfg = copy.deepcopy(ref_configs.QS_INPUT_CONFIG)
    v = cfg["ObjectConfig"]["Children"]["Voting"]["Parameters"]
    v["BinOrBandwidthType"], v["BinOrBandwidthFactor"], v["MaxFilterType"], v["SingleObjectMaxType"] = \
        "BoundingBoxMedian", 0.5, "Merge", "ModelRadiusVotes"
    rc, out, err = _params_of(ref_configs.write(cfg, str(tmp_path / "h.ism")))
    assert rc == 0, err
    assert "radius_type=2 radius_factor=0.5 single_object_max_type=3" in out and "max_filter_type=2" in out


def test_detection_metrics_known_answers(tmp_path):
    """host/eval_detection.h (eval_tool_detection's metrics): greedy matching, per-class AP, dataset sweep, parsers."""
    tool = os.path.join(HOST, "eval_detection_selftest")
    if not os.path.exists(tool):
        subprocess.check_call(["make", "-C", HOST, "-s", "eval_detection_selftest"])
    ann = tmp_path / "scene1.txt"
    ann.write_text("chair (0.25) 1.0 2.0 3.0\n\ntable (0.0) 0.5 0.5 3.5 1 1 1 1 0 0 0\nbook (0.1) 0 0 0\n")
    lst = tmp_path / "list.txt"
    lst.write_text("# test detection\nscene1.pcd scene1.txt\n#skipped.pcd skipped.txt\nscene2.pcd scene2.txt\n")
    out = subprocess.check_output([tool, str(ann), str(lst)]).decode()
    assert "selftest ok" in out
