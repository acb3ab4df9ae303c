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
