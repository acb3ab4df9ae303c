"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol include/pcdb200.h declares, and
the ctypes mirrors have the C structs' sizes.  No compute call is made (there is no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

from pcdb200 import api
from pcdb200.structs import MAXIMUM_DTYPE, VOTE_DTYPE, Params, Stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcdb200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcdb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = api.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libpcdb200.so does not export %s" % name
    assert sorted(api.SYMBOLS) == declared, "pcdb200/api.py binds a different symbol set than the header declares"
    lib.pcdb_abi_version.restype = C.c_int
    assert lib.pcdb_abi_version() == 5


def test_struct_layouts_match_the_header():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "pcdb200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(pcdb_params), sizeof(pcdb_vote), sizeof(pcdb_maximum),
         sizeof(pcdb_stats), offsetof(pcdb_params, bandwidth), offsetof(pcdb_maximum, vote_begin),
         offsetof(pcdb_vote, bbox_quat));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    sizes = [int(x) for x in out]
    assert sizes[0] == C.sizeof(Params)
    assert sizes[1] == VOTE_DTYPE.itemsize
    assert sizes[2] == MAXIMUM_DTYPE.itemsize
    assert sizes[3] == C.sizeof(Stats)
    assert sizes[4] == Params.bandwidth.offset
    assert sizes[5] == MAXIMUM_DTYPE.fields["vote_begin"][1]
    assert sizes[6] == VOTE_DTYPE.fields["bbox_quat"][1]


def test_header_compiles_as_c_and_cpp():
    with tempfile.TemporaryDirectory() as d:
        for ext, cc in ((".c", "/usr/bin/gcc"), (".cpp", "/usr/bin/g++")):
            f = os.path.join(d, "h" + ext)
            open(f, "w").write('#include "pcdb200.h"\nint main(void){return PCDB_ABI_VERSION - 2;}\n')
            subprocess.check_call([cc, "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), f, "-o", f + ".out"])


def test_no_fallback_without_gpu():
    """pcdb_create must fail loudly when there is no usable sm_100 device (this container has none)."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    try:
        api.Context()
    except api.PcdbError as e:
        assert e.code == -2  # PCDB_E_NO_DEVICE
    else:
        raise AssertionError("pcdb_create succeeded without a GPU")


def test_product_does_not_touch_the_oracle():
    """Nothing under point-cloud-donkey_b200/ may import, include or link the oracle."""
    pkg = os.path.join(ROOT, "point-cloud-donkey_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle_py" not in text and "liboracle" not in text and "pcd_oracle" not in text, \
                    "%s references the oracle" % os.path.join(dirpath, fn)
