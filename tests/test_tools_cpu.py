"""CPU tests of the measurement helpers under tools/ (they turn ncu output into the files committed under profiles/)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _row(i, kernel, ns):
    return '"%d","1","python","127.0.0.1","%s","1","7","(1, 1, 1)","(128, 1, 1)","0","10.0","Command line profiler metrics",' \
           '"gpu__time_duration.sum","ns","%s"' % (i, kernel, ns)


def test_launch_list_picks_the_second_full_batch_call(tmp_path):
    hdr = '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device",' \
          '"CC","Section Name","Metric Name","Metric Unit","Metric Value"'
    lines = ["==PROF== Connected to process 1", hdr]
    i = 0

    def call(gemm_ns, extra=0):
        nonlocal i
        for k, ns in (("void <unnamed>::k_flags(float const*, int)", "1,000"),
                      ("void <unnamed>::k_knn_gemm<(bool)1, (int)1, (bool)1>(CUtensorMap_st, GemmArgs)", gemm_ns),
                      ("void cub::CUB_200_1::DeviceScanKernel<int, long>(int*)", "2,000"),
                      ("void cub::CUB_200_1::DeviceScanKernel<int, long>(int*)", "3,000")) + \
                (("void <unnamed>::k_extra(int)", "5,000"),) * extra:
            lines.append(_row(i, k, ns))
            i += 1
    call("500,000")          # a small (single-cloud) call: below the threshold
    call("40,000,000")       # warm-up step
    call("60,000,000", 1)    # the timed step
    call("61,000,000")       # a later step
    p = tmp_path / "launches.csv"
    p.write_text("\n".join(lines) + "\n")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_list.py"), str(p), "30"],
                         capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0].startswith("#") and "5 launches" in out[0] and "3 full-batch calls" in out[0]
    assert out[1] == "kernel,launches,ms,share"
    first = out[2].split(",")
    assert first[0].startswith("k_knn_gemm<") and first[-3] == "1" and abs(float(first[-2]) - 60.0) < 1e-6
    rows = {r.rsplit(",", 3)[0]: r.rsplit(",", 3)[1:] for r in out[2:]}
    assert rows["cub::DeviceScanKernel"][0] == "2" and abs(float(rows["cub::DeviceScanKernel"][1]) - 0.005) < 1e-9
    assert "k_extra" in rows and abs(sum(float(v[2]) for v in rows.values()) - 1.0) < 1e-3
