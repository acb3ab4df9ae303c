"""Minimal PCD writer/reader used to feed the C++ eval_tool from synthetic clouds (tests, examples)."""
import numpy as np

_FIELDS = "x y z rgb normal_x normal_y normal_z curvature"


def write_pcd(path, xyz, normals, rgb, ascii=False):
    n = xyz.shape[0]
    if normals is None:  # raw scan: x y z rgb only (the reference then estimates the normals)
        head = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F U\n"
                "COUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n"
                % (n, n, "ascii" if ascii else "binary"))
        with open(path, "wb") as f:
            f.write(head.encode())
            if ascii:
                for i in range(n):
                    f.write(("%.9g %.9g %.9g %d\n" % (xyz[i, 0], xyz[i, 1], xyz[i, 2], int(rgb[i]))).encode())
            else:
                rec = np.zeros(n, dtype=[("xyz", np.float32, 3), ("rgb", np.uint32)])
                rec["xyz"], rec["rgb"] = xyz, rgb
                f.write(rec.tobytes())
        return
    head = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE 4 4 4 4 4 4 4 4\n"
            "TYPE F F F U F F F F\nCOUNT 1 1 1 1 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA %s\n"
            % (_FIELDS, n, n, "ascii" if ascii else "binary"))
    with open(path, "wb") as f:
        f.write(head.encode())
        if ascii:
            for i in range(n):
                f.write(("%.9g %.9g %.9g %d %.9g %.9g %.9g 0\n" % (xyz[i, 0], xyz[i, 1], xyz[i, 2], int(rgb[i]),
                                                                  normals[i, 0], normals[i, 1], normals[i, 2])).encode())
        else:
            rec = np.zeros(n, dtype=[("xyz", np.float32, 3), ("rgb", np.uint32), ("n", np.float32, 3), ("c", np.float32)])
            rec["xyz"], rec["rgb"], rec["n"] = xyz, rgb, normals
            f.write(rec.tobytes())
