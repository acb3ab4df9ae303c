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


def lzf_compress(data: bytes) -> bytes:
    """Small LZF compressor (stream format of liblzf): greedy 3-byte hash matching, literal runs of <= 32 bytes,
    back references of 3..264 bytes at distances <= 8192.  Only used to write binary_compressed test files."""
    n = len(data)
    out = bytearray()
    lit = bytearray()
    table = {}
    i = 0

    def flush():
        for s in range(0, len(lit), 32):
            chunk = lit[s:s + 32]
            out.append(len(chunk) - 1)
            out.extend(chunk)
        lit.clear()

    while i < n:
        match_len = 0
        if i + 2 < n:
            key = data[i:i + 3]
            j = table.get(key)
            table[key] = i
            if j is not None and 0 < i - j <= 8192:
                m = 3
                while i + m < n and m < 264 and data[j + m] == data[i + m]:
                    m += 1
                match_len, dist = m, i - j - 1
        if match_len >= 3:
            flush()
            ln = match_len - 2
            if ln < 7:
                out.append((ln << 5) | (dist >> 8))
            else:
                out.append((7 << 5) | (dist >> 8))
                out.append(ln - 7)
            out.append(dist & 0xFF)
            i += match_len
        else:
            lit.append(data[i])
            i += 1
    flush()
    return bytes(out)


def write_pcd_compressed(path, xyz, normals, rgb):
    """DATA binary_compressed: uint32 compressed size, uint32 raw size, LZF stream of the structure-of-arrays payload."""
    n = xyz.shape[0]
    head = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE 4 4 4 4 4 4 4 4\n"
            "TYPE F F F U F F F F\nCOUNT 1 1 1 1 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\n"
            "DATA binary_compressed\n" % (_FIELDS, n, n))
    cols = [xyz[:, 0], xyz[:, 1], xyz[:, 2], rgb.astype(np.uint32), normals[:, 0], normals[:, 1], normals[:, 2],
            np.zeros(n, np.float32)]
    raw = b"".join(np.ascontiguousarray(c, dtype=(np.uint32 if c.dtype == np.uint32 else np.float32)).tobytes() for c in cols)
    comp = lzf_compress(raw)
    with open(path, "wb") as f:
        f.write(head.encode())
        f.write(np.array([len(comp), len(raw)], np.uint32).tobytes())
        f.write(comp)


def write_ply(path, xyz, normals, rgb, fmt="binary_little_endian", with_faces=False):
    """PLY vertex cloud the way MeshLab / pcl::io::savePLYFile lay it out: x y z nx ny nz red green blue."""
    n = xyz.shape[0]
    head = "ply\nformat %s 1.0\ncomment test cloud\nelement vertex %d\n" % (fmt, n)
    head += "property float x\nproperty float y\nproperty float z\n"
    if normals is not None:
        head += "property float nx\nproperty float ny\nproperty float nz\n"
    head += "property uchar red\nproperty uchar green\nproperty uchar blue\n"
    if with_faces:
        head += "element face 1\nproperty list uchar int vertex_indices\n"
    head += "end_header\n"
    r, g, b = (rgb >> 16) & 255, (rgb >> 8) & 255, rgb & 255
    with open(path, "wb") as f:
        f.write(head.encode())
        if fmt == "ascii":
            for i in range(n):
                vals = ["%.9g" % v for v in xyz[i]]
                if normals is not None:
                    vals += ["%.9g" % v for v in normals[i]]
                vals += [str(int(r[i])), str(int(g[i])), str(int(b[i]))]
                f.write((" ".join(vals) + "\n").encode())
            if with_faces:
                f.write(b"3 0 1 2\n")
        else:
            e = "<" if fmt == "binary_little_endian" else ">"
            fields = [("xyz", e + "f4", 3)] + ([("n", e + "f4", 3)] if normals is not None else []) + [("c", "u1", 3)]
            rec = np.zeros(n, dtype=fields)
            rec["xyz"] = xyz
            if normals is not None:
                rec["n"] = normals
            rec["c"] = np.stack([r, g, b], 1).astype(np.uint8)
            f.write(rec.tobytes())
            if with_faces:
                f.write(b"\x03" + np.array([0, 1, 2], e + "i4").tobytes())
