"""Seeded synthetic clouds of the shapes BASELINE.json names.

A class = a fixed random "prototype" (union of 4-8 primitives inside the unit cube); a cloud = P points
sampled by area on the prototype + Gaussian jitter + a random rigid rotation, with analytic outward normals
(so the normals stage is bypassed, as with add_normals-preprocessed data) and a class-seeded colour field.
Host-side numpy only: this is input generation, not part of the measured path.
"""
import numpy as np

from .structs import (MAXFILTER_SIMPLE, DIST_CHISQUARED, DIST_EUCLIDEAN, FEATURE_CSHOT, FEATURE_SHOT, default_params)


def _rand_rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


class Prototype:
    def __init__(self, class_id, seed=0):
        rng = np.random.default_rng([seed, class_id, 7919])
        self.class_id = class_id
        n = int(rng.integers(4, 9))
        self.prims = []
        for _ in range(n):
            kind = ("sphere", "box", "cylinder", "torus")[int(rng.integers(0, 4))]
            c = rng.uniform(-0.28, 0.28, 3)
            R = _rand_rot(rng)
            if kind == "sphere":
                p = (rng.uniform(0.08, 0.22),)
                area = 4 * np.pi * p[0] ** 2
            elif kind == "box":
                p = tuple(rng.uniform(0.06, 0.2, 3))
                area = 8 * (p[0] * p[1] + p[1] * p[2] + p[0] * p[2])
            elif kind == "cylinder":
                p = (rng.uniform(0.05, 0.15), rng.uniform(0.08, 0.25))
                area = 4 * np.pi * p[0] * p[1] + 2 * np.pi * p[0] ** 2
            else:
                R_ = rng.uniform(0.1, 0.2)
                p = (R_, rng.uniform(0.03, 0.45 * R_))
                area = 4 * np.pi ** 2 * p[0] * p[1]
            self.prims.append((kind, c, R, p, area))
        a = np.array([pr[4] for pr in self.prims])
        self.prob = a / a.sum()
        self.col_A = rng.normal(scale=4.0, size=(3, 3))
        self.col_phi = rng.uniform(0, 2 * np.pi, 3)

    @staticmethod
    def _sample_prim(kind, p, n, rng):
        if kind == "sphere":
            d = rng.normal(size=(n, 3))
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            return d * p[0], d
        if kind == "box":
            a, b, c = p
            fa = np.array([b * c, b * c, a * c, a * c, a * b, a * b])
            face = rng.choice(6, size=n, p=fa / fa.sum())
            u = rng.uniform(-1, 1, (n, 3)) * np.array([a, b, c])
            nrm = np.zeros((n, 3))
            ax = face // 2
            sg = np.where(face % 2 == 0, 1.0, -1.0)
            u[np.arange(n), ax] = sg * np.array([a, b, c])[ax]
            nrm[np.arange(n), ax] = sg
            return u, nrm
        if kind == "cylinder":
            r, h = p
            side = 4 * np.pi * r * h
            cap = np.pi * r * r
            which = rng.choice(3, size=n, p=np.array([side, cap, cap]) / (side + 2 * cap))
            th = rng.uniform(0, 2 * np.pi, n)
            pts = np.zeros((n, 3))
            nrm = np.zeros((n, 3))
            s = which == 0
            pts[s] = np.stack([r * np.cos(th[s]), r * np.sin(th[s]), rng.uniform(-h, h, s.sum())], 1)
            nrm[s] = np.stack([np.cos(th[s]), np.sin(th[s]), np.zeros(s.sum())], 1)
            for w, sg in ((1, 1.0), (2, -1.0)):
                m = which == w
                rr = r * np.sqrt(rng.uniform(0, 1, m.sum()))
                pts[m] = np.stack([rr * np.cos(th[m]), rr * np.sin(th[m]), np.full(m.sum(), sg * h)], 1)
                nrm[m, 2] = sg
            return pts, nrm
        R_, r = p
        out_p, out_n = [], []
        need = n
        while need > 0:
            m = int(need * 1.6) + 8
            u = rng.uniform(0, 2 * np.pi, m)
            v = rng.uniform(0, 2 * np.pi, m)
            keep = rng.uniform(0, 1, m) < (R_ + r * np.cos(v)) / (R_ + r)
            u, v = u[keep][:need], v[keep][:need]
            out_p.append(np.stack([(R_ + r * np.cos(v)) * np.cos(u), (R_ + r * np.cos(v)) * np.sin(u), r * np.sin(v)], 1))
            out_n.append(np.stack([np.cos(v) * np.cos(u), np.cos(v) * np.sin(u), np.sin(v)], 1))
            need -= len(u)
        return np.concatenate(out_p), np.concatenate(out_n)

    def sample(self, P, rng, jitter=0.002, rotate=True, scale=1.0, color_noise=0.03):
        which = rng.choice(len(self.prims), size=P, p=self.prob)
        pts = np.zeros((P, 3))
        nrm = np.zeros((P, 3))
        for i, (kind, c, R, p, _) in enumerate(self.prims):
            m = np.nonzero(which == i)[0]
            if len(m) == 0:
                continue
            lp, ln = self._sample_prim(kind, p, len(m), rng)
            pts[m] = lp @ R.T + c
            nrm[m] = ln @ R.T
        col = 0.5 + 0.5 * np.sin(pts @ self.col_A.T + self.col_phi)
        col = np.clip(col + rng.normal(scale=color_noise, size=col.shape), 0, 1)
        rgb8 = (col * 255.0).astype(np.uint32)
        rgb = (rgb8[:, 0] << 16) | (rgb8[:, 1] << 8) | rgb8[:, 2]
        pts = pts + rng.normal(scale=jitter, size=pts.shape)
        if rotate:
            G = _rand_rot(rng)
            pts = pts @ G.T
            nrm = nrm @ G.T
        return (pts * scale).astype(np.float32), nrm.astype(np.float32), rgb.astype(np.uint32)


def make_clouds(class_ids, seeds, P, proto_seed=0, scale=1.0, jitter=0.002, rotate=True):
    """Concatenated clouds: xyz (n,3) f32, normals (n,3) f32, rgb (n,) u32, cloud_off (B+1,) i64."""
    protos = {}
    xs, ns, cs = [], [], []
    off = [0]
    for cid, sd in zip(class_ids, seeds):
        if cid not in protos:
            protos[cid] = Prototype(cid, proto_seed)
        x, n, c = protos[cid].sample(P, np.random.default_rng([int(sd), 104729]), jitter=jitter, rotate=rotate,
                                     scale=scale)
        xs.append(x)
        ns.append(n)
        cs.append(c)
        off.append(off[-1] + P)
    return np.concatenate(xs), np.concatenate(ns), np.concatenate(cs), np.asarray(off, np.int64)


def make_scene(class_ids, seed, P_obj, spacing=2.5, plane_points=6000, clutter_points=1500, scale=1.0, jitter=0.002):
    """C5-shaped cluttered scene: object instances on a grid above a table plane plus uniform clutter.
    Returns xyz, normals, rgb and the ground truth (class id, instance centre) per object."""
    rng = np.random.default_rng([int(seed), 15485863])
    xs, ns, cs, truth = [], [], [], []
    side = int(np.ceil(np.sqrt(len(class_ids))))
    for i, cid in enumerate(class_ids):
        x, n, c = Prototype(cid, 0).sample(P_obj, np.random.default_rng([int(seed) + i, 104729]), jitter=jitter,
                                           rotate=True, scale=scale)
        shift = np.array([(i % side) * spacing, (i // side) * spacing, 0.0], np.float32) * scale
        lo, hi = x.min(0), x.max(0)
        shift[2] = -lo[2]  # rest on the table z = 0
        xs.append((x + shift).astype(np.float32))
        ns.append(n)
        cs.append(c)
        truth.append((cid, (0.5 * (lo + hi) + shift).astype(np.float32)))
    ext = side * spacing * scale
    pl = np.zeros((plane_points, 3), np.float32)
    pl[:, :2] = rng.uniform(-0.5 * spacing * scale, ext, (plane_points, 2))
    pl[:, 2] = rng.normal(0, jitter * scale, plane_points)
    xs.append(pl)
    ns.append(np.tile(np.array([0, 0, 1], np.float32), (plane_points, 1)))
    cs.append(np.full(plane_points, 0x808080, np.uint32))
    cl = rng.uniform([-0.5 * spacing * scale] * 2 + [0], [ext, ext, scale], (clutter_points, 3)).astype(np.float32)
    cn = rng.normal(size=(clutter_points, 3)).astype(np.float32)
    cn /= np.linalg.norm(cn, axis=1, keepdims=True)
    xs.append(cl)
    ns.append(cn)
    cs.append(rng.integers(0, 1 << 24, clutter_points).astype(np.uint32))
    return np.concatenate(xs), np.concatenate(ns), np.concatenate(cs), truth


# ---- workload definitions (configs of BASELINE.json) -------------------------------------
WORKLOADS = {
    # C1: quick-start stand-in; radii of config/qs_input_config.ism (mm-like scale x250)
    "c1": dict(n_classes=5, P=5000, scale=250.0, train_per_class=1, n_test=5, cshot=False,
               radius=60.0, lrf_radius=50.0, leaf=50.0, bandwidth=50.0, dist=DIST_CHISQUARED, n_words=None),
    # C2: ModelNet10-shaped
    # radii as shipped in config/default.ism (Radius 0.40, ReferenceFrameRadius 0.30) on unit-size objects;
    # LeafSize 0.08 gives ~256 keypoints per 2k-point cloud, Bandwidth 0.30
    "c2": dict(n_classes=10, P=2048, scale=1.0, train_per_class=None, n_test=1000, cshot=False,
               radius=0.40, lrf_radius=0.30, leaf=0.08, bandwidth=0.30, dist=DIST_EUCLIDEAN, n_words=200_000),
    # C3: ModelNet40-shaped
    "c3": dict(n_classes=40, P=2048, scale=1.0, train_per_class=None, n_test=4096, cshot=False,
               radius=0.40, lrf_radius=0.30, leaf=0.08, bandwidth=0.30, dist=DIST_EUCLIDEAN, n_words=1_000_000),
    # C4: Washington-shaped CSHOT (default_config_kinect.ism radii at object scale 0.2 m)
    "c4": dict(n_classes=51, P=8192, scale=0.2, train_per_class=None, n_test=512, cshot=True,
               radius=0.05, lrf_radius=0.05, leaf=0.02, bandwidth=0.045, dist=DIST_EUCLIDEAN, n_words=1_000_000),
    # C5: cluttered-scene localisation: one ~300k-point scene per step (25 objects of 8192 points on a table plane plus
    # clutter), about 2.5e4 keypoints in one cloud, multi-class maxima (SingleObjectMode=false, cross-class filter)
    "c5": dict(n_classes=10, P=8192, scale=1.0, train_per_class=None, n_test=8, cshot=False,
               radius=0.40, lrf_radius=0.30, leaf=0.08, bandwidth=0.30, dist=DIST_EUCLIDEAN, n_words=200_000,
               scene_objects=25, plane_points=80_000, clutter_points=15_000),
}


def workload_params(name, **over):
    w = WORKLOADS[name]
    p = default_params(
        feature_type=FEATURE_CSHOT if w["cshot"] else FEATURE_SHOT,
        feature_radius=w["radius"], lrf_radius=w["lrf_radius"], leaf_size=w["leaf"], bandwidth=w["bandwidth"],
        distance_type=w["dist"], knn_k=1, average_rotation=1, single_object_mode=1,
    )
    if "scene_objects" in w:
        p.single_object_mode = 0
        p.min_votes_threshold = 5
        p.max_filter_type = MAXFILTER_SIMPLE
    for k, v in over.items():
        setattr(p, k, v)
    return p
