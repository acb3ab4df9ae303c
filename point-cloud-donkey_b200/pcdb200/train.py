"""Training-side driver: ImplicitShapeModel::train + Codebook::activate on top of the CUDA library.

Reference: src/implicit_shape_model/implicit_shape_model.cpp:252-500 (per-cloud features, AABB, feature order),
codebook/codebook.cpp:64-368 (activation, class variances, k=1 clean-up), codebook/codeword_distribution.cpp:37-71
(vote = R_train (bbox centre - keypoint), bbox quaternion), :171-243 (per-vote weights).  Clustering "None" and
feature ranking "Uniform" (all shipped configs).

Feature extraction, the N x N activation (the same tcgen05 activation as detection — what makes a 1M-word codebook
trainable in seconds) and the pairwise functor distances run on the GPU through the C-ABI; what is left here is the
reference's bookkeeping over std::map, restated with float32 numpy in the reference's operation order.
"""
import numpy as np

from .structs import Codebook

F32 = np.float32


def aabb(xyz):
    """Utils::computeAABB (utils/utils.cpp:221-233): pos(3), quat wxyz(4), size(3)."""
    mn, mx = xyz.min(0).astype(F32), xyz.max(0).astype(F32)
    size = mx - mn
    pos = mn + size / F32(2)
    return np.concatenate([pos, np.array([1, 0, 0, 0], F32), size]).astype(F32)


# ---- boost::math::quaternion<float> arithmetic, vectorised over rows (w, x, y, z) -------------------------------------
def qmul(l, r):
    a, b, c, d = l[..., 0], l[..., 1], l[..., 2], l[..., 3]
    ar, br, cr, dr = r[..., 0], r[..., 1], r[..., 2], r[..., 3]
    return np.stack([
        a * ar - b * br - c * cr - d * dr,
        a * br + b * ar + c * dr - d * cr,
        a * cr - b * dr + c * ar + d * br,
        a * dr + b * cr - c * br + d * ar,
    ], -1).astype(F32)


def qconj(q):
    return (q * np.array([1, -1, -1, -1], F32)).astype(F32)


def lrf_quat(rf):
    """Utils::getRotQuaternion + matrix2Quat (utils/utils.cpp:136-152,342-394): rows of the matrix = LRF axes."""
    m = rf.reshape(-1, 3, 3).astype(F32)
    n = m.shape[0]
    quat = np.zeros((n, 4), F32)  # x y z w
    trace = (m[:, 0, 0] + m[:, 1, 1]) + m[:, 2, 2]
    pos = trace > 0
    with np.errstate(invalid="ignore", divide="ignore"):
        root = np.sqrt(trace + F32(1.0)).astype(F32)
        w = F32(0.5) * root
        r2 = F32(0.5) / root
        qa = np.stack([(m[:, 2, 1] - m[:, 1, 2]) * r2, (m[:, 0, 2] - m[:, 2, 0]) * r2, (m[:, 1, 0] - m[:, 0, 1]) * r2, w], 1)
    quat[pos] = qa[pos]
    idx = np.nonzero(~pos)[0]
    for t in idx:  # rare branch: plain scalar code
        mm = m[t]
        i = 0
        if mm[1, 1] > mm[0, 0]:
            i = 1
        if mm[2, 2] > mm[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        arg = F32(np.float64((mm[i, i] - mm[j, j]) - mm[k, k]) + 1.0)
        root = np.sqrt(arg).astype(F32)
        q = np.zeros(4, F32)
        q[i] = F32(0.5) * root
        root = F32(0.5) / root
        q[3] = (mm[k, j] - mm[j, k]) * root
        q[j] = (mm[j, i] + mm[i, j]) * root
        q[k] = (mm[k, i] + mm[i, k]) * root
        quat[t] = q
    return np.stack([quat[:, 3], quat[:, 0], quat[:, 1], quat[:, 2]], 1).astype(F32)


def quat_rotate(q, p):
    """q p q*  (Utils::rotateInto)."""
    pq = np.concatenate([np.zeros((len(p), 1), F32), p.astype(F32)], 1)
    return qmul(qmul(q, pq), qconj(q))[:, 1:]


def quat_rotate_inv(q, p):
    """q* p q  (Utils::rotateBack)."""
    pq = np.concatenate([np.zeros((len(p), 1), F32), p.astype(F32)], 1)
    return qmul(qmul(qconj(q), pq), q)[:, 1:]


def _seq_sum_f32(x):
    return np.cumsum(x.astype(F32), dtype=F32)[-1] if len(x) else F32(0)


def extract_features(ctx, xyz, normals, rgb, cloud_off, batch=128):
    """computeFeatures for many clouds, batched through pcdb_compute_features."""
    B = len(cloud_off) - 1
    fx, fl, fd, counts = [], [], [], []
    for b0 in range(0, B, batch):
        b1 = min(B, b0 + batch)
        s, e = int(cloud_off[b0]), int(cloud_off[b1])
        a = ctx.compute_features(xyz[s:e], normals[s:e], None if rgb is None else rgb[s:e],
                                 cloud_off[b0:b1 + 1] - cloud_off[b0])
        fx.append(a[0])
        fl.append(a[1])
        fd.append(a[2])
        counts.append(np.diff(a[3]))
    off = np.concatenate([[0], np.cumsum(np.concatenate(counts))]).astype(np.int64)
    return np.concatenate(fx), np.concatenate(fl), np.concatenate(fd), off


def train_codebook(ctx, prm, feat_xyz, feat_lrf, feat_desc, feat_off, cloud_class, cloud_instance, cloud_bbox,
                   n_classes, knn_mode=0):
    """Codebook from training features (class-major clouds).  Returns pcdb200.structs.Codebook."""
    feat_xyz, feat_lrf = feat_xyz.astype(F32), feat_lrf.astype(F32)
    feat_desc = np.ascontiguousarray(feat_desc, F32)
    cloud_class = np.asarray(cloud_class, np.uint32)
    cloud_instance = np.asarray(cloud_instance, np.uint32)
    cloud_bbox = np.asarray(cloud_bbox, F32).reshape(-1, 10)
    if (np.diff(cloud_class.astype(np.int64)) < 0).any():
        raise ValueError("training clouds must be class-major (the reference iterates a std::map by class id)")
    n_clouds = len(feat_off) - 1
    Fn, D = feat_desc.shape
    k = prm.knn_k
    feat_cloud = np.repeat(np.arange(n_clouds), np.diff(feat_off))

    # codewords = features (clustering None, implicit_shape_model.cpp:447-475); activation against all of them
    tmp = Codebook(feat_desc, np.arange(Fn + 1), np.zeros((Fn, 3)), np.ones(Fn), np.zeros(Fn), np.zeros(Fn),
                   np.zeros((Fn, 7)), np.ones(Fn), feat_xyz, np.arange(Fn), np.ones(n_classes))
    tprm = prm.copy()
    tprm.use_distance_ratio = 0  # detection-only (activation_strategy_knn.h:67)
    ctx.set_params(tprm)
    ctx.set_codebook(tmp)
    idx, dist, cnt = ctx.knn(feat_desc, k=k, dist_type=prm.distance_type, mode=knn_mode)
    ctx.set_params(prm)

    # CodewordDistribution::addCodeword for every (feature, activated codeword), in feature order
    fi = np.repeat(np.arange(Fn), cnt)                        # activating feature
    first = np.concatenate([[0], np.cumsum(cnt)[:-1]]) if Fn else np.zeros(0, np.int64)
    rank = np.arange(int(cnt.sum())) - np.repeat(first, cnt)
    cw = idx[fi, rank].astype(np.int64)                       # activated codeword (== row of the temporary codebook)
    centre = cloud_bbox[feat_cloud[fi], :3]
    vote = (centre - feat_xyz[fi]).astype(F32)
    rq = lrf_quat(feat_lrf[fi])
    vote_rot = quat_rotate(rq, vote)
    bq = qmul(cloud_bbox[feat_cloud[fi], 3:7], qconj(rq))
    vbbox = np.concatenate([bq, cloud_bbox[feat_cloud[fi], 7:10]], 1).astype(F32)

    # class variances (codebook.cpp:94-193)
    sigma2 = np.ones(n_classes, F32)
    for cls in np.unique(cloud_class):
        clouds = np.nonzero(cloud_class == cls)[0]
        num_features = int(feat_off[clouds[-1] + 1] - feat_off[clouds[0]])
        max_elements = int(np.sqrt(num_features))
        # "if (all.size() < max_elements) append": whole features / whole models are appended while the list is short
        f0, f1 = int(feat_off[clouds[0]]), int(feat_off[clouds[-1] + 1])
        before = np.concatenate([[0], np.cumsum(cnt[f0:f1])[:-1]]) if f1 > f0 else np.zeros(0, np.int64)
        take = np.nonzero(before < max_elements)[0] + f0
        activated = [int(r) for f in take for r in idx[f, :cnt[f]]]
        model_feats = []
        for c in clouds:
            if len(model_feats) < max_elements:
                model_feats.extend(range(int(feat_off[c]), int(feat_off[c + 1])))
        if not model_feats or not activated:
            continue
        a = np.repeat(np.asarray(model_feats), len(activated))
        b = np.tile(np.asarray(activated), len(model_feats))
        d = np.concatenate([ctx.distance_pairs(feat_desc[a[s:s + 65536]], feat_desc[b[s:s + 65536]], prm.distance_type)
                            for s in range(0, len(a), 65536)])
        num = len(a)
        mean = _seq_sum_f32(d) / F32(num)
        diff = (d - mean).astype(F32)
        with np.errstate(divide="ignore", invalid="ignore"):
            var = _seq_sum_f32(diff * diff) / F32(num - 1)
        if cls < n_classes:
            sigma2[cls] = var

    # group by codeword id (std::map order); KNN k == 1 keeps only codewords with exactly one vote (:201-224)
    order = np.argsort(cw, kind="stable")
    cw_s = cw[order]
    ids, start, count = np.unique(cw_s, return_index=True, return_counts=True)
    if k == 1:
        keep = count == 1
        ids, start, count = ids[keep], start[keep], count[keep]
    grp_first = np.concatenate([[0], np.cumsum(count)[:-1]]) if len(ids) else np.zeros(0, np.int64)
    sel = order[np.repeat(start, count) + (np.arange(int(count.sum())) - np.repeat(grp_first, count))]
    vote_off = np.concatenate([[0], np.cumsum(count)]).astype(np.int64)

    # computeWeights (:171-243): median over the activating features of exp(-|centre' - centre|^2 / 0.25)
    vweight = np.ones(len(sel), F32)
    sigma = F32(0.5)
    single = np.repeat(count == 1, count)
    if len(sel):
        back = feat_xyz[fi[sel]] + quat_rotate_inv(rq[sel], vote_rot[sel])
        dd = back - centre[sel]
        dist1 = np.sqrt((dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]) + dd[:, 2] * dd[:, 2]).astype(F32)
        vweight[single] = np.exp((F32(-1) * (dist1 * dist1)) / (sigma * sigma)).astype(F32)[single]
        for w_i in np.nonzero(count > 1)[0]:  # k > 1: the general median
            rows = sel[vote_off[w_i]:vote_off[w_i + 1]]
            for vi, r in enumerate(rows):
                c2 = feat_xyz[fi[rows]] + quat_rotate_inv(rq[rows], np.tile(vote_rot[r], (len(rows), 1)))
                dj = c2 - centre[r]
                dn = np.sqrt((dj[:, 0] * dj[:, 0] + dj[:, 1] * dj[:, 1]) + dj[:, 2] * dj[:, 2]).astype(F32)
                lw = np.sort(np.exp((F32(-1) * (dn * dn)) / (sigma * sigma)).astype(F32))
                m = len(lw)
                med = (lw[m // 2 - 1] + lw[m // 2]) / F32(2) if m % 2 == 0 else lw[m // 2]
                vweight[vote_off[w_i] + vi] = med
    return Codebook(
        words=feat_desc[ids], vote_off=vote_off, vote_xyz=vote_rot[sel], vote_weight=vweight,
        vote_class=cloud_class[feat_cloud[fi[sel]]], vote_instance=cloud_instance[feat_cloud[fi[sel]]],
        vote_bbox=vbbox[sel], vote_class_weight=np.ones(len(sel), F32), kp_train=feat_xyz[ids],
        codeword_ids=ids.astype(np.int32), sigma2=sigma2)
