"""CPU oracle for SURVEY.md section 8(f) row N1: ``Solver.clustering`` (utils/modeler.py:762-899) --
C-alpha candidates from the stitched volumes (:767-860) and their neighbour graph (:862-899).

TEST INFRASTRUCTURE ONLY (same rule as oracle/mica_oracle.py: nothing under ``mica_b200/``
imports this; only tests/, smoke() and bench.py's CPU legs do, as the checker).

Each function restates one block of the reference method in NumPy with the logging removed.
Pinning: oracle/make_golden.py runs the UNMODIFIED ``Solver.clustering`` through
oracle/ref_harness.py::solver_clustering and asserts that ``ca_candidates`` below reproduces
``CA_cands``, ``CA_cands_AAProb``, ``CA_cands_AA`` and ``CAProb_clusted`` bit for bit; the
reference's outputs are committed as tests/golden/candidates.npz.

Two things are NOT pinned by the reference alone and are said so here:
* DBSCAN is Open3D's (utils/modeler.py:768-770), absent from the image.  ``dbscan`` restates
  the published algorithm with Open3D's conventions and is checked against scikit-learn's
  implementation (same conventions) -- "parity unpinned" against Open3D itself.
* ``np.argsort(-scores)`` (utils/modeler.py:815) is an unstable sort: the order of voxels
  with EQUAL probability is whatever NumPy's build does.  The oracle (and the CUDA path)
  breaks ties by the ``np.where`` order of the voxels, i.e. a stable sort; on tie-free inputs
  this is the reference's order exactly, and the golden fixture is tie-free among the
  candidates that matter (asserted when it is generated).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


# utils/modeler.py:767
def threshold_points(ca_prob, thr):
    """``np.array(np.where(CAProb > thr)).T``: int64 [n,3] in C (x-major) order.  ``thr`` is a
    python float, so NumPy 2 compares in float32."""
    return np.array(np.where(ca_prob > thr)).T


# utils/modeler.py:768-770 (Open3D PointCloud.cluster_dbscan)
def dbscan(points, eps, min_points):
    """DBSCAN with Open3D's conventions: closed eps-ball that includes the point itself, core
    iff the ball holds >= min_points points, clusters numbered by their first core point in
    input order and grown to completion one at a time, a border point keeps the label of the
    first cluster that reaches it, noise = -1."""
    from scipy.spatial import cKDTree
    pts = np.asarray(points, dtype=np.float64)
    n = len(pts)
    labels = np.full(n, -2, dtype=np.int64)
    if n == 0:
        return labels
    nbs = cKDTree(pts).query_ball_point(pts, float(eps))
    core = np.array([len(v) >= min_points for v in nbs])
    cluster = 0
    for i in range(n):
        if labels[i] != -2:
            continue
        if not core[i]:
            labels[i] = -1
            continue
        labels[i] = cluster
        stack = list(nbs[i])
        while stack:
            q = stack.pop()
            if labels[q] == -1:
                labels[q] = cluster
            if labels[q] != -2:
                continue
            labels[q] = cluster
            if core[q]:
                stack.extend(nbs[q])
        cluster += 1
    return labels


# utils/modeler.py:775-797
def valid_clusters(points, labels, bb_prob):
    """Boolean mask over the points: clusters whose summed backbone probability exceeds a
    tenth of the best sum get their mean as score, the others 0; clusters scoring above half
    the best score are valid."""
    labels = np.asarray(labels)
    n_lab = int(labels.max()) + 1 if len(labels) else 0
    sums = []
    for lab in range(n_lab):
        p = points[np.where(labels == lab)]
        sums.append(np.sum(bb_prob[p[:, 0], p[:, 1], p[:, 2]]))
    avgs = []
    for lab in range(n_lab):
        if sums[lab] > np.max(sums) / 10:
            p = points[np.where(labels == lab)]
            avgs.append(np.mean(bb_prob[p[:, 0], p[:, 1], p[:, 2]]))
        else:
            avgs.append(0)
    val = np.zeros_like(labels).astype(bool)
    best = np.max(avgs)
    for lab in range(n_lab):
        if avgs[lab] > best / 2:
            val[np.where(labels == lab)] = True
    return val, np.asarray(sums, dtype=np.float32), np.asarray(avgs, dtype=np.float32)


# utils/modeler.py:800-802
def clustered_volume(ca_prob, points, val):
    out = np.zeros_like(ca_prob)
    c = points[np.where(val)]
    out[c[:, 0], c[:, 1], c[:, 2]] = ca_prob[c[:, 0], c[:, 1], c[:, 2]]
    return out


# utils/modeler.py:805-832
def nms(points, scores, nms_radius, thr):
    """Greedy non-maximum suppression over the valid points, best score first; a pick deletes
    every point whose SQUARED distance is <= nms_radius (the reference compares the squared
    distance with ``nms_radius`` itself, :827-829).  Stable order among equal scores."""
    order = np.argsort(-scores.astype(np.float64), kind='stable')
    pts = points[order].astype(np.float64)
    sc = scores[order].astype(np.float64)
    picks = []
    while len(pts) > 0 and sc[0] >= thr:
        picks.append([int(pts[0, 0]), int(pts[0, 1]), int(pts[0, 2])])
        d2 = (pts[:, 0] - pts[0, 0]) ** 2 + (pts[:, 1] - pts[0, 1]) ** 2 + (pts[:, 2] - pts[0, 2]) ** 2
        keep = d2 > nms_radius
        pts, sc = pts[keep], sc[keep]
    return np.asarray(picks, dtype=np.int64).reshape(-1, 3)


def pairwise_sum_27(a):
    """``np.sum`` of a 3x3x3 float32 view = NumPy's pairwise kernel on the 27 ravelled values:
    eight running sums over the first 24, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
    then the last three added one by one (checked against np.sum in tests/test_oracle_next.py)."""
    a = np.asarray(a, dtype=np.float32).ravel()
    r = [f32(a[j]) for j in range(8)]
    for i in (8, 16):
        for j in range(8):
            r[j] = f32(r[j] + a[i + j])
    res = f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])))
    for i in (24, 25, 26):
        res = f32(res + a[i])
    return res


# utils/modeler.py:837-860
def refine(cands, ca_prob, aa_prob, aa_pred):
    """3x3x3 probability-weighted centroid and amino-acid profile per pick.  A pick on the
    border of the volume raises inside the reference's try block (empty or short slice) and is
    skipped (:856-857).  Returns (CA_cands float64 [m,3], CA_cands_AAProb float32 [20,m],
    CA_cands_AA float32 [m], kept bool [len(cands)])."""
    X, Y, Z = ca_prob.shape
    new_c, new_a, kept = [], [], []
    for c in np.asarray(cands, dtype=np.int64):
        inside = (1 <= c[0] <= X - 2) and (1 <= c[1] <= Y - 2) and (1 <= c[2] <= Z - 2)
        kept.append(bool(inside))
        if not inside:
            continue
        win = ca_prob[c[0] - 1:c[0] + 2, c[1] - 1:c[1] + 2, c[2] - 1:c[2] + 2]
        w = win / pairwise_sum_27(win)                        # float32 / float32
        coord = np.zeros(3, dtype=np.float64)
        acc = None
        for di in (-1, 0, 1):
            for dj in (-1, 0, 1):
                for dk in (-1, 0, 1):
                    t = c + [di, dj, dk]
                    wk = w[di + 1, dj + 1, dk + 1]
                    coord = coord + t * wk                    # int64 * float32 -> float64
                    term = aa_prob[:, t[0], t[1], t[2]] * wk  # float32
                    acc = term if acc is None else acc + term  # np.sum(list, axis=0): row by row
        new_c.append(coord)
        new_a.append(acc)
    ca_cands = np.array(new_c, dtype=np.float64).reshape(-1, 3)
    aap = np.array(new_a, dtype=np.float32).reshape(-1, 20).T
    rc = np.round(ca_cands).astype(int)
    aa = aa_pred[rc[:, 0], rc[:, 1], rc[:, 2]] if len(rc) else np.zeros(0, aa_pred.dtype)
    return ca_cands, aap, aa, np.asarray(kept, dtype=bool)


def ca_candidates(ca_prob, bb_prob, aa_prob, aa_pred, ca_score_thrh=0.3, cluster_eps=10,
                  cluster_min_points=10, nms_radius=9, labels=None):
    """utils/modeler.py:767-860 end to end."""
    pts = threshold_points(ca_prob, ca_score_thrh)
    if labels is None:
        labels = dbscan(pts, cluster_eps, cluster_min_points)
    val, sums, avgs = valid_clusters(pts, labels, bb_prob)
    vp = pts[val]
    picks = nms(vp, ca_prob[vp[:, 0], vp[:, 1], vp[:, 2]], nms_radius, ca_score_thrh)
    ca_cands, aap, aa, kept = refine(picks, ca_prob, aa_prob, aa_pred)
    return dict(points=pts, labels=np.asarray(labels), valid=val, cluster_sums=sums, cluster_avgs=avgs,
                picks=picks, picks_kept=kept, CA_cands=ca_cands, CA_cands_AAProb=aap, CA_cands_AA=aa,
                CAProb_clusted=clustered_volume(ca_prob, pts, val))


# utils/modeler.py:862-899 (the neighbour graph the tracer starts from)
def neighbor_scores(ca_cands, bb_prob):
    """cand_self_dis (float64 [m,m]), neigh_mat (float64 [m,m]) and best_neigh."""
    m = len(ca_cands)
    diff = ca_cands[:, None, :] - ca_cands[None, :, :]
    dis = np.linalg.norm(diff, axis=2)
    neigh = np.zeros_like(dis)
    for a in range(m):
        for b in np.where((dis[a] <= 6) * (dis[a] >= 2))[0]:
            d = max(0, abs(dis[a, b] - 3.8) - 0.5)
            dis_score = max(0, 1 - d / 2)
            dens = 0
            for j in range(1, 5):
                c = np.round(j / 5 * ca_cands[b] + (5 - j) / 5 * ca_cands[a]).astype(int)
                dens += bb_prob[c[0], c[1], c[2]]
            neigh[a, b] = (dis_score + dens / 4) / 2
    lists = dict(neighbors2to6=[np.where((dis[a] <= 6) * (dis[a] >= 2))[0] for a in range(m)],
                 neighbors0to6=[np.where(dis[a] <= 6)[0] for a in range(m)],
                 neighbors0to7=[np.where(dis[a] <= 7)[0] for a in range(m)],
                 neighbors2to7=[np.where((dis[a] <= 7) * (dis[a] >= 2))[0] for a in range(m)])
    neighbor_scores.lists = lists                        # :866-873 (kept as an attribute: callers unpack three)
    best = []
    for a in range(m):
        lst = []
        second, first = neigh[a].argsort()[-2:]
        if neigh[a, first] != 0:
            lst.append(int(first))
        if neigh[a, second] != 0:
            lst.append(int(second))
        best.append(lst)
    return dis, neigh, best
