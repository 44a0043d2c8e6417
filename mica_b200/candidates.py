"""C-alpha candidates from the stitched volumes, on the device.

Drop-in for the head of the reference's ``Solver.clustering`` (utils/modeler.py:762-860):

    pcd_numpy = np.where(CAProb > thr)          :767      -> ordered compaction
    labels    = open3d cluster_dbscan(eps, m)   :768-770  -> lattice DBSCAN (or a caller-supplied function)
    cluster score filter                        :775-797  -> per-label sums on the device, decisions on the host
    NNPred.CAProb_clusted                       :800-802  -> produced on request
    greedy NMS, best probability first          :805-832  -> parallel greedy independent set
    3x3x3 weighted refinement, AA profile       :837-860  -> one warp per pick

The volumes stay in HBM where ``CryoEMPredictor`` stitched them; only the picks (a few thousand rows)
are copied to the host -- the 20-channel ``amino_acid_probability`` volume (8.8 GB at 480^3), whose only
consumer in the reference is line :850, never crosses PCIe.  There is no CPU path: host arrays are
uploaded, and everything below raises without a GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, ops
from ._lib import lib, check
from .ops import _dev, _stream, device_guard

VOLUME_KEYS = ('backbone_probability', 'carbon_alpha_probability', 'amino_acid_prediction',
               'amino_acid_probability')


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _scalar_i64(t: torch.Tensor) -> int:
    """Read a device int64 scalar (synchronises the current stream)."""
    return int(t.cpu().item())


def _as_device(v, device):
    if hasattr(v, 'tensor') and isinstance(getattr(v, 'tensor'), torch.Tensor):
        v = v.tensor                         # predict.DeviceVolume: the volume never left HBM
    if isinstance(v, torch.Tensor):
        if not v.is_cuda:
            ops.require_gpu()
            v = v.to(device)
        return v.contiguous().float() if v.dtype != torch.float32 else v.contiguous()
    ops.require_gpu()
    return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(device)


@device_guard
def threshold_points(vol: torch.Tensor, thr: float):
    """``np.where(vol > thr)`` (utils/modeler.py:767) -> (lin int64 [n], xyz int32 [n,3]) on the device, in
    NumPy's order.  ``thr`` is compared in float32 like NumPy 2 compares a python float with a float32 array."""
    p = _dev(vol, torch.float32, 'vol')
    X, Y, Z = (int(s) for s in vol.shape)
    n_vox = X * Y * Z
    dev = vol.device
    nbytes = lib.mica_cand_threshold_workspace_bytes(n_vox)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    t = float(np.float32(thr))
    check(lib.mica_cand_threshold_count(p, n_vox, t, _ptr(ws), nbytes, _ptr(count), _stream()), 'threshold_count')
    n = _scalar_i64(count)
    lin = torch.empty(n, dtype=torch.int64, device=dev)
    xyz = torch.empty((n, 3), dtype=torch.int32, device=dev)
    if n:
        check(lib.mica_cand_threshold_write(p, X, Y, Z, t, _ptr(ws), _ptr(lin), _ptr(xyz), n, _stream()),
              'threshold_write')
    return lin, xyz


@device_guard
def gather(vol: torch.Tensor, lin: torch.Tensor) -> torch.Tensor:
    out = torch.empty(lin.shape[0], dtype=torch.float32, device=vol.device)
    if lin.shape[0]:
        check(lib.mica_gather_f32(_dev(vol, torch.float32, 'vol'), _dev(lin, torch.int64, 'lin'), lin.shape[0],
                                  _ptr(out), _stream()), 'gather')
    return out


@device_guard
def dbscan_lattice(lin: torch.Tensor, shape_xyz, eps, min_points):
    """Open3D ``cluster_dbscan(eps, min_points)`` (utils/modeler.py:770) for distinct lattice points given as
    ascending linear indices.  Returns (labels int32 [n] on the device, number of clusters)."""
    X, Y, Z = (int(s) for s in shape_xyz)
    n = int(lin.shape[0])
    dev = lin.device
    labels = torch.empty(n, dtype=torch.int32, device=dev)
    ncl = torch.zeros(1, dtype=torch.int64, device=dev)
    if n == 0:
        return labels, 0
    eps_sq = int(np.floor(float(eps) * float(eps) + 1e-9))
    nbytes = lib.mica_dbscan_workspace_bytes(X, Y, Z, n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    check(lib.mica_dbscan_lattice(_dev(lin, torch.int64, 'lin'), n, X, Y, Z, eps_sq, int(min_points), _ptr(ws),
                                  nbytes, _ptr(labels), _ptr(ncl), _stream()), 'dbscan_lattice')
    return labels, _scalar_i64(ncl)


@device_guard
def cluster_scores(bb_at: torch.Tensor, labels: torch.Tensor, n_labels: int):
    """Per-label (sum float64, count int64) of the backbone probability at the points -> host arrays."""
    dev = bb_at.device
    sums = torch.zeros(max(n_labels, 1), dtype=torch.float64, device=dev)
    counts = torch.zeros(max(n_labels, 1), dtype=torch.int64, device=dev)
    check(lib.mica_cand_cluster_scores(_dev(bb_at, torch.float32, 'vals'), _dev(labels, torch.int32, 'labels'),
                                       bb_at.shape[0], n_labels, _ptr(sums), _ptr(counts), _stream()),
          'cluster_scores')
    return sums[:n_labels].cpu().numpy(), counts[:n_labels].cpu().numpy()


def valid_labels(sums, counts):
    """utils/modeler.py:781-797 on the per-label statistics: a cluster whose summed backbone probability is
    above a tenth of the best sum scores its mean (else 0); clusters above half the best score are valid.
    Raises ValueError where the reference's ``np.max`` of an empty list does (no cluster at all)."""
    sums32 = np.asarray(sums, dtype=np.float64).astype(np.float32)          # np.sum of float32 -> float32
    if len(sums32) == 0:
        raise ValueError('zero-size array to reduction operation maximum which has no identity')
    with np.errstate(invalid='ignore', divide='ignore'):
        means = (np.asarray(sums, dtype=np.float64) / np.maximum(counts, 1)).astype(np.float32)
    avgs = np.where(sums32 > sums32.max() / np.float32(10), means, np.float32(0))
    return avgs > avgs.max() / 2, sums32, avgs


class Candidates(dict):
    """Result of ``find_candidates``: the reference's attribute names as keys (host arrays) plus the
    device-side intermediates under ``device``."""
    __getattr__ = dict.__getitem__


def find_candidates(volumes, CA_score_thrh=0.3, cluster_eps=10, cluster_min_points=10, nms_radius=9,
                    labels_fn=None, device='cuda', want_clustered=False) -> Candidates:
    """utils/modeler.py:767-860.  ``volumes``: an ``ops.StitchedVolumes``, or a mapping with the four keys
    ``CryoEMPredictor.run_prediction`` returns (device tensors are used in place, host arrays uploaded).
    ``labels_fn(points int64 [n,3]) -> labels`` replaces the built-in lattice DBSCAN (e.g. to call Open3D).
    """
    ops.require_gpu()
    if isinstance(volumes, ops.StitchedVolumes):
        volumes = volumes.as_dict()
    dev = torch.device(device)
    for v in volumes.values():
        v = getattr(v, 'tensor', v)
        if isinstance(v, torch.Tensor) and v.is_cuda:
            dev = v.device
            break
    ca = _as_device(volumes['carbon_alpha_probability'], dev)
    bb = _as_device(volumes['backbone_probability'], dev)
    aap = _as_device(volumes['amino_acid_probability'], dev)
    aapred = _as_device(volumes['amino_acid_prediction'], dev)
    X, Y, Z = (int(s) for s in ca.shape)
    if tuple(bb.shape) != (X, Y, Z) or tuple(aapred.shape) != (X, Y, Z) or tuple(aap.shape) != (20, X, Y, Z):
        raise _lib.MicaError('volume shapes do not match')
    with torch.cuda.device(dev):
        lin, xyz = threshold_points(ca, CA_score_thrh)                                   # :767
        n = int(lin.shape[0])
        if labels_fn is None:                                                            # :768-770
            labels, n_labels = dbscan_lattice(lin, (X, Y, Z), cluster_eps, cluster_min_points)
        else:
            host_labels = np.asarray(labels_fn(xyz.cpu().numpy().astype(np.int64)), dtype=np.int32)
            labels = torch.from_numpy(host_labels).to(dev)
            n_labels = int(host_labels.max()) + 1 if n else 0
        if n == 0:
            raise ValueError('zero-size array to reduction operation maximum which has no identity')  # labels.max()
        bb_at = gather(bb, lin)                                                          # :779
        sums, counts = cluster_scores(bb_at, labels, n_labels)
        ok, sums32, avgs = valid_labels(sums, counts)                                    # :781-797
        label_ok = torch.from_numpy(ok.astype(np.uint8)).to(dev)
        valid = torch.empty(n, dtype=torch.uint8, device=dev)
        check(lib.mica_cand_valid_points(_ptr(labels), _ptr(label_ok), n, n_labels, _ptr(valid), _stream()),
              'valid_points')
        # :805-832
        work = torch.empty(X * Y * Z, dtype=torch.float32, device=dev)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        rounds = C.c_int(0)
        check(lib.mica_cand_nms(_ptr(ca), X, Y, Z, _ptr(lin), _ptr(valid), n, int(nms_radius), _ptr(work),
                                _ptr(flag), C.byref(rounds), _stream()), 'nms')
        cap = n
        ws_bytes = lib.mica_cand_picks_workspace_bytes(cap)
        picks_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        pick_lin = torch.empty(cap, dtype=torch.int64, device=dev)
        pick_xyz = torch.empty((cap, 3), dtype=torch.int32, device=dev)
        n_picks = torch.zeros(1, dtype=torch.int64, device=dev)
        check(lib.mica_cand_nms_picks(_ptr(work), Y, Z, _ptr(lin), _ptr(valid), n, cap, _ptr(picks_ws), ws_bytes,
                                      _ptr(n_picks), _ptr(pick_lin), _ptr(pick_xyz), _stream()), 'nms_picks')
        m = _scalar_i64(n_picks)
        pick_lin, pick_xyz = pick_lin[:m], pick_xyz[:m]
        # :837-860
        out_xyz = torch.zeros((m, 3), dtype=torch.float64, device=dev)
        out_aap = torch.zeros((m, 20), dtype=torch.float32, device=dev)
        out_aa = torch.zeros(m, dtype=torch.float32, device=dev)
        out_ok = torch.zeros(m, dtype=torch.uint8, device=dev)
        check(lib.mica_cand_refine(_ptr(ca), _ptr(aap), _ptr(aapred), X, Y, Z, _ptr(pick_lin), m, _ptr(out_xyz),
                                   _ptr(out_aap), _ptr(out_aa), _ptr(out_ok), _stream()), 'refine')
        kept = out_ok.cpu().numpy().astype(bool)
        res = Candidates(
            CA_cands=out_xyz.cpu().numpy()[kept],
            CA_cands_AAProb=np.ascontiguousarray(out_aap.cpu().numpy()[kept].T),
            CA_cands_AA=out_aa.cpu().numpy()[kept],
            picks=pick_xyz.cpu().numpy().astype(np.int64), picks_kept=kept,
            n_points=n, n_clusters=n_labels, cluster_sums=sums32, cluster_avgs=avgs, nms_rounds=int(rounds.value),
            device=dict(lin=lin, xyz=xyz, labels=labels, valid=valid, pick_lin=pick_lin, bb=bb))
        if want_clustered:                                                               # :800-802
            res['CAProb_clusted'] = clustered_volume(ca, lin, valid)
    return res


@device_guard
def clustered_volume(ca: torch.Tensor, lin: torch.Tensor, valid: torch.Tensor) -> torch.Tensor:
    """``NNPred.CAProb_clusted`` (utils/modeler.py:800-802) on the device."""
    out = torch.empty_like(ca)
    check(lib.mica_cand_clustered_volume(_dev(ca, torch.float32, 'ca'), ca.numel(), _ptr(lin), _ptr(valid),
                                         lin.shape[0], _ptr(out), _stream()), 'clustered_volume')
    return out


def neighbor_graph(ca_cands, bb, device='cuda'):
    """utils/modeler.py:862-897 for the refined picks ``ca_cands`` (float64 [m,3]) and the backbone volume
    ``bb`` (device tensor or host array).  Returns dict(cand_self_dis, neigh_mat: float64 [m,m] host arrays;
    neighbors2to6 / neighbors0to6 / neighbors0to7 / neighbors2to7: lists of index arrays; best_neigh: list of
    lists) -- the attributes ``Solver.clustering`` leaves behind."""
    ops.require_gpu()
    dev = bb.device if isinstance(bb, torch.Tensor) and bb.is_cuda else torch.device(device)
    bb = _as_device(bb, dev)
    X, Y, Z = (int(s) for s in bb.shape)
    ca_cands = np.ascontiguousarray(ca_cands, dtype=np.float64).reshape(-1, 3)
    m = len(ca_cands)
    if m < 2:
        raise ValueError('not enough values to unpack (expected 2)')      # `second, first = ...argsort()[-2:]`
    with torch.cuda.device(dev):
        xyz = torch.from_numpy(ca_cands).to(dev)
        dis = torch.empty((m, m), dtype=torch.float64, device=dev)
        neigh = torch.empty((m, m), dtype=torch.float64, device=dev)
        check(lib.mica_cand_neighbor_graph(_ptr(xyz), m, _ptr(bb), X, Y, Z, _ptr(dis), _ptr(neigh), _stream()),
              'neighbor_graph')
        best = torch.empty((m, 2), dtype=torch.int32, device=dev)
        check(lib.mica_cand_best_neighbors(_ptr(neigh), m, _ptr(best), _stream()), 'best_neighbors')
        cap = 64
        while True:
            idx = torch.empty((m, cap), dtype=torch.int32, device=dev)
            cnt = torch.empty(m, dtype=torch.int32, device=dev)
            check(lib.mica_cand_neighbor_lists(_ptr(dis), m, 7.0, cap, _ptr(idx), _ptr(cnt), _stream()),
                  'neighbor_lists')
            counts = cnt.cpu().numpy()
            if int(counts.max()) <= cap:
                break
            cap = int(counts.max())
        idx_h, dis_h, neigh_h = idx.cpu().numpy(), dis.cpu().numpy(), neigh.cpu().numpy()
    mask = np.arange(cap)[None, :] < counts[:, None]
    rows = np.broadcast_to(np.arange(m)[:, None], (m, cap))[mask]
    cols = idx_h[mask].astype(np.int64)
    d = dis_h[rows, cols]                                # distances of the listed (<= 7 A) pairs

    def split(sel):                                      # per-row index arrays of the selected (row, col) pairs
        per_row = np.bincount(rows[sel], minlength=m)
        return np.split(cols[sel], np.cumsum(per_row)[:-1])

    lists = dict(neighbors2to6=split((d <= 6) & (d >= 2)), neighbors0to6=split(d <= 6),      # :866-873
                 neighbors0to7=split(np.ones(len(d), bool)), neighbors2to7=split(d >= 2))
    best_h = best.cpu().numpy()
    best_neigh = [[int(v) for v in row if v >= 0] for row in best_h]
    return dict(cand_self_dis=dis_h, neigh_mat=neigh_h, best_neigh=best_neigh, **lists)


def clustering(solver, nnpred=None, volumes=None, labels_fn=None):
    """The whole of ``Solver.clustering`` (utils/modeler.py:762-899): ``clustering_head`` + the neighbour graph,
    setting every attribute the reference method sets."""
    res = clustering_head(solver, nnpred, volumes, labels_fn)
    bb = res['device']['bb']
    g = neighbor_graph(solver.CA_cands, bb)
    solver.cand_self_dis, solver.neigh_mat, solver.best_neigh = g['cand_self_dis'], g['neigh_mat'], g['best_neigh']
    for k in ('neighbors2to6', 'neighbors0to6', 'neighbors0to7', 'neighbors2to7'):
        if isinstance(getattr(solver, k, None), list):
            getattr(solver, k).extend(g[k])            # the reference appends to the lists Solver.__init__ made
        else:
            setattr(solver, k, g[k])
    return res


def clustering_head(solver, nnpred=None, volumes=None, labels_fn=None):
    """What a maintainer calls in place of utils/modeler.py:767-860 inside ``Solver.clustering``: reads the
    same configuration (``solver.cluster_eps``, ``cluster_min_points``, ``nms_radius``,
    ``modeling_config.CA_score_thrh``) and sets the same attributes (``CA_cands``, ``CA_cands_AAProb``,
    ``CA_cands_AA`` and, when ``nnpred`` is given, ``NNPred.CAProb_clusted``).  ``volumes`` defaults to the
    device-resident volumes the predictor registered for ``modeling_config.output_path``."""
    if volumes is None:
        import os
        from . import session
        reg = session.get(os.path.join(str(solver.modeling_config.output_path), 'results', 'device_volumes'))
        if reg is None:
            volumes = {'carbon_alpha_probability': solver.CAProb, 'backbone_probability': nnpred.BBProb,
                       'amino_acid_probability': nnpred.AAProb, 'amino_acid_prediction': solver.AAPred}
        else:
            volumes = reg
    res = find_candidates(volumes, CA_score_thrh=solver.modeling_config.CA_score_thrh,
                          cluster_eps=solver.cluster_eps, cluster_min_points=solver.cluster_min_points,
                          nms_radius=solver.nms_radius, labels_fn=labels_fn, want_clustered=nnpred is not None)
    solver.CA_cands = res['CA_cands']
    solver.CA_cands_AAProb = res['CA_cands_AAProb']
    solver.CA_cands_AA = res['CA_cands_AA']
    if nnpred is not None:
        nnpred.CAProb_clusted = res['CAProb_clusted'].cpu().numpy()
    return res
