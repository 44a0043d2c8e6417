"""GPU parity of the SURVEY 8(f) rows through the C ABI: N1 C-alpha candidates (utils/modeler.py:762-860),
N3 label masks (scripts_for_training_data/create_*_mask.py), N4 docking masks (utils/dock_in_map.py:248-364).
Index work is compared bit for bit with the oracle and with the golden outputs of the unmodified reference."""
import os
import types

import numpy as np
import pytest
import torch

from mica_b200 import synthetic
from oracle import candidates_oracle as co
from oracle import masks_oracle as mo

from _next_cases import CANDIDATE_CASES, DOCK_CASES, candidate_volumes, dock_structure, mask_case

pytestmark = pytest.mark.gpu


def _dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


# ------------------------------------------------------------------------------------------ N1
@pytest.mark.parametrize('shape', [(13, 7, 5), (33, 31, 30), (64, 64, 65), (1, 1, 4097)])
def test_threshold_points_is_np_where(cuda, shape):
    from mica_b200 import candidates as cd
    rng = np.random.default_rng(sum(shape))
    vol = rng.random(shape, dtype=np.float32)
    vol.ravel()[::97] = np.nan                                        # NaN > thr is False
    vol.ravel()[5::89] = np.float32(0.3)                              # == thr is not above it
    for thr in (0.3, 0.97, -1.0, 2.0):
        lin, xyz = cd.threshold_points(_dev(vol, cuda), thr)
        want = np.array(np.where(vol > thr)).T
        assert np.array_equal(xyz.cpu().numpy(), want.reshape(-1, 3))
        assert np.array_equal(lin.cpu().numpy(), np.flatnonzero(vol > thr))


def test_dbscan_lattice_matches_the_sequential_algorithm(cuda):
    from mica_b200 import candidates as cd
    rng = np.random.default_rng(5)
    for trial, (dens, eps, mp) in enumerate([(0.02, 3, 4), (0.01, 10, 10), (0.08, 2, 6), (0.004, 5, 3),
                                             (0.03, 2.5, 5), (0.002, 10, 2)]):
        shape = (30, 26, 34) if trial % 2 == 0 else (19, 40, 67)              # 67: rows span three 32-bit words
        occ = rng.random(shape) < dens
        occ[4:9, 5:9, 6:12] |= rng.random((5, 4, 6)) < 0.7
        pts = np.array(np.where(occ)).T
        want = co.dbscan(pts, eps, mp)
        lin = _dev(np.flatnonzero(occ).astype(np.int64), cuda)
        labels, n_clusters = cd.dbscan_lattice(lin, shape, eps, mp)
        assert np.array_equal(labels.cpu().numpy(), want), trial
        assert n_clusters == want.max() + 1


@pytest.mark.parametrize('n', range(len(CANDIDATE_CASES)))
def test_find_candidates_matches_reference_golden(cuda, golden_dir, n):
    from mica_b200 import candidates as cd
    g = np.load(os.path.join(golden_dir, 'candidates.npz'))
    p = candidate_volumes(CANDIDATE_CASES[n])
    vols = {k: _dev(v, cuda) for k, v in p.items()}
    res = cd.find_candidates(vols, want_clustered=True)
    assert np.array_equal(res['device']['labels'].cpu().numpy(), g[f'c{n}_labels'])    # lattice DBSCAN
    assert np.array_equal(res['picks'], g[f'c{n}_picks'])                              # NMS order
    assert np.array_equal(res['CA_cands'], g[f'c{n}_CA_cands'])                        # float64, bit for bit
    assert np.array_equal(res['CA_cands_AAProb'], g[f'c{n}_CA_cands_AAProb'])
    assert np.array_equal(res['CA_cands_AA'], g[f'c{n}_CA_cands_AA'])
    assert np.array_equal(np.flatnonzero(res['CAProb_clusted'].cpu().numpy()), g[f'c{n}_clusted_lin'])
    ca = p['carbon_alpha_probability']
    cl = res['CAProb_clusted'].cpu().numpy()
    assert np.array_equal(cl[cl != 0], ca[cl != 0])
    # a caller-supplied DBSCAN (Open3D in the reference) gives the same result
    res2 = cd.find_candidates(vols, labels_fn=lambda pts: co.dbscan(pts, 10, 10))
    assert np.array_equal(res2['CA_cands'], res['CA_cands'])
    # cluster scores: float64 sums on the device vs the reference's float32 pairwise sums
    o = co.ca_candidates(ca, p['backbone_probability'], p['amino_acid_probability'], p['amino_acid_prediction'])
    assert np.allclose(res['cluster_sums'], o['cluster_sums'], rtol=1e-5)
    assert np.allclose(res['cluster_avgs'], o['cluster_avgs'], rtol=1e-5)


def test_find_candidates_other_parameters_and_host_arrays(cuda):
    from mica_b200 import candidates as cd
    p = synthetic.synthetic_predictions((44, 40, 36), (40, 20), seed=9)
    kw = dict(ca_score_thrh=0.35, cluster_eps=8, cluster_min_points=6, nms_radius=6)
    o = co.ca_candidates(p['carbon_alpha_probability'], p['backbone_probability'], p['amino_acid_probability'],
                         p['amino_acid_prediction'], **kw)
    res = cd.find_candidates(p, CA_score_thrh=0.35, cluster_eps=8, cluster_min_points=6, nms_radius=6)  # host arrays
    for k in ('CA_cands', 'CA_cands_AAProb', 'CA_cands_AA', 'picks'):
        assert np.array_equal(res[k], o[k]), k


def test_nms_with_ties_is_the_stable_greedy_order(cuda):
    from mica_b200 import candidates as cd
    rng = np.random.default_rng(3)
    shape = (48, 40, 44)
    ca = (np.round(rng.random(shape), 2) * (rng.random(shape) < 0.2)).astype(np.float32)   # many equal scores
    bb = np.ones(shape, np.float32)
    aa = rng.random((20,) + shape, dtype=np.float32)
    pred = aa.argmax(0).astype(np.float32)
    pts = co.threshold_points(ca, 0.3)
    lab = np.zeros(len(pts), np.int32)                                                  # one cluster: all valid
    want = co.nms(pts, ca[pts[:, 0], pts[:, 1], pts[:, 2]], 9, 0.3)
    res = cd.find_candidates(dict(carbon_alpha_probability=ca, backbone_probability=bb, amino_acid_probability=aa,
                                  amino_acid_prediction=pred), labels_fn=lambda p_: lab)
    assert np.array_equal(res['picks'], want)
    assert res['nms_rounds'] >= 8


def test_candidates_at_full_size_properties(cuda):
    """480^3 (BASELINE configs[1]): threshold == torch.nonzero; picks form the greedy independent set."""
    from mica_b200 import candidates as cd
    g = torch.Generator(device=cuda).manual_seed(2022)
    shape = (480, 480, 480)
    ca = torch.rand(shape, generator=g, device=cuda)
    ca = torch.where(ca > 0.996, torch.rand(shape, generator=g, device=cuda), torch.zeros((), device=cuda))
    lin, xyz = cd.threshold_points(ca, 0.3)
    thr32 = torch.tensor(0.3, dtype=torch.float32, device=cuda)                     # NumPy 2 compares in float32
    want = torch.nonzero(ca.reshape(-1) > thr32).flatten()
    assert torch.equal(lin, want)
    assert torch.equal(xyz.long(), torch.nonzero(ca > thr32))
    labels, n_clusters = cd.dbscan_lattice(lin, shape, 10, 10)
    assert n_clusters >= 1 and int(labels.max()) == n_clusters - 1
    vols = dict(carbon_alpha_probability=ca, backbone_probability=ca,
                amino_acid_probability=ca.unsqueeze(0).expand(20, -1, -1, -1).contiguous() / 20,
                amino_acid_prediction=torch.zeros(shape, device=cuda))
    res = cd.find_candidates(vols, nms_radius=9)
    picks = res['picks']
    valid = res['device']['valid'].bool()
    vp = xyz[valid].cpu().numpy().astype(np.int64)
    sc = ca.reshape(-1)[lin[valid]].cpu().numpy()
    ps = ca.cpu().numpy()[picks[:, 0], picks[:, 1], picks[:, 2]]
    assert np.all(np.diff(ps) <= 0)                                                     # best first
    from scipy.spatial import cKDTree
    assert len(cKDTree(picks).query_pairs(3.0)) == 0                                     # independent
    # every valid point that was not picked has a picked point of higher priority within the radius
    tree = cKDTree(picks)
    rank = {tuple(p): i for i, p in enumerate(picks)}
    lin_of = lambda q: (q[0] * 480 + q[1]) * 480 + q[2]
    for q, s in zip(vp[::7], sc[::7]):
        if tuple(q) in rank:
            continue
        near = tree.query_ball_point(q, 3.0)
        assert any(ps[j] > s or (ps[j] == s and lin_of(picks[j]) < lin_of(q)) for j in near)
    assert len(res['CA_cands']) == int(res['picks_kept'].sum())


def test_clustering_head_sets_the_solver_attributes(cuda):
    from mica_b200 import candidates as cd
    p = candidate_volumes(CANDIDATE_CASES[0])
    solver = types.SimpleNamespace(cluster_eps=10, cluster_min_points=10, nms_radius=9,
                                   modeling_config=types.SimpleNamespace(CA_score_thrh=0.3, output_path='/nonexistent'),
                                   CAProb=p['carbon_alpha_probability'], AAPred=p['amino_acid_prediction'])
    nnpred = types.SimpleNamespace(BBProb=p['backbone_probability'], AAProb=p['amino_acid_probability'],
                                   CAProb_clusted=None)
    cd.clustering_head(solver, nnpred)
    o = co.ca_candidates(p['carbon_alpha_probability'], p['backbone_probability'], p['amino_acid_probability'],
                         p['amino_acid_prediction'])
    assert np.array_equal(solver.CA_cands, o['CA_cands'])
    assert np.array_equal(solver.CA_cands_AAProb, o['CA_cands_AAProb'])
    assert np.array_equal(solver.CA_cands_AA, o['CA_cands_AA'])
    assert np.array_equal(nnpred.CAProb_clusted, o['CAProb_clusted'])


# ------------------------------------------------------------------------------------------ N3
def test_label_masks_match_reference_golden(cuda, golden_dir, tmp_path):
    from mica_b200 import label_masks as lm, mrc
    g = np.load(os.path.join(golden_dir, 'label_masks.npz'))
    shape, origin, st = mask_case()
    mp, pp = str(tmp_path / 'norm.mrc'), str(tmp_path / 's.pdb')
    mrc.write_mrc(mp, mrc.MrcMap(data=np.zeros(shape, np.float32), origin=origin))
    synthetic.write_pdb(pp, st)
    bb = lm.BackboneMask(mp).generate_mask(pp)
    ca = lm.CarbonAlphaMask(mp).generate_mask(pp)
    gen = lm.AminoAcidMaskGenerator(mp)
    aa = gen.generate_mask(pp)
    assert bb.dtype == np.int32 and bb.shape == shape
    assert np.array_equal(bb, g['backbone']) and np.array_equal(ca, g['carbon_alpha'])
    assert np.array_equal(aa, g['amino_acid'])
    out = str(tmp_path / 'amino_acid_mask.mrc')
    gen.save_mask(aa, out)
    back = mrc.read_mrc(out)
    assert np.array_equal(back.data, aa.astype(np.float32)) and tuple(back.origin) == tuple(origin)


def test_label_masks_under_heavy_collisions(cuda):
    """Many atoms per voxel, voxels on every face, coordinates on .5 (round half to even)."""
    from mica_b200 import label_masks as lm
    rng = np.random.default_rng(17)
    shape = (9, 7, 8)
    origin = (np.float32(0.25), np.float32(-1.5), np.float32(2.0))
    for trial in range(12):
        n = int(rng.integers(1, 400))
        coords = (rng.integers(-2, 20, (n, 3)) * 0.5).astype(np.float32) + np.array(origin, np.float32)
        coords[:, 0] = np.minimum(coords[:, 0], origin[0] + shape[2] - 1)      # keep x inside nx (no IndexError)
        coords[:, 2] = np.minimum(coords[:, 2], origin[2] + shape[0] - 1)
        is_cls = rng.random(n) < 0.4
        pos = mo.atom_positions(coords, origin, shape)
        mask, status = lm.class_mask(_dev(coords, cuda), _dev(is_cls.astype(np.uint8), cuda), origin, shape)
        assert int(status.item()) == 0
        assert np.array_equal(mask.cpu().numpy(), mo.atom_class_mask(pos, is_cls, shape)), trial
        labs = rng.integers(1, 21, n).astype(np.int32)
        am, status = lm.aa_mask(_dev(coords, cuda), _dev(labs, cuda), origin, shape)
        assert int(status.item()) == 0
        assert np.array_equal(am.cpu().numpy(), mo.amino_acid_mask(pos, labs.tolist(), shape)), trial
    # an x index that the mis-ordered clip lets through raises IndexError in the reference (D7)
    coords = np.array([[origin[0] + 8.0, origin[1] + 1, origin[2] + 1]], np.float32)   # x = 8 >= nx = 8, < nz = 9
    _, status = lm.class_mask(_dev(coords, cuda), _dev(np.ones(1, np.uint8), cuda), origin, shape)
    assert int(status.item()) == 1
    _, status = lm.aa_mask(_dev(coords, cuda), _dev(np.ones(1, np.int32), cuda), origin, shape)
    assert int(status.item()) == 1
    empty, status = lm.class_mask(torch.zeros((0, 3), device=cuda), torch.zeros(0, dtype=torch.uint8, device=cuda),
                                  origin, shape)
    assert int(empty.abs().sum()) == 0


def test_label_masks_at_config3_size(cuda):
    """BASELINE configs[2]: 20 000 residues in a 480^3 grid -- cross-checked against torch reductions."""
    from mica_b200 import label_masks as lm
    from mica_b200.pdb import channel_codes
    st = synthetic.synthetic_structure(20000, (480, 480, 480), seed=2022)
    coords = _dev(st['coords'], cuda)
    names = np.array(st['atom_names'])
    is_bb = _dev(np.isin(names, ['N', 'CA', 'C', 'O']).astype(np.uint8), cuda)
    mask, status = lm.class_mask(coords, is_bb, (0, 0, 0), (480, 480, 480))
    assert int(status.item()) == 0
    idx = torch.round(coords).long().clamp_(0, 479)
    lin = (idx[:, 2] * 480 + idx[:, 1]) * 480 + idx[:, 0]
    flat = mask.reshape(-1)
    assert bool((flat[lin] >= 2).all())                                       # every atom voxel is 2 or 3
    assert int((flat >= 2).sum()) == int(torch.unique(lin).numel())
    # last writer wins: scatter in file order with torch (deterministic on a sorted, stable key)
    order = torch.arange(lin.numel(), device=cuda)
    last = torch.zeros(480 ** 3, dtype=torch.long, device=cuda).scatter_reduce_(0, lin, order, 'amax', include_self=False)
    want = torch.where(is_bb[last[lin]].bool(), 3, 2).int()
    assert torch.equal(flat[lin], want)
    occ = (flat >= 2).reshape(1, 1, 480, 480, 480).float()
    dil = torch.nn.functional.max_pool3d(occ, 3, 1, 1).reshape(-1) > 0
    assert torch.equal(flat == 1, dil & (flat < 2))                           # 1 = 26-neighbourhood minus atoms


# ------------------------------------------------------------------------------------------ N4
def test_docking_masks_match_reference_golden(cuda, golden_dir, tmp_path):
    from mica_b200 import dock_masks as dm, mrc
    g = np.load(os.path.join(golden_dir, 'docking_masks.npz'))
    shape, origin, _ = mask_case()
    src = synthetic.synthetic_map(shape, seed=3)
    proc = dm.DockingMapMasks()
    for n, (vox, radius) in enumerate(DOCK_CASES):
        inp, thr_p, out_p, pp = (str(tmp_path / f) for f in ('in.mrc', 'thr.mrc', 'out.mrc', 's.pdb'))
        mrc.write_mrc(inp, mrc.MrcMap(data=src, voxel_size=vox, origin=origin))
        synthetic.write_pdb(pp, dock_structure(vox, origin))
        assert proc.initial_map_processing(inp, thr_p, 0.1) == thr_p
        thr = mrc.read_mrc(thr_p)
        assert np.array_equal(np.flatnonzero(thr.data), g['thr_nonzero'])
        assert np.array_equal(thr.data, mo.contour_threshold(src, 0.1))
        assert np.array_equal(np.array(thr.voxel_size, np.float32), g[f'd{n}_voxel'])
        assert proc.subsequent_map_processing(thr_p, pp, out_p, radius=radius) == out_p
        masked = mrc.read_mrc(out_p).data
        assert np.array_equal(np.flatnonzero(masked != thr.data), g[f'd{n}_zeroed']), n
        assert np.array_equal(masked[masked != 0], thr.data[masked != 0])


def test_zero_around_atoms_edge_cases(cuda):
    from mica_b200 import dock_masks as dm
    rng = np.random.default_rng(23)
    shape = (20, 24, 28)
    data = (rng.random(shape, dtype=np.float32) + np.float32(0.5))
    origin = (np.float32(1.5), np.float32(-2.0), np.float32(0.75))
    for vox, radius in [((1.3, 0.7, 2.1), 2.6), ((1.0, 1.0, 1.0), 0.0), ((0.25, 0.25, 0.25), 1.0), ((1.0, 1.0, 1.0), 5.0)]:
        coords = (rng.random((60, 3)) * np.array([19, 23, 19]) * np.array(vox) + np.array(origin)).astype(np.float32)
        coords[:5] -= 40.0                                                     # outside: dropped by the bounds test
        want = mo.mask_around_atoms(data, coords, vox, origin, radius)
        got = _dev(data, cuda)
        status = dm.zero_around_atoms(got, _dev(coords, cuda), vox, origin, radius)
        assert int(status.item()) == 0
        assert np.array_equal(got.cpu().numpy(), want), (vox, radius)
    # x index in [nx, nz) would pass the reference's test only when nz > nx; here nz < nx, so z in [nz, nx) raises
    coords = np.array([[origin[0] + 3, origin[1] + 3, origin[2] + 22]], np.float32)     # z = 22 >= nz = 20, < nx = 28
    status = dm.zero_around_atoms(_dev(data, cuda), _dev(coords, cuda), (1, 1, 1), origin, 2.0)
    assert int(status.item()) == 1
    thr = dm.contour_threshold(_dev(np.array([np.nan, 0.05, 0.1, 0.2, -1.0], np.float32), cuda), 0.1).cpu().numpy()
    assert np.isnan(thr[0]) and list(thr[1:]) == [0.0, np.float32(0.1), np.float32(0.2), 0.0]


def test_predictor_keeps_volumes_resident_for_the_clustering_head(cuda, tmp_path):
    """Solver.nnPred -> Solver.clustering without the 20-channel volume crossing PCIe: the predictor leaves
    amino_acid_probability in HBM (keep_on_device, host_volumes) and clustering_head picks it up from the
    session; results equal the oracle run on the predictor's full host output."""
    from mica_b200 import candidates as cd, mrc, session
    from mica_b200.create_grids import GridCreator
    from mica_b200.predict import CryoEMPredictor
    session.clear()
    p = candidate_volumes(CANDIDATE_CASES[0])
    X, Y, Z = p['carbon_alpha_probability'].shape

    class Replay(torch.nn.Module):
        """Stands where MICA stands: logits whose post-processing reproduces the synthetic volumes at the
        cube cores (3-way softmax of (0, 0, log(2p/(1-p))) has p as its last entry)."""

        def __init__(self, ijk, W, pad):
            super().__init__()
            self.ijk, self.W, self.pad, self.pos = ijk, W, pad, 0

        def forward(self, x, af):
            out = []
            for b in range(x.shape[0]):
                i, j, k = (int(v) - self.pad for v in self.ijk[self.pos + b])
                def window(vol):
                    w = np.zeros((self.W,) * 3, np.float32)
                    xs, ys, zs = (slice(max(0, a), min(n, a + self.W)) for a, n in ((i, X), (j, Y), (k, Z)))
                    w[xs.start - i:xs.stop - i, ys.start - j:ys.stop - j, zs.start - k:zs.stop - k] = vol[xs, ys, zs]
                    return w
                def two_way(prob):
                    q = np.clip(window(prob), 1e-6, 1 - 1e-6).astype(np.float64)
                    l3 = np.log(2 * q / (1 - q)).astype(np.float32)
                    z = np.zeros_like(l3)
                    return np.stack([z, z, z, l3])                 # classes 0, (dropped 1), 2, 3
                aa = np.stack([np.zeros((self.W,) * 3, np.float32)] +
                              [np.log(np.clip(window(p['amino_acid_probability'][c]), 1e-6, 1)) for c in range(20)])
                out.append((two_way(p['backbone_probability']), two_way(p['carbon_alpha_probability']), aa))
            self.pos += x.shape[0]
            return tuple(torch.from_numpy(np.stack([o[t] for o in out])).to(x.device) for t in range(3))

    norm_path = str(tmp_path / 'resampled_normalized_map.mrc')
    mrc.write_mrc(norm_path, mrc.MrcMap(data=np.zeros((Z, Y, X), np.float32)))        # cube space is [x,y,z]
    grids = str(tmp_path / 'grids' / 'ID') + '/'
    r = GridCreator(quiet=True).create_normalized_map_grids(norm_path, os.path.join(grids, 'normalized_map_grids'))
    assert r['success']
    out_path = str(tmp_path / 'out')
    pr = CryoEMPredictor('unused', grids, out_path, save_output=False, quiet=True, keep_on_device=True,
                         host_volumes=('backbone_probability', 'carbon_alpha_probability', 'amino_acid_prediction'))
    assert pr.select_processing_strategy()
    pr.model = Replay(pr._source['ijk'], 64, 8)
    ok, host = pr.run_prediction()
    assert ok and set(host) == {'backbone_probability', 'carbon_alpha_probability', 'amino_acid_prediction',
                                'amino_acid_probability'}
    from mica_b200.predict import DeviceVolume                      # the 20-channel volume stays in HBM
    assert isinstance(host['amino_acid_probability'], DeviceVolume) and host['amino_acid_probability'].tensor.is_cuda
    assert all(isinstance(host[k], np.ndarray) for k in host if k != 'amino_acid_probability')
    reg = session.get(os.path.join(out_path, 'results', 'device_volumes'))
    assert reg is not None and reg['amino_acid_probability'].is_cuda and tuple(reg['amino_acid_probability'].shape) == (20, X, Y, Z)
    full = {k: v.cpu().numpy() for k, v in reg.items()}
    assert np.abs(full['carbon_alpha_probability'] - p['carbon_alpha_probability']).max() < 1e-4
    solver = types.SimpleNamespace(cluster_eps=10, cluster_min_points=10, nms_radius=9,
                                   modeling_config=types.SimpleNamespace(CA_score_thrh=0.3, output_path=out_path))
    res = cd.clustering_head(solver)
    o = co.ca_candidates(full['carbon_alpha_probability'], full['backbone_probability'], full['amino_acid_probability'],
                         full['amino_acid_prediction'])
    assert len(solver.CA_cands) > 20
    assert np.array_equal(solver.CA_cands, o['CA_cands'])
    assert np.array_equal(solver.CA_cands_AAProb, o['CA_cands_AAProb'])
    assert np.array_equal(solver.CA_cands_AA, o['CA_cands_AA'])
    session.clear()


def _unflatten(g, prefix):
    return np.split(g[prefix + '_flat'], np.cumsum(g[prefix + '_len'])[:-1])


@pytest.mark.parametrize('n', range(len(CANDIDATE_CASES)))
def test_neighbor_graph_matches_reference_golden(cuda, golden_dir, n):
    """utils/modeler.py:862-899 (distance matrix, neighbour lists, neigh_mat, best_neigh), bit for bit --
    including the reference's mixed float32/float64 score arithmetic."""
    from mica_b200 import candidates as cd
    g = np.load(os.path.join(golden_dir, 'candidates.npz'))
    p = candidate_volumes(CANDIDATE_CASES[n])
    res = cd.neighbor_graph(g[f'c{n}_CA_cands'], _dev(p['backbone_probability'], cuda))
    assert np.array_equal(res['cand_self_dis'], g[f'c{n}_cand_self_dis'])
    assert np.array_equal(res['neigh_mat'], g[f'c{n}_neigh_mat'])
    want_best = [[int(v) for v in row if v >= 0] for row in g[f'c{n}_best_neigh']]
    assert res['best_neigh'] == want_best
    for key in ('neighbors2to6', 'neighbors0to6', 'neighbors0to7', 'neighbors2to7'):
        want = _unflatten(g, f'c{n}_{key}')
        assert len(want) == len(res[key]) and all(np.array_equal(a, b) for a, b in zip(res[key], want)), key


def test_neighbor_graph_dense_points_and_whole_clustering(cuda):
    from mica_b200 import candidates as cd
    rng = np.random.default_rng(8)
    shape = (30, 28, 26)
    bb = rng.random(shape, dtype=np.float32)
    pts = rng.random((700, 3)) * (np.array(shape) - 3) + 1.0              # ~100 points within 7 A: list regrowth
    res = cd.neighbor_graph(pts, bb)
    dis, nm, best = co.neighbor_scores(pts, bb)
    assert np.array_equal(res['cand_self_dis'], dis) and np.array_equal(res['neigh_mat'], nm)
    assert max(len(v) for v in res['neighbors0to7']) > 64
    for key, want in co.neighbor_scores.lists.items():
        assert all(np.array_equal(a, b) for a, b in zip(res[key], want)), key
    for a, (got, want) in enumerate(zip(res['best_neigh'], best)):
        row = nm[a]
        top = np.sort(row[row > 0])[-3:]
        if len(np.unique(top)) == len(top):                                # untied rows: the reference's order
            assert got == want, a
        else:                                    # ties (coarse float32 scores): same scores, stable choice of index
            assert [row[j] for j in got] == [row[j] for j in want], a
    # the whole method through the Solver-facing wrapper
    p = candidate_volumes(CANDIDATE_CASES[1])
    solver = types.SimpleNamespace(cluster_eps=10, cluster_min_points=10, nms_radius=9,
                                   modeling_config=types.SimpleNamespace(CA_score_thrh=0.3, output_path='/nonexistent'),
                                   CAProb=p['carbon_alpha_probability'], AAPred=p['amino_acid_prediction'],
                                   neighbors2to6=[], neighbors0to6=[], neighbors0to7=[], neighbors2to7=[])
    nnpred = types.SimpleNamespace(BBProb=p['backbone_probability'], AAProb=p['amino_acid_probability'])
    cd.clustering(solver, nnpred)
    o = co.ca_candidates(p['carbon_alpha_probability'], p['backbone_probability'], p['amino_acid_probability'],
                         p['amino_acid_prediction'])
    dis, nm, best = co.neighbor_scores(o['CA_cands'], p['backbone_probability'])
    assert np.array_equal(solver.CA_cands, o['CA_cands']) and np.array_equal(solver.neigh_mat, nm)
    assert np.array_equal(solver.cand_self_dis, dis) and solver.best_neigh == best
    assert len(solver.neighbors2to6) == len(o['CA_cands'])
