"""Pins the oracles of the SURVEY 8(f) rows -- N1 candidates (utils/modeler.py:762-899), N3 label masks
(scripts_for_training_data/create_*_mask.py), N4 docking masks (utils/dock_in_map.py:248-364) -- against
the golden outputs of the UNMODIFIED reference (tests/golden/{candidates,label_masks,docking_masks}.npz),
against the installed libraries they lean on, and, in the build container, against the reference live."""
import os
import tempfile

import numpy as np
import pytest

from mica_b200 import synthetic
from oracle import candidates_oracle as co
from oracle import masks_oracle as mo
from oracle import ref_harness as rh

from _next_cases import CANDIDATE_CASES, DOCK_CASES, candidate_volumes, dock_structure, mask_case


@pytest.mark.parametrize('n', range(len(CANDIDATE_CASES)))
def test_candidates_oracle_matches_reference_golden(golden_dir, n):
    g = np.load(os.path.join(golden_dir, 'candidates.npz'))
    p = candidate_volumes(CANDIDATE_CASES[n])
    ca, bb = p['carbon_alpha_probability'], p['backbone_probability']
    assert np.float64(ca.astype(np.float64).sum()) == g[f'c{n}_ca_crc'], 'synthetic recipe drifted'
    o = co.ca_candidates(ca, bb, p['amino_acid_probability'], p['amino_acid_prediction'])
    assert np.array_equal(o['CA_cands'], g[f'c{n}_CA_cands'])                 # float64, bit for bit
    assert np.array_equal(o['CA_cands_AAProb'], g[f'c{n}_CA_cands_AAProb'])
    assert np.array_equal(o['CA_cands_AA'], g[f'c{n}_CA_cands_AA'])
    assert np.array_equal(np.flatnonzero(o['CAProb_clusted']), g[f'c{n}_clusted_lin'])
    _, nm, _ = co.neighbor_scores(o['CA_cands'], bb)
    assert np.array_equal(nm, g[f'c{n}_neigh_mat'])
    if n == 2:
        assert (~o['picks_kept']).sum() > 0                                   # picks on the border are skipped


def test_dbscan_restated_matches_scikit_learn():
    from sklearn.cluster import DBSCAN
    rng = np.random.default_rng(5)
    noise = 0
    for trial, (dens, eps, mp) in enumerate([(0.02, 3, 4), (0.01, 10, 10), (0.08, 2, 6), (0.004, 5, 3)]):
        occ = rng.random((30, 26, 34)) < dens
        occ[4:9, 5:9, 6:12] |= rng.random((5, 4, 6)) < 0.7                    # a dense blob with border points
        pts = np.array(np.where(occ)).T
        want = DBSCAN(eps=eps, min_samples=mp).fit(pts).labels_
        got = co.dbscan(pts, eps, mp)
        assert np.array_equal(got, want), trial
        assert got.max() >= 0
        noise += int((got == -1).sum())
    assert noise > 0


def test_pairwise_sum_27_is_numpy_sum():
    rng = np.random.default_rng(0)
    v = rng.random((20, 21, 22)).astype(np.float32)
    for _ in range(2000):
        c = rng.integers(1, 19, 3)
        win = v[c[0] - 1:c[0] + 2, c[1] - 1:c[1] + 2, c[2] - 1:c[2] + 2]
        assert co.pairwise_sum_27(win) == np.sum(win)


def test_nms_is_the_unique_greedy_independent_set():
    rng = np.random.default_rng(3)
    pts = np.array(np.where(rng.random((24, 24, 24)) < 0.15)).T
    sc = np.round(rng.random(len(pts)), 2).astype(np.float32) + np.float32(0.3)       # many ties
    picks = co.nms(pts, sc, 9, 0.3)
    order = {tuple(p): i for i, p in enumerate(pts)}
    prio = {tuple(p): (-float(s), order[tuple(p)]) for p, s in zip(pts, sc)}
    picked = {tuple(p) for p in picks}
    for a in picked:
        for b in picked:
            assert a == b or sum((x - y) ** 2 for x, y in zip(a, b)) > 9
    for p in map(tuple, pts):
        if p not in picked:
            assert any(sum((x - y) ** 2 for x, y in zip(p, q)) <= 9 and prio[q] < prio[p] for q in picked)


def test_label_mask_oracles_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'label_masks.npz'))
    shape, origin, st = mask_case()
    pos = mo.atom_positions(st['coords'], origin, shape)
    names, resn = np.array(st['atom_names']), np.array(st['res_names'])
    assert np.array_equal(mo.atom_class_mask(pos, np.isin(names, ['N', 'CA', 'C', 'O']), shape), g['backbone'])
    assert np.array_equal(mo.atom_class_mask(pos, names == 'CA', shape), g['carbon_alpha'])
    sel = [i for i in range(len(names)) if names[i] == 'CA' and resn[i] in mo.AA_LABELS]
    labs = [mo.AA_LABELS[resn[i]] for i in sel]
    assert np.array_equal(mo.amino_acid_mask(pos[sel], labs, shape), g['amino_acid'])
    assert np.array_equal(mo.amino_acid_mask_closed_form(pos[sel], labs, shape), g['amino_acid'])
    assert set(np.unique(g['backbone'])) == {0, 1, 2, 3}


def test_amino_acid_closed_form_equals_sequential_under_collisions():
    rng = np.random.default_rng(11)
    shape = (7, 6, 8)
    for trial in range(60):
        n = int(rng.integers(1, 60))
        pos = np.stack([rng.integers(0, s, n) for s in shape], axis=1)             # crowded: many shared voxels
        labs = rng.integers(1, 21, n).tolist()
        assert np.array_equal(mo.amino_acid_mask(pos, labs, shape),
                              mo.amino_acid_mask_closed_form(pos, labs, shape)), trial


def test_docking_mask_oracles_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'docking_masks.npz'))
    shape, origin, _ = mask_case()
    src = synthetic.synthetic_map(shape, seed=3)
    thr = mo.contour_threshold(src, 0.1)
    assert np.array_equal(np.flatnonzero(thr), g['thr_nonzero'])
    for n, (vox, radius) in enumerate(DOCK_CASES):
        vs = tuple(g[f'd{n}_voxel'])
        sel = mo.select_central_atoms(dock_structure(vox, origin)['coords'])
        for fn in (mo.mask_around_atoms, mo.mask_around_atoms_restated):
            masked = fn(thr, sel, vs, origin, radius)
            assert np.array_equal(np.flatnonzero(masked != thr), g[f'd{n}_zeroed']), (n, fn.__name__)


@pytest.mark.skipif(not rh.available(), reason='needs /root/reference (build container only)')
def test_next_row_oracles_against_the_live_reference():
    p = synthetic.synthetic_predictions((44, 40, 36), (40, 20), seed=9)
    ca, bb, aa, ap = (p[k] for k in ('carbon_alpha_probability', 'backbone_probability', 'amino_acid_probability',
                                     'amino_acid_prediction'))
    r = rh.solver_clustering(ca, bb, aa, ap, ca_score_thrh=0.35, cluster_eps=8, cluster_min_points=6, nms_radius=6)
    o = co.ca_candidates(ca, bb, aa, ap, ca_score_thrh=0.35, cluster_eps=8, cluster_min_points=6, nms_radius=6)
    for k in ('CA_cands', 'CA_cands_AAProb', 'CA_cands_AA', 'CAProb_clusted'):
        assert np.array_equal(r[k], o[k]), k
    shape, origin, st = mask_case(60)
    with tempfile.TemporaryDirectory() as td:
        mp, pp = os.path.join(td, 'norm.mrc'), os.path.join(td, 's.pdb')
        rh._write_mrc(mp, np.zeros(shape, np.float32), (1, 1, 1), origin)
        synthetic.write_pdb(pp, st)
        bbm, cam, aam = rh.label_masks(mp, pp)
    pos = mo.atom_positions(st['coords'], origin, shape)
    names = np.array(st['atom_names'])
    assert np.array_equal(bbm, mo.atom_class_mask(pos, np.isin(names, ['N', 'CA', 'C', 'O']), shape))
    assert np.array_equal(cam, mo.atom_class_mask(pos, names == 'CA', shape))
