"""Host-side I/O of the drop-in: MRC2014 round trips against the
independent stand-in the reference harness uses, and the fixed-column PDB reader."""
import os
import sys

import numpy as np
import pytest

from mica_b200 import mrc, pdb, synthetic
from mica_b200.pipeline import MapHeader, zoom_factors
from oracle import mica_oracle as orc

STANDINS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', 'standins')


def _standin_mrcfile():
    sys.path.insert(0, STANDINS)
    try:
        import importlib
        return importlib.import_module('mrcfile')
    finally:
        sys.path.remove(STANDINS)


def test_mrc_roundtrip_and_cross_read(tmp_path):
    rng = np.random.default_rng(0)
    data = rng.normal(size=(5, 7, 9)).astype(np.float32)
    m = mrc.MrcMap(data=data, voxel_size=(np.float32(1.06), np.float32(1.1), np.float32(0.9)),
                   origin=(np.float32(-3.5), np.float32(2.25), np.float32(10)), mapc=2, mapr=3, maps=1,
                   nxstart=4, nystart=-5, nzstart=6)
    p = str(tmp_path / 'a.mrc')
    mrc.write_mrc(p, m)
    r = mrc.read_mrc(p)
    assert np.array_equal(r.data, data) and (r.mapc, r.mapr, r.maps) == (2, 3, 1)
    assert (r.nxstart, r.nystart, r.nzstart) == (4, -5, 6)
    assert r.origin == m.origin
    mf = _standin_mrcfile()
    with mf.open(p) as f:                                       # what the reference would see
        assert np.array_equal(f.data, data)
        assert (int(f.header.mapc), int(f.header.nzstart)) == (2, 6)
        assert np.float32(f.voxel_size.x) == r.voxel_size[0] and np.float32(f.voxel_size.z) == r.voxel_size[2]
        assert np.float32(f.header.origin.y) == np.float32(2.25)
    p2 = str(tmp_path / 'b.mrc')
    with mf.new(p2, overwrite=True) as f:                       # what the reference would write
        f.set_data(data)
        f.voxel_size = (1.2, 1.2, 1.2)
        f.header.origin.x = 7
        f.header.nystart = 3
    r2 = mrc.read_mrc(p2)
    assert np.array_equal(r2.data, data) and r2.voxel_size[1] == np.float32(np.float32(1.2 * 7) / np.float32(7))
    assert r2.origin[0] == 7 and r2.nystart == 3


def test_pdb_reader_matches_biopython_standin(tmp_path):
    st = synthetic.synthetic_structure(50, (30, 30, 30), seed=3, hetero_every=6, unknown_every=5)
    p = str(tmp_path / 's.pdb')
    synthetic.write_pdb(p, st)
    coords, bb, aa, nres = pdb.read_pdb_atoms(p)
    keep = ~st['hetero']
    assert np.array_equal(coords, st['coords'][keep])           # %8.3f text round trip is exact
    obb, oaa = orc.channel_codes([a for a, k in zip(st['atom_names'], keep) if k],
                                 [r for r, k in zip(st['res_names'], keep) if k])
    assert np.array_equal(bb, obb) and np.array_equal(aa, oaa)
    sys.path.insert(0, STANDINS)
    try:
        from Bio import PDB
        structure = PDB.PDBParser(QUIET=True).get_structure('x', p)
    finally:
        sys.path.remove(STANDINS)
    atoms = [(a.get_name(), r.get_resname(), a.get_coord()) for m in structure for c in m for r in c
             if r.get_id()[0] == ' ' for a in r]
    assert len(atoms) == len(coords)
    assert np.array_equal(np.stack([a[2] for a in atoms]), coords)


def test_header_transpose_and_zoom_bookkeeping():
    for axes in [(1, 2, 3), (2, 3, 1), (3, 1, 2), (1, 3, 2), (2, 1, 3), (3, 2, 1)]:
        h = MapHeader(mapc=axes[0], mapr=axes[1], maps=axes[2], nxstart=4, nystart=-7, nzstart=11)
        perm, off = h.transpose_order()
        operm, ooff = orc.transpose_order(*axes, (11, -7, 4))
        assert list(perm) == list(operm) and off == ooff
    v = (np.float32(1.06), np.float32(1.13), np.float32(0.97))
    assert [float(a) for a in zoom_factors(v)] == [float(a) for a in orc.zoom_factors(v)]


# ---------------------------------------------------------------- host logic of the SURVEY 8(f) mirrors (no GPU)
def test_valid_labels_follow_the_reference_cluster_filter():
    """candidates.valid_labels (host part of utils/modeler.py:781-797) against the oracle's restatement."""
    from mica_b200 import candidates as cd
    from oracle import candidates_oracle as co
    rng = np.random.default_rng(4)
    bb = rng.random((12, 12, 12), dtype=np.float32)
    pts = np.array(np.where(rng.random((12, 12, 12)) < 0.3)).T
    labels = rng.integers(-1, 6, len(pts))
    labels[:40] = 5                                                   # one dominant cluster
    bb[tuple(pts[labels == 2].T)] *= 0.05                             # one weak cluster (dropped by the sum test)
    val, sums, avgs = co.valid_clusters(pts, labels, bb)
    at = bb[pts[:, 0], pts[:, 1], pts[:, 2]].astype(np.float64)
    s64 = np.array([at[labels == k].sum() for k in range(6)])
    cnt = np.array([(labels == k).sum() for k in range(6)])
    ok, s32, a32 = cd.valid_labels(s64, cnt)
    assert np.allclose(s32, sums, rtol=1e-6) and np.allclose(a32, avgs, rtol=1e-6)
    assert np.array_equal(ok[labels[labels >= 0]], val[labels >= 0]) and not val[labels < 0].any()
    assert not ok.all() and ok.any()
    with pytest.raises(ValueError):                                   # np.max of an empty list in the reference
        cd.valid_labels(np.zeros(0), np.zeros(0, np.int64))


def test_pdb_records_and_central_atom_selection(tmp_path):
    from mica_b200 import dock_masks as dm, pdb, synthetic
    from oracle import masks_oracle as mo
    st = synthetic.synthetic_structure(40, (30, 30, 30), seed=3, hetero_every=5, unknown_every=7)
    p = str(tmp_path / 's.pdb')
    synthetic.write_pdb(p, st)
    rec = pdb.read_pdb_records(p)
    assert np.array_equal(rec['coords'], st['coords'])                # ATOM and HETATM records, file order
    assert rec['atom_names'] == list(st['atom_names']) and rec['res_names'] == list(st['res_names'])
    assert rec['res_index'][0] == 0 and rec['res_index'][-1] == 39 and np.all(np.diff(rec['res_index']) >= 0)
    atoms_only, _, _, n_res = pdb.read_pdb_atoms(p)
    assert len(atoms_only) < len(rec['coords'])                       # the AF3 encoder skips hetero residues
    for method in ('median', 'mean'):
        assert np.array_equal(dm.select_central_atoms(rec['coords'], 40, method),
                              mo.select_central_atoms(rec['coords'], 40, method))
    with pytest.raises(ValueError):
        dm.select_central_atoms(rec['coords'], 40, 'mode')


def test_next_row_mirrors_fail_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from mica_b200 import _lib, candidates as cd, dock_masks as dm, label_masks as lm, mrc
    vols = {k: np.zeros((4, 4, 4), np.float32) for k in ('carbon_alpha_probability', 'backbone_probability',
                                                        'amino_acid_prediction')}
    vols['amino_acid_probability'] = np.zeros((20, 4, 4, 4), np.float32)
    with pytest.raises(_lib.MicaError):
        cd.find_candidates(vols)
    with pytest.raises(_lib.MicaError):
        cd.neighbor_graph(np.zeros((3, 3)), np.zeros((4, 4, 4), np.float32))
    with pytest.raises(_lib.MicaError):
        lm.class_mask(torch.zeros((1, 3)), torch.zeros(1, dtype=torch.uint8), (0, 0, 0), (4, 4, 4))
    with pytest.raises(_lib.MicaError):
        dm.contour_threshold(torch.zeros(8), 0.1)
    p = str(tmp_path / 'm.mrc')
    mrc.write_mrc(p, mrc.MrcMap(data=np.zeros((4, 4, 4), np.float32)))
    with pytest.raises(_lib.MicaError):
        lm.BackboneMask(p)
    with pytest.raises(_lib.MicaError):
        dm.DockingMapMasks()


def test_native_pdb_parser_does_not_depend_on_the_number_of_pieces(tmp_path, monkeypatch):
    """mica_parse_pdb cuts the text into pieces of whole lines, one host thread each; residue runs, MODEL
    counts and duplicate-name detection that straddle a cut are stitched.  Every cut position must give the
    result of the single-piece parse (and of the NumPy parser)."""
    st = synthetic.synthetic_structure(40, (30, 30, 30), seed=5, hetero_every=7, unknown_every=5)
    p = str(tmp_path / 's.pdb')
    synthetic.write_pdb(p, st)
    lines = open(p).read().splitlines()
    atom_lines = [i for i, l in enumerate(lines) if l.startswith('ATOM')]
    # an alternate location (same residue, same name) and two MODEL records in the middle of the file
    dup = lines[atom_lines[20]]
    dup = dup[:16] + 'B' + dup[17:54] + '  0.40' + dup[60:]
    with_dup = lines[:atom_lines[20] + 1] + [dup] + lines[atom_lines[20] + 1:]
    with_models = ['MODEL        1'] + lines[:atom_lines[60]] + ['ENDMDL', 'MODEL        2'] + lines[atom_lines[60]:]
    for name, body in (('plain', lines), ('dup', with_dup), ('models', with_models)):
        q = str(tmp_path / f'{name}.pdb')
        with open(q, 'w') as fh:
            fh.write('\n'.join(body) + '\n')
        results = []
        for threads, piece in ((1, 1 << 20), (2, 64), (5, 64), (16, 64), (64, 40)):
            monkeypatch.setenv('MICA_PDB_THREADS', str(threads))
            monkeypatch.setenv('MICA_PDB_MIN_PIECE', str(piece))
            got = pdb._records_native(q, True)
            extras = pdb._records_native.last
            results.append(tuple(np.array(a) for a in got) + (extras['bb'].copy(), extras['aa'].copy(),
                                                              extras['n_res'], extras['dup']))
        for r in results[1:]:
            for a, b in zip(results[0], r):
                assert np.array_equal(a, b), name
        assert results[0][-1] == (name == 'dup')
        ref = pdb._records_numpy(q, True)
        for a, b in zip(results[0][:4], ref):
            assert np.array_equal(a, b), name
        monkeypatch.delenv('MICA_PDB_THREADS')
        monkeypatch.delenv('MICA_PDB_MIN_PIECE')
        assert len(pdb.read_pdb_atoms(q)[0]) > 0
