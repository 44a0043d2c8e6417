"""The reference-facing classes (DataPreprocessor / GridCreator / CryoEMPredictor
mirrors) driven the way Solver.getData / Solver.nnPred drive the originals
(utils/modeler.py:673-760), checked against the oracle."""
import os

import numpy as np
import pytest
import torch

from mica_b200 import mrc, pdb, session, synthetic
from mica_b200.create_grids import GridCreator
from mica_b200.predict import CryoEMPredictor
from mica_b200.preprocessing import DataPreprocessor
from oracle import mica_oracle as orc

pytestmark = pytest.mark.gpu


class PointwiseModel(torch.nn.Module):
    """Deterministic stand-in for MICA with its (exp_map, af_features) -> (bb, ca, aa)
    signature; elementwise, so the CPU oracle can evaluate it on the oracle's cubes."""

    def forward(self, x, af):
        s = af.sum(dim=1, keepdim=True)
        bb = torch.cat([x, -x, 2 * x - 0.5 + s, x * x], dim=1)
        ca = torch.cat([0.5 - x, x, x * 3 - 1, 1.5 * x + af[:, :1]], dim=1)
        aa = torch.cat([x * (0.1 * t) + af[:, t % 24:t % 24 + 1] * (t % 3) + np.sin(t) for t in range(21)], dim=1)
        return bb, ca, aa


def _inputs(tmp_path, shape=(40, 40, 40), voxel=1.1):
    case = tmp_path / 'input' / 'ID'
    os.makedirs(case / 'AF3_results', exist_ok=True)
    src = synthetic.synthetic_map(shape, voxel=voxel, seed=31)
    origin = (np.float32(12.5), np.float32(-4.0), np.float32(0.75))
    map_path = str(case / 'map.mrc')
    mrc.write_mrc(map_path, mrc.MrcMap(data=src, voxel_size=(np.float32(voxel),) * 3, origin=origin,
                                       nxstart=2, nystart=3, nzstart=5))
    n = round(shape[0] * voxel)
    st = synthetic.synthetic_structure(150, (n, n, n), seed=31, origin_xyz=origin, hetero_every=13)
    pdb_path = str(case / 'ID_af3_docked.pdb')
    synthetic.write_pdb(pdb_path, st)
    return src, origin, map_path, pdb_path, str(case / 'AF3_results'), str(case / 'grids')   # as run.py:108-112


def test_get_data_then_nn_pred_like_the_solver(cuda, tmp_path):
    session.clear()
    src, origin, map_path, pdb_path, af3_results, grids_path = _inputs(tmp_path)
    voxel = (np.float32(1.1),) * 3
    # ---- Solver.getData
    dp = DataPreprocessor(map_path=map_path, AF3_results=af3_results, quiet=True)
    assert dp.resample_and_normalize_map() is None
    assert dp.normalized_map_path.endswith('resampled_normalized_map.mrc')
    o_norm, _, _ = orc.normalize(orc.resample(src, voxel))
    written = mrc.read_mrc(dp.normalized_map_path)
    assert np.abs(written.data - o_norm).max() <= 1e-5
    assert written.voxel_size == (1.0, 1.0, 1.0) and written.origin == origin and written.nzstart == 5
    assert dp.create_AF3_encodings(pdb_path) is True
    coords, bb_ch, aa_ch, _ = pdb.read_pdb_atoms(pdb_path)
    o_af3, ok = orc.af3_encode(coords, bb_ch, aa_ch, origin, o_norm.shape)
    assert ok
    for c, name in enumerate(pdb.CHANNEL_NAMES):
        got = mrc.read_mrc(os.path.join(dp.AF3_encodings, f'{name}_encoding.mrc')).data
        assert np.array_equal(got, o_af3[c]), name
    gc = GridCreator(quiet=True)
    r1 = gc.create_normalized_map_grids(dp.normalized_map_path, os.path.join(grids_path, 'normalized_map_grids'))
    r2 = gc.create_AF3_encodings_grids(dp.AF3_encodings, os.path.join(grids_path, 'AF3_encoding_grids'))
    o_cubes, o_meta, o_shape, o_off = orc.extract_cubes(o_norm, nstart_zyx=(5, 3, 2))
    assert r1['success'] and r1['grid_count'] == len(o_cubes) and r1['offset'] == o_off
    assert set(r1) == {'success', 'grid_count', 'offset', 'output_directory', 'processing_time', 'input_file'}
    assert r2['success'] and r2['successful_channels'] == 24 and r2['total_grids'] == 24 * len(o_cubes)
    # ---- Solver.nnPred
    pr = CryoEMPredictor(model_path='unused', grids_path=grids_path, output_path=str(tmp_path / 'out'),
                         save_output=True, device='cuda', quiet=True, model=PointwiseModel())
    ok, vols = pr.run_prediction()
    assert ok and set(vols) == {'backbone_probability', 'carbon_alpha_probability', 'amino_acid_prediction',
                                'amino_acid_probability'}
    # oracle: same model on the oracle's cubes, reference post-processing + stitching
    norm_dev = torch.from_numpy(written.data)
    o_x = torch.from_numpy(orc.extract_cubes(written.data)[0][:, None])
    o_af = torch.from_numpy(np.stack([orc.extract_cubes(o_af3[c])[0] for c in range(24)], axis=1))
    with torch.no_grad():
        bb, ca, aa = PointwiseModel()(o_x, o_af)
    want = orc.postprocess_and_stitch(bb, ca, aa, o_meta, o_shape)
    for k in ('backbone_probability', 'carbon_alpha_probability', 'amino_acid_probability'):
        assert vols[k].shape == want[k].shape and np.abs(vols[k] - want[k]).max() <= 2e-5, k
    assert (vols['amino_acid_prediction'] == want['amino_acid_prediction']).mean() > 0.999
    assert os.path.exists(os.path.join(str(tmp_path / 'out'), 'results', 'ID', 'amino_acid_probability.npy'))


def test_materialized_npz_files_and_file_fallback(cuda, tmp_path):
    """materialize=True writes the reference's per-cube files; a predictor in a fresh
    process (empty session) consumes them and gives the same volumes."""
    session.clear()
    rng = np.random.default_rng(4)
    vol = rng.random((20, 52, 30), dtype=np.float32)
    p = str(tmp_path / 'resampled_normalized_map.mrc')
    mrc.write_mrc(p, mrc.MrcMap(data=vol, nxstart=1, nystart=2, nzstart=3))
    grids = str(tmp_path / 'grids' / 'ID') + '/'
    gc = GridCreator(quiet=True, materialize=True)
    res = gc.create_normalized_map_grids(p, os.path.join(grids, 'normalized_map_grids'))
    o_cubes, o_meta, o_shape, o_off = orc.extract_cubes(vol, nstart_zyx=(3, 2, 1))
    assert res['grid_count'] == len(o_cubes)
    for c, m in zip(o_cubes, o_meta):
        d = np.load(os.path.join(grids, 'normalized_map_grids', f'normalized_map_grid_i{m[0]}_j{m[1]}_k{m[2]}.npz'))
        assert np.array_equal(d['grid'], c) and (int(d['di']), int(d['dj']), int(d['dk'])) == tuple(m[3:])
        assert tuple(d['orig_shape']) == tuple(o_shape) and int(d['padding']) == 8
    model = PointwiseModel()
    ok1, v1 = CryoEMPredictor('unused', grids, str(tmp_path / 'o1'), save_output=False, model=model,
                              quiet=True).run_prediction()
    session.clear()
    ok2, v2 = CryoEMPredictor('unused', grids, str(tmp_path / 'o2'), save_output=False, model=model,
                              quiet=True).run_prediction()
    assert ok1 and ok2
    for k in v1:
        assert np.array_equal(v1[k], v2[k]), k


def test_failure_conventions(cuda, tmp_path):
    """Never raise at the boundary: return None / False / (False, {}) like the reference."""
    session.clear()
    dp = DataPreprocessor(map_path=str(tmp_path / 'missing.mrc'), AF3_results=str(tmp_path / 'AF3_results') + '/')
    assert dp.resample_and_normalize_map() is None and dp.normalized_map_path is None
    flat = str(tmp_path / 'flat.mrc')
    mrc.write_mrc(flat, mrc.MrcMap(data=np.zeros((8, 8, 8), np.float32)))
    dp = DataPreprocessor(map_path=flat, AF3_results=str(tmp_path / 'AF3_results') + '/')
    dp.resample_and_normalize_map()
    assert dp.normalized_map_path is None                       # no positive values -> nothing written
    gc = GridCreator(quiet=True)
    assert gc.create_normalized_map_grids(str(tmp_path / 'nope.mrc'), str(tmp_path / 'g'))['success'] is False
    assert gc.create_grids_from_mrc(str(tmp_path / 'nope.mrc'), str(tmp_path / 'g')) == (0, None)
    ok, vols = CryoEMPredictor('unused', str(tmp_path / 'empty') + '/', str(tmp_path / 'o')).run_prediction()
    assert ok is False and vols == {}


def test_zero_af3_cubes_are_batched_apart(cuda, tmp_path):
    """D8: with batching on, a cube without AF3 signal must see the model's zero-AF3 branch."""
    session.clear()
    rng = np.random.default_rng(9)
    vol = rng.random((20, 20, 100), dtype=np.float32)
    af3 = np.zeros((24, 20, 20, 100), np.float32)
    af3[3, 5, 5, 10] = 1.0                                       # only the first cube along x has AF3 signal
    p = str(tmp_path / 'resampled_normalized_map.mrc')
    mrc.write_mrc(p, mrc.MrcMap(data=vol))
    enc = tmp_path / 'AF3_encodings'
    os.makedirs(enc)
    for c, name in enumerate(pdb.CHANNEL_NAMES):
        mrc.write_mrc(str(enc / f'{name}_encoding.mrc'), mrc.MrcMap(data=af3[c]))
    grids = str(tmp_path / 'grids' / 'ID') + '/'
    gc = GridCreator(quiet=True)
    gc.create_normalized_map_grids(p, os.path.join(grids, 'normalized_map_grids'))
    gc.create_AF3_encodings_grids(str(enc), os.path.join(grids, 'AF3_encoding_grids'))
    seen = []

    class Probe(PointwiseModel):
        def forward(self, x, af):
            seen.append((x.shape[0], bool(af.abs().sum() < 1e-6)))
            return super().forward(x, af)

    pr = CryoEMPredictor('unused', grids, str(tmp_path / 'o'), save_output=False, model=Probe(), quiet=True)
    pr.batch_threshold = 1                                      # force the batched strategy
    ok, _ = pr.run_prediction()
    assert ok and sorted(seen) == [(1, False), (2, True)]
