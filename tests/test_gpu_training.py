"""The training-data builders (scripts_for_training_data/ mirrors in mica_b200/training_data.py)
against the outputs of the unmodified reference scripts (tests/golden/training_twins.npz,
tests/golden/cubes.npz) and the oracle."""
import glob
import os

import numpy as np
import pytest

from mica_b200 import mrc, pdb, synthetic
from mica_b200.training_data import FeaturesEncoder, MapProcessor, build_grids, create_and_save_grids
from oracle import mica_oracle as orc

pytestmark = pytest.mark.gpu


def _write(path, data, voxel=1.0, origin=(0, 0, 0), **kw):
    mrc.write_mrc(str(path), mrc.MrcMap(data=np.asarray(data, np.float32), voxel_size=(np.float32(voxel),) * 3,
                                        origin=tuple(np.float32(v) for v in origin), **kw))


def test_map_processor_matches_reference_script(cuda, golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, 'training_twins.npz'))
    inp, out = tmp_path / 'emd_0000.map', tmp_path / 'resampled_normalized_map.mrc'
    _write(inp, g['src'], voxel=float(g['voxel'][0]), origin=g['origin'], nxstart=3, nystart=-2, nzstart=7)
    mp = MapProcessor(str(inp))
    assert np.float32(mp.voxel_size.x) == g['voxel'][0] and mp.nzstart == 7
    mp.process_map(str(out), target_voxel_size=1.0)
    got = mrc.read_mrc(str(out))
    assert got.data.shape == g['normalized'].shape
    assert np.abs(got.data - g['normalized']).max() <= 1e-5          # resample tolerance on the [0,1] scale
    assert tuple(got.voxel_size) == (1.0, 1.0, 1.0) and tuple(got.origin) == tuple(g['origin'])
    assert (got.nxstart, got.nystart, got.nzstart) == (3, -2, 7)
    # normalise alone, on the reference's own resampled input: bit-exact (R2/R3)
    res = orc.resample(g['src'], g['voxel'])
    assert np.array_equal(mp.normalize(res), g['normalized'])


def test_map_processor_failure_modes(cuda, tmp_path, capsys):
    inp = tmp_path / 'emd_1.map'
    _write(inp, np.full((12, 12, 12), 3.0, np.float32))              # no value above the median
    mp = MapProcessor(str(inp))
    mp.resample()
    assert mp.normalize() is None and 'Error during normalization' in capsys.readouterr().out
    with pytest.raises(ValueError):
        mp.save_map(str(tmp_path / 'never.mrc'))
    mp.process_map(str(tmp_path / 'never.mrc'))                      # prints, writes nothing
    assert not os.path.exists(tmp_path / 'never.mrc')


def test_features_encoder_matches_reference_script(cuda, golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, 'training_twins.npz'))
    mpath, ppath = tmp_path / 'resampled_normalized_map.mrc', tmp_path / 'x_af3_docked.pdb'
    _write(mpath, g['enc_map'], origin=g['origin'])
    ppath.write_text(str(g['pdb_text']))
    enc = FeaturesEncoder(str(mpath))
    vol = enc.encode_structure(str(ppath))
    assert vol.shape == (24,) + g['enc_map'].shape and vol.dtype == np.float32
    assert np.array_equal(np.argwhere(vol > 0).astype(np.int32), g['enc_nonzero'])      # bit-exact occupancy
    assert enc.get_channel_names() == pdb.CHANNEL_NAMES and enc.get_aa_channel_index('TRP') == 22
    assert enc.get_aa_channel_index('UNK') == -1
    assert list(enc.transform_coordinates(np.array([2.5, 3.5, -0.5], np.float32) + g['origin'])) == [2, 4, 0]
    out = tmp_path / 'CA_encoding.mrc'
    enc.save_channel_as_mrc(vol, str(out), channel_idx=0)
    back = mrc.read_mrc(str(out))
    assert np.array_equal(back.data, vol[0]) and tuple(back.origin) == tuple(g['origin'])


def test_features_encoder_index_error_on_non_cubic_map(cuda, tmp_path):
    """The (nz,ny,nx)-vs-(x,y,z) clip (D7): an atom at x beyond nz-1 is silently clamped, one whose
    z index still exceeds nz raises IndexError, which the script's main() catches."""
    mpath, ppath = tmp_path / 'm.mrc', tmp_path / 'p.pdb'
    _write(mpath, np.zeros((10, 20, 30), np.float32))
    st = synthetic.synthetic_structure(3, (30, 20, 10), seed=1)
    st['coords'][:] = np.float32(1.0)
    st['coords'][0] = (25.0, 5.0, 3.0)         # x clamps to nz-1 = 9
    synthetic.write_pdb(str(ppath), st)
    vol = FeaturesEncoder(str(mpath)).encode_structure(str(ppath))
    assert vol[:, 3, 5, 9].any() and not vol[:, 3, 5, 25].any()
    st['coords'][1] = (2.0, 2.0, 15.0)         # z = 15 passes the clip (bound nx-1 = 29) but nz = 10
    synthetic.write_pdb(str(ppath), st)
    with pytest.raises(IndexError):
        FeaturesEncoder(str(mpath)).encode_structure(str(ppath))


def test_create_and_save_grids_match_reference_script(cuda, golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, 'cubes.npz'))
    p = tmp_path / 'resampled_normalized_map.mrc'
    _write(p, g['t_vol'], origin=(1.5, 2.5, 3.5), mapc=1, mapr=2, maps=3)
    n = create_and_save_grids(str(p), str(tmp_path / 'g'), 8, 2, min_max=0.01)
    assert n == len(g['t_meta'])                                      # the reference script's count
    want, meta, shp, _ = orc.extract_cubes(g['t_vol'], grid_size=8, padding=2, transpose=False, drop_below=0.01)
    files = sorted(glob.glob(str(tmp_path / 'g' / 'grid_i*.npz')))
    assert len(files) == n
    for c, m in zip(want, meta):
        d = np.load(str(tmp_path / 'g' / f'grid_i{m[0]}_j{m[1]}_k{m[2]}.npz'), allow_pickle=True)
        assert np.array_equal(d['grid'], c)
        assert (int(d['di']), int(d['dj']), int(d['dk'])) == tuple(m[3:6])
        assert tuple(d['orig_shape']) == g['t_vol'].shape and int(d['grid_size']) == 8 and int(d['padding']) == 2
        assert float(d['origin']['x']) == 1.5 and int(d['maps']) == 3
    # the mask / encoding variants keep every window
    n_all = create_and_save_grids(str(p), str(tmp_path / 'all'), 8, 2)
    assert n_all == len(orc.extract_cubes(g['t_vol'], grid_size=8, padding=2, transpose=False)[0]) > n
    # int8 label masks keep their dtype
    lab = (np.random.default_rng(0).integers(0, 4, size=(9, 10, 11))).astype(np.int8)
    mp = tmp_path / 'backbone_mask.mrc'
    mrc.write_mrc(str(mp), mrc.MrcMap(data=lab), dtype=np.int8)
    create_and_save_grids(str(mp), str(tmp_path / 'bb'), 8, 2)
    d = np.load(str(tmp_path / 'bb' / 'grid_i8_j8_k8.npz'), allow_pickle=True)
    assert d['grid'].dtype == np.int8 and np.array_equal(d['grid'][2:3, 2:4, 2:5], lab[8:9, 8:10, 8:11])


def test_build_grids_walks_the_dataset_tree(cuda, tmp_path):
    base = tmp_path / 'Processed_Data'
    for emd in ('0001', '0002'):
        os.makedirs(base / emd)
        v = np.random.default_rng(int(emd)).random((10, 12, 9)).astype(np.float32)
        _write(base / emd / 'resampled_normalized_map.mrc', v)
        _write(base / emd / 'CA_encoding.mrc', (v > 0.9))
        _write(base / emd / 'ALA_encoding.mrc', (v > 0.95))
    assert build_grids('normalized_map', str(base), str(tmp_path / 'G' / 'normalized_maps'), 8, 2) == 2 * 8
    assert build_grids('AF3_encodings', str(base), str(tmp_path / 'G'), 8, 2) == 2 * 2 * 8
    assert os.path.exists(tmp_path / 'G' / 'CA_encodings' / '0002' / 'grid_i8_j8_k8.npz')
