"""Executes the UNMODIFIED reference (/root/reference) in the build container.

TEST INFRASTRUCTURE ONLY (see mica_oracle.py).  /root/reference does not exist
on the GPU box, so nothing at run time may import this module; it is used by
oracle/make_golden.py to produce tests/golden/*.npz and by the container-only
tests that are skipped when /root/reference is absent.

The reference needs ``mrcfile`` and ``Bio.PDB`` (absent from the image); the
I/O-only stand-ins under oracle/standins/ are put on sys.path when the real
packages cannot be imported.  No reference source is copied: the modules are
imported from where they lie.
"""
from __future__ import annotations

import contextlib
import glob
import io
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_STANDINS = os.path.join(_HERE, 'standins')
#: the staged archive of the unmodified reference (oracle/make_ref.py; travels to the GPU box)
STAGED_ARCHIVE = os.path.join(_HERE, '_ref', 'reference_py.zip')


def _find_root():
    root = os.environ.get('MICA_REFERENCE_ROOT', '/root/reference')
    if os.path.isfile(os.path.join(root, 'utils', 'preprocessing.py')):
        return root
    if os.path.isfile(STAGED_ARCHIVE):
        return STAGED_ARCHIVE                       # imported through zipimport
    return root


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return REFERENCE_ROOT == STAGED_ARCHIVE or os.path.isfile(os.path.join(REFERENCE_ROOT, 'utils', 'preprocessing.py'))


def live_tree() -> bool:
    """True when the reference's source tree itself is present (the build container)."""
    return REFERENCE_ROOT != STAGED_ARCHIVE and available()


def _setup_path():
    try:
        import mrcfile  # noqa: F401
        import Bio.PDB  # noqa: F401
    except ImportError:
        if _STANDINS not in sys.path:
            sys.path.insert(0, _STANDINS)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


@contextlib.contextmanager
def _quiet():
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


def _write_mrc(path, data, voxel_xyz=(1.0, 1.0, 1.0), origin_xyz=(0, 0, 0),
               nstart_xyz=(0, 0, 0), axes=(1, 2, 3)):
    _setup_path()
    import mrcfile
    with mrcfile.new(path, overwrite=True) as m:
        m.set_data(np.asarray(data, dtype=np.float32))
        m.voxel_size = tuple(voxel_xyz)
        m.header.origin.x, m.header.origin.y, m.header.origin.z = origin_xyz
        m.header.nxstart, m.header.nystart, m.header.nzstart = nstart_xyz
        m.header.mapc, m.header.mapr, m.header.maps = axes
        m.update_header_stats()


def _read_mrc(path):
    _setup_path()
    import mrcfile
    with mrcfile.open(path) as m:
        return np.array(m.data, dtype=np.float32)


def resample_and_normalize(src, voxel_xyz, workdir, origin_xyz=(0, 0, 0),
                           nstart_xyz=(0, 0, 0), axes=(1, 2, 3)):
    """DataPreprocessor.resample_and_normalize_map (utils/preprocessing.py:80).
    Returns (normalised volume or None, path of the written MRC or None)."""
    _setup_path()
    from utils.preprocessing import DataPreprocessor
    map_path = os.path.join(workdir, 'input_map.mrc')
    af3_dir = os.path.join(workdir, 'AF3_results', 'x')
    os.makedirs(af3_dir, exist_ok=True)
    _write_mrc(map_path, src, voxel_xyz, origin_xyz, nstart_xyz, axes)
    with _quiet():
        dp = DataPreprocessor(map_path=map_path, AF3_results=af3_dir + '/', quiet=True)
        dp.logger.disabled = True
        dp.resample_and_normalize_map()
    out = dp.normalized_map_path
    if out is None or not os.path.exists(out):
        return None, None, dp
    return _read_mrc(out), out, dp


def af3_encodings(dp, pdb_path):
    """DataPreprocessor.create_AF3_encodings (utils/preprocessing.py:225) on the
    preprocessor returned by resample_and_normalize.  Returns (ok, [24,nz,ny,nx] or None)."""
    _setup_path()
    from mica_b200.pdb import CHANNEL_NAMES
    with _quiet():
        ok = dp.create_AF3_encodings(pdb_path)
    if not ok:
        return False, None
    vols = [_read_mrc(os.path.join(dp.AF3_encodings, f'{n}_encoding.mrc')) for n in CHANNEL_NAMES]
    return True, np.stack(vols)


def grids_from_mrc(mrc_path, outdir, grid_size=48, padding=8, prefix='grid'):
    """GridCreator.create_grids_from_mrc (utils/create_grids.py:89).  Returns
    (grid_count, offset, {(i,j,k): (cube, di, dj, dk)}, orig_shape)."""
    _setup_path()
    from utils.create_grids import GridCreator
    with _quiet():
        gc = GridCreator(quiet=True)
        gc.logger.disabled = True
        count, offset = gc.create_grids_from_mrc(mrc_path, outdir, grid_size, padding, prefix)
    cubes, orig_shape = {}, None
    for f in sorted(glob.glob(os.path.join(outdir, f'{prefix}_i*.npz'))):
        d = np.load(f, allow_pickle=True)
        cubes[(int(d['i']), int(d['j']), int(d['k']))] = (
            np.array(d['grid']), int(d['di']), int(d['dj']), int(d['dk']))
        orig_shape = tuple(int(v) for v in d['orig_shape'])
    return count, offset, cubes, orig_shape


def training_grids_from_mrc(mrc_path, outdir, grid_size=48, padding=8):
    """scripts_for_training_data/create_grids_for_normalized_map.py:18 (no
    transpose, drops cubes whose max < 0.01)."""
    _setup_path()
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, 'scripts_for_training_data'))
    import create_grids_for_normalized_map as m
    count = m.create_and_save_grids(mrc_path, outdir, grid_size, padding)
    cubes = {}
    for f in sorted(glob.glob(os.path.join(outdir, 'grid_i*.npz'))):
        d = np.load(f, allow_pickle=True)
        cubes[(int(d['i']), int(d['j']), int(d['k']))] = (
            np.array(d['grid']), int(d['di']), int(d['dj']), int(d['dk']))
    return count, cubes


def training_map_processor(src, voxel_xyz, workdir, origin_xyz=(0, 0, 0)):
    """scripts_for_training_data/create_normalized_map.py::MapProcessor.process_map.
    Returns the normalised volume read back from the MRC it wrote (None if it wrote none)."""
    _setup_path()
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, 'scripts_for_training_data'))
    import create_normalized_map as m
    inp = os.path.join(workdir, 'emd_0000.map')
    out = os.path.join(workdir, 'tw_resampled_normalized_map.mrc')
    _write_mrc(inp, src, voxel_xyz, origin_xyz)
    with _quiet():
        m.MapProcessor(inp).process_map(out, target_voxel_size=1.0)
    return (_read_mrc(out), out) if os.path.exists(out) else (None, None)


def training_features_encoder(normalized_map_path, pdb_path):
    """scripts_for_training_data/create_AF3_encodings.py::FeaturesEncoder.encode_structure.
    Returns the float32 (24,nz,ny,nx) volume; IndexError propagates as in the reference."""
    _setup_path()
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, 'scripts_for_training_data'))
    import create_AF3_encodings as m
    with _quiet():
        enc = m.FeaturesEncoder(normalized_map_path)
        vol = enc.encode_structure(pdb_path)
    return vol.astype(np.float32), enc.get_channel_names()


class _ReplayModel:
    """Stands where MICA stands in run_inference (utils/predict.py:339): returns
    pre-generated logits for the cubes of the batch, keyed by the map cube's
    content hash so that the DataLoader order does not matter."""

    def __init__(self, table):
        self.table = table

    def eval(self):
        return self

    def __call__(self, x, af3):
        import torch
        rows = [self.table[x[b].numpy().tobytes()] for b in range(x.shape[0])]
        self.last_af3 = af3
        return tuple(torch.from_numpy(np.stack([r[t] for r in rows])) for t in range(3))


def predict_and_stitch(grids_path, output_path, logits_for_cube, batch_threshold=None):
    """CryoEMPredictor.{select_processing_strategy, prepare_data, run_inference,
    reconstruct_and_save_volumes} (utils/predict.py:176-587) on CPU with the
    model replaced by a logits replay.  ``logits_for_cube``: dict map-cube-bytes
    -> (bb[4,W,W,W], ca[4,...], aa[21,...]).  Returns (ok, volumes dict)."""
    _setup_path()
    from utils.predict import CryoEMPredictor
    with _quiet():
        pr = CryoEMPredictor(model_path='unused', grids_path=grids_path.rstrip('/') + '/',
                             output_path=output_path, save_output=False, device='cpu', quiet=True)
        pr.logger.disabled = True
        if batch_threshold is not None:
            pr.batch_threshold = batch_threshold
        assert pr.select_processing_strategy()
        pr.model = _ReplayModel(logits_for_cube)
        ok, loader = pr.prepare_data()
        assert ok
        assert pr.run_inference(loader)
        ok, vols = pr.reconstruct_and_save_volumes()
    return ok, vols


# ---------------------------------------------------------------------------------------
# N1 (Solver.clustering head), N3 (label masks), N4 (docking masks)
# ---------------------------------------------------------------------------------------
def solver_clustering(ca_prob, bb_prob, aa_prob, aa_pred, ca_score_thrh=0.3, cluster_eps=10,
                      cluster_min_points=10, nms_radius=9):
    """``Solver.clustering`` (utils/modeler.py:762-899), unmodified, on in-memory volumes.

    The method is called unbound on a bare namespace carrying exactly the attributes it reads
    (the Solver constructor needs a full modelling configuration, FASTA files and Phenix).
    ``open3d`` is absent: DBSCAN comes from oracle/standins/open3d (scikit-learn).  Returns a
    dict with every intermediate the reference leaves behind."""
    import logging
    import types
    _setup_path()
    if _STANDINS not in sys.path:                    # open3d / superpose3d stand-ins
        sys.path.insert(0, _STANDINS)
    import utils.modeler as modeler
    log = logging.getLogger('ref_harness.clustering')
    log.disabled = True
    s = types.SimpleNamespace(
        logger=log, cluster_eps=cluster_eps, cluster_min_points=cluster_min_points, nms_radius=nms_radius,
        modeling_config=types.SimpleNamespace(CA_score_thrh=ca_score_thrh),
        CAProb=ca_prob, AAPred=aa_pred, neighbors2to6=[], neighbors0to6=[], neighbors0to7=[], neighbors2to7=[])
    modeler.NNPred.BBProb, modeler.NNPred.AAProb = bb_prob, aa_prob
    with _quiet():
        modeler.Solver.clustering(s)
    return dict(CA_cands=s.CA_cands, CA_cands_AAProb=s.CA_cands_AAProb, CA_cands_AA=s.CA_cands_AA,
                CAProb_clusted=modeler.NNPred.CAProb_clusted, cand_self_dis=s.cand_self_dis,
                neigh_mat=s.neigh_mat, best_neigh=[list(map(int, b)) for b in s.best_neigh],
                neighbors2to6=[np.asarray(v) for v in s.neighbors2to6],
                neighbors0to6=[np.asarray(v) for v in s.neighbors0to6],
                neighbors0to7=[np.asarray(v) for v in s.neighbors0to7],
                neighbors2to7=[np.asarray(v) for v in s.neighbors2to7])


def label_masks(normalized_map_path, pdb_path):
    """scripts_for_training_data/create_{backbone,carbon_alpha,amino_acid}_mask.py
    ``generate_mask`` (:120-177, :120-177, :128-183), unmodified.  Returns three int32 volumes."""
    _setup_path()
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, 'scripts_for_training_data'))
    import create_backbone_mask as mb
    import create_carbon_alpha_mask as mc
    import create_amino_acid_mask as ma
    with _quiet():
        return (mb.BackboneMask(normalized_map_path).generate_mask(pdb_path),
                mc.CarbonAlphaMask(normalized_map_path).generate_mask(pdb_path),
                ma.AminoAcidMaskGenerator(normalized_map_path).generate_mask(pdb_path))


def docking_masks(src, voxel_xyz, origin_xyz, pdb_path, workdir, contour_level, radius=2.0, percentage=40,
                  centroid_method='median'):
    """``PhenixDockingProcessor.initial_map_processing`` then ``subsequent_map_processing``
    (utils/dock_in_map.py:248-283, 285-364), unmodified, called unbound (the constructor wants a
    Phenix installation).  Returns (thresholded map, masked map, voxel size as stored) read back from the MRCs."""
    import logging
    import types
    _setup_path()
    from utils.dock_in_map import PhenixDockingProcessor as P
    log = logging.getLogger('ref_harness.docking')
    log.disabled = True
    me = types.SimpleNamespace(logger=log)
    inp = os.path.join(workdir, 'dock_in.mrc')
    thr = os.path.join(workdir, 'dock_thr.mrc')
    out = os.path.join(workdir, 'dock_masked.mrc')
    _write_mrc(inp, src, voxel_xyz, origin_xyz)
    with _quiet():
        P.initial_map_processing(me, inp, thr, contour_level)
        P.subsequent_map_processing(me, thr, pdb_path, out, radius=radius, percentage=percentage,
                                    centroid_method=centroid_method)
    import mrcfile
    with mrcfile.open(thr) as m:                     # what the reference saw: cella / m{x,y,z} in float32
        vs = (np.float32(m.voxel_size.x), np.float32(m.voxel_size.y), np.float32(m.voxel_size.z))
    return _read_mrc(thr), _read_mrc(out), vs
