"""The reference's own CPU path for the whole hot path, run AS IT IS: Solver.getData + Solver.nnPred
(utils/modeler.py:673-760) with their .mrc / .npz files and worker pools, the model replaced by a
ring of pre-generated logits (north_star keeps the model out of the path, and the GPU arm replaces it
the same way).

TEST / BASELINE INFRASTRUCTURE ONLY: used by ``bench.py --impl reference`` and the ``cpu_baseline``
leg.  The reference modules are imported unmodified -- from /root/reference in the build container,
from oracle/_ref/reference_py.zip (oracle/make_ref.py) on the GPU box -- with the I/O-only stand-ins of
oracle/standins for the absent ``mrcfile`` / ``Bio.PDB`` packages."""
from __future__ import annotations

import os
import shutil
import tempfile
import time

import numpy as np

from . import ref_harness


class RingModel:
    """Stands where MICA stands in run_inference (utils/predict.py:339): returns the first ``b`` cubes of a
    fixed ring of logits, whatever the input."""

    def __init__(self, ring):
        import torch
        self.ring = [torch.from_numpy(np.ascontiguousarray(a)) for a in ring]

    def eval(self):
        return self

    def __call__(self, x, af3):
        b = x.shape[0]
        return tuple(t[:b] for t in self.ring)


def scratch_root():
    """tmpfs when there is one (the reference's per-cube files are pure overhead to be measured, not the disk)."""
    return '/dev/shm' if os.path.isdir('/dev/shm') and os.access('/dev/shm', os.W_OK) else None


def prepare(workdir, src, voxel_xyz, structure):
    """Untimed set-up: the input map as MRC and the docked structure as PDB, laid out as run.py:108-112."""
    from mica_b200 import synthetic
    case = os.path.join(workdir, 'input', 'ID')
    os.makedirs(os.path.join(case, 'AF3_results'), exist_ok=True)
    map_path = os.path.join(case, 'map.mrc')
    ref_harness._write_mrc(map_path, src, voxel_xyz)
    pdb_path = os.path.join(case, 'ID_af3_docked.pdb')
    synthetic.write_pdb(pdb_path, dict(structure, hetero=np.zeros(len(structure['coords']), bool),
                                       res_ids=structure.get('res_ids', np.arange(len(structure['coords'])))))
    return dict(map_path=map_path, pdb_path=pdb_path, af3_results=os.path.join(case, 'AF3_results') + '/',
                grids=os.path.join(case, 'grids') + '/', out=os.path.join(workdir, 'out'))


def run_as_is(paths, ring, grid_size=48, padding=8, parallel=True):
    """One pass of the unmodified reference: DataPreprocessor -> GridCreator -> CryoEMPredictor.  Returns
    (seconds, volumes dict, cube count).  Files it leaves behind are removed afterwards (untimed), as
    Solver.nnPred does."""
    ref_harness._setup_path()
    from utils.create_grids import GridCreator
    from utils.predict import CryoEMPredictor
    from utils.preprocessing import DataPreprocessor
    model = RingModel(ring)
    t0 = time.perf_counter()
    with ref_harness._quiet():
        dp = DataPreprocessor(map_path=paths['map_path'], AF3_results=paths['af3_results'], quiet=True)
        dp.logger.disabled = True
        dp.resample_and_normalize_map()
        if dp.normalized_map_path is None or not os.path.exists(dp.normalized_map_path):
            raise RuntimeError('reference normalisation failed')
        ok = dp.create_AF3_encodings(paths['pdb_path'])
        gc = GridCreator(quiet=True)
        gc.logger.disabled = True
        r1 = gc.create_normalized_map_grids(dp.normalized_map_path,
                                            os.path.join(paths['grids'], 'normalized_map_grids'), grid_size, padding)
        if ok:
            gc.create_AF3_encodings_grids(dp.AF3_encodings, os.path.join(paths['grids'], 'AF3_encoding_grids'),
                                          grid_size, padding, parallel=parallel)
        pr = CryoEMPredictor(model_path='unused', grids_path=paths['grids'], output_path=paths['out'],
                             save_output=False, device='cpu', quiet=True)
        pr.logger.disabled = True
        if not pr.select_processing_strategy():
            raise RuntimeError('reference strategy selection failed')
        pr.model = model
        good, loader = pr.prepare_data()
        if not good or not pr.run_inference(loader):
            raise RuntimeError('reference inference failed')
        good, vols = pr.reconstruct_and_save_volumes()
    dt = time.perf_counter() - t0
    for d in (paths['grids'], os.path.join(paths['out'], 'results'), getattr(dp, 'AF3_encodings', None)):
        if d and os.path.isdir(d):
            shutil.rmtree(d, ignore_errors=True)
    if dp.normalized_map_path and os.path.exists(dp.normalized_map_path):
        os.remove(dp.normalized_map_path)
    return dt, vols, r1['grid_count']


def make_workdir():
    return tempfile.mkdtemp(prefix='mica_ref_', dir=scratch_root())
