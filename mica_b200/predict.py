"""Drop-in for the reference's ``utils/predict.py::CryoEMPredictor``.

Same constructor and ``run_prediction() -> (bool, {four volumes})`` contract
(utils/predict.py:48,589; consumer: Solver.nnPred, utils/modeler.py:722-734).  The
model (models/model.py::MICA, PyTorch convolutions) is used as it is; everything
around it -- cube assembly, softmax/argmax, stitching -- runs in libmica_b200.so with
no per-cube file: cubes come from the volumes GridCreator registered (or, when only
the reference's .npz files exist on disk, are uploaded from them)."""
from __future__ import annotations

import glob
import logging
import os
import time

import numpy as np
import torch

from . import ops, pdb, session

MAP_TYPES = ('backbone_probability', 'carbon_alpha_probability', 'amino_acid_prediction',
             'amino_acid_probability')


class CryoEMPredictor:
    def __init__(self, model_path, grids_path, output_path, save_output=True, device='cuda', quiet=False,
                 model=None, keep_on_device=False, host_volumes=MAP_TYPES):
        """``keep_on_device`` / ``host_volumes`` are additions to the reference signature: with
        ``keep_on_device`` the stitched volumes stay registered in HBM (``session`` key
        ``<output_path>/results/device_volumes``) for ``mica_b200.candidates.clustering_head``, and
        ``host_volumes`` names the volumes copied to the host (default: all four, as the reference returns);
        leaving out ``amino_acid_probability`` saves 20 of the 23 channels of PCIe traffic."""
        self.keep_on_device = bool(keep_on_device)
        self.host_volumes = tuple(host_volumes)
        self.model_path = model_path
        self.grids_path = grids_path
        self.output_path = output_path
        self.temp_output_path = os.path.join(output_path, 'results', 'predicted_grids')
        parts = str(grids_path).split('/')
        self.reconstruction_path = os.path.join(output_path, 'results', parts[-2] if len(parts) > 1 else parts[-1])
        self.save_output = save_output.lower() == 'true' if isinstance(save_output, str) else bool(save_output)
        self.device = device
        self.quiet = quiet
        self.batch_threshold = 200                      # utils/predict.py:72
        self.model = model
        self.use_optimized_batching = False
        self.sample_count = 0
        self.optimal_batch_size = 1
        self.timing_stats = {k: 0 for k in ('strategy_selection', 'model_loading', 'data_loading', 'inference',
                                            'reconstruction', 'saving', 'total')}
        self.logger = logging.getLogger(__name__)
        self._source = None

    def _print_status(self, message):
        if not self.quiet:
            print(message)

    # ------------------------------------------------------------------ data source
    def _map_dir(self):
        return os.path.join(str(self.grids_path).rstrip('/'), 'normalized_map_grids')

    def _af3_dir(self):
        return os.path.join(str(self.grids_path).rstrip('/'), 'AF3_encoding_grids')

    def _resolve_source(self):
        """Resident volumes registered by GridCreator, else the reference's files on disk."""
        m = session.get(self._map_dir())
        if m is not None:
            a = session.get(self._af3_dir())
            self._source = {'kind': 'resident', 'map': m, 'af3': a, 'ijk': m['ijk'], 'cube_shape': m['cube_shape'],
                            'grid_size': m['grid_size'], 'padding': m['padding']}
            return len(m['ijk'])
        files = sorted(glob.glob(os.path.join(self._map_dir(), '*.npz')))
        if not files:
            return 0
        first = np.load(files[0])
        gs, pad = int(first['grid_size']), int(first['padding'])
        ijk = np.array([[int(np.load(f)[k]) for k in ('i', 'j', 'k')] for f in files], dtype=np.int32)
        self._source = {'kind': 'files', 'files': files, 'ijk': ijk,
                        'cube_shape': tuple(int(v) for v in np.asarray(first['orig_shape']).ravel()),
                        'grid_size': gs, 'padding': pad}
        return len(files)

    # utils/predict.py:156-215
    def _calculate_optimal_batch_size(self):
        per_sample_gb = (64 ** 3 * 25 * 5) / 1024 ** 3
        try:
            free = torch.cuda.get_device_properties(torch.device(self.device)).total_memory / 1024 ** 3 * 0.7
        except Exception:
            return 1
        return int(min(max(1, free / per_sample_gb), 8))      # capped at 8 as in the reference (:174)

    def select_processing_strategy(self):
        t0 = time.time()
        try:
            self.sample_count = self._resolve_source()
            if self.sample_count == 0:
                self.logger.error(f'No grid files found in: {self._map_dir()}/')
                return False
            self.use_optimized_batching = self.sample_count > self.batch_threshold
            if self.use_optimized_batching:
                self.optimal_batch_size = self._calculate_optimal_batch_size()
            return True
        except Exception as e:
            self.logger.error(f'Strategy selection failed: {e}')
            return False
        finally:
            self.timing_stats['strategy_selection'] = time.time() - t0

    # utils/predict.py:217-258
    def load_model(self):
        t0 = time.time()
        try:
            if self.model is not None:
                return True
            if not os.path.exists(self.model_path):
                self.logger.error(f'Model file not found: {self.model_path}')
                return False
            from models.model import MICA                       # the reference's model, unchanged
            model = MICA().to(self.device)
            checkpoint = torch.load(self.model_path, map_location=self.device)
            state = {k.replace('module.', ''): v.to(self.device) for k, v in checkpoint['model_state_dict'].items()}
            model.load_state_dict(state, strict=False)
            model.eval()
            self.model = model
            return True
        except Exception as e:
            self.logger.error(f'Model loading failed: {e}')
            return False
        finally:
            self.timing_stats['model_loading'] = time.time() - t0

    # ------------------------------------------------------------------ inference + stitch
    def _batches(self, B):
        n = self.sample_count
        return [(b0, min(n, b0 + B)) for b0 in range(0, n, B)]

    def _fetch(self, b0, b1, dev):
        """(exp_map [B,1,W^3], af_features [B,24,W^3], nonzero flags [B]) on the device."""
        s = self._source
        gs, pad = s['grid_size'], s['padding']
        W = gs + 2 * pad
        B = b1 - b0
        if s['kind'] == 'resident':
            ijk = self._ijk_dev[b0:b1]
            x = ops.extract_cubes(s['map']['volume'], ijk, gs, pad, s['map']['perm'])
            flags = torch.zeros(B, dtype=torch.int32, device=dev)
            if s['af3'] is not None:
                af = ops.extract_cubes(s['af3']['volume'], ijk, gs, pad, s['af3']['perm'], nonzero=flags)
            else:
                af = torch.zeros((B, 24, W, W, W), dtype=torch.float32, device=dev)
            return x, af, flags
        xs, afs = [], []
        for f in s['files'][b0:b1]:                              # dataset/dataset.py:194-224
            xs.append(np.load(f)['grid'])
            try:
                feats = []
                for name in pdb.CHANNEL_NAMES:
                    p = f.replace('normalized_map_grids', f'AF3_encoding_grids/{name}_grids')
                    feats.append(np.load(p.replace('normalized_map', name))['grid'])
                afs.append(np.stack(feats))
            except Exception:
                afs.append(np.zeros((24, W, W, W), np.float32))
        x = torch.from_numpy(np.stack(xs)[:, None].astype(np.float32)).to(dev)
        af = torch.from_numpy(np.stack(afs).astype(np.float32)).to(dev)
        flags = (af.abs().reshape(B, -1).sum(1) > 0).to(torch.int32)
        return x, af, flags

    def run_inference_and_reconstruct(self):
        """run_inference + reconstruct_volume (utils/predict.py:307-512) fused: logits are
        post-processed and their cores pasted into the four volumes as each batch finishes."""
        dev = torch.device(self.device)
        s = self._source
        gs, pad = s['grid_size'], s['padding']
        self._ijk_dev = torch.from_numpy(np.ascontiguousarray(s['ijk'], dtype=np.int32)).to(dev)
        vols = ops.StitchedVolumes(s['cube_shape'], dev)
        B = self.optimal_batch_size if self.use_optimized_batching else 1
        model = self.model
        if hasattr(model, 'eval'):
            model.eval()
        with torch.no_grad():
            for b0, b1 in self._batches(B):
                x, af, flags = self._fetch(b0, b1, dev)
                ijk = self._ijk_dev[b0:b1]
                if b1 - b0 == 1:
                    groups = [torch.arange(1, device=dev)]
                else:
                    # D8: MICA tests `af_features.abs().sum() < 1e-6` over the whole batch
                    # (models/model.py:60-63); cubes without AF3 signal get their own batch so the
                    # result does not depend on what they happen to be batched with.
                    nz = flags != 0
                    groups = [g for g in (torch.nonzero(nz).flatten(), torch.nonzero(~nz).flatten()) if len(g)]
                for g in groups:
                    whole = len(g) == (b1 - b0)
                    gx, gaf, gijk = (x, af, ijk) if whole else (x[g].contiguous(), af[g].contiguous(),
                                                               ijk[g].contiguous())
                    bb, ca, aa = model(gx, gaf)
                    ops.postproc_stitch(bb.contiguous().float(), ca.contiguous().float(), aa.contiguous().float(),
                                        gijk, vols, gs, pad)
        return vols

    # utils/predict.py:589-634
    def run_prediction(self):
        t_total = time.time()
        try:
            if not self.select_processing_strategy():
                return False, {}
            if not self.load_model():
                return False, {}
            t0 = time.time()
            vols = self.run_inference_and_reconstruct()
            torch.cuda.synchronize()
            self.timing_stats['inference'] = time.time() - t0
            t0 = time.time()
            if self.keep_on_device:
                session.put(os.path.join(str(self.output_path), 'results', 'device_volumes'), **vols.as_dict())
            volumes = {k: v.cpu().numpy() for k, v in vols.as_dict().items() if k in self.host_volumes}
            self.timing_stats['reconstruction'] = time.time() - t0
            if self.save_output:
                t0 = time.time()
                os.makedirs(self.reconstruction_path, exist_ok=True)
                for k in MAP_TYPES:
                    if k in volumes:
                        np.save(f'{self.reconstruction_path}/{k}.npy', volumes[k])
                self.timing_stats['saving'] = time.time() - t0
            self.timing_stats['total'] = time.time() - t_total
            return True, volumes
        except Exception as e:                                   # reference: logged, (False, {}) (:632-634)
            self.logger.error(f'Prediction pipeline failed: {e}')
            return False, {}
