"""Drop-in for the reference's ``utils/preprocessing.py::DataPreprocessor``.

Same constructor, methods, return values and never-raise convention
(utils/preprocessing.py:42,80,225 and SURVEY.md 8b); the arithmetic runs in
libmica_b200.so on the GPU and the results stay resident (mica_b200.session) for
GridCreator / CryoEMPredictor.  The .mrc side effects of the reference are kept
(``write_files=True``) because they are documented outputs of these methods."""
from __future__ import annotations

import logging
import os

import numpy as np
import torch

from . import mrc, ops, pdb, session
from .pipeline import MapHeader, MapPipeline


class DataPreprocessor:
    def __init__(self, map_path, AF3_results, normalized_map_path=None, quiet=True, device='cuda',
                 write_files=True, order=3):
        self.map_path = map_path
        self.AF3_results = AF3_results
        self.normalized_map_path = normalized_map_path
        self.quiet = quiet
        self.device = torch.device(device)
        self.write_files = write_files
        self.order = order
        self.logger = logging.getLogger(__name__)

    def print_clean(self, message):
        if not self.quiet:
            print(message)

    # utils/preprocessing.py:80-170
    def resample_and_normalize_map(self, target_voxel_size=1.0):
        success = False
        try:
            m = mrc.read_mrc(self.map_path)
            header = MapHeader(voxel_size=m.voxel_size, origin=m.origin, mapc=m.mapc, mapr=m.mapr, maps=m.maps,
                               nxstart=m.nxstart, nystart=m.nystart, nzstart=m.nzstart)
            pipe = MapPipeline(self.device, order=self.order, target_voxel_size=target_voxel_size)
            src = torch.from_numpy(np.array(m.data, dtype=np.float32)).to(self.device)
            if pipe.resample_and_normalize(src, header):
                out_path = os.path.join(os.path.dirname(self.AF3_results), 'resampled_normalized_map.mrc')
                out_header = MapHeader(voxel_size=(np.float32(target_voxel_size),) * 3, origin=m.origin,
                                       mapc=m.mapc, mapr=m.mapr, maps=m.maps, nxstart=m.nxstart,
                                       nystart=m.nystart, nzstart=m.nzstart)
                if self.write_files:
                    os.makedirs(os.path.dirname(out_path) or '.', exist_ok=True)
                    mrc.write_mrc(out_path, mrc.MrcMap(
                        data=pipe.normalized.cpu().numpy(), voxel_size=out_header.voxel_size, origin=m.origin,
                        mapc=m.mapc, mapr=m.mapr, maps=m.maps, nxstart=m.nxstart, nystart=m.nystart,
                        nzstart=m.nzstart))
                session.put(out_path, kind='normalized_map', volume=pipe.normalized, header=out_header,
                            median=pipe.median, p999=pipe.p999)
                self.normalized_map_path = out_path
                success = True
            else:
                reason = {1: 'No positive values found after thresholding',
                          2: 'Percentile value is zero - cannot normalize'}.get(pipe.norm_status, 'unknown')
                self.logger.error(f'Normalization failed: {reason}')
        except Exception as e:                                  # reference: swallowed and logged (:159-161)
            self.logger.error(f'Map processing failed: {e}')
        self.print_clean('Map successfully resampled and normalized' if success
                         else 'Map Resampling and Normalization Failed')

    def _reference_map(self):
        entry = session.get(self.normalized_map_path)
        if entry is not None:
            return tuple(entry['volume'].shape), entry['header']
        m = mrc.read_mrc(self.normalized_map_path)
        return m.data.shape, MapHeader(voxel_size=m.voxel_size, origin=m.origin, mapc=m.mapc, mapr=m.mapr,
                                       maps=m.maps, nxstart=m.nxstart, nystart=m.nystart, nzstart=m.nzstart)

    # utils/preprocessing.py:225-347
    def create_AF3_encodings(self, combined_docked_model_path):
        success = False
        try:
            shape, header = self._reference_map()
            coords, bb_ch, aa_ch, n_res = pdb.read_pdb_atoms(combined_docked_model_path)
            dev = self.device
            vol, status = ops.af3_encode(torch.from_numpy(coords).to(dev), torch.from_numpy(bb_ch).to(dev),
                                         torch.from_numpy(aa_ch).to(dev), header.origin, shape)
            if int(status.item()) != 0:
                raise IndexError('atom index out of bounds for the map axis (reference clip quirk)')
            self.AF3_encodings = os.path.join(os.path.dirname(self.AF3_results), 'AF3_encodings')
            if self.write_files:
                os.makedirs(self.AF3_encodings, exist_ok=True)
                host = vol.cpu().numpy()
                for c, name in enumerate(pdb.CHANNEL_NAMES):
                    mrc.write_mrc(os.path.join(self.AF3_encodings, f'{name}_encoding.mrc'), mrc.MrcMap(
                        data=host[c], voxel_size=(np.float32(1),) * 3, origin=header.origin, mapc=header.mapc,
                        mapr=header.mapr, maps=header.maps, nxstart=header.nxstart, nystart=header.nystart,
                        nzstart=header.nzstart))
            session.put(self.AF3_encodings, kind='af3_encodings', volume=vol, header=header,
                        atoms=len(coords), residues=n_res)
            success = True
        except Exception as e:                                  # reference: swallowed -> False (:344-347)
            self.print_clean(f'Encoding failed: AF3 encoding failed: {e}')
        return success
