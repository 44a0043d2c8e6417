"""Drop-ins for the two map-masking steps of the docking driver.

    PhenixDockingProcessor.initial_map_processing      utils/dock_in_map.py:248-283
    PhenixDockingProcessor.subsequent_map_processing   utils/dock_in_map.py:285-364

Same names, arguments, return value (the output path) and MRC written; ``DockingMapMasks`` can be mixed
into (or called from) the reference's processor, whose Phenix orchestration is out of scope.  The
contour threshold is one elementwise kernel; the radius mask zeroes the ball of every seed voxel
instead of running a Euclidean distance transform over the whole map -- identical result, the distance
being evaluated exactly as ``distance_transform_edt(~mask, sampling=voxel_size)`` does.  The choice of
the atoms (centroid, distance sort, ``percentage`` cut) is host logic over the atom list, as in the
reference."""
from __future__ import annotations

import ctypes as C
import logging
import os

import numpy as np
import torch

from . import mrc, ops, pdb
from ._lib import lib, check
from .ops import _dev, _stream, device_guard

_F3 = C.c_float * 3


@device_guard
def contour_threshold(data: torch.Tensor, contour_level: float, out: torch.Tensor | None = None) -> torch.Tensor:
    """``np.where(data < contour_level, 0, data)`` (utils/dock_in_map.py:269); float32 comparison."""
    p = _dev(data, torch.float32, 'data')
    if out is None:
        out = torch.empty_like(data)
    check(lib.mica_contour_threshold_f32(p, _dev(out, torch.float32, 'out'), data.numel(),
                                         float(np.float32(contour_level)), _stream()), 'contour_threshold')
    return out


@device_guard
def zero_around_atoms(map_data: torch.Tensor, coords: torch.Tensor, voxel_size_xyz, origin_xyz, radius=2.0):
    """utils/dock_in_map.py:330-352 in place on ``map_data`` (device float32 [nz,ny,nx]).  Returns the device
    status word (1 where the reference's ``mask[z, y, x] = True`` would raise IndexError)."""
    p = _dev(map_data, torch.float32, 'map_data')
    nz, ny, nx = (int(v) for v in map_data.shape)
    n = int(coords.shape[0])
    p_xyz = _dev(coords, torch.float32, 'coords') if n else None
    status = torch.zeros(1, dtype=torch.int32, device=map_data.device)
    check(lib.mica_zero_around_atoms(p_xyz, n, _F3(*(float(np.float32(v)) for v in origin_xyz)),
                                     _F3(*(float(np.float32(v)) for v in voxel_size_xyz)), float(radius), nz, ny, nx,
                                     p, C.c_void_p(status.data_ptr()), _stream()), 'zero_around_atoms')
    return status


def select_central_atoms(coords, percentage=40, centroid_method='median'):
    """utils/dock_in_map.py:314-327 (host logic on the atom list)."""
    coords = np.asarray(coords)
    if centroid_method == 'mean':
        centroid = np.mean(coords, axis=0)
    elif centroid_method == 'median':
        centroid = np.median(coords, axis=0)
    else:
        raise ValueError(f'Unknown centroid method: {centroid_method}')
    distances = np.sqrt(np.sum((coords - centroid) ** 2, axis=1))
    n_use = int(len(coords) * (percentage / 100.0))
    return coords[np.argsort(distances)[:n_use]]


class DockingMapMasks:
    """The two map-processing methods of ``PhenixDockingProcessor`` (utils/dock_in_map.py:36)."""

    def __init__(self, device='cuda', logger=None):
        ops.require_gpu()
        self.device = torch.device(device)
        self.logger = logger or logging.getLogger(__name__)

    def initial_map_processing(self, input_map, output_map_path, contour_level):
        self.logger.info(f'Processing initial map with contour level {contour_level}')
        try:
            m = mrc.read_mrc(input_map)
            data = torch.from_numpy(np.array(m.data, dtype=np.float32)).to(self.device)
            clipped = contour_threshold(data, contour_level)
            # the reference copies voxel_size and origin only (:272-275)
            mrc.write_mrc(output_map_path, mrc.MrcMap(data=clipped.cpu().numpy(), voxel_size=m.voxel_size,
                                                      origin=m.origin))
            return output_map_path
        except Exception as e:
            self.logger.error(f'Error processing initial map: {str(e)}')
            raise

    def subsequent_map_processing(self, input_map_path, pdb_file_path, output_map_path, radius=2.0, percentage=40,
                                  centroid_method='median'):
        model_name = os.path.basename(pdb_file_path)
        try:
            coords = pdb.read_pdb_records(pdb_file_path)['coords']
            selected = select_central_atoms(coords, percentage, centroid_method)
            m = mrc.read_mrc(input_map_path)
            data = torch.from_numpy(np.array(m.data, dtype=np.float32)).to(self.device)
            sel = torch.from_numpy(np.ascontiguousarray(selected, dtype=np.float32).reshape(-1, 3)).to(self.device)
            status = zero_around_atoms(data, sel, m.voxel_size, m.origin, radius)
            if int(status.item()) != 0:
                raise IndexError('index is out of bounds for the map axis')
            mrc.write_mrc(output_map_path, mrc.MrcMap(data=data.cpu().numpy(), voxel_size=m.voxel_size,
                                                      origin=m.origin))
            return output_map_path
        except Exception as e:
            self.logger.error(f'Error in map masking for {model_name}: {str(e)}')
            raise
