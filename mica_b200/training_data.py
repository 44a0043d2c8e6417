"""Drop-ins for the reference's training-data builders (scripts_for_training_data/).

    create_normalized_map.py           -> MapProcessor            (resample + normalise, R1/R2 twins)
    create_AF3_encodings.py            -> FeaturesEncoder         (24-channel rasteriser, R4 twin)
    create_grids_for_normalized_map.py -> create_and_save_grids(..., min_max=0.01)
    create_grids_for_{AF3_encodings,BB_mask,CA_mask,AA_mask}.py -> create_and_save_grids(...)   (R5 twins)

Same class / method / function names, arguments, return values and files written as the
reference scripts; the arithmetic runs in libmica_b200.so on the GPU.  Differences to the
inference path that these twins keep: cubes are cut from the raw (z,y,x)
array -- no axis transpose (create_grids_for_normalized_map.py:40-54) --, the map variant
drops cubes whose maximum is < 0.01 (:78), and FeaturesEncoder lets the IndexError of the
mis-ordered clip (D7) propagate to its caller (create_AF3_encodings.py:181-188 catches it).
The scripts' ``main()`` loops over ``Training_Dataset/Raw_Data/*`` are ``build_*`` below."""
from __future__ import annotations

import os
from glob import glob

import numpy as np
import torch

from . import mrc, ops, pdb
from ._lib import NORM_OK

_REC = [('x', '<f4'), ('y', '<f4'), ('z', '<f4')]


def _rec(xyz):
    """mrcfile hands voxel_size / origin out as (x, y, z) float32 records."""
    return np.rec.array(tuple(np.float32(v) for v in xyz), dtype=_REC)


def _device(device):
    ops.require_gpu()
    return torch.device(device)


class MapProcessor:
    """scripts_for_training_data/create_normalized_map.py:19-115."""

    def __init__(self, input_map, device='cuda', order=3):
        self.input_map = input_map
        self.device = _device(device)
        self.order = order
        m = mrc.read_mrc(input_map)
        self._map = m
        self.data = m.data
        self.voxel_size = _rec(m.voxel_size)
        self.origin = _rec(m.origin)
        self.mapc, self.mapr, self.maps = m.mapc, m.mapr, m.maps
        self.nxstart, self.nystart, self.nzstart = m.nxstart, m.nystart, m.nzstart

    def _resample_device(self):
        src = torch.from_numpy(np.array(self.data, dtype=np.float32)).to(self.device)
        zf = [np.float32(self.voxel_size.x), np.float32(self.voxel_size.y), np.float32(self.voxel_size.z)]  # :40
        if all(float(z) == 1.0 for z in zf):                       # SciPy early exit (D10)
            return src.clone()
        return ops.resample(src, ops.zoom_output_shape(src.shape, zf), order=self.order)

    def resample(self, target_voxel_size=1.0):
        """:37-46 -- zoom(data, [vx, vy, vz], order=3); returns the resampled array (host)."""
        self._resampled_dev = self._resample_device()
        self.resampled_data = self._resampled_dev.cpu().numpy()
        self.target_voxel_size = target_voxel_size
        return self.resampled_data

    def normalize(self, data=None):
        """:48-79 -- returns the normalised array, or None (after printing the reference's
        message) when there is no positive value or the percentile is zero."""
        if data is None:
            dev = getattr(self, '_resampled_dev', None)
            if dev is None:
                dev = torch.from_numpy(np.array(self.data, dtype=np.float32)).to(self.device)
        else:
            dev = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).to(self.device)
        norm, stats = ops.normalize(dev)
        self.median, self.percentile_value, _, status = stats.result()
        if status != NORM_OK:
            print('Error during normalization!!!')
            return None
        self._normalized_dev = norm
        self.normalized_data = norm.cpu().numpy()
        return self.normalized_data

    def process_map(self, output_path, target_voxel_size=1.0):
        """:81-97"""
        self.resample(target_voxel_size)
        if self.normalize() is None:
            print('Processing failed during normalization')
            return
        self.save_map(output_path)

    def save_map(self, output_path):
        """:99-115"""
        if not hasattr(self, 'normalized_data'):
            raise ValueError('No processed data available to save')
        mrc.write_mrc(output_path, mrc.MrcMap(
            data=self.normalized_data.astype(np.float32), voxel_size=(np.float32(1),) * 3,
            origin=self._map.origin, mapc=self.mapc, mapr=self.mapr, maps=self.maps,
            nxstart=self.nxstart, nystart=self.nystart, nzstart=self.nzstart))


class FeaturesEncoder:
    """scripts_for_training_data/create_AF3_encodings.py:19-125."""

    def __init__(self, map_file, device='cuda'):
        self.device = _device(device)
        m = mrc.read_mrc(map_file)
        self._map = m
        self.map_data = m.data
        self.shape = m.data.shape
        self.voxel_size = _rec(m.voxel_size)
        self.origin = _rec(m.origin)
        self.mapc, self.mapr, self.maps = m.mapc, m.mapr, m.maps
        self.nxstart, self.nystart, self.nzstart = m.nxstart, m.nystart, m.nzstart
        self.backbone_atoms = list(pdb.BACKBONE_ATOMS)
        self.amino_acids = list(pdb.AMINO_ACIDS)
        self.num_channels = len(self.backbone_atoms) + len(self.amino_acids)

    def transform_coordinates(self, coord):
        """:52-60 (host helper, one coordinate; the kernel applies the same arithmetic per atom:
        float32 subtract, round half to even, clip with (nz,ny,nx) bounds against (x,y,z))."""
        shifted = np.asarray(coord, dtype=np.float32) - np.array((self.origin.x, self.origin.y, self.origin.z))
        indices = np.round(shifted / 1.0).astype(int)
        return np.clip(indices, 0, np.array(self.shape) - 1)

    def get_aa_channel_index(self, residue_name):
        try:
            return len(self.backbone_atoms) + self.amino_acids.index(residue_name)
        except ValueError:
            return -1

    def encode_structure(self, pdb_file):
        """:68-106 -- returns the (24, nz, ny, nx) occupancy volume (host float32; the reference
        builds float64 and casts to float32 when saving).  Raises IndexError where the reference's
        ``feature_volume[..., idx[2], idx[1], idx[0]] = 1`` would (non-cubic maps, D7)."""
        coords, bb_ch, aa_ch, _ = pdb.read_pdb_atoms(pdb_file)
        dev = self.device
        vol, status = ops.af3_encode(torch.from_numpy(coords).to(dev), torch.from_numpy(bb_ch).to(dev),
                                     torch.from_numpy(aa_ch).to(dev), self._map.origin, self.shape)
        if int(status.item()) != 0:
            raise IndexError('index is out of bounds for the map axis (clip bounds are (nz,ny,nx) against (x,y,z))')
        self._feature_dev = vol
        return vol.cpu().numpy()

    def save_channel_as_mrc(self, feature_volume, output_path, channel_idx=0):
        """:108-121"""
        mrc.write_mrc(output_path, mrc.MrcMap(
            data=np.asarray(feature_volume[channel_idx], dtype=np.float32), voxel_size=(np.float32(1),) * 3,
            origin=self._map.origin, mapc=self.mapc, mapr=self.mapr, maps=self.maps,
            nxstart=self.nxstart, nystart=self.nystart, nzstart=self.nzstart))

    def get_channel_names(self):
        return self.backbone_atoms + self.amino_acids


def create_and_save_grids(mrc_file, output_dir, grid_size=48, padding=8, min_max=None, device='cuda', batch=64):
    """The ``create_and_save_grids`` of the five create_grids_for_*.py scripts: cut W^3 windows
    (W = grid_size + 2 * padding) at stride grid_size from the raw (z,y,x) array, zero outside the
    map, one ``grid_i{i}_j{j}_k{k}.npz`` per window with the reference's keys.  ``min_max=0.01`` is
    the normalised-map variant (create_grids_for_normalized_map.py:78: windows whose maximum is
    below it are not written).  Returns the number of files written."""
    dev = _device(device)
    os.makedirs(output_dir, exist_ok=True)
    m = mrc.read_mrc(mrc_file)
    dtype = m.data.dtype
    vol = torch.from_numpy(np.array(m.data, dtype=np.float32)).to(dev)
    orig_shape = tuple(int(v) for v in m.data.shape)
    perm = (0, 1, 2)                                              # no transpose in the training builders
    ijk = ops.cube_origins(orig_shape, grid_size)
    d_ijk = torch.from_numpy(ijk).to(dev)
    voxel, origin = _rec(m.voxel_size), _rec(m.origin)
    grid_count = 0
    for b0 in range(0, len(ijk), batch):
        sel = d_ijk[b0:b0 + batch]
        cmax = torch.empty(sel.shape[0], dtype=torch.float32, device=dev)
        cubes = ops.extract_cubes(vol, sel, grid_size, padding, perm, cube_max=cmax)
        keep = np.ones(sel.shape[0], bool) if min_max is None else (cmax.cpu().numpy() >= np.float32(min_max))
        host = cubes[:, 0].cpu().numpy()
        for n, (i, j, k) in enumerate(ijk[b0:b0 + batch]):
            if not keep[n]:
                continue
            i, j, k = int(i), int(j), int(k)
            np.savez(os.path.join(output_dir, f'grid_i{i}_j{j}_k{k}.npz'), grid=host[n].astype(dtype, copy=False),
                     i=i, j=j, k=k, di=min(grid_size, orig_shape[0] - i), dj=min(grid_size, orig_shape[1] - j),
                     dk=min(grid_size, orig_shape[2] - k), orig_shape=orig_shape, grid_size=grid_size,
                     padding=padding, voxel_size=voxel, origin=origin, mapc=np.int32(m.mapc),
                     mapr=np.int32(m.mapr), maps=np.int32(m.maps))
            grid_count += 1
    return grid_count


# ------------------------------------------------------------------ the scripts' main() loops
def build_normalized_maps(base_dir='Training_Dataset/Raw_Data', output_dir='Training_Dataset/Processed_Data',
                          device='cuda'):
    """create_normalized_map.py:143-160"""
    done = 0
    for i, directory in enumerate(sorted(glob(f'{base_dir}/*'))):
        emd_id = directory.split('/')[-1]
        try:
            out = f'{output_dir}/{emd_id}'
            os.makedirs(out, exist_ok=True)
            MapProcessor(f'{directory}/emd_{emd_id}.map', device=device).process_map(
                f'{out}/resampled_normalized_map.mrc', target_voxel_size=1.0)
            print(f'Created resampled and normalized density map for EMD ID: {emd_id} | Completed {i + 1} density maps ...')
            done += 1
        except Exception:
            print(f'Failed for normalizing map for EMD ID {emd_id}')
    return done


def build_af3_encodings(base_dir='Training_Dataset/Raw_Data', output_dir='Training_Dataset/Processed_Data',
                        device='cuda'):
    """create_AF3_encodings.py:160-188"""
    done = 0
    for i, directory in enumerate(sorted(glob(f'{base_dir}/*'))):
        emd_id = directory.split('/')[-1]
        pdb_files = glob(f'{directory}/*af3_docked*.pdb')
        if not pdb_files:
            print(f'No AF3 docked PDB file found for EMD ID: {emd_id}')
            continue
        out = f'{output_dir}/{emd_id}'
        os.makedirs(out, exist_ok=True)
        try:
            enc = FeaturesEncoder(f'{out}/resampled_normalized_map.mrc', device=device)
            vol = enc.encode_structure(pdb_files[0])
            for j, name in enumerate(enc.get_channel_names()):
                enc.save_channel_as_mrc(vol, f'{out}/{name}_encoding.mrc', channel_idx=j)
            print(f'Generated feature encodings for EMD ID: {emd_id} | Completed {i + 1} density maps ...')
            done += 1
        except Exception as e:
            print(f'Failed for density map with EMD ID: {emd_id} - Error: {str(e)}')
    return done


def build_grids(kind, base_dir='Training_Dataset/Processed_Data', output_dir=None, grid_size=48, padding=8,
                device='cuda'):
    """The main() loops of the five create_grids_for_*.py scripts.  ``kind`` is one of
    'normalized_map', 'AF3_encodings', 'BB_mask', 'CA_mask', 'AA_mask'."""
    single = {'normalized_map': ('resampled_normalized_map.mrc', 'Training_Dataset/Grids/normalized_maps', 0.01),
              'BB_mask': ('backbone_mask.mrc', 'Training_Dataset/Grids/BB_masks', None),
              'CA_mask': ('carbon_alpha_mask.mrc', 'Training_Dataset/Grids/CA_masks', None),
              'AA_mask': ('amino_acid_mask.mrc', 'Training_Dataset/Grids/AA_masks', None)}
    total = 0
    for directory in sorted(glob(f'{base_dir}/*')):
        emd_id = directory.split('/')[-1]
        try:
            if kind == 'AF3_encodings':                    # create_grids_for_AF3_encodings.py:149-167
                root = output_dir or 'Training_Dataset/Grids'
                for f in glob(f'{directory}/*encoding*.mrc'):
                    enc_type = os.path.basename(f).split('_')[0]
                    total += create_and_save_grids(f, f'{root}/{enc_type}_encodings/{emd_id}', grid_size, padding,
                                                   device=device)
            else:
                fname, default_out, min_max = single[kind]
                total += create_and_save_grids(f'{directory}/{fname}', f'{output_dir or default_out}/{emd_id}',
                                               grid_size, padding, min_max=min_max, device=device)
        except Exception as e:
            print(f'Grid creation failed for EMD ID: {emd_id} - Error: {str(e)}')
    return total
