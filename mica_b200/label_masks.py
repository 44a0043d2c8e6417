"""Drop-ins for the reference's training label-mask builders.

    scripts_for_training_data/create_backbone_mask.py      -> BackboneMask
    scripts_for_training_data/create_carbon_alpha_mask.py  -> CarbonAlphaMask
    scripts_for_training_data/create_amino_acid_mask.py    -> AminoAcidMaskGenerator

Same class / method names, arguments, return values (an int32 (nz,ny,nx) volume) and files written
(float32 MRC with the map's header, ``save_mask``).  The per-atom Python loops and position
dictionaries are replaced by order-free CUDA kernels (mica_b200/csrc/masks.cu) that reproduce the
reference's last-writer-wins / sequential-zeroing results bit for bit.  Like the reference,
``generate_mask`` raises (IndexError) when an atom index exceeds its real axis -- on non-cubic maps
the reference clips (x,y,z) with (nz,ny,nx)."""
from __future__ import annotations

import ctypes as C
import os
from glob import glob

import numpy as np
import torch

from . import mrc, ops, pdb
from ._lib import lib, check
from .ops import _stream, device_guard

AA_MAPPING = {'ALA': 1, 'CYS': 2, 'ASP': 3, 'GLU': 4, 'PHE': 5, 'GLY': 6, 'HIS': 7, 'ILE': 8, 'LYS': 9,
              'LEU': 10, 'MET': 11, 'ASN': 12, 'PRO': 13, 'GLN': 14, 'ARG': 15, 'SER': 16, 'THR': 17,
              'VAL': 18, 'TRP': 19, 'TYR': 20}                    # create_amino_acid_mask.py:40-45
_REC = [('x', '<f4'), ('y', '<f4'), ('z', '<f4')]


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else None


@device_guard
def class_mask(coords: torch.Tensor, is_class: torch.Tensor, origin_xyz, shape_zyx, clip_hi_xyz=None):
    """create_backbone_mask.py:136-172 on the device: (mask int32 [nz,ny,nx], status int32 [1])."""
    nz, ny, nx = (int(v) for v in shape_zyx)
    if not coords.is_cuda:
        raise ops._lib.MicaError('coords must be a CUDA tensor (mica_b200 has no CPU fallback)')
    if clip_hi_xyz is None:
        clip_hi_xyz = (nz - 1, ny - 1, nx - 1)
    dev = coords.device
    mask = torch.empty((nz, ny, nx), dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ox, oy, oz = (float(np.float32(v)) for v in origin_xyz)
    check(lib.mica_label_class_mask(_ptr(coords.contiguous()), _ptr(is_class.contiguous()), coords.shape[0], ox, oy, oz,
                                    int(clip_hi_xyz[0]), int(clip_hi_xyz[1]), int(clip_hi_xyz[2]), nz, ny, nx,
                                    _ptr(mask), _ptr(status), _stream()), 'label_class_mask')
    return mask, status


@device_guard
def aa_mask(ca_coords: torch.Tensor, labels: torch.Tensor, origin_xyz, shape_zyx, clip_hi_xyz=None):
    """create_amino_acid_mask.py:151-177 on the device: (mask int32 [nz,ny,nx], status int32 [1])."""
    nz, ny, nx = (int(v) for v in shape_zyx)
    if not ca_coords.is_cuda:
        raise ops._lib.MicaError('coords must be a CUDA tensor (mica_b200 has no CPU fallback)')
    if clip_hi_xyz is None:
        clip_hi_xyz = (nz - 1, ny - 1, nx - 1)
    dev = ca_coords.device
    mask = torch.empty((nz, ny, nx), dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = lib.mica_label_aa_mask_workspace_bytes(nz, ny, nx)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    ox, oy, oz = (float(np.float32(v)) for v in origin_xyz)
    check(lib.mica_label_aa_mask(_ptr(ca_coords.contiguous()), _ptr(labels.contiguous()), ca_coords.shape[0], ox, oy,
                                 oz, int(clip_hi_xyz[0]), int(clip_hi_xyz[1]), int(clip_hi_xyz[2]), nz, ny, nx,
                                 _ptr(ws), nbytes, _ptr(mask), _ptr(status), _stream()), 'label_aa_mask')
    return mask, status


class _MaskBase:
    """Constructor, coordinate transform, neighbour helper and save_mask shared by the three scripts
    (create_backbone_mask.py:34-118,179-208)."""

    def __init__(self, map_path, device='cuda'):
        ops.require_gpu()
        self.device = torch.device(device)
        self.map_path = map_path
        m = mrc.read_mrc(map_path)
        self._map = m
        self.map_data = m.data
        self.shape = m.data.shape
        self.voxel_size = np.rec.array(tuple(np.float32(v) for v in m.voxel_size), dtype=_REC)
        self.origin = np.rec.array(tuple(np.float32(v) for v in m.origin), dtype=_REC)
        self.mapc, self.mapr, self.maps = m.mapc, m.mapr, m.maps
        self.nxstart, self.nystart, self.nzstart = m.nxstart, m.nystart, m.nzstart

    def transform_coordinates(self, coord):
        """Host helper for one coordinate (the kernels apply the same arithmetic per atom)."""
        shifted = np.asarray(coord, dtype=np.float32) - np.array([self.origin.x, self.origin.y, self.origin.z])
        indices = np.round(shifted / 1.0).astype(int)
        return np.clip(indices, 0, np.array(self.shape) - 1)

    def get_neighbors(self, center):
        x, y, z = center
        out = []
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    if dx == 0 and dy == 0 and dz == 0:
                        continue
                    q = (x + dx, y + dy, z + dz)
                    if all(0 <= q[a] < self.shape[a] for a in range(3)):
                        out.append(q)
        return out

    def save_mask(self, mask, output_path):
        mrc.write_mrc(output_path, mrc.MrcMap(
            data=np.asarray(mask).astype(np.float32), voxel_size=(np.float32(1),) * 3, origin=self._map.origin,
            mapc=self.mapc, mapr=self.mapr, maps=self.maps, nxstart=self.nxstart, nystart=self.nystart,
            nzstart=self.nzstart))

    def _origin(self):
        return (self.origin.x, self.origin.y, self.origin.z)

    @staticmethod
    def _raise_if_oob(status):
        if int(status.item()) != 0:
            raise IndexError('index is out of bounds for the map axis (clip bounds are (nz,ny,nx) against (x,y,z))')


class _ClassMask(_MaskBase):
    CLASS_ATOMS: tuple = ()

    def generate_mask(self, pdb_file):
        """0 background, 1 neighbour voxel, 2 other atoms, 3 class atoms (host int32 volume)."""
        try:
            rec = pdb.read_pdb_records(pdb_file)
            is_cls = np.fromiter((a in self.CLASS_ATOMS for a in rec['atom_names']), dtype=np.uint8,
                                 count=len(rec['atom_names']))
            mask, status = class_mask(torch.from_numpy(rec['coords']).to(self.device),
                                      torch.from_numpy(is_cls).to(self.device), self._origin(), self.shape)
            self._raise_if_oob(status)
            self._mask_dev = mask
            return mask.cpu().numpy()
        except Exception as e:
            print(f'Error generating mask: {str(e)}')
            raise


class BackboneMask(_ClassMask):
    """scripts_for_training_data/create_backbone_mask.py:25-208."""
    CLASS_ATOMS = ('N', 'CA', 'C', 'O')


class CarbonAlphaMask(_ClassMask):
    """scripts_for_training_data/create_carbon_alpha_mask.py:25-209."""
    CLASS_ATOMS = ('CA',)


class AminoAcidMaskGenerator(_MaskBase):
    """scripts_for_training_data/create_amino_acid_mask.py:23-215."""

    def __init__(self, map_path, device='cuda'):
        super().__init__(map_path, device)
        self.aa_mapping = dict(AA_MAPPING)

    def generate_mask(self, pdb_path):
        try:
            rec = pdb.read_pdb_records(pdb_path)
            # first C-alpha of every residue whose name is one of the 20 (:156-163)
            seen, rows, labs = set(), [], []
            for a, (name, resn, r) in enumerate(zip(rec['atom_names'], rec['res_names'], rec['res_index'])):
                if name == 'CA' and resn in self.aa_mapping and r not in seen:
                    seen.add(r)
                    rows.append(a)
                    labs.append(self.aa_mapping[resn])
            coords = torch.from_numpy(np.ascontiguousarray(rec['coords'][rows]).reshape(-1, 3)).to(self.device)
            labels = torch.from_numpy(np.asarray(labs, dtype=np.int32)).to(self.device)
            mask, status = aa_mask(coords, labels, self._origin(), self.shape)
            self._raise_if_oob(status)
            self._mask_dev = mask
            return mask.cpu().numpy()
        except Exception as e:
            print(f'Error generating mask: {str(e)}')
            raise


def build_masks(kind, base_dir='Training_Dataset/Raw_Data', processed_dir='Training_Dataset/Processed_Data',
                device='cuda'):
    """The ``main()`` loops of the three scripts (create_backbone_mask.py:236-266).  ``kind`` is
    'backbone', 'carbon_alpha' or 'amino_acid'.  Returns the number of masks written."""
    cls, fname = {'backbone': (BackboneMask, 'backbone_mask.mrc'),
                  'carbon_alpha': (CarbonAlphaMask, 'carbon_alpha_mask.mrc'),
                  'amino_acid': (AminoAcidMaskGenerator, 'amino_acid_mask.mrc')}[kind]
    done = 0
    for i, directory in enumerate(sorted(glob(f'{base_dir}/*'))):
        emd_id = directory.split('/')[-1]
        pdb_file = next((p for p in glob(f'{directory}/*.pdb') if len(p.split('/')[-1]) == 8), None)   # "1abc.pdb"
        if pdb_file is None:
            print(f'No suitable PDB file found for EMD ID: {emd_id}')
            continue
        out = f'{processed_dir}/{emd_id}'
        os.makedirs(out, exist_ok=True)
        try:
            gen = cls(f'{out}/resampled_normalized_map.mrc', device=device)
            gen.save_mask(gen.generate_mask(pdb_file), f'{out}/{fname}')
            done += 1
        except Exception as e:
            print(f'Failed for density map with EMD ID: {emd_id} - Error: {str(e)}')
    return done
