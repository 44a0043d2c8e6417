"""CPU oracle for SURVEY.md section 8(f) rows N3 (training label masks) and N4 (docking masks).

TEST INFRASTRUCTURE ONLY (same rule as oracle/mica_oracle.py).

N3: scripts_for_training_data/create_backbone_mask.py:120-177, create_carbon_alpha_mask.py:120-177,
    create_amino_acid_mask.py:128-183 -- restated on atom arrays (the PDB text is already parsed).
N4: utils/dock_in_map.py:248-283 (contour threshold) and :285-364 (zero the map within ``radius`` of
    the ``percentage`` % of atoms closest to the structure's centroid).

Pinning: oracle/make_golden.py runs the unmodified reference classes through
oracle/ref_harness.py::{label_masks, docking_masks}, asserts these functions reproduce them bit for
bit, and commits the reference's outputs as tests/golden/label_masks.npz and docking_masks.npz.
"""
from __future__ import annotations

import numpy as np

AA_LABELS = {'ALA': 1, 'CYS': 2, 'ASP': 3, 'GLU': 4, 'PHE': 5, 'GLY': 6, 'HIS': 7, 'ILE': 8, 'LYS': 9,
             'LEU': 10, 'MET': 11, 'ASN': 12, 'PRO': 13, 'GLN': 14, 'ARG': 15, 'SER': 16, 'THR': 17,
             'VAL': 18, 'TRP': 19, 'TYR': 20}           # create_amino_acid_mask.py:40-45


def atom_positions(coords, origin_xyz, shape):
    """``transform_coordinates`` + ``pos = (idx[2], idx[1], idx[0])`` (create_backbone_mask.py:65-88,
    :150-151): float32 subtract, round half to even, clip (x,y,z) against (shape[0],shape[1],shape[2])
    = (nz,ny,nx) -- the mis-ordered clip of SURVEY D7.  Returns int64 [A,3] as (z,y,x)."""
    shifted = np.asarray(coords, dtype=np.float32) - np.array([np.float32(v) for v in origin_xyz])
    idx = np.round(shifted / 1.0).astype(int)
    idx = np.clip(idx, 0, np.array(shape) - 1)
    return idx[:, ::-1]


def _neighbors(pos, shape):
    z, y, x = pos
    out = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dz == 0 and dy == 0 and dx == 0:
                    continue
                q = (z + dz, y + dy, x + dx)
                if 0 <= q[0] < shape[0] and 0 <= q[1] < shape[1] and 0 <= q[2] < shape[2]:
                    out.append(q)
    return out


def atom_class_mask(pos_zyx, is_class, shape):
    """create_backbone_mask.py:136-172 / create_carbon_alpha_mask.py:136-173.  Atoms in file order:
    3 where ``is_class`` (backbone atom / C-alpha), else 2 -- the LAST atom on a voxel wins; then
    every in-bounds 26-neighbour of an atom voxel that holds no atom becomes 1.  An index beyond its
    real axis raises IndexError exactly as ``mask[pos] = ...`` does."""
    mask = np.zeros(shape, dtype=np.int32)
    assigned = {}
    for p, c in zip(pos_zyx, is_class):
        p = (int(p[0]), int(p[1]), int(p[2]))
        mask[p] = 3 if c else 2
        assigned[p] = mask[p]
    for p in list(assigned.keys()):
        for q in _neighbors(p, shape):
            if q not in assigned:
                mask[q] = 1
                assigned[q] = 1
    return mask


def amino_acid_mask(ca_pos_zyx, aa_label, shape):
    """create_amino_acid_mask.py:151-177.  C-alphas in file order; each gives its label to the
    in-bounds 26-neighbours that are unassigned or hold a LARGER label, then zeroes its own voxel
    (without forgetting the label that voxel was assigned)."""
    mask = np.zeros(shape, dtype=np.int32)
    assigned = {}
    for p, a in zip(ca_pos_zyx, aa_label):
        p = (int(p[0]), int(p[1]), int(p[2]))
        for q in _neighbors(p, shape):
            if q not in assigned or a < assigned[q]:
                mask[q] = a
                assigned[q] = a
        mask[p] = 0
    return mask


def amino_acid_mask_closed_form(ca_pos_zyx, aa_label, shape):
    """The order-free statement the CUDA kernels implement (checked equal to ``amino_acid_mask``):
    for a voxel v let a_all = min label over the C-alphas that have v as a neighbour, T = the index
    of the LAST C-alpha sitting on v (none: v keeps a_all), a_0 = min label over the neighbouring
    C-alphas that come BEFORE T.  v ends as a_all if a later C-alpha lowered the minimum
    (a_all < a_0), else 0."""
    INF = 1 << 30
    a_all = np.full(shape, INF, dtype=np.int64)
    a_0 = np.full(shape, INF, dtype=np.int64)
    last = np.zeros(shape, dtype=np.int64)
    pos = [(int(p[0]), int(p[1]), int(p[2])) for p in ca_pos_zyx]
    for t, (p, a) in enumerate(zip(pos, aa_label)):
        last[p] = max(last[p], t + 1)
        for q in _neighbors(p, shape):
            a_all[q] = min(a_all[q], a)
    for t, (p, a) in enumerate(zip(pos, aa_label)):
        for q in _neighbors(p, shape):
            if last[q] and t + 1 < last[q]:
                a_0[q] = min(a_0[q], a)
    out = np.where(a_all == INF, 0, np.where(last == 0, a_all, np.where(a_all < a_0, a_all, 0)))
    return out.astype(np.int32)


# ------------------------------------------------------------------------------------ N4
def contour_threshold(data, contour_level):
    """utils/dock_in_map.py:269: ``np.where(data < contour_level, 0, data)`` (float32 compare)."""
    return np.where(data < contour_level, 0, data).astype(np.float32)


def select_central_atoms(coords, percentage=40, centroid_method='median'):
    """utils/dock_in_map.py:314-327: the ``percentage`` % of atoms closest to the centroid."""
    coords = np.asarray(coords)
    centroid = np.mean(coords, axis=0) if centroid_method == 'mean' else np.median(coords, axis=0)
    d = np.sqrt(np.sum((coords - centroid) ** 2, axis=1))
    n_use = int(len(coords) * (percentage / 100.0))
    return coords[np.argsort(d)[:n_use]]


def mask_around_atoms(map_data, selected_coords, voxel_size_xyz, origin_xyz, radius=2.0):
    """utils/dock_in_map.py:330-352: seeds = truncated voxel coordinates of the atoms that pass the
    (mis-ordered, D7-like) bounds test; every voxel whose Euclidean distance transform to the seeds
    (``sampling=voxel_size``: [vx,vy,vz] applied to axes (z,y,x)) is <= radius is zeroed."""
    from scipy.ndimage import distance_transform_edt
    voxel = np.array([np.float32(v) for v in voxel_size_xyz])
    origin = np.array([np.float32(v) for v in origin_xyz])
    vc = ((np.asarray(selected_coords) - origin) / voxel).astype(int)
    mask = np.zeros_like(map_data, dtype=bool)
    ok = ((vc >= 0) & (vc < np.array(map_data.shape))).all(axis=1)
    vc = vc[ok]
    mask[vc[:, 2], vc[:, 1], vc[:, 0]] = True
    dist = distance_transform_edt(~mask, sampling=voxel)
    out = map_data.copy()
    out[dist <= radius] = 0
    return out.astype(np.float32)


def mask_around_atoms_restated(map_data, selected_coords, voxel_size_xyz, origin_xyz, radius=2.0):
    """The same without the distance transform (what the CUDA kernel does): a voxel is zeroed iff
    some seed lies within ``radius``, the distance computed as SciPy does -- float64,
    sqrt(((dz*s0)^2 + (dy*s1)^2) + (dx*s2)^2) with s = float64([vx,vy,vz])."""
    voxel = np.array([np.float32(v) for v in voxel_size_xyz])
    origin = np.array([np.float32(v) for v in origin_xyz])
    vc = ((np.asarray(selected_coords) - origin) / voxel).astype(int)
    nz, ny, nx = map_data.shape
    ok = ((vc >= 0) & (vc < np.array(map_data.shape))).all(axis=1)
    vc = vc[ok]
    if len(vc) and (vc[:, 0].max() >= nx or vc[:, 2].max() >= nz or vc[:, 1].max() >= ny):
        raise IndexError('seed index out of bounds for its real axis')
    s = voxel.astype(np.float64)
    out = map_data.copy()
    if radius < 0:
        return out.astype(np.float32)
    rz, ry, rx = (int(np.floor(radius / s[a])) + 1 for a in range(3))
    for x, y, z in np.unique(vc, axis=0):
        z0, z1 = max(0, z - rz), min(nz, z + rz + 1)
        y0, y1 = max(0, y - ry), min(ny, y + ry + 1)
        x0, x1 = max(0, x - rx), min(nx, x + rx + 1)
        dz = (np.arange(z0, z1) - z)[:, None, None] * s[0]
        dy = (np.arange(y0, y1) - y)[None, :, None] * s[1]
        dx = (np.arange(x0, x1) - x)[None, None, :] * s[2]
        d = np.sqrt((dz * dz + dy * dy) + dx * dx)
        out[z0:z1, y0:y1, x0:x1][d <= radius] = 0
    return out.astype(np.float32)
