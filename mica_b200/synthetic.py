"""Seeded synthetic inputs for tests and bench.py.

Not reference behaviour: the reference ships no data.  Shapes follow
BASELINE.json ``configs``; everything derives from ``np.random.default_rng(seed)``
(seed 2022, the reference's default ``--seed``, run.py:86).
"""
from __future__ import annotations

import numpy as np

AMINO_ACIDS = ['ALA', 'CYS', 'ASP', 'GLU', 'PHE', 'GLY', 'HIS', 'ILE', 'LYS', 'LEU',
               'MET', 'ASN', 'PRO', 'GLN', 'ARG', 'SER', 'THR', 'VAL', 'TRP', 'TYR']

#: heavy-atom names per residue (backbone first)
HEAVY_ATOMS = {
    'ALA': ['N', 'CA', 'C', 'O', 'CB'],
    'CYS': ['N', 'CA', 'C', 'O', 'CB', 'SG'],
    'ASP': ['N', 'CA', 'C', 'O', 'CB', 'CG', 'OD1', 'OD2'],
    'GLU': ['N', 'CA', 'C', 'O', 'CB', 'CG', 'CD', 'OE1', 'OE2'],
    'PHE': ['N', 'CA', 'C', 'O', 'CB', 'CG', 'CD1', 'CD2', 'CE1', 'CE2', 'CZ'],
    'GLY': ['N', 'CA', 'C', 'O'],
    'HIS': ['N', 'CA', 'C', 'O', 'CB', 'CG', 'ND1', 'CD2', 'CE1', 'NE2'],
    'ILE': ['N', 'CA', 'C', 'O', 'CB', 'CG1', 'CG2', 'CD1'],
    'LYS': ['N', 'CA', 'C', 'O', 'CB', 'CG', 'CD', 'CE', 'NZ'],
    'LEU': ['N', 'CA', 'C', 'O', 'CB', 'CG', 'CD1', 'CD2'],
    'MET': ['N', 'CA', 'C', 'O', 'CB', 'CG', 'SD', 'CE'],
    'ASN': ['N', 'CA', 'C', 'O', 'CB', 'CG', 'OD1', 'ND2'],
    'PRO': ['N', 'CA', 'C', 'O', 'CB', 'CG', 'CD'],
    'GLN': ['N', 'CA', 'C', 'O', 'CB', 'CG', 'CD', 'OE1', 'NE2'],
    'ARG': ['N', 'CA', 'C', 'O', 'CB', 'CG', 'CD', 'NE', 'CZ', 'NH1', 'NH2'],
    'SER': ['N', 'CA', 'C', 'O', 'CB', 'OG'],
    'THR': ['N', 'CA', 'C', 'O', 'CB', 'OG1', 'CG2'],
    'VAL': ['N', 'CA', 'C', 'O', 'CB', 'CG1', 'CG2'],
    'TRP': ['N', 'CA', 'C', 'O', 'CB', 'CG', 'CD1', 'CD2', 'NE1', 'CE2', 'CE3', 'CZ2', 'CZ3', 'CH2'],
    'TYR': ['N', 'CA', 'C', 'O', 'CB', 'CG', 'CD1', 'CD2', 'CE1', 'CE2', 'CZ', 'OH'],
}


def synthetic_structure(n_residues, box_xyz, seed=2022, origin_xyz=(0.0, 0.0, 0.0),
                        margin=4.0, hetero_every=0, unknown_every=0):
    """Random-walk C-alpha chain (3.8 A steps, reflected at the box walls) with a
    full heavy-atom set per residue scattered within ~4 A of its C-alpha (atoms are kept
    inside the box: on a non-cubic grid the reference's clip quirk D7 turns an atom beyond
    the x range into an IndexError).

    Returns dict(coords float32 [A,3] in (x,y,z) Angstrom, atom_names [A],
    res_names [A], res_ids int [A], hetero bool [A]).  Coordinates are rounded to
    3 decimals so that a PDB text round trip is exact.  ``hetero_every`` /
    ``unknown_every`` sprinkle HETATM residues (skipped by the encoder,
    utils/preprocessing.py:279) and non-standard residue names (no amino-acid
    channel, :180-185) for edge-case tests."""
    rng = np.random.default_rng(seed)
    box = np.asarray(box_xyz, dtype=np.float64)
    lo, hi = margin, box - margin
    steps = rng.normal(size=(n_residues, 3))
    steps *= 3.8 / np.linalg.norm(steps, axis=1, keepdims=True)
    ca = np.empty((n_residues, 3))
    pos = lo + rng.random(3) * (hi - lo)
    for r in range(n_residues):
        pos = pos + steps[r]
        for a in range(3):                      # reflect at the walls
            if pos[a] < lo:
                pos[a] = 2 * lo - pos[a]
            elif pos[a] > hi[a]:
                pos[a] = 2 * hi[a] - pos[a]
        ca[r] = pos
    types = rng.integers(0, 20, size=n_residues)
    coords, atom_names, res_names, res_ids, hetero = [], [], [], [], []
    for r in range(n_residues):
        name = AMINO_ACIDS[types[r]]
        names = HEAVY_ATOMS[name]
        off = rng.normal(scale=1.6, size=(len(names), 3))
        off[names.index('CA')] = 0.0
        xyz = ca[r] + off
        is_het = bool(hetero_every) and (r % hetero_every == hetero_every - 1)
        if unknown_every and (r % unknown_every == unknown_every - 1):
            name = 'UNK'
        coords.append(xyz)
        atom_names += names
        res_names += [name] * len(names)
        res_ids += [r + 1] * len(names)
        hetero += [is_het] * len(names)
    coords = np.clip(np.concatenate(coords), 0.0, box - 1.0) + np.asarray(origin_xyz, dtype=np.float64)
    coords = np.round(coords, 3).astype(np.float32)
    return dict(coords=coords, atom_names=atom_names, res_names=res_names,
                res_ids=np.asarray(res_ids), hetero=np.asarray(hetero, dtype=bool))


def write_pdb(path, structure):
    """Fixed-column PDB text (one chain per 9999 residues)."""
    with open(path, 'w') as f:
        for n, (xyz, an, rn, rid, het) in enumerate(zip(
                structure['coords'], structure['atom_names'], structure['res_names'],
                structure['res_ids'], structure['hetero'])):
            chain = chr(ord('A') + ((int(rid) - 1) // 9999) % 26)
            rec = 'HETATM' if het else 'ATOM  '
            aname = an if len(an) == 4 else ' ' + an
            f.write('%s%5d %-4s %3s %s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f          %2s\n' % (
                rec, (n + 1) % 100000, aname, rn, chain, (int(rid) - 1) % 9999 + 1,
                xyz[0], xyz[1], xyz[2], 1.0, 0.0, an[0]))
        f.write('END\n')


def synthetic_map(shape_zyx, voxel=1.06, resolution=3.7, seed=2022, n_atoms=None,
                  noise=0.05):
    """Density-like float32 volume (z,y,x): unit point masses at random positions
    blurred by a Gaussian of sigma = 0.225 * resolution (in Angstrom) plus
    N(0, noise) -- positive tail, smooth, mostly background (section 8d)."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    shape = tuple(int(s) for s in shape_zyx)
    nvox = int(np.prod(shape))
    if n_atoms is None:
        n_atoms = max(16, nvox // 400)
    vol = np.zeros(shape, dtype=np.float32)
    centre = np.array(shape) / 2.0
    # atoms concentrated in a central blob (a "particle" in solvent)
    pts = rng.normal(loc=centre, scale=np.array(shape) / 6.0, size=(n_atoms, 3))
    pts = np.clip(np.round(pts).astype(np.int64), 0, np.array(shape) - 1)
    np.add.at(vol, (pts[:, 0], pts[:, 1], pts[:, 2]), 1.0)
    sigma = 0.225 * resolution / voxel
    vol = gaussian_filter(vol, sigma=sigma, mode='constant')
    vol /= max(float(vol.max()), 1e-12)
    vol += rng.normal(scale=noise, size=shape).astype(np.float32)
    return vol.astype(np.float32)


def synthetic_logits(n_cubes, window=64, seed=2022):
    """Stand-in model outputs [n,4,W^3], [n,4,W^3], [n,21,W^3] (float32)."""
    rng = np.random.default_rng(seed)
    vox = (window, window, window)

    def draw(c):
        a = rng.standard_normal((n_cubes, c) + vox, dtype=np.float32)
        a *= np.float32(2.0)
        return a

    return draw(4), draw(4), draw(21)


def synthetic_predictions(shape_xyz, n_residues=(60, 25), seed=2022, spurious=2, wall_margin=3.0):
    """Stitched-volume look-alikes for the candidate step (utils/modeler.py:762-858): one
    random-walk C-alpha chain per entry of ``n_residues`` (each confined to its own part of the
    box so that DBSCAN sees separate clusters) rendered as Gaussian peaks into
    ``carbon_alpha_probability``, a wider tube for ``backbone_probability``, smooth random
    ``amino_acid_probability`` [20,X,Y,Z] and its arg-max as float32 (utils/predict.py:462).
    ``spurious`` adds bright C-alpha blobs with no backbone support (clusters the reference's
    score filter must drop).  A little noise breaks ties between voxels.  ``wall_margin=0`` lets the
    chains touch the faces of the box (picks on the border, which the reference skips)."""
    rng = np.random.default_rng(seed)
    X, Y, Z = (int(v) for v in shape_xyz)
    gx, gy, gz = np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing='ij')
    ca = np.zeros((X, Y, Z), np.float64)
    bb = np.zeros((X, Y, Z), np.float64)

    def splat(vol, p, sigma, amp):
        lo = np.maximum(np.floor(p - 4 * sigma).astype(int), 0)
        hi = np.minimum(np.ceil(p + 4 * sigma).astype(int) + 1, [X, Y, Z])
        sl = tuple(slice(a, b) for a, b in zip(lo, hi))
        d2 = (gx[sl] - p[0]) ** 2 + (gy[sl] - p[1]) ** 2 + (gz[sl] - p[2]) ** 2
        np.maximum(vol[sl], amp * np.exp(-d2 / (2 * sigma * sigma)), out=vol[sl])

    n_chains = len(n_residues)
    for c, n in enumerate(n_residues):
        x_lo, x_hi = 3 + c * (X - 6) / n_chains, 3 + (c + 1) * (X - 6) / n_chains - 12
        lo = np.array([x_lo, wall_margin, wall_margin])
        hi = np.array([max(x_hi, x_lo + 6), Y - 1.0 - wall_margin, Z - 1.0 - wall_margin])
        pos = lo + rng.random(3) * (hi - lo)
        for _ in range(n):
            step = rng.normal(size=3)
            pos = pos + 3.8 * step / np.linalg.norm(step)
            pos = np.where(pos < lo, 2 * lo - pos, pos)
            pos = np.where(pos > hi, 2 * hi - pos, pos)
            splat(ca, pos, 0.9, 0.55 + 0.44 * rng.random())
            splat(bb, pos, 1.8, 0.95)
    for _ in range(spurious):
        p = np.array([X - 5.0, 4.0 + rng.random() * (Y - 8), 4.0 + rng.random() * (Z - 8)])
        for _ in range(3):
            splat(ca, p + rng.normal(scale=1.5, size=3), 1.2, 0.9)
    ca = np.clip(ca + rng.normal(scale=0.01, size=ca.shape), 0.0, 1.0).astype(np.float32)
    bb = np.clip(bb + rng.normal(scale=0.01, size=bb.shape), 0.0, 1.0).astype(np.float32)
    aa = rng.random((20, X, Y, Z), dtype=np.float32) ** 4
    aa /= aa.sum(axis=0, keepdims=True)
    aa_pred = aa.argmax(axis=0).astype(np.float32)
    return dict(carbon_alpha_probability=ca, backbone_probability=bb,
                amino_acid_probability=aa.astype(np.float32), amino_acid_prediction=aa_pred)


def synthetic_map_device(shape_zyx, device, voxel=1.06, resolution=3.7, seed=2022, noise=0.05):
    """``synthetic_map`` built on the GPU with torch ops (bench inputs of 512^3 .. 720^3: the NumPy
    version takes half a minute there).  Same recipe -- point masses in a central blob, separable
    Gaussian blur of sigma = 0.225 * resolution / voxel, N(0, noise) -- but a different random stream,
    so it is NOT voxel-identical to ``synthetic_map``; every rank of a multi-GPU run that uses the
    same seed gets the same map.  Input synthesis only: no part of the measured path."""
    import torch
    import torch.nn.functional as F
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(int(seed))
    shape = tuple(int(s) for s in shape_zyx)
    nvox = shape[0] * shape[1] * shape[2]
    n_atoms = max(16, nvox // 400)
    dims = torch.tensor(shape, dtype=torch.float32, device=dev)
    pts = torch.randn((n_atoms, 3), generator=gen, device=dev) * (dims / 6.0) + dims / 2.0
    pts = torch.minimum(torch.clamp(torch.round(pts), min=0), dims - 1).long()
    lin = (pts[:, 0] * shape[1] + pts[:, 1]) * shape[2] + pts[:, 2]
    vol = torch.zeros(nvox, dtype=torch.float32, device=dev)
    vol.index_add_(0, lin, torch.ones(n_atoms, dtype=torch.float32, device=dev))
    vol = vol.view(1, 1, *shape)
    sigma = 0.225 * resolution / voxel
    r = max(1, int(4.0 * sigma + 0.5))
    t = torch.arange(-r, r + 1, dtype=torch.float32, device=dev)
    k = torch.exp(-0.5 * (t / sigma) ** 2)
    k = k / k.sum()
    for axis in range(3):
        ks = [1, 1, 1]
        ks[axis] = 2 * r + 1
        pad = [0, 0, 0]
        pad[axis] = r
        vol = F.conv3d(vol, k.view(1, 1, *ks), padding=tuple(pad))
    vol = vol.view(shape)
    vol /= vol.max().clamp_min(1e-12)
    vol += torch.randn(shape, generator=gen, device=dev) * noise
    return vol.contiguous()


def pointwise_model(x, af):
    """Deterministic elementwise stand-in for ``MICA.forward(exp_map, af_features) -> (bb, ca, aa)``:
    the logits of a voxel depend only on that voxel's inputs, so two runs that cut the same cube get the
    same logits whatever the batching, the rank or the call order (parity legs of bench.py / tests)."""
    import math
    import torch
    s = af.sum(dim=1, keepdim=True)
    bb = torch.cat([x, -x, 2 * x - 0.5 + s, x * x], dim=1)
    ca = torch.cat([0.5 - x, x, x * 3 - 1, 1.5 * x + af[:, :1]], dim=1)
    aa = torch.cat([x * (0.1 * t) + af[:, t % 24:t % 24 + 1] * (t % 3) + math.sin(t) for t in range(21)], dim=1)
    return bb, ca, aa
