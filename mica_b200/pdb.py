"""Fixed-column PDB reader producing the atom arrays the AF3 rasteriser consumes.

Stands where ``Bio.PDB.PDBParser.get_structure`` stands in the reference
(utils/preprocessing.py:269,275-298): keeps atoms of residues whose hetero flag
is ' ' (ATOM records, :279), coordinates as float32 (Bio.PDB stores float32),
atom name = columns 13-16 stripped, residue name = columns 18-20.  Text munging
only; the channel mapping mirrors :254-263,180-185,292-298."""
from __future__ import annotations

import numpy as np

BACKBONE_ATOMS = ['CA', 'N', 'C', 'O']
AMINO_ACIDS = ['ALA', 'CYS', 'ASP', 'GLU', 'PHE', 'GLY', 'HIS', 'ILE', 'LYS', 'LEU',
               'MET', 'ASN', 'PRO', 'GLN', 'ARG', 'SER', 'THR', 'VAL', 'TRP', 'TYR']
CHANNEL_NAMES = BACKBONE_ATOMS + AMINO_ACIDS
_BB = {n: i for i, n in enumerate(BACKBONE_ATOMS)}
_AA = {n: 4 + i for i, n in enumerate(AMINO_ACIDS)}


def channel_codes(atom_names, res_names):
    """int8 backbone channel (0..3 | -1) and amino-acid channel (4..23 | -1) per atom."""
    bb = np.fromiter((_BB.get(a, -1) for a in atom_names), dtype=np.int8, count=len(atom_names))
    aa = np.fromiter((_AA.get(r, -1) for r in res_names), dtype=np.int8, count=len(res_names))
    return bb, aa


def read_pdb_atoms(path):
    """Returns (coords float32 [A,3] x,y,z; bb_ch int8 [A]; aa_ch int8 [A];
    n_residues) for the ATOM records of ``path``."""
    xs, names, resn = [], [], []
    n_res, last = 0, None
    with open(path) as f:
        for line in f:
            if not line.startswith('ATOM  '):
                continue
            key = (line[21], line[22:27])
            if key != last:
                n_res += 1
                last = key
            xs.append((line[30:38], line[38:46], line[46:54]))
            names.append(line[12:16].strip())
            resn.append(line[17:20].strip())
    coords = np.array(xs, dtype=np.float64).astype(np.float32).reshape(-1, 3)
    bb, aa = channel_codes(names, resn)
    return coords, bb, aa, n_res


def read_pdb_records(path):
    """Every ATOM and HETATM record in file order -- what iterating a Bio.PDB structure
    ``for model / chain / residue / atom`` visits in the label-mask builders
    (scripts_for_training_data/create_backbone_mask.py:143-147; no hetero filter there).
    Returns dict(coords float32 [A,3], atom_names [A], res_names [A], res_index int64 [A]);
    ``res_index`` numbers the residues (a new one starts whenever chain / hetero flag /
    sequence number / insertion code changes)."""
    xs, names, resn, ridx = [], [], [], []
    n_res, last = -1, None
    with open(path) as f:
        for line in f:
            rec = line[:6]
            if rec.startswith('MODEL'):
                last = None
                continue
            if rec not in ('ATOM  ', 'HETATM'):
                continue
            key = (line[21], rec, line[22:27])
            if key != last:
                n_res += 1
                last = key
            xs.append((line[30:38], line[38:46], line[46:54]))
            names.append(line[12:16].strip())
            resn.append(line[17:20].strip())
            ridx.append(n_res)
    coords = np.array(xs, dtype=np.float64).astype(np.float32).reshape(-1, 3)
    return dict(coords=coords, atom_names=names, res_names=resn, res_index=np.asarray(ridx, dtype=np.int64))
