"""Minimal stand-in for ``Bio.PDB`` used ONLY by oracle/ref_harness.py to run the
unmodified reference (utils/preprocessing.py:52-53,269,275-298).  Fixed-column
ATOM/HETATM parser: coordinates float32 from columns 31-38/39-46/47-54, atom
name = columns 13-16 stripped, residue name = columns 18-20, hetero flag ' '
for ATOM records."""
import numpy as np


class _Atom:
    def __init__(self, name, coord):
        self._name, self._coord = name, coord

    def get_coord(self):
        return self._coord

    def get_name(self):
        return self._name


class _Residue(list):
    def __init__(self, rid, resname):
        super().__init__()
        self._id, self._resname = rid, resname
        self.id = rid

    def __contains__(self, key):                 # ``'CA' in residue`` (create_amino_acid_mask.py:156)
        if isinstance(key, str):
            return any(a.get_name() == key for a in self)
        return list.__contains__(self, key)

    def __getitem__(self, key):                  # ``residue['CA']``: first atom of that name
        if isinstance(key, str):
            for a in self:
                if a.get_name() == key:
                    return a
            raise KeyError(key)
        return list.__getitem__(self, key)

    def get_id(self):
        return self._id

    def get_resname(self):
        return self._resname


class PDBParser:
    def __init__(self, QUIET=False, **kw):
        pass

    def get_structure(self, name, path):
        models, chains, cur_res_key = [], None, None
        with open(path) as f:
            for line in f:
                rec = line[:6]
                if rec.startswith('MODEL') or chains is None:
                    chains = {}
                    models.append(chains)
                    cur_res_key = None
                    if rec.startswith('MODEL'):
                        continue
                if rec not in ('ATOM  ', 'HETATM'):
                    continue
                chain_id = line[21]
                resname = line[17:20].strip()
                resseq, icode = int(line[22:26]), line[26]
                het = ' ' if rec == 'ATOM  ' else ('W' if resname in ('HOH', 'WAT') else 'H_' + resname)
                key = (chain_id, het, resseq, icode)
                chain = chains.setdefault(chain_id, [])
                if key != cur_res_key:
                    chain.append(_Residue((het, resseq, icode), resname))
                    cur_res_key = key
                coord = np.array([float(line[30:38]), float(line[38:46]), float(line[46:54])], 'f')
                chain[-1].append(_Atom(line[12:16].strip(), coord))
        return _Structure(list(m.values()) for m in models)


class _Structure(list):
    """models -> chains -> residues -> atoms, iterable like Bio.PDB's Structure."""

    def get_atoms(self):                         # utils/dock_in_map.py:312
        for model in self:
            for chain in model:
                for residue in chain:
                    yield from residue


class PDBIO:
    pass
