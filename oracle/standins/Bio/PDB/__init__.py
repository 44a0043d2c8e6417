"""Minimal stand-in for ``Bio.PDB`` (see _impl.py), used ONLY by oracle/ref_harness.py.  Laid out like
Biopython: the parser class lives in the submodule ``Bio.PDB.PDBParser`` (utils/modeler.py:19 imports it
by that path) and is re-exported here (``from Bio.PDB import *`` in the training scripts)."""
from ._impl import PDBIO, _Atom, _Residue, _Structure  # noqa: F401
from .PDBParser import PDBParser  # noqa: F401  (rebinds the name from the submodule to the class)

__all__ = ['PDBParser', 'PDBIO']
