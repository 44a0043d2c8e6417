"""Stand-in submodule ``Bio.PDB.PDBParser`` (utils/modeler.py:19)."""
from ._impl import PDBParser  # noqa: F401
