"""Stand-in namespace for Biopython (absent from this image); see Bio/PDB."""
