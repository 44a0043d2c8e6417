"""Process-level registry of device-resident pipeline products, keyed by the file or
directory path under which the reference would have stored them.

The reference hands data from stage to stage through the file system
(resampled_normalized_map.mrc -> AF3_encodings/*.mrc -> grids/*/*.npz ->
results/predicted_grids/*.npz, utils/modeler.py:673-760).  The drop-in classes keep
those paths as *names* and pass the tensors through this registry instead, so a
caller written against the reference (Solver.getData / Solver.nnPred) works unchanged
while nothing but the final volumes ever leaves HBM."""
from __future__ import annotations

import os

_REGISTRY: dict[str, dict] = {}


def _key(path) -> str:
    return os.path.abspath(str(path)).rstrip('/')


def put(path, **entry):
    _REGISTRY[_key(path)] = entry
    return entry


def get(path):
    return _REGISTRY.get(_key(path))


def drop(path):
    _REGISTRY.pop(_key(path), None)


def clear():
    _REGISTRY.clear()
