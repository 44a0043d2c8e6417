"""Process-level registry of device-resident pipeline products, keyed by the file or
directory path under which the reference would have stored them.

The reference hands data from stage to stage through the file system
(resampled_normalized_map.mrc -> AF3_encodings/*.mrc -> grids/*/*.npz ->
results/predicted_grids/*.npz, utils/modeler.py:673-760).  The drop-in classes keep
those paths as *names* and pass the tensors through this registry instead, so a
caller written against the reference (Solver.getData / Solver.nnPred) works unchanged
while nothing but the final volumes ever leaves HBM.

Lifetime: an entry pins its tensors in HBM until it is dropped.  ``CryoEMPredictor.run_prediction``
drops the entries it consumed (normalised map, AF3 atoms, both cube directories) when it is done --
the point at which ``Solver.nnPred`` deletes the corresponding files (utils/modeler.py:753-758) --
so a process that runs many maps does not accumulate them; ``drop`` / ``release_under`` / ``clear``
release by hand.  An entry also stands in for the file of the same name, therefore ``get`` checks the
file: if somebody (re)wrote it after the entry was registered, the entry is stale and is discarded."""
from __future__ import annotations

import os
import time

_REGISTRY: dict[str, dict] = {}
_META: dict[str, tuple] = {}          # key -> (time of registration, mtime of the file then or None)


def _key(path) -> str:
    return os.path.abspath(str(path)).rstrip('/')


def _mtime(key):
    try:
        return os.stat(key).st_mtime
    except OSError:
        return None


def put(path, **entry):
    key = _key(path)
    _META[key] = (time.time(), _mtime(key))
    _REGISTRY[key] = entry
    return entry


def get(path):
    key = _key(path)
    entry = _REGISTRY.get(key)
    if entry is None:
        return None
    now = _mtime(key)
    if now is not None and os.path.isfile(key):
        registered, seen = _META.get(key, (0.0, None))
        # a file that appeared or changed after registration (1 s slack for coarse clocks and for the
        # drop-in's own write_files=True output, written just before the entry is registered)
        if (seen is None and now > registered + 1.0) or (seen is not None and now > seen + 1e-6):
            drop(key)
            return None
    return entry


def drop(path):
    _REGISTRY.pop(_key(path), None)
    _META.pop(_key(path), None)


def release_under(path):
    """Drop every entry at or below ``path`` (e.g. a map's working directory)."""
    root = _key(path)
    for k in [k for k in _REGISTRY if k == root or k.startswith(root + os.sep)]:
        drop(k)


def clear():
    _REGISTRY.clear()
    _META.clear()
