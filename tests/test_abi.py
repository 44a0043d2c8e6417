"""The C-ABI library loads, exports every symbol include/mica_b200.h declares, and
fails loudly (no CPU fallback) when no GPU is present.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'mica_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mica_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from mica_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(_lib.lib, s), f'{s} declared in include/mica_b200.h but not exported'
    assert sorted(_lib.SIGNATURES) == syms, 'ctypes signatures out of sync with the header'
    assert _lib.lib.mica_version() >= 100


def test_sm100a_only_cubin():
    import shutil
    import subprocess
    from mica_b200 import _lib
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    out = subprocess.run([cuobjdump, '-lelf', _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_\d+a?', out))
    assert archs == {'sm_100a'}, archs


def test_host_only_entry_points():
    from mica_b200 import ops
    assert ops.zoom_output_shape((200, 200, 200), (1.06, 1.06, 1.06)) == (212, 212, 212)
    assert ops.zoom_output_shape((400, 400, 400), (1.2, 1.2, 1.2)) == (480, 480, 480)
    # banker's rounding of the float32 product: 10 * 1.25 = 12.5 -> 12; 6 * 1.25 = 7.5 -> 8
    assert ops.zoom_output_shape((10, 6, 2), (1.25, 1.25, 1.25)) == (12, 8, 2)
    ijk = ops.cube_origins((70, 100, 50), 48)
    assert ijk.shape == (12, 3) and ijk.dtype == np.int32 and list(ijk[1]) == [0, 0, 48]


def test_no_cpu_fallback():
    import torch
    from mica_b200 import _lib, ops
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    assert _lib.lib.mica_device_count() == -4
    with pytest.raises(_lib.MicaError):
        ops.require_gpu()
    with pytest.raises(_lib.MicaError):
        ops.normalize(torch.zeros(8))
    with pytest.raises(_lib.MicaError):
        ops.extract_cubes(torch.zeros(4, 4, 4), torch.zeros((1, 3), dtype=torch.int32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'mica_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', text, flags=re.M), f
                assert 'scipy.ndimage import zoom' not in text, f
