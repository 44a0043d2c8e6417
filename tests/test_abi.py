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


def _prototypes():
    src = open(os.path.join(ROOT, 'include', 'mica_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'\b(mica_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', src):
        args = m.group(2).strip()
        protos[m.group(1)] = [] if args in ('', 'void') else [a.strip() for a in args.split(',')]
    return protos


def test_ctypes_signatures_match_the_header_prototypes():
    """Argument count and the pointer / integer / float class of every parameter: a drifted binding would
    otherwise only show up as garbage on the GPU box."""
    from mica_b200 import _lib
    protos = _prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES)
    for name, params in protos.items():
        argtypes = _lib.SIGNATURES[name][1]
        assert len(params) == len(argtypes), f'{name}: header has {len(params)} parameters, ctypes {len(argtypes)}'
        for decl, ct in zip(params, argtypes):
            is_ptr = '*' in decl or '[' in decl or 'mica_stream_t' in decl
            ct_ptr = ct is ctypes.c_void_p or ct is ctypes.c_char_p or hasattr(ct, 'contents') or \
                (isinstance(ct, type) and issubclass(ct, ctypes._Pointer))
            assert is_ptr == ct_ptr, f'{name}: `{decl}` bound as {ct}'
            if not is_ptr:
                base = decl.replace('const', '').split()[0]
                want = {'int': ctypes.c_int, 'int64_t': ctypes.c_int64, 'float': ctypes.c_float,
                        'double': ctypes.c_double, 'size_t': ctypes.c_size_t}[base]
                assert ct is want, f'{name}: `{decl}` bound as {ct}'


def test_new_entry_points_reject_bad_arguments_before_touching_the_device():
    from mica_b200 import _lib
    lib = _lib.lib
    assert lib.mica_cand_threshold_count(None, 64, 0.3, None, 0, None, None) == -1
    assert b'null' in lib.mica_last_error()
    assert lib.mica_dbscan_lattice(None, 5, 4, 4, 4, 100, 10, None, 0, None, None, None) == -1
    assert lib.mica_cand_neighbor_graph(None, 70000, None, 4, 4, 4, None, None, None) == -1
    assert lib.mica_label_class_mask(None, None, 1 << 30, 0, 0, 0, 3, 3, 3, 4, 4, 4, None, None, None) == -1
    assert lib.mica_label_aa_mask(None, None, 0, 0, 0, 0, 3, 3, 3, 0, 4, 4, None, 0, None, None, None) == -1
    assert lib.mica_zero_around_atoms(None, 1, None, None, 2.0, 4, 4, 4, None, None, None) == -1
    assert lib.mica_contour_threshold_f32(None, None, 8, 0.1, None) == -1
    assert lib.mica_contour_threshold_f32(None, None, 0, 0.1, None) == 0          # nothing to do
    assert lib.mica_cand_threshold_workspace_bytes(480 ** 3) > 27000 * 12
    assert lib.mica_label_aa_mask_workspace_bytes(4, 5, 6) == 3 * 4 * 120
