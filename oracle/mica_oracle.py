"""CPU oracle for the MICA voxel-parallel map pipeline.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mica_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` do, and there only as the checker / the timed
CPU reference -- never as a product path.

Every function restates, in NumPy/SciPy/torch-CPU, the arithmetic of one
reference function (cited ``file:line`` relative to ``/root/reference``) with the
file I/O (``.mrc`` / ``.npz`` round trips) removed.  Where the reference calls an
un-vendored third-party routine (``scipy.ndimage.zoom``, ``np.median``,
``np.percentile``) the oracle calls the very same installed routine, and a second
"restated" function spells the published algorithm out so that it can be checked
against the installed library (tests/test_oracle.py) and against golden vectors
produced by running the unmodified reference (oracle/make_golden.py ->
tests/golden/*.npz).

Pinning: parity is pinned against (a) the reference's own code executed in the
build container through oracle/ref_harness.py (fixtures committed under
tests/golden/), and (b) the installed NumPy 2.3 / SciPy 1.18 -- the reference
itself ships no tests or golden vectors.
"""
from __future__ import annotations

import math

import numpy as np

# ----------------------------------------------------------------------------
# constants of the reference
# ----------------------------------------------------------------------------
#: channel order, utils/preprocessing.py:254-263 == dataset/dataset.py:184-188
BACKBONE_ATOMS = ['CA', 'N', 'C', 'O']
AMINO_ACIDS = ['ALA', 'CYS', 'ASP', 'GLU', 'PHE', 'GLY', 'HIS', 'ILE', 'LYS', 'LEU',
               'MET', 'ASN', 'PRO', 'GLN', 'ARG', 'SER', 'THR', 'VAL', 'TRP', 'TYR']
CHANNEL_NAMES = BACKBONE_ATOMS + AMINO_ACIDS


# ----------------------------------------------------------------------------
# R1  resample  (utils/preprocessing.py:112-117; create_normalized_map.py:37-46)
# ----------------------------------------------------------------------------
def zoom_factors(voxel_size_xyz, target_voxel_size=1.0):
    """utils/preprocessing.py:112-114.  ``voxel_size.{x,y,z}`` are np.float32
    scalars (mrcfile recarray); dividing by a python float keeps float32 under
    NumPy 2.  The list [vx, vy, vz] is applied to array axes (0,1,2)=(z,y,x) (D9)."""
    vx, vy, vz = (np.float32(v) for v in voxel_size_xyz)
    return [vx / target_voxel_size, vy / target_voxel_size, vz / target_voxel_size]


def zoom_output_shape(in_shape, zf):
    """scipy.ndimage.zoom: ``int(round(ii * jj))`` -- int * np.float32 is float32
    under NumPy 2, round() is banker's rounding."""
    return tuple(int(round(ii * jj)) for ii, jj in zip(in_shape, zf))


def resample(data, voxel_size_xyz, target_voxel_size=1.0, order=3):
    """The reference's call: ``zoom(data, zoom_factors, order=3)``
    (utils/preprocessing.py:117).  order=1 is the north-star trilinear variant."""
    from scipy.ndimage import zoom
    return zoom(data, zoom_factors(voxel_size_xyz, target_voxel_size), order=order)


_POLE = math.sqrt(3.0) - 2.0


def _prefilter_axis_restated(c, axis):
    """Cubic B-spline IIR prefilter, mirror boundary, exact-sum causal init
    (SciPy ni_splines.c: apply_filter / _init_causal_mirror / _init_anticausal_mirror)."""
    c = np.moveaxis(c, axis, 0)
    n = c.shape[0]
    if n < 2:
        return np.moveaxis(c, 0, axis)
    z = _POLE
    c *= (1.0 - z) * (1.0 - 1.0 / z)
    # causal init: c0 = sum_k (z^k + z^(2n-2-k)) c[k] / (1 - z^(2n-2)), summed the
    # way SciPy's _init_causal_mirror does (pairs k and n-1-k share the factor z^k)
    z_n_1 = z ** (n - 1)
    acc = c[0] + z_n_1 * c[n - 1]
    z_i = z
    for k in range(1, n - 1):
        acc = acc + z_i * (c[k] + z_n_1 * c[n - 1 - k])
        z_i *= z
        if abs(z_i) < 1e-300:
            break
    c[0] = acc / (1.0 - z_n_1 * z_n_1)
    for k in range(1, n):
        c[k] += z * c[k - 1]
    c[n - 1] = (z / (z * z - 1.0)) * (c[n - 1] + z * c[n - 2])
    for k in range(n - 2, -1, -1):
        c[k] = z * (c[k + 1] - c[k])
    return np.moveaxis(c, 0, axis)


def _axis_taps(n_in, n_out, order):
    """Per-output-index tap start / weights along one axis (align-corners map)."""
    if n_out > 1:
        scale = (n_in - 1) / (n_out - 1)
    else:
        scale = 1.0
    k = np.arange(n_out, dtype=np.float64)
    x = k * scale
    if order == 3:
        f = np.floor(x)
        t = x - f
        w = np.stack([(1 - t) ** 3 / 6.0,
                      (3 * t ** 3 - 6 * t ** 2 + 4) / 6.0,
                      (-3 * t ** 3 + 3 * t ** 2 + 3 * t + 1) / 6.0,
                      t ** 3 / 6.0], axis=1)
        start = f.astype(np.int64) - 1
        ntap = 4
    elif order == 1:
        f = np.floor(x)
        t = x - f
        w = np.stack([1 - t, t], axis=1)
        start = f.astype(np.int64)
        ntap = 2
    else:
        raise ValueError(order)
    idx = start[:, None] + np.arange(ntap)[None, :]
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= n_in, 2 * (n_in - 1) - idx, idx)
    idx = np.clip(idx, 0, n_in - 1)
    return idx, w


def resample_restated(data, zf, order=3):
    """Published algorithm of ``scipy.ndimage.zoom(data, zf, order)`` with the
    defaults the reference uses (mode='constant', cval=0, prefilter=True,
    grid_mode=False); SURVEY.md Appendix A.  Float64 throughout, float32 result."""
    data = np.asarray(data)
    out_shape = zoom_output_shape(data.shape, zf)
    if all(float(z) == 1.0 for z in zf):       # SciPy >= 1.14 early exit (D10)
        return data.astype(data.dtype, copy=True)
    c = data.astype(np.float64)
    if order > 1:
        for a in range(3):
            c = _prefilter_axis_restated(c, a)
    # separable gather: contract one axis at a time (z, then y, then x)
    for a in range(3):
        idx, w = _axis_taps(data.shape[a], out_shape[a], order)
        c = np.moveaxis(c, a, 0)
        acc = np.zeros((out_shape[a],) + c.shape[1:], dtype=np.float64)
        for l in range(idx.shape[1]):
            acc += w[:, l].reshape((-1,) + (1,) * (c.ndim - 1)) * c[idx[:, l]]
        c = np.moveaxis(acc, 0, a)
    return c.astype(data.dtype if data.dtype == np.float32 else np.float64)


# ----------------------------------------------------------------------------
# R2/R3  normalise  (utils/preprocessing.py:121-133; create_normalized_map.py:48-79)
# ----------------------------------------------------------------------------
def normalize(resampled):
    """Verbatim arithmetic of utils/preprocessing.py:122-133.
    Returns (normalised float32 volume or None, median, percentile_value)."""
    norm_data = np.nan_to_num(resampled)
    median = np.median(norm_data)
    map_data_ = (norm_data > median) * (norm_data - median)
    positive_values = map_data_[np.where(map_data_ > 0)]
    if len(positive_values) == 0:
        return None, median, None
    percentile_value = np.percentile(positive_values, 99.9)
    if percentile_value == 0:
        return None, median, percentile_value
    map_data_ = (map_data_ < percentile_value) * map_data_ + \
                (map_data_ >= percentile_value) * percentile_value
    map_data_ /= percentile_value
    return map_data_.astype(np.float32), median, percentile_value


def order_stats_restated(x):
    """SURVEY.md section 8(a) R3 recipe: what the installed NumPy 2.x returns for
    ``np.median(x)`` and ``np.percentile(pos, 99.9)`` on float32 input, spelt out
    with a full sort.  Returns (median f32, percentile f32 or None, n_pos)."""
    f32 = np.float32
    x = np.nan_to_num(np.asarray(x, dtype=np.float32)).ravel()
    s = np.sort(x)
    n = s.size
    if n % 2:
        med = s[n // 2]
    else:
        med = f32(f32(s[n // 2 - 1] + s[n // 2]) / f32(2))
    n_le = int(np.searchsorted(s, med, side='right'))       # count(x <= med)
    npos = n - n_le
    if npos == 0:
        return med, None, 0
    q = f32(f32(99.9) / f32(100))
    vi = f32(f32(npos - 1) * q)
    prev = np.floor(vi)
    nxt = f32(prev + f32(1))
    if vi >= npos - 1:
        lo = hi = npos - 1
    else:
        lo, hi = int(prev), int(nxt)
    g = f32(vi - prev)
    a = f32(s[n_le + lo] - med)
    b = f32(s[n_le + hi] - med)
    d = f32(b - a)
    r = f32(a + f32(d * g))
    if g >= 0.5:
        r = f32(b - f32(d * f32(f32(1) - g)))
    return med, r, npos


# ----------------------------------------------------------------------------
# R4  AF3 24-channel rasteriser  (utils/preprocessing.py:172-178, 275-298)
# ----------------------------------------------------------------------------
def transform_coordinates(coord, origin_xyz, shape):
    """utils/preprocessing.py:172-178 verbatim (note the (z,y,x)-ordered ``shape``
    used to clip (x,y,z)-ordered indices -- D7)."""
    coord_shifted = coord - np.array(origin_xyz)
    indices = coord_shifted / 1.0
    indices = np.round(indices).astype(int)
    indices = np.clip(indices, 0, np.array(shape) - 1)
    return indices


def channel_codes(atom_names, res_names):
    """(bb_ch, aa_ch) int8 per atom: backbone channel 0..3 or -1
    (utils/preprocessing.py:292-293) and amino-acid channel 4..23 or -1 (:180-185)."""
    bb = np.array([BACKBONE_ATOMS.index(a) if a in BACKBONE_ATOMS else -1
                   for a in atom_names], dtype=np.int8)
    aa = np.array([4 + AMINO_ACIDS.index(r) if r in AMINO_ACIDS else -1
                   for r in res_names], dtype=np.int8)
    return bb, aa


def af3_indices(coords, bb_ch, aa_ch, origin_xyz, shape):
    """The voxels utils/preprocessing.py:275-298 sets to 1, as sorted unique linear indices into
    the C-ordered [24,nz,ny,nx] volume (or None on the IndexError path, :344-347).  Same index
    arithmetic as the loop; lets full-size grids be checked without the dense float64 volume."""
    coords = np.asarray(coords, dtype=np.float32)
    origin = np.array([np.float32(o) for o in origin_xyz], dtype=np.float32)
    # float32 - float32 stays float32 (Bio.PDB coords are float32; the reference's
    # np.array((origin.x, origin.y, origin.z)) of np.float32 scalars is float32)
    idx = np.round(coords - origin[None, :]).astype(int)
    idx = np.clip(idx, 0, np.array(shape)[None, :] - 1)
    nz, ny, nx = (int(v) for v in shape)
    # numpy negative indices cannot occur after the clip at 0; overflow raises
    bad = (idx[:, 2] >= nz) | (idx[:, 1] >= ny) | (idx[:, 0] >= nx)
    if bad.any():
        return None
    lin = (idx[:, 2].astype(np.int64) * ny + idx[:, 1]) * nx + idx[:, 0]
    n = np.int64(nz) * ny * nx
    bb_ch, aa_ch = np.asarray(bb_ch), np.asarray(aa_ch)
    parts = [bb_ch[bb_ch >= 0].astype(np.int64) * n + lin[bb_ch >= 0],
             aa_ch[aa_ch >= 0].astype(np.int64) * n + lin[aa_ch >= 0]]
    return np.unique(np.concatenate(parts))


def af3_encode(coords, bb_ch, aa_ch, origin_xyz, shape, dtype=np.float32):
    """utils/preprocessing.py:268-298 for atoms already filtered to standard
    residues (``residue.id[0] == ' '``).  ``coords`` float32 [A,3] (x,y,z),
    ``origin_xyz`` three np.float32.  Returns (volume [24,nz,ny,nx], ok) where
    ok=False reproduces the IndexError -> ``return False`` path (:344-347)."""
    vol = np.zeros((24,) + tuple(shape), dtype=dtype)
    lin = af3_indices(coords, bb_ch, aa_ch, origin_xyz, shape)
    if lin is None:
        return vol, False
    vol.reshape(-1)[lin] = 1.0
    return vol, True


# ----------------------------------------------------------------------------
# R5  cube extraction  (utils/create_grids.py:67-184)
# ----------------------------------------------------------------------------
def transpose_order(mapc, mapr, maps, nstart_zyx):
    """utils/create_grids.py:67-87,120-122: returns (trans_order, trans_offset)."""
    axis_order = [int(maps) - 1, int(mapr) - 1, int(mapc) - 1]
    offset = [float(v) for v in nstart_zyx]
    trans_offset, trans_order = [], []
    for i in range(3):
        for j in range(len(axis_order)):
            if axis_order[j] == i:
                trans_offset.append(offset[j])
                trans_order.append(j)
    return trans_order, trans_offset


def cube_origins(shape, grid_size):
    """Loop order of utils/create_grids.py:143-145 -> int list of (i,j,k)."""
    return [(i, j, k)
            for i in range(0, shape[0], grid_size)
            for j in range(0, shape[1], grid_size)
            for k in range(0, shape[2], grid_size)]


def extract_cubes(volume, mapc=1, mapr=2, maps=3, nstart_zyx=(0, 0, 0),
                  grid_size=48, padding=8, transpose=True, drop_below=None):
    """utils/create_grids.py:119-176 (``transpose=True``) or the training twin
    scripts_for_training_data/create_grids_for_normalized_map.py:40-100
    (``transpose=False``; ``drop_below=0.01`` reproduces its ``grid.max() >= 0.01``
    filter).  Returns (cubes [n,W,W,W], meta int64 [n,6]=(i,j,k,di,dj,dk),
    orig_shape, offset)."""
    if transpose:
        order, offset = transpose_order(mapc, mapr, maps, nstart_zyx)
        density_map = np.transpose(volume, order)
    else:
        density_map, offset = volume, None
    orig_shape = density_map.shape
    window = grid_size + 2 * padding
    pads = [(padding, window - (orig_shape[a] % grid_size)) for a in range(3)]
    padded = np.pad(density_map, pads, 'constant')
    cubes, meta = [], []
    for (i, j, k) in cube_origins(orig_shape, grid_size):
        di = min(grid_size, orig_shape[0] - i)
        dj = min(grid_size, orig_shape[1] - j)
        dk = min(grid_size, orig_shape[2] - k)
        grid = padded[i:i + window, j:j + window, k:k + window]
        if grid.shape != (window, window, window):
            continue
        if drop_below is not None and not (grid.max() >= drop_below):
            continue
        cubes.append(grid)
        meta.append((i, j, k, di, dj, dk))
    cubes = np.stack(cubes) if cubes else np.zeros((0, window, window, window), volume.dtype)
    return cubes, np.array(meta, dtype=np.int64).reshape(-1, 6), orig_shape, offset


# ----------------------------------------------------------------------------
# R7  post-processing of the model outputs  (utils/predict.py:342-349)
# ----------------------------------------------------------------------------
def postprocess(bb_logits, ca_logits, aa_logits):
    """torch-CPU restatement of utils/predict.py:342-349.  Inputs are numpy or
    torch [B,4,W,W,W], [B,4,W,W,W], [B,21,W,W,W]; returns numpy
    (bb_prob [B,W,W,W], ca_prob, aa_prob [B,20,W,W,W], aa_pred int64 [B,W,W,W])."""
    import torch
    softmax = torch.nn.Softmax(dim=1)
    bb = torch.as_tensor(bb_logits)
    ca = torch.as_tensor(ca_logits)
    aa = torch.as_tensor(aa_logits)
    bb = torch.cat((bb[:, :1], bb[:, 2:]), dim=1)
    bb_scores = softmax(bb)
    ca = torch.cat((ca[:, :1], ca[:, 2:]), dim=1)
    ca_scores = softmax(ca)
    aa_scores = softmax(aa[:, 1:, :, :, :])
    aa_predictions = torch.max(aa_scores, 1)[1]
    return (bb_scores[:, 2].numpy(), ca_scores[:, 2].numpy(),
            aa_scores.numpy(), aa_predictions.numpy())


# ----------------------------------------------------------------------------
# R8  stitching  (utils/predict.py:439-512)
# ----------------------------------------------------------------------------
def stitch(cube_preds, meta, orig_shape, map_type, padding=8, volume=None):
    """utils/predict.py:458-501: centre-crop paste of the disjoint cores.
    ``cube_preds`` [n,W,W,W] (or [n,20,W,W,W] for 'amino_acid_probability');
    ``meta`` rows (i,j,k,di,dj,dk).  The volume is float32 for every map type
    (int64 predictions are cast on assignment, :462).  ``volume``: paste into this
    array (a chunk of the cubes at a time) instead of a fresh zero volume."""
    if volume is None:
        if map_type == 'amino_acid_probability':
            volume = np.zeros((20, *orig_shape), dtype=np.float32)
        else:
            volume = np.zeros(tuple(orig_shape), dtype=np.float32)
    p = padding
    for grid, (i, j, k, di, dj, dk) in zip(cube_preds, meta):
        if map_type == 'amino_acid_probability':
            volume[:, i:i + di, j:j + dj, k:k + dk] = grid[:, p:p + di, p:p + dj, p:p + dk]
        else:
            volume[i:i + di, j:j + dj, k:k + dk] = grid[p:p + di, p:p + dj, p:p + dk]
    return volume


def postprocess_and_stitch(bb_logits, ca_logits, aa_logits, meta, orig_shape, padding=8):
    """R7 + R8 without the per-cube .npz round trip: the four volumes
    ``CryoEMPredictor.run_prediction`` returns (utils/predict.py:526-531)."""
    bb, ca, aa_prob, aa_pred = postprocess(bb_logits, ca_logits, aa_logits)
    return {
        'backbone_probability': stitch(bb, meta, orig_shape, 'backbone_probability', padding),
        'carbon_alpha_probability': stitch(ca, meta, orig_shape, 'carbon_alpha_probability', padding),
        'amino_acid_prediction': stitch(aa_pred, meta, orig_shape, 'amino_acid_prediction', padding),
        'amino_acid_probability': stitch(aa_prob, meta, orig_shape, 'amino_acid_probability', padding),
    }


def overlap_window(kind, grid_size, padding):
    """1-D window weights of the overlap-weighted stitch (NOT reference behaviour; BASELINE.json north_star
    variant): 'core' = the reference's crop (1 on the core, 0 on the halo), 'uniform' = plain
    average over every window that covers a voxel, 'triangle' = linear ramp peaking at the window centre."""
    W = grid_size + 2 * padding
    u = np.arange(W, dtype=np.float64)
    if kind == 'core':
        w = ((u >= padding) & (u < padding + grid_size)).astype(np.float64)
    elif kind == 'uniform':
        w = np.ones(W)
    elif kind == 'triangle':
        w = np.minimum(u + 1, W - u) / (W / 2.0)
    else:
        raise ValueError(kind)
    return w.astype(np.float32)


def postprocess_and_stitch_overlap(bb_logits, ca_logits, aa_logits, meta, orig_shape, grid_size, padding, w1):
    """NumPy statement of the overlap-weighted mode: vol = sum_cubes w * prob / sum_cubes w with the separable
    window w = w1[a] w1[b] w1[c]; amino_acid_prediction = arg-max of the averaged 20 probabilities.  float64
    accumulation (the kernel adds float32 atomically in arbitrary order: compare with a tolerance)."""
    bb, ca, aa_prob, _ = postprocess(bb_logits, ca_logits, aa_logits)
    X, Y, Z = (int(v) for v in orig_shape)
    W = grid_size + 2 * padding
    num = np.zeros((22, X, Y, Z), np.float64)
    den = np.zeros((X, Y, Z), np.float64)
    w3 = (w1[:, None, None].astype(np.float64) * w1[None, :, None] * w1[None, None, :])
    for n, row in enumerate(meta):
        i, j, k = (int(v) for v in row[:3])
        lo = [i - padding, j - padding, k - padding]
        sl_v, sl_c = [], []
        for a, (o, dim) in enumerate(zip(lo, (X, Y, Z))):
            v0, v1 = max(0, o), min(dim, o + W)
            sl_v.append(slice(v0, v1))
            sl_c.append(slice(v0 - o, v1 - o))
        sv, sc = tuple(sl_v), tuple(sl_c)
        w = w3[sc]
        num[(0,) + sv] += w * bb[n][sc]
        num[(1,) + sv] += w * ca[n][sc]
        num[(slice(2, 22),) + sv] += w[None] * aa_prob[n][(slice(None),) + sc]
        den[sv] += w
    ok = den > 0
    out = np.where(ok[None], num / np.where(ok, den, 1.0)[None], 0.0).astype(np.float32)
    pred = np.where(ok, out[2:].argmax(axis=0), 0).astype(np.float32)
    return {'backbone_probability': out[0], 'carbon_alpha_probability': out[1],
            'amino_acid_prediction': pred, 'amino_acid_probability': out[2:]}


# ----------------------------------------------------------------------------
# whole path, in memory (what bench.py's CPU arm times)
# ----------------------------------------------------------------------------
def pipeline_front(src, voxel_size_xyz, coords, bb_ch, aa_ch, origin_xyz,
                   grid_size=48, padding=8, order=3):
    """map -> resample -> normalise -> AF3 encode -> 25-channel cubes (R1..R6)."""
    res = resample(src, voxel_size_xyz, order=order)
    norm, med, p = normalize(res)
    if norm is None:
        raise RuntimeError('normalisation failed')
    af3, ok = af3_encode(coords, bb_ch, aa_ch, origin_xyz, norm.shape)
    if not ok:
        af3 = np.zeros_like(af3)
    cubes, meta, orig_shape, offset = extract_cubes(norm, grid_size=grid_size, padding=padding)
    af3_cubes = np.stack([extract_cubes(af3[c], grid_size=grid_size, padding=padding)[0]
                          for c in range(24)], axis=1)
    return norm, af3, cubes[:, None], af3_cubes, meta, orig_shape, offset


def pipeline_whole_streamed(src, voxel_size_xyz, coords, bb_ch, aa_ch, origin_xyz, logits_ring,
                            grid_size=48, padding=8, order=3):
    """The whole path map -> four stitched volumes with the cubes handled a chunk at a time, so a
    large sample fits in host memory (bench.py's CPU arm).  Same arithmetic as pipeline_front +
    postprocess_and_stitch; ``logits_ring`` = (bb [c,4,W^3], ca [c,4,W^3], aa [c,21,W^3]) stands
    where the model stands and is reused for every chunk of c cubes, as the GPU arm's ring is.
    Returns (volumes dict, number of working-grid voxels, number of cubes)."""
    res = resample(src, voxel_size_xyz, order=order)
    norm, med, p = normalize(res)
    if norm is None:
        raise RuntimeError('normalisation failed')
    af3, ok = af3_encode(coords, bb_ch, aa_ch, origin_xyz, norm.shape)
    if not ok:
        af3 = np.zeros_like(af3)
    order_, _ = transpose_order(1, 2, 3, (0, 0, 0))
    vol_t = np.transpose(norm, order_)                      # utils/create_grids.py:119-122
    af3_t = np.transpose(af3, (0,) + tuple(1 + o for o in order_))
    orig_shape = vol_t.shape
    W = grid_size + 2 * padding
    pads = [(padding, W - (orig_shape[a] % grid_size)) for a in range(3)]
    padded = np.pad(vol_t, pads, 'constant')               # :135-139, once per channel in the reference
    padded_af3 = np.pad(af3_t, [(0, 0)] + pads, 'constant')
    origins = cube_origins(orig_shape, grid_size)
    chunk = logits_ring[0].shape[0]
    vols = {'backbone_probability': np.zeros(orig_shape, np.float32),
            'carbon_alpha_probability': np.zeros(orig_shape, np.float32),
            'amino_acid_prediction': np.zeros(orig_shape, np.float32),
            'amino_acid_probability': np.zeros((20,) + tuple(orig_shape), np.float32)}
    for c0 in range(0, len(origins), chunk):
        sel = origins[c0:c0 + chunk]
        # the model's two inputs (dataset/dataset.py:194-224); materialised as the reference does
        x = np.stack([padded[i:i + W, j:j + W, k:k + W] for i, j, k in sel])[:, None]
        af = np.stack([padded_af3[:, i:i + W, j:j + W, k:k + W] for i, j, k in sel])
        assert x.shape[1:] == (1, W, W, W) and af.shape[1:] == (24, W, W, W)
        n = len(sel)
        bb, ca, aa_prob, aa_pred = postprocess(logits_ring[0][:n], logits_ring[1][:n], logits_ring[2][:n])
        meta = [(i, j, k, min(grid_size, orig_shape[0] - i), min(grid_size, orig_shape[1] - j),
                 min(grid_size, orig_shape[2] - k)) for i, j, k in sel]
        for name, pred in (('backbone_probability', bb), ('carbon_alpha_probability', ca),
                           ('amino_acid_prediction', aa_pred), ('amino_acid_probability', aa_prob)):
            stitch(pred, meta, orig_shape, name, padding, volume=vols[name])
    return vols, int(norm.size), len(origins)
