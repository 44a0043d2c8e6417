"""TEST INFRASTRUCTURE ONLY (never imported by mica_b200/): a NumPy statement of the finite-horizon segment
scheme the fast resample path uses for the z / y prefilter (mica_b200/csrc/resample.cu: reg_segment,
cols_reg_segment, cols_reg_kernel, launch_cols_reg), next to what it replaces -- SciPy's whole-line recursion
(scipy.ndimage.spline_filter1d(order=3, mode='mirror'), the prefilter inside the zoom() call at
/root/reference/utils/preprocessing.py:117).

What it pins on the CPU:
  * accuracy of the scheme (16-sample horizon, float32 far run-ins, float32 storage) against SciPy;
  * the property the z-slab partition relies on: segments are numbered on the WHOLE line, so a block of the
    line that holds the windows of the segments it needs (mica_resample_slab_source_planes) yields the same
    coefficients as the whole line, bit for bit -- and a block cut into its OWN segments does not.

The arithmetic follows the kernel operation by operation, except that NumPy has no fused multiply-add: a * b + c
is rounded twice here, once on the GPU.  The bit-identity property does not depend on that (same operations on the
same inputs in the same order), the accuracy figure does not either (1e-16 against 6e-8)."""
import numpy as np

POLE = np.float64(-0.26794919243112270647)          # sqrt(3) - 2
GAIN = (1.0 - POLE) * (1.0 - 1.0 / POLE)            # = 6
COL_LEN, COL_H, NEAR = 26, 16, 4                    # kColLen, kColH, kNear


def geometry(ng):
    """launch_cols_reg: number and length of the segments of a line of ng samples."""
    n_seg = -(-ng // COL_LEN)
    return n_seg, -(-ng // n_seg)


def _mirror(idx, n):
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx > n - 1, 2 * (n - 1) - idx, idx)
    return np.maximum(idx, 0)


def _segment(x32, k0, k1, g0, ng):
    """reg_segment for the segment [k0, k1) (global numbers) of lines x32 [lines, n] = samples [g0, g0 + n).
    Returns float32 [lines, k1 - k0]."""
    n = x32.shape[1]
    H, F = COL_H, COL_H - NEAR
    W = COL_LEN + 2 * H
    idx = _mirror(np.arange(k0 - H, k0 - H + W), ng)        # the window, mirrored at the ends of the LINE
    loc = _mirror(idx - g0, n)                              # ... read from the block (reflected if it is too short)
    x = x32[:, loc]                                         # float32 [lines, W]
    zf, z = np.float32(POLE), POLE
    run = np.zeros(x.shape[0], np.float32)
    for i in range(F):                                      # causal run-in, far part: float32
        run = (zf.astype(np.float64) * run + x[:, i]).astype(np.float32)
    st = run.astype(np.float64)
    for i in range(F, H):                                   # near part: float64
        st = z * st + x[:, i]
    cp = np.empty((x.shape[0], COL_LEN + NEAR))
    for i in range(COL_LEN + NEAR):
        st = z * st + x[:, H + i]
        cp[:, i] = st
    run = st.astype(np.float32)
    ahead = np.empty((x.shape[0], F), np.float32)
    for i in range(F):                                      # causal values of the far look-ahead samples: float32
        run = (zf.astype(np.float64) * run + x[:, H + COL_LEN + NEAR + i]).astype(np.float32)
        ahead[:, i] = run
    back = np.zeros(x.shape[0], np.float32)
    for i in range(F - 1, -1, -1):                          # anticausal run-in, far part
        back = (zf.astype(np.float64) * back + ahead[:, i]).astype(np.float32)
    d = back.astype(np.float64)
    for i in range(COL_LEN + NEAR - 1, COL_LEN - 1, -1):
        d = z * d + cp[:, i]
    out = np.empty((x.shape[0], COL_LEN), np.float32)
    scale = -z * GAIN
    for i in range(COL_LEN - 1, -1, -1):
        d = z * d + cp[:, i]
        out[:, i] = (d * scale).astype(np.float32)
    return out[:, :k1 - k0]


def prefilter_block(block32, g0, ng, need_lo, need_hi, own_segments=False):
    """Coefficients of global samples of the segments that hold [need_lo, need_hi], computed from the block
    ``block32`` [lines, n] = samples [g0, g0 + n) of lines of ``ng`` samples.  Returns (first_sample, float32
    [lines, count]).  own_segments=True is the round-2-early behaviour: the block is cut into ITS OWN segments
    and mirrored at its own ends (what made slabs differ from the whole map in the last bits)."""
    if own_segments:
        n = block32.shape[1]
        first, coeff = prefilter_block(block32, 0, n, max(0, need_lo - g0), min(n - 1, need_hi - g0))
        return first + g0, coeff
    _, seg = geometry(ng)
    parts = []
    for s in range(need_lo // seg, need_hi // seg + 1):
        k0 = s * seg
        parts.append(_segment(block32, k0, min(ng, k0 + seg), g0, ng))
    return need_lo // seg * seg, np.concatenate(parts, axis=1)
