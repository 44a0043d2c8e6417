"""CPU pin of the finite-horizon segment scheme of the fast resample path (oracle/segment_scheme.py): accuracy
against SciPy's whole-line prefilter, and the bit-identity of a z-slab's coefficients with the whole line's that
SlabPlan + mica_resample_slab_source_planes are there to guarantee (the GPU statement of the same property:
tests/test_gpu_fullsize.py::test_config4_720_grid_slab_ranks_reproduce_one_gpu)."""
import numpy as np
import pytest
from scipy import ndimage

from mica_b200 import ops
from mica_b200.slab import SlabPlan
from oracle import segment_scheme as seg


def _lines(n, lines=64, seed=0):
    g = np.random.default_rng(seed)
    x = g.normal(size=(lines, n)).astype(np.float32)
    x[: lines // 2] = ndimage.gaussian_filter1d(x[: lines // 2], 3.0, axis=1)      # half smooth, half white noise
    return x


@pytest.mark.parametrize('n', [96, 400, 679])
def test_scheme_matches_scipy_prefilter(n):
    x = _lines(n, seed=n)
    first, got = seg.prefilter_block(x, 0, n, 0, n - 1)
    want = ndimage.spline_filter1d(x.astype(np.float64), order=3, axis=1, mode='mirror')
    assert first == 0 and got.shape == want.shape
    # float32 storage (6e-8 relative to the coefficient) + horizon 0.27^16 = 7e-10 + float32 far run-ins 2e-10
    assert np.abs(got - want).max() <= 1.5e-7 * np.abs(want).max()


@pytest.mark.parametrize('src,voxel,world', [(679, 1.06, 8), (679, 1.06, 2), (400, 1.2, 3), (333, 0.9, 4)])
def test_slab_blocks_give_the_whole_lines_bits(src, voxel, world):
    x = _lines(src, seed=src + world)
    _, whole = seg.prefilter_block(x, 0, src, 0, src - 1)
    plan = SlabPlan((src, 8, 8), (np.float32(voxel),) * 3, 48, 8, world)
    nz = plan.out_shape[0]
    scale = (src - 1) / (nz - 1)
    differs_with_own_segments = 0
    for me in plan.ranks:
        lo_tap = max(0, int(np.floor(me.ext_lo * scale)) - 1)
        hi_tap = min(src - 1, int(np.floor((me.ext_hi - 1) * scale)) + 2)
        assert (me.src_lo, me.src_hi)[0] <= ops.resample_slab_source_planes(src, nz, me.ext_lo, me.ext_hi - me.ext_lo)[0]
        block = x[:, me.src_lo:me.src_hi]
        first, got = seg.prefilter_block(block, me.src_lo, src, lo_tap, hi_tap)
        assert np.array_equal(got, whole[:, first:first + got.shape[1]]), me          # the SAME bits
        first, own = seg.prefilter_block(block, me.src_lo, src, lo_tap, hi_tap, own_segments=True)
        own = own[:, lo_tap - first:hi_tap - first + 1]                               # the planes the taps read
        ref = whole[:, lo_tap:hi_tap + 1]
        assert np.abs(own - ref).max() <= 1e-6 * np.abs(whole).max()                  # close ...
        differs_with_own_segments += int((own != ref).sum())
    assert differs_with_own_segments > 0                                               # ... but not the same


def test_a_block_with_a_short_halo_is_reflected_not_read_out_of_bounds():
    """A caller that supplies only the 16 planes of horizon (SlabPlan(aligned=False), any external caller) still
    gets every coefficient its taps need to 1e-9 of the line's scale -- the reflected part of the window lies
    >= 16 samples away."""
    src, g0, n = 400, 120, 150
    x = _lines(src, seed=5)
    want = ndimage.spline_filter1d(x.astype(np.float64), order=3, axis=1, mode='mirror')
    need_lo, need_hi = g0 + 16, g0 + n - 1 - 16
    first, got = seg.prefilter_block(x[:, g0:g0 + n], g0, src, need_lo, need_hi)
    sel = slice(need_lo - first, need_hi - first + 1)
    assert np.abs(got[:, sel] - want[:, need_lo:need_hi + 1]).max() <= 1.5e-7 * np.abs(want).max()
