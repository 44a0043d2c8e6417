"""BASELINE.json configs 1-5 at their full sizes (config 1 = configs[0], the reference's own CPU-runnable case,
is small enough for the oracle to run the whole path).  Where the CPU oracle finishes in seconds it is
used directly (order statistics of 110.6 M voxels, atom indices of 20 k residues); elsewhere the
check is a size-independent property of the path: extract -> stitch is the identity, the sparse
AF3 cube fill equals the dense volume's windows, resampling is linear and reproduces constants,
a pointwise model commutes with cutting and stitching, N slab ranks reproduce one GPU."""
import numpy as np
import pytest
import torch

from mica_b200 import ops, synthetic
from mica_b200.pdb import channel_codes
from mica_b200.pipeline import MapHeader, MapPipeline
from mica_b200.slab import SlabPipeline, SlabPlan
from oracle import mica_oracle as orc
from test_gpu_slab import _lockstep_stats

pytestmark = pytest.mark.gpu


def _smooth_random(shape, device, seed):
    """Density-like volume made on the device: positive blobs + noise (positive tail, many distinct values)."""
    g = torch.Generator(device=device).manual_seed(seed)
    v = torch.randn(shape, generator=g, device=device) * 0.05
    blobs = torch.rand(shape, generator=g, device=device)
    v += torch.where(blobs > 0.97, blobs * 4 - 3.5, torch.zeros((), device=device))
    return torch.nn.functional.avg_pool3d(v[None, None], 3, 1, 1)[0, 0].contiguous()


# ----------------------------------------------------------------------- config 1: 200^3 @ 1.06 A -> 212^3
def test_config1_200_map_whole_path_against_the_oracle(cuda):
    """BASELINE configs[0]: synthetic 200^3 map, 1.06 A voxel, 3.7 A resolution, reference defaults
    (grid_size 48, padding 8 -> 125 cubes): every stage of the GPU path against the oracle's."""
    src = synthetic.synthetic_map((200, 200, 200), voxel=1.06, resolution=3.7, seed=2022)
    voxel = (np.float32(1.06),) * 3
    hdr = MapHeader(voxel_size=voxel)
    st = synthetic.synthetic_structure(1700, (212, 212, 212), seed=2022)
    bb_ch, aa_ch = channel_codes(st['atom_names'], st['res_names'])
    atoms = tuple(torch.from_numpy(a).to(cuda) for a in (st['coords'], bb_ch, aa_ch))
    ring = synthetic.synthetic_logits(16, 64, seed=2022)
    d_ring = [torch.from_numpy(a).to(cuda) for a in ring]
    seen = {}

    def model_fn(x, af):
        seen[len(seen)] = (x[:2].cpu().numpy(), af[:2].cpu().numpy())       # first two cubes of every batch
        return tuple(t[:x.shape[0]] for t in d_ring)

    pipe = MapPipeline(cuda, 48, 8, batch_cubes=16)
    vols = pipe.run(torch.from_numpy(src).to(cuda), hdr, atoms, model_fn)
    assert tuple(pipe.normalized.shape) == (212, 212, 212) and len(pipe.ijk_host) == 125
    # R1-R3
    o_norm, med, p = orc.normalize(orc.resample(src, voxel))
    assert np.abs(pipe.normalized.cpu().numpy() - o_norm).max() <= 1e-5
    # R4 (bit-exact) and R5/R6 on the cubes the model saw
    o_af3, ok = orc.af3_encode(st['coords'], bb_ch, aa_ch, (0.0, 0.0, 0.0), o_norm.shape)
    assert ok
    o_cubes, meta, shp, _ = orc.extract_cubes(o_norm)
    for b, (x, af) in seen.items():
        for r in range(len(x)):
            c = 16 * b + r
            assert np.abs(x[r, 0] - o_cubes[c]).max() <= 1e-5
            i, j, k = (int(v) for v in meta[c][:3])
            want = np.zeros((24, 64, 64, 64), np.float32)
            t = np.transpose(o_af3, (0, 3, 2, 1))
            xs, ys, zs = (slice(max(0, a - 8), min(212, a + 56)) for a in (i, j, k))
            want[:, xs.start - i + 8:xs.stop - i + 8, ys.start - j + 8:ys.stop - j + 8,
                 zs.start - k + 8:zs.stop - k + 8] = t[:, xs, ys, zs]
            assert np.array_equal(af[r], want), c
    # R7/R8: reference post-processing + stitching of the same logits, 16 cubes at a time
    want = {'backbone_probability': np.zeros(shp, np.float32), 'carbon_alpha_probability': np.zeros(shp, np.float32),
            'amino_acid_prediction': np.zeros(shp, np.float32), 'amino_acid_probability': np.zeros((20,) + tuple(shp), np.float32)}
    for c0 in range(0, 125, 16):
        n = min(16, 125 - c0)
        bb, ca, aa_prob, aa_pred = orc.postprocess(ring[0][:n], ring[1][:n], ring[2][:n])
        for name, pred in (('backbone_probability', bb), ('carbon_alpha_probability', ca),
                           ('amino_acid_prediction', aa_pred), ('amino_acid_probability', aa_prob)):
            orc.stitch(pred, meta[c0:c0 + n], shp, name, 8, volume=want[name])
    for k in ('backbone_probability', 'carbon_alpha_probability', 'amino_acid_probability'):
        assert np.abs(vols.as_dict()[k].cpu().numpy() - want[k]).max() <= 1e-5, k
    top2 = np.sort(want['amino_acid_probability'], axis=0)[-2:]
    clear = (top2[1] - top2[0]) > 4e-6                             # equal wherever the top-2 gap is not rounding
    assert np.array_equal(vols.amino_acid_prediction.cpu().numpy()[clear], want['amino_acid_prediction'][clear])
    assert clear.mean() > 0.999


# ----------------------------------------------------------------------- config 2: 400^3 -> 480^3
@pytest.fixture(scope='module')
def map400(cuda):
    return _smooth_random((400, 400, 400), cuda, 2022)


def test_config2_resample_is_linear_and_reproduces_constants(cuda, map400):
    out_shape = ops.zoom_output_shape((400, 400, 400), [np.float32(1.2)] * 3)
    assert out_shape == (480, 480, 480)
    f = map400
    g = torch.roll(map400, (17, -5, 3), (0, 1, 2)) * 0.5
    rf, rg = ops.resample(f, out_shape), ops.resample(g, out_shape)
    rc = ops.resample(2 * f + 3 * g, out_shape)
    scale = float(rc.abs().max())
    assert float((rc - (2 * rf + 3 * rg)).abs().max()) <= 2e-6 * scale
    const = ops.resample(torch.full((400, 400, 400), 0.375, device=cuda), out_shape)
    # mirror-extended constants are reproduced exactly by the prefilter + B-spline partition of unity
    assert float((const - 0.375).abs().max()) <= 1e-7
    # the end points of every axis are interpolation nodes: corners are copied
    for c in ((0, 0, 0), (-1, 0, -1), (0, -1, -1), (-1, -1, -1)):
        assert abs(float(rf[c]) - float(f[c])) <= 1e-6 * max(1.0, abs(float(f[c])))


def test_config2_long_lines_match_scipy(cuda):
    """400-sample lines through the segment-parallel prefilter and the TMA-fed march, against SciPy."""
    src = np.random.default_rng(5).normal(size=(48, 400, 200)).astype(np.float32)
    voxel = (np.float32(1.2),) * 3
    want = orc.resample(src, voxel)
    got = ops.resample(torch.from_numpy(src).to(cuda), want.shape).cpu().numpy()
    assert np.abs(got - want).max() <= 2e-6 * float(np.abs(want).max())


def test_fast_path_all_axes_match_scipy(cuda):
    """Every axis >= 96 samples: float32 z / y prefilter passes (register-only), fused x prefilter +
    interpolation, TMA-fed y/z march -- against SciPy, incl. an odd row length (no 16-byte aligned rows)
    and anisotropic zoom."""
    rng = np.random.default_rng(17)
    for shape, voxel in (((112, 130, 141), (1.2, 1.2, 1.2)), ((100, 97, 128), (1.06, 1.3, 0.9))):
        src = rng.normal(size=shape).astype(np.float32)
        src[10:20, 30:40, 50:60] += 5.0
        voxel = tuple(np.float32(v) for v in voxel)
        want = orc.resample(src, voxel)
        got = ops.resample(torch.from_numpy(src).to(cuda), want.shape).cpu().numpy()
        assert np.abs(got - want).max() <= 2e-6 * float(np.abs(want).max()), shape


def test_config2_full_size_resample_matches_scipy(cuda):
    """The whole 400^3 -> 480^3 map of BASELINE configs[1] against scipy.ndimage.zoom (about a minute of
    CPU, once), and the normalised result against the oracle on the [0,1] scale (north_star: 1e-5)."""
    src = synthetic.synthetic_map((400, 400, 400), voxel=1.2, seed=2022)
    voxel = (np.float32(1.2),) * 3
    want = orc.resample(src, voxel)
    assert want.shape == (480, 480, 480)
    d = ops.resample(torch.from_numpy(src).to(cuda), want.shape)
    err = float(np.abs(d.cpu().numpy() - want).max())
    assert err <= 2e-6 * float(np.abs(want).max()), err
    o_norm, _, _ = orc.normalize(want)
    norm, st = ops.normalize(d)
    assert st.result()[3] == 0
    assert float(np.abs(norm.cpu().numpy() - o_norm).max()) <= 1e-5


def test_config2_order_statistics_of_110M_voxels_are_numpy_exact(cuda, map400):
    """np.median / np.percentile on the full 480^3 working grid (N = 110 592 000 > 2^24: the float32
    virtual-index path of NumPy 2) -- thresholds and the normalised volume bit for bit."""
    res = ops.resample(map400, (480, 480, 480))
    host = res.cpu().numpy()
    want, med, p = orc.normalize(host)
    norm, st = ops.normalize(res)
    gmed, gp, npos, status = st.result()
    assert status == 0
    assert np.float32(gmed).tobytes() == np.float32(med).tobytes() and np.float32(gp).tobytes() == np.float32(p).tobytes()
    assert npos == int((host > med).sum())
    assert np.array_equal(norm.cpu().numpy().view(np.uint32), want.view(np.uint32))
    # heavy ties: a masked map (half the voxels exactly 0) and a two-valued map
    masked = res * (torch.rand(res.shape, device=cuda) > 0.5)
    want, med, p = orc.normalize(masked.cpu().numpy())
    norm, st = ops.normalize(masked)
    gmed, gp, _, status = st.result()
    assert status == 0 and np.float32(gmed) == np.float32(med) and np.float32(gp) == np.float32(p)
    assert np.array_equal(norm.cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize('gs,pad', [(32, 16), (48, 8)])
def test_config2_extract_then_stitch_is_the_identity(cuda, gs, pad):
    g = torch.Generator(device=cuda).manual_seed(7)
    vol = torch.rand((480, 480, 480), generator=g, device=cuda)
    perm = ops.STANDARD_PERM
    shape = ops.cube_space_shape(vol.shape, perm)
    ijk = torch.from_numpy(ops.cube_origins(shape, gs)).to(cuda)
    assert len(ijk) == (3375 if gs == 32 else 1000)
    out = torch.zeros((1,) + shape, device=cuda)
    W = gs + 2 * pad
    core_sum = torch.zeros((), dtype=torch.int64, device=cuda)         # checksum of the float bit patterns
    for b0 in range(0, len(ijk), 125):
        sel = ijk[b0:b0 + 125]
        cubes = ops.extract_cubes(vol, sel, gs, pad, perm)
        assert cubes.shape == (len(sel), 1, W, W, W)
        core_sum += cubes[:, :, pad:pad + gs, pad:pad + gs, pad:pad + gs].contiguous().view(torch.int32).sum(dtype=torch.int64)
        ops.stitch_cubes(cubes, sel, shape, gs, pad, out=out)
    assert torch.equal(out[0], vol.permute(2, 1, 0))                 # stitched volumes are indexed [x,y,z]
    assert int(core_sum) == int(vol.view(torch.int32).sum(dtype=torch.int64))   # cores tile the volume exactly once
    # the first cube's halo is the zero padding of np.pad
    first = ops.extract_cubes(vol, ijk[:1], gs, pad, perm)[0, 0]
    assert float(first[:pad].abs().max()) == 0 and float(first[:, :pad].abs().max()) == 0


# ----------------------------------------------------------------------- config 3: 20 k residues -> 480^3
def test_config3_af3_encoding_of_20k_residues(cuda):
    shape = (480, 480, 480)
    st = synthetic.synthetic_structure(20000, shape[::-1], seed=2022)
    bb_ch, aa_ch = channel_codes(st['atom_names'], st['res_names'])
    assert len(st['coords']) > 150000
    want = orc.af3_indices(st['coords'], bb_ch, aa_ch, (0.0, 0.0, 0.0), shape)
    atoms = tuple(torch.from_numpy(a).to(cuda) for a in (st['coords'], bb_ch, aa_ch))
    vol, status = ops.af3_encode(*atoms, (0.0, 0.0, 0.0), shape)
    assert int(status.item()) == 0
    got = torch.nonzero(vol.reshape(-1)).flatten().cpu().numpy()
    assert np.array_equal(got, want)                                  # occupancy bit-exact (sorted linear indices)
    assert float(vol.reshape(-1)[torch.from_numpy(want).to(cuda)].min()) == 1.0
    # sparse per-cube fill == windows of the dense volume, every cube, both strides
    for gs, pad, B in ((32, 16, 135), (48, 8, 125)):
        ijk = torch.from_numpy(ops.cube_origins(ops.cube_space_shape(shape), gs)).to(cuda)
        filler = ops.Af3CubeFiller(cuda, B, gs, pad)
        assert int(filler.bin(*atoms, (0.0, 0.0, 0.0), shape).item()) == 0
        dense = torch.empty((B, 24, gs + 2 * pad, gs + 2 * pad, gs + 2 * pad), device=cuda)
        for b0 in range(0, len(ijk), B):
            sel = ijk[b0:b0 + B]
            ops.extract_cubes(vol, sel, gs, pad, out=dense[:len(sel)])
            assert torch.equal(filler.fill(sel), dense[:len(sel)]), (gs, b0)
        del filler, dense


# ----------------------------------------------------------------------- config 4: 720^3 over 2/4/8 slabs
@pytest.mark.parametrize('src_edge,voxel', [(720, 1.0), (679, 1.06)])
def test_config4_720_grid_slab_ranks_reproduce_one_gpu(cuda, src_edge, voxel):
    """Each rank (emulated one after the other on this GPU, histograms summed in lock-step as NCCL
    would) resamples and normalises only its z-slab; planes and thresholds must equal the 1-GPU run
    bit for bit."""
    src = _smooth_random((src_edge,) * 3, cuda, 4)
    hdr = MapHeader(voxel_size=(np.float32(voxel),) * 3)
    single = MapPipeline(cuda, 48, 8)
    assert single.resample_and_normalize(src, hdr)
    want = single.normalized
    assert tuple(want.shape) == (720, 720, 720)
    for world in (2, 4, 8):
        pipes = [SlabPipeline(cuda, r, world, 48, 8, global_src_shape=tuple(src.shape)) for r in range(world)]
        res, owned = [], []
        plan = SlabPlan(tuple(src.shape), hdr.voxel_size, 48, 8, world)
        assert [r.out_hi - r.out_lo for r in plan.ranks] == {2: [384, 336], 4: [192, 192, 192, 144],
                                                             8: [96] * 7 + [48]}[world]
        for p in pipes:
            me = plan.ranks[p.rank]
            p._exchange = lambda own, me=me: src[me.src_lo:me.src_hi]
            r_, o_ = p.slab_resample(src[me.own_lo:me.own_hi], hdr)
            res.append(r_)
            owned.append(o_)
        stats = _lockstep_stats(pipes, owned, want.numel())
        for p, r_, s in zip(pipes, res, stats):
            p.stats = s
            p.slab_normalize(r_)
            assert p.check_status()
            assert p.median == single.median and p.p999 == single.p999          # thresholds agree exactly
            me = p.plan.ranks[p.rank]
            # the slab's block holds whole prefilter windows (SlabPlan aligned=True): not merely close, the SAME bits,
            # cube halo planes included
            assert torch.equal(p.normalized, want[me.ext_lo:me.ext_hi]), (world, p.rank)
        del pipes, res, owned, stats


# ----------------------------------------------------------------------- config 5: 512^3 through the model loop
class _Pointwise(torch.nn.Module):
    """Stands where MICA stands (models/model.py:331: (exp_map, af_features) -> bb, ca, aa logits) but
    is voxel-wise, so model(cut(volume)) stitched must equal post-processing the whole volume at once."""

    def forward(self, x, af):
        s = af.sum(dim=1, keepdim=True)
        bb = torch.cat([x, -x, 2 * x - 0.5 + s, x * x], dim=1)
        ca = torch.cat([0.5 - x, x, x * 3 - 1, 1.5 * x + af[:, :1]], dim=1)
        aa = torch.cat([x * (0.1 * t) + af[:, t % 24:t % 24 + 1] * (t % 3) + float(np.sin(t)) for t in range(21)], dim=1)
        return bb, ca, aa


def test_config5_512_map_cut_model_stitch_commutes_with_a_pointwise_model(cuda):
    src = _smooth_random((512, 512, 512), cuda, 5)
    hdr = MapHeader(voxel_size=(np.float32(1.0),) * 3)                 # zoom 1: SciPy's copy path (D10)
    st = synthetic.synthetic_structure(6400, (512, 512, 512), seed=5)    # ~50 k atoms
    bb_ch, aa_ch = channel_codes(st['atom_names'], st['res_names'])
    atoms = tuple(torch.from_numpy(a).to(cuda) for a in (st['coords'], bb_ch, aa_ch))
    model = _Pointwise()
    pipe = MapPipeline(cuda, 48, 8, batch_cubes=11)
    with torch.no_grad():
        vols = pipe.run(src, hdr, atoms, model)
    assert len(pipe.ijk_host) == 1331 and tuple(vols.shape) == (512, 512, 512)
    # whole-volume evaluation, 64 x-planes at a time (one W^3-free "cube" per slab)
    dense, status = ops.af3_encode(*atoms, hdr.origin, (512, 512, 512))
    assert int(status.item()) == 0
    norm_t = pipe.normalized.permute(2, 1, 0)                           # [x,y,z]
    af_t = dense.permute(0, 3, 2, 1)
    with torch.no_grad():
        for x0 in range(0, 512, 64):
            bb, ca, aa = model(norm_t[None, None, x0:x0 + 64], af_t[None, :, x0:x0 + 64])
            wb = torch.softmax(torch.cat([bb[:, :1], bb[:, 2:]], 1), 1)[0, 2]
            wc = torch.softmax(torch.cat([ca[:, :1], ca[:, 2:]], 1), 1)[0, 2]
            wa = torch.softmax(aa[:, 1:], 1)[0]
            assert float((vols.backbone_probability[x0:x0 + 64] - wb).abs().max()) <= 1e-5
            assert float((vols.carbon_alpha_probability[x0:x0 + 64] - wc).abs().max()) <= 1e-5
            assert float((vols.amino_acid_probability[:, x0:x0 + 64] - wa).abs().max()) <= 1e-5
            top2 = wa.topk(2, 0).values
            clear = (top2[0] - top2[1]) > 4e-6
            pred = vols.amino_acid_prediction[x0:x0 + 64]
            assert bool((pred[clear] == wa.argmax(0)[clear].float()).all())
