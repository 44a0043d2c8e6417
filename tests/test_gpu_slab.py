"""The z-slab partitioned path, emulated rank by rank on ONE GPU (the histogram all-reduce
is replaced by an explicit sum between the hist and pick kernels of every rank, in
lock-step), must reproduce the single-GPU pipeline and the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from mica_b200 import _lib, ops, synthetic
from mica_b200.pdb import channel_codes
from mica_b200.pipeline import MapHeader, MapPipeline
from mica_b200.slab import SlabPipeline, SlabPlan
from oracle import mica_oracle as orc

pytestmark = pytest.mark.gpu


def _lockstep_stats(ranks, owned, n_total):
    """What ops.OrderStats.run(all_reduce=...) does on every rank of a real group."""
    lib = _lib.lib
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    stats = [ops.OrderStats(o.device) for o in owned]
    for s in stats:
        _lib.check(lib.mica_select_init(s._p, n_total, st))
    for step in range(_lib.SELECT_PASSES):
        for s, o in zip(stats, owned):
            _lib.check(lib.mica_select_hist(C.c_void_p(o.data_ptr()), o.numel(), s._p, step, st))
        total = sum(s.hist_view().clone() for s in stats)
        for s in stats:
            s.hist_view().copy_(total)
            _lib.check(lib.mica_select_pick(s._p, step, st))
    return stats


@pytest.mark.parametrize('world', [2, 3])
@pytest.mark.parametrize('src_shape,voxel,gs,pad', [((120, 40, 36), 1.2, 32, 16), ((144, 32, 32), 1.0, 48, 8)])
def test_emulated_slab_ranks_match_single_gpu(cuda, world, src_shape, voxel, gs, pad):
    src = synthetic.synthetic_map(src_shape, voxel=voxel, seed=13)
    hdr = MapHeader(voxel_size=(np.float32(voxel),) * 3)
    d_src = torch.from_numpy(src).to(cuda)
    single = MapPipeline(cuda, gs, pad, batch_cubes=4)
    assert single.resample_and_normalize(d_src, hdr)
    n_out = tuple(single.normalized.shape)
    st = synthetic.synthetic_structure(150, n_out[::-1], seed=13)
    bb_ch, aa_ch = channel_codes(st['atom_names'], st['res_names'])
    atoms = tuple(torch.from_numpy(a).to(cuda) for a in (st['coords'], bb_ch, aa_ch))
    assert single.encode_af3(*atoms)
    ijk_all = single.cube_index()
    logits = [torch.from_numpy(a).to(cuda) for a in synthetic.synthetic_logits(len(ijk_all), gs + 2 * pad, seed=5)]
    lut = {tuple(v): n for n, v in enumerate(ijk_all)}

    def model_for(pipe):
        state = {'b': 0}

        def fn(x, af):
            rows = [lut[tuple(v)] for v in pipe.ijk_host[state['b']:state['b'] + x.shape[0]]]
            state['b'] += x.shape[0]
            idx = torch.tensor(rows, device=cuda)
            return tuple(t[idx] for t in logits)
        return fn

    want = single.predict_and_stitch(model_for(single))

    # ---- emulated ranks
    pipes = [SlabPipeline(cuda, r, world, gs, pad, batch_cubes=4, global_src_shape=src_shape) for r in range(world)]
    res, owned = [], []
    for p in pipes:
        p._exchange = lambda own, p=p: d_src[p.plan.ranks[p.rank].src_lo:p.plan.ranks[p.rank].src_hi].contiguous()
        # each rank is handed its own block of source planes; the plan is rebuilt inside slab_resample
        pl = SlabPlan(src_shape, hdr.voxel_size, gs, pad, world)
        me = pl.ranks[p.rank]
        r_, o_ = p.slab_resample(d_src[me.own_lo:me.own_hi].contiguous(), hdr)
        res.append(r_)
        owned.append(o_)
    stats = _lockstep_stats(pipes, owned, int(np.prod(n_out)))
    got_norm = torch.empty(n_out, device=cuda)
    vols = {k: torch.zeros_like(v) for k, v in want.as_dict().items()}
    for p, r_, s in zip(pipes, res, stats):
        p.stats = s
        p.slab_normalize(r_)
        assert p.check_status()
        assert p.median == single.median and p.p999 == single.p999       # thresholds agree exactly
        me = p.plan.ranks[p.rank]
        got_norm[me.out_lo:me.out_hi] = p.normalized[me.out_lo - me.ext_lo:me.out_hi - me.ext_lo]
        assert p.encode_af3(*atoms)
        v = p.predict_and_stitch(model_for(p))
        (o0, o1, o2), (e0, e1, e2) = p.box
        for k, t in v.as_dict().items():
            vols[k][..., o2:o2 + e2] = t
    assert (got_norm - single.normalized).abs().max().item() <= 1e-6
    for k, t in want.as_dict().items():
        if k == 'amino_acid_prediction':
            assert (vols[k] == t).float().mean().item() == 1.0
        else:
            assert torch.equal(vols[k], t), k          # same logits, same cores -> identical volumes
    o_norm, _, _ = orc.normalize(orc.resample(src, hdr.voxel_size))
    assert np.abs(got_norm.cpu().numpy() - o_norm).max() <= 1e-5


@pytest.mark.parametrize('world', [2, 3, 8])
def test_peer_memory_histogram_exchange_equals_single_gpu(cuda, world):
    """mica_select_peer_reduce (publish / signal / wait / sum in one kernel over peer-mapped buffers)
    with `world` emulated ranks on their own streams: thresholds must equal the whole-array run."""
    import ctypes as C
    from mica_b200.peer import PeerHistogram
    lib = _lib.lib
    g = torch.Generator(device=cuda).manual_seed(world)
    x = torch.randn(3_000_001, generator=g, device=cuda) * 0.05
    x[::50] += torch.rand(x[::50].shape, generator=g, device=cuda)
    want = ops.OrderStats(cuda).run(x).result()
    bounds = [0] + sorted(int(v) // 4 * 4 for v in torch.randint(1, x.numel() - 1, (world - 1,)).tolist()) + [x.numel()]
    parts = [x[bounds[r]:bounds[r + 1]] for r in range(world)]
    groups = PeerHistogram.emulate(cuda, world)
    streams = [torch.cuda.Stream(cuda) for _ in range(world)]
    stats = [ops.OrderStats(cuda) for _ in range(world)]
    torch.cuda.synchronize()
    for repeat in range(2):                              # epochs keep counting across maps
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                st = C.c_void_p(streams[r].cuda_stream)
                _lib.check(lib.mica_select_init(stats[r]._p, x.numel(), st))
        for rnd in range(_lib.SELECT_PASSES):
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    st = C.c_void_p(streams[r].cuda_stream)
                    _lib.check(lib.mica_select_hist(C.c_void_p(parts[r].data_ptr()), parts[r].numel(), stats[r]._p, rnd, st))
                    groups[r].reduce(stats[r], rnd, streams[r].cuda_stream)
                    _lib.check(lib.mica_select_pick(stats[r]._p, rnd, st))
        torch.cuda.synchronize()
        for r in range(world):
            got = stats[r].result()
            assert got[3] == 0, f'rank {r}: status {got[3]}'
            assert got[:3] == want[:3], (r, got, want)
    groups[0].close()


@pytest.mark.parametrize('world', [2, 3, 8])
def test_peer_memory_halo_exchange_assembles_the_source_planes(cuda, world):
    """mica_halo_publish / mica_halo_pull with `world` emulated ranks (own streams, plain pointers instead of
    IPC): every rank's assembled buffer must equal its planes [src_lo, src_hi) of the whole map, over
    several epochs (slot parity) and with maps that change from epoch to epoch."""
    from mica_b200.peer import PeerHalo
    src_shape, voxel, gs, pad = (96 * world, 24, 20), 1.1, 32, 16
    plan = SlabPlan(src_shape, (np.float32(voxel),) * 3, gs, pad, world)
    plane = src_shape[1] * src_shape[2]
    need = 0
    for r in range(world):
        np_ = PeerHalo.neighbour_plan(plan, r)
        assert np_ is not None
        need = max([need] + [b - a for rng in np_ if rng is not None for a, b in [rng]])
    groups = PeerHalo.emulate(cuda, world, need * plane)
    assert all(g.fits(plan, g.rank) for g in groups)
    streams = [torch.cuda.Stream(cuda) for _ in range(world)]
    g = torch.Generator(device=cuda).manual_seed(3)
    for epoch in range(3):
        full = torch.rand(src_shape, generator=g, device=cuda)
        torch.cuda.synchronize()
        owns, bufs = [], []
        for r in range(world):                            # publish never waits: issue all of them first
            me = plan.ranks[r]
            own = full[me.own_lo:me.own_hi].contiguous()
            owns.append(own)
            with torch.cuda.stream(streams[r]):
                groups[r].publish(own, plan, streams[r].cuda_stream)
        for r in range(world):
            me = plan.ranks[r]
            with torch.cuda.stream(streams[r]):
                buf = torch.full((me.src_hi - me.src_lo,) + src_shape[1:], -1.0, device=cuda)
                lo, hi = max(me.src_lo, me.own_lo), min(me.src_hi, me.own_hi)
                buf[lo - me.src_lo:hi - me.src_lo].copy_(owns[r][lo - me.own_lo:hi - me.own_lo])
                groups[r].pull(buf, plan, streams[r].cuda_stream)
            bufs.append(buf)
        torch.cuda.synchronize()
        for r in range(world):
            me = plan.ranks[r]
            assert not groups[r].timed_out()
            assert torch.equal(bufs[r], full[me.src_lo:me.src_hi]), (epoch, r)
    groups[0].close()


def test_next_map_halo_prefetch_is_consumed_and_bit_identical(cuda):
    """SlabPipeline.prefetch_source: the halo of map k+1 is exchanged on a side stream while map k is still
    in flight; the next slab_resample must take that buffer (no second exchange) and produce the planes of the
    whole-map resample bit for bit (aligned plan, fast path).  A prefetch for a block that is not the one that
    follows is dropped and the exchange runs again.  Two emulated ranks on their own streams."""
    from mica_b200.peer import PeerHalo
    world, src_shape, voxel = 2, (208, 96, 100), 1.1
    hdr = MapHeader(voxel_size=(np.float32(voxel),) * 3)
    plan = SlabPlan(src_shape, hdr.voxel_size, 48, 8, world)
    plane = src_shape[1] * src_shape[2]
    need = max(b - a for r in range(world) for rng in PeerHalo.neighbour_plan(plan, r) if rng is not None
               for a, b in [rng])
    groups = PeerHalo.emulate(cuda, world, need * plane)
    calls = [0] * world
    pipes = []
    for r in range(world):
        p = SlabPipeline(cuda, r, world, 48, 8, global_src_shape=src_shape, group=False)
        p.peer_halo = groups[r]
        real = groups[r].exchange

        def counted(own, plan_, buf=None, *, real=real, r=r):
            calls[r] += 1
            return real(own, plan_, buf)
        groups[r].exchange = counted
        pipes.append(p)
    streams = [torch.cuda.Stream(cuda) for _ in range(world)]
    g = torch.Generator(device=cuda).manual_seed(11)
    maps = [torch.rand(src_shape, generator=g, device=cuda) for _ in range(4)]
    owns = [[m[me.own_lo:me.own_hi].contiguous() for me in plan.ranks] for m in maps]
    want = [ops.resample(m, plan.out_shape) for m in maps]
    torch.cuda.synchronize()
    got = {}
    # maps 0 -> 1 -> 2 with the next map announced; then map 3 although map 0 was announced (stale prefetch)
    for k, nxt in ((0, 1), (1, 2), (2, 0), (3, None)):
        for r, p in enumerate(pipes):
            with torch.cuda.stream(streams[r]):
                got[k, r], _ = p.slab_resample(owns[k][r], hdr)
                if nxt is not None:
                    p.prefetch_source(owns[nxt][r], hdr)
    torch.cuda.synchronize()
    assert calls == [5, 5]            # 0: exchange, 1 and 2: prefetched, 0 again: prefetched and dropped, 3: exchange
    for (k, r), t in got.items():
        me = plan.ranks[r]
        assert not groups[r].timed_out()
        assert torch.equal(t, want[k][me.ext_lo:me.ext_hi]), (k, r)
    groups[0].close()


def test_peer_halo_times_out_instead_of_hanging(cuda, monkeypatch):
    """A neighbour that never publishes turns into a status word after the bounded spin (MICA_PEER_TIMEOUT_MS;
    60 s by default -- ranks of a real job can be seconds apart)."""
    from mica_b200.peer import PeerHalo
    monkeypatch.setenv('MICA_PEER_TIMEOUT_MS', '1000')
    plan = SlabPlan((80, 8, 8), (np.float32(1.1),) * 3, 32, 16, 2)
    groups = PeerHalo.emulate(cuda, 2, 64 * 64)
    me = plan.ranks[0]
    buf = torch.zeros((me.src_hi - me.src_lo, 8, 8), device=cuda)
    groups[0].epoch = 1                                   # rank 1 never published epoch 1
    import time
    t0 = time.time()
    groups[0].pull(buf, plan)
    torch.cuda.synchronize()
    assert groups[0].timed_out() and time.time() - t0 < 20
    groups[0].close()


@pytest.mark.parametrize('world', [2, 3])
def test_balanced_cube_ranks_store_cores_into_the_owners_volumes(cuda, world):
    """config 5 dataflow, emulated: the cubes of ONE map dealt out evenly, every core stored into the volume
    block of the rank that owns its x range (postproc_stitch_peer); the union of the blocks must be bit-equal
    to the single-GPU volumes."""
    from mica_b200.peer import PeerVolumes
    from mica_b200.slab import BalancedCubePipeline
    src = synthetic.synthetic_map((56, 40, 100), voxel=1.0, seed=21)
    hdr = MapHeader(voxel_size=(np.float32(1.0),) * 3)
    d_src = torch.from_numpy(src).to(cuda)
    st = synthetic.synthetic_structure(200, (100, 40, 56), seed=21)
    bb_ch, aa_ch = channel_codes(st['atom_names'], st['res_names'])
    atoms = tuple(torch.from_numpy(a).to(cuda) for a in (st['coords'], bb_ch, aa_ch))
    single = MapPipeline(cuda, 32, 16, batch_cubes=5)
    with torch.no_grad():
        want = single.run(d_src, hdr, atoms, synthetic.pointwise_model)
    X, Y, Z = single.cube_shape
    pipes = [BalancedCubePipeline(cuda, r, world, 32, 16, batch_cubes=5, group=False) for r in range(world)]
    for p in pipes:
        assert p.resample_and_normalize(d_src, hdr) and p.encode_af3(*atoms)
        p.cube_index()
    n_max = max(pipes[0].x_bounds[r + 1] - pipes[0].x_bounds[r] for r in range(world)) * Y * Z
    groups = PeerVolumes.emulate(cuda, world, n_max)
    vols = []
    for p, g_ in zip(pipes, groups):                       # every owner clears its block BEFORE anybody stores
        p.peer_volumes = g_
        vols.append(p._new_volumes())
    assert sum(len(p.ijk_host) for p in pipes) == len(single.ijk_host)
    assert max(p.cubes_per_rank) - min(p.cubes_per_rank) <= 1
    with torch.no_grad():
        for p, v in zip(pipes, vols):
            p.predict_and_stitch(synthetic.pointwise_model, v, model_batch=2, d8='split')
    torch.cuda.synchronize()
    for r, (p, v) in enumerate(zip(pipes, vols)):
        x0, x1 = p.x_bounds[r], p.x_bounds[r + 1]
        for k, t in v.as_dict().items():
            ref = want.as_dict()[k]
            ref = ref[:, x0:x1] if ref.dim() == 4 else ref[x0:x1]
            assert torch.equal(t, ref), (r, k)
    groups[0].close()
