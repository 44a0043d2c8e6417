"""Host-side logic of the z-slab partition on CPU: plan arithmetic, and
the source-halo exchange + histogram all-reduce over a world_size-2 gloo group."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mica_b200.slab import SlabPlan, exchange_source_halo


@pytest.mark.parametrize('world', [1, 2, 3, 4, 8])
@pytest.mark.parametrize('src,voxel,gs,pad', [((400, 40, 40), 1.2, 32, 16), ((720, 16, 16), 1.0, 48, 8),
                                             ((679, 16, 16), 1.06, 48, 8), ((100, 8, 8), 0.83, 32, 16)])
def test_plan_partitions_every_plane_once(world, src, voxel, gs, pad):
    plan = SlabPlan(src, (np.float32(voxel),) * 3, gs, pad, world)
    nz, sz = plan.out_shape[0], src[0]
    owned = np.zeros(nz, int)
    have = np.zeros(sz, int)
    scale = (sz - 1) / (nz - 1)
    for r in plan.ranks:
        owned[r.out_lo:r.out_hi] += 1
        have[r.own_lo:r.own_hi] += 1
        if r.out_hi > r.out_lo:
            assert r.out_lo % gs == 0 and (r.out_hi % gs == 0 or r.out_hi == nz)  # slab-aligned cores
        assert r.ext_lo == max(0, r.out_lo - pad) and r.ext_hi == min(nz, r.out_hi + pad) or r.out_hi == r.out_lo
        if r.out_hi > r.out_lo and not plan.identity:
            # every tap of every resampled plane, plus the prefilter horizon, is inside the slab or at a true edge
            lo_tap = int(np.floor(r.ext_lo * scale)) - 1
            hi_tap = int(np.floor((r.ext_hi - 1) * scale)) + 2
            assert r.src_lo <= max(0, lo_tap - 16) and r.src_hi >= min(sz, hi_tap + 16 + 1)
    assert (owned == 1).all() and (have == 1).all()
    # transfers are consistent: what r receives from p is what p sends to r
    for r in range(world):
        sends, recvs = plan.transfers(r)
        for p, a, b in recvs:
            assert (r, a, b) in plan.transfers(p)[0]


def test_identity_zoom_needs_only_cube_halo():
    plan = SlabPlan((720, 8, 8), (np.float32(1.0),) * 3, 48, 8, 4)
    assert plan.identity and plan.out_shape == (720, 8, 8)
    r = plan.ranks[1]
    assert (r.src_lo, r.src_hi) == (r.ext_lo, r.ext_hi) == (r.out_lo - 8, r.out_hi + 8)


def _worker(rank, world, port, src_shape, voxel, tmpdir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = np.random.default_rng(7).normal(size=src_shape).astype(np.float32)       # same on every rank
        plan = SlabPlan(src_shape, (np.float32(voxel),) * 3, 32, 16, world)
        me = plan.ranks[rank]
        own = torch.from_numpy(g[me.own_lo:me.own_hi].copy())
        buf = exchange_source_halo(own, plan, rank)
        assert np.array_equal(buf.numpy(), g[me.src_lo:me.src_hi]), 'halo exchange assembled the wrong planes'
        # histogram all-reduce: int64 sum of per-rank digit histograms equals the global one
        keys = (g.view(np.uint32) >> 21).astype(np.int64)
        local = torch.from_numpy(np.bincount(keys[me.own_lo:me.own_hi].ravel(), minlength=2048))
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
        assert np.array_equal(local.numpy(), np.bincount(keys.ravel(), minlength=2048))
        open(os.path.join(tmpdir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('src_shape,voxel', [((64, 6, 5), 1.2), ((50, 4, 4), 1.0), ((90, 3, 7), 0.9)])
def test_halo_exchange_and_hist_allreduce_gloo_world2(tmp_path, src_shape, voxel):
    port = 29500 + (os.getpid() + src_shape[0]) % 2000
    mp.spawn(_worker, args=(2, port, src_shape, voxel, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / 'ok0') and os.path.exists(tmp_path / 'ok1')
