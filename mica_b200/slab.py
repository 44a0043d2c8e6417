"""z-slab partition of the map pipeline over the GPUs of one box.

The working grid is cut along memory axis 0 (z; cube axis k under the standard MRC
axis order) at multiples of ``grid_size`` so that every cube core -- and therefore
every stitched voxel -- has exactly one owner.  Per stage:

  resample      each rank resamples its owned output planes plus ``padding`` halo planes
                (the cube windows reach that far) from its source planes plus the
                interpolation taps plus ``halo_k`` planes of prefilter horizon; the source
                halo comes from the neighbouring ranks by send/recv (NCCL over NVLink).
  order stats   local histograms over OWNED planes only, all-reduced (int64 sum) between
                the hist and pick kernels of each of the 5 radix rounds: exact.
  AF3 encode    atoms are replicated (a few MB); each rank rasterises its slab + halo.
  extract       a rank's cubes are those whose core lies in its slab; no exchange (the halo
                planes were resampled and normalised locally).
  stitch        cores are slab-aligned: each rank writes only its own [X, Y, z-range] box.

There is no reference counterpart (the reference is single-process); results are checked
against the single-GPU path and the CPU oracle."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import ops
from ._lib import MicaError, NORM_OK
from .pipeline import MapHeader, MapPipeline, zoom_factors


def _split_even(n_items: int, parts: int):
    base, rem = divmod(n_items, parts)
    bounds = [0]
    for r in range(parts):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


@dataclass
class RankSlab:
    out_lo: int          # owned output planes [out_lo, out_hi)
    out_hi: int
    ext_lo: int          # resampled planes incl. the cube halo [ext_lo, ext_hi)
    ext_hi: int
    src_lo: int          # source planes needed [src_lo, src_hi)
    src_hi: int
    own_lo: int          # source planes this rank holds before the exchange [own_lo, own_hi)
    own_hi: int


class SlabPlan:
    """Pure host arithmetic of the partition (unit-tested on CPU)."""

    def __init__(self, src_shape, voxel_size_xyz, grid_size, padding, world, target_voxel_size=1.0,
                 halo_k=16, order=3):
        self.src_shape = tuple(int(v) for v in src_shape)
        self.world = int(world)
        self.grid_size, self.padding, self.halo_k, self.order = int(grid_size), int(padding), int(halo_k), order
        zf = zoom_factors(voxel_size_xyz, target_voxel_size)
        self.identity = all(float(z) == 1.0 for z in zf)
        self.out_shape = ops.zoom_output_shape(self.src_shape, zf)
        sz, nz = self.src_shape[0], self.out_shape[0]
        layers = -(-nz // self.grid_size)
        lb = _split_even(layers, self.world)
        ob = _split_even(sz, self.world)
        scale = (sz - 1) / (nz - 1) if nz > 1 else 1.0
        taps_lo, taps_hi = (1, 2) if order == 3 else (0, 1)
        k = self.halo_k if (order == 3 and not self.identity) else 0
        self.ranks = []
        for r in range(self.world):
            out_lo, out_hi = min(nz, lb[r] * self.grid_size), min(nz, lb[r + 1] * self.grid_size)
            ext_lo, ext_hi = max(0, out_lo - self.padding), min(nz, out_hi + self.padding)
            if out_hi <= out_lo:
                ext_lo = ext_hi = out_lo
                src_lo = src_hi = 0
            elif self.identity:
                src_lo, src_hi = ext_lo, ext_hi
            else:
                src_lo = max(0, int(np.floor(ext_lo * scale)) - taps_lo - k)
                src_hi = min(sz, int(np.floor((ext_hi - 1) * scale)) + taps_hi + k + 1)
            self.ranks.append(RankSlab(out_lo, out_hi, ext_lo, ext_hi, src_lo, src_hi, ob[r], ob[r + 1]))

    def transfers(self, rank):
        """(sends, recvs) for ``rank``: lists of (peer, plane_lo, plane_hi) in global source planes."""
        me = self.ranks[rank]
        sends, recvs = [], []
        for p, other in enumerate(self.ranks):
            if p == rank:
                continue
            lo, hi = max(me.src_lo, other.own_lo), min(me.src_hi, other.own_hi)
            if hi > lo:
                recvs.append((p, lo, hi))
            lo, hi = max(other.src_lo, me.own_lo), min(other.src_hi, me.own_hi)
            if hi > lo:
                sends.append((p, lo, hi))
        return sends, recvs


def exchange_source_halo(own: torch.Tensor, plan: SlabPlan, rank: int, group=None) -> torch.Tensor:
    """Assemble the source planes [src_lo, src_hi) this rank needs from its own block
    ``own`` (= global planes [own_lo, own_hi)) and its neighbours' (send/recv)."""
    import torch.distributed as dist
    me = plan.ranks[rank]
    n = max(0, me.src_hi - me.src_lo)
    buf = torch.empty((n,) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device)
    lo, hi = max(me.src_lo, me.own_lo), min(me.src_hi, me.own_hi)
    if hi > lo:
        buf[lo - me.src_lo:hi - me.src_lo].copy_(own[lo - me.own_lo:hi - me.own_lo])
    sends, recvs = plan.transfers(rank)
    ops_ = []
    keep = []
    for p, a, b in sends:
        t = own[a - me.own_lo:b - me.own_lo].contiguous()
        keep.append(t)
        ops_.append(dist.P2POp(dist.isend, t, p, group))
    for p, a, b in recvs:
        ops_.append(dist.P2POp(dist.irecv, buf[a - me.src_lo:b - me.src_lo], p, group))
    if ops_:
        for req in dist.batch_isend_irecv(ops_):
            req.wait()
    return buf


class SlabPipeline(MapPipeline):
    """One rank's share of a z-slab partitioned map.  ``run`` takes this rank's OWN block
    of source planes (global planes [own_lo, own_hi) of ``global_src_shape``)."""

    def __init__(self, device, rank, world, grid_size=48, padding=8, order=3, batch_cubes=16,
                 target_voxel_size=1.0, halo_k=16, global_src_shape=None, group=None, af3_mode='sparse',
                 hist_exchange='peer'):
        super().__init__(device, grid_size, padding, order, batch_cubes, target_voxel_size, af3_mode)
        self.rank, self.world, self.halo_k, self.group = int(rank), int(world), halo_k, group
        self.global_src_shape = global_src_shape
        self.plan = None
        #: 'peer' = one fused publish/signal/wait/sum kernel over NVLink peer memory per radix round
        #: (peer.PeerHistogram, built on first use); 'nccl' = torch.distributed.all_reduce
        if hist_exchange not in ('peer', 'nccl'):
            raise MicaError(f'hist_exchange must be peer or nccl, got {hist_exchange!r}')
        self.hist_exchange = hist_exchange
        self.peer = None

    def _peer_group(self):
        if self.peer is None:
            from .peer import PeerHistogram
            self.peer = PeerHistogram(self.device, self.rank, self.world, self.group)
        return self.peer

    # -- collectives (torch.distributed over NCCL; injectable for single-process emulation)
    def _all_reduce_hist(self, hist):
        import torch.distributed as dist
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=self.group)

    def _exchange(self, own):
        return exchange_source_halo(own, self.plan, self.rank, self.group)

    def make_plan(self, own_shape, header):
        gshape = self.global_src_shape
        if gshape is None:                      # weak-scaling default: equal blocks stacked along z
            gshape = (own_shape[0] * self.world, own_shape[1], own_shape[2])
        self.plan = SlabPlan(gshape, header.voxel_size, self.grid_size, self.padding, self.world,
                             self.target_voxel_size, self.halo_k, self.order)
        return self.plan

    def slab_resample(self, own_src, header=None):
        """Halo exchange + resample of this rank's planes.  Returns (res, owned): the local
        resampled planes [ext_lo, ext_hi) and the view of the owned ones [out_lo, out_hi)."""
        if header is not None:
            self.header = header
        plan = self.make_plan(tuple(own_src.shape), self.header)
        me = plan.ranks[self.rank]
        nz, ny, nx = plan.out_shape
        with self.timer('halo_exchange'):
            src = self._exchange(own_src)
        with self.timer('resample'):
            if plan.identity:
                res = src.clone()
            else:
                res = ops.resample(src, plan.out_shape, order=self.order, src_z0=me.src_lo,
                                   src_shape=plan.src_shape, dst_z0=me.ext_lo, dst_nz_local=me.ext_hi - me.ext_lo)
        self.z0, self.global_nz = me.ext_lo, nz
        self.owned_voxels = (me.out_hi - me.out_lo) * ny * nx
        return res, res[me.out_lo - me.ext_lo:me.out_hi - me.ext_lo]

    def slab_normalize(self, res):
        """Apply the (globally agreed) thresholds to the local planes, halo included."""
        with self.timer('normalize_apply'):
            self.normalized = self.stats.apply(res, res)
        self.norm_status = None

    def resample_and_normalize(self, own_src, header=None, defer_status=False):
        res, owned = self.slab_resample(own_src, header)
        nz, ny, nx = self.plan.out_shape
        with self.timer('order_stats'):
            if self.hist_exchange == 'peer' and self.world > 1:
                self.stats = ops.OrderStats(self.device).run(owned, n_total=nz * ny * nx, peer=self._peer_group())
            else:
                self.stats = ops.OrderStats(self.device).run(owned, n_total=nz * ny * nx,
                                                             all_reduce=self._all_reduce_hist)
        self.slab_normalize(res)
        return True if defer_status else self.check_status()

    def _global_shape(self):
        return tuple(self.plan.out_shape)

    def encode_af3(self, coords, bb_ch, aa_ch, defer_status=False):
        if self.af3_mode == 'sparse':            # atoms are binned on the global cube grid
            return super().encode_af3(coords, bb_ch, aa_ch, defer_status)
        nz, ny, nx = self.plan.out_shape
        with self.timer('af3_encode'):
            vol, status = ops.af3_encode(coords, bb_ch, aa_ch, self.header.origin, (nz, ny, nx),
                                         clip_hi_xyz=self.af3_clip, z0=self.z0, nz_local=self.normalized.shape[0])
        self.af3, self._af3_status, self._atoms_binned = vol, status, False
        if defer_status:
            return True
        ok = int(status.item()) == 0
        self.af3 = vol if ok else None
        return ok

    def cube_index(self):
        perm, offset = self.header.transpose_order()
        if perm[2] != 0:
            raise MicaError('slab partition needs memory axis 0 to be cube axis k (standard MRC axis order)')
        self.perm, self.offset = perm, offset
        me = self.plan.ranks[self.rank]
        self.cube_shape = ops.cube_space_shape(self.plan.out_shape, perm)
        ijk = ops.cube_origins(self.cube_shape, self.grid_size)
        mine = (ijk[:, 2] >= me.out_lo) & (ijk[:, 2] < me.out_hi)
        self._set_cube_origins(ijk[mine], (self.cube_shape, self.grid_size, me.out_lo, me.out_hi))
        self.box = ((0, 0, me.out_lo), (self.cube_shape[0], self.cube_shape[1], me.out_hi - me.out_lo))
        return self.ijk_host

    def _extract_map(self, ijk, x):
        ops.extract_cubes(self.normalized, ijk, self.grid_size, self.padding, self.perm, out=x,
                          global_nz=self.global_nz, z0=self.z0)

    def _extract_af3(self, ijk, af, nonzero):
        ops.extract_cubes(self.af3, ijk, self.grid_size, self.padding, self.perm, out=af,
                          global_nz=self.global_nz, z0=self.z0, nonzero=nonzero)

    def _new_volumes(self):
        org, ext = self.box
        return ops.StitchedVolumes(self.cube_shape, self.device, org=org, ext=ext)
