#!/usr/bin/env python3
"""Times the pre-phase stages in isolation (CUDA events, median of R runs) -- the loop used while
working on a kernel:  python tools/bench_stages.py [--edge 400 --voxel 1.2] [--stages resample,stats,extract]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mica_b200 import ops, synthetic  # noqa: E402


def timeit(fn, reps=9, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--edge', type=int, default=400)
    ap.add_argument('--voxel', type=float, default=1.2)
    ap.add_argument('--stages', default='resample,stats,normalize,extract,extract24,stitch')
    ap.add_argument('--grid-size', type=int, default=32)
    ap.add_argument('--padding', type=int, default=16)
    ap.add_argument('--batch', type=int, default=256)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    stages = set(args.stages.split(','))
    src = synthetic.synthetic_map_device((args.edge,) * 3, dev, voxel=args.voxel)
    zf = [np.float32(args.voxel)] * 3
    out_shape = ops.zoom_output_shape(src.shape, zf)
    n, ns = int(np.prod(out_shape)), src.numel()
    peak = 6554.6
    res = ops.resample(src, out_shape) if args.voxel != 1.0 else src.clone()
    print(f'source {args.edge}^3 -> grid {out_shape}, {n / 1e6:.1f} MVox')
    if 'resample' in stages and args.voxel != 1.0:
        out = torch.empty_like(res)
        med, best = timeit(lambda: ops.resample(src, out_shape, out=out))
        print(f'resample        {med:8.3f} ms (best {best:.3f})  {(4 * ns + 4 * n) / med / 1e6:8.1f} GB/s alg  '
              f'frac {(4 * ns + 4 * n) / med / 1e6 / peak:.3f}')
    if 'stats' in stages:
        st = ops.OrderStats(dev)
        med, best = timeit(lambda: st.run(res))
        info = st.compact_info()
        print(f'order_stats     {med:8.3f} ms (best {best:.3f})  {4 * n / med / 1e6:8.1f} GB/s alg  compact {info}')
        if 'normalize' in stages:
            y = torch.empty_like(res)
            med, best = timeit(lambda: st.apply(res, y))
            print(f'normalize_apply {med:8.3f} ms (best {best:.3f})  {8 * n / med / 1e6:8.1f} GB/s  frac {8 * n / med / 1e6 / peak:.3f}')
    if 'writepeak' in stages:
        buf = torch.empty(1 << 31, dtype=torch.uint8, device=dev)          # 2 GiB
        other = torch.empty_like(buf)
        med, best = timeit(lambda: buf.zero_())
        print(f'memset 2 GiB    {med:8.3f} ms (best {best:.3f})  {buf.numel() / med / 1e6:8.1f} GB/s written')
        med, best = timeit(lambda: other.copy_(buf))
        print(f'copy 2 GiB      {med:8.3f} ms (best {best:.3f})  {2 * buf.numel() / med / 1e6:8.1f} GB/s read+write')
        del buf, other
    gs, pad, B = args.grid_size, args.padding, args.batch
    W = gs + 2 * pad
    cs = ops.cube_space_shape(res.shape)
    ijk_all = ops.cube_origins(cs, gs)
    ijk = torch.from_numpy(ijk_all[:B]).to(dev)
    if 'extract' in stages:
        x = torch.empty((len(ijk), 1, W, W, W), device=dev)
        med, best = timeit(lambda: ops.extract_cubes(res, ijk, gs, pad, out=x))
        nb = len(ijk) * 4 * (gs ** 3 + W ** 3)
        print(f'extract map x{len(ijk)} {med:8.3f} ms (best {best:.3f})  {nb / med / 1e6:8.1f} GB/s alg  frac {nb / med / 1e6 / peak:.3f}')
    if 'extract24' in stages:
        Bc = min(B, int(os.environ.get('MICA_BENCH_E24', '64')))
        vol = torch.rand((24,) + tuple(res.shape), device=dev)
        x = torch.empty((Bc, 24, W, W, W), device=dev)
        med, best = timeit(lambda: ops.extract_cubes(vol, ijk[:Bc], gs, pad, out=x))
        nb = Bc * 24 * 4 * (gs ** 3 + W ** 3)
        print(f'extract 24ch x{Bc} {med:8.3f} ms (best {best:.3f})  {nb / med / 1e6:8.1f} GB/s alg  frac {nb / med / 1e6 / peak:.3f}')
        del vol, x
    if 'stitch' in stages:
        g = torch.Generator(device=dev).manual_seed(1)
        bb, ca, aa = (torch.randn((len(ijk), c, W, W, W), generator=g, device=dev) for c in (4, 4, 21))
        vols = ops.StitchedVolumes(cs, dev)
        med, best = timeit(lambda: ops.postproc_stitch(bb, ca, aa, ijk, vols, gs, pad))
        nb = len(ijk) * gs ** 3 * 208
        print(f'postproc_stitch x{len(ijk)} {med:8.3f} ms (best {best:.3f})  {nb / med / 1e6:8.1f} GB/s alg  frac {nb / med / 1e6 / peak:.3f}')


if __name__ == '__main__':
    main()
