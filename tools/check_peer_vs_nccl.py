#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/check_peer_vs_nccl.py : the slab pipeline's thresholds and owned
planes must be identical with the peer-memory histogram exchange and with the NCCL all-reduce."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from mica_b200 import synthetic
from mica_b200.pipeline import MapHeader
from mica_b200.slab import SlabPipeline

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
sys.stdout.flush()
saved = os.dup(1)
os.dup2(2, 1)
dist.init_process_group('nccl', device_id=dev)
src = torch.from_numpy(synthetic.synthetic_map((96, 80, 72), voxel=1.2, seed=5 + rank)).to(dev)
hdr = MapHeader(voxel_size=(np.float32(1.2),) * 3)
out = {}
for mode in ('nccl', 'peer', 'peer'):
    p = SlabPipeline(dev, rank, world, 32, 16, batch_cubes=8, hist_exchange=mode)
    assert p.resample_and_normalize(src, hdr)
    out.setdefault(mode, []).append((p.median, p.p999, p.n_pos, p.normalized.clone()))
torch.cuda.synchronize()
a, b, c = out['nccl'][0], out['peer'][0], out['peer'][1]
ok = a[:3] == b[:3] == c[:3] and torch.equal(a[3], b[3]) and torch.equal(a[3], c[3])
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
os.dup2(saved, 1)
if rank == 0:
    print('peer == nccl on every rank:', bool(flag.item()), 'median', a[0], 'p999', a[1], 'n_pos', a[2])
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
