#!/usr/bin/env python3
"""torchrun experiment: cost of the NCCL source-halo exchange with the ranks aligned by a barrier."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from mica_b200.pipeline import MapHeader
from mica_b200.slab import SlabPipeline, exchange_source_halo

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev = torch.device('cuda', local)
os.dup2(2, 1)
dist.init_process_group('nccl', device_id=dev)
src = torch.randn((400, 400, 400), device=dev)
p = SlabPipeline(dev, rank, world, 32, 16)
plan = p.make_plan(tuple(src.shape), MapHeader(voxel_size=(np.float32(1.2),) * 3))
for i in range(5):
    exchange_source_halo(src, plan, rank)
torch.cuda.synchronize()
res = []
for i in range(10):
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    buf = exchange_source_halo(src, plan, rank)
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    res.append((a.elapsed_time(b), (t1 - t0) * 1e3))
print(f'rank {rank}: gpu ms {np.median([r[0] for r in res]):.3f} host ms {np.median([r[1] for r in res]):.3f} sends/recvs {plan.transfers(rank)}', file=sys.stderr)
# only the local copy part
torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); x = torch.empty_like(src); x.copy_(src); b.record(); torch.cuda.synchronize()
print(f'rank {rank}: plain 256 MB copy {a.elapsed_time(b):.3f} ms', file=sys.stderr)
dist.destroy_process_group()
