#!/usr/bin/env python3
"""torchrun experiment: where do the ranks of the slab pipeline wait for each other?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from mica_b200 import ops, synthetic
from mica_b200.pdb import channel_codes
from mica_b200.pipeline import MapHeader, StageTimer
from mica_b200.slab import SlabPipeline

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev = torch.device('cuda', local)
os.dup2(2, 1)
dist.init_process_group('nccl', device_id=dev)
e = 400
src = torch.from_numpy(synthetic.synthetic_map((e, e, e), voxel=1.2, seed=2022 + rank)).to(dev)
header = MapHeader(voxel_size=(np.float32(1.2),) * 3)
n_out = ops.zoom_output_shape(src.shape, [np.float32(1.2)] * 3)
st = synthetic.synthetic_structure(20000 * world, (n_out[2], n_out[1], n_out[0] * world), seed=2022)
bb_ch, aa_ch = channel_codes(st['atom_names'], st['res_names'])
atoms = tuple(torch.from_numpy(a).to(dev) for a in (st['coords'], bb_ch, aa_ch))
pipe = SlabPipeline(dev, rank, world, 32, 16, batch_cubes=128)
pipe.af3_clip = (n_out[2] - 1, n_out[1] - 1, n_out[0] * world - 1)
B, W = 128, 64
ring = tuple(torch.randn((B, c, W, W, W), device=dev) for c in (4, 4, 21))
model_fn = lambda x, af: tuple(t[:x.shape[0]] for t in ring)
vols = None
for _ in range(3):
    vols = pipe.run(src, header, atoms, model_fn, vols)
dist.barrier(); torch.cuda.synchronize()
rows = []
for it in range(8):
    timer = StageTimer()
    pipe.timer = timer
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.time()
    ev0.record()
    vols = pipe.run(src, header, atoms, model_fn, vols)
    ev1.record()
    t_host1 = time.time()
    torch.cuda.synchronize()
    s = {k: v[1] for k, v in timer.summary().items()}
    # offsets of the stage starts relative to the step start
    first = {name: ev0.elapsed_time(a) for name, a, b in timer.spans}
    rows.append((t_host0, (t_host1 - t_host0) * 1e3, ev0.elapsed_time(ev1), s, first))
for r in range(world):
    dist.barrier()
    if r == rank:
        for t0, host_ms, gpu_ms, s, first in rows:
            print(f'rank {rank} start {t0 % 100:9.5f} s host {host_ms:6.2f} ms gpu {gpu_ms:6.2f} ms  ' +
                  ' '.join(f'{k}={v:.2f}' for k, v in s.items()), file=sys.stderr)
        sys.stderr.flush()
dist.destroy_process_group()
