import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mica_b200 import ops
shape = tuple(int(v) for v in sys.argv[1:4])
voxel = [np.float32(v) for v in sys.argv[4:7]]
src = np.random.default_rng(0).normal(size=shape).astype(np.float32)
out_shape = ops.zoom_output_shape(shape, voxel)
print(shape, out_shape, flush=True)
r = ops.resample(torch.from_numpy(src).cuda(), out_shape)
torch.cuda.synchronize()
print('ok', float(r.abs().max()))
