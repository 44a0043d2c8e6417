#!/usr/bin/env python3
"""Experiment: does cudaLimitMaxL2FetchGranularity change postproc_stitch's time (over-fetch of
the 128-byte core rows that start 64 bytes into a 256-byte logit row)?"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mica_b200 import ops

rt = None
for name in ('libcudart.so.12', 'libcudart.so'):
    try:
        rt = ctypes.CDLL(name); break
    except OSError:
        pass
dev = torch.device('cuda:0')
torch.zeros(1, device=dev)
B, W, S, pad = 32, 64, 32, 16
ring = [tuple(torch.randn((B, c, W, W, W), device=dev) for c in (4, 4, 21)) for _ in range(3)]
shape = (128, 128, 128 * 2)
ijk = torch.from_numpy(ops.cube_origins(shape, S)[:B].copy()).to(dev)
vols = ops.StitchedVolumes(shape, dev)


def timeit():
    for r in ring:
        ops.postproc_stitch(*r, ijk, vols, S, pad)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 30
    for i in range(n):
        ops.postproc_stitch(*ring[i % 3], ijk, vols, S, pad)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


val = ctypes.c_size_t()
if rt is not None:
    rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
print('default L2 fetch granularity', val.value, 'stitch us', timeit())
for g in (32, 64, 128):
    if rt is not None:
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(g))
        rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
    print('set', g, 'rc', rc, 'now', val.value, 'stitch us', timeit())
