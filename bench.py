#!/usr/bin/env python3
"""bench.py -- GVox/s of the map -> cubes -> stitched-volume hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the whole hot path over one synthetic map: B-spline resample to
the 1 A grid, exact median/p99.9 normalisation, 24-channel AF3 rasterisation, 64^3 cube
extraction into the model's input batch, and softmax/argmax + stitching of the model's
logits into the four output volumes.  The model itself (models/model.py, PyTorch
convolutions) is out of scope per north_star and is replaced by a ring of pre-generated
logits larger than L2, so every byte the post-processing reads comes from HBM.

N = 1 workload: BASELINE.json configs[1] -- synthetic 400^3 map at 1.2 A -> 480^3 working
grid, 64^3 cubes at stride 32 (grid_size=32, padding=16), ~158 k atoms.
N > 1: the same per-GPU slab (weak scaling): a (400 N) x 400 x 400 map z-slab partitioned
over N ranks, source-halo exchange and histogram all-reduce over NCCL.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (the reference's
NumPy/SciPy/torch-CPU arithmetic, in memory) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'map_to_stitched_volume_throughput'
UNIT = 'GVox/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--src-edge', type=int, default=400)
    ap.add_argument('--voxel', type=float, default=1.2)
    ap.add_argument('--grid-size', type=int, default=32)
    ap.add_argument('--padding', type=int, default=16)
    ap.add_argument('--batch-cubes', type=int, default=256)
    ap.add_argument('--cpu-edge', type=int, default=0, help='source edge of the CPU sample (0 = auto)')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--af3-mode', default='sparse', choices=['sparse', 'dense'],
                    help='sparse: AF3 cube channels written from per-cube atom bins (default); dense: the '
                         "reference's dataflow (24-channel volume, then window extraction)")
    ap.add_argument('--no-variant', action='store_true', help='skip timing the other af3 mode')
    ap.add_argument('--maps-in-flight', type=int, default=0,
                    help='also time K steps with this many maps in flight (one pipeline + stream each); 0 = skip')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def algorithmic_bytes(n_src, n_vox, n_atoms, f):
    """B(N) = 4 Ns + N (420 + 100 f) (+16 per atom)."""
    return {
        'resample': 4 * n_src + 4 * n_vox,
        'normalize': 12 * n_vox,
        'af3_encode': 96 * n_vox + 16 * n_atoms,
        'extract': 100 * n_vox * (1 + f),
        'postproc_stitch': 208 * n_vox,
    }


# ----------------------------------------------------------------------------- CPU arm
def cpu_sample_edge(args):
    if args.cpu_edge:
        return args.cpu_edge
    # the oracle costs ~1.75 s per 100^3 source map on one core and scales with the volume: size the
    # sample so that all runs together stay near 150 s (one run of ~14-24 s for the cpu_baseline leg)
    runs = args.steps + args.warmup if args.impl == 'reference' else 1
    edge = 100.0 * (150.0 / runs / 1.75) ** (1.0 / 3.0)
    return int(max(60, min(240 if runs > 1 else 200, edge // 20 * 20)))


def cpu_workload(edge, args):
    from mica_b200 import synthetic
    from oracle import mica_oracle as orc
    src = synthetic.synthetic_map((edge,) * 3, voxel=args.voxel, seed=2022)
    voxel = (np.float32(args.voxel),) * 3
    n_out = orc.zoom_output_shape(src.shape, orc.zoom_factors(voxel))
    st = synthetic.synthetic_structure(max(50, int(np.prod(n_out)) // 5500), n_out[::-1], seed=2022)
    bb_ch, aa_ch = orc.channel_codes(st['atom_names'], st['res_names'])
    n_cubes = int(np.prod([-(-n // args.grid_size) for n in n_out]))
    # 16 cubes of stand-in logits, reused for every chunk (the GPU arm reuses its ring the same way)
    ring = synthetic.synthetic_logits(16, args.grid_size + 2 * args.padding, seed=2022)
    return dict(src=src, voxel=voxel, coords=st['coords'], bb_ch=bb_ch, aa_ch=aa_ch, ring=ring,
                n_out=n_out, n_cubes=n_cubes)


def cpu_step(w, args):
    """The reference's arithmetic for the whole path, in memory (no .mrc/.npz I/O)."""
    from oracle import mica_oracle as orc
    t0 = time.perf_counter()
    vols, nvox, ncubes = orc.pipeline_whole_streamed(
        w['src'], w['voxel'], w['coords'], w['bb_ch'], w['aa_ch'], (0.0, 0.0, 0.0), w['ring'],
        args.grid_size, args.padding)
    dt = time.perf_counter() - t0
    assert ncubes == w['n_cubes']
    return dt, nvox


def cpu_baseline(args, steps=1, warmup=0):
    import torch
    edge = cpu_sample_edge(args)
    w = cpu_workload(edge, args)
    for _ in range(warmup):
        cpu_step(w, args)
    times = []
    for _ in range(steps):
        dt, nvox = cpu_step(w, args)
        times.append(dt)
    t = float(np.mean(times))
    return {
        'value': nvox / t / 1e9, 'unit': UNIT, 'cores': int(torch.get_num_threads()), 'kind': 'port',
        'host_cpus': os.cpu_count(),
        'sample': f'{edge}^3 map @ {args.voxel} A -> {"x".join(map(str, w["n_out"]))} grid, {w["n_cubes"]} cubes '
                  f'(grid_size={args.grid_size}, padding={args.padding}), {len(w["coords"])} atoms; oracle '
                  f'(scipy zoom + numpy + torch-CPU softmax) in memory, no file I/O; {t:.2f} s/step',
    }, t


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cb, t = cpu_baseline(args, steps=args.steps, warmup=args.warmup)
    f = ((args.grid_size + 2 * args.padding) / args.grid_size) ** 3
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, 1) | {'sampled': cb['sample']},
        'cpu_baseline': cb,
        'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def workload_config(args, n_gpus):
    e = args.src_edge
    return {
        'workload': f'BASELINE configs[1]: synthetic {e}^3 map @ {args.voxel} A -> 1 A grid, 64^3 cubes at stride '
                    f'{args.grid_size} (grid_size={args.grid_size}, padding={args.padding})'
                    + (f'; weak scaling: ({e}x{n_gpus})x{e}x{e} map z-slab partitioned over {n_gpus} GPUs'
                       if n_gpus > 1 else ''),
        'resample': 'cubic B-spline (scipy.ndimage.zoom order=3 semantics)',
        'model': 'excluded (north_star): logits come from a pre-generated HBM ring',
        'batch_cubes': args.batch_cubes,
        'l2': 'no flush needed: every stage streams buffers far larger than the 126 MB L2',
    }


class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--id={index}', f'--query-gpu={self.FIELDS}',
                                       '--format=csv,noheader,nounits', '-lms', '20'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(',') for r in open(self.f.name).read().strip().splitlines() if r.count(',') >= 6]
        os.unlink(self.f.name)
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        sm = [float(r[0]) for r in rows]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].strip().lower() == 'active' for r in rows)]
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(rows[0][1]), 'samples': len(rows),
                'power_w_max': max(float(r[2]) for r in rows), 'reasons': reasons}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mica_b200 import ops, synthetic
    from mica_b200.pdb import channel_codes
    from mica_b200.pipeline import MapHeader, MapPipeline, StageTimer, _no_timer, run_map_pipeline_host

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torch.distributed.run for --gpus > 1')
    ops.require_gpu()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout; the contract is ONE JSON line there
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=dev)

    # ---------------- synthetic inputs (host), seed 2022
    e = args.src_edge
    src_np = synthetic.synthetic_map((e, e, e), voxel=args.voxel, seed=2022 + rank)
    header = MapHeader(voxel_size=(np.float32(args.voxel),) * 3)
    n_out = ops.zoom_output_shape(src_np.shape, [np.float32(args.voxel)] * 3)
    # atoms are replicated on every rank (a few MB) and span the whole (stacked) working grid: one
    # 20 k-residue chain per slab, so that every rank has the same AF3 work (weak scaling)
    parts = [synthetic.synthetic_structure(20000, (n_out[2], n_out[1], n_out[0]), seed=2022 + r) for r in range(world)]
    for r, part in enumerate(parts):
        part['coords'][:, 2] += np.float32(r * n_out[0])
    st = {k: (np.concatenate([p_[k] for p_ in parts]) if isinstance(parts[0][k], np.ndarray)
              else sum((list(p_[k]) for p_ in parts), [])) for k in parts[0]}
    bb_ch, aa_ch = channel_codes(st['atom_names'], st['res_names'])
    src_host = torch.from_numpy(src_np).pin_memory()
    atoms_host = (torch.from_numpy(st['coords']).pin_memory(), torch.from_numpy(bb_ch).pin_memory(),
                  torch.from_numpy(aa_ch).pin_memory())
    src = src_host.to(dev)
    atoms = tuple(t.to(dev) for t in atoms_host)

    def make_pipe(af3_mode):
        if world > 1:
            from mica_b200.slab import SlabPipeline
            p = SlabPipeline(dev, rank, world, grid_size=args.grid_size, padding=args.padding,
                             batch_cubes=args.batch_cubes, af3_mode=af3_mode)
            # the stacked map is not cubic: use the geometrically meant clip bounds instead of the
            # reference's (z,y,x)-vs-(x,y,z) mix-up (D7), which would squash every atom onto z <= nx-1
            p.af3_clip = (n_out[2] - 1, n_out[1] - 1, n_out[0] * world - 1)
            return p
        return MapPipeline(dev, grid_size=args.grid_size, padding=args.padding, batch_cubes=args.batch_cubes,
                           af3_mode=af3_mode)

    # ---------------- logits ring (stands where MICA.forward stands), >> L2
    W, B = args.grid_size + 2 * args.padding, args.batch_cubes
    gen = torch.Generator(device=dev).manual_seed(2022)
    ring = [tuple(torch.randn((B, c, W, W, W), generator=gen, device=dev) * 2 for c in (4, 4, 21)) for _ in range(2)]
    state = {'i': 0}

    def model_fn(x, af):
        bb, ca, aa = ring[state['i'] % len(ring)]
        state['i'] += 1
        b = x.shape[0]
        return bb[:b], ca[:b], aa[:b]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(pipe, with_clocks, only=None):
        """W warm-up steps, then exactly K timed steps between barrier+synchronize, CUDA events on the
        launching stream, max over ranks.  ``only``: the stages whose launches are bracketed by events
        inside the timed region (every event pair costs ~4 us of stream time; 81 launches per step)."""
        vols = None
        for _ in range(args.warmup):
            vols = pipe.run(src, header, atoms, model_fn, vols)      # checked at once: a broken setup fails here
        sync()
        timer = StageTimer(only)
        pipe.timer = timer
        if os.environ.get('MICA_NO_PREFETCH'):            # experiment knob
            pipe.prefetch = False
        launches0 = ops.launch_count()
        clocks = ClockSampler(local_rank) if (with_clocks and rank == 0) else None
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        ev0.record()
        t_host = time.perf_counter()
        for _ in range(args.steps):
            # status words go to pinned memory in stream order and are checked after the loop (finish()):
            # no host synchronisation between the maps of the stream
            vols = pipe.run(src, header, atoms, model_fn, vols, defer_check=True)
        ev1.record()
        t_host = (time.perf_counter() - t_host) / args.steps * 1e3      # host time spent enqueueing one step
        sync()
        pipe.finish()
        ms = ev0.elapsed_time(ev1)
        launches = ops.launch_count() - launches0
        clk = clocks.stop() if clocks else None
        stages = timer.summary()
        pipe.timer = _no_timer
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / args.steps, stages, launches, clk, vols, t_host

    pipe = make_pipe(args.af3_mode)
    # pass 1 (not the headline): every stage bracketed by events -> the per-stage breakdown and which
    # single-kernel stage dominates.  pass 2 (the headline): the K timed steps with events around the
    # dominant kernel's launches only, as the roofline line needs its live launch duration.
    _, stages_all, _, _, vols, _ = measure(pipe, False)
    single_kernel_stages = ('postproc_stitch', 'extract_af3', 'extract_map', 'normalize_apply')
    dom_stage = max((k for k in single_kernel_stages if k in stages_all), key=lambda k: stages_all[k][1])
    del vols
    ms_per_step, stages, launches, clk, vols, host_ms = measure(pipe, True, only={dom_stage})
    n_vox_rank = int(np.prod(pipe.normalized.shape)) if world == 1 else pipe.owned_voxels
    n_vox = n_vox_rank * world
    value = n_vox / (ms_per_step * 1e-3) / 1e9
    n_cubes = len(pipe.ijk_host)

    variant = None
    if not args.no_variant and world == 1:
        other = 'dense' if args.af3_mode == 'sparse' else 'sparse'
        del vols
        vols = None
        pipe2 = make_pipe(other)
        ms2, stages2, _, _, vols2, _ = measure(pipe2, False)
        variant = {'af3_mode': other, 'ms_per_step': ms2, 'value': n_vox / (ms2 * 1e-3) / 1e9, 'unit': UNIT,
                   'stage_ms_per_step': {k: round(v[1] / args.steps, 4) for k, v in stages2.items()}}
        del pipe2, vols2
        torch.cuda.empty_cache()
        vols = pipe.run(src, header, atoms, model_fn, None)

    # ---------------- several maps in flight (a stream of maps: the resample / order-statistics kernels of one
    # map are issue-bound, the cube loop of another is DRAM-bound, so they overlap).  Reported as a variant:
    # the headline above stays one map at a time.
    in_flight = None
    if args.maps_in_flight > 1 and world == 1:
        n_f = args.maps_in_flight
        pipes = [pipe] + [make_pipe(args.af3_mode) for _ in range(n_f - 1)]
        streams = [torch.cuda.Stream(dev) for _ in range(n_f)]
        fv = [vols] + [None] * (n_f - 1)
        for p_ in pipes:
            p_.timer = _no_timer

        def run_many(k):
            start = torch.cuda.Event(enable_timing=True)
            start.record()
            for st_ in streams:
                st_.wait_event(start)
            for s_ in range(k):
                i_ = s_ % n_f
                with torch.cuda.stream(streams[i_]):
                    fv[i_] = pipes[i_].run(src, header, atoms, model_fn, fv[i_], defer_check=True)
            for st_ in streams:
                torch.cuda.current_stream().wait_stream(st_)
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            return start, end

        run_many(max(args.warmup, args.steps))      # also fills every pipeline's pool of pinned status records
        sync()
        for p_ in pipes:
            p_.finish()
        a_, b_ = run_many(args.steps)
        sync()
        for p_ in pipes:
            p_.finish()
        ms_f = a_.elapsed_time(b_) / args.steps
        in_flight = {'maps_in_flight': n_f, 'ms_per_step': ms_f, 'value': n_vox / (ms_f * 1e-3) / 1e9, 'unit': UNIT,
                     'note': 'K steps issued round-robin on %d independent pipelines/streams; throughput of a stream of '
                             'maps, not the latency of one' % n_f}
        vols = fv[0]
        del pipes, fv
        torch.cuda.empty_cache()

    # ---------------- end to end through the public API with host buffers
    # (every rank copies its own source block in and its own slab of the four volumes out)
    out_host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in vols.as_dict().items()}
    dev_vols = vols                                                 # device volumes reused by every e2e step
    run_map_pipeline_host(src_host, header, atoms_host, model_fn, pipe, out_host, dev_vols)       # warm
    sync()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        _, h2d, d2h = run_map_pipeline_host(src_host, header, atoms_host, model_fn, pipe, out_host, dev_vols)
    sync()
    dt = (time.perf_counter() - t0) / args.e2e_steps
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {'value': n_vox / dt / 1e9, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d) * world,
           'd2h_bytes_per_step': int(d2h) * world, 'ms_per_step': dt * 1e3,
           'note': 'pinned host map + atoms -> device -> four stitched volumes -> pinned host (finished x-layers are '
                   'copied out on side streams while later cube batches run); '
                   'host wall clock between device synchronisations, max over ranks'}

    # the same call with amino_acid_probability (20 of the 23 channels) left in HBM: what the drop-in for the
    # head of Solver.clustering (mica_b200/candidates.py) makes possible -- its only consumer in
    # the reference (utils/modeler.py:850) then runs on the device.  Reported next to e2e, never instead of it.
    e2e3 = None
    if world == 1:
        out3 = {k: v for k, v in out_host.items() if k != 'amino_acid_probability'}
        run_map_pipeline_host(src_host, header, atoms_host, model_fn, pipe, out3, dev_vols)
        sync()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            _, h2d3, d2h3 = run_map_pipeline_host(src_host, header, atoms_host, model_fn, pipe, out3, dev_vols)
        sync()
        dt3 = (time.perf_counter() - t0) / args.e2e_steps
        e2e3 = {'value': n_vox / dt3 / 1e9, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d3), 'd2h_bytes_per_step': int(d2h3),
                'ms_per_step': dt3 * 1e3,
                'note': 'as e2e, but amino_acid_probability stays on the device for mica_b200.candidates '
                        '(utils/modeler.py:767-860 on the GPU); backbone / C-alpha / amino_acid_prediction go to the host'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (the single-kernel stage with the most time)
    peak, peak_src = peaks()
    S = args.grid_size
    f = (W / S) ** 3
    n_src = src.numel()
    n_atoms = atoms[0].shape[0]
    alg = algorithmic_bytes(n_src, n_vox_rank, n_atoms, f)
    stage_ms = {k: round(v[1] / args.steps, 4) for k, v in stages_all.items()}
    # stage -> (kernel, SURVEY 8(d) algorithmic bytes per step of that kernel)
    single = {
        'postproc_stitch': ('postproc_stitch_kernel', alg['postproc_stitch'], '208 B/voxel: 29 logit channels of '
                            'each core voxel read (116 B), 23 output channels written (92 B)'),
        'extract_af3': ('extract_tma_kernel<4> (24 AF3 channels)', 96 * n_vox_rank * (1 + f),
                        '96 B/voxel x (1 + f): 24 channels read once, written f = (W/S)^3 times'),
        'extract_map': ('extract_tma_kernel<4> (map channel)', 4 * n_vox_rank * (1 + f),
                        '4 B/voxel x (1 + f): map read once, written f = (W/S)^3 times'),
        'normalize_apply': ('normalize_apply_kernel', 8 * n_vox_rank, '8 B/voxel: read + write in place'),
    }
    dom = dom_stage
    calls, tot_ms = stages[dom]
    per_launch_bytes = single[dom][1] / (calls / args.steps)
    dur_ms = tot_ms / calls
    achieved = per_launch_bytes / (dur_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):      # dram bytes per launch from the committed `ncu --set full` capture
        t = json.load(open(tpath)).get(single[dom][0].split(' ')[0])
        if t and t.get('batch_cubes') == args.batch_cubes and t.get('grid_size') == S:
            traffic = t['dram_bytes_per_launch']
    stage_frac = {}
    for name, keys in (('resample', ['resample']), ('normalize', ['order_stats', 'normalize_apply']),
                       ('af3_encode', ['af3_encode']), ('extract', ['extract_map', 'extract_af3', 'af3_fill_cubes']),
                       ('postproc_stitch', ['postproc_stitch'])):
        t_ms = sum(stage_ms.get(k, 0.0) for k in keys)
        if t_ms:
            stage_frac[name] = round(alg[name] / (t_ms * 1e-3) / 1e9 / peak, 4)
    total_alg = sum(alg.values())
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic (seed 2022; random-init stand-in logits)',
        'config': workload_config(args, world) | {'working_grid': list(pipe.normalized.shape), 'cubes': n_cubes,
                                                   'atoms': int(n_atoms)},
        'roofline': {
            'bound': 'hbm', 'kernel': single[dom][0], 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
            'frac': achieved / peak, 'traffic': traffic,
            'peak_source': peak_src, 'launch_ms': dur_ms, 'launches_per_step': calls / args.steps,
            'algorithmic_bytes_per_launch': per_launch_bytes, 'algorithmic_bytes_rule': single[dom][2],
            'share_of_step': tot_ms / args.steps / ms_per_step,
            # the whole path against SURVEY 8(d)'s B(N), which counts the dataflow the reference
            # materialises (dense 24-channel AF3 volume and its windows); the sparse AF3 path moves
            # fewer bytes than B(N), so this figure can exceed 1 -- it is not a kernel roofline
            'whole_path': {'algorithmic_bytes_per_voxel': total_alg / n_vox_rank,
                           'achieved': total_alg / (ms_per_step * 1e-3) / 1e9,
                           'frac': total_alg / (ms_per_step * 1e-3) / 1e9 / peak},
            'stage_ms_per_step': stage_ms, 'stage_frac_of_peak': stage_frac,
            'stage_note': 'stage times come from a separate fully instrumented pass of the same K steps; with the '
                          'prefetch stream extract / fill overlap the stitch, so they do not add up to ms_per_step',
        },
        'af3_mode': args.af3_mode, 'variant': variant,
        'clocks': clk, 'gpu_launches': int(launches), 'host_enqueue_ms_per_step': host_ms, 'host_loop_enqueue_ms': getattr(pipe, 'last_loop_enqueue_ms', None), 'e2e': e2e,
        'e2e_aa_prob_resident': e2e3, 'variant_maps_in_flight': in_flight,
    }
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'], _ = cpu_baseline(args)
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
