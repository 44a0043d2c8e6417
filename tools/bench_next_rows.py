#!/usr/bin/env python3
"""Times the SURVEY 8(f) rows at BASELINE sizes on one B200 (CUDA events, warm, median of 5):

  N1  C-alpha candidates on 480^3 stitched volumes of a 20 000-residue synthetic structure
  N3  label masks for the same structure in a 480^3 grid (configs[2] geometry)
  N4  contour threshold + 2 A masking around 40 % of the atoms on a 480^3 map

Prints one JSON object; `python tools/bench_next_rows.py > profiles/rNN_next_rows.json` on the GPU box."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mica_b200 import candidates as cd, dock_masks as dm, label_masks as lm, ops, synthetic   # noqa: E402

EDGE = int(os.environ.get('EDGE', '480'))
N_RES = int(os.environ.get('N_RES', '20000'))


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


def device_predictions(st, dev):
    """Gaussian C-alpha peaks / backbone tube rendered on the device (7^3 stencil per atom)."""
    shape = (EDGE,) * 3
    names = np.array(st['atom_names'])
    ca_xyz = torch.from_numpy(st['coords'][names == 'CA']).to(dev)
    bb_xyz = torch.from_numpy(st['coords'][np.isin(names, ['N', 'CA', 'C', 'O'])]).to(dev)
    g = torch.Generator(device=dev).manual_seed(2022)

    def splat(xyz, sigma, amp_lo, amp_hi):
        vol = torch.zeros(EDGE ** 3, device=dev)
        r = torch.arange(-3, 4, device=dev)
        off = torch.stack(torch.meshgrid(r, r, r, indexing='ij'), -1).reshape(-1, 3)
        base = torch.round(xyz).long()
        idx = (base[:, None, :] + off[None]).clamp_(0, EDGE - 1)
        d2 = ((idx.float() - xyz[:, None, :]) ** 2).sum(-1)
        amp = amp_lo + (amp_hi - amp_lo) * torch.rand(xyz.shape[0], 1, generator=g, device=dev)
        val = amp * torch.exp(-d2 / (2 * sigma * sigma))
        lin = (idx[..., 0] * EDGE + idx[..., 1]) * EDGE + idx[..., 2]
        vol.scatter_reduce_(0, lin.reshape(-1), val.reshape(-1), 'amax')
        vol += 0.01 * torch.rand(EDGE ** 3, generator=g, device=dev)
        return vol.clamp_(0, 1).reshape(shape)

    ca = splat(ca_xyz, 0.9, 0.55, 0.99)
    bb = splat(bb_xyz, 1.2, 0.9, 0.99)
    aa = torch.rand((20,) + shape, generator=g, device=dev)
    aa /= aa.sum(0, keepdim=True)
    pred = aa.argmax(0).float()
    return dict(carbon_alpha_probability=ca, backbone_probability=bb, amino_acid_probability=aa,
                amino_acid_prediction=pred)


def main():
    ops.require_gpu()
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    st = synthetic.synthetic_structure(N_RES, (EDGE,) * 3, seed=2022)
    vols = device_predictions(st, dev)
    ca = vols['carbon_alpha_probability']
    n_vox = EDGE ** 3
    out = {'grid': [EDGE] * 3, 'residues': N_RES, 'atoms': int(len(st['coords'])), 'timing': 'CUDA events, median of 5 after 1 warm-up'}

    # ---- N1
    l0 = ops.launch_count()
    t_thr, (lin, xyz) = timed(lambda: cd.threshold_points(ca, 0.3))
    t_db, (labels, ncl) = timed(lambda: cd.dbscan_lattice(lin, (EDGE,) * 3, 10, 10))
    t_all, res = timed(lambda: cd.find_candidates(vols))
    t0 = time.perf_counter()
    res = cd.find_candidates(vols)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    graph = cd.neighbor_graph(res['CA_cands'], vols['backbone_probability'])
    graph_wall = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    graph = cd.neighbor_graph(res['CA_cands'], vols['backbone_probability'])
    graph_wall = min(graph_wall, (time.perf_counter() - t0) * 1e3)
    out['N1_neighbor_graph'] = {'picks': int(len(res['CA_cands'])), 'wall_ms_incl_two_dense_matrices_to_host': graph_wall,
                                'matrix_bytes_to_host': int(graph['cand_self_dis'].nbytes + graph['neigh_mat'].nbytes),
                                'scored_pairs': int((graph['neigh_mat'] > 0).sum())}
    del graph
    out['N1_candidates'] = {
        'points_above_threshold': int(lin.shape[0]), 'clusters': int(ncl), 'picks': int(len(res['picks'])),
        'candidates': int(len(res['CA_cands'])), 'nms_rounds': int(res['nms_rounds']),
        'threshold_ms': t_thr, 'threshold_GBps_of_4N_read': 4 * n_vox / (t_thr * 1e-3) / 1e9,
        'dbscan_ms': t_db, 'find_candidates_ms': t_all, 'find_candidates_wall_ms': wall,
        'd2h_bytes': int(res['CA_cands'].nbytes + res['CA_cands_AAProb'].nbytes + res['CA_cands_AA'].nbytes
                         + res['picks'].nbytes),
        'avoided_d2h_bytes': int(vols['amino_acid_probability'].numel() * 4),
        'launches': int(ops.launch_count() - l0)}

    # ---- N3
    coords = torch.from_numpy(st['coords']).to(dev)
    names = np.array(st['atom_names'])
    is_bb = torch.from_numpy(np.isin(names, ['N', 'CA', 'C', 'O']).astype(np.uint8)).to(dev)
    resn = np.array(st['res_names'])
    sel = names == 'CA'
    ca_xyz = torch.from_numpy(st['coords'][sel]).to(dev)
    labs = torch.from_numpy(np.array([lm.AA_MAPPING[r] for r in resn[sel]], dtype=np.int32)).to(dev)
    t_cls, _ = timed(lambda: lm.class_mask(coords, is_bb, (0, 0, 0), (EDGE,) * 3))
    t_aa, _ = timed(lambda: lm.aa_mask(ca_xyz, labs, (0, 0, 0), (EDGE,) * 3))
    out['N3_label_masks'] = {'class_mask_ms': t_cls, 'class_mask_GBps_of_4N_written': 4 * n_vox / (t_cls * 1e-3) / 1e9,
                             'aa_mask_ms': t_aa, 'aa_mask_note': '4 int32 volumes initialised (16 B/voxel) + 3 atom passes',
                             'aa_mask_GBps_of_16N_written': 16 * n_vox / (t_aa * 1e-3) / 1e9}

    # ---- N4
    m = torch.rand((EDGE,) * 3, device=dev)
    t_thr4, thr = timed(lambda: dm.contour_threshold(m, 0.1))
    sel_atoms = torch.from_numpy(dm.select_central_atoms(st['coords'])).to(dev)
    t_zero, _ = timed(lambda: dm.zero_around_atoms(thr, sel_atoms, (1.0, 1.0, 1.0), (0, 0, 0), 2.0))
    out['N4_docking_masks'] = {'contour_threshold_ms': t_thr4, 'contour_threshold_GBps_of_8N': 8 * n_vox / (t_thr4 * 1e-3) / 1e9,
                               'zero_around_atoms_ms': t_zero, 'selected_atoms': int(sel_atoms.shape[0])}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
