"""Shared recipes of the SURVEY 8(f) tests: the seeded inputs the golden fixtures were generated from
(oracle/make_golden.py::CANDIDATE_CASES / _mask_case) -- kept in one place so that the CPU tests, the GPU
tests and the generator cannot drift apart."""
import numpy as np

from mica_b200 import synthetic

CANDIDATE_CASES = [dict(shape=(56, 44, 40), n_residues=(60, 25), seed=1, wall_margin=3.0),
                   dict(shape=(40, 64, 48), n_residues=(30, 30, 20), seed=2, wall_margin=3.0),
                   dict(shape=(70, 30, 34), n_residues=(90,), seed=3, wall_margin=0.0)]

DOCK_CASES = [((1.0, 1.0, 1.0), 2.0), ((1.06, 0.93, 1.2), 3.3), ((0.83, 0.83, 0.83), 2.0), ((0.5, 0.5, 0.5), 2.0)]


def candidate_volumes(case):
    return synthetic.synthetic_predictions(case['shape'], case['n_residues'], seed=case['seed'],
                                           wall_margin=case['wall_margin'])


def mask_case(n_residues=160):
    shape = (40, 36, 32)
    origin = (np.float32(-2.5), np.float32(3.25), np.float32(1.0))
    st = synthetic.synthetic_structure(n_residues, (32, 36, 40), seed=5, origin_xyz=origin, hetero_every=7,
                                       unknown_every=11)
    return shape, origin, st


def dock_structure(vox, origin):
    box = (31 * vox[0], 35 * vox[1], 31 * vox[2])
    return synthetic.synthetic_structure(120, box, seed=5, origin_xyz=origin, hetero_every=7)
