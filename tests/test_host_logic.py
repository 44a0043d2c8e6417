"""Host-side logic of round 2 that needs no GPU: model feeding (D8), the multi-GPU plans, the session
registry's lifetime rules, the array-like device volume, and the input edge cases of the readers."""
import gzip
import os
import time

import numpy as np
import pytest
import torch

from mica_b200 import mrc, pdb, session
from mica_b200.peer import PeerHalo
from mica_b200.pipeline import run_model_chunks
from mica_b200.predict import DeviceVolume
from mica_b200.slab import SlabPlan, _split_even


# ------------------------------------------------------------------ D8: how a super-batch is fed to the model
def _feed(flags, model_batch, d8):
    B = len(flags)
    x = torch.arange(B, dtype=torch.float32).view(B, 1, 1, 1, 1)
    af = torch.tensor(flags, dtype=torch.float32).view(B, 1, 1, 1, 1)
    ijk = torch.arange(3 * B, dtype=torch.int32).view(B, 3)
    calls, stitched = [], []

    def model(gx, gaf):
        calls.append((gx.flatten().tolist(), bool(gaf.abs().sum() < 1e-6)))
        return gx, gx, gx

    def stitch(bb, ca, aa, gijk):
        stitched.extend((int(v), int(i)) for v, i in zip(bb.flatten().tolist(), (gijk[:, 0] // 3).tolist()))

    run_model_chunks(model, x, af, np.asarray(flags), ijk, stitch, model_batch, d8)
    return calls, stitched


def test_split_mode_never_mixes_zero_and_nonzero_af3_cubes():
    flags = [1, 0, 0, 1, 1, 0, 0, 0, 1, 0]
    calls, stitched = _feed(flags, 4, 'split')
    for cubes, is_zero in calls:
        assert all((flags[int(c)] == 0) == is_zero for c in cubes)          # homogeneous, and the branch matches
        assert len(cubes) <= 4
    assert sorted(stitched) == [(b, b) for b in range(10)]                  # every cube once, with its own ijk
    # chunks follow the reference's consecutive batches of model_batch cubes
    assert sorted(sum((c for c, _ in calls[:2]), [])) == [0.0, 1.0, 2.0, 3.0]


def test_reference_mode_keeps_the_mixed_batches():
    flags = [1, 0, 0, 1, 1, 0, 0, 0, 1, 0]
    calls, stitched = _feed(flags, 4, 'reference')
    assert [c for c, _ in calls] == [[0.0, 1.0, 2.0, 3.0], [4.0, 5.0, 6.0, 7.0], [8.0, 9.0]]
    assert [z for _, z in calls] == [False, False, False]                    # a mixed batch is never "zero"
    assert sorted(stitched) == [(b, b) for b in range(10)]
    calls1, _ = _feed(flags, 1, 'split')                                     # batch 1: nothing to split
    assert [len(c) for c, _ in calls1] == [1] * 10


def test_whole_super_batch_when_no_model_batch():
    calls, _ = _feed([0, 1, 0], None, 'none')
    assert len(calls) == 1 and len(calls[0][0]) == 3
    with pytest.raises(Exception):
        _feed([0], 1, 'sometimes')


# ------------------------------------------------------------------ multi-GPU plans
@pytest.mark.parametrize('src,voxel,gs,pad', [((679, 8, 8), 1.06, 48, 8), ((720, 8, 8), 1.0, 48, 8),
                                             ((3200, 8, 8), 1.2, 32, 16), ((512, 8, 8), 1.0, 48, 8)])
@pytest.mark.parametrize('world', [2, 4, 8])
def test_bench_plans_need_only_the_direct_neighbours(src, voxel, gs, pad, world):
    """The peer-memory halo exchange serves plans whose halos come from rank +-1 only: true for every
    multi-GPU configuration of bench.py, because a rank holds the source planes under its output slab."""
    plan = SlabPlan(src, (np.float32(voxel),) * 3, gs, pad, world)
    for r in range(world):
        np_ = PeerHalo.neighbour_plan(plan, r)
        assert np_ is not None
        s_lo, s_hi, r_lo, r_hi = np_
        me = plan.ranks[r]
        if r_lo is not None:                             # what I receive from below is what rank r-1 sends up
            assert PeerHalo.neighbour_plan(plan, r - 1)[1] == r_lo and r_lo[1] == me.own_lo
        if r_hi is not None:
            assert PeerHalo.neighbour_plan(plan, r + 1)[0] == r_hi and r_hi[0] == me.own_hi
        if s_lo is not None:
            assert s_lo[0] == me.own_lo                   # a prefix / suffix of my own block
        if s_hi is not None:
            assert s_hi[1] == me.own_hi


@pytest.mark.parametrize('src,voxel,gs,pad', [((679, 8, 8), 1.06, 48, 8), ((400, 8, 8), 1.2, 32, 16),
                                             ((3200, 8, 8), 1.2, 32, 16), ((333, 8, 8), 0.9, 48, 8)])
@pytest.mark.parametrize('world', [2, 3, 8])
def test_slab_blocks_hold_whole_prefilter_windows(src, voxel, gs, pad, world):
    """Bit-identity of the N-rank map rests on this: the z prefilter cuts a line into segments of
    ceil(n / ceil(n / 26)) samples with a window of 16 before and 26 + 16 from the segment start (resample.cu,
    cols_reg_kernel); a rank's block must contain the window of every segment that holds a plane its taps read."""
    plan = SlabPlan(src, (np.float32(voxel),) * 3, gs, pad, world)
    sz, nz = src[0], plan.out_shape[0]
    n_seg = -(-sz // 26)
    seg = -(-sz // n_seg)
    scale = (sz - 1) / (nz - 1)
    for me in plan.ranks:
        if me.out_hi <= me.out_lo:
            continue
        lo_tap = max(0, int(np.floor(me.ext_lo * scale)) - 1)
        hi_tap = min(sz - 1, int(np.floor((me.ext_hi - 1) * scale)) + 2)
        for s in range(lo_tap // seg, hi_tap // seg + 1):
            assert me.src_lo <= max(0, s * seg - 16) and me.src_hi >= min(sz, s * seg + 26 + 16), (me, s)
        assert me.src_hi - me.src_lo <= (hi_tap - lo_tap + 1) + 2 * (16 + seg) + 4      # and not much more than that
    loose = SlabPlan(src, (np.float32(voxel),) * 3, gs, pad, world, aligned=False)
    assert all(a.src_lo <= b.src_lo and a.src_hi >= b.src_hi for a, b in zip(plan.ranks, loose.ranks))


def test_aligned_plans_contain_the_loose_ones_for_any_shape():
    """Random shapes / voxel sizes / geometries / rank counts: the segment-aligned block always contains the
    16-plane-horizon block, stays inside the map, and the host entry point never refuses a slab the plan makes."""
    rng = np.random.default_rng(1)
    for _ in range(600):
        sz, vox = int(rng.integers(4, 2200)), float(rng.uniform(0.45, 2.6))
        gs, pad = [(32, 16), (48, 8), (16, 8)][int(rng.integers(3))]
        world = int(rng.integers(1, 9))
        tight = SlabPlan((sz, 8, 8), (np.float32(vox),) * 3, gs, pad, world)
        loose = SlabPlan((sz, 8, 8), (np.float32(vox),) * 3, gs, pad, world, aligned=False)
        for a, b in zip(tight.ranks, loose.ranks):
            if a.out_hi > a.out_lo:
                assert 0 <= a.src_lo <= b.src_lo and b.src_hi <= a.src_hi <= sz, (sz, vox, gs, pad, world, a, b)


def test_far_halos_are_reported_so_the_caller_falls_back():
    plan = SlabPlan((160, 8, 8), (np.float32(1.1),) * 3, 32, 16, 8)           # blocks thinner than the halo
    assert any(PeerHalo.neighbour_plan(plan, r) is None for r in range(8))


def test_balanced_cube_partition_is_even_and_covers_the_map():
    for n, world in ((1331, 8), (1331, 3), (44, 8), (5, 8)):
        b = _split_even(n, world)
        sizes = [b[r + 1] - b[r] for r in range(world)]
        assert b[0] == 0 and b[-1] == n and max(sizes) - min(sizes) <= 1
    xb = _split_even(512, 8)
    assert xb == [64 * r for r in range(9)]


# ------------------------------------------------------------------ session registry lifetime
def test_session_entry_is_invalidated_when_the_file_is_rewritten(tmp_path):
    session.clear()
    p = tmp_path / 'resampled_normalized_map.mrc'
    session.put(str(p), volume='resident')
    assert session.get(str(p))['volume'] == 'resident'                        # no file: the entry stands for it
    p.write_bytes(b'x')
    os.utime(p, (time.time() + 5, time.time() + 5))                           # somebody wrote the file later
    assert session.get(str(p)) is None
    p2 = tmp_path / 'b.mrc'
    p2.write_bytes(b'y')
    session.put(str(p2), volume=1)                                            # written, then registered: valid
    assert session.get(str(p2)) is not None
    os.utime(p2, (time.time() + 9, time.time() + 9))
    assert session.get(str(p2)) is None
    session.put(str(tmp_path / 'grids' / 'a'), x=1)
    session.put(str(tmp_path / 'grids' / 'b'), x=2)
    session.put(str(tmp_path / 'other'), x=3)
    session.release_under(str(tmp_path / 'grids'))
    assert session.get(str(tmp_path / 'grids' / 'a')) is None and session.get(str(tmp_path / 'other')) is not None
    session.clear()


def test_device_volume_behaves_like_the_array():
    a = np.random.default_rng(0).random((20, 6, 5, 4)).astype(np.float32)
    v = DeviceVolume(torch.from_numpy(a.copy()))
    assert v.shape == a.shape and v.ndim == 4 and v.dtype == np.float32 and len(v) == 20 and v.size == a.size
    ix = (np.array([1, 3]), np.array([0, 4]), np.array([2, 2]))
    assert np.array_equal(v[:, ix[0], ix[1], ix[2]], a[:, ix[0], ix[1], ix[2]])
    assert np.array_equal(v[3], a[3]) and v[1, 2, 3, 1] == a[1, 2, 3, 1]
    assert np.array_equal(np.asarray(v), a) and np.array_equal(np.argmax(v, axis=0), np.argmax(a, axis=0))
    assert np.array_equal(v[:, 1:3], a[:, 1:3])                               # served from the host copy now


# ------------------------------------------------------------------ reader edge cases
def _atom(rec, serial, name, alt, resn, chain, resseq, xyz, occ):
    return '%-6s%5d %-4s%s%3s %s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f\n' % (rec, serial, name, alt, resn, chain, resseq,
                                                                       xyz[0], xyz[1], xyz[2], occ, 0.0)


def test_pdb_reader_keeps_one_atom_per_name_like_biopython(tmp_path):
    txt = ''.join([
        _atom('ATOM', 1, ' N  ', ' ', 'ALA', 'A', 1, (1, 2, 3), 1.0),
        _atom('ATOM', 2, ' CA ', 'A', 'ALA', 'A', 1, (4, 5, 6), 0.4),
        _atom('ATOM', 3, ' CA ', 'B', 'ALA', 'A', 1, (7, 8, 9), 0.6),        # higher occupancy wins
        _atom('ATOM', 4, ' C  ', ' ', 'ALA', 'A', 1, (1, 1, 1), 1.0),
        _atom('ATOM', 5, ' C  ', ' ', 'ALA', 'A', 1, (2, 2, 2), 1.0),        # "defined twice": ignored
        _atom('ATOM', 6, ' O  ', 'A', 'ALA', 'A', 1, (3, 3, 3), 0.5),
        _atom('ATOM', 7, ' O  ', 'B', 'ALA', 'A', 1, (4, 4, 4), 0.5),        # a tie: the first stands
        _atom('HETATM', 8, ' O  ', ' ', 'HOH', 'A', 2, (9, 9, 9), 1.0),
        _atom('ATOM', 9, ' CA ', ' ', 'GLY', 'B', 1, (-0.5, 100.25, -12.125), 1.0),
    ])
    p = tmp_path / 'a.pdb'
    p.write_text(txt)
    coords, bb, aa, n_res = pdb.read_pdb_atoms(str(p))
    assert coords.tolist() == [[1, 2, 3], [7, 8, 9], [1, 1, 1], [3, 3, 3], [-0.5, 100.25, -12.125]]
    assert bb.tolist() == [1, 0, 2, 3, 0] and aa.tolist() == [4, 4, 4, 4, 9] and n_res == 2
    rec = pdb.read_pdb_records(str(p))
    assert rec['atom_names'] == ['N', 'CA', 'C', 'O', 'O', 'CA'] and rec['res_index'].tolist() == [0, 0, 0, 0, 1, 2]
    # ragged lines (no fixed width) take the general path and agree
    p2 = tmp_path / 'b.pdb'
    p2.write_text(''.join(ln.rstrip() + '\n' for ln in txt.splitlines()) + 'END\n')
    c2, bb2, aa2, _ = pdb.read_pdb_atoms(str(p2))
    assert np.array_equal(c2, coords) and np.array_equal(bb2, bb) and np.array_equal(aa2, aa)


def test_pdb_coordinates_equal_python_float_parsing(tmp_path):
    rng = np.random.default_rng(5)
    vals = np.round(rng.uniform(-999, 9999, size=(500, 3)), 3)
    lines = [_atom('ATOM', i + 1, ' CA ', ' ', 'ALA', 'A', i % 9999 + 1, v, 1.0) for i, v in enumerate(vals)]
    p = tmp_path / 'c.pdb'
    p.write_text(''.join(lines))
    coords, _, _, _ = pdb.read_pdb_atoms(str(p))
    want = np.array([[float(ln[30:38]), float(ln[38:46]), float(ln[46:54])] for ln in lines]).astype(np.float32)
    assert np.array_equal(coords, want)


def test_mrc_reader_byte_order_compression_and_mode(tmp_path):
    vol = np.random.default_rng(1).random((4, 5, 6)).astype(np.float32)
    p = str(tmp_path / 'le.mrc')
    mrc.write_mrc(p, mrc.MrcMap(data=vol, voxel_size=(np.float32(1.5),) * 3, origin=(1.0, 2.0, 3.0), nzstart=7))
    raw = bytearray(open(p, 'rb').read())
    # big-endian twin: every 4-byte header word and sample swapped, machine stamp 0x11 0x11
    be = bytearray(raw)
    words = np.frombuffer(bytes(raw[:1024]), dtype='<u4').byteswap().tobytes()
    be[:1024] = words
    be[104:112] = raw[104:112]                                 # exttyp / nversion bytes are not read
    be[208:212] = b'MAP '
    be[212:216] = bytes([0x11, 0x11, 0, 0])
    be[1024:] = np.frombuffer(bytes(raw[1024:]), dtype='<f4').astype('>f4').tobytes()
    pb = str(tmp_path / 'be.mrc')
    open(pb, 'wb').write(bytes(be))
    m = mrc.read_mrc(pb)
    assert np.array_equal(m.data, vol) and m.voxel_size == (1.5, 1.5, 1.5) and m.nzstart == 7 and m.mode == 2
    pz = str(tmp_path / 'le.mrc.gz')
    with gzip.open(pz, 'wb') as f:
        f.write(bytes(raw))
    mz = mrc.read_mrc(pz)
    assert np.array_equal(mz.data, vol) and mz.origin == (1.0, 2.0, 3.0)
    pi = str(tmp_path / 'int.mrc')
    mrc.write_mrc(pi, mrc.MrcMap(data=(vol * 100).astype(np.int16)), dtype=np.int16)
    assert mrc.read_mrc(pi).mode == 1                           # DataPreprocessor rejects these (float32 only)
