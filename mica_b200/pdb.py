"""Fixed-column PDB reader producing the atom arrays the AF3 rasteriser consumes.

Stands where ``Bio.PDB.PDBParser.get_structure`` stands in the reference
(utils/preprocessing.py:269,275-298): keeps atoms of residues whose hetero flag
is ' ' (ATOM records, :279), coordinates as float32 (Bio.PDB stores float32),
atom name = columns 13-16 stripped, residue name = columns 18-20.  Text munging
only; the channel mapping mirrors :254-263,180-185,292-298.

The records are parsed by ``mica_parse_pdb`` in libmica_b200.so (a 160 k-atom docked model: a few ms;
a column-wise NumPy reader with identical results serves when the library is not built).  Like Bio.PDB's StructureBuilder, one atom per (model, chain, residue id,
atom name) survives: of alternate locations the one with the highest occupancy (the first of
equals), and a second record of the same name with a blank altloc is dropped (Bio.PDB warns
"defined twice" and ignores it).  The survivor sits where the name first appeared, which is
where iterating the Bio.PDB residue visits it.  Point mutations (Bio.PDB DisorderedResidue)
are not emulated."""
from __future__ import annotations

import numpy as np

BACKBONE_ATOMS = ['CA', 'N', 'C', 'O']
AMINO_ACIDS = ['ALA', 'CYS', 'ASP', 'GLU', 'PHE', 'GLY', 'HIS', 'ILE', 'LYS', 'LEU',
               'MET', 'ASN', 'PRO', 'GLN', 'ARG', 'SER', 'THR', 'VAL', 'TRP', 'TYR']
CHANNEL_NAMES = BACKBONE_ATOMS + AMINO_ACIDS
_BB = {n: i for i, n in enumerate(BACKBONE_ATOMS)}
_AA = {n: 4 + i for i, n in enumerate(AMINO_ACIDS)}
_COLS = 60          # columns 1-60 hold everything read here (through the occupancy)


def channel_codes(atom_names, res_names):
    """int8 backbone channel (0..3 | -1) and amino-acid channel (4..23 | -1) per atom."""
    bb = np.fromiter((_BB.get(a, -1) for a in atom_names), dtype=np.int8, count=len(atom_names))
    aa = np.fromiter((_AA.get(r, -1) for r in res_names), dtype=np.int8, count=len(res_names))
    return bb, aa


def _codes_of(field4, table):
    """Map a uint8 [A,4] text field through ``table`` (stripped str -> int) -> int8, -1 when absent."""
    if len(field4) == 0:
        return np.zeros(0, np.int8)
    uniq, inv = np.unique(np.ascontiguousarray(field4).view('<u4').ravel(), return_inverse=True)
    lut = np.array([table.get(u.tobytes().decode('ascii', 'replace').strip(), -1) for u in uniq.astype('<u4')],
                   dtype=np.int8)
    return lut[inv]


def _fixed3(txt8):
    """``%8.3f`` fields (uint8 [A,8]: `` *-?[0-9]+`` right-aligned in columns 1-4, '.', three digits) ->
    float64, bit-identical to ``float(text)``: the digits form an exact integer and ONE IEEE division by
    1000 is the correctly rounded decimal.  Returns None when a field does not have that shape (the
    caller then parses the text)."""
    t = np.ascontiguousarray(txt8.T)                      # [8, A]: one contiguous row per column
    dig = (t >= 48) & (t <= 57)
    minus = t[:4] == 45
    if not ((t[4] == 46).all() and dig[5:8].all() and dig[3].all()):
        return None
    if not (dig[:3] <= dig[1:4]).all():                   # digits are contiguous up to the point
        return None
    if not (dig[:4] | minus | (t[:4] == 32)).all():
        return None
    if (minus[3]).any() or (minus[:3] & ~dig[1:4]).any():  # a sign sits directly before the digits
        return None
    if (minus[1:4] & (t[0:3] != 32)).any():               # and only blanks before the sign
        return None
    d = (t - 48) * dig                                    # uint8 digit values, 0 where blank / sign
    n = d[0].astype(np.int32)
    for j in (1, 2, 3, 5, 6, 7):
        n *= 10
        n += d[j]
    val = n.astype(np.float64) / 1000.0
    return np.where(minus.any(axis=0), -val, val)


def _records_numpy(path, with_hetatm):
    """Column-wise parse of the ATOM (and HETATM) records.  Returns a dict of per-record arrays in file
    order after the Bio.PDB de-duplication: cols (uint8 [A,60]), is_het, model, coords float32 [A,3]."""
    with open(path, 'rb') as fh:
        data = fh.read()
    if not data:
        return None
    raw = np.frombuffer(data, dtype=np.uint8)
    L = data.find(b'\n') + 1
    n_full = raw.size // L if L else 0
    tail = raw[n_full * L:] if L else raw
    fixed = (L >= _COLS + 1 and n_full > 0 and bool((raw[L - 1:n_full * L:L] == 10).all())
             and data.count(b'\n') == n_full + int(tail.size > 0 and tail[-1] == 10))
    has_cr = data.find(b'\r') >= 0
    if fixed:
        # every line has the same length (what PDB writers produce): the file IS the column table
        table = raw[:n_full * L].reshape(n_full, L)
        if tail.size >= 6:                               # a last line without the full width
            last = np.full((1, L), 32, np.uint8)
            last[0, :min(tail.size, L)] = tail[:L]
            last[last == 10] = 32
            table = np.concatenate([table, last])
        tag = np.ascontiguousarray(table[:, :6]).view('S6').ravel()
        is_atom, is_het = tag == b'ATOM  ', tag == b'HETATM'
        sel = is_atom | (is_het if with_hetatm else False)
        is_model = np.char.startswith(tag, b'MODEL') if not sel.all() else np.zeros(len(tag), bool)
        model = np.cumsum(is_model)[sel]
        cols = np.ascontiguousarray(table[sel, :_COLS])
        het = is_het[sel]
    else:
        nl = np.flatnonzero(raw == 10)
        starts = np.concatenate(([0], nl + 1))
        ends = np.concatenate((nl, [raw.size]))
        keep = ends > starts
        starts, ends = starts[keep], ends[keep]
        head = np.full((len(starts), 6), 32, dtype=np.uint8)
        for c in range(6):
            ok = starts + c < ends
            head[ok, c] = raw[starts[ok] + c]
        tag = head.view('S6').ravel()
        is_atom, is_het = tag == b'ATOM  ', tag == b'HETATM'
        is_model = np.char.startswith(tag, b'MODEL')
        sel = is_atom | (is_het if with_hetatm else False)
        model = np.cumsum(is_model)[sel]                 # records of one MODEL share a serial
        s, e = starts[sel], ends[sel]
        idx = s[:, None] + np.arange(_COLS)[None, :]
        inside = idx < e[:, None]
        cols = np.where(inside, raw[np.minimum(idx, raw.size - 1)], 32).astype(np.uint8)
        het = is_het[sel]
    if has_cr:
        cols[cols == 13] = 32                            # CR of CRLF files
    if len(cols):
        txt = np.ascontiguousarray(cols[:, 30:54]).reshape(-1, 8)      # x, y, z fields of every record
        v = _fixed3(txt)
        if v is None:
            v = txt.view('S8').ravel().astype(np.float64)
        xyz = v.reshape(-1, 3).astype(np.float32)
    else:
        xyz = np.zeros((0, 3), np.float32)
    n = len(cols)
    fields = np.zeros((n, 16), dtype=np.uint8)
    occ = np.zeros(n, np.float32)
    if n:
        fields[:, 0:4] = cols[:, 12:16]
        fields[:, 4] = cols[:, 16]
        fields[:, 5:8] = cols[:, 17:20]
        fields[:, 8] = cols[:, 21]
        fields[:, 9:14] = cols[:, 22:27]
        fields[:, 14] = het
        occ_txt = np.char.strip(np.ascontiguousarray(cols[:, 54:60]).view('S6').ravel())
        occ = np.array([float(t) if t else 0.0 for t in occ_txt], dtype=np.float32)
    return fields, np.ascontiguousarray(xyz), occ, model.astype(np.int32)


def _records_native(path, with_hetatm):
    """The same four arrays from libmica_b200.so's ``mica_parse_pdb`` (a 160 k-atom model: ~3 ms instead of
    ~40 ms).  Returns None when the library is not available (then the NumPy reader does the work)."""
    try:
        from . import _lib
    except ImportError:
        return None
    import ctypes as C
    with open(path, 'rb') as fh:
        data = fh.read()
    if not data:
        return np.zeros((0, 16), np.uint8), np.zeros((0, 3), np.float32), np.zeros(0, np.float32), np.zeros(0, np.int32)
    cap = len(data) // 54 + 1                             # a record needs its 54 columns
    xyz = np.empty((cap, 3), np.float32)
    fields = np.empty((cap, 16), np.uint8)
    occ = np.empty(cap, np.float32)
    model = np.empty(cap, np.int32)
    bb, aa = np.empty(cap, np.int8), np.empty(cap, np.int8)
    info = np.zeros(2, np.int64)
    n = _lib.lib.mica_parse_pdb(data, len(data), int(bool(with_hetatm)), cap, xyz.ctypes.data, fields.ctypes.data,
                                occ.ctypes.data, model.ctypes.data, bb.ctypes.data, aa.ctypes.data, info.ctypes.data)
    if n < 0:
        raise ValueError(f'{path}: {_lib.last_error()}')
    _records_native.last = dict(bb=bb[:n], aa=aa[:n], n_res=int(info[0]), dup=bool(info[1]))
    return fields[:n], xyz[:n], occ[:n], model[:n]


def _records(path, with_hetatm, native=True):
    """The ATOM (and HETATM) records after the Bio.PDB de-duplication, in file order: dict(fields uint8
    [A,16] -- 0-3 atom name with its spacing, 4 altloc, 5-7 residue name, 8 chain, 9-13 resSeq + iCode,
    14 HETATM flag --, coords float32 [A,3], model int32 [A])."""
    got = _records_native(path, with_hetatm) if native else None
    if got is None:
        got = _records_numpy(path, with_hetatm)
        if got is None:
            return None
    fields, xyz, occ, model = got
    n = len(fields)
    extras = getattr(_records_native, 'last', None) if native and got is not None else None
    _records_native.last = None
    if extras is not None and len(extras['bb']) == n and not extras['dup']:
        return dict(fields=fields, coords=np.ascontiguousarray(xyz), model=model, native=extras)
    if n:
        # ---- Bio.PDB de-duplication: one atom per (model, chain, het, resseq+icode, full atom name)
        key = np.zeros((n, 24), dtype=np.uint8)
        key[:, 0:4] = fields[:, 0:4]
        key[:, 4:10] = fields[:, 8:14]                    # chain, resSeq + iCode
        key[:, 10] = fields[:, 14]
        key[:, 12:16] = model.astype('<u4').view(np.uint8).reshape(-1, 4)
        h = key[:, :16].copy().view('<u8')
        hs = np.sort(h[:, 0] * np.uint64(0x9E3779B97F4A7C15) ^ h[:, 1])
        cnt = np.zeros(1, np.int64)
        if len(hs) > 1 and (hs[1:] == hs[:-1]).any():      # (a hash collision only costs the exact check)
            kv = key.view('S24').ravel()
            _, first, inv, cnt = np.unique(kv, return_index=True, return_inverse=True, return_counts=True)
        if (cnt > 1).any():
            altloc = fields[:, 4]
            winner = first.copy()                         # per group: record whose coordinates survive
            for g in np.flatnonzero(cnt > 1):
                members = np.flatnonzero(inv == g)
                dis = members[altloc[members] != 32]
                if len(dis) == 0:
                    continue                              # blank-altloc duplicates: the first stands
                best, best_occ = -1, -np.inf              # DisorderedAtom.disordered_add: strictly greater wins;
                pool = list(dis)                          # a blank-altloc record met by an altloc one joins the pool
                if altloc[members[0]] == 32:
                    pool = list(dis[:1]) + [members[0]] + list(dis[1:])
                for m_ in pool:
                    if occ[m_] > best_occ:
                        best, best_occ = m_, occ[m_]
                winner[g] = best
            keep_pos = np.sort(first)                     # survivors sit where the name first appeared
            src = winner[inv[keep_pos]]
            fields, xyz, model = fields[keep_pos], xyz[src], model[keep_pos]
    return dict(fields=fields, coords=np.ascontiguousarray(xyz), model=model)


def _residue_index(rec):
    """Number the residues in file order: a new one starts whenever model / chain / record type /
    sequence number / insertion code changes."""
    f = rec['fields']
    n = len(f)
    if n == 0:
        return np.zeros(0, np.int64)
    key = np.zeros((n, 12), dtype=np.uint8)
    key[:, 0:7] = f[:, 8:15]                              # chain, resSeq + iCode, HETATM flag
    key[:, 7:11] = rec['model'].astype('<u4').view(np.uint8).reshape(-1, 4)
    kv = key.view('S12').ravel()
    change = np.concatenate(([True], kv[1:] != kv[:-1]))
    return np.cumsum(change) - 1


def _names(rec):
    f = rec['fields']
    atom = np.char.strip(np.ascontiguousarray(f[:, 0:4]).view('S4').ravel())
    res = np.char.strip(np.ascontiguousarray(f[:, 5:8]).view('S3').ravel())
    return atom, res


def _res_field4(fields):
    f = np.full((len(fields), 4), 32, np.uint8)
    f[:, :3] = fields[:, 5:8]
    return f


def read_pdb_atoms(path):
    """Returns (coords float32 [A,3] x,y,z; bb_ch int8 [A]; aa_ch int8 [A];
    n_residues) for the ATOM records of ``path``."""
    rec = _records(path, with_hetatm=False)
    if rec is None or len(rec['fields']) == 0:
        return np.zeros((0, 3), np.float32), np.zeros(0, np.int8), np.zeros(0, np.int8), 0
    fast = rec.get('native')
    if fast is not None:                                   # no duplicates: the parser's own codes and count stand
        return rec['coords'], fast['bb'], fast['aa'], fast['n_res']
    ridx = _residue_index(rec)
    f = rec['fields']
    return rec['coords'], _codes_of(f[:, 0:4], _BB), _codes_of(_res_field4(f), _AA), int(ridx[-1]) + 1


def read_pdb_records(path):
    """Every ATOM and HETATM record in file order -- what iterating a Bio.PDB structure
    ``for model / chain / residue / atom`` visits in the label-mask builders
    (scripts_for_training_data/create_backbone_mask.py:143-147; no hetero filter there).
    Returns dict(coords float32 [A,3], atom_names [A], res_names [A], res_index int64 [A]);
    ``res_index`` numbers the residues (a new one starts whenever chain / hetero flag /
    sequence number / insertion code changes)."""
    rec = _records(path, with_hetatm=True)
    if rec is None or len(rec['fields']) == 0:
        return dict(coords=np.zeros((0, 3), np.float32), atom_names=[], res_names=[],
                    res_index=np.zeros(0, np.int64))
    atom, res = _names(rec)
    return dict(coords=rec['coords'], atom_names=[a.decode('ascii', 'replace') for a in atom],
                res_names=[r.decode('ascii', 'replace') for r in res], res_index=_residue_index(rec))
