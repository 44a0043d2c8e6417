"""MRC2014 reader/writer for the pipeline's on-disk products.

Replaces the ``mrcfile`` calls at utils/preprocessing.py:98-107,138-148,196-206
and utils/create_grids.py:108-117: float32 mode-2 maps, ``data`` shaped
(nz,ny,nx), ``voxel_size = cella / m{x,y,z}`` as float32, origin / axis order /
n*start carried through.  Pure I/O -- no arithmetic of the hot path lives here."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

_MODES = {0: np.int8, 1: np.int16, 2: np.float32, 6: np.uint16, 12: np.float16}


@dataclass
class MrcMap:
    data: np.ndarray                         # (nz, ny, nx)
    voxel_size: tuple = (1.0, 1.0, 1.0)       # (x, y, z) as np.float32
    origin: tuple = (0.0, 0.0, 0.0)           # (x, y, z) as np.float32
    mapc: int = 1
    mapr: int = 2
    maps: int = 3
    nxstart: int = 0
    nystart: int = 0
    nzstart: int = 0
    mode: int = 2                             # MRC sample type of the file (2 = float32)
    extra: dict = field(default_factory=dict)


def _open_bytes(path):
    """(header bytes, payload source).  gzip / bzip2 files are recognised by their magic numbers, as
    ``mrcfile.open`` does, and decompressed into memory; plain files are memory-mapped."""
    with open(path, 'rb') as f:
        magic = f.read(3)
    if magic[:2] == b'\x1f\x8b':
        import gzip
        with gzip.open(path, 'rb') as f:
            return f.read()
    if magic == b'BZh':
        import bz2
        with bz2.open(path, 'rb') as f:
            return f.read()
    return None


def read_mrc(path) -> MrcMap:
    """Header + payload of an MRC2014 file.  The byte order comes from the machine stamp (0x11 0x11 =
    big-endian, anything else is read little-endian like the files every writer of this pipeline
    produces).  ``data`` keeps the file's sample type (``mode`` says which); for plain files it is a
    copy-on-write memory map, so nothing is read before the caller touches it."""
    blob = _open_bytes(path)
    if blob is None:
        with open(path, 'rb') as f:
            hdr = f.read(1024)
    else:
        hdr = blob[:1024]
    if len(hdr) < 1024:
        raise ValueError(f'{path}: truncated MRC header')
    bo = '>' if hdr[212:214] == b'\x11\x11' else '<'
    w = np.frombuffer(hdr, dtype=bo + 'i4', count=56)
    fl = np.frombuffer(hdr, dtype=bo + 'f4', count=56)
    nx, ny, nz, mode = (int(v) for v in w[0:4])
    if mode not in _MODES or min(nx, ny, nz) <= 0:
        raise ValueError(f'{path}: unsupported MRC (mode={mode}, shape={(nz, ny, nx)})')
    nsymbt = max(int(w[23]), 0)
    dt = np.dtype(_MODES[mode]).newbyteorder(bo)
    count = nx * ny * nz
    if blob is None:
        import os
        if os.path.getsize(path) < 1024 + nsymbt + dt.itemsize * count:
            raise ValueError(f'{path}: truncated MRC payload')
        data = np.memmap(path, dtype=dt, mode='c', offset=1024 + nsymbt, shape=(count,))
    else:
        data = np.frombuffer(blob, dtype=dt, count=-1, offset=1024 + nsymbt)[:count]
        if data.size != count:
            raise ValueError(f'{path}: truncated MRC payload')
    if bo == '>':
        data = data.astype(dt.newbyteorder('<'))
    mx, my, mz = (int(v) for v in w[7:10])
    cella = fl[10:13]
    f32 = np.float32
    vs = tuple(f32(cella[a]) / f32(m) if m else f32(0) for a, m in enumerate((mx, my, mz)))
    return MrcMap(data=data.reshape(nz, ny, nx), voxel_size=vs,
                  origin=tuple(f32(v) for v in fl[49:52]),
                  mapc=int(w[16]), mapr=int(w[17]), maps=int(w[18]),
                  nxstart=int(w[4]), nystart=int(w[5]), nzstart=int(w[6]), mode=mode)


def write_mrc(path, m: MrcMap, dtype=np.float32):
    """``dtype``: float32 (mode 2, what every reference writer produces) or one of the other
    MRC2014 sample types (int8 / int16 / uint16 / float16), e.g. for label masks."""
    mode = {np.dtype(v): k for k, v in _MODES.items()}[np.dtype(dtype)]
    data = np.ascontiguousarray(m.data, dtype=np.dtype(dtype).newbyteorder('<'))
    nz, ny, nx = data.shape
    hdr = np.zeros(256, dtype='<i4')
    fl = hdr.view('<f4')
    hdr[0:4] = (nx, ny, nz, mode)
    hdr[4:7] = (m.nxstart, m.nystart, m.nzstart)
    hdr[7:10] = (nx, ny, nz)
    vx, vy, vz = (np.float32(v) for v in m.voxel_size)
    fl[10:13] = (vx * nx, vy * ny, vz * nz)
    fl[13:16] = 90.0
    hdr[16:19] = (m.mapc, m.mapr, m.maps)
    fl[19] = data.min() if data.size else 0
    fl[20] = data.max() if data.size else 0
    fl[21] = data.mean(dtype=np.float64) if data.size else 0
    hdr[22] = 1
    hdr[27] = 20140
    fl[49:52] = [np.float32(v) for v in m.origin]
    raw = bytearray(hdr.tobytes())
    raw[208:212] = b'MAP '
    raw[212:216] = bytes([0x44, 0x44, 0, 0])
    raw[216:220] = np.float32(data.std(dtype=np.float64) if data.size else 0).tobytes()
    with open(path, 'wb') as f:
        f.write(bytes(raw))
        f.write(data.tobytes())
