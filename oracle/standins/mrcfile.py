"""Minimal stand-in for the third-party ``mrcfile`` package (absent from this
image), used ONLY by oracle/ref_harness.py to execute the unmodified reference
in the build container.  Implements the subset the reference touches
(utils/preprocessing.py:98-107,138-148,196-206; utils/create_grids.py:108-117):
``open`` / ``new`` context managers, ``.data`` (nz,ny,nx), ``.voxel_size``,
``.header.<field>``, ``.set_data``, ``.update_header_stats``.  MRC2014 layout:
1024-byte little-endian header + float32 payload."""
import builtins

import numpy as np

_HDR = np.dtype([
    ('nx', '<i4'), ('ny', '<i4'), ('nz', '<i4'), ('mode', '<i4'),
    ('nxstart', '<i4'), ('nystart', '<i4'), ('nzstart', '<i4'),
    ('mx', '<i4'), ('my', '<i4'), ('mz', '<i4'),
    ('cella', [('x', '<f4'), ('y', '<f4'), ('z', '<f4')]),
    ('cellb', [('alpha', '<f4'), ('beta', '<f4'), ('gamma', '<f4')]),
    ('mapc', '<i4'), ('mapr', '<i4'), ('maps', '<i4'),
    ('dmin', '<f4'), ('dmax', '<f4'), ('dmean', '<f4'),
    ('ispg', '<i4'), ('nsymbt', '<i4'),
    ('extra1', 'V8'), ('exttyp', 'S4'), ('nversion', '<i4'), ('extra2', 'V84'),
    ('origin', [('x', '<f4'), ('y', '<f4'), ('z', '<f4')]),
    ('map', 'S4'), ('machst', 'u1', (4,)), ('rms', '<f4'), ('nlabl', '<i4'),
    ('label', 'S80', (10,)),
])
assert _HDR.itemsize == 1024


class _Mrc:
    def __init__(self, path, mode):
        self._path, self._mode = path, mode
        if mode == 'r':
            with builtins.open(path, 'rb') as f:
                self.header = np.frombuffer(f.read(1024), dtype=_HDR)[0].copy().view(np.recarray)
                f.seek(1024 + int(self.header.nsymbt))
                n = int(self.header.nx) * int(self.header.ny) * int(self.header.nz)
                self.data = np.frombuffer(f.read(4 * n), dtype='<f4').reshape(
                    int(self.header.nz), int(self.header.ny), int(self.header.nx))
        else:
            self.header = np.zeros((), dtype=_HDR).view(np.recarray)
            self.header.mode = 2
            self.header.mapc, self.header.mapr, self.header.maps = 1, 2, 3
            self.header.cellb.alpha = self.header.cellb.beta = self.header.cellb.gamma = 90.0
            self.header.map = b'MAP '
            self.header.machst = [0x44, 0x44, 0, 0]
            self.header.nversion = 20140
            self.header.ispg = 1
            self.data = None

    # --- voxel size: cella / m{x,y,z} as a float32 record -------------------
    @property
    def voxel_size(self):
        h = self.header
        v = np.zeros((), dtype=[('x', '<f4'), ('y', '<f4'), ('z', '<f4')]).view(np.recarray)
        v.x = h.cella.x / h.mx if h.mx else 0
        v.y = h.cella.y / h.my if h.my else 0
        v.z = h.cella.z / h.mz if h.mz else 0
        return v

    @voxel_size.setter
    def voxel_size(self, size):
        if getattr(getattr(size, 'dtype', None), 'names', None):      # the (x, y, z) record voxel_size returns
            sx, sy, sz = size['x'], size['y'], size['z']
        else:
            try:
                sx, sy, sz = size
            except TypeError:
                sx = sy = sz = size
        h = self.header
        h.cella.x, h.cella.y, h.cella.z = sx * h.mx, sy * h.my, sz * h.mz

    def set_data(self, data):
        data = np.ascontiguousarray(data, dtype=np.float32)
        self.data = data
        h = self.header
        h.nz, h.ny, h.nx = data.shape
        h.mz, h.my, h.mx = data.shape
        h.cella.x, h.cella.y, h.cella.z = data.shape[2], data.shape[1], data.shape[0]

    def update_header_stats(self):
        h = self.header
        h.dmin, h.dmax = self.data.min(), self.data.max()
        h.dmean = self.data.mean(dtype=np.float64)
        h.rms = self.data.std(dtype=np.float64)

    def close(self):
        if self._mode == 'w':
            with builtins.open(self._path, 'wb') as f:
                f.write(np.asarray(self.header).tobytes())
                f.write(self.data.tobytes())

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def open(path, mode='r', permissive=False):
    return _Mrc(path, 'r')


def new(path, data=None, overwrite=False):
    m = _Mrc(path, 'w')
    if data is not None:
        m.set_data(data)
    return m
