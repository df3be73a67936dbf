"""TEST INFRASTRUCTURE ONLY -- minimal numpy-backed subset of the GDAL python API (see package docstring).

Container format (for every raster regardless of file extension): b"LBRS" + 1 byte codec (0 raw, 1 zlib)
followed by an .npy payload holding a CHW array.  `bin/gdal_translate` converts between codec 0 ("GTiff")
and codec 1 (a lossless stand-in for "JP2OpenJPEG ... REVERSIBLE=YES").
"""
import io
import zlib

import numpy as np

GDT_Byte, GDT_UInt16, GDT_Float32, GDT_Float64 = 1, 2, 6, 7
_NP = {GDT_Byte: np.uint8, GDT_UInt16: np.uint16, GDT_Float32: np.float32, GDT_Float64: np.float64}
_GDT = {np.dtype(v): k for k, v in _NP.items()}
_MAGIC = b"LBRS"


def UseExceptions():
    return None


def _load(path):
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:4] != _MAGIC:
        raise RuntimeError(f"{path}: not a shim raster")
    payload = blob[5:]
    if blob[4] == 1:
        payload = zlib.decompress(payload)
    arr = np.load(io.BytesIO(payload), allow_pickle=False)
    return arr.reshape((-1,) + arr.shape[-2:])


def _store(path, arr, codec=0):
    buf = io.BytesIO()
    np.save(buf, np.ascontiguousarray(arr), allow_pickle=False)
    payload = buf.getvalue()
    if codec == 1:
        payload = zlib.compress(payload, 6)
    with open(path, "wb") as f:
        f.write(_MAGIC + bytes([codec]) + payload)


class _Band:
    def __init__(self, ds, idx):
        self._ds, self._idx = ds, idx
        self.DataType = _GDT[ds._arr.dtype]

    def WriteArray(self, array, xoff=0, yoff=0):
        a = np.asarray(array)
        self._ds._arr[self._idx, yoff:yoff + a.shape[0], xoff:xoff + a.shape[1]] = a
        self._ds._dirty = True
        return 0

    def ReadAsArray(self):
        return self._ds._arr[self._idx].copy()


class Dataset:
    def __init__(self, arr, path=None, writable=False):
        self._arr, self._path, self._writable, self._dirty = arr, path, writable, writable
        self.RasterCount, self.RasterYSize, self.RasterXSize = arr.shape

    def ReadAsArray(self):
        a = self._arr.copy()
        return a[0] if a.shape[0] == 1 else a

    def GetRasterBand(self, i):
        return _Band(self, i - 1)

    def WriteArray(self, array, xoff=0, yoff=0):
        a = np.asarray(array)
        a = a.reshape((-1,) + a.shape[-2:])
        self._arr[:, yoff:yoff + a.shape[1], xoff:xoff + a.shape[2]] = a
        self._dirty = True
        return 0

    def FlushCache(self):
        if self._writable and self._path is not None and self._dirty:
            _store(self._path, self._arr)
            self._dirty = False

    def __del__(self):
        try:
            self.FlushCache()
        except Exception:
            pass


class _Driver:
    def Create(self, path, xsize, ysize, bands=1, etype=GDT_Byte):
        return Dataset(np.zeros((bands, ysize, xsize), dtype=_NP[etype]), path, writable=True)


def GetDriverByName(name):
    return _Driver()


def Open(path, *args):
    return Dataset(_load(path), path)


def Translate(dest, src, srcWin=None, **kw):
    ds = Open(src) if isinstance(src, str) else src
    a = ds._arr
    if srcWin is not None:
        x, y, w, h = srcWin
        a = a[:, y:y + h, x:x + w]
    _store(dest, a)
    return Open(dest)
