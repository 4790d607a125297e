"""The GTiff-writing calls of dataset/utils.py:47-85 (save_image) on top of the tifffile shim; geo-referencing is ignored."""
import numpy as np

import tifffile

GDT_Byte, GDT_UInt16, GDT_Float32 = 1, 2, 6
_DT = {GDT_Byte: np.uint8, GDT_UInt16: np.uint16, GDT_Float32: np.float32}


class _Band:
    def __init__(self, ds, idx):
        self._ds, self._idx = ds, idx

    def WriteArray(self, array):
        a = np.asarray(array)
        dt = self._ds._dtype
        if np.issubdtype(dt, np.integer):
            info = np.iinfo(dt)
            a = np.clip(np.nan_to_num(a), info.min, info.max)      # GDAL clamps on conversion to the band type
        self._ds._planes[self._idx] = a.astype(dt)
        self._ds._dirty = True

    def FlushCache(self):
        self._ds.FlushCache()


class _Dataset:
    def __init__(self, path, cols, rows, chans, dtype):
        self._path, self._dtype = path, np.dtype(dtype)
        self._planes = [np.zeros((rows, cols), dtype) for _ in range(chans)]
        self._dirty = True

    def SetGeoTransform(self, t):
        pass

    def SetProjection(self, wkt):
        pass

    def GetRasterBand(self, i):
        return _Band(self, i - 1)

    def FlushCache(self):
        if self._dirty:
            if len(self._planes) == 1:
                tifffile.imwrite(self._path, self._planes[0])
            else:
                tifffile.imwrite(self._path, np.stack(self._planes, axis=0), planar=True)
            self._dirty = False

    def __del__(self):
        try:
            self.FlushCache()
        except Exception:
            pass


class _Driver:
    def Create(self, path, cols, rows, chans, dtype=GDT_Byte):
        return _Dataset(path, cols, rows, chans, _DT[dtype])


def GetDriverByName(name):
    if name != "GTiff":
        raise ValueError("the gdal shim only writes GTiff")
    return _Driver()


class _Opened:
    def __init__(self, path):
        self._a = tifffile.imread(path)

    def ReadAsArray(self):
        a = self._a
        return a.transpose(2, 0, 1) if a.ndim == 3 and a.shape[2] <= 16 and a.shape[0] > 16 else a


def Open(path):
    return _Opened(path)
