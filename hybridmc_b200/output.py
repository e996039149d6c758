"""Output files: HDF5 with the reference's dataset / attribute names (src/writer.h,
src/config.cc:18-60, src/main.cc:383-409).

Writing goes through the library's dependency-free HDF5 writer (C-ABI `hmc_h5_*`,
hybridmc_b200/csrc/host/h5lite.h); reading uses the pure-Python parser in `h5lite` (or,
for the legacy HMCB container of earlier builds, `hmcb`).  `File` mimics the h5py access
pattern the reference's tools use: `f['s_bias'][0]`, `f['dist'][:]`,
`f['unique_config']['pos']`, `f.attrs['nonlocal_bonds']`.
"""
import ctypes as C

import numpy as np

from . import h5lite
from .engine import EngineError, load_library

_TYPES = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.uint64): 2,
          np.dtype(np.uint32): 3}


def _put(lib, fn, handle, name, value):
    if isinstance(value, str):
        raw = value.encode()
        rc = fn(handle, name.encode(), 4, 0, None, raw, C.c_uint64(len(raw)))
    else:
        arr = np.asarray(value)
        if arr.dtype not in _TYPES:
            if np.issubdtype(arr.dtype, np.floating):
                arr = arr.astype(np.float64)
            elif np.issubdtype(arr.dtype, np.integer) or arr.dtype == np.bool_:
                arr = arr.astype(np.uint64)
            else:
                raise TypeError("%s: unsupported dtype %s" % (name, arr.dtype))
        if not arr.flags.c_contiguous:  # (ascontiguousarray would turn 0-d into 1-d)
            arr = arr.copy(order="C")
        dims = (C.c_uint64 * max(arr.ndim, 1))(*arr.shape)
        rc = fn(handle, name.encode(), _TYPES[arr.dtype], arr.ndim, dims,
                arr.ctypes.data_as(C.c_void_p), C.c_uint64(arr.nbytes))
    if rc != 0:
        raise EngineError(lib.hmc_last_error(None).decode())


def write_hdf5(path, datasets, attrs=None, fsync=False):
    """datasets: {"name" or "group/name": ndarray (f64/f32/u64/u32) | str}; attrs: root
    attributes.  The whole file is written in one call."""
    lib = load_library()
    lib.hmc_h5_dataset.argtypes = lib.hmc_h5_attribute.argtypes = [
        C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]
    lib.hmc_h5_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.hmc_h5_destroy.argtypes = [C.c_void_p]
    h = C.c_void_p()
    if lib.hmc_h5_create(C.byref(h)) != 0:
        raise EngineError(lib.hmc_last_error(None).decode())
    try:
        for name, value in datasets.items():
            _put(lib, lib.hmc_h5_dataset, h, name, value)
        for name, value in (attrs or {}).items():
            _put(lib, lib.hmc_h5_attribute, h, name, value)
        if lib.hmc_h5_write(h, str(path).encode(), int(fsync)) != 0:
            raise EngineError(lib.hmc_last_error(None).decode())
    finally:
        lib.hmc_h5_destroy(h)


class _Group(dict):
    def __getitem__(self, key):
        if key in self.keys():
            return dict.__getitem__(self, key)
        sub = {k[len(key) + 1:]: v for k, v in self.items() if k.startswith(key + "/")}
        if not sub:
            raise KeyError(key)
        return _Group(sub)


class File(_Group):
    """Read an output / snapshot file written by the engine (HDF5, or legacy HMCB)."""

    def __init__(self, path):
        super().__init__()
        if h5lite.is_hdf5(path):
            datasets, self.attrs = h5lite.read(path)
            self.update(datasets)
        else:
            from . import hmcb
            f = hmcb.File(path)
            self.update(dict.items(f))
            self.attrs = f.attrs

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
