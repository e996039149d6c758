"""Reader for the HMCB container the `hybridmc` executable writes while libhdf5 is not
available (hybridmc_b200/csrc/host/hybridmc_main.cpp).  It holds the reference's dataset
and attribute names and shapes; `File` mimics the h5py access pattern the reference's
tools use (`f['s_bias'][0]`, `f['dist'][:]`, `f['unique_config']['pos']`, `f.attrs`)."""
import struct

import numpy as np

_DT = {0: np.float64, 1: np.float32, 2: np.uint64, 3: np.uint32}


class _Group(dict):
    def __getitem__(self, key):
        if key in self.keys():
            return dict.__getitem__(self, key)
        sub = {k[len(key) + 1:]: v for k, v in self.items() if k.startswith(key + "/")}
        if not sub:
            raise KeyError(key)
        return _Group(sub)


class File(_Group):
    def __init__(self, path):
        super().__init__()
        self.attrs = {}
        with open(path, "rb") as f:
            data = f.read()
        if data[:6] != b"HMCB1\n":
            raise ValueError("%s is not an HMCB container" % path)
        i = 6
        while i < len(data):
            kind = data[i]
            (nl,) = struct.unpack_from("<H", data, i + 1)
            name = data[i + 3:i + 3 + nl].decode()
            i += 3 + nl
            dtype, rank = data[i], data[i + 1]
            dims = struct.unpack_from("<%dQ" % rank, data, i + 2)
            i += 2 + 8 * rank
            (nb,) = struct.unpack_from("<Q", data, i)
            i += 8
            raw = data[i:i + nb]
            i += nb
            if dtype == 4:
                val = raw.decode()
            else:
                val = np.frombuffer(raw, dtype=_DT[dtype]).reshape(dims) if rank else \
                    np.frombuffer(raw, dtype=_DT[dtype])[0]
            (self.attrs if kind == 1 else self)[name] = val

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# chunk shapes / filters of the reference's writers (src/writer.h:14-220, src/config.cc:34-46)
_REF_CHUNKS = {
    "config/int": lambda shp: (1024,),
    "unique_config/pos": lambda shp: (10,) + tuple(shp[1:]),
    "unique_config/vel": lambda shp: (10,) + tuple(shp[1:]),
    "unique_config/config": lambda shp: (1000,),
    "config_count": lambda shp: (100,) + tuple(shp[1:]),
    "dist": lambda shp: (10000,) + tuple(shp[1:]),
}


def to_hdf5(src, dst):
    """Convert an HMCB container into a real HDF5 file with the reference's layout
    (names, dtypes, unlimited first dimension, chunking, deflate 6 on pos/vel) so that the
    reference's tools/*.py read it.  Needs h5py, which this image does not ship: run it
    wherever the analysis tools run.  `python -m hybridmc_b200.hmcb in.h5 out.h5`."""
    try:
        import h5py
    except ImportError as exc:  # pragma: no cover - h5py is absent from the build image
        raise RuntimeError("to_hdf5 needs h5py (not installed here)") from exc
    f = File(src)
    with h5py.File(dst, "w") as out:
        for name, val in dict.items(f):
            if isinstance(val, str):
                out.create_dataset(name, data=val, dtype=h5py.string_dtype())
                continue
            arr = np.asarray(val)
            if name in _REF_CHUNKS and arr.ndim >= 1:
                kw = dict(maxshape=(None,) + arr.shape[1:], chunks=_REF_CHUNKS[name](arr.shape))
                if name in ("unique_config/pos", "unique_config/vel"):
                    kw.update(compression="gzip", compression_opts=6)
                out.create_dataset(name, data=arr, **kw)
            else:
                out.create_dataset(name, data=arr)
        for name, val in f.attrs.items():
            out.attrs[name] = val


if __name__ == "__main__":
    import sys
    to_hdf5(sys.argv[1], sys.argv[2])
