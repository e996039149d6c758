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
