"""HDF5 output without libhdf5: the C++ writer
(hybridmc_b200/csrc/host/h5lite.h, C-ABI hmc_h5_*) against two independent readers — the
pure-Python parser (hybridmc_b200/h5lite.py) and the C++ reader the executable uses — and
both readers against a file written by libhdf5 itself
(tests/golden/libhdf5_written_matlab73.mat: scipy's MATLAB-7.3 fixture, an HDF5 file behind
a 512-byte user block), which pins the byte-level conventions of the format."""
import os
import struct
import subprocess

import numpy as np
import pytest

from hybridmc_b200 import h5lite, output

HERE = os.path.dirname(os.path.abspath(__file__))
LIBHDF5_FILE = os.path.join(HERE, "golden", "libhdf5_written_matlab73.mat")


def fnv1a(raw):
    h = 1469598103934665603
    for c in raw:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.fixture(scope="module")
def dump_tool(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("h5") / "h5dump")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-DH5LITE_WITH_ZLIB", "-o", exe,
                    os.path.join(HERE, "h5lite_dump.cpp"), "-l:libstdc++.so.6", "-lz"], check=True)

    def run(path):
        r = subprocess.run([exe, path], capture_output=True, text=True)
        assert r.returncode == 0, (r.returncode, r.stderr)
        out = {}
        for line in r.stdout.splitlines():
            kind, name, dtype, dims, nbytes, digest = line.split()
            dims = tuple(int(x) for x in dims[1:-1].split(",") if x)
            out[(kind, name)] = (int(dtype), dims, int(nbytes), int(digest, 16))
        return out
    return run


def example():
    rng = np.random.default_rng(7)
    nb = 20
    datasets = {
        "s_bias": rng.random(2).astype(np.float32),
        "config_count": rng.integers(0, 1 << 40, (7, 2)).astype(np.uint64),
        "dist": rng.random((1000, 12)),
        "config/int": rng.integers(0, 4, 5000).astype(np.uint64),
        "unique_config/pos": rng.random((1, nb, 3)),
        "unique_config/vel": rng.random((1, nb, 3)),
        "unique_config/config": np.array([3], dtype=np.uint64),
        "never_written": np.zeros((0, 5)),
        "mt": " ".join(str(x) for x in rng.integers(0, 2 ** 32, 625)),
    }
    for k in range(21):  # more links than one symbol-table node holds (8): three nodes
        datasets["many/d%02d" % k] = np.arange(k, dtype=np.uint32)
    attrs = {"sigma": np.float64(1.0), "length": np.float64(17.5), "nsteps": np.uint32(77),
             "nonlocal_bonds": np.array([[2, 6], [6, 10]], dtype=np.uint32),
             "permanent_bonds": np.zeros((0, 2), dtype=np.uint32)}
    return datasets, attrs


def test_python_reader_against_libhdf5_file():
    ds, attrs = h5lite.read(LIBHDF5_FILE)
    assert list(ds) == ["testdouble"] and attrs == {}
    assert ds["testdouble"].shape == (9, 1)
    assert np.array_equal(ds["testdouble"][:, 0], np.arange(9) * np.pi / 4)
    # the attribute libhdf5 attached to the dataset: pins the attribute-message layout
    # (name / datatype / dataspace each padded to 8, scalar dataspace, fixed string type)
    assert h5lite.read_attrs(LIBHDF5_FILE) == {"testdouble": {"MATLAB_class": b"double"}}
    st = h5lite.structure(LIBHDF5_FILE)
    # the conventions the writer copies: K values, cached root scratch pad, heap free list
    assert (st["leaf_k"], st["internal_k"], st["root_cache_type"]) == (4, 16, 1)
    assert st["root_scratch"] == st["root_stab"]
    assert st["heap_free_next"] == 1 and st["heap_free_offset"] + st["heap_free_len"] == st["heap_size"]
    assert st["root_nmsgs_declared"] == st["root_nmsgs_found"]


def test_cpp_reader_against_libhdf5_file(dump_tool):
    got = dump_tool(LIBHDF5_FILE)
    raw = (np.arange(9) * np.pi / 4).tobytes()
    assert got == {("dataset", "testdouble"): (0, (9, 1), 72, fnv1a(raw))}


def test_datatype_message_bytes_equal_libhdf5(tmp_path):
    """the IEEE-double datatype message libhdf5 wrote, found verbatim in our file"""
    with open(LIBHDF5_FILE, "rb") as f:
        ref = f.read()
    f64_msg = bytes.fromhex("11203f0008000000000040 00340b0034ff030000".replace(" ", ""))
    assert f64_msg in ref
    path = str(tmp_path / "a.h5")
    output.write_hdf5(path, {"x": np.arange(3.0)})
    with open(path, "rb") as f:
        mine = f.read()
    assert f64_msg in mine
    # ... and the fill-value message (header + body) of its contiguous dataset
    fill_msg = bytes.fromhex("0500080001000000" "0102020100000000")
    assert fill_msg in ref and fill_msg in mine
    assert mine[:8] == h5lite.SIGNATURE and struct.unpack_from("<Q", mine, 40)[0] == len(mine)


def test_writer_round_trip_both_readers(tmp_path, dump_tool):
    datasets, attrs = example()
    path = str(tmp_path / "out.h5")
    output.write_hdf5(path, datasets, attrs, fsync=True)
    ds, at = h5lite.read(path)
    assert set(ds) == set(datasets) and set(at) == set(attrs)
    for k, v in datasets.items():
        if isinstance(v, str):
            assert ds[k] == v
        else:
            assert ds[k].dtype == v.dtype and ds[k].shape == v.shape and np.array_equal(ds[k], v), k
    for k, v in attrs.items():
        assert np.asarray(at[k]).dtype == v.dtype and np.shape(at[k]) == np.shape(v)
        assert np.array_equal(at[k], v), k
    st = h5lite.structure(path)
    assert st["base"] == 0 and st["eof"] == st["file_size"] == os.path.getsize(path)
    assert (st["leaf_k"], st["internal_k"], st["root_cache_type"]) == (4, 16, 1)
    assert st["root_scratch"] == st["root_stab"]
    assert st["heap_free_next"] == 1 and st["heap_free_offset"] + st["heap_free_len"] == st["heap_size"]
    assert st["root_nmsgs_declared"] == st["root_nmsgs_found"] == 1 + len(attrs)
    codes = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.uint64): 2,
             np.dtype(np.uint32): 3}
    got = dump_tool(path)
    assert len(got) == len(datasets) + len(attrs)
    for k, v in datasets.items():
        if isinstance(v, str):
            assert got[("dataset", k)] == (4, (), len(v), fnv1a(v.encode()))
        else:
            assert got[("dataset", k)] == (codes[v.dtype], v.shape, v.nbytes, fnv1a(v.tobytes())), k
    for k, v in attrs.items():
        assert got[("attr", k)] == (codes[v.dtype], np.shape(v), v.nbytes, fnv1a(v.tobytes())), k


def test_file_view_matches_h5py_access_pattern(tmp_path):
    datasets, attrs = example()
    path = str(tmp_path / "out.h5")
    output.write_hdf5(path, datasets, attrs)
    with output.File(path) as f:
        assert f["s_bias"][0] == datasets["s_bias"][0]
        assert np.array_equal(f["unique_config"]["pos"], datasets["unique_config/pos"])
        assert np.array_equal(f["config"]["int"][:], datasets["config/int"])
        assert f.attrs["nonlocal_bonds"].tolist() == [[2, 6], [6, 10]]
        assert sorted(f["many"]) == ["d%02d" % k for k in range(21)]


def test_symbol_nodes_sorted_and_keys_consistent(tmp_path):
    """libhdf5 finds a link by binary search over the B-tree keys and inside the node, so
    names must ascend across the whole group and key[i+1] must name the last entry of child i"""
    datasets, _ = example()
    path = str(tmp_path / "out.h5")
    output.write_hdf5(path, datasets)
    with open(path, "rb") as f:
        d = f.read()
    p = h5lite._Parser(d)
    ds_msgs = p.messages(p.root_header)
    stab = [m for m in ds_msgs if m[0] == 0x11][0][2]
    # descend into group "many"
    names_seen = []

    def group_nodes(btree, heap):
        heap_data = p.q(heap + 24)
        assert d[btree:btree + 4] == b"TREE" and d[btree + 5] == 0
        used = p.h(btree + 6)
        assert p.q(btree + 8) == p.q(btree + 16) == h5lite.UNDEF  # no siblings
        assert p.q(btree + 24) == 0  # key[0] = offset of ""
        out = []
        for k in range(used):
            child, key = p.q(btree + 32 + 16 * k), p.q(btree + 40 + 16 * k)
            n = p.h(child + 6)
            assert 1 <= n <= 8
            ents = []
            for e in range(n):
                off = heap_data + p.q(child + 8 + 40 * e)
                ents.append((d[off:d.index(b"\0", off)].decode(), child + 8 + 40 * e))
            assert heap_data + key == heap_data + p.q(ents[-1][1])  # key names the last entry
            out.extend(ents)
        return out
    root = group_nodes(p.q(stab), p.q(stab + 8))
    names = [n for n, _ in root]
    assert names == sorted(names) and "many" in names
    ent = dict(root)["many"]
    assert p.i(ent + 16) == 1  # groups cache their B-tree / heap in the scratch pad
    many = group_nodes(p.q(ent + 24), p.q(ent + 32))
    names_seen = [n for n, _ in many]
    assert names_seen == ["d%02d" % k for k in range(21)]


def test_writer_errors(tmp_path):
    from hybridmc_b200.engine import EngineError
    with pytest.raises(EngineError, match="cannot write"):
        output.write_hdf5(str(tmp_path / "no_such_dir" / "x.h5"), {"a": np.arange(2.0)})
    import ctypes as C
    from hybridmc_b200.engine import load_library
    lib = load_library()
    h = C.c_void_p()
    assert lib.hmc_h5_create(C.byref(h)) == 0
    lib.hmc_h5_dataset.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_uint64]
    dims = (C.c_uint64 * 1)(3)
    buf = (C.c_double * 2)()
    assert lib.hmc_h5_dataset(h, b"x", 0, 1, dims, buf, 16) < 0  # 3 doubles announced, 2 given
    assert b"does not match" in lib.hmc_last_error(None)
    assert lib.hmc_h5_dataset(h, b"x", 9, 1, dims, buf, 16) < 0
    lib.hmc_h5_destroy.argtypes = [C.c_void_p]
    lib.hmc_h5_destroy(h)


def _chunked_file(datasets):
    """A minimal HDF5 file assembled by hand from the format specification, with the layout
    the stock reference executable uses for its extensible datasets (src/writer.h): dataspace
    with an unlimited first dimension, layout v3 CHUNKED, version-1 chunk B-tree, optional
    deflate filter.  datasets: {name: (array, chunk_rows, deflate)}; flat names only."""
    import zlib
    buf = bytearray(96)  # superblock, patched at the end
    UNDEF = 0xFFFFFFFFFFFFFFFF

    def align():
        buf.extend(b"\0" * (-len(buf) % 8))

    def msg(mtype, body, flags=0):
        body = body + b"\0" * (-len(body) % 8)
        return struct.pack("<HHBBH", mtype, len(body), flags, 0, 0) + body

    headers = {}
    for name, (arr, crows, deflate) in datasets.items():
        rank = arr.ndim
        cshape = (crows,) + arr.shape[1:]
        # chunks: partial edge chunks are stored at full size
        keys = []
        for r0 in range(0, arr.shape[0], crows):
            chunk = np.zeros(cshape, dtype=arr.dtype)
            part = arr[r0:r0 + crows]
            chunk[:len(part)] = part
            raw = chunk.tobytes()
            if deflate:
                raw = zlib.compress(raw, 6)
            align()
            keys.append((len(raw), 0, (r0,) + (0,) * (rank - 1) + (0,), len(buf)))
            buf.extend(raw)
        align()
        btree = len(buf)
        node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(keys), UNDEF, UNDEF)
        for size, mask, offs, addr in keys:
            node += struct.pack("<II", size, mask) + struct.pack("<%dQ" % (rank + 1), *offs)
            node += struct.pack("<Q", addr)
        node += struct.pack("<II", 0, 0) + struct.pack("<%dQ" % (rank + 1), arr.shape[0], *([0] * rank))
        buf.extend(node)
        space = struct.pack("<BBBBI", 1, rank, 1, 0, 0) + struct.pack("<%dQ" % rank, *arr.shape) + \
            struct.pack("<%dQ" % rank, UNDEF, *arr.shape[1:])
        if arr.dtype == np.float64:
            dtype = bytes.fromhex("11203f000800000000004000340b0034ff030000")
        else:
            dtype = struct.pack("<BBBBIHH", 0x10, 0, 0, 0, arr.itemsize, 0, 8 * arr.itemsize)
        msgs = msg(0x0001, space) + msg(0x0003, dtype, 1) + msg(0x0005, bytes([2, 3, 2, 1, 0, 0, 0, 0]), 1)
        n = 4
        if deflate:
            msgs += msg(0x000B, struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, 8, 1, 1) +
                        b"deflate\0" + struct.pack("<II", 6, 0))
            n += 1
        msgs += msg(0x0008, struct.pack("<BBBQ", 3, 2, rank + 1, btree) +
                    struct.pack("<%dI" % (rank + 1), *cshape, arr.itemsize))
        align()
        headers[name] = len(buf)
        buf.extend(struct.pack("<BBHII4x", 1, 0, n, 1, len(msgs)) + msgs)
    # root group: heap, one symbol node, B-tree, object header
    names = sorted(headers)
    heap_data = bytearray(8)
    offs = []
    for nm in names:
        offs.append(len(heap_data))
        heap_data.extend(nm.encode() + b"\0")
        heap_data.extend(b"\0" * (-len(heap_data) % 8))
    free = len(heap_data)
    heap_data.extend(struct.pack("<QQ", 1, 16))
    align()
    heap = len(buf)
    buf.extend(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free, heap + 32) + heap_data)
    align()
    snod = len(buf)
    node = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for nm, off in zip(names, offs):
        node += struct.pack("<QQII16x", off, headers[nm], 0, 0)
    buf.extend(node + b"\0" * (328 - len(node)))
    btree = len(buf)
    node = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod, offs[-1])
    buf.extend(node + b"\0" * (544 - len(node)))
    root = len(buf)
    m = msg(0x0011, struct.pack("<QQ", btree, heap))
    buf.extend(struct.pack("<BBHII4x", 1, 0, 1, 1, len(m)) + m)
    sb = SIG + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0) + \
        struct.pack("<QQQQ", 0, UNDEF, len(buf), UNDEF) + struct.pack("<QQII", 0, root, 1, 0) + \
        struct.pack("<QQ", btree, heap)
    assert len(sb) == 96
    buf[:96] = sb
    return bytes(buf)


SIG = h5lite.SIGNATURE


def test_reads_chunked_deflated_datasets_like_the_stock_executable_writes(tmp_path, dump_tool):
    """(format per specification; NOT pinned against libhdf5 — no chunked file exists here)"""
    rng = np.random.default_rng(5)
    dist = rng.random((23, 4))           # chunk (10, 4): two full chunks + a partial one
    cfg = rng.integers(0, 4, 2500).astype(np.uint64)  # chunk 1024, unfiltered
    path = tmp_path / "stock_like.h5"
    path.write_bytes(_chunked_file({"dist": (dist, 10, True), "cfg": (cfg, 1024, False)}))
    ds, attrs = h5lite.read(str(path))
    assert attrs == {} and sorted(ds) == ["cfg", "dist"]
    assert ds["dist"].dtype == np.float64 and np.array_equal(ds["dist"], dist)
    assert ds["cfg"].dtype == np.uint64 and np.array_equal(ds["cfg"], cfg)
    with output.File(str(path)) as f:
        assert np.array_equal(f["dist"][:], dist)
    # the C++ reader the executable uses for --input-file
    got = dump_tool(str(path))
    assert got[("dataset", "dist")] == (0, dist.shape, dist.nbytes, fnv1a(dist.tobytes()))
    assert got[("dataset", "cfg")] == (2, cfg.shape, cfg.nbytes, fnv1a(cfg.tobytes()))
    pos = rng.random((3, 5, 3))  # unique_config/pos-like: 3-D, chunk (2, 5, 3), deflate
    p3 = tmp_path / "pos.h5"
    p3.write_bytes(_chunked_file({"pos": (pos, 2, True)}))
    assert np.array_equal(h5lite.read(str(p3))[0]["pos"], pos)
    assert dump_tool(str(p3))[("dataset", "pos")] == (0, pos.shape, pos.nbytes, fnv1a(pos.tobytes()))


def test_cpp_reader_survives_corrupted_files(tmp_path):
    """the executable reads --input-file / --snapshot-file with this parser: truncated and
    bit-flipped copies of a libhdf5-written file and of one of our own files must end in a clean
    error (or parse), never in a memory error — checked under AddressSanitizer + UBSan"""
    import random
    exe = str(tmp_path / "h5dump_asan")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined",
                        "-fno-omit-frame-pointer", "-DH5LITE_WITH_ZLIB", "-o", exe,
                        os.path.join(HERE, "h5lite_dump.cpp"), "-lz"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("no sanitizer runtime in this toolchain")
    from hybridmc_b200.output import write_hdf5
    own = str(tmp_path / "own.h5")
    write_hdf5(own, {"s_bias": np.array([3.5, 0.0], np.float32),
                     "config_count": np.arange(8, dtype=np.uint64).reshape(4, 2),
                     "dist": np.random.default_rng(0).random((50, 3)),
                     "unique_config/pos": np.zeros((1, 8, 3)), "config/int": np.arange(30, dtype=np.uint64)},
               {"sigma": np.float64(1.0), "nonlocal_bonds": np.array([[2, 6]], np.uint32)})
    rng = random.Random(5)
    for src_path in (own, os.path.join(HERE, "golden", "libhdf5_written_matlab73.mat")):
        src = open(src_path, "rb").read()
        for k in range(60):
            b = bytearray(src)
            if k % 3 == 0:
                b = b[:rng.randrange(8, len(b))]
            else:
                for _ in range(rng.randrange(1, 8)):
                    b[rng.randrange(min(len(b), 3000))] ^= 1 << rng.randrange(8)
            path = str(tmp_path / "fz.h5")
            open(path, "wb").write(b)
            res = subprocess.run([exe, path], capture_output=True, timeout=60)
            err = res.stderr.decode("latin1")
            assert "AddressSanitizer" not in err and "runtime error" not in err, err[:800]
            assert res.returncode in (0, 1, 2), (res.returncode, err[:400])
