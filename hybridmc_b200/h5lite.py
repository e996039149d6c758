"""Pure-Python reader for the HDF5 subset the engine's writer emits
(hybridmc_b200/csrc/host/h5lite.h; the image has neither libhdf5 nor h5py) — written
independently of the C++ code, from the HDF5 File Format Specification, so that each
checks the other, and pinned against a libhdf5-written file (tests/test_h5lite.py).

Understands: superblock v0/v1 (8-byte offsets, optional user block), version-1 object
headers with continuation blocks, symbol-table groups (v1 B-tree of any depth, local
heap), dataspace v1/v2, fixed-point / IEEE float / fixed string / variable-length string
datatypes, layout v1-v3 CONTIGUOUS or COMPACT, attribute messages v1, and — for the files
the stock reference executable writes (src/writer.h: extensible datasets) — layout v3
CHUNKED through the version-1 chunk B-tree with the deflate / shuffle / fletcher32 filters
(filter pipeline message v1/v2).  The chunked path follows the specification but could not
be pinned against a libhdf5-written chunked file here (none exists in the image; the test
builds one by hand), unlike everything else in this module.
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


def is_hdf5(path):
    with open(path, "rb") as f:
        return f.read(8) == SIGNATURE


class _Parser:
    def __init__(self, data):
        self.d = data
        sb = 0
        while data[sb:sb + 8] != SIGNATURE:
            sb = 512 if sb == 0 else sb * 2
            if sb + 8 > len(data):
                raise ValueError("not an HDF5 file")
        ver = data[sb + 8]
        if ver > 1:
            raise NotImplementedError("superblock version %d" % ver)
        if data[sb + 13] != 8 or data[sb + 14] != 8:
            raise NotImplementedError("only 8-byte offsets and lengths")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", data, sb + 16)
        p = sb + 24 + (4 if ver == 1 else 0)
        self.base, _free, self.eof, _drv = struct.unpack_from("<4Q", data, p)
        self.root_entry = p + 32
        self.root_header = self.base + self.q(self.root_entry + 8)

    def q(self, at):
        return struct.unpack_from("<Q", self.d, at)[0]

    def h(self, at):
        return struct.unpack_from("<H", self.d, at)[0]

    def i(self, at):
        return struct.unpack_from("<I", self.d, at)[0]

    def messages(self, header):
        """[(type, flags, body offset, body size)] of a version-1 object header."""
        d = self.d
        if d[header] != 1:
            raise NotImplementedError("object header version %d" % d[header])
        nmsgs = self.h(header + 2)
        blocks = [(header + 16, self.i(header + 8))]
        out = []
        count = 0
        k = 0
        while k < len(blocks):
            p, size = blocks[k]
            end = p + size
            while p + 8 <= end and count < nmsgs:
                mtype, msize, flags = self.h(p), self.h(p + 2), d[p + 4]
                body = p + 8
                count += 1
                if mtype == 0x0010:
                    blocks.append((self.base + self.q(body), self.q(body + 8)))
                else:
                    out.append((mtype, flags, body, msize))
                p = body + msize
            k += 1
        return out

    def datatype(self, at):
        """-> (kind, numpy dtype or None, element size)"""
        d = self.d
        cls, ver = d[at] & 15, d[at] >> 4
        if ver not in (1, 2, 3):
            raise NotImplementedError("datatype version %d" % ver)
        bits0 = d[at + 1]
        size = self.i(at + 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            signed = bool(bits0 & 8)
            return "num", np.dtype("%s%s%d" % (order, "i" if signed else "u", size)), size
        if cls == 1:
            return "num", np.dtype("%sf%d" % (order, size)), size
        if cls == 3:
            return "str", np.dtype("S%d" % size), size
        if cls == 9:
            if bits0 & 15 != 1:
                raise NotImplementedError("variable-length sequences")
            return "vstr", None, size
        raise NotImplementedError("datatype class %d" % cls)

    def dataspace(self, at):
        ver, rank = self.d[at], self.d[at + 1]
        p = at + (8 if ver == 1 else 4)
        return tuple(struct.unpack_from("<%dQ" % rank, self.d, p))

    def vstr(self, at):
        length, gcol, index = self.i(at), self.base + self.q(at + 4), self.i(at + 12)
        if self.d[gcol:gcol + 4] != b"GCOL":
            raise ValueError("bad global heap collection")
        end = gcol + self.q(gcol + 8)
        p = gcol + 16
        while p + 16 <= end:
            idx, size = self.h(p), self.q(p + 8)
            if idx == index:
                return self.d[p + 16:p + 16 + length].decode()
            if idx == 0:
                break
            p += 16 + ((size + 7) & ~7)
        raise ValueError("global heap object %d not found" % index)

    def value(self, kind, dt, esize, shape, at):
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if kind == "vstr":
            vals = [self.vstr(at + 16 * k) for k in range(n)]
            return vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)
        arr = np.frombuffer(self.d, dtype=dt, count=n, offset=at) if n else np.empty(0, dt)
        if kind == "str":
            arr = np.char.rstrip(arr, b"\0")
        return arr.reshape(shape).copy() if shape else arr[0]

    def attribute(self, body):
        if self.d[body] != 1:
            raise NotImplementedError("attribute message version %d" % self.d[body])
        nl, tl, sl = self.h(body + 2), self.h(body + 4), self.h(body + 6)
        p = body + 8
        name = self.d[p:p + nl].split(b"\0")[0].decode()
        p += (nl + 7) & ~7
        kind, dt, esize = self.datatype(p)
        p += (tl + 7) & ~7
        shape = self.dataspace(p)
        p += (sl + 7) & ~7
        return name, self.value(kind, dt, esize, shape, p)

    def dataset(self, msgs, path):
        kind = dt = esize = shape = None
        layout = None
        for mtype, _flags, body, _size in msgs:
            if mtype == 0x0003:
                kind, dt, esize = self.datatype(body)
            elif mtype == 0x0001:
                shape = self.dataspace(body)
            elif mtype == 0x0008:
                layout = body
        if kind is None or shape is None or layout is None:
            raise ValueError("incomplete dataset %s" % path)
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        d = self.d
        ver = d[layout]
        if ver == 3:
            cls = d[layout + 1]
            if cls == 1:
                addr = self.q(layout + 2)
                if n == 0 or addr == UNDEF:
                    return np.zeros(shape, dtype=dt) if dt is not None else None
                return self.value(kind, dt, esize, shape, self.base + addr)
            if cls == 0:
                return self.value(kind, dt, esize, shape, layout + 4)
        elif ver in (1, 2):
            cls = d[layout + 2]
            if cls == 1:
                addr = self.q(layout + 8)
                if n == 0 or addr == UNDEF:
                    return np.zeros(shape, dtype=dt)
                return self.value(kind, dt, esize, shape, self.base + addr)
            if cls == 0:
                rank = d[layout + 1]
                return self.value(kind, dt, esize, shape, layout + 8 + 4 * rank + 4)
        else:
            raise NotImplementedError("layout message version %d" % ver)
        if ver == 3 and d[layout + 1] == 2:
            if kind == "vstr":
                raise NotImplementedError("chunked variable-length strings")
            return self.chunked(layout, msgs, dt, esize, shape)
        raise NotImplementedError("dataset %s: layout not supported" % path)

    def filters(self, msgs):
        """[(filter id, client data)] of the filter pipeline message (IV.A.2.l), in
        application order"""
        out = []
        for mtype, _f, body, _s in msgs:
            if mtype != 0x000B:
                continue
            ver, n = self.d[body], self.d[body + 1]
            p = body + (8 if ver == 1 else 2)
            for _ in range(n):
                fid = self.h(p)
                if ver == 1 or fid >= 256:
                    nlen = self.h(p + 2)
                    p += 4
                else:
                    nlen = 0
                    p += 2
                _flags, ncd = self.h(p), self.h(p + 2)
                p += 4
                p += (nlen + 7) & ~7 if ver == 1 else nlen
                cd = struct.unpack_from("<%dI" % ncd, self.d, p)
                p += 4 * ncd
                if ver == 1 and ncd % 2:
                    p += 4
                out.append((fid, cd))
        return out

    def chunked(self, layout, msgs, dt, esize, shape):
        """layout v3 class 2: rank+1, B-tree address, chunk dims (the last one = element size)"""
        rank1 = self.d[layout + 2]
        btree = self.base + self.q(layout + 3)
        cdims = struct.unpack_from("<%dI" % rank1, self.d, layout + 11)
        rank = rank1 - 1
        if rank != len(shape) or cdims[-1] != esize:
            raise ValueError("inconsistent chunked layout")
        cshape = tuple(cdims[:rank])
        filters = self.filters(msgs)
        out = np.zeros(shape, dtype=dt)
        if self.q(layout + 3) == UNDEF or 0 in shape:
            return out
        for offs, size, mask, addr in self.chunks(btree, rank1):
            raw = self.d[addr:addr + size]
            for k in range(len(filters) - 1, -1, -1):  # undo the pipeline back to front
                if mask & (1 << k):
                    continue
                fid, cd = filters[k]
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:  # shuffle: byte planes -> elements
                    n = len(raw) // esize
                    raw = np.frombuffer(raw[:n * esize], np.uint8).reshape(esize, n).T.tobytes()
                elif fid == 3:  # fletcher32: 4-byte checksum appended
                    raw = raw[:-4]
                else:
                    raise NotImplementedError("filter %d" % fid)
            chunk = np.frombuffer(raw, dtype=dt, count=int(np.prod(cshape))).reshape(cshape)
            sel_out = tuple(slice(o, min(o + c, s_)) for o, c, s_ in zip(offs, cshape, shape))
            sel_in = tuple(slice(0, s.stop - s.start) for s in sel_out)
            if all(s.stop > s.start for s in sel_out):
                out[sel_out] = chunk[sel_in]
        return out

    def chunks(self, node, rank1):
        """leaves of the version-1 chunk B-tree (III.A.1, node type 1):
        key = chunk size, filter mask, rank+1 offsets; keys and child pointers alternate"""
        d = self.d
        if d[node:node + 4] != b"TREE" or d[node + 4] != 1:
            raise ValueError("bad chunk B-tree node")
        level, used = d[node + 5], self.h(node + 6)
        ksize = 8 + 8 * rank1
        p = node + 24
        for _ in range(used):
            size, mask = self.i(p), self.i(p + 4)
            offs = struct.unpack_from("<%dQ" % rank1, d, p + 8)[:-1]
            child = self.base + self.q(p + ksize)
            if level:
                yield from self.chunks(child, rank1)
            else:
                yield offs, size, mask, child
            p += ksize + 8

    def walk(self, header, path, datasets, attrs):
        msgs = self.messages(header)
        stab = [m for m in msgs if m[0] == 0x0011]
        for mtype, _f, body, _s in msgs:
            if mtype == 0x000C:
                name, val = self.attribute(body)
                attrs.setdefault(path, {})[name] = val
        if not stab:
            datasets[path] = self.dataset(msgs, path)
            return
        body = stab[0][2]
        btree, heap = self.base + self.q(body), self.base + self.q(body + 8)
        if self.d[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        heap_data = self.base + self.q(heap + 24)
        self.btree(btree, heap_data, path + "/" if path else "", datasets, attrs)

    def btree(self, node, heap_data, prefix, datasets, attrs):
        d = self.d
        if d[node:node + 4] != b"TREE" or d[node + 4] != 0:
            raise ValueError("bad group B-tree node")
        level, used = d[node + 5], self.h(node + 6)
        for k in range(used):
            child = self.base + self.q(node + 24 + 8 + 16 * k)
            if level:
                self.btree(child, heap_data, prefix, datasets, attrs)
                continue
            if d[child:child + 4] != b"SNOD":
                raise ValueError("bad symbol table node")
            for e in range(self.h(child + 6)):
                ent = child + 8 + 40 * e
                off = heap_data + self.q(ent)
                name = d[off:d.index(b"\0", off)].decode()
                self.walk(self.base + self.q(ent + 8), prefix + name, datasets, attrs)


def read(path):
    """-> (datasets {"group/name": ndarray | scalar | str}, attrs {name: value} of the root)"""
    with open(path, "rb") as f:
        data = f.read()
    p = _Parser(data)
    datasets, attrs = {}, {}
    p.walk(p.root_header, "", datasets, attrs)
    return datasets, attrs.get("", {})


def read_attrs(path):
    """-> {object path ("" = root): {attribute name: value}} for every group and dataset"""
    with open(path, "rb") as f:
        data = f.read()
    p = _Parser(data)
    datasets, attrs = {}, {}
    p.walk(p.root_header, "", datasets, attrs)
    return attrs


def structure(path):
    """Structural facts the writer's tests pin against libhdf5's conventions."""
    with open(path, "rb") as f:
        data = f.read()
    p = _Parser(data)
    msgs = p.messages(p.root_header)
    stab = [m for m in msgs if m[0] == 0x0011][0][2]
    heap = p.base + p.q(stab + 8)
    heap_data = p.base + p.q(heap + 24)
    free_off = p.q(heap + 16)
    return dict(base=p.base, eof=p.eof, file_size=len(data), leaf_k=p.leaf_k,
                internal_k=p.internal_k, root_cache_type=p.i(p.root_entry + 16),
                root_scratch=(p.q(p.root_entry + 24), p.q(p.root_entry + 32)),
                root_stab=(p.q(stab), p.q(stab + 8)),
                heap_size=p.q(heap + 8), heap_free_offset=free_off,
                heap_free_next=p.q(heap_data + free_off), heap_free_len=p.q(heap_data + free_off + 8),
                root_nmsgs_declared=p.h(p.root_header + 2), root_nmsgs_found=len(msgs))
