"""Minimal HDF5 writer / reader for the optimiser checkpoints (src/mpi_optimize.jl:26-28 writes
`h5write(loc * "iter_$k.h5", "data", x)`; scripts/post_rect.jl:31-43 reads them back with `h5read(file, "data")`).

h5py / libhdf5 are not available in this image, so this module writes the file format directly, restricted to what a
checkpoint needs: ONE file = superblock version 0, a root group in the classic layout (version-1 object header with a
symbol-table message, version-1 B-tree with one leaf, one symbol-table node, local heap) and little-endian
IEEE float64 / int32 / int64 datasets with contiguous storage.  That is the layout libhdf5 itself produces with default
settings (library format "earliest"), which every HDF5 release reads.

The reader understands the same subset plus user blocks, several datasets per group, nested groups, version-1/2
dataspaces, compact storage and 32-bit floats; tests/test_capi_cpu.py walks a file written by the real library
(scipy's bundled MATLAB-7.3 test file) with it, so writer and reader are checked against libhdf5's own output and
against each other.  Arrays are stored in C order with their numpy shape (a Julia reader sees the dimensions reversed,
as with any HDF5 file written from a row-major language; the checkpoint vector is 1-D).
"""
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16          # library defaults (group leaf / internal node K)

_DTYPES = {   # numpy dtype -> datatype message body (class+version, 3 class-bit bytes, size, properties)
    "<f8": struct.pack("<B3BI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023),
    "<f4": struct.pack("<B3BI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127),
    "<i4": struct.pack("<B3BI", 0x10, 0x08, 0x00, 0x00, 4) + struct.pack("<HH", 0, 32),
    "<i8": struct.pack("<B3BI", 0x10, 0x08, 0x00, 0x00, 8) + struct.pack("<HH", 0, 64),
}


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _message(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages):
    data = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(data)) + data


def write_datasets(path, datasets):
    """datasets: {name: array}.  Writes one HDF5 file with these datasets in the root group."""
    names = sorted(datasets)                     # symbol-table entries are ordered by name
    if not names or len(names) > 2 * LEAF_K:
        raise ValueError("between 1 and %d datasets per file" % (2 * LEAF_K))
    arrays = {}
    for nm in names:
        a = np.ascontiguousarray(datasets[nm])
        key = a.dtype.newbyteorder("<").str if a.dtype.byteorder != "|" else a.dtype.str
        if key not in _DTYPES:
            raise TypeError("unsupported dtype %s" % a.dtype)
        arrays[nm] = (a.astype(key, copy=False), key)
    # local heap data segment: "" at offset 0 (the root's own name), then the dataset names, then one free block
    heap = bytearray(b"\0" * 8)
    name_off = {}
    for nm in names:
        name_off[nm] = len(heap)
        heap += _pad8(nm.encode() + b"\0")
    free_off = len(heap)
    heap += struct.pack("<QQ", 1, 16)            # last free block: next = 1 (H5HL_FREE_NULL), size 16
    # addresses
    pos = 96                                     # superblock
    root_oh = _object_header([_message(0x0011, struct.pack("<QQ", 0, 0))])      # patched below
    a_root = pos; pos += len(root_oh)
    btree_size = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
    a_btree = pos; pos += btree_size
    a_heap = pos; pos += 32
    a_heapdata = pos; pos += len(heap)
    snod_size = 8 + 2 * LEAF_K * 40
    a_snod = pos; pos += snod_size
    headers, a_oh, a_data = {}, {}, {}
    for nm in names:
        a, key = arrays[nm]
        space = struct.pack("<BBB5x", 1, a.ndim, 0) + b"".join(struct.pack("<Q", d) for d in a.shape)
        msgs = [_message(0x0001, space), _message(0x0003, _DTYPES[key], flags=1),
                _message(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0)),                  # fill value v2: late allocation, written if set, default value
                _message(0x0008, struct.pack("<BBQQ", 3, 1, 0, a.nbytes))]              # layout v3 contiguous, address patched below
        headers[nm] = msgs
        a_oh[nm] = pos
        pos += len(_object_header(msgs))
    for nm in names:
        a_data[nm] = pos
        pos += arrays[nm][0].nbytes + (-arrays[nm][0].nbytes % 8)
    eof = pos
    # assemble
    out = bytearray()
    out += SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    out += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    out += struct.pack("<QQII", 0, a_root, 1, 0) + struct.pack("<QQ", a_btree, a_heap)      # root symbol-table entry, cached
    assert len(out) == 96
    out += _object_header([_message(0x0011, struct.pack("<QQ", a_btree, a_heap))])
    bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, a_snod, name_off[names[-1]])
    out += bt + b"\0" * (btree_size - len(bt))
    out += b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, a_heapdata)
    out += heap
    sn = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for nm in names:
        sn += struct.pack("<QQII16x", name_off[nm], a_oh[nm], 0, 0)
    out += sn + b"\0" * (snod_size - len(sn))
    for nm in names:
        a, key = arrays[nm]
        msgs = headers[nm][:3] + [_message(0x0008, struct.pack("<BBQQ", 3, 1, a_data[nm], a.nbytes))]
        assert len(out) == a_oh[nm]
        out += _object_header(msgs)
    for nm in names:
        assert len(out) == a_data[nm]
        raw = arrays[nm][0].tobytes()
        out += raw + b"\0" * (-len(raw) % 8)
    assert len(out) == eof
    with open(path, "wb") as fh:
        fh.write(out)


def write_dataset(path, name, array):
    write_datasets(path, {name: array})


# ------------------------------------------------------------------------------------------ reader
class _File:
    def __init__(self, path):
        self.buf = open(path, "rb").read()
        base = 0
        while self.buf[base:base + 8] != SIGNATURE:        # a user block puts the superblock at 512, 1024, ...
            base = 512 if base == 0 else base * 2
            if base + 8 > len(self.buf):
                raise ValueError("not an HDF5 file")
        ver = self.buf[base + 8]
        if ver not in (0, 1):
            raise ValueError("superblock version %d not supported" % ver)
        so, sl = self.buf[base + 13], self.buf[base + 14]
        if (so, sl) != (8, 8):
            raise ValueError("only 8-byte offsets / lengths")
        self.leaf_k, self.int_k = struct.unpack_from("<HH", self.buf, base + 16)
        p = base + 24 + (4 if ver == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack_from("<QQQQ", self.buf, p)     # all addresses are relative to the base address
        self.root = self._entry(p + 32)

    def _entry(self, p):
        name_off, oh, ctype = struct.unpack_from("<QQI", self.buf, p)
        scratch = struct.unpack_from("<QQ", self.buf, p + 24)
        return {"name_off": name_off, "oh": oh, "ctype": ctype, "scratch": scratch}

    def messages(self, addr):
        """(type, flags, body) of a version-1 object header, following continuation messages."""
        p = self.base + addr
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", self.buf, p)
        if ver != 1:
            raise ValueError("object header version %d not supported" % ver)
        chunks, out = [(p + 16, size)], []
        while chunks and len(out) < nmsg:
            q, left = chunks.pop(0)
            end = q + left
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", self.buf, q)
                body = self.buf[q + 8:q + 8 + msize]
                q += 8 + msize
                if mtype == 0x0010:                           # continuation
                    off, ln = struct.unpack_from("<QQ", body)
                    chunks.append((self.base + off, ln))
                out.append((mtype, flags, body))
        return out

    def links(self, group_oh):
        """{name: object header address} of a classic group."""
        st = [b for t, _, b in self.messages(group_oh) if t == 0x0011]
        if not st:
            raise ValueError("not a classic (symbol-table) group")
        btree, heap = struct.unpack_from("<QQ", st[0])
        hp = self.base + heap
        if self.buf[hp:hp + 4] != b"HEAP":
            raise ValueError("bad local heap")
        hsize, _, hdata = struct.unpack_from("<QQQ", self.buf, hp + 8)
        seg = self.buf[self.base + hdata:self.base + hdata + hsize]
        out = {}

        def walk(addr):
            p = self.base + addr
            if self.buf[p:p + 4] == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", self.buf, p + 4)
                if ntype != 0:
                    raise ValueError("not a group B-tree")
                for e in range(used):
                    walk(struct.unpack_from("<Q", self.buf, p + 24 + 8 + e * 16)[0])
            elif self.buf[p:p + 4] == b"SNOD":
                nsym = struct.unpack_from("<H", self.buf, p + 6)[0]
                for e in range(nsym):
                    ent = self._entry(p + 8 + e * 40)
                    nm = seg[ent["name_off"]:seg.index(b"\0", ent["name_off"])].decode()
                    out[nm] = ent["oh"]
            else:
                raise ValueError("bad group node")
        walk(btree)
        return out

    def dataset(self, oh):
        shape = dtype = data = None
        for t, _, b in self.messages(oh):
            if t == 0x0001:
                ver, rank = b[0], b[1]
                shape = struct.unpack_from("<%dQ" % rank, b, 8 if ver == 1 else 4)
            elif t == 0x0003:
                cls, size = b[0] & 0x0F, struct.unpack_from("<I", b, 4)[0]
                if b[1] & 1:
                    raise ValueError("big-endian data not supported")
                if cls == 1: dtype = {8: "<f8", 4: "<f4"}[size]
                elif cls == 0: dtype = ("<i" if b[1] & 0x08 else "<u") + str(size)
                else: raise ValueError("datatype class %d not supported" % cls)
            elif t == 0x0008:
                ver, lclass = b[0], b[1]
                if ver in (1, 2):                              # HDF5 1.6 and older: rank + 1 four-byte sizes, the last the element size
                    nd, lclass = b[1], b[2]
                    if lclass != 1:
                        raise ValueError("only contiguous storage in version-%d layouts" % ver)
                    addr = struct.unpack_from("<Q", b, 8)[0]
                    size = int(np.prod(struct.unpack_from("<%dI" % nd, b, 16)))
                    data = self.buf[self.base + addr:self.base + addr + size]
                    continue
                if ver != 3:
                    raise ValueError("layout version %d not supported" % ver)
                if lclass == 1:
                    addr, size = struct.unpack_from("<QQ", b, 2)
                    data = self.buf[self.base + addr:self.base + addr + size] if addr != UNDEF else b""
                elif lclass == 0:
                    size = struct.unpack_from("<H", b, 2)[0]
                    data = b[4:4 + size]
                else:
                    raise ValueError("chunked storage not supported")
        if shape is None or dtype is None or data is None:
            raise ValueError("not a dataset")
        n = int(np.prod(shape)) if shape else 1
        return np.frombuffer(data, dtype=dtype, count=n).reshape(shape).copy()


def list_names(path, group="/"):
    f = _File(path)
    oh = f.root["oh"]
    for part in [p for p in group.split("/") if p]:
        oh = f.links(oh)[part]
    return sorted(f.links(oh))


def read_dataset(path, name):
    f = _File(path)
    parts = [p for p in name.split("/") if p]
    oh = f.root["oh"]
    for part in parts[:-1]:
        oh = f.links(oh)[part]
    return f.dataset(f.links(oh)[parts[-1]])
