#!/usr/bin/env python
"""Minimal pure-python reader of the HDF5 subset ws_export_hdf5 writes (superblock v0, object headers v1,
symbol-table groups, contiguous datasets, v1 attributes), written from the HDF5 File Format Specification
independently of the writer -- the repo's check that the exported file is well formed where no libhdf5/h5py
exists.  With h5py available `python tools/h5min.py file.h5 --h5py` cross-checks every object against it.

    python tools/h5min.py map.h5            # list groups, datasets (shape, dtype), attributes
"""
import struct
import sys

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


class File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        if b[:8] != SIG:
            raise H5Error("bad signature")
        (sb_ver, fs_ver, root_ver, _r0, shm_ver, so, sl, _r1, self.leaf_k, self.int_k, flags) = struct.unpack_from("<8B2HI", b, 8)
        if (sb_ver, fs_ver, root_ver, shm_ver) != (0, 0, 0, 0) or (so, sl) != (8, 8):
            raise H5Error("unsupported superblock")
        base, fsaddr, self.eof, drv = struct.unpack_from("<4Q", b, 24)
        if base != 0 or fsaddr != UNDEF or drv != UNDEF or self.eof != len(b):
            raise H5Error("superblock addresses: base %d eof %d len %d" % (base, self.eof, len(b)))
        self.root = self._symbol_entry(56)

    # ---- low level ----
    def _symbol_entry(self, off):
        name_off, hdr, cache, _res = struct.unpack_from("<QQII", self.b, off)
        bt, heap = struct.unpack_from("<QQ", self.b, off + 24)
        return {"name_off": name_off, "header": hdr, "cache": cache, "btree": bt, "heap": heap}

    def _heap_name(self, heap_addr, off):
        b = self.b
        if b[heap_addr:heap_addr + 4] != b"HEAP" or b[heap_addr + 4] != 0:
            raise H5Error("bad local heap")
        size, free, data = struct.unpack_from("<QQQ", b, heap_addr + 8)
        if data + size > len(b) or off >= size:
            raise H5Error("heap bounds")
        # the free list must stay inside the data segment (libhdf5 checks this when it loads a heap)
        f = free
        while f != 1:
            if f % 8 or f + 16 > size:
                raise H5Error("heap free list")
            nxt, sz = struct.unpack_from("<QQ", b, data + f)
            if sz < 16 or f + sz > size:
                raise H5Error("heap free block")
            f = nxt
        end = b.index(b"\0", data + off)
        return b[data + off:end].decode()

    def _messages(self, addr):
        b = self.b
        if addr % 8:
            raise H5Error("object header alignment")
        ver, _r, nmsg, refcnt, size = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise H5Error("object header version %d" % ver)
        p, end, out = addr + 16, addr + 16 + size, []
        while p < end:
            mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
            if msize % 8:
                raise H5Error("message size not a multiple of 8")
            out.append((mtype, b[p + 8:p + 8 + msize]))
            p += 8 + msize
        if p != end or len(out) != nmsg:
            raise H5Error("object header: %d messages declared, %d found" % (nmsg, len(out)))
        return out

    @staticmethod
    def _dtype(d):
        cls, ver = d[0] & 0x0F, d[0] >> 4
        size = struct.unpack_from("<I", d, 4)[0]
        if ver != 1:
            raise H5Error("datatype version")
        if cls == 0:
            off, prec = struct.unpack_from("<HH", d, 8)
            if d[1] & 1 or off != 0 or prec != 8 * size:
                raise H5Error("fixed-point layout")
            return np.dtype("<%s%d" % ("i" if d[1] & 8 else "u", size)), 12
        if cls == 1:
            off, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", d, 8)
            if (size, off, prec, eloc, esize, mloc, msize, bias, d[1], d[2]) != (4, 0, 32, 23, 8, 0, 23, 127, 0x20, 0x1F):
                raise H5Error("not IEEE float32 LE")
            return np.dtype("<f4"), 20
        raise H5Error("datatype class %d" % cls)

    @staticmethod
    def _dataspace(d):
        ver, rank, flags = d[0], d[1], d[2]
        if ver != 1 or flags != 0:
            raise H5Error("dataspace")
        return tuple(struct.unpack_from("<%dQ" % rank, d, 8)), 8 + 8 * rank

    # ---- objects ----
    def _group_links(self, btree, heap):
        b, out = self.b, []

        def walk(addr, lo_key):
            if b[addr:addr + 4] != b"TREE" or b[addr + 4] != 0:
                raise H5Error("bad B-tree node")
            level, used = b[addr + 5], struct.unpack_from("<H", b, addr + 6)[0]
            if used > 2 * self.int_k:
                raise H5Error("B-tree node overfull")
            p = addr + 24
            key = struct.unpack_from("<Q", b, p)[0]
            for i in range(used):
                child, nkey = struct.unpack_from("<QQ", b, p + 8 + 16 * i)
                if level > 0:
                    walk(child, key)
                else:
                    if b[child:child + 4] != b"SNOD" or b[child + 4] != 1:
                        raise H5Error("bad symbol table node")
                    n = struct.unpack_from("<H", b, child + 6)[0]
                    if n > 2 * self.leaf_k:
                        raise H5Error("symbol table node overfull")
                    names = []
                    for k in range(n):
                        e = self._symbol_entry(child + 8 + 40 * k)
                        e["name"] = self._heap_name(heap, e["name_off"])
                        names.append(e["name"])
                        out.append(e)
                    lo, hi = self._heap_name(heap, key), self._heap_name(heap, nkey)
                    if names != sorted(names) or not all(lo < x <= hi for x in names):
                        raise H5Error("names outside their B-tree keys")
                key = nkey

        walk(btree, None)
        names = [e["name"] for e in out]
        if names != sorted(names):
            raise H5Error("group links not sorted")
        return out

    def open(self, path):
        node = dict(self.root)
        for part in [p for p in path.split("/") if p]:
            obj = self.read_object(node["header"])
            if obj["kind"] != "group":
                raise H5Error("%s is not a group" % part)
            hit = [e for e in obj["links"] if e["name"] == part]
            if not hit:
                raise KeyError(path)
            node = hit[0]
        return self.read_object(node["header"])

    def read_object(self, addr):
        msgs = self._messages(addr)
        attrs, shape, dtype, layout, stab = {}, None, None, None, None
        for t, d in msgs:
            if t == 0x0011:
                stab = struct.unpack_from("<QQ", d, 0)
            elif t == 0x0001:
                shape, _ = self._dataspace(d)
            elif t == 0x0003:
                dtype, _ = self._dtype(d)
            elif t == 0x0008:
                if d[0] != 3 or d[1] != 1:
                    raise H5Error("layout")
                layout = struct.unpack_from("<QQ", d, 2)
            elif t == 0x000C:
                ver, _r, nsz, tsz, ssz = struct.unpack_from("<BBHHH", d, 0)
                if ver != 1:
                    raise H5Error("attribute version")
                pad = lambda n: (n + 7) & ~7
                p = 8
                name = d[p:p + nsz].rstrip(b"\0").decode(); p += pad(nsz)
                adt, used = self._dtype(d[p:p + pad(tsz)])
                if used != tsz:
                    raise H5Error("attribute datatype size")
                p += pad(tsz)
                ashape, used = self._dataspace(d[p:p + pad(ssz)])
                if used != ssz:
                    raise H5Error("attribute dataspace size")
                p += pad(ssz)
                n = int(np.prod(ashape)) if ashape else 1
                val = np.frombuffer(d, adt, n, p)
                attrs[name] = val[0] if not ashape else val.reshape(ashape)
            elif t in (0x0005, 0x0000, 0x0012):
                pass
            else:
                raise H5Error("unexpected message 0x%04x" % t)
        if stab is not None:
            return {"kind": "group", "links": self._group_links(*stab), "attrs": attrs}
        if None in (shape, dtype, layout):
            raise H5Error("incomplete dataset header")
        addr, nbytes = layout
        if nbytes != int(np.prod(shape)) * dtype.itemsize or addr + nbytes > len(self.b):
            raise H5Error("dataset extent")
        return {"kind": "dataset", "data": np.frombuffer(self.b, dtype, int(np.prod(shape)), addr).reshape(shape), "attrs": attrs}

    def walk(self, path="/"):
        obj = self.open(path)
        yield path, obj
        if obj["kind"] == "group":
            for e in obj["links"]:
                yield from self.walk(path.rstrip("/") + "/" + e["name"])


def main():
    f = File(sys.argv[1])
    use_h5py = "--h5py" in sys.argv
    h = None
    if use_h5py:
        import h5py
        h = h5py.File(sys.argv[1], "r")
    for path, obj in f.walk():
        if obj["kind"] == "group":
            print("%-28s group  %d links  attrs %s" % (path, len(obj["links"]), {k: v.item() for k, v in obj["attrs"].items()}))
            if h is not None and path != "/":
                assert sorted(h[path].keys()) == [e["name"] for e in obj["links"]]
                assert {k: v for k, v in h[path].attrs.items()} == {k: v.item() for k, v in obj["attrs"].items()}
        else:
            d = obj["data"]
            print("%-28s dataset %s %s" % (path, d.shape, d.dtype))
            if h is not None:
                assert np.array_equal(h[path][...], d) and h[path].dtype == d.dtype
    if h is not None:
        print("h5py agrees")


if __name__ == "__main__":
    main()
