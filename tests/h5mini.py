"""Minimal HDF5 reader - TEST INFRASTRUCTURE for host/pinc_h5.c (tests/test_h5_writer.py).

Written from the HDF5 File Format Specification, independently of the writer, and pinned to a file produced by the real
HDF5 library (scipy's MATLAB v7.3 fixture, see test_h5_writer.py::test_reader_on_a_file_written_by_libhdf5), because neither
libhdf5 nor h5py exist in this image.  It understands what libhdf5 1.8 emits with default settings:
superblock v0/v1 (optionally behind a user block), version-1 object headers with continuation blocks, old-style groups
(Symbol Table message -> v1 B-tree -> symbol-table nodes + local heap), Dataspace v1/v2, Datatype classes 0 (integer),
1 (floating point), 3 (string) and 7 (reference), Data Layout v3 compact / contiguous / chunked-without-filters (v1 B-tree,
node type 1), Attribute messages v1-v3.
"""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


class Dataset:
    def __init__(self, shape, dtype, data, attrs, layout):
        self.shape, self.dtype, self.data, self.attrs, self.layout = shape, dtype, data, attrs, layout


class Group:
    def __init__(self, links, attrs):
        self.links, self.attrs = links, attrs       # links: name -> Group | Dataset, in the order the file stores them


class File:
    def __init__(self, path):
        self.buf = open(path, "rb").read()
        off = 0
        while self.buf[off:off + 8] != SIG:
            off = 512 if off == 0 else off * 2
            if off >= len(self.buf):
                raise H5Error("no HDF5 signature")
        self.sb = off
        b = self.buf
        ver = b[off + 8]
        if ver not in (0, 1):
            raise H5Error("superblock version %d" % ver)
        self.so, self.sl = b[off + 13], b[off + 14]
        if (self.so, self.sl) != (8, 8):
            raise H5Error("only 8-byte offsets/lengths")
        self.leaf_k, self.int_k = struct.unpack_from("<HH", b, off + 16)
        p = off + 24 + (4 if ver == 1 else 0)
        self.base, self.freespace, self.eof, self.driver = struct.unpack_from("<QQQQ", b, p)
        p += 32
        name_off, oh, cache, _res = struct.unpack_from("<QQII", b, p)
        self.root_cache = (cache, struct.unpack_from("<QQ", b, p + 24))
        self.names_checked = 0
        self.root = self._object(oh)

    # ---- low level -----------------------------------------------------------------------------------------------
    def _at(self, addr):
        return self.base + addr

    def _messages(self, addr):
        b, p = self.buf, self._at(addr)
        if b[p] != 1:
            raise H5Error("object header version %d at %d" % (b[p], addr))
        nmsg, refc, size = struct.unpack_from("<HII", b, p + 2)
        blocks = [(p + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            q, left = blocks.pop(0)
            end = q + left
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, q)
                data = b[q + 8:q + 8 + msize]
                if mtype == 0x0010:                      # continuation
                    caddr, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((self._at(caddr), clen))
                out.append((mtype, flags, data))
                q += 8 + msize
        return out

    def _heap(self, addr):
        b, p = self.buf, self._at(addr)
        if b[p:p + 4] != b"HEAP":
            raise H5Error("bad local heap signature")
        size, free, seg = struct.unpack_from("<QQQ", b, p + 8)
        if free != 1 and free >= size:
            raise H5Error("bad heap free list (libhdf5 refuses this file)")
        return self._at(seg), size

    def _name(self, heap, off):
        seg, size = heap
        if off >= size:
            raise H5Error("name offset outside the heap")
        e = self.buf.index(b"\0", seg + off)
        return self.buf[seg + off:e].decode()

    def _btree_group(self, addr, heap, out, lo_key=None):
        """Depth-first walk of a group B-tree; checks signatures, key order and the bracketing of every child's names."""
        b, p = self.buf, self._at(addr)
        if b[p:p + 4] != b"TREE" or b[p + 4] != 0:
            raise H5Error("bad group B-tree node")
        level, used = b[p + 5], struct.unpack_from("<H", b, p + 6)[0]
        if used > 2 * self.int_k:
            raise H5Error("B-tree node over-full")
        q = p + 24
        keys = [struct.unpack_from("<Q", b, q + 16 * i)[0] for i in range(used + 1)]
        kids = [struct.unpack_from("<Q", b, q + 16 * i + 8)[0] for i in range(used)]
        for i, kid in enumerate(kids):
            lo, hi = self._name(heap, keys[i]), self._name(heap, keys[i + 1])
            if level > 0:
                self._btree_group(kid, heap, out)
                continue
            s = self._at(kid)
            if b[s:s + 4] != b"SNOD" or b[s + 4] != 1:
                raise H5Error("bad symbol-table node")
            n = struct.unpack_from("<H", b, s + 6)[0]
            if n > 2 * self.leaf_k:
                raise H5Error("symbol-table node over-full")
            prev = None
            for j in range(n):
                noff, oh, cache, _r = struct.unpack_from("<QQII", b, s + 8 + 40 * j)
                name = self._name(heap, noff)
                if not (lo < name <= hi) or (prev is not None and not prev < name):       # libhdf5 finds names by these keys
                    raise H5Error("name %r outside its B-tree keys (%r, %r] or out of order" % (name, lo, hi))
                prev = name
                self.names_checked += 1
                out.append((name, oh))

    def _dataspace(self, d):
        ver, rank, flags = d[0], d[1], d[2]
        p = 8 if ver == 1 else 4
        dims = struct.unpack_from("<%dQ" % rank, d, p)
        maxd = struct.unpack_from("<%dQ" % rank, d, p + 8 * rank) if flags & 1 else None
        return tuple(dims), maxd

    def _datatype(self, d):
        cls, ver = d[0] & 0x0F, d[0] >> 4
        bits = d[1] | (d[2] << 8) | (d[3] << 16)
        size = struct.unpack_from("<I", d, 4)[0]
        order = ">" if bits & 1 else "<"
        if cls == 0:
            return np.dtype("%s%s%d" % (order, "i" if bits & 8 else "u", size)), 8 + 4
        if cls == 1:
            boff, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", d, 8)
            ieee = {4: (32, 23, 8, 0, 23, 127), 8: (64, 52, 11, 0, 52, 1023)}
            if ieee.get(size) != (prec, eloc, esize, mloc, msize, bias) or boff != 0 or ((bits >> 8) & 0xFF) != 8 * size - 1 or ((bits >> 4) & 3) != 2:
                raise H5Error("floating point type is not IEEE 754 binary%d" % (8 * size))
            return np.dtype("%sf%d" % (order, size)), 8 + 12
        if cls == 3:
            return np.dtype("S%d" % size), 8
        if cls == 7:
            return np.dtype("V%d" % size), 8
        raise H5Error("datatype class %d" % cls)

    def _attribute(self, d):
        ver = d[0]
        if ver == 1:
            nlen, tlen, slen = struct.unpack_from("<HHH", d, 2)
            p = 8
            pad = lambda n: (n + 7) & ~7
        else:
            nlen, tlen, slen = struct.unpack_from("<HHH", d, 2)
            p = 8 + (1 if ver == 3 else 0)
            pad = lambda n: n
        name = d[p:p + nlen].split(b"\0")[0].decode(); p += pad(nlen)
        dt, _ = self._datatype(d[p:p + tlen]); p += pad(tlen)
        shape, _ = self._dataspace(d[p:p + slen]) if slen >= 8 or (slen and d[p] == 2) else ((), None); p += pad(slen)
        n = int(np.prod(shape)) if shape else 1
        val = np.frombuffer(d, dtype=dt, count=n, offset=p).reshape(shape)
        return name, val

    def _chunked(self, addr, shape, chunk, dt):
        """Layout class 2 without filters: v1 B-tree (node type 1), keys = chunk size, filter mask, offsets."""
        out = np.zeros(shape, dtype=dt)
        rank = len(shape)

        def walk(a):
            b, p = self.buf, self._at(a)
            if b[p:p + 4] != b"TREE" or b[p + 4] != 1:
                raise H5Error("bad chunk B-tree node")
            level, used = b[p + 5], struct.unpack_from("<H", b, p + 6)[0]
            q = p + 24
            ksz = 8 + 8 * (rank + 1)
            for i in range(used):
                csize, mask = struct.unpack_from("<II", b, q)
                offs = struct.unpack_from("<%dQ" % (rank + 1), b, q + 8)
                child = struct.unpack_from("<Q", b, q + ksz)[0]
                q += ksz + 8
                if level > 0:
                    walk(child)
                    continue
                if mask:
                    raise H5Error("filtered chunk")
                c = np.frombuffer(b, dtype=dt, count=int(np.prod(chunk)), offset=self._at(child)).reshape(chunk)
                sl = tuple(slice(o, min(o + cs, s)) for o, cs, s in zip(offs[:rank], chunk, shape))
                out[sl] = c[tuple(slice(0, s.stop - s.start) for s in sl)]
        if addr != UNDEF:
            walk(addr)
        return out

    def _object(self, addr):
        msgs = self._messages(addr)
        attrs = dict(self._attribute(d) for t, _f, d in msgs if t == 0x000C)
        kinds = {t for t, _f, _d in msgs}
        if 0x0011 in kinds:
            d = next(d for t, _f, d in msgs if t == 0x0011)
            bt, hp = struct.unpack_from("<QQ", d, 0)
            heap = self._heap(hp)
            if self._name(heap, 0) != "":
                raise H5Error("heap offset 0 is not the empty name")
            entries = []
            self._btree_group(bt, heap, entries)
            return Group({name: self._object(oh) for name, oh in entries}, attrs)
        if 0x0008 in kinds:
            shape, maxd = self._dataspace(next(d for t, _f, d in msgs if t == 0x0001))
            dt, _ = self._datatype(next(d for t, _f, d in msgs if t == 0x0003))
            lay = next(d for t, _f, d in msgs if t == 0x0008)
            n = int(np.prod(shape)) if shape else 1
            if lay[0] in (1, 2):
                # versions 1 and 2 (files of older libraries): dimensionality, class, 5 reserved bytes, [address], 4-byte sizes
                rank1, cls = lay[1], lay[2]
                p = 8
                a = UNDEF
                if cls != 0:
                    a = struct.unpack_from("<Q", lay, p)[0]; p += 8
                sizes = struct.unpack_from("<%dI" % rank1, lay, p); p += 4 * rank1
                if cls == 1:
                    data = np.frombuffer(self.buf, dtype=dt, count=n, offset=self._at(a)).reshape(shape) if n else np.zeros(shape, dt)
                    return Dataset(shape, dt, data, attrs, "contiguous")
                if cls == 2:
                    return Dataset(shape, dt, self._chunked(a, shape, tuple(sizes[:-1]), dt), attrs, "chunked")
                csize = struct.unpack_from("<I", lay, p)[0]
                return Dataset(shape, dt, np.frombuffer(lay, dtype=dt, count=n, offset=p + 4).reshape(shape), attrs, "compact")
            if lay[0] != 3:
                raise H5Error("layout version %d" % lay[0])
            if lay[1] == 0:
                size = struct.unpack_from("<H", lay, 2)[0]
                data = np.frombuffer(lay, dtype=dt, count=n, offset=4).reshape(shape)
                kind = "compact"
            elif lay[1] == 1:
                a, size = struct.unpack_from("<QQ", lay, 2)
                if n and size != n * dt.itemsize:
                    raise H5Error("contiguous size %d != %d elements" % (size, n))
                if n and self._at(a) + size > len(self.buf):
                    raise H5Error("raw data beyond the end of the file")
                data = np.frombuffer(self.buf, dtype=dt, count=n, offset=self._at(a)).reshape(shape) if n else np.zeros(shape, dt)
                kind = "contiguous"
            else:
                rank1 = lay[2]
                a = struct.unpack_from("<Q", lay, 3)[0]
                cdims = struct.unpack_from("<%dI" % rank1, lay, 11)
                data = self._chunked(a, shape, tuple(cdims[:-1]), dt)
                kind = "chunked"
            return Dataset(shape, dt, data, attrs, kind)
        raise H5Error("object at %d is neither an old-style group nor a dataset (messages %s)" % (addr, sorted(kinds)))

    def get(self, path):
        o = self.root
        for part in [p for p in path.split("/") if p]:
            o = o.links[part]
        return o
