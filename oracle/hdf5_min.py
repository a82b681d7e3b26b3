"""Minimal read-only HDF5 parser (TEST INFRASTRUCTURE - not part of the product path).

h5py is not available in the build container, and the reference's golden regression
files (/root/reference/tests/references/*.hdf5) are classic HDF5 1.8-style files:
superblock v0, v1 object headers, symbol-table groups, layout-v3 contiguous or
chunked (+fletcher32) datasets.  This module reads exactly that subset so that
tests/golden/make_golden.py can turn the goldens into .npz fixtures.

It follows the public HDF5 file-format specification (v1.1 "Disk Format: Level 0-2"),
nothing from the reference repo.
"""
import struct
import numpy as np


class H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        b = self.b
        assert b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        assert b[8] == 0, "only superblock v0 supported"
        assert b[13] == 8 and b[14] == 8, "only 8-byte offsets/lengths supported"
        # v0 superblock: 24 bytes fixed + 4 addresses, then the root symbol-table entry
        root_entry = 24 + 4 * 8
        self.root = self._read_group_entry(root_entry)

    # ------------------------------------------------------------------ low level
    def _u(self, off, n):
        return int.from_bytes(self.b[off:off + n], "little")

    def _read_group_entry(self, off):
        """symbol table entry -> dict(name_off, ohdr, cache, btree, heap)"""
        name_off = self._u(off, 8)
        ohdr = self._u(off + 8, 8)
        cache = self._u(off + 16, 4)
        btree = heap = None
        if cache == 1:
            btree = self._u(off + 24, 8)
            heap = self._u(off + 32, 8)
        return dict(name_off=name_off, ohdr=ohdr, cache=cache, btree=btree, heap=heap)

    def _messages(self, ohdr):
        """iterate (type, flags, bytes) over a v1 object header incl. continuations"""
        b = self.b
        assert b[ohdr] == 1, "only v1 object headers supported (got %d)" % b[ohdr]
        nmsg = self._u(ohdr + 2, 2)
        hsize = self._u(ohdr + 8, 4)
        blocks = [(ohdr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype = self._u(pos, 2)
                msize = self._u(pos + 2, 2)
                mflags = b[pos + 4]
                data = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:  # continuation
                    blocks.append((int.from_bytes(data[:8], "little"), int.from_bytes(data[8:16], "little")))
                out.append((mtype, mflags, data))
        return out

    def _heap_str(self, heap, off):
        assert self.b[heap:heap + 4] == b"HEAP"
        dseg = self._u(heap + 24, 8)
        s = dseg + off
        e = self.b.index(b"\x00", s)
        return self.b[s:e].decode()

    def _group_children(self, btree, heap):
        """walk a v1 group B-tree -> {name: symbol table entry}"""
        res = {}
        b = self.b
        assert b[btree:btree + 4] == b"TREE", "bad group btree"
        level = b[btree + 5]
        nent = self._u(btree + 6, 2)
        pos = btree + 24
        children = []
        for i in range(nent):
            pos += 8  # key
            children.append(self._u(pos, 8))
            pos += 8
        for c in children:
            if level > 0:
                res.update(self._group_children(c, heap))
            else:
                assert b[c:c + 4] == b"SNOD"
                nsym = self._u(c + 6, 2)
                for s in range(nsym):
                    e = self._read_group_entry(c + 8 + 40 * s)
                    res[self._heap_str(heap, e["name_off"])] = e
        return res

    # ------------------------------------------------------------------ public
    def _entry_for(self, path):
        e = self.root
        for part in [p for p in path.split("/") if p]:
            bt, hp = e["btree"], e["heap"]
            if bt is None:
                for mtype, _, data in self._messages(e["ohdr"]):
                    if mtype == 0x11:
                        bt = int.from_bytes(data[:8], "little")
                        hp = int.from_bytes(data[8:16], "little")
            if bt is None:
                raise KeyError(path)
            ch = self._group_children(bt, hp)
            if part not in ch:
                raise KeyError(path)
            e = ch[part]
        return e

    def keys(self, path="/"):
        e = self._entry_for(path)
        bt, hp = e["btree"], e["heap"]
        if bt is None:
            for mtype, _, data in self._messages(e["ohdr"]):
                if mtype == 0x11:
                    bt = int.from_bytes(data[:8], "little")
                    hp = int.from_bytes(data[8:16], "little")
        if bt is None:
            return []
        return sorted(self._group_children(bt, hp).keys())

    def is_group(self, path):
        e = self._entry_for(path)
        if e["btree"] is not None:
            return True
        return any(m[0] == 0x11 for m in self._messages(e["ohdr"]))

    @staticmethod
    def _dtype(data):
        cls = data[0] & 0x0F
        bits0 = data[1]
        size = int.from_bytes(data[4:8], "little")
        order = ">" if (bits0 & 1) else "<"
        if cls == 0:
            signed = bool(bits0 & 0x08)
            return np.dtype("%s%s%d" % (order, "i" if signed else "u", size))
        if cls == 1:
            return np.dtype("%sf%d" % (order, size))
        raise NotImplementedError("datatype class %d" % cls)

    @staticmethod
    def _dims(data):
        ver, rank, flags = data[0], data[1], data[2]
        off = 8 if ver == 1 else 4
        return tuple(int.from_bytes(data[off + 8 * i: off + 8 * i + 8], "little") for i in range(rank))

    def attrs(self, path):
        e = self._entry_for(path)
        res = {}
        for mtype, _, d in self._messages(e["ohdr"]):
            if mtype != 0x0C:
                continue
            ver = d[0]
            assert ver == 1, "attribute message v%d not supported" % ver
            nsz, tsz, ssz = (int.from_bytes(d[2:4], "little"), int.from_bytes(d[4:6], "little"),
                             int.from_bytes(d[6:8], "little"))
            pad = lambda x: (x + 7) // 8 * 8
            p = 8
            name = d[p:p + nsz].split(b"\x00")[0].decode()
            p += pad(nsz)
            tmsg = d[p:p + tsz]
            p += pad(tsz)
            smsg = d[p:p + ssz]
            p += pad(ssz)
            try:
                dt = self._dtype(tmsg)
            except NotImplementedError:
                cls = tmsg[0] & 0x0F
                if cls == 3:  # fixed-length string
                    size = int.from_bytes(tmsg[4:8], "little")
                    res[name] = d[p:p + size].split(b"\x00")[0].decode(errors="replace")
                continue
            dims = self._dims(smsg) if len(smsg) >= 4 and smsg[1] > 0 else ()
            cnt = int(np.prod(dims)) if dims else 1
            arr = np.frombuffer(d, dtype=dt, count=cnt, offset=p)
            res[name] = arr.reshape(dims) if dims else arr[0]
        return res

    def dataset(self, path):
        e = self._entry_for(path)
        dt = dims = None
        layout = None
        filters = []
        for mtype, _, d in self._messages(e["ohdr"]):
            if mtype == 0x01:
                dims = self._dims(d)
            elif mtype == 0x03:
                dt = self._dtype(d)
            elif mtype == 0x08:
                assert d[0] == 3, "only layout v3 supported"
                cls = d[1]
                if cls == 1:
                    layout = ("contiguous", int.from_bytes(d[2:10], "little"), int.from_bytes(d[10:18], "little"))
                elif cls == 2:
                    rank = d[2]
                    bt = int.from_bytes(d[3:11], "little")
                    cd = tuple(int.from_bytes(d[11 + 4 * i: 15 + 4 * i], "little") for i in range(rank))
                    layout = ("chunked", bt, cd)
                elif cls == 0:
                    sz = int.from_bytes(d[2:4], "little")
                    layout = ("compact", d[4:4 + sz])
            elif mtype == 0x0B:
                ver = d[0]
                nf = d[1]
                p = 8 if ver == 1 else 2
                for _ in range(nf):
                    fid = int.from_bytes(d[p:p + 2], "little")
                    if ver == 1 or fid >= 256:
                        nlen = int.from_bytes(d[p + 2:p + 4], "little")
                    else:
                        nlen = 0
                    ncv = int.from_bytes(d[p + 6:p + 8], "little")
                    p += 8
                    if ver == 1:
                        p += (nlen + 7) // 8 * 8
                    else:
                        p += nlen
                    p += 4 * ncv
                    if ver == 1 and ncv % 2:
                        p += 4
                    filters.append(fid)
        assert dt is not None and dims is not None and layout is not None, path
        n = int(np.prod(dims))
        if layout[0] == "contiguous":
            return np.frombuffer(self.b, dtype=dt, count=n, offset=layout[1]).reshape(dims).copy()
        if layout[0] == "compact":
            return np.frombuffer(layout[1], dtype=dt, count=n).reshape(dims).copy()
        # chunked: walk the v1 raw-data B-tree
        out = np.zeros(dims, dtype=dt)
        _, bt, cd = layout
        rank = len(dims)
        cshape = cd[:rank]
        for f in filters:
            assert f == 3, "only fletcher32 filter supported (got id %d)" % f

        def walk(node):
            b = self.b
            assert b[node:node + 4] == b"TREE" and b[node + 4] == 1
            level = b[node + 5]
            nent = self._u(node + 6, 2)
            pos = node + 24
            ksz = 8 + 8 * (rank + 1)
            for i in range(nent):
                csize = self._u(pos, 4)
                offs = tuple(self._u(pos + 8 + 8 * j, 8) for j in range(rank))
                child = self._u(pos + ksz, 8)
                pos += ksz + 8
                if level > 0:
                    walk(child)
                else:
                    raw = b[child:child + csize]
                    if 3 in filters:
                        raw = raw[:-4]
                    chunk = np.frombuffer(raw, dtype=dt, count=int(np.prod(cshape))).reshape(cshape)
                    sl = tuple(slice(o, min(o + c, d)) for o, c, d in zip(offs, cshape, dims))
                    out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        walk(bt)
        return out

    def walk(self, path="/"):
        """yield (path, 'group'|'dataset')"""
        for k in self.keys(path):
            p = path.rstrip("/") + "/" + k
            if self.is_group(p):
                yield p, "group"
                yield from self.walk(p)
            else:
                yield p, "dataset"
