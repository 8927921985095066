"""Minimal pure-Python HDF5 reader for UEDGE save files (no h5py in the image).

Supports exactly what the reference's ``hdf5_save`` files use
(src/uedge/hdf5.py:83-187): superblock v0, old-style groups (B-tree v1 + SNOD
symbol nodes + local heap), object headers v1 with continuation blocks,
simple dataspaces, fixed-point / IEEE-float datatypes and contiguous or
compact, unfiltered layouts.  Chunked data raises NotImplementedError.

``read_h5(path)`` returns ``{"group/name": ndarray}`` for every dataset.
"""
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class _H5:
    def __init__(self, buf):
        self.b = buf
        if buf[:8] != _SIG:
            raise ValueError("not an HDF5 file")
        ver = buf[8]
        if ver not in (0, 1):
            raise NotImplementedError("superblock version %d" % ver)
        self.so = buf[13]  # size of offsets
        self.sl = buf[14]  # size of lengths
        p = 24 if ver == 0 else 28
        self.base = self._off(p)
        p += 4 * self.so  # base, free-space, eof, driver-info
        # root symbol table entry
        self.root = self._symentry(p)

    def _uint(self, p, n):
        return int.from_bytes(self.b[p : p + n], "little")

    def _off(self, p):
        return self._uint(p, self.so)

    def _len(self, p):
        return self._uint(p, self.sl)

    def _symentry(self, p):
        so = self.so
        name_off = self._off(p)
        ohdr = self._off(p + so)
        cache = self._uint(p + 2 * so, 4)
        scratch = p + 2 * so + 8
        ent = {"name_off": name_off, "ohdr": ohdr, "cache": cache}
        if cache == 1:
            ent["btree"] = self._off(scratch)
            ent["heap"] = self._off(scratch + so)
        return ent

    # ---- object header ---------------------------------------------------
    def _messages(self, addr):
        b = self.b
        ver = b[addr]
        if ver != 1:
            raise NotImplementedError("object header version %d" % ver)
        nmsg = self._uint(addr + 2, 2)
        hsize = self._uint(addr + 8, 4)
        blocks = [(addr + 16, hsize)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(msgs) < nmsg:
                mtype = self._uint(p, 2)
                msize = self._uint(p + 2, 2)
                data = p + 8
                msgs.append((mtype, data, msize))
                if mtype == 0x10:
                    blocks.append((self._off(data), self._len(data + self.so)))
                p = data + msize
        return msgs

    def _group_children(self, btree, heap):
        b = self.b
        if b[heap : heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        hdata = self._off(heap + 8 + 2 * self.sl)
        out = []

        def name_at(off):
            s = hdata + off
            e = b.index(b"\0", s)
            return b[s:e].decode()

        def walk(node):
            if b[node : node + 4] == b"TREE":
                level = b[node + 5]
                nent = self._uint(node + 6, 2)
                p = node + 8 + 2 * self.so
                p += self.sl  # key 0
                for _ in range(nent):
                    child = self._off(p)
                    p += self.so + self.sl
                    walk(child)
            elif b[node : node + 4] == b"SNOD":
                nsym = self._uint(node + 6, 2)
                p = node + 8
                esz = 2 * self.so + 4 + 4 + 16
                for i in range(nsym):
                    ent = self._symentry(p + i * esz)
                    out.append((name_at(ent["name_off"]), ent))
            else:
                raise ValueError("bad group node")

        walk(btree)
        return out

    def _dataset(self, msgs):
        b = self.b
        shape = None
        dtype = None
        layout = None
        for mtype, p, n in msgs:
            if mtype == 0x01:
                ver, rank = b[p], b[p + 1]
                q = p + (8 if ver == 1 else 4)
                shape = tuple(self._len(q + i * self.sl) for i in range(rank))
            elif mtype == 0x03:
                cls = b[p] & 0x0F
                bits0 = b[p + 1]
                size = self._uint(p + 4, 4)
                endian = ">" if (bits0 & 1) else "<"
                if cls == 0:
                    signed = (bits0 >> 3) & 1
                    dtype = np.dtype("%s%s%d" % (endian, "i" if signed else "u", size))
                elif cls == 1:
                    dtype = np.dtype("%sf%d" % (endian, size))
                elif cls == 3:
                    dtype = np.dtype("S%d" % size)
                else:
                    dtype = None
            elif mtype == 0x08:
                ver = b[p]
                if ver == 3:
                    lclass = b[p + 1]
                    if lclass == 1:
                        layout = ("contig", self._off(p + 2), self._len(p + 2 + self.so))
                    elif lclass == 0:
                        sz = self._uint(p + 2, 2)
                        layout = ("compact", p + 4, sz)
                    else:
                        layout = ("chunked",)
                else:
                    rank = b[p + 1]
                    lclass = b[p + 2]
                    q = p + 8
                    if lclass == 1:
                        layout = ("contig", self._off(q), None)
                    elif lclass == 0:
                        q += 4 * rank
                        sz = self._uint(q, 4)
                        layout = ("compact", q + 4, sz)
                    else:
                        layout = ("chunked",)
        if shape is None or dtype is None or layout is None:
            return None
        if layout[0] == "chunked":
            raise NotImplementedError("chunked dataset")
        count = int(np.prod(shape)) if shape else 1
        addr = layout[1]
        if layout[0] == "contig":
            if addr == (1 << (8 * self.so)) - 1:
                return np.zeros(shape, dtype=dtype)
            addr += self.base
        arr = np.frombuffer(self.b, dtype=dtype, count=count, offset=addr)
        return arr.reshape(shape).copy()

    def walk(self):
        out = {}

        def visit(prefix, ent):
            msgs = self._messages(ent["ohdr"])
            bt = hp = None
            if ent.get("cache") == 1:
                bt, hp = ent["btree"], ent["heap"]
            else:
                for mtype, p, n in msgs:
                    if mtype == 0x11:
                        bt, hp = self._off(p), self._off(p + self.so)
            if bt is not None:
                for name, child in self._group_children(bt, hp):
                    visit(prefix + [name], child)
            else:
                ds = self._dataset(msgs)
                if ds is not None:
                    out["/".join(prefix)] = ds

        visit([], self.root)
        return out


def read_h5(path):
    with open(path, "rb") as f:
        buf = f.read()
    return _H5(buf).walk()
