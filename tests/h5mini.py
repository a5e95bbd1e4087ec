"""Minimal reader for the reference's golden HDF5 fixtures (TEST INFRASTRUCTURE: h5py is not part of the image).

The files under ``tests/test_data`` of the reference are tiny (8-34 KB), written by h5py with the oldest format:
superblock version 0, old-style groups (``TREE`` / ``SNOD`` / ``HEAP``), version-1 object headers, contiguous or
compact little-endian float / integer datasets, no filters.  That subset of the HDF5 file format specification is
what is implemented here; anything else raises ``NotImplementedError``.

    with open(path, "rb") as f: datasets = read_datasets(f.read())      # {"group/name": numpy array}
"""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class _File:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        if buf[8] != 0:
            raise NotImplementedError(f"superblock version {buf[8]}")
        self.O, self.L = buf[13], buf[14]
        if (self.O, self.L) != (8, 8):
            raise NotImplementedError("offsets / lengths are not 8 bytes")
        p = 24  # signature 8, versions 4, sizes 2 + reserved, K values 4, flags 4
        self.base, _free, _eof, _drv = struct.unpack_from("<4Q", buf, p)
        p += 32
        self.root = self._symbol_entry(p)

    def _symbol_entry(self, p):
        name_off, hdr, cache, _res = struct.unpack_from("<QQII", self.b, p)
        scratch = self.b[p + 24:p + 40]
        return dict(name_off=name_off, header=hdr, cache=cache, scratch=scratch)

    # ---- groups -------------------------------------------------------------------------------------------------
    def _heap_name(self, heap_addr, off):
        assert self.b[heap_addr:heap_addr + 4] == b"HEAP"
        data_addr = struct.unpack_from("<Q", self.b, heap_addr + 24)[0]
        s = data_addr + off
        e = self.b.index(b"\x00", s)
        return self.b[s:e].decode()

    def _btree_leaves(self, addr):
        assert self.b[addr:addr + 4] == b"TREE", "group B-tree expected"
        node_type, level, used = struct.unpack_from("<BBH", self.b, addr + 4)
        if node_type != 0:
            raise NotImplementedError("non-group B-tree")
        p = addr + 8 + 16  # signature + type/level/entries + two sibling addresses
        children = []
        for _ in range(used):
            p += 8  # key
            children.append(struct.unpack_from("<Q", self.b, p)[0])
            p += 8
        for c in children:
            if level > 0:
                yield from self._btree_leaves(c)
            else:
                yield c

    def _group_entries(self, btree, heap):
        for snod in self._btree_leaves(btree):
            assert self.b[snod:snod + 4] == b"SNOD"
            n = struct.unpack_from("<H", self.b, snod + 6)[0]
            for i in range(n):
                e = self._symbol_entry(snod + 8 + 40 * i)
                yield self._heap_name(heap, e["name_off"]), e

    # ---- object headers -----------------------------------------------------------------------------------------
    def _messages(self, addr):
        ver, _r, nmsg, _refs, size = struct.unpack_from("<BBHII", self.b, addr)
        if ver != 1:
            raise NotImplementedError(f"object header version {ver}")
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", self.b, p)
                data = self.b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:  # continuation
                    off, ln = struct.unpack_from("<QQ", data, 0)
                    blocks.append((off, ln))
                out.append((mtype, data))
        return out

    def _dataset(self, msgs):
        shape = dtype = raw = None
        for mtype, d in msgs:
            if mtype == 0x01:
                ver, rank, flags = d[0], d[1], d[2]
                p = 8 if ver == 1 else 4
                shape = struct.unpack_from(f"<{rank}Q", d, p) if rank else ()
            elif mtype == 0x03:
                cls = d[0] & 0x0F
                size = struct.unpack_from("<I", d, 4)[0]
                if d[1] & 1:
                    raise NotImplementedError("big-endian data")
                if cls == 1:
                    dtype = {4: "<f4", 8: "<f8"}[size]
                elif cls == 0:
                    dtype = ("<i" if d[1] & 8 else "<u") + str(size)
                else:
                    dtype = None  # strings etc.: skipped
            elif mtype == 0x08:
                ver = d[0]
                if ver == 3:
                    lclass = d[1]
                    if lclass == 1:
                        a, n = struct.unpack_from("<QQ", d, 2)
                        raw = None if a == UNDEF else self.b[a:a + n]
                    elif lclass == 0:
                        n = struct.unpack_from("<H", d, 2)[0]
                        raw = d[4:4 + n]
                    else:
                        raise NotImplementedError("chunked dataset")
                elif ver in (1, 2):
                    rank, lclass = d[1], d[2]
                    if lclass != 1:
                        raise NotImplementedError("layout v1/v2 other than contiguous")
                    a = struct.unpack_from("<Q", d, 8)[0]
                    raw = ("addr", a)
                else:
                    raise NotImplementedError(f"layout version {ver}")
        if shape is None or dtype is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        itemsize = np.dtype(dtype).itemsize
        if isinstance(raw, tuple):
            raw = self.b[raw[1]:raw[1] + n * itemsize]
        if raw is None:
            return np.zeros(shape, dtype)
        return np.frombuffer(raw[:n * itemsize], dtype=dtype).reshape(shape).copy()

    def walk(self, entry=None, prefix=""):
        entry = entry or self.root
        msgs = self._messages(entry["header"])
        stab = [d for t, d in msgs if t == 0x11]
        if stab or entry["cache"] == 1:
            if stab:
                btree, heap = struct.unpack_from("<QQ", stab[0], 0)
            else:
                btree, heap = struct.unpack_from("<QQ", entry["scratch"], 0)
            for name, e in self._group_entries(btree, heap):
                yield from self.walk(e, f"{prefix}{name}/")
            return
        arr = self._dataset(msgs)
        if arr is not None:
            yield prefix.rstrip("/"), arr


def read_datasets(buf: bytes) -> dict:
    return dict(_File(buf).walk())
