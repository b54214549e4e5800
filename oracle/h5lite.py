"""TEST INFRASTRUCTURE ONLY -- minimal reader for the reference's HDF5 fixtures.

h5py is not installed.  The four files under ``/root/reference/tests/*.h5`` use
the oldest on-disk format only (superblock v0, v1 object headers, symbol-table
groups, contiguous little-endian float64 datasets, no filters), which this
~100-line parser walks.  Used by ``oracle/make_goldens.py`` to convert them to
``tests/golden/*.npz`` so the GPU box (no reference tree) can check against them.
"""
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class _File(object):
    def __init__(self, buf):
        self.b = buf
        assert buf[:8] == _SIG, "not an HDF5 file"
        assert buf[8] == 0, "only superblock v0 is supported"
        self.so, self.sl = buf[13], buf[14]   # size of offsets / lengths
        assert self.so == 8 and self.sl == 8
        # v0 superblock: 24 bytes header, then base/free/eof/driver addresses,
        # then the root group symbol table entry
        self.root_entry = 24 + 4*8

    def u(self, off, n):
        return int.from_bytes(self.b[off:off + n], "little")

    # symbol table entry: link name offset(8) obj header addr(8) cache type(4) rsvd(4) scratch(16)
    def entry(self, off):
        name_off = self.u(off, 8)
        ohdr = self.u(off + 8, 8)
        cache = self.u(off + 16, 4)
        btree = heap = None
        if cache == 1:
            btree, heap = self.u(off + 24, 8), self.u(off + 32, 8)
        return name_off, ohdr, btree, heap

    def messages(self, ohdr):
        """Yield (type, offset, size) of a version-1 object header."""
        assert self.b[ohdr] == 1
        nmsg = self.u(ohdr + 2, 2)
        size = self.u(ohdr + 8, 4)
        blocks = [(ohdr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            pos, sz = blocks.pop(0)
            end = pos + sz
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize = self.u(pos, 2), self.u(pos + 2, 2)
                body = pos + 8
                if mtype == 0x10:   # continuation
                    blocks.append((self.u(body, 8), self.u(body + 8, 8)))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    def heap_data(self, heap):
        assert self.b[heap:heap + 4] == b"HEAP"
        return self.u(heap + 8 + 2*8, 8)

    def group_entries(self, btree, heap):
        data = self.heap_data(heap)
        out = []

        def walk(node):
            assert self.b[node:node + 4] == b"TREE"
            level, used = self.b[node + 5], self.u(node + 6, 2)
            pos = node + 8 + 2*8
            for i in range(used):
                child = self.u(pos + 8 + i*16, 8)
                if level > 0:
                    walk(child)
                else:
                    assert self.b[child:child + 4] == b"SNOD"
                    nsym = self.u(child + 6, 2)
                    for k in range(nsym):
                        e = child + 8 + k*40
                        name_off, ohdr, bt, hp = self.entry(e)
                        s = data + name_off
                        name = self.b[s:self.b.index(b"\0", s)].decode()
                        out.append((name, ohdr, bt, hp))
        walk(btree)
        return out

    def read_object(self, ohdr, btree=None, heap=None):
        msgs = self.messages(ohdr)
        if btree is None:
            for mtype, body, _ in msgs:
                if mtype == 0x11:   # symbol table message
                    btree, heap = self.u(body, 8), self.u(body + 8, 8)
        if btree is not None:
            return {name: self.read_object(o, bt, hp)
                    for name, o, bt, hp in self.group_entries(btree, heap)}
        shape = dtype = addr = None
        for mtype, body, _ in msgs:
            if mtype == 0x01:      # dataspace v1
                assert self.b[body] == 1
                rank, flags = self.b[body + 1], self.b[body + 2]
                shape = tuple(self.u(body + 8 + 8*i, 8) for i in range(rank))
            elif mtype == 0x03:    # datatype: class 1 (float), 8 bytes, little endian
                cls = self.b[body] & 0x0f
                size = self.u(body + 4, 4)
                assert cls == 1 and size == 8 and (self.b[body + 1] & 1) == 0
                dtype = "<f8"
            elif mtype == 0x08:    # layout
                ver = self.b[body]
                if ver == 3:
                    assert self.b[body + 1] == 1, "contiguous layout only"
                    addr = self.u(body + 2, 8)
                else:
                    assert ver in (1, 2) and self.b[body + 2] == 1
                    addr = self.u(body + 8, 8)
        n = int(np.prod(shape)) if shape else 1
        if n == 0:     # no storage allocated (address undefined)
            return np.zeros(shape)
        return np.frombuffer(self.b, dtype=dtype, count=n, offset=addr).reshape(shape).copy()


def read(path):
    """Return the file content as nested dicts of numpy arrays."""
    with open(path, "rb") as fh:
        f = _File(fh.read())
    _, ohdr, btree, heap = f.entry(f.root_entry)
    return f.read_object(ohdr, btree, heap)


def flat(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(flat(v, prefix + k + "/"))
        else:
            out[prefix + k] = v
    return out
