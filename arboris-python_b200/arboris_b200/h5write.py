"""Minimal HDF5 writer for the trajectory files of ``arboris_b200.observers``.

The reference's ``Hdf5Logger`` (observers.py:133-289) needs h5py, which is not
available here; its files only use the oldest on-disk structures (superblock
version 0, version-1 object headers, symbol-table groups -- B-tree v1 + local
heap + symbol-table nodes -- and contiguous little-endian float64 datasets),
which is what this module emits from a nested ``dict`` of numpy arrays.  Format
reference: "HDF5 File Format Specification Version 1.1", sections III.A
(B-trees), III.B (symbol-table nodes), III.D (local heaps), IV.A.1 (object
header), IV.A.2.b/d/i/r (dataspace, datatype, layout, symbol-table messages).
"""
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K = 64          # a symbol-table node holds up to 2*_LEAF_K entries
_INTERNAL_K = 16      # a B-tree node points to up to 2*_INTERNAL_K symbol-table nodes


def _pad8(b):
    return b + b"\0"*((-len(b)) % 8)


class _Writer(object):
    def __init__(self):
        self.buf = bytearray(96)        # superblock, filled in at the end

    def alloc(self, data):
        """Append ``data`` 8-byte aligned, return its address."""
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    # ---- object headers ---------------------------------------------------------------
    @staticmethod
    def _message(mtype, body):
        body = _pad8(body)
        return struct.pack("<HHB3x", mtype, len(body), 0) + body

    def _object_header(self, messages):
        data = b"".join(messages)
        head = struct.pack("<BxHII4x", 1, len(messages), 1, len(data))
        return self.alloc(head + data)

    def dataset(self, arr):
        arr = np.ascontiguousarray(arr, dtype="<f8")
        addr = self.alloc(arr.tobytes()) if arr.size else _UNDEF
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        # IEEE 754 little-endian binary64: class 1 (floating point), version 1
        dtype = struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + \
            struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        # fill value (version 2): allocation time late, write time if set, undefined value
        fill = struct.pack("<BBBB", 2, 2, 2, 0)
        layout = struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)
        return self._object_header([self._message(0x0001, space), self._message(0x0003, dtype),
                                    self._message(0x0005, fill), self._message(0x0008, layout)])

    # ---- groups -----------------------------------------------------------------------
    def group(self, tree):
        """Write ``tree`` (dict name -> dict | array); returns (object header, btree, heap)."""
        entries = []
        for name in sorted(tree, key=lambda s: s.encode()):
            v = tree[name]
            if isinstance(v, dict):
                entries.append((name, ) + self.group(v))
            else:
                entries.append((name, self.dataset(v), None, None))
        if len(entries) > 4*_LEAF_K*_INTERNAL_K:
            raise ValueError("too many entries in one group for a single-level B-tree")
        # local heap: "" at offset 0, then the names, each padded to 8 bytes
        heap_data = bytearray(8)
        offs = []
        for name, _, _, _ in entries:
            offs.append(len(heap_data))
            heap_data += _pad8(name.encode() + b"\0")
        free = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)         # one free block of 16 bytes at the end
        data_addr = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free, data_addr))
        # symbol-table nodes (sorted by name), each allocated at full capacity
        cap = 2*_LEAF_K
        nodes, keys = [], [0]
        for i in range(0, max(len(entries), 1), cap):
            chunk = entries[i:i + cap]
            body = b"SNOD" + struct.pack("<BxH", 1, len(chunk))
            for k, (name, ohdr, bt, hp) in enumerate(chunk):
                if bt is None:
                    body += struct.pack("<QQII16x", offs[i + k], ohdr, 0, 0)
                else:
                    body += struct.pack("<QQIIQQ", offs[i + k], ohdr, 1, 0, bt, hp)
            body += b"\0"*(8 + cap*40 - len(body))
            nodes.append(self.alloc(body))
            keys.append(offs[i + len(chunk) - 1] if chunk else 0)
        tree_node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(nodes), _UNDEF, _UNDEF)
        for i, n in enumerate(nodes):
            tree_node += struct.pack("<QQ", keys[i], n)
        tree_node += struct.pack("<Q", keys[len(nodes)])
        tree_node += b"\0"*(24 + (2*_INTERNAL_K)*16 + 8 - len(tree_node))
        btree = self.alloc(tree_node)
        ohdr = self._object_header([self._message(0x0011, struct.pack("<QQ", btree, heap))])
        return ohdr, btree, heap

    def finish(self, root):
        ohdr, btree, heap = root
        while len(self.buf) % 8:
            self.buf.append(0)
        sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, len(self.buf), _UNDEF)
        sb += struct.pack("<QQIIQQ", 0, ohdr, 1, 0, btree, heap)
        assert len(sb) == 96
        self.buf[0:96] = sb
        return bytes(self.buf)


def dumps(tree):
    """Serialise a nested dict of float arrays as an HDF5 file image."""
    w = _Writer()
    return w.finish(w.group(tree))


def write(path, tree):
    with open(path, "wb") as fh:
        fh.write(dumps(tree))
