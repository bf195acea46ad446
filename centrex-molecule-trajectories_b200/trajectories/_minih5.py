"""A small, dependency-free HDF5 reader/writer with an h5py-shaped interface.

The reference persists results through h5py (trajectory_simulator.py:218-254, molecule.py:97-106,169-185,
apertures.py:56-80, utils.py:15-159).  Neither h5py nor libhdf5 exists in the build image, so this module
implements the part of the HDF5 file format (version 1.8 "File Format Specification", the structures every
libhdf5 release reads) that the result layout needs:

  writer   superblock version 0; old-style groups (symbol-table message, version-1 B-tree of group nodes,
           symbol-table nodes "SNOD", local heap "HEAP"); version-1 object headers; contiguous datasets of
           little-endian integers / IEEE floats / fixed strings; attributes (message 0x000C version 1) holding
           scalars, arrays, variable-length UTF-8 strings in a global heap collection "GCOL" (what h5py writes
           for a Python str) and the FALSE/TRUE enum h5py uses for bool.
  reader   the same structures as written by libhdf5 itself, plus what old files add: a user block, object
           header continuation blocks, version-1/2 layout messages, compact datasets, big-endian numbers.
           Chunked datasets and new-style (link-message / fractal-heap) groups raise NotImplementedError.

The interface is the subset of h5py the package uses: `File(path, mode)` as a context manager, `create_group`,
`create_dataset`, item access by path, `del`, `keys/items/values`, `in`, `.attrs` (dict-like), `Dataset[()]`.
A file opened for writing is held as a tree in memory (dataset contents are read lazily) and written out as a
whole when it is closed; that is the right trade for this layout, whose files are written once per run.
`trajectories._hdf.h5py()` returns the real h5py when it is importable and this module otherwise, so files
written here are meant to be read by h5py and vice versa.

tests/test_minih5.py pins the reader against a file written by libhdf5 itself (a MATLAB v7.3 file shipped
with scipy's test data) and checks the writer's bytes against the specification structure by structure.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16               # group leaf / internal node K of the superblock (library defaults)
SNOD_SIZE = 8 + 2 * LEAF_K * 40
TREE_SIZE = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
GCOL_MIN = 4096

MSG_NIL, MSG_DATASPACE, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LAYOUT = 0x0, 0x1, 0x3, 0x4, 0x5, 0x8
MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMBOL_TABLE, MSG_MTIME = 0xC, 0x10, 0x11, 0x12
MSG_LINK_INFO, MSG_LINK = 0x2, 0x6


def _pad8(n: int) -> int:
    return (n + 7) & ~7


def _contig(a: np.ndarray) -> np.ndarray:
    """C-contiguous view or copy that keeps 0-d arrays 0-d (np.ascontiguousarray would make them 1-d)."""
    return a if a.flags.c_contiguous else np.array(a, order="C")


class _VlenStr:
    """Marker dtype: variable-length UTF-8 string (one global-heap reference per element)."""


# ---------------------------------------------------------------------------
# in-memory tree
# ---------------------------------------------------------------------------
class _Node:
    def __init__(self):
        self.attrs: Dict[str, object] = {}


class _GroupNode(_Node):
    def __init__(self):
        super().__init__()
        self.children: Dict[str, _Node] = {}


class _DatasetNode(_Node):
    def __init__(self, data=None, lazy=None):
        super().__init__()
        self._data = data
        self._lazy = lazy              # (path, offset, shape, dtype) of a contiguous block in an existing file

    @property
    def shape(self):
        return self._data.shape if self._data is not None else self._lazy[2]

    @property
    def dtype(self):
        dt = np.dtype(self._data.dtype if self._data is not None else self._lazy[3])
        return dt if dt.isnative else dt.newbyteorder("=")          # what load() returns

    def load(self) -> np.ndarray:
        if self._data is None:
            path, offset, shape, dtype = self._lazy
            count = int(np.prod(shape, dtype=np.int64))
            with open(path, "rb") as f:
                f.seek(offset)
                self._data = np.fromfile(f, dtype=dtype, count=count).reshape(shape)
            if not self._data.dtype.isnative:
                self._data = self._data.astype(self._data.dtype.newbyteorder("="))
        return self._data


# ---------------------------------------------------------------------------
# datatype / dataspace encoding
# ---------------------------------------------------------------------------
def _encode_dtype(dt) -> bytes:
    """Datatype message body (version 1) for a numpy dtype, bool, or _VlenStr."""
    if dt is _VlenStr:
        base = struct.pack("<BBBBI", 0x10, 0x00, 0, 0, 1) + struct.pack("<HH", 0, 8)        # unsigned char
        return struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + base
    dt = np.dtype(dt)
    if dt == np.bool_:
        # h5py's bool: enum of a signed 1-byte integer, members FALSE = 0, TRUE = 1
        base = struct.pack("<BBBBI", 0x10, 0x08, 0, 0, 1) + struct.pack("<HH", 0, 8)
        names = b"FALSE\0\0\0" + b"TRUE\0\0\0\0"
        return struct.pack("<BBBBI", 0x18, 2, 0, 0, 1) + base + names + b"\x00\x01"
    if dt.kind in "iu":
        flags = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10, flags, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize == 8:
        return struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind == "f" and dt.itemsize == 4:
        return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, max(dt.itemsize, 1))                  # null-terminated ASCII
    raise TypeError(f"dtype {dt} cannot be stored")


def _encode_dataspace(shape) -> bytes:
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _normalise_value(value):
    """What h5py would store for `attrs[key] = value`: (array-or-list-of-str, dtype marker, shape)."""
    if isinstance(value, (str, np.str_)):
        return [str(value)], _VlenStr, ()
    if isinstance(value, (bytes, np.bytes_)):
        a = np.array(bytes(value), dtype=f"S{max(len(value), 1)}")
        return a, a.dtype, ()
    if isinstance(value, (bool, np.bool_)):
        return np.array(bool(value)), np.dtype(np.bool_), ()
    if isinstance(value, (int, np.integer)) and not isinstance(value, np.generic):
        return np.array(value, dtype=np.int64), np.dtype(np.int64), ()
    if isinstance(value, float):
        return np.array(value, dtype=np.float64), np.dtype(np.float64), ()
    a = np.asarray(value)
    if a.dtype.kind == "U" or (a.dtype.kind == "O" and all(isinstance(v, str) for v in a.reshape(-1))):
        flat = [str(v) for v in a.reshape(-1)]
        return flat, _VlenStr, a.shape
    if a.dtype.kind not in "iufSb":
        raise TypeError(f"Object dtype {a.dtype} has no native HDF5 equivalent")
    if a.dtype.kind in "iuf" and not a.dtype.isnative:
        a = a.astype(a.dtype.newbyteorder("="))
    return _contig(a), a.dtype, a.shape


# ---------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------
class _GlobalHeap:
    """One global heap collection holding every variable-length string of the file."""

    def __init__(self):
        self.objects: List[bytes] = []
        self.address = UNDEF

    def add(self, data: bytes) -> int:
        self.objects.append(data)
        return len(self.objects)               # heap object index (1-based; 0 is the free space)

    def size(self) -> int:
        used = 16 + sum(16 + _pad8(len(o)) for o in self.objects)
        return max(GCOL_MIN, _pad8(used + 16))

    def encode(self) -> bytes:
        total = self.size()
        out = bytearray(b"GCOL" + struct.pack("<B3xQ", 1, total))
        for k, o in enumerate(self.objects, 1):
            out += struct.pack("<HH4xQ", k, 1, len(o)) + o + b"\0" * (_pad8(len(o)) - len(o))
        free = total - len(out)
        out += struct.pack("<HH4xQ", 0, 0, free) + b"\0" * (free - 16)     # object 0: the free space (size counts its header)
        return bytes(out)


class _Writer:
    def __init__(self, root: _GroupNode):
        self.root = root
        self.gheap = _GlobalHeap()
        self.cursor = 96                       # after the superblock
        self.parts: List[Tuple[int, object]] = []

    def alloc(self, size: int) -> int:
        addr = self.cursor
        self.cursor += _pad8(size)
        return addr

    # -- messages ----------------------------------------------------------
    @staticmethod
    def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
        body = body + b"\0" * (_pad8(len(body)) - len(body))
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def _attribute(self, name: str, value) -> bytes:
        data, dt, shape = _normalise_value(value)
        nm = name.encode() + b"\0"
        dtm, dsm = _encode_dtype(dt), _encode_dataspace(shape)
        if dt is _VlenStr:
            raw = b"".join(self._vlen_ref(s) for s in data)
        elif np.dtype(dt) == np.bool_:
            raw = np.asarray(data, dtype=np.int8).tobytes()
        else:
            raw = _contig(np.asarray(data)).tobytes()
        body = struct.pack("<BxHHH", 1, len(nm), len(dtm), len(dsm))
        for piece in (nm, dtm, dsm):
            body += piece + b"\0" * (_pad8(len(piece)) - len(piece))
        return self._message(MSG_ATTRIBUTE, body + raw)

    def _vlen_ref(self, s: str) -> bytes:
        b = s.encode("utf-8")
        idx = self.gheap.add(b)
        return ("vlen", len(b), idx)           # patched once the heap's address is known

    def _header(self, messages: List[bytes]) -> bytes:
        body = b"".join(messages)
        return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body

    # -- layout pass -------------------------------------------------------
    def _attr_messages(self, node: _Node) -> List[object]:
        out = []
        for k, v in node.attrs.items():
            data, dt, shape = _normalise_value(v)
            if dt is _VlenStr:
                out.append(("attr_vlen", k, [(len(s.encode("utf-8")), self.gheap.add(s.encode("utf-8"))) for s in data], shape))
            else:
                out.append(self._attribute(k, v))
        return out

    def _attr_size(self, m) -> int:
        if isinstance(m, bytes):
            return len(m)
        _, name, refs, shape = m
        nm = len(name.encode()) + 1
        return 8 + _pad8(8 + _pad8(nm) + _pad8(len(_encode_dtype(_VlenStr))) + _pad8(len(_encode_dataspace(shape))) + 16 * len(refs))

    def _attr_bytes(self, m) -> bytes:
        if isinstance(m, bytes):
            return m
        _, name, refs, shape = m
        nm = name.encode() + b"\0"
        dtm, dsm = _encode_dtype(_VlenStr), _encode_dataspace(shape)
        body = struct.pack("<BxHHH", 1, len(nm), len(dtm), len(dsm))
        for piece in (nm, dtm, dsm):
            body += piece + b"\0" * (_pad8(len(piece)) - len(piece))
        raw = b"".join(struct.pack("<IQI", ln, self.gheap.address, idx) for ln, idx in refs)
        return self._message(MSG_ATTRIBUTE, body + raw)

    def plan(self, node: _Node) -> dict:
        """Assign addresses depth first; returns the plan record of `node`."""
        attrs = self._attr_messages(node)
        asize = sum(self._attr_size(m) for m in attrs)
        if isinstance(node, _GroupNode):
            names = sorted(node.children, key=lambda s: s.encode())
            rec = {"node": node, "attrs": attrs, "names": names}
            rec["header"] = self.alloc(16 + 24 + asize)
            heap_data = 8 + sum(_pad8(len(n.encode()) + 1) for n in names) + 16      # "" + names + one free block
            rec["heap_data_size"] = heap_data
            rec["heap"] = self.alloc(32 + heap_data)
            n_snod = max(1, -(-len(names) // (2 * LEAF_K))) if names else 0
            rec["snods"] = [self.alloc(SNOD_SIZE) for _ in range(n_snod)]
            # B-tree levels, bottom up
            levels, width = [], n_snod
            while True:
                n_nodes = max(1, -(-width // (2 * INTERNAL_K)))
                levels.append([self.alloc(TREE_SIZE) for _ in range(n_nodes)])
                if n_nodes == 1:
                    break
                width = n_nodes
            rec["levels"] = levels
            rec["children"] = [self.plan(node.children[n]) for n in names]
            return rec
        rec = {"node": node, "attrs": attrs}
        dt = np.dtype(node.dtype)
        msgs = 8 + _pad8(len(_encode_dataspace(node.shape))) + 8 + _pad8(len(_encode_dtype(dt))) + 16 + 8 + 24
        rec["header"] = self.alloc(16 + msgs + asize)
        nbytes = int(np.prod(node.shape, dtype=np.int64)) * dt.itemsize
        rec["nbytes"] = nbytes
        rec["data"] = self.alloc(nbytes) if nbytes else UNDEF
        return rec

    # -- emission pass -----------------------------------------------------
    def emit(self, rec: dict) -> None:
        node = rec["node"]
        attr_msgs = [self._attr_bytes(m) for m in rec["attrs"]]
        if isinstance(node, _GroupNode):
            names, kids = rec["names"], rec["children"]
            root_tree = rec["levels"][-1][0]
            head = self._header([self._message(MSG_SYMBOL_TABLE, struct.pack("<QQ", root_tree, rec["heap"]), 0)] + attr_msgs)
            self.parts.append((rec["header"], head))
            # local heap: offset 0 holds the empty string every B-tree's first key points at
            data = bytearray(8)
            offs = []
            for n in names:
                offs.append(len(data))
                b = n.encode() + b"\0"
                data += b + b"\0" * (_pad8(len(b)) - len(b))
            free_at = len(data)
            data += struct.pack("<QQ", 1, 16)                                        # the only free block: next = 1 (none), size 16
            assert len(data) == rec["heap_data_size"]
            heap = b"HEAP" + struct.pack("<B3xQQQ", 0, len(data), free_at, rec["heap"] + 32) + bytes(data)
            self.parts.append((rec["heap"], heap))
            # symbol table nodes
            last_name_off = []
            for s, addr in enumerate(rec["snods"]):
                lo, hi = s * 2 * LEAF_K, min(len(names), (s + 1) * 2 * LEAF_K)
                body = bytearray(b"SNOD" + struct.pack("<BxH", 1, hi - lo))
                for k in range(lo, hi):
                    body += struct.pack("<QQII16x", offs[k], kids[k]["header"], 0, 0)
                body += b"\0" * (SNOD_SIZE - len(body))
                self.parts.append((addr, bytes(body)))
                last_name_off.append(offs[hi - 1])
            # B-tree: level 0 points at the symbol table nodes; key[i + 1] = heap offset of the largest name below child i
            child_addrs, child_keys = list(rec["snods"]), last_name_off
            for level, addrs in enumerate(rec["levels"]):
                next_addrs, next_keys = [], []
                for k, addr in enumerate(addrs):
                    lo, hi = k * 2 * INTERNAL_K, min(len(child_addrs), (k + 1) * 2 * INTERNAL_K)
                    left = addrs[k - 1] if k > 0 else UNDEF
                    right = addrs[k + 1] if k + 1 < len(addrs) else UNDEF
                    body = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, level, hi - lo, left, right))
                    first_key = 0 if lo == 0 else child_keys[lo - 1]
                    body += struct.pack("<Q", first_key)
                    for c in range(lo, hi):
                        body += struct.pack("<QQ", child_addrs[c], child_keys[c])
                    body += b"\0" * (TREE_SIZE - len(body))
                    self.parts.append((addr, bytes(body)))
                    next_addrs.append(addr)
                    next_keys.append(child_keys[hi - 1] if hi > lo else 0)
                child_addrs, child_keys = next_addrs, next_keys
            for kid in kids:
                self.emit(kid)
            return
        dt = np.dtype(node.dtype)
        msgs = [
            self._message(MSG_FILL, bytes([1, 2, 2, 1, 0, 0, 0, 0]), 1),
            self._message(MSG_DATATYPE, _encode_dtype(dt), 1),
            self._message(MSG_DATASPACE, _encode_dataspace(node.shape), 0),
            self._message(MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, rec["data"], rec["nbytes"]), 0),
        ] + attr_msgs
        self.parts.append((rec["header"], self._header(msgs)))
        if rec["nbytes"]:
            self.parts.append((rec["data"], node))

    def write(self, path: str) -> None:
        plan = self.plan(self.root)
        self.gheap.address = self.alloc(self.gheap.size()) if self.gheap.objects else UNDEF
        self.emit(plan)
        if self.gheap.objects:
            self.parts.append((self.gheap.address, self.gheap.encode()))
        eof = self.cursor
        root = plan
        sb = (SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
              + struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
              + struct.pack("<QQII", 0, root["header"], 1, 0) + struct.pack("<QQ", root["levels"][-1][0], root["heap"]))
        assert len(sb) == 96
        tmp = f"{path}.tmp{os.getpid()}"
        sources: Dict[str, object] = {}
        with open(tmp, "wb") as f:
            f.write(sb)
            pos = 96
            for addr, part in sorted(self.parts, key=lambda p: p[0]):
                if addr > pos:
                    f.write(b"\0" * (addr - pos))
                    pos = addr
                assert addr == pos, (addr, pos)
                if isinstance(part, _DatasetNode):
                    if part._data is None and np.dtype(part._lazy[3]).isnative:
                        # untouched dataset of the file being rewritten: copy its bytes, one shared handle per source
                        src_path, offset, shape, dtype = part._lazy
                        src = sources.get(src_path)
                        if src is None:
                            src = sources[src_path] = open(src_path, "rb")
                        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
                        src.seek(offset)
                        left = nbytes
                        while left:
                            chunk = src.read(min(left, 1 << 24))
                            if not chunk:
                                raise OSError("source file truncated while copying a dataset")
                            f.write(chunk)
                            left -= len(chunk)
                        pos += nbytes
                        continue
                    arr = _contig(part.load())
                    if arr.dtype == np.bool_:
                        arr = arr.astype(np.int8)
                    arr.tofile(f)
                    pos += arr.nbytes
                else:
                    f.write(part)
                    pos += len(part)
            if pos < eof:
                f.write(b"\0" * (eof - pos))
        for src in sources.values():
            src.close()
        os.replace(tmp, path)


# ---------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------
class _Reader:
    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as f:
            self.buf = f.read()
        self.base = self._find_superblock()
        self._heaps: Dict[int, bytes] = {}

    def _find_superblock(self) -> int:
        off = 0
        while off + 8 <= len(self.buf):
            if self.buf[off:off + 8] == SIGNATURE:
                return off
            off = 512 if off == 0 else off * 2
        raise OSError("Unable to open file (file signature not found)")

    def u(self, fmt: str, addr: int):
        return struct.unpack_from("<" + fmt, self.buf, self.base + addr)

    def root(self) -> _GroupNode:
        version = self.buf[self.base + 8]
        if version not in (0, 1):
            raise NotImplementedError(f"superblock version {version} (written with libver='latest') is not supported")
        so, sl = self.buf[self.base + 13], self.buf[self.base + 14]
        if (so, sl) != (8, 8):
            raise NotImplementedError("only 8-byte offsets and lengths are supported")
        self.leaf_k, self.internal_k = self.u("HH", 16)
        pos = 24 + (4 if version == 1 else 0)
        base_addr, _free, self.eof, _drv = self.u("QQQQ", pos)
        # addresses in the file are relative to the base address, which for a user block equals its size
        _name, header, _cache = self.u("QQI", pos + 32)
        return self.read_object(header)

    # -- object headers ------------------------------------------------------
    def messages(self, addr: int) -> Iterator[Tuple[int, int, bytes]]:
        version = self.buf[self.base + addr]
        if version != 1:
            raise NotImplementedError("version-2 object headers (libver='latest') are not supported")
        n_msgs, _refs, size = self.u("xxHII", addr)
        blocks = [(addr + 16, size)]
        seen = 0
        while blocks and seen < n_msgs:
            pos, left = blocks.pop(0)
            while left >= 8 and seen < n_msgs:
                mtype, msize, flags = self.u("HHB", pos)
                body = self.buf[self.base + pos + 8: self.base + pos + 8 + msize]
                seen += 1
                if mtype == MSG_CONTINUATION:
                    blocks.append(struct.unpack("<QQ", body[:16]))
                else:
                    yield mtype, flags, body
                pos += 8 + msize
                left -= 8 + msize

    def read_object(self, addr: int) -> _Node:
        msgs = list(self.messages(addr))
        kinds = {m[0] for m in msgs}
        if MSG_SYMBOL_TABLE in kinds:
            node: _Node = _GroupNode()
            btree, heap = struct.unpack("<QQ", next(b for t, _, b in msgs if t == MSG_SYMBOL_TABLE)[:16])
            for name, child in self.group_entries(btree, heap):
                node.children[name] = _LazyChild(self, child)
        elif MSG_LINK_INFO in kinds or MSG_LINK in kinds:
            raise NotImplementedError("new-style groups (link messages) are not supported; write with libver='earliest'")
        elif MSG_LAYOUT in kinds:
            node = self.read_dataset(msgs)
        else:
            node = _GroupNode()
        for t, _, body in msgs:
            if t == MSG_ATTRIBUTE:
                name, value = self.read_attribute(body)
                node.attrs[name] = value
        return node

    # -- groups ---------------------------------------------------------------
    def heap_data(self, addr: int) -> Tuple[int, int]:
        assert self.buf[self.base + addr: self.base + addr + 4] == b"HEAP", "bad local heap signature"
        size, _free, data = self.u("QQQ", addr + 8)
        return data, size

    def group_entries(self, btree: int, heap: int) -> List[Tuple[str, int]]:
        data, _ = self.heap_data(heap)
        out: List[Tuple[str, int]] = []

        def name_at(off: int) -> str:
            start = self.base + data + off
            return self.buf[start: self.buf.index(b"\0", start)].decode("utf-8")

        def walk(addr: int):
            sig = self.buf[self.base + addr: self.base + addr + 4]
            if sig == b"TREE":
                ntype, level, used = self.u("BBH", addr + 4)
                assert ntype == 0, "not a group B-tree"
                for k in range(used):
                    (child,) = self.u("Q", addr + 24 + 8 + 16 * k)
                    walk(child)
            elif sig == b"SNOD":
                (count,) = self.u("H", addr + 6)
                for k in range(count):
                    off, header = self.u("QQ", addr + 8 + 40 * k)
                    out.append((name_at(off), header))
            else:
                raise OSError(f"bad group node signature {sig!r} at {addr:#x}")

        walk(btree)
        return out

    # -- datatypes -------------------------------------------------------------
    def decode_dtype(self, b: bytes):
        """-> (numpy dtype | _VlenStr | ('enum_bool', base dtype), bytes consumed)"""
        cls, ver = b[0] & 0x0F, b[0] >> 4
        bits = b[1] | (b[2] << 8) | (b[3] << 16)
        (size,) = struct.unpack_from("<I", b, 4)
        order = ">" if bits & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits & 0x08 else 'u'}{size}"), 12
        if cls == 1:
            return np.dtype(f"{order}f{size}"), 20
        if cls == 3:
            return np.dtype(f"S{size}"), 8
        if cls == 9:
            if (bits & 0x0F) != 1:
                raise NotImplementedError("variable-length sequences are not supported")
            _, used = self.decode_dtype(b[8:])
            return _VlenStr, 8 + used
        if cls == 8:
            n_members = bits & 0xFFFF
            base, used = self.decode_dtype(b[8:])
            pos, names = 8 + used, []
            for _ in range(n_members):
                end = b.index(b"\0", pos)
                names.append(b[pos:end].decode())
                pos = end + 1 if ver >= 3 else pos + _pad8(end + 1 - pos)
            values = np.frombuffer(b, dtype=base, count=n_members, offset=pos)
            if names == ["FALSE", "TRUE"] and values.tolist() == [0, 1]:
                return ("enum_bool", base), pos + n_members * base.itemsize
            return base, pos + n_members * base.itemsize
        raise NotImplementedError(f"datatype class {cls} is not supported")

    @staticmethod
    def decode_dataspace(b: bytes) -> Tuple[int, ...]:
        version, rank, flags = b[0], b[1], b[2]
        if version == 1:
            return tuple(struct.unpack_from(f"<{rank}Q", b, 8)) if rank else ()
        if version == 2:
            if b[3] == 2:          # null dataspace
                return (0,)
            return tuple(struct.unpack_from(f"<{rank}Q", b, 4)) if rank else ()
        raise NotImplementedError(f"dataspace version {version}")

    def global_heap_object(self, addr: int, index: int) -> bytes:
        assert self.buf[self.base + addr: self.base + addr + 4] == b"GCOL", "bad global heap signature"
        (total,) = self.u("Q", addr + 8)
        pos, end = addr + 16, addr + total
        while pos + 16 <= end:
            idx, _refs, size = self.u("HH4xQ", pos)
            if idx == 0:
                break
            if idx == index:
                return self.buf[self.base + pos + 16: self.base + pos + 16 + size]
            pos += 16 + _pad8(size)
        raise OSError(f"global heap object {index} not found")

    def decode_values(self, dt, shape, raw: bytes):
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if dt is _VlenStr:
            vals = []
            for k in range(count):
                ln, addr, idx = struct.unpack_from("<IQI", raw, 16 * k)
                vals.append(self.global_heap_object(addr, idx)[:ln].decode("utf-8") if ln else "")
            return vals[0] if shape == () else np.array(vals, dtype=object).reshape(shape)
        if isinstance(dt, tuple):
            arr = np.frombuffer(raw, dtype=dt[1], count=count).astype(np.bool_)
        else:
            arr = np.frombuffer(raw, dtype=dt, count=count)
            if not arr.dtype.isnative:
                arr = arr.astype(arr.dtype.newbyteorder("="))
        if shape == ():
            v = arr[0]
            return bytes(v) if arr.dtype.kind == "S" else v
        return arr.reshape(shape).copy()

    def read_attribute(self, b: bytes):
        version = b[0]
        if version == 1:
            nsz, dsz, ssz = struct.unpack_from("<HHH", b, 2)
            pos = 8
            name = b[pos: pos + nsz].split(b"\0")[0].decode("utf-8")
            pos += _pad8(nsz)
            dt, _ = self.decode_dtype(b[pos: pos + dsz])
            pos += _pad8(dsz)
            shape = self.decode_dataspace(b[pos: pos + ssz])
            pos += _pad8(ssz)
        elif version in (2, 3):
            nsz, dsz, ssz = struct.unpack_from("<HHH", b, 2)
            pos = 8 + (1 if version == 3 else 0)
            name = b[pos: pos + nsz].split(b"\0")[0].decode("utf-8")
            pos += nsz
            dt, _ = self.decode_dtype(b[pos: pos + dsz])
            pos += dsz
            shape = self.decode_dataspace(b[pos: pos + ssz])
            pos += ssz
        else:
            raise NotImplementedError(f"attribute message version {version}")
        return name, self.decode_values(dt, shape, b[pos:])

    # -- datasets ---------------------------------------------------------------
    def read_dataset(self, msgs) -> _DatasetNode:
        dt = shape = layout = None
        for t, _, body in msgs:
            if t == MSG_DATATYPE:
                dt, _ = self.decode_dtype(body)
            elif t == MSG_DATASPACE:
                shape = self.decode_dataspace(body)
            elif t == MSG_LAYOUT:
                layout = body
        if dt is _VlenStr:
            raise NotImplementedError("variable-length string datasets are not supported")
        as_bool = isinstance(dt, tuple)
        if as_bool:
            dt = dt[1]
        version = layout[0]
        if version == 3:
            cls = layout[1]
            if cls == 1:
                addr, _size = struct.unpack_from("<QQ", layout, 2)
            elif cls == 0:
                (size,) = struct.unpack_from("<H", layout, 2)
                data = np.frombuffer(layout[4: 4 + size], dtype=dt).reshape(shape).copy()
                return _DatasetNode(data=data.astype(np.bool_) if as_bool else data)
            else:
                raise NotImplementedError("chunked datasets are not supported")
        elif version in (1, 2):
            rank, cls = layout[1], layout[2]
            if cls != 1:
                raise NotImplementedError("only contiguous datasets are supported for layout versions 1 and 2")
            (addr,) = struct.unpack_from("<Q", layout, 8)
        else:
            raise NotImplementedError(f"layout message version {version}")
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if addr == UNDEF or count == 0:
            data = np.zeros(shape, dtype=dt)
            return _DatasetNode(data=data.astype(np.bool_) if as_bool else data)
        if as_bool:
            data = np.frombuffer(self.buf, dtype=dt, count=count, offset=self.base + addr).reshape(shape).astype(np.bool_)
            return _DatasetNode(data=data)
        return _DatasetNode(lazy=(self.path, self.base + addr, tuple(shape), np.dtype(dt)))


class _LazyChild(_Node):
    """A child that has not been parsed yet (its object header address is known)."""

    def __init__(self, reader: _Reader, addr: int):
        self.reader, self.addr = reader, addr

    def resolve(self) -> _Node:
        return self.reader.read_object(self.addr)


# ---------------------------------------------------------------------------
# h5py-shaped front end
# ---------------------------------------------------------------------------
class AttributeManager:
    def __init__(self, node: _Node, file: "File"):
        self._node, self._file = node, file

    def __getitem__(self, key):
        v = self._node.attrs[key]
        return v

    def __setitem__(self, key, value):
        self._file._require_writable()
        data, dt, shape = _normalise_value(value)          # validates the type like h5py (TypeError)
        if dt is _VlenStr:
            stored = data[0] if shape == () else np.array(data, dtype=object).reshape(shape)
        elif shape == ():
            stored = bytes(data) if np.dtype(dt).kind == "S" else np.asarray(data).reshape(())[()]
        else:
            stored = np.array(data)
        self._node.attrs[key] = stored
        self._file._dirty = True

    def __delitem__(self, key):
        self._file._require_writable()
        del self._node.attrs[key]
        self._file._dirty = True

    def __contains__(self, key):
        return key in self._node.attrs

    def __iter__(self):
        return iter(self._node.attrs)

    def __len__(self):
        return len(self._node.attrs)

    def keys(self):
        return self._node.attrs.keys()

    def values(self):
        return self._node.attrs.values()

    def items(self):
        return self._node.attrs.items()

    def get(self, key, default=None):
        return self._node.attrs.get(key, default)


class Dataset:
    def __init__(self, node: _DatasetNode, file: "File", name: str):
        self._node, self._file, self.name = node, file, name

    @property
    def attrs(self) -> AttributeManager:
        return AttributeManager(self._node, self._file)

    @property
    def shape(self):
        return tuple(self._node.shape)

    @property
    def dtype(self):
        return np.dtype(self._node.dtype)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key):
        data = self._node.load()
        out = data[key]
        return out.copy() if isinstance(out, np.ndarray) else out

    def __array__(self, dtype=None, copy=None):
        a = self._node.load()
        return a.astype(dtype) if dtype is not None else a


def _split(path: str) -> List[str]:
    return [p for p in str(path).split("/") if p]


class Group:
    def __init__(self, node: _GroupNode, file: "File", name: str):
        self._node, self._file, self.name = node, file, name

    @property
    def attrs(self) -> AttributeManager:
        return AttributeManager(self._node, self._file)

    # -- navigation -----------------------------------------------------------
    def _child(self, node: _GroupNode, part: str) -> _Node:
        child = node.children[part]
        if isinstance(child, _LazyChild):
            child = node.children[part] = child.resolve()
        return child

    def _walk(self, path: str) -> _Node:
        node: _Node = self._file._root if str(path).startswith("/") else self._node
        for part in _split(path):
            if not isinstance(node, _GroupNode) or part not in node.children:
                raise KeyError(f"Unable to open object (object '{part}' doesn't exist)")
            node = self._child(node, part)
        return node

    def _wrap(self, node: _Node, path: str):
        full = "/" + "/".join(_split(self.name) + _split(path)) if not str(path).startswith("/") else "/" + "/".join(_split(path))
        return Group(node, self._file, full) if isinstance(node, _GroupNode) else Dataset(node, self._file, full)

    def __getitem__(self, path):
        return self._wrap(self._walk(path), path)

    def __contains__(self, path) -> bool:
        try:
            self._walk(path)
            return True
        except KeyError:
            return False

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._node.children)

    def keys(self):
        return sorted(self._node.children, key=lambda s: s.encode())     # the order of a group B-tree (and of h5py)

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def get(self, path, default=None):
        return self[path] if path in self else default

    # -- modification -----------------------------------------------------------
    def _parent_of(self, path: str, create: bool) -> Tuple[_GroupNode, str]:
        parts = _split(path)
        if not parts:
            raise ValueError("Unable to create group (name already exists)")
        node: _Node = self._file._root if str(path).startswith("/") else self._node
        for part in parts[:-1]:
            if part in node.children:
                node = self._child(node, part)
                if not isinstance(node, _GroupNode):
                    raise ValueError(f"Unable to create group ('{part}' is a dataset)")
            elif create:
                new = _GroupNode()
                node.children[part] = new
                node = new
            else:
                raise KeyError(f"Unable to open object (object '{part}' doesn't exist)")
        return node, parts[-1]

    def create_group(self, path) -> "Group":
        self._file._require_writable()
        parent, leaf = self._parent_of(path, create=True)
        if leaf in parent.children:
            raise ValueError("Unable to create group (name already exists)")
        node = _GroupNode()
        parent.children[leaf] = node
        self._file._dirty = True
        return self._wrap(node, path)

    def require_group(self, path) -> "Group":
        return self[path] if path in self else self.create_group(path)

    def create_dataset(self, path, shape=None, dtype=None, data=None) -> Dataset:
        self._file._require_writable()
        parent, leaf = self._parent_of(path, create=True)
        if leaf in parent.children:
            raise ValueError("Unable to create dataset (name already exists)")
        if data is None:
            data = np.zeros(shape if shape is not None else (), dtype=dtype or np.float32)
        arr = np.asarray(data, dtype=dtype)
        if arr.dtype.kind in "UO":
            raise TypeError("string datasets are not supported by this writer (use bytes)")
        if arr.dtype.kind in "iuf" and not arr.dtype.isnative:
            arr = arr.astype(arr.dtype.newbyteorder("="))
        _encode_dtype(arr.dtype)                                          # validates
        if shape is not None and tuple(np.atleast_1d(shape)) != arr.shape:
            arr = arr.reshape(shape)
        node = _DatasetNode(data=_contig(arr))
        parent.children[leaf] = node
        self._file._dirty = True
        return self._wrap(node, path)

    def __setitem__(self, path, value):
        self.create_dataset(path, data=value)

    def __delitem__(self, path):
        self._file._require_writable()
        parent, leaf = self._parent_of(path, create=False)
        if leaf not in parent.children:
            raise KeyError(f"Couldn't delete link (name '{leaf}' doesn't exist)")
        del parent.children[leaf]
        self._file._dirty = True


class File(Group):
    """`File(path, mode)` with h5py's modes: r, r+, a (default here as in the reference's calls), w, w- / x."""

    def __init__(self, path, mode: str = "r"):
        self.filename = os.fspath(path)
        self.mode = mode
        exists = os.path.exists(self.filename)
        if mode in ("r", "r+") and not exists:
            raise FileNotFoundError(f"Unable to open file (unable to open file: name = '{self.filename}')")
        if mode in ("w-", "x") and exists:
            raise FileExistsError(f"Unable to create file (file exists): '{self.filename}'")
        if mode not in ("r", "r+", "a", "w", "w-", "x"):
            raise ValueError(f"Invalid mode {mode!r}")
        self._writable = mode != "r"
        self._open = True
        if exists and mode in ("r", "r+", "a"):
            self._root = _Reader(self.filename).root()
            self._dirty = False
        else:
            self._root = _GroupNode()
            self._dirty = True
        super().__init__(self._root, self, "/")

    def _require_writable(self):
        if not self._open:
            raise ValueError("Invalid file (the file is closed)")
        if not self._writable:
            raise ValueError("Unable to modify file (no write intent on file)")

    def _resolve_all(self, node: _GroupNode):
        for k, child in list(node.children.items()):
            if isinstance(child, _LazyChild):
                child = node.children[k] = child.resolve()
            if isinstance(child, _GroupNode):
                self._resolve_all(child)

    def flush(self):
        if self._open and self._writable and self._dirty:
            self._resolve_all(self._root)
            _Writer(self._root).write(self.filename)
            self._dirty = False

    def close(self):
        if self._open:
            self.flush()
            self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
