"""trajectories._minih5: the dependency-free HDF5 writer/reader behind SimulationResult.save_to_hdf
(reference trajectory_simulator.py:218-254, utils.py:15-159) when h5py is not installed.

Pinned three ways: the reader against a file written by libhdf5 itself (a MATLAB v7.3 file in scipy's test
data: user block, superblock 0, old-style group, version-1 object header, layout version 2), the writer's bytes
against the HDF5 file-format specification structure by structure (with an independent parser written here),
and writer -> reader round trips over every type the result layout stores."""
import struct
from pathlib import Path

import numpy as np
import pytest

from trajectories import _minih5 as h5

UNDEF = 0xFFFFFFFFFFFFFFFF


def scipy_sample():
    import scipy.io

    p = Path(scipy.io.__file__).parent / "matlab" / "tests" / "data" / "testhdf5_7.4_GLNX86.mat"
    if not p.exists():
        pytest.skip("scipy's HDF5 test file is not installed")
    return p


def test_reads_file_written_by_libhdf5():
    with h5.File(scipy_sample(), "r") as f:
        assert f.keys() == ["testdouble"]
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.float64
        np.testing.assert_array_equal(d[()].ravel(), np.linspace(0, 2 * np.pi, 9))
        assert dict(d.attrs.items()) == {"MATLAB_class": b"double"}
        assert d[3, 0] == np.linspace(0, 2 * np.pi, 9)[3]


# ---------------------------------------------------------------------------
# an independent walk over the bytes, following the specification
# ---------------------------------------------------------------------------
class Spec:
    def __init__(self, path):
        self.b = Path(path).read_bytes()

    def u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self.b, off)

    def superblock(self):
        b = self.b
        assert b[:8] == b"\x89HDF\r\n\x1a\n"
        ver_sb, ver_fs, ver_root, _r0, ver_shm, size_off, size_len, _r1 = b[8:16]
        assert (ver_sb, ver_fs, ver_root, ver_shm) == (0, 0, 0, 0) and (size_off, size_len) == (8, 8)
        leaf_k, int_k, flags = self.u("HHI", 16)
        assert (leaf_k, int_k, flags) == (4, 16, 0)
        base, free, eof, driver = self.u("QQQQ", 24)
        assert base == 0 and free == UNDEF and driver == UNDEF and eof == len(b)
        name_off, header, cache, _res, btree, heap = self.u("QQIIQQ", 56)
        assert name_off == 0 and cache == 1
        return header, btree, heap

    def messages(self, addr):
        version, _res, n_msgs, refs, size = self.u("BBHII", addr)
        assert version == 1 and refs == 1 and addr % 8 == 0
        pos, end, out = addr + 16, addr + 16 + size, []
        while pos < end:
            mtype, msize, flags = self.u("HHB", pos)
            assert msize % 8 == 0, "message bodies are padded to 8 bytes in version-1 headers"
            out.append((mtype, flags, self.b[pos + 8: pos + 8 + msize]))
            pos += 8 + msize
        assert pos == end and len(out) == n_msgs
        return out

    def heap(self, addr):
        assert self.b[addr: addr + 4] == b"HEAP" and self.b[addr + 4] == 0
        size, free, data = self.u("QQQ", addr + 8)
        assert data == addr + 32 and size % 8 == 0
        assert self.b[data: data + 8] == bytes(8)                  # offset 0: the empty name
        # the free list: a chain of (next, size) blocks, terminated by next == 1
        while free != 1:
            assert free % 8 == 0 and free + 16 <= size
            nxt, fsize = self.u("QQ", data + free)
            assert fsize >= 16 and free + fsize <= size
            free = nxt
        return data

    def name(self, heap_data, off):
        end = self.b.index(b"\0", heap_data + off)
        return self.b[heap_data + off: end].decode()

    def tree(self, addr, heap_data, lo_name=""):
        """Names below the node at `addr`, checking every B-tree invariant on the way."""
        assert self.b[addr: addr + 4] == b"TREE"
        ntype, level, used, left, right = self.u("BBHQQ", addr + 4)
        assert ntype == 0 and 1 <= used <= 32 or (used == 0 and level == 0)
        keys = [self.u("Q", addr + 24 + 16 * k)[0] for k in range(used + 1)]
        kids = [self.u("Q", addr + 32 + 16 * k)[0] for k in range(used)]
        names = []
        for k, child in enumerate(kids):
            lo, hi = self.name(heap_data, keys[k]), self.name(heap_data, keys[k + 1])
            if level > 0:
                sub = self.tree(child, heap_data)
            else:
                assert self.b[child: child + 4] == b"SNOD" and self.b[child + 4] == 1
                (count,) = self.u("H", child + 6)
                assert 1 <= count <= 8
                sub = []
                for e in range(count):
                    noff, header, cache, _r = self.u("QQII", child + 8 + 40 * e)
                    sub.append((self.name(heap_data, noff), header))
            # a child holds the names in (key[k], key[k+1]]; the right key IS its largest name
            assert all(lo.encode() < n.encode() <= hi.encode() for n, _ in sub)
            assert sub[-1][0] == hi
            names += sub
        assert [n for n, _ in names] == sorted((n for n, _ in names), key=str.encode)
        return names


def test_writer_structures_follow_the_specification(tmp_path):
    path = tmp_path / "spec.h5"
    x = np.arange(18.0).reshape(9, 2)
    with h5.File(path, "w") as f:
        g = f.create_group("run/trajectories/molecule_0")
        g.create_dataset("x", data=x)
        g.attrs["aperture_hit"] = "Detected"
        g.attrs["alive"] = True
        f["run"].attrs["n"] = 7
        f["run"].attrs["z0"] = 0.25
    s = Spec(path)
    root_header, root_btree, root_heap = s.superblock()
    msgs = s.messages(root_header)
    assert msgs[0][0] == 0x11 and struct.unpack("<QQ", msgs[0][2]) == (root_btree, root_heap)
    heap_data = s.heap(root_heap)
    (name, run_header), = s.tree(root_btree, heap_data)
    assert name == "run"
    run_msgs = s.messages(run_header)
    assert [m[0] for m in run_msgs] == [0x11, 0x0C, 0x0C]
    # attribute message, version 1: int64 scalar `n` = 7
    body = run_msgs[1][2]
    ver, _r, nsz, dsz, ssz = struct.unpack_from("<BBHHH", body, 0)
    assert (ver, nsz, dsz, ssz) == (1, 2, 12, 8) and body[8:10] == b"n\0"
    assert body[16:28] == bytes([0x10, 0x08, 0, 0, 8, 0, 0, 0, 0, 0, 64, 0])      # fixed point, signed, LE, 8 bytes, 64 bits
    assert body[32:40] == bytes([1, 0, 0, 0, 0, 0, 0, 0])                          # scalar dataspace, version 1
    assert struct.unpack_from("<q", body, 40) == (7,)
    # float64 scalar `z0`: the datatype bytes are the ones libhdf5 itself wrote into scipy's sample file
    body = run_msgs[2][2]
    want_f64 = bytes.fromhex("11203f0008000000000040003 40b0034ff030000".replace(" ", ""))
    assert body[16:36] == want_f64 and struct.unpack_from("<d", body, 48) == (0.25,)
    sample = scipy_sample().read_bytes()
    assert want_f64 in sample
    # down to the dataset
    btree, heap = struct.unpack("<QQ", run_msgs[0][2])
    (_, traj_header), = s.tree(btree, s.heap(heap))
    btree, heap = struct.unpack("<QQ", s.messages(traj_header)[0][2])
    (_, mol_header), = s.tree(btree, s.heap(heap))
    mol_msgs = s.messages(mol_header)
    assert [m[0] for m in mol_msgs] == [0x11, 0x0C, 0x0C]
    # a Python str is a variable-length UTF-8 string: a (length, collection address, index) reference into a global heap
    body = mol_msgs[1][2]
    assert body[8:21] == b"aperture_hit\0"
    assert body[24:32] == bytes([0x19, 0x01, 0x01, 0, 16, 0, 0, 0])
    ln, gcol, idx = struct.unpack_from("<IQI", body, 24 + 24 + 8)
    assert s.b[gcol: gcol + 4] == b"GCOL" and s.b[gcol + 4] == 1 and ln == 8
    (gsize,) = s.u("Q", gcol + 8)
    assert gsize >= 4096 and gsize % 8 == 0 and gcol + gsize <= len(s.b)
    pos = gcol + 16
    found = None
    while pos < gcol + gsize:
        oidx, refs, osize = s.u("HH4xQ", pos)
        if oidx == 0:
            assert pos + osize == gcol + gsize                   # the free space runs to the end of the collection
            break
        if oidx == idx:
            found = s.b[pos + 16: pos + 16 + osize]
        pos += 16 + (osize + 7) // 8 * 8
    assert found == b"Detected"
    # bool: the FALSE/TRUE enum over a signed byte that h5py writes
    body = mol_msgs[2][2]
    assert body[8:14] == b"alive\0" and body[16] == 0x18 and body[17] == 2
    assert b"FALSE\0\0\0TRUE\0\0\0\0\x00\x01" in body
    # dataset: fill value, datatype, dataspace, contiguous layout version 3
    btree, heap = struct.unpack("<QQ", mol_msgs[0][2])
    (name, x_header), = s.tree(btree, s.heap(heap))
    assert name == "x"
    x_msgs = s.messages(x_header)
    assert [m[0] for m in x_msgs] == [0x05, 0x03, 0x01, 0x08]
    assert x_msgs[0][2] == bytes([1, 2, 2, 1, 0, 0, 0, 0]) and x_msgs[0][2] in sample      # as libhdf5 writes it
    assert x_msgs[1][2][:20] == want_f64
    assert x_msgs[2][2] == struct.pack("<BBBB4xQQ", 1, 2, 0, 0, 9, 2)
    ver, cls, addr, size = struct.unpack_from("<BBQQ", x_msgs[3][2], 0)
    assert (ver, cls, size) == (3, 1, x.nbytes) and addr % 8 == 0
    np.testing.assert_array_equal(np.frombuffer(s.b, dtype="<f8", count=18, offset=addr).reshape(9, 2), x)


def test_many_children_make_a_multi_level_tree(tmp_path):
    path = tmp_path / "many.h5"
    n = 3000
    with h5.File(path, "w") as f:
        g = f.create_group("trajectories")
        for i in range(n):
            g.create_dataset(f"molecule_{i}", data=np.array([float(i)]))
    s = Spec(path)
    _, btree, heap = s.superblock()
    (_, header), = s.tree(btree, s.heap(heap))
    btree, heap = struct.unpack("<QQ", s.messages(header)[0][2])
    assert s.b[btree + 5] >= 1                                   # 375 symbol-table nodes need an internal level
    names = s.tree(btree, s.heap(heap))
    assert len(names) == n
    with h5.File(path, "r") as f:
        g = f["trajectories"]
        assert len(g) == n and g.keys() == sorted((f"molecule_{i}" for i in range(n)), key=str.encode)
        for i in (0, 7, 8, 9, 255, 256, 257, 2999):
            assert g[f"molecule_{i}"][()] == [float(i)]


def test_round_trip_of_every_stored_type(tmp_path):
    path = tmp_path / "rt.h5"
    values = {
        "class": "RectangularAperture", "name": "DR aperture", "unicode": "größe µ", "empty": "",
        "f": 0.018, "i": 5, "neg": -3, "big": 2 ** 40, "flag": True, "off": np.bool_(False),
        "np_f": np.float64(1.5), "np_i32": np.int32(-9), "np_u8": np.uint8(200), "f32": np.float32(0.5),
        "vec": np.array([1.0, 2.0, 3.0]), "ints": [1, 2, 3], "mat": np.arange(6).reshape(2, 3), "raw": b"abc",
        "names": np.array(["a", "bc"], dtype=object),
    }
    with h5.File(path, "a") as f:
        g = f.create_group("lens run/2022/beamline/")          # '/' nests; trailing and doubled slashes are ignored
        e = f.create_group("lens run/2022/beamline//DR aperture")
        for k, v in values.items():
            e.attrs[k] = v
        with pytest.raises(TypeError):
            e.attrs["bad"] = object()
        with pytest.raises(ValueError):
            f.create_group("lens run/2022/beamline")
        f.create_dataset("d/f64", data=np.linspace(0, 1, 7).reshape(7, 1))
        f.create_dataset("d/i64", data=np.arange(5))
        f.create_dataset("d/u8", data=np.arange(4, dtype=np.uint8))
        f.create_dataset("d/f32", data=np.ones((2, 2), dtype=np.float32))
        f.create_dataset("d/empty", data=np.empty((0, 3)))
        f.create_dataset("d/scalar", data=np.float64(2.5))
        f["d/assigned"] = np.arange(3.0)
        assert g.name == "/lens run/2022/beamline"
    with h5.File(path, "r") as f:
        a = f["lens run/2022/beamline/DR aperture"].attrs
        assert set(a.keys()) == set(values)
        for k in ("class", "name", "unicode", "empty"):
            assert a[k] == values[k] and type(a[k]) is str
        assert a["f"] == 0.018 and isinstance(a["f"], np.float64)
        assert a["i"] == 5 and a["neg"] == -3 and a["big"] == 2 ** 40 and isinstance(a["i"], np.int64)
        assert a["flag"] is np.True_ and a["off"] is np.False_
        assert a["np_i32"] == -9 and a["np_i32"].dtype == np.int32 and a["np_u8"] == 200 and a["f32"].dtype == np.float32
        np.testing.assert_array_equal(a["vec"], [1.0, 2.0, 3.0])
        np.testing.assert_array_equal(a["ints"], [1, 2, 3])
        np.testing.assert_array_equal(a["mat"], np.arange(6).reshape(2, 3))
        assert a["raw"] == b"abc" and a["names"].tolist() == ["a", "bc"]
        assert f["d"].keys() == ["assigned", "empty", "f32", "f64", "i64", "scalar", "u8"]
        np.testing.assert_array_equal(f["d/f64"][()], np.linspace(0, 1, 7).reshape(7, 1))
        np.testing.assert_array_equal(f["d/f64"][2:4, 0], np.linspace(0, 1, 7)[2:4])
        assert f["d/i64"].dtype == np.int64 and f["d/u8"].dtype == np.uint8 and f["d/f32"].dtype == np.float32
        assert f["d/empty"].shape == (0, 3) and f["d/empty"][()].shape == (0, 3)
        assert f["d/scalar"][()] == 2.5 and f["d/scalar"].shape == ()
        assert "d/f64" in f and "d/nope" not in f and "lens run" in f
        with pytest.raises(KeyError):
            f["d/nope"]
        with pytest.raises(ValueError):
            f.create_group("x")                                  # opened read-only
    # append: untouched datasets are carried over without being loaded, deletions and additions take effect
    with h5.File(path, "a") as f:
        del f["d/i64"]
        f.create_group("second run").attrs["n"] = 1
    with h5.File(path, "r") as f:
        assert f.keys() == ["d", "lens run", "second run"] and "i64" not in f["d"]
        np.testing.assert_array_equal(f["d/u8"][()], np.arange(4, dtype=np.uint8))
        assert f["lens run/2022/beamline/DR aperture"].attrs["unicode"] == "größe µ"
    with pytest.raises(FileNotFoundError):
        h5.File(tmp_path / "missing.h5", "r")
    with pytest.raises(OSError):
        (tmp_path / "junk.h5").write_bytes(b"not an hdf5 file" * 10)
        h5.File(tmp_path / "junk.h5", "r")
