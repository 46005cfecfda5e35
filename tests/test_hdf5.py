"""CPU tests of the dependency-free HDF5 reader / writer (chapel/src/MyHDF5.chpl).

The reader is pinned on ``tests/golden/libhdf5_written.mat``: a file written by the real libhdf5 (a MATLAB 7.3 file from
scipy's test data, scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat, BSD licence) -- 512-byte user block, superblock 0,
symbol-table root group, version-1 object header, contiguous f64[9, 1] holding k * pi / 4.  The writer is checked through
the reader and field by field against the structures of that file.  Chunked / filtered / new-style files are built by
hand below from the published format specification (no libhdf5-written sample of those exists in this image)."""
from __future__ import annotations

import struct
import zlib
from pathlib import Path

import numpy as np
import pytest

from lattice_symmetries_b200 import hdf5 as H

GOLDEN = Path(__file__).parent / "golden" / "libhdf5_written.mat"


# ---- reading what libhdf5 wrote ---------------------------------------------------------------------------------------
def test_reads_a_file_written_by_libhdf5():
    with H.File(GOLDEN) as f:
        assert (f.superblock_version, f.O, f.L, f.base) == (0, 8, 8, 512)
        assert (f.group_leaf_k, f.group_internal_k) == (4, 16)
        assert f.datasets() == ["/testdouble"]
        assert f.exists("testdouble") and f.exists("/testdouble") and not f.exists("/nothing")
        assert f.shape("/testdouble") == (9, 1) and f.dtype("/testdouble") == np.dtype("<f8")
        a = f.read("/testdouble")
        assert a.shape == (9, 1) and np.array_equal(a[:, 0], np.arange(9) * (np.pi / 4))
        assert f.data_offset("/testdouble") == 4096
        with pytest.raises(KeyError):
            f.read("/missing")
    assert H.list_datasets(GOLDEN) == ["/testdouble"]
    assert np.array_equal(H.read_dataset(GOLDEN, "testdouble")[:, 0], np.arange(9) * (np.pi / 4))


def test_rejects_what_is_not_hdf5(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(H.Hdf5Error):
        H.File(p)
    p.write_bytes(H.SIGNATURE + bytes([9]) + bytes(200))
    with pytest.raises(H.Hdf5Error):
        H.File(p)


# ---- writing ----------------------------------------------------------------------------------------------------------
def test_round_trip_of_the_reference_datasets(tmp_path):
    rng = np.random.default_rng(0)
    reps = np.sort(rng.choice(1 << 40, size=10_000, replace=False).astype(np.uint64))
    x = rng.standard_normal((1, 10_000))
    evals = rng.standard_normal(3)
    evecs = rng.standard_normal((3, 10_000))
    p = tmp_path / "out.h5"
    H.write_file(p, {"/representatives": reps, "/x": x, "hamiltonian/eigenvalues": evals,
                     "hamiltonian/eigenvectors": evecs, "hamiltonian/residuals": np.abs(evals) * 1e-9,
                     "basis/representatives": reps})
    with H.File(p) as f:
        assert f.datasets() == ["/basis/representatives", "/hamiltonian/eigenvalues", "/hamiltonian/eigenvectors",
                                "/hamiltonian/residuals", "/representatives", "/x"]
        assert f.exists("basis/representatives") and f.exists("hamiltonian") and not f.exists("basis/x")
        got = f.read("/representatives")
        assert got.dtype == np.uint64 and np.array_equal(got, reps)
        assert np.array_equal(f.read("basis/representatives"), reps)
        assert f.shape("/x") == (1, 10_000) and np.array_equal(f.read("/x"), x)
        assert np.array_equal(f.read("hamiltonian/eigenvectors"), evecs)
        assert np.array_equal(f.read("hamiltonian/eigenvalues"), evals)
        # blocks of the last axis (readDatasetAsBlocks): only those bytes are read
        assert np.array_equal(f.read("/representatives", rows=(17, 4242)), reps[17:4242])
        assert np.array_equal(f.read("hamiltonian/eigenvectors", rows=(100, 200)), evecs[:, 100:200])
        assert f.read("/x", rows=(5, 5)).shape == (1, 0)
        with pytest.raises(IndexError):
            f.read("/x", rows=(0, 10_001))
        assert f.eof == p.stat().st_size
        for name in f.datasets():
            assert f.data_offset(name) % 4096 == 0


@pytest.mark.parametrize("dtype", ["u1", "i1", "u2", "i2", "u4", "i4", "u8", "i8", "f4", "f8", "c16"])
def test_every_element_type(tmp_path, dtype):
    rng = np.random.default_rng(1)
    a = (rng.standard_normal((2, 33)) * 100).astype(dtype)
    if dtype == "c16":
        a = a + 1j * rng.standard_normal((2, 33))
    p = tmp_path / "t.h5"
    H.write_file(p, {"a": a, "scalar": np.asarray(a[0, 0]), "empty": a[:, :0]})
    with H.File(p) as f:
        assert f.dtype("a") == np.dtype(dtype).newbyteorder("<") and np.array_equal(f.read("a"), a)
        assert f.read("scalar").shape == () and f.read("scalar") == a[0, 0]
        assert f.read("empty").shape == (2, 0)


def test_unwritable_types(tmp_path):
    for bad in (np.zeros(3, dtype="c8"), np.zeros(3, dtype="f2"), np.zeros(3, dtype="S4"), np.zeros(3, dtype=bool)):
        with pytest.raises(H.Hdf5Error):
            H.write_file(tmp_path / "bad.h5", {"a": bad})
    with pytest.raises(H.Hdf5Error):
        H.create(tmp_path / "bad.h5", {"a": H.DatasetSpec((3,), "f8"), "a/b": H.DatasetSpec((3,), "f8")})
    with pytest.raises(H.Hdf5Error):
        H.create(tmp_path / "bad.h5", {"/": H.DatasetSpec((3,), "f8")})


def test_many_links_span_several_symbol_table_nodes(tmp_path):
    arrays = {f"g/d{k:03d}": np.full(k % 5 + 1, k, dtype=np.int32) for k in range(100)}
    arrays.update({f"top{k}": np.arange(k, dtype=np.float64) for k in range(1, 12)})
    p = tmp_path / "many.h5"
    H.write_file(p, arrays)
    with H.File(p) as f:
        names = f.datasets()
        assert len(names) == 111 and names == sorted(names)
        for name, a in arrays.items():
            assert np.array_equal(f.read(name), a)
        # B-tree keys: key[i+1] is the heap offset of the LARGEST name of child i (the invariant libhdf5 searches by)
        links = f._links(f.root_header)
        msgs = dict(f._messages(links["g"]))
        tree, heap = f._offset(msgs[0x11], 0), f._offset(msgs[0x11], 8)
        node = f._at(tree, H._TREE_SIZE)
        assert node[:4] == b"TREE" and node[4] == 0 and node[5] == 0
        used = struct.unpack_from("<H", node, 6)[0]
        assert used == 13   # 100 links, 8 per symbol-table node
        last = ""
        for i in range(used):
            child = struct.unpack_from("<Q", node, 24 + 8 + 16 * i)[0]
            key = struct.unpack_from("<Q", node, 24 + 16 * (i + 1))[0]
            snod = f._at(child, H._SNOD_SIZE)
            count = struct.unpack_from("<H", snod, 6)[0]
            inside = [f._heap_name(heap, struct.unpack_from("<Q", snod, 8 + 40 * k)[0]) for k in range(count)]
            assert inside == sorted(inside) and inside[0] > last
            assert f._heap_name(heap, key) == inside[-1]
            last = inside[-1]
    with pytest.raises(H.Hdf5Error):
        H.create(tmp_path / "toomany.h5", {f"d{k}": H.DatasetSpec((1,), "f8") for k in range(257)})


def test_written_structures_match_the_libhdf5_file_field_by_field(tmp_path):
    """Same content as the golden file -> the same bytes in every structure libhdf5 would parse, addresses aside."""
    p = tmp_path / "same.h5"
    H.write_file(p, {"testdouble": (np.arange(9) * (np.pi / 4)).reshape(9, 1)})
    g, m = H.File(GOLDEN), H.File(p)
    try:
        graw, mraw = GOLDEN.read_bytes()[512:], p.read_bytes()
        # superblock: versions, sizes, K values (bytes 8..20), undefined free-space / driver addresses
        assert graw[:8] == mraw[:8] == H.SIGNATURE and graw[8:20] == mraw[8:20]
        assert mraw[24:32] == bytes(8) and mraw[32:40] == mraw[48:56] == b"\xff" * 8
        # root symbol-table entry: cache type 1 with the B-tree and heap addresses in the scratch pad
        for raw, f in ((graw, g), (mraw, m)):
            name_off, header, cache = struct.unpack_from("<QQI", raw, 56)
            tree, heap = struct.unpack_from("<QQ", raw, 80)
            assert (name_off, header, cache) == (0, f.root_header, 1)
            msg = dict(f._messages(f.root_header))[0x11]
            assert struct.unpack("<QQ", msg) == (tree, heap)
            # root object header: version 1, two messages (symbol table, flagged constant; NIL), 32 bytes of messages
            assert raw[header:header + 16] == bytes.fromhex("01000200010000002000000000000000")
            assert raw[header + 16:header + 24] == bytes.fromhex("1100100001000000") and raw[header + 40:header + 48] == bytes(8)
            # local heap: header, then "" at 0, the name at 8, ONE free block (next = 1, size = the rest)
            h = raw[heap:heap + 32]
            size, free, data = struct.unpack_from("<QQQ", h, 8)
            assert h[:8] == b"HEAP\0\0\0\0" and free == 24
            seg = raw[data:data + size]
            assert seg[:8] == bytes(8) and seg[8:24] == b"testdouble\0\0\0\0\0\0"
            assert struct.unpack_from("<QQ", seg, 24) == (1, size - 24)
            # B-tree: one leaf, keys ("" , "testdouble"); symbol-table node: one entry, cache type 0
            t = raw[tree:tree + 56]
            assert t[:8] == b"TREE\0\0\1\0" and t[8:24] == b"\xff" * 16
            key0, child, key1 = struct.unpack_from("<QQQ", t, 24)
            assert (key0, key1) == (0, 8)
            s = raw[child:child + 48]
            assert s[:8] == b"SNOD\1\0\1\0"
            assert struct.unpack_from("<QQII", s, 8)[0] == 8 and struct.unpack_from("<QQII", s, 8)[2:] == (0, 0)
            assert s[32:48] == bytes(16)
        # dataset header: the messages libhdf5 needs, byte for byte where the versions agree
        gm, mm = dict(g._messages(g._resolve("testdouble"))), dict(m._messages(m._resolve("testdouble")))
        assert gm[0x01] == mm[0x01]   # dataspace
        assert gm[0x03] == mm[0x03]   # datatype (IEEE f64, little-endian)
        assert gm[0x05] == mm[0x05]   # fill value
        assert mm[0x08][:2] == bytes([3, 1]) and struct.unpack_from("<QQ", mm[0x08], 2) == (4096, 72)   # layout 3, contiguous
        assert np.array_equal(g.read("testdouble"), m.read("testdouble"))
    finally:
        g.close()
        m.close()


def test_ranks_fill_their_own_rows(tmp_path):
    """``writeDatasetAsBlocks``: the file is created once, every rank writes its block of the last axis."""
    rng = np.random.default_rng(2)
    dim, world = 100_003, 5
    reps = np.sort(rng.choice(1 << 50, size=dim, replace=False).astype(np.uint64))
    vecs = rng.standard_normal((2, dim))
    bounds = [dim * r // world for r in range(world + 1)]
    p = tmp_path / "sharded.h5"
    offsets = H.create(p, {"basis/representatives": H.DatasetSpec((dim,), np.uint64),
                           "hamiltonian/eigenvectors": H.DatasetSpec((2, dim), np.float64)})
    assert set(offsets) == {"basis/representatives", "hamiltonian/eigenvectors"}
    assert np.array_equal(H.read_dataset(p, "basis/representatives"), np.zeros(dim, dtype=np.uint64))   # not written yet
    for r in reversed(range(world)):   # any order
        H.write_rows(p, "basis/representatives", reps[bounds[r]:bounds[r + 1]], bounds[r])
        H.write_rows(p, "hamiltonian/eigenvectors", vecs[:, bounds[r]:bounds[r + 1]], bounds[r])
    assert np.array_equal(H.read_dataset(p, "basis/representatives"), reps)
    assert np.array_equal(H.read_dataset(p, "hamiltonian/eigenvectors"), vecs)
    for r in range(world):
        assert np.array_equal(H.read_dataset(p, "hamiltonian/eigenvectors", rows=(bounds[r], bounds[r + 1])),
                              vecs[:, bounds[r]:bounds[r + 1]])
    with pytest.raises(H.Hdf5Error):
        H.write_rows(p, "basis/representatives", reps[:10], dim - 5)
    with pytest.raises(H.Hdf5Error):
        H.write_rows(p, "basis/representatives", reps[:10].astype(np.float64), 0)
    with pytest.raises(H.Hdf5Error):
        H.write_rows(p, "hamiltonian/eigenvectors", vecs[:1, :10], 0)


# ---- files this writer does not produce: built by hand from the format specification --------------------------------------
def _header_v1(messages):
    data = b"".join(struct.pack("<HHB3x", t, len(H._pad8(b)), 0) + H._pad8(b) for t, b in messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(data)) + data


def _chunked_file(path, a, chunk, deflate=False, shuffle=False, fletcher=False, superblock=0):
    """A chunked (optionally shuffled / deflated / checksummed) 2-D dataset "/c" behind a one-level chunk B-tree."""
    a = np.ascontiguousarray(a)
    item = a.dtype.itemsize
    filters = []
    if shuffle:
        filters.append((2, (item,)))
    if deflate:
        filters.append((1, (6,)))
    if fletcher:
        filters.append((3, ()))
    chunks = []
    for i in range(0, a.shape[0], chunk[0]):
        for j in range(0, a.shape[1], chunk[1]):
            block = np.zeros(chunk, dtype=a.dtype)
            part = a[i:i + chunk[0], j:j + chunk[1]]
            block[:part.shape[0], :part.shape[1]] = part
            raw = block.tobytes()
            if shuffle:
                raw = np.frombuffer(raw, dtype=np.uint8).reshape(-1, item).T.copy().tobytes()
            if deflate:
                raw = zlib.compress(raw, 6)
            if fletcher:
                raw += b"\xde\xad\xbe\xef"   # (the reader does not verify it)
            chunks.append(((i, j), raw))
    body = bytearray(4096)
    # root group with one link message (new-style compact group) when superblock 2, else a symbol table
    pos_data = 2048
    tree_addr = 1024
    keys = b""
    for (i, j), raw in chunks:
        keys += struct.pack("<IIQQQ", len(raw), 0, i, j, 0) + struct.pack("<Q", pos_data)
        body[pos_data:pos_data + len(raw)] = raw
        pos_data += len(raw) + (-len(raw) % 8)
        if pos_data > len(body) - 1024:
            body.extend(bytes(4096))
    keys += struct.pack("<IIQQQ", 0, 0, a.shape[0], a.shape[1], 0)
    tree = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(chunks), H.UNDEF, H.UNDEF) + keys
    body[tree_addr:tree_addr + len(tree)] = tree
    space = struct.pack("<BBB5x", 1, 2, 0) + struct.pack("<QQ", *a.shape)
    layout = struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", tree_addr) + struct.pack("<III", chunk[0], chunk[1], item)
    messages = [(0x01, space), (0x03, H._datatype_message(a.dtype)), (0x08, layout)]
    if filters:
        pipeline = struct.pack("<BB6x", 1, len(filters))
        for fid, vals in filters:
            pipeline += struct.pack("<HHHH", fid, 0, 0, len(vals)) + b"".join(struct.pack("<I", v) for v in vals)
            if len(vals) % 2:
                pipeline += bytes(4)
        messages.append((0x0B, pipeline))
    dset = _header_v1(messages)
    dset_addr = 512
    body[dset_addr:dset_addr + len(dset)] = dset
    if superblock == 0:
        heap_addr, heap_data, btree, snod, root = 96, 128, 224, 320, 900   # no overlap: sizes 32, 88, 48, 48, 40
        seg = bytearray(88)
        seg[8:10] = b"c\0"
        seg[16:32] = struct.pack("<QQ", 1, 72)
        body[heap_addr:heap_addr + 32] = b"HEAP" + struct.pack("<B3xQQQ", 0, 88, 16, heap_data)
        body[heap_data:heap_data + 88] = seg
        body[btree:btree + 48] = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, H.UNDEF, H.UNDEF) + struct.pack("<QQQ", 0, snod, 8)
        body[snod:snod + 48] = b"SNOD" + struct.pack("<BxH", 1, 1) + struct.pack("<QQII16x", 8, dset_addr, 0, 0)
        rooth = _header_v1([(0x11, struct.pack("<QQ", btree, heap_addr))])
        body[root:root + len(rooth)] = rooth
        body[0:96] = H.SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) + \
            struct.pack("<QQQQ", 0, H.UNDEF, len(body), H.UNDEF) + struct.pack("<QQII", 0, root, 1, 0) + \
            struct.pack("<QQ", btree, heap_addr)
    else:
        # superblock 2 + a version-2 object header for the root group holding one hard-link message
        link = struct.pack("<BB", 1, 0) + struct.pack("<B", 1) + b"c" + struct.pack("<Q", dset_addr)
        msgs = struct.pack("<BHB", 0x06, len(link), 0) + link
        root = 96
        rooth = b"OHDR" + struct.pack("<BB", 2, 0) + struct.pack("<B", len(msgs)) + msgs + bytes(4)
        body[root:root + len(rooth)] = rooth
        body[0:48] = H.SIGNATURE + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", 0, H.UNDEF, len(body), root) + bytes(4)
    Path(path).write_bytes(bytes(body))


@pytest.mark.parametrize("deflate,shuffle,fletcher", [(False, False, False), (True, False, False), (True, True, True)])
@pytest.mark.parametrize("superblock", [0, 2])
def test_reads_chunked_and_filtered_datasets(tmp_path, deflate, shuffle, fletcher, superblock):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 1 << 40, size=(5, 37), dtype=np.uint64)
    p = tmp_path / "chunked.h5"
    _chunked_file(p, a, (2, 16), deflate, shuffle, fletcher, superblock)
    with H.File(p) as f:
        assert f.superblock_version == superblock and f.datasets() == ["/c"]
        assert f.shape("c") == (5, 37)
        assert np.array_equal(f.read("c"), a)
        assert np.array_equal(f.read("c", rows=(3, 30)), a[:, 3:30])
        with pytest.raises(H.Hdf5Error):
            f.data_offset("c")


# ---- through storage.py: the reference's file names and layouts ---------------------------------------------------------
def test_storage_exports_the_reference_files(tmp_path):
    from lattice_symmetries_b200 import storage as S
    rng = np.random.default_rng(4)
    dim, world = 5000, 3
    reps = np.sort(rng.choice(1 << 36, size=dim, replace=False).astype(np.uint64))
    x = rng.standard_normal(dim)
    y = rng.standard_normal(dim)
    bounds = [dim * r // world for r in range(world + 1)]
    p = tmp_path / "golden.h5"
    for r in range(world):   # rank 0 first: it creates the file (a barrier separates the two steps in a real run)
        S.save_block_h5(p, "/representatives", reps[bounds[r]:bounds[r + 1]], bounds[r], dim, r, world)
    for r in range(world):
        S.save_block_h5(p, "/x", x[None, bounds[r]:bounds[r + 1]], bounds[r], dim, r, world)
        S.save_block_h5(p, "/y", y[None, bounds[r]:bounds[r + 1]], bounds[r], dim, r, world)
    # what chapel/test/TestStatesEnumeration.chpl:23-25 and TestMatrixVectorProduct.chpl:7-11 read
    assert np.array_equal(H.read_dataset(p, "/representatives"), reps)
    assert H.File(p).shape("/x") == (1, dim) and np.array_equal(H.read_dataset(p, "/x")[0], x)
    assert np.array_equal(H.read_dataset(p, "/y")[0], y)
    for r in range(world):
        block, lo = S.load_block_h5(p, "/representatives", r, world)
        assert lo == bounds[r] and np.array_equal(block, reps[bounds[r]:bounds[r + 1]])
        vec, lo = S.load_block_h5(p, "/x", r, world, bounds=bounds)
        assert vec.shape == (1, bounds[r + 1] - bounds[r]) and np.array_equal(vec[0], x[bounds[r]:bounds[r + 1]])
    # block -> hashed, as makeBasisStates does on a multi-locale run (Diagonalize.chpl:227-235)
    masks = S.locale_index_of(H.read_dataset(p, "/representatives"), 4)
    parts = S.block_to_hashed(H.read_dataset(p, "/x")[0], masks, 4)
    assert np.array_equal(S.hashed_to_block(parts, masks), x)
