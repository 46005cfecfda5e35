"""A dependency-free reader / writer for the HDF5 files the reference exchanges.

The reference stores its golden data and its caches in HDF5 through libhdf5 with default property lists
(chapel/src/MyHDF5.chpl:27-142 read, :219-326 create / write): ``/representatives`` u64[dim]
(chapel/test/TestStatesEnumeration.chpl:23-25), ``/x`` ``/y`` f64[1, dim] (chapel/test/TestMatrixVectorProduct.chpl:7-11,
34, 43), ``basis/representatives``, ``hamiltonian/eigenvectors|eigenvalues|residuals``
(chapel/src/Diagonalize.chpl:227-256).  There is no libhdf5 / h5py in this image, so this module implements the part of
the HDF5 file format those files use -- it is host-side I/O, not a kernel:

reading   superblock 0/1 (what libhdf5 writes by default) and 2/3; object headers version 1 and 2 with continuation
          blocks; old-style groups (symbol-table message -> B-tree v1 -> SNOD -> local heap) and compact new-style groups
          (link messages); dataspace 1/2; fixed-point, IEEE floating-point and the {r, i} compound h5py uses for complex
          numbers; compact, contiguous and chunked layouts (B-tree v1 chunk index; deflate / shuffle / fletcher32
          filters).  Dense new-style groups (fractal heaps) and every other datatype class raise ``Hdf5Error``.
writing   what ``H5Fcreate`` + ``H5Dcreate(H5P_DEFAULT)`` + ``H5Dwrite`` produce: superblock 0, version-1 object
          headers, old-style groups, contiguous little-endian datasets.  The data of a dataset is one contiguous run
          of the file, so ``create`` lays out the whole file first and ranks fill their own row ranges with plain
          positional writes (``write_rows``) -- the reference's ``writeDatasetAsBlocks`` (MyHDF5.chpl:266-326) without
          a parallel HDF5 library.

Parity: the READER is pinned on a file written by the real libhdf5 (tests/golden/libhdf5_written.mat, a MATLAB 7.3 file
from scipy's test data: user block, superblock 0, symbol-table groups, chunked + contiguous datasets).  The WRITER is
checked through the reader and structure by structure against that file; no libhdf5 is present to open its output, so
its acceptance by libhdf5 is *unpinned* (stated in DESIGN.md).
"""
from __future__ import annotations

import struct
import zlib
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

__all__ = ["Hdf5Error", "File", "read_dataset", "list_datasets", "create", "write_rows", "write_file", "DatasetSpec"]

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(ValueError):
    pass


# ======================================================================================================================
# reading
# ======================================================================================================================
class _Dataset:
    """What the object header of a dataset says (nothing is read until :meth:`File.read`)."""

    def __init__(self):
        self.shape: Tuple[int, ...] = ()
        self.dtype: Optional[np.dtype] = None
        self.layout: Optional[tuple] = None     # ("contiguous", address, size) | ("compact", bytes) | ("chunked", btree, chunk dims)
        self.filters: List[Tuple[int, Tuple[int, ...]]] = []
        self.fill: Optional[bytes] = None


class File:
    """Read access to one HDF5 file: ``File(path).read("/x")``, ``.datasets()``, ``.shape("/x")``, ``.dtype("/x")``."""

    def __init__(self, path):
        self.path = Path(path)
        self._f = open(self.path, "rb")
        try:
            self._read_superblock()
        except (struct.error, IndexError) as e:
            self._f.close()
            raise Hdf5Error(f"{self.path}: truncated or corrupt superblock") from e
        except Exception:
            self._f.close()
            raise

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- low level -------------------------------------------------------------------------------------------------
    def _at(self, address: int, size: int) -> bytes:
        self._f.seek(self.base + address)
        data = self._f.read(size)
        if len(data) != size:
            raise Hdf5Error(f"{self.path}: read past the end of the file (address {address}, {size} bytes)")
        return data

    def _offset(self, buf: bytes, pos: int) -> int:
        return int.from_bytes(buf[pos:pos + self.O], "little")

    def _length(self, buf: bytes, pos: int) -> int:
        return int.from_bytes(buf[pos:pos + self.L], "little")

    def _undefined(self, address: int) -> bool:
        return address == (1 << (8 * self.O)) - 1

    def _read_superblock(self):
        # the signature sits at 0, 512, 1024, ... (a user block precedes it in e.g. MATLAB files)
        self._f.seek(0, 2)
        size = self._f.tell()
        start = 0
        while True:
            self._f.seek(start)
            if self._f.read(8) == SIGNATURE:
                break
            start = 512 if start == 0 else start * 2
            if start >= size:
                raise Hdf5Error(f"{self.path}: not an HDF5 file")
        self._f.seek(start)
        head = self._f.read(128)
        version = head[8]
        self.superblock_version = version
        if version in (0, 1):
            self.O, self.L = head[13], head[14]
            self.group_leaf_k, self.group_internal_k = struct.unpack_from("<HH", head, 16)
            pos = 24 + (4 if version == 1 else 0)
            self.base = 0
            base = self._offset(head, pos)
            self.eof = self._offset(head, pos + 2 * self.O)
            entry = pos + 4 * self.O
            self.base = base
            self.root_header = self._offset(head, entry + self.O)
        elif version in (2, 3):
            self.O, self.L = head[9], head[10]
            self.base = 0
            base = self._offset(head, 12)
            self.eof = self._offset(head, 12 + 2 * self.O)
            self.root_header = self._offset(head, 12 + 3 * self.O)
            self.base = base
        else:
            raise Hdf5Error(f"{self.path}: superblock version {version} is not supported")
        if self.base == 0 and start != 0:
            self.base = start   # (files whose base address was left 0 behind a user block)

    # ---- object headers -----------------------------------------------------------------------------------------------
    def _messages(self, address: int) -> List[Tuple[int, bytes]]:
        """(type, body) of every header message of the object at ``address``, continuation blocks included."""
        head = self._at(address, 16)
        out: List[Tuple[int, bytes]] = []
        if head[:4] == b"OHDR":
            return self._messages_v2(address)
        if head[0] != 1:
            raise Hdf5Error(f"{self.path}: object header version {head[0]} at {address}")
        count = struct.unpack_from("<H", head, 2)[0]
        size = struct.unpack_from("<I", head, 8)[0]
        blocks = [(address + 16, size)]
        while blocks and len(out) < count:
            addr, length = blocks.pop(0)
            buf = self._at(addr, length)
            pos = 0
            while pos + 8 <= length and len(out) < count:
                mtype, msize = struct.unpack_from("<HH", buf, pos)
                body = buf[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:
                    blocks.append((self._offset(body, 0), self._length(body, self.O)))
                out.append((mtype, body))
        return out

    def _messages_v2(self, address: int) -> List[Tuple[int, bytes]]:
        head = self._at(address, 64)
        flags = head[5]
        pos = 6
        if flags & 0x20:
            pos += 16
        if flags & 0x10:
            pos += 4
        width = 1 << (flags & 3)
        chunk = int.from_bytes(head[pos:pos + width], "little")
        pos += width
        out: List[Tuple[int, bytes]] = []
        blocks = [(address + pos, chunk)]
        tracked = bool(flags & 0x04)
        while blocks:
            addr, length = blocks.pop(0)
            buf = self._at(addr, length)
            p = 0
            while p + 4 <= length:
                mtype = buf[p]
                msize = struct.unpack_from("<H", buf, p + 1)[0]
                p += 4 + (2 if tracked else 0)
                body = buf[p:p + msize]
                p += msize
                if mtype == 0x10:
                    caddr, clen = self._offset(body, 0), self._length(body, self.O)
                    blocks.append((caddr + 4, clen - 8))   # "OCHK" ... checksum
                out.append((mtype, body))
        return out

    # ---- groups -----------------------------------------------------------------------------------------------------
    def _heap_name(self, heap_address: int, offset: int) -> str:
        head = self._at(heap_address, 8 + 2 * self.L + self.O)
        if head[:4] != b"HEAP":
            raise Hdf5Error(f"{self.path}: local heap signature missing at {heap_address}")
        size = self._length(head, 8)
        data = self._at(self._offset(head, 8 + 2 * self.L), size)
        end = data.index(b"\0", offset)
        return data[offset:end].decode("utf-8")

    def _group_btree(self, address: int, heap: int, out: Dict[str, int]):
        node = self._at(address, 8 + 2 * self.O)
        if node[:4] == b"SNOD":
            count = struct.unpack_from("<H", node, 6)[0]
            entry = 2 * self.O + 24
            body = self._at(address + 8, count * entry)
            for i in range(count):
                name = self._heap_name(heap, self._offset(body, i * entry))
                out[name] = self._offset(body, i * entry + self.O)
            return
        if node[:4] != b"TREE" or node[4] != 0:
            raise Hdf5Error(f"{self.path}: group B-tree node expected at {address}")
        used = struct.unpack_from("<H", node, 6)[0]
        body = self._at(address + 8 + 2 * self.O, (used + 1) * self.L + used * self.O)
        for i in range(used):
            child = self._offset(body, (i + 1) * self.L + i * self.O)
            self._group_btree(child, heap, out)

    def _links(self, address: int) -> Dict[str, int]:
        """name -> object header address of the members of the group at ``address``."""
        out: Dict[str, int] = {}
        for mtype, body in self._messages(address):
            if mtype == 0x11:   # symbol table: B-tree + local heap
                self._group_btree(self._offset(body, 0), self._offset(body, self.O), out)
            elif mtype == 0x06:   # link message (compact new-style group)
                flags = body[1]
                pos = 2
                ltype = 0
                if flags & 0x08:
                    ltype = body[pos]
                    pos += 1
                if flags & 0x04:
                    pos += 8
                if flags & 0x10:
                    pos += 1
                width = 1 << (flags & 3)
                n = int.from_bytes(body[pos:pos + width], "little")
                pos += width
                name = body[pos:pos + n].decode("utf-8")
                pos += n
                if ltype == 0:
                    out[name] = self._offset(body, pos)
            elif mtype == 0x02:   # link info: a fractal heap address means dense storage
                flags = body[1]
                pos = 2 + (8 if flags & 1 else 0)
                if not self._undefined(self._offset(body, pos)):
                    raise Hdf5Error(f"{self.path}: dense (fractal-heap) groups are not supported")
        return out

    def _resolve(self, name: str) -> int:
        parts = [p for p in name.split("/") if p]
        address = self.root_header
        for i, part in enumerate(parts):
            links = self._links(address)
            if part not in links:
                raise KeyError(f"{self.path}: no object '{'/'.join(parts[:i + 1])}'")
            address = links[part]
        return address

    def exists(self, name: str) -> bool:
        """``doesObjectExist`` of the reference (chapel/src/Diagonalize.chpl:228)."""
        try:
            self._resolve(name)
            return True
        except KeyError:
            return False

    def datasets(self, group: str = "/") -> List[str]:
        """Full names of all datasets below ``group`` (depth first, sorted by name)."""
        out: List[str] = []
        seen = set()

        def walk(prefix: str, address: int):
            if address in seen:
                return
            seen.add(address)
            for name, child in sorted(self._links(address).items()):
                types = {t for t, _ in self._messages(child)}
                if 0x08 in types:
                    out.append(f"{prefix}/{name}")
                elif types & {0x11, 0x02, 0x06}:
                    walk(f"{prefix}/{name}", child)

        prefix = "/" + "/".join(p for p in group.split("/") if p)
        walk("" if prefix == "/" else prefix, self._resolve(group))
        return out

    # ---- datasets -----------------------------------------------------------------------------------------------------
    def _datatype(self, body: bytes) -> Tuple[np.dtype, int]:
        """(numpy dtype, bytes consumed)."""
        cls, version = body[0] & 0x0F, body[0] >> 4
        bits = body[1] | (body[2] << 8) | (body[3] << 16)
        size = struct.unpack_from("<I", body, 4)[0]
        order = ">" if bits & 1 else "<"
        if cls == 0:
            offset, precision = struct.unpack_from("<HH", body, 8)
            if offset != 0 or precision != 8 * size or size not in (1, 2, 4, 8):
                raise Hdf5Error(f"{self.path}: fixed-point type with padding bits is not supported")
            return np.dtype(f"{order}{'i' if bits & 8 else 'u'}{size}"), 12
        if cls == 1:
            offset, precision, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", body, 8)
            if (size, precision, eloc, esize, mloc, msize, bias) == (8, 64, 52, 11, 0, 52, 1023):
                return np.dtype(f"{order}f8"), 20
            if (size, precision, eloc, esize, mloc, msize, bias) == (4, 32, 23, 8, 0, 23, 127):
                return np.dtype(f"{order}f4"), 20
            raise Hdf5Error(f"{self.path}: non-IEEE floating-point type")
        if cls == 6:
            members = bits & 0xFFFF
            pos = 8
            fields = []
            for _ in range(members):
                end = body.index(b"\0", pos)
                mname = body[pos:end].decode("utf-8")
                if version < 3:
                    pos += (end - pos + 8) // 8 * 8
                    moffset = struct.unpack_from("<I", body, pos)[0]
                    pos += 4
                    if version == 1:
                        pos += 28   # dimensionality, reserved, permutation, reserved, 4 dimension sizes
                else:
                    pos = end + 1
                    width = max(1, (max(size - 1, 1).bit_length() + 7) // 8)
                    moffset = int.from_bytes(body[pos:pos + width], "little")
                    pos += width
                mtype, used = self._datatype(body[pos:])
                pos += used
                fields.append((mname, mtype, moffset))
            names = [f[0] for f in fields]
            if names in (["r", "i"], ["real", "imag"]) and fields[0][1] == fields[1][1] and fields[0][1].kind == "f" \
                    and fields[0][2] == 0 and fields[1][2] == fields[0][1].itemsize and size == 2 * fields[0][1].itemsize:
                return np.dtype(f"{fields[0][1].byteorder if fields[0][1].byteorder != '=' else '<'}c{size}"), pos
            return np.dtype({"names": names, "formats": [f[1] for f in fields], "offsets": [f[2] for f in fields],
                             "itemsize": size}), pos
        raise Hdf5Error(f"{self.path}: datatype class {cls} is not supported")

    def _dataset(self, name: str) -> _Dataset:
        d = _Dataset()
        have_space = False
        for mtype, body in self._messages(self._resolve(name)):
            if mtype == 0x01:
                version, rank, flags = body[0], body[1], body[2]
                pos = 8 if version == 1 else 4
                if version == 2 and body[3] == 2:
                    raise Hdf5Error(f"{self.path}: '{name}' has a null dataspace")
                d.shape = tuple(self._length(body, pos + i * self.L) for i in range(rank))
                have_space = True
            elif mtype == 0x03:
                d.dtype, _ = self._datatype(body)
            elif mtype == 0x08:
                version = body[0]
                if version == 3:
                    cls = body[1]
                    if cls == 0:
                        n = struct.unpack_from("<H", body, 2)[0]
                        d.layout = ("compact", body[4:4 + n])
                    elif cls == 1:
                        d.layout = ("contiguous", self._offset(body, 2), self._length(body, 2 + self.O))
                    elif cls == 2:
                        nd = body[2]
                        tree = self._offset(body, 3)
                        dims = struct.unpack_from(f"<{nd}I", body, 3 + self.O)
                        d.layout = ("chunked", tree, tuple(dims))
                    else:
                        raise Hdf5Error(f"{self.path}: layout class {cls}")
                elif version in (1, 2):
                    nd, cls = body[1], body[2]
                    pos = 8
                    address = None
                    if cls != 0:
                        address = self._offset(body, pos)
                        pos += self.O
                    dims = struct.unpack_from(f"<{nd}I", body, pos)
                    pos += 4 * nd
                    if cls == 0:
                        n = struct.unpack_from("<I", body, pos)[0]
                        d.layout = ("compact", body[pos + 4:pos + 4 + n])
                    elif cls == 1:
                        d.layout = ("contiguous", address, None)
                    else:
                        d.layout = ("chunked", address, tuple(dims))
                else:
                    raise Hdf5Error(f"{self.path}: data layout message version {version} is not supported")
            elif mtype == 0x0B:
                version, count = body[0], body[1]
                pos = 8 if version == 1 else 2
                for _ in range(count):
                    fid = struct.unpack_from("<H", body, pos)[0]
                    pos += 2
                    nlen = 0
                    if version == 1 or fid >= 256:
                        nlen = struct.unpack_from("<H", body, pos)[0]
                        pos += 2
                    _flags, nvals = struct.unpack_from("<HH", body, pos)
                    pos += 4
                    pos += (nlen + 7) // 8 * 8 if version == 1 else nlen
                    vals = struct.unpack_from(f"<{nvals}I", body, pos)
                    pos += 4 * nvals
                    if version == 1 and nvals % 2:
                        pos += 4
                    d.filters.append((fid, tuple(vals)))
            elif mtype == 0x05 and len(body) >= 4:
                version = body[0]
                if version in (1, 2):
                    defined = body[3]
                    if (version == 1 or defined) and len(body) >= 8:
                        n = struct.unpack_from("<I", body, 4)[0]
                        if n:
                            d.fill = body[8:8 + n]
                elif version == 3 and body[1] & 0x20:
                    n = struct.unpack_from("<I", body, 2)[0]
                    d.fill = body[6:6 + n]
        if not have_space or d.dtype is None or d.layout is None:
            raise Hdf5Error(f"{self.path}: '{name}' is not a dataset")
        return d

    def shape(self, name: str) -> Tuple[int, ...]:
        return self._dataset(name).shape

    def dtype(self, name: str) -> np.dtype:
        return self._dataset(name).dtype

    def data_offset(self, name: str) -> int:
        """Byte position in the FILE of the first element of a contiguous dataset (``write_rows`` uses it)."""
        d = self._dataset(name)
        if d.layout[0] != "contiguous" or self._undefined(d.layout[1]):
            raise Hdf5Error(f"{self.path}: '{name}' is not stored contiguously")
        return self.base + d.layout[1]

    def _unfilter(self, raw: bytes, d: _Dataset, mask: int) -> bytes:
        for k in range(len(d.filters) - 1, -1, -1):
            if mask & (1 << k):
                continue
            fid, vals = d.filters[k]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                width = vals[0] if vals else d.dtype.itemsize
                n = len(raw) // width
                a = np.frombuffer(raw[:n * width], dtype=np.uint8).reshape(width, n).T
                raw = np.ascontiguousarray(a).tobytes() + raw[n * width:]
            elif fid == 3:
                raw = raw[:-4]
            else:
                raise Hdf5Error(f"{self.path}: filter {fid} is not supported")
        return raw

    def _chunks(self, address: int, nd: int) -> Iterable[Tuple[Tuple[int, ...], int, int, int]]:
        """(offsets, filter mask, address, size) of every stored chunk below the B-tree node at ``address``."""
        node = self._at(address, 8 + 2 * self.O)
        if node[:4] != b"TREE" or node[4] != 1:
            raise Hdf5Error(f"{self.path}: chunk B-tree node expected at {address}")
        level = node[5]
        used = struct.unpack_from("<H", node, 6)[0]
        key = 8 + 8 * nd
        body = self._at(address + 8 + 2 * self.O, (used + 1) * key + used * self.O)
        for i in range(used):
            pos = i * (key + self.O)
            size, mask = struct.unpack_from("<II", body, pos)
            offsets = struct.unpack_from(f"<{nd}Q", body, pos + 8)
            child = self._offset(body, pos + key)
            if level > 0:
                yield from self._chunks(child, nd)
            else:
                yield offsets[:-1], mask, child, size

    def read(self, name: str, rows: Optional[Tuple[int, int]] = None) -> np.ndarray:
        """The whole dataset, or ``rows = (begin, end)`` of its LAST axis (the reference reads blocks of the last axis:
        ``readDatasetAsBlocks``, MyHDF5.chpl:145-215) -- for contiguous 1-D / [k, dim] data only the bytes needed are read."""
        d = self._dataset(name)
        dtype = d.dtype
        count = int(np.prod(d.shape, dtype=np.int64)) if d.shape else 1
        kind = d.layout[0]
        if kind == "contiguous" and rows is not None and len(d.shape) >= 1 and not self._undefined(d.layout[1]):
            lo, hi = rows
            if not 0 <= lo <= hi <= d.shape[-1]:
                raise IndexError(f"rows {rows} outside [0, {d.shape[-1]}]")
            lead = int(np.prod(d.shape[:-1], dtype=np.int64)) if len(d.shape) > 1 else 1
            out = np.empty((lead, hi - lo), dtype=dtype)
            for k in range(lead):
                self._f.seek(self.base + d.layout[1] + (k * d.shape[-1] + lo) * dtype.itemsize)
                got = np.fromfile(self._f, dtype=dtype, count=hi - lo)
                if got.shape[0] != hi - lo:
                    raise Hdf5Error(f"{self.path}: '{name}' is truncated")
                out[k] = got
            return out.reshape(d.shape[:-1] + (hi - lo,)).astype(dtype.newbyteorder("="), copy=False)
        if kind == "contiguous":
            if self._undefined(d.layout[1]):   # never written: the fill value
                a = np.zeros(count, dtype=dtype)
                if d.fill:
                    a[:] = np.frombuffer(d.fill, dtype=dtype, count=1)[0]
            else:
                self._f.seek(self.base + d.layout[1])
                a = np.fromfile(self._f, dtype=dtype, count=count)
                if a.shape[0] != count:
                    raise Hdf5Error(f"{self.path}: '{name}' is truncated")
        elif kind == "compact":
            a = np.frombuffer(d.layout[1], dtype=dtype, count=count).copy()
        else:
            tree, dims = d.layout[1], d.layout[2]
            chunk = dims[:-1]
            if len(chunk) != len(d.shape):
                raise Hdf5Error(f"{self.path}: chunk rank does not match the dataspace of '{name}'")
            a = np.zeros(d.shape, dtype=dtype)
            if d.fill:
                a[...] = np.frombuffer(d.fill, dtype=dtype, count=1)[0]
            if not self._undefined(tree):
                for offsets, mask, address, size in self._chunks(tree, len(dims)):
                    raw = self._unfilter(self._at(address, size), d, mask)
                    block = np.frombuffer(raw, dtype=dtype, count=int(np.prod(chunk, dtype=np.int64))).reshape(chunk)
                    sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offsets, chunk, d.shape))
                    a[sel] = block[tuple(slice(0, s.stop - s.start) for s in sel)]
            a = a.reshape(-1)
        a = a.reshape(d.shape).astype(dtype.newbyteorder("="), copy=False)
        if rows is not None:
            a = a[..., rows[0]:rows[1]]
        return a


def read_dataset(path, name: str, rows: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """``readDataset`` / ``readDatasetAsBlocks`` of the reference (MyHDF5.chpl:64-215)."""
    with File(path) as f:
        return f.read(name, rows)


def list_datasets(path) -> List[str]:
    with File(path) as f:
        return f.datasets()


# ======================================================================================================================
# writing
# ======================================================================================================================
class DatasetSpec:
    """Shape and element type of a dataset to create."""

    def __init__(self, shape: Sequence[int], dtype):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        if self.dtype.kind not in "iufc" or self.dtype.itemsize not in (1, 2, 4, 8, 16) or \
                (self.dtype.kind == "c" and self.dtype.itemsize != 16) or (self.dtype.kind == "f" and self.dtype.itemsize < 4):
            raise Hdf5Error(f"element type {self.dtype} cannot be written")

    @property
    def nbytes(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize if self.shape else self.dtype.itemsize


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _float_type(size: int) -> bytes:
    if size == 8:
        return struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)


def _datatype_message(dtype: np.dtype) -> bytes:
    if dtype.kind in "iu":
        return struct.pack("<BBBBI", 0x10, 0x08 if dtype.kind == "i" else 0x00, 0, 0, dtype.itemsize) + \
            struct.pack("<HH", 0, 8 * dtype.itemsize)
    if dtype.kind == "f":
        return _float_type(dtype.itemsize)
    # complex128: the compound {r: f64 @ 0, i: f64 @ 8} h5py reads as complex (datatype version 1)
    out = struct.pack("<BBBBI", 0x16, 2, 0, 0, 16)
    for name, offset in ((b"r", 0), (b"i", 8)):
        out += _pad8(name + b"\0") + struct.pack("<IB3xI4x4I", offset, 0, 0, 0, 0, 0, 0) + _float_type(8)
    return out


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages: List[bytes]) -> bytes:
    data = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(data)) + data


class _Group:
    def __init__(self):
        self.children: Dict[str, Union["_Group", DatasetSpec]] = {}
        self.header = self.btree = self.heap = self.heap_data = 0
        self.snods: List[int] = []
        self.name_offsets: Dict[str, int] = {}
        self.heap_size = 0


_LEAF_K = 4
_INTERNAL_K = 16
_SNOD_SIZE = 8 + 2 * _LEAF_K * 40
_TREE_SIZE = 24 + (2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8
_GROUP_HEADER_SIZE = 16 + (8 + 16) + 8   # prefix + symbol-table message + an empty NIL message, as libhdf5 lays it out


def _dataset_header(spec: DatasetSpec, data_address: int) -> bytes:
    rank = len(spec.shape)
    space = struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", s) for s in spec.shape)
    # byte for byte what libhdf5 put into the golden file: version 1, allocation time "late", fill time "if set",
    # fill value defined with size 0 (= the default, zero)
    fill = struct.pack("<BBBBI", 1, 2, 2, 1, 0)
    layout = struct.pack("<BBQQ", 3, 1, data_address, spec.nbytes)
    return _object_header([_message(0x01, space), _message(0x03, _datatype_message(spec.dtype), 1),
                           _message(0x05, fill, 1), _message(0x08, layout)])


def create(path, datasets: Dict[str, DatasetSpec], align: int = 4096) -> Dict[str, int]:
    """Create ``path`` holding the (empty, zero-filled) datasets ``{"basis/representatives": DatasetSpec(...)}``; groups
    are created as needed (``H5Fcreate`` + ``H5Dcreate``, MyHDF5.chpl:283-326).  Returns name -> byte position of the
    dataset's first element in the file, for :func:`write_rows`.  The file is laid out completely here: metadata first,
    then every dataset's contiguous data on an ``align`` boundary."""
    root = _Group()
    flat: List[Tuple[str, DatasetSpec]] = []
    for name, spec in datasets.items():
        parts = [p for p in name.split("/") if p]
        if not parts:
            raise Hdf5Error("empty dataset name")
        g = root
        for part in parts[:-1]:
            child = g.children.setdefault(part, _Group())
            if not isinstance(child, _Group):
                raise Hdf5Error(f"'{part}' is both a dataset and a group")
            g = child
        if parts[-1] in g.children:
            raise Hdf5Error(f"'{name}' defined twice")
        g.children[parts[-1]] = spec
        flat.append(("/".join(parts), spec))

    # ---- addresses ----
    pos = 96   # superblock 0 (56 bytes) + root symbol-table entry (40 bytes)
    groups: List[_Group] = []

    def place_group(g: _Group):
        nonlocal pos
        if len(g.children) > 2 * _LEAF_K * 2 * _INTERNAL_K:
            raise Hdf5Error("more than 256 links in one group")
        groups.append(g)
        g.header = pos
        pos += _GROUP_HEADER_SIZE
        g.btree = pos
        pos += _TREE_SIZE
        g.heap = pos
        pos += 32
        names = sorted(g.children, key=lambda s: s.encode("utf-8"))
        offset = 8   # offset 0: the empty string (padded to 8)
        for n in names:
            g.name_offsets[n] = offset
            offset += len(_pad8(n.encode("utf-8") + b"\0"))
        g.heap_size = max(88, offset + 16)   # room for one free block at the end
        g.heap_size += -g.heap_size % 8
        g.heap_data = pos
        pos += g.heap_size
        for _ in range(max(1, (len(names) + 2 * _LEAF_K - 1) // (2 * _LEAF_K))):
            g.snods.append(pos)
            pos += _SNOD_SIZE
        for n in names:
            if isinstance(g.children[n], _Group):
                place_group(g.children[n])

    place_group(root)
    headers: Dict[int, int] = {}
    for g in groups:
        for n, c in g.children.items():
            if isinstance(c, DatasetSpec):
                headers[id(c)] = pos
                pos += len(_dataset_header(c, 0))
    meta_end = pos
    data: Dict[int, int] = {}
    for _, spec in flat:
        pos += -pos % align
        data[id(spec)] = pos
        pos += spec.nbytes
    eof = pos

    # ---- bytes ----
    out = bytearray(meta_end)

    def put(address: int, b: bytes):
        out[address:address + len(b)] = b

    def entry(name_offset: int, child, g: Optional[_Group] = None) -> bytes:
        if isinstance(child, _Group):
            return struct.pack("<QQII", name_offset, child.header, 1, 0) + struct.pack("<QQ", child.btree, child.heap)
        return struct.pack("<QQII16x", name_offset, headers[id(child)], 0, 0)

    put(0, SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INTERNAL_K, 0) +
        struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) + entry(0, root))
    for g in groups:
        put(g.header, _object_header([_message(0x11, struct.pack("<QQ", g.btree, g.heap), 1), _message(0x00, b"")]))
        names = sorted(g.children, key=lambda s: s.encode("utf-8"))
        heap = bytearray(g.heap_size)
        used = 8
        for n in names:
            b = _pad8(n.encode("utf-8") + b"\0")
            heap[g.name_offsets[n]:g.name_offsets[n] + len(b)] = b
            used = g.name_offsets[n] + len(b)
        free = g.heap_size - used
        heap[used:used + 16] = struct.pack("<QQ", 1, free)   # the only free block: next = 1 (none), its size
        put(g.heap, b"HEAP" + struct.pack("<B3xQQQ", 0, g.heap_size, used, g.heap_data))
        put(g.heap_data, bytes(heap))
        per = 2 * _LEAF_K
        leaves = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        keys = [0]
        for leaf, address in zip(leaves, g.snods):
            put(address, b"SNOD" + struct.pack("<BxH", 1, len(leaf)) +
                b"".join(entry(g.name_offsets[n], g.children[n]) for n in leaf))
            keys.append(g.name_offsets[leaf[-1]] if leaf else 0)
        used_entries = len(leaves) if names else 0
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, used_entries, UNDEF, UNDEF)
        for i in range(used_entries):
            tree += struct.pack("<QQ", keys[i], g.snods[i])
        tree += struct.pack("<Q", keys[used_entries] if used_entries else 0)
        put(g.btree, tree)
        for n, c in g.children.items():
            if isinstance(c, DatasetSpec):
                put(headers[id(c)], _dataset_header(c, data[id(c)]))
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "wb") as f:
        f.write(bytes(out))
        f.truncate(eof)   # datasets read as zeros until written (sparse on most file systems)
    return {name: data[id(spec)] for name, spec in flat}


def write_rows(path, name: str, array, row_begin: int = 0) -> None:
    """Write ``array`` ([..., rows]) into columns ``[row_begin, row_begin + rows)`` of the LAST axis of the existing
    contiguous dataset ``name``: what one locale's hyperslab write does in ``writeDatasetAsBlocks``
    (MyHDF5.chpl:217-262).  Ranks may call it concurrently on their own row ranges."""
    a = np.asarray(array)
    with File(path) as f:
        shape, dtype, start = f.shape(name), f.dtype(name), f.data_offset(name)
    if a.ndim == 0 and shape == ():
        a, shape = a.reshape(1), (1,)
    if a.ndim == 0 or tuple(a.shape[:-1]) != tuple(shape[:-1]) or row_begin < 0 or row_begin + a.shape[-1] > shape[-1]:
        raise Hdf5Error(f"a block of shape {a.shape} at row {row_begin} does not fit '{name}' of shape {shape}")
    if np.dtype(a.dtype).newbyteorder("=") != dtype.newbyteorder("="):
        raise Hdf5Error(f"'{name}' holds {dtype}, not {a.dtype}")
    lead = int(np.prod(shape[:-1], dtype=np.int64)) if len(shape) > 1 else 1
    a2 = a.reshape(lead, a.shape[-1])
    chunk = 1 << 24
    with open(path, "r+b") as f:
        for k in range(lead):
            f.seek(start + (k * shape[-1] + row_begin) * dtype.itemsize)
            for lo in range(0, a2.shape[1], chunk):
                f.write(np.ascontiguousarray(a2[k, lo:lo + chunk]).astype(dtype, copy=False).tobytes())


def write_file(path, arrays: Dict[str, np.ndarray]) -> None:
    """Create ``path`` with the given datasets and write them whole (``writeDataset``, MyHDF5.chpl:283-310)."""
    specs = {name: DatasetSpec(np.asarray(a).shape, np.asarray(a).dtype) for name, a in arrays.items()}
    create(path, specs)
    for name, a in arrays.items():
        if np.asarray(a).size:
            write_rows(path, name, a, 0)
