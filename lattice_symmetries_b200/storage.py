"""On-disk representatives / vectors and the block <-> hashed layout converters.

The reference keeps representatives and vectors in HDF5 files (``/representatives`` u64[dim], ``/x`` ``/y``
f64[1, dim]: chapel/test/TestStatesEnumeration.chpl:23-25, TestMatrixVectorProduct.chpl:7-11, 34, 43;
``basis/representatives`` cached by chapel/src/Diagonalize.chpl:227-256) and converts between the *block* layout of
those files (one array in sorted order) and its in-memory *hashed* layout (state -> locale by
``hash64_01(state) % numLocales``, chapel/src/StatesEnumeration.chpl:198-212) with
chapel/src/BlockToHashed.chpl:106-189 / HashedToBlock.chpl:79-150.

Two containers.  (1) The native one, for checkpoints: a little-endian raw array behind a one-line JSON header (dtype,
shape, layout, CRC-32 of the payload); a sharded basis / vector is one file per rank (``name.r3of8.lsb``), each holding
that rank's contiguous range of the sorted order -- concatenated in rank order they are the block layout, no
conversion needed.  (2) The reference's own: HDF5, through the dependency-free reader / writer in ``hdf5.py`` (there is
no libhdf5 in this image) -- ``save_block_h5`` / ``load_block_h5`` write and read the very datasets named above, every
rank its own rows of ONE file.  The converters below are only for exchanging data with a reference run that uses the
hashed layout.
"""
from __future__ import annotations

import json
import zlib
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

__all__ = [
    "MAGIC", "save_array", "load_array", "rank_path", "save_representatives", "load_representatives",
    "load_all_representatives", "save_vector", "load_vector", "hash64_01", "locale_index_of", "block_to_hashed", "hashed_to_block",
    "save_block_h5", "load_block_h5",
]

MAGIC = "lattice-symmetries-b200/array/1"
_CHUNK = 1 << 24  # elements per write / CRC step


def _le(dtype) -> np.dtype:
    return np.dtype(dtype).newbyteorder("<")


def save_array(path, array, **meta) -> Path:
    """``array`` (numpy, or anything ``np.asarray`` can view, e.g. the managed ``basis.states``) -> header line +
    raw little-endian payload; written in chunks, so a 10 GB block needs no second copy in host memory."""
    path = Path(path)
    a = np.asarray(array)
    flat = a.reshape(-1)
    crc = 0
    for lo in range(0, flat.shape[0], _CHUNK):
        crc = zlib.crc32(np.ascontiguousarray(flat[lo:lo + _CHUNK]).astype(_le(a.dtype), copy=False).tobytes(), crc)
    header = dict(meta, magic=MAGIC, dtype=np.dtype(a.dtype).name, shape=list(a.shape), crc32=crc)
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "wb") as f:
        f.write((json.dumps(header, sort_keys=True) + "\n").encode())
        for lo in range(0, flat.shape[0], _CHUNK):
            f.write(np.ascontiguousarray(flat[lo:lo + _CHUNK]).astype(_le(a.dtype), copy=False).tobytes())
    return path


def load_array(path, verify: bool = True) -> Tuple[np.ndarray, dict]:
    with open(path, "rb") as f:
        header = json.loads(f.readline().decode())
        if header.get("magic") != MAGIC:
            raise ValueError(f"{path}: not a {MAGIC} file")
        dtype = _le(header["dtype"])
        count = int(np.prod(header["shape"], dtype=np.int64)) if header["shape"] else 1
        a = np.fromfile(f, dtype=dtype, count=count)
    if a.shape[0] != count:
        raise ValueError(f"{path}: truncated ({a.shape[0]} of {count} elements)")
    if verify:
        crc = 0
        for lo in range(0, count, _CHUNK):
            crc = zlib.crc32(a[lo:lo + _CHUNK].tobytes(), crc)
        if crc != header["crc32"]:
            raise ValueError(f"{path}: checksum mismatch")
    return a.astype(np.dtype(header["dtype"]), copy=False).reshape(header["shape"]), header


def rank_path(path, rank: int, world: int) -> Path:
    """File of one rank of a sharded array: ``basis.lsb`` -> ``basis.r3of8.lsb``."""
    path = Path(path)
    return path if world == 1 else path.with_name(f"{path.stem}.r{rank}of{world}{path.suffix}")


# ---- representatives ----------------------------------------------------------------------------------------------
def save_representatives(path, basis) -> Path:
    """The (local block of the) sorted representatives of a built basis -- the payload of the reference's
    ``/representatives`` dataset.  Under an active communicator every rank writes its own file."""
    from .distributed import layout_of
    try:
        L = layout_of(basis)
        world, rank, dim, lo, hi = L.world, L.rank, L.dim, L.row_begin, L.row_end
    except ValueError:
        world, rank, dim, lo, hi = 1, 0, basis.number_states, 0, basis.number_states
    return save_array(rank_path(path, rank, world), basis.states, kind="representatives", layout="block",
                      number_sites=basis.number_sites, number_bits=basis.number_bits, dim=dim, row_begin=lo, row_end=hi,
                      world=world, rank=rank)


def load_representatives(path, rank: int = 0, world: int = 1) -> Tuple[np.ndarray, dict]:
    """One rank's block; feed it to ``basis.unchecked_set_representatives`` (kernels/reference.c:196-211)."""
    states, header = load_array(rank_path(path, rank, world))
    if header.get("kind") != "representatives" or states.dtype != np.uint64:
        raise ValueError(f"{path}: not a representatives file")
    if states.size > 1 and not bool(np.all(states[1:] > states[:-1])):
        raise ValueError(f"{path}: representatives are not strictly ascending")
    return states, header


def load_all_representatives(path, world: int) -> np.ndarray:
    """All ranks' files in rank order = the block layout of the whole basis."""
    return np.concatenate([load_representatives(path, r, world)[0] for r in range(world)])


# ---- vectors ---------------------------------------------------------------------------------------------------------
def save_vector(path, vector, row_begin: int = 0, dim: Optional[int] = None, rank: int = 0, world: int = 1) -> Path:
    """A (block of a) vector or of a batch of vectors ([k, rows], like the reference's ``/x`` f64[1, dim])."""
    v = vector.detach().cpu().numpy() if hasattr(vector, "detach") else np.asarray(vector)
    rows = v.shape[-1]
    return save_array(rank_path(path, rank, world), v, kind="vector", layout="block", row_begin=int(row_begin),
                      row_end=int(row_begin) + rows, dim=int(dim if dim is not None else rows), world=world, rank=rank)


def load_vector(path, rank: int = 0, world: int = 1) -> Tuple[np.ndarray, dict]:
    v, header = load_array(rank_path(path, rank, world))
    if header.get("kind") != "vector":
        raise ValueError(f"{path}: not a vector file")
    return v, header


# ---- the reference's hashed layout ---------------------------------------------------------------------------------
def hash64_01(x) -> np.ndarray:
    """chapel/src/StatesEnumeration.chpl:198-203 (the splitmix64 finaliser), vectorised; wraps modulo 2^64."""
    x = np.asarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def locale_index_of(states, number_locales: int) -> np.ndarray:
    """``localeIdxOf`` (StatesEnumeration.chpl:205-212): the ``masks`` array of the reference's converters."""
    if number_locales <= 1:
        return np.zeros(np.asarray(states).shape[0], dtype=np.uint8)
    return (hash64_01(states) % np.uint64(number_locales)).astype(np.uint8)


def block_to_hashed(array, masks, number_locales: int) -> List[np.ndarray]:
    """``arrFromBlockToHashed`` (BlockToHashed.chpl:106-189): element i of the block array (last axis) goes to locale
    ``masks[i]``, order preserved within a locale.  Returns one array per locale ([..., count_l])."""
    a = np.asarray(array)
    masks = np.asarray(masks)
    if a.shape[-1] != masks.shape[0]:
        raise ValueError("masks must have one entry per element of the last axis")
    return [np.ascontiguousarray(a[..., masks == l]) for l in range(number_locales)]


def hashed_to_block(parts: Sequence[np.ndarray], masks) -> np.ndarray:
    """``arrFromHashedToBlock`` (HashedToBlock.chpl:79-150): the inverse of :func:`block_to_hashed`."""
    masks = np.asarray(masks)
    first = np.asarray(parts[0])
    out = np.empty(first.shape[:-1] + (masks.shape[0],), dtype=first.dtype)
    for l, p in enumerate(parts):
        sel = masks == l
        if int(sel.sum()) != np.asarray(p).shape[-1]:
            raise ValueError(f"locale {l}: {np.asarray(p).shape[-1]} elements, masks say {int(sel.sum())}")
        out[..., sel] = p
    return out


# ---- the reference's own container: HDF5 -----------------------------------------------------------------------------
def save_block_h5(path, dataset: str, block, row_begin: int, dim: int, rank: int = 0, world: int = 1) -> None:
    """One rank's block (last axis = rows ``[row_begin, row_begin + n)``) of a dataset of the reference's HDF5 files
    (``/representatives`` u64[dim], ``/x`` f64[1, dim], ``basis/representatives``, ``hamiltonian/eigenvectors``):
    ``writeDatasetAsBlocks`` (chapel/src/MyHDF5.chpl:266-326, chapel/src/Diagonalize.chpl:227-256).  Rank 0 creates the
    file / adds the dataset at its full size, then every rank writes its own rows -- the caller puts a barrier between
    rank 0's call and the others' (one file on a shared file system; contiguous ranges in rank order ARE the block
    layout).  An existing file keeps its other datasets (it is re-laid-out: HDF5 metadata precedes the data here)."""
    from . import hdf5
    path = Path(path)
    a = block.detach().cpu().numpy() if hasattr(block, "detach") else np.asarray(block)
    name = "/".join(p for p in dataset.split("/") if p)
    if rank == 0:
        specs, keep = {}, {}
        if path.exists():
            with hdf5.File(path) as f:
                for other in f.datasets():
                    if other.strip("/") != name:
                        keep[other.strip("/")] = f.read(other)
        for other, arr in keep.items():
            specs[other] = hdf5.DatasetSpec(arr.shape, arr.dtype)
        specs[name] = hdf5.DatasetSpec(tuple(a.shape[:-1]) + (int(dim),), a.dtype)
        tmp = path.with_name(path.name + ".tmp")
        hdf5.create(tmp, specs)
        for other, arr in keep.items():
            if arr.size:
                hdf5.write_rows(tmp, other, arr, 0)
        tmp.replace(path)
    hdf5.write_rows(path, name, a, int(row_begin))


def load_block_h5(path, dataset: str, rank: int = 0, world: int = 1, bounds: Optional[Sequence[int]] = None):
    """(block, row_begin): rank ``rank``'s rows of the last axis of ``dataset`` -- ``readDatasetAsBlocks``
    (chapel/src/MyHDF5.chpl:145-215).  ``bounds`` (world + 1 row numbers) selects the row ranges of a balanced layout;
    by default the even split the reference's block distribution uses."""
    from . import hdf5
    with hdf5.File(path) as f:
        dim = f.shape(dataset)[-1]
        if bounds is None:
            bounds = [dim * r // world for r in range(world + 1)]
        if len(bounds) != world + 1 or bounds[0] != 0 or bounds[-1] != dim:
            raise ValueError("bounds must run from 0 to the number of rows")
        return f.read(dataset, rows=(int(bounds[rank]), int(bounds[rank + 1]))), int(bounds[rank])
