"""CPU tests of the on-disk format and the block <-> hashed converters (SURVEY 8f-3;
chapel/src/StatesEnumeration.chpl:198-212, BlockToHashed.chpl:106-189, HashedToBlock.chpl:79-150)."""
from __future__ import annotations

import numpy as np
import pytest

from lattice_symmetries_b200 import storage as S


def _hash_scalar(x: int) -> int:
    """The reference's hash64_01, spelled out with Python integers (independent of the numpy version)."""
    M = (1 << 64) - 1
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
    return x ^ (x >> 31)


def test_hash64_01_matches_the_reference_formula():
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.array([0, 1, 2, 0xFFFFFFFFFFFFFFFF, 1 << 63], dtype=np.uint64),
                         rng.integers(0, 1 << 63, size=200, dtype=np.uint64)])
    want = np.array([_hash_scalar(int(x)) for x in xs], dtype=np.uint64)
    assert np.array_equal(S.hash64_01(xs), want)
    # splitmix64 known answers (seed 0 stream: the finaliser applied to k * 0x9E3779B97F4A7C15)
    g = 0x9E3779B97F4A7C15
    assert _hash_scalar(g) == 0xE220A8397B1DCDAF and _hash_scalar((2 * g) & ((1 << 64) - 1)) == 0x6E789E6AA1B965F4
    for locales in (1, 2, 3, 8):
        m = S.locale_index_of(xs, locales)
        assert m.dtype == np.uint8 and m.max() < locales
        if locales > 1:
            assert np.array_equal(m, (want % np.uint64(locales)).astype(np.uint8))


@pytest.mark.parametrize("locales", [1, 2, 5])
def test_block_hashed_round_trip(locales):
    rng = np.random.default_rng(1)
    states = np.sort(rng.choice(1 << 30, size=5000, replace=False).astype(np.uint64))
    masks = S.locale_index_of(states, locales)
    parts = S.block_to_hashed(states, masks, locales)
    assert sum(p.shape[0] for p in parts) == states.shape[0]
    for l, p in enumerate(parts):   # every locale holds exactly its states, still ascending
        assert np.all(S.locale_index_of(p, locales) == l) and np.all(p[1:] > p[:-1])
    assert np.array_equal(S.hashed_to_block(parts, masks), states)
    x = rng.standard_normal((3, states.shape[0]))   # a batch of vectors, like the reference's [k, dim] arrays
    xp = S.block_to_hashed(x, masks, locales)
    assert all(p.shape == (3, q.shape[0]) for p, q in zip(xp, parts))
    assert np.array_equal(S.hashed_to_block(xp, masks), x)
    with pytest.raises(ValueError):
        S.hashed_to_block([p[..., :-1] if p.shape[-1] else p for p in xp], masks)


def test_array_files_round_trip_and_detect_corruption(tmp_path):
    rng = np.random.default_rng(2)
    for a in (rng.integers(0, 1 << 62, size=1000, dtype=np.uint64), rng.standard_normal((2, 77)),
              rng.standard_normal(5) + 1j * rng.standard_normal(5), np.zeros(0, dtype=np.uint64)):
        p = S.save_array(tmp_path / "a.lsb", a, note="x")
        b, header = S.load_array(p)
        assert b.dtype == a.dtype and b.shape == a.shape and np.array_equal(a, b) and header["note"] == "x"
    p = S.save_array(tmp_path / "c.lsb", np.arange(100, dtype=np.uint64))
    raw = bytearray(p.read_bytes())
    raw[-5] ^= 0x10
    p.write_bytes(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        S.load_array(p)
    p.write_bytes(bytes(raw[:-8]))
    with pytest.raises(ValueError, match="truncated"):
        S.load_array(p)


def test_sharded_representatives_concatenate(tmp_path):
    states = np.arange(10, 1000, 3, dtype=np.uint64)
    bounds = [0, 100, 100, 250, states.shape[0]]
    world = len(bounds) - 1
    for r in range(world):
        S.save_array(S.rank_path(tmp_path / "basis.lsb", r, world), states[bounds[r]:bounds[r + 1]],
                     kind="representatives", layout="block", dim=int(states.shape[0]), row_begin=bounds[r],
                     row_end=bounds[r + 1], world=world, rank=r)
    assert S.rank_path(tmp_path / "basis.lsb", 2, 4).name == "basis.r2of4.lsb"
    assert np.array_equal(S.load_all_representatives(tmp_path / "basis.lsb", world), states)
    part, header = S.load_representatives(tmp_path / "basis.lsb", 2, world)
    assert (header["row_begin"], header["row_end"]) == (100, 250) and np.array_equal(part, states[100:250])
    S.save_array(tmp_path / "bad.lsb", states[::-1].copy(), kind="representatives")
    with pytest.raises(ValueError, match="ascending"):
        S.load_representatives(tmp_path / "bad.lsb")
    v = np.random.default_rng(3).standard_normal((1, 40))
    S.save_vector(tmp_path / "x.lsb", v, row_begin=7, dim=99, rank=1, world=2)
    w, header = S.load_vector(tmp_path / "x.lsb", 1, 2)
    assert np.array_equal(v, w) and header["row_begin"] == 7 and header["row_end"] == 47 and header["dim"] == 99
