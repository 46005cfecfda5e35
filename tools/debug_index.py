import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import helpers as H
from oracle import ls_oracle as oracle
p = H.chain10_getting_started()
ob, reps, index, off, diag = p.oracle_setup(oracle)
basis = p.product_basis(); basis.build()
for trial in range(3):
    rng = np.random.default_rng(7)
    present = reps[rng.integers(0, reps.shape[0], size=min(50000, 4 * reps.shape[0]))]
    absent = present ^ np.uint64(1)
    junk = rng.integers(0, 2 ** min(63, ob.number_bits), size=1000, dtype=np.uint64)
    needles = np.concatenate([reps[:1], reps[-1:], present, absent, junk])
    got = basis.index(needles); want = index(needles)
    bad = np.nonzero(got != want)[0]
    print("trial", trial, "mismatches", bad.size, [(int(i), int(needles[i]), int(got[i]), int(want[i])) for i in bad[:10]])
print("reps", reps)
print("lop3 peak Tops/s", __import__("lattice_symmetries_b200")._lib.lib.ls_b200_measure_lop3_peak() / 1e12)
