"""Small driver for ncu captures: build one workload, run a few matvecs.
    ncu ... python tools/profile_workload.py kagome36 2
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from lattice_symmetries_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "kagome36"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model, desc = bench.make_model(name)
basis = model.basis()
basis.build()
op = model.operator(basis)
dim = basis.number_states
x = _lib.DeviceArray.from_numpy(np.random.default_rng(0).standard_normal(dim))
y = _lib.DeviceArray(dim, np.float64)
for _ in range(reps):
    op.matvec_device(x.ptr, y.ptr, sync=True)
ms = _lib.lib.ls_b200_last_kernel_ms
print(name, "dim", dim, "matvec ms %.2f" % ms(b"matvec"), "orbit ms %.2f x%d" % (ms(b"orbit"), ms(b"orbit_launches")),
      "gather ms %.2f x%d" % (ms(b"gather"), ms(b"gather_launches")), "combine ms %.2f" % ms(b"combine"),
      "build ms %.2f" % ms(b"build"))
# matrix elements of the launch that `ncu -s 1 -c 1` captures (the second chunk of the first product): lets bench.py turn the
# captured instruction count into instructions per matrix element
T = max(1, op.number_off_diag_terms)
chunk_rows = max(1, min(dim, (1 << 27) // T))
second = op.count_matrix_elements(min(dim, chunk_rows), min(dim, 2 * chunk_rows)) if dim > chunk_rows else op.count_matrix_elements(0, dim)
print("launch_matrix_elements", second)
