"""Block matvec timing: k vectors through one pass vs k single products.  python tools/ab_block.py kagome36"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from lattice_symmetries_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "kagome36"
model, desc = bench.make_model(name)
basis = model.basis()
basis.build()
op = model.operator(basis)
dim = basis.number_states
ms = _lib.lib.ls_b200_last_kernel_ms
for k in (1, 2, 4, 8):
    x = _lib.DeviceArray.from_numpy(np.random.default_rng(0).standard_normal(k * dim))
    y = _lib.DeviceArray(k * dim, np.float64)
    for _ in range(3):
        op.matvec_block_device(k, x.ptr, dim, y.ptr, dim, sync=True)
    t = ms(b"matvec")
    print(f"{name} dim {dim}: {k} vector(s) {t:.2f} ms = {t / k:.2f} ms per vector")
    del x, y
