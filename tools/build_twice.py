"""Build the workload's basis several times in one process: cold vs warm allocation behaviour."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from lattice_symmetries_b200 import _lib  # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "kagome36"
model, _ = bench.make_model(name)
times = []
for i in range(4):
    basis = model.basis()
    t0 = time.perf_counter()
    basis.build()
    times.append((time.perf_counter() - t0) * 1e3)
    del basis
print(name, "build wall ms:", ["%.1f" % t for t in times])
