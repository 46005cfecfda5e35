"""Matrix elements per contiguous row shard: python tools/nnz_balance.py kagome36 8"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "kagome36"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
model, _ = bench.make_model(name)
basis = model.basis()
basis.build()
op = model.operator(basis)
dim = basis.number_states
chunk = -(-dim // world)
counts = [op.count_matrix_elements(min(r * chunk, dim), min((r + 1) * chunk, dim)) for r in range(world)]
total = sum(counts)
print(name, "world", world, "elements per rank / mean:", [round(c * world / total, 3) for c in counts])
