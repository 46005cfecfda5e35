"""Ground state of the 36-site kagome Heisenberg antiferromagnet (Gamma, A1, spin-inversion-even sector) by on-device
Lanczos: python tools/kagome36_ground_state.py.  Literature: E0/N = -0.438377 J per site in the S.S convention
(Leung & Elser 1993; Waldtmann et al. 1998); the model here is written with Pauli matrices, sigma.sigma = 4 S.S."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from lattice_symmetries_b200.lanczos import lanczos_ground_state  # noqa: E402

import os  # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "kagome36"
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
if world > 1:  # torchrun: rows sharded over the ranks, vectors replicated by all-gather
    import torch
    import torch.distributed as dist
    from lattice_symmetries_b200.distributed import build_sharded, init_process
    init_process(int(os.environ.get("LOCAL_RANK", "0")))  # select this rank's GPU BEFORE the library touches a device
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
model, desc = bench.make_model(name)
basis = model.basis()
t0 = time.perf_counter()
if world > 1:
    build_sharded(basis)
else:
    basis.build()
t1 = time.perf_counter()
op = model.operator(basis)
res = lanczos_ground_state(op, max_iters=400, tol=1e-10)
t2 = time.perf_counter()
n = model.number_sites
if rank == 0:
    print(f"{world} GPU(s)")
if rank == 0:
    print(f"{desc}: dim {basis.number_states}, build {t1 - t0:.2f} s, Lanczos {res.iterations} iterations in {t2 - t1:.2f} s")
if rank == 0:
    print(f"E0 = {res.energy:.10f}  (converged={res.converged}, residual {res.residual:.2e});  E0 / (4 N) = {res.energy / (4 * n):.8f} per site")
