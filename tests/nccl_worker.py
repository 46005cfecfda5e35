"""Real ranks for tests/test_gpu_distributed.py::test_nccl_ranks -- run under torchrun, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/nccl_worker.py

Every rank builds its block of the basis through the reference-facing entry point
(``ls_hs_build_representatives`` under an active library communicator), runs both distributed product forms on
device-resident blocks and the host-pointer product (``ls_chpl_matrix_vector_product`` with this rank's blocks of
x and y), and compares its rows with the CPU oracle's full product.  Prints NCCL_WORKER_OK on rank 0.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main() -> int:
    import torch
    import torch.distributed as dist
    import lattice_symmetries_b200 as ls
    from lattice_symmetries_b200 import lattices as L
    from lattice_symmetries_b200.distributed import (ALLGATHER, ALLTOALL, DistributedOperator, init_communicator,
                                                     init_process, layout_of)
    from lattice_symmetries_b200.lanczos import lanczos_ground_state
    from oracle import ls_oracle as oracle
    import helpers as H

    os.environ["LS_B200_PROFILE"] = "1"   # kernel times per chunk of rows: what ls_b200_dist_rebalance balances with
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    init_process(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world, rank = init_communicator()
    oracle.build()
    failures = []

    def check(what, ok):
        if not ok:
            failures.append(what)
            print(f"[rank {rank}] FAILED: {what}", flush=True)

    models = [L.heisenberg_chain(24), L.kagome_heisenberg(24, spin_inversion=1), L.kagome_heisenberg(18),
              L.hubbard_square(2, 4)]
    for model in models:
        particle = 0 if model.particle == "spin-1/2" else 1
        p = H.Problem(model.name, model.number_sites, model.expression, particle=particle,
                      hamming_weight=model.hamming_weight, number_particles=model.number_particles,
                      spin_inversion=model.spin_inversion, symmetries=model.symmetries)
        ob, reps, index, off, diag = p.oracle_setup(oracle)
        dim = reps.shape[0]
        basis = model.basis()
        basis.build()                       # ls_hs_build_representatives: sharded, a collective
        lay = layout_of(basis)
        check(f"{model.name}: layout", (lay.world, lay.rank, lay.dim) == (world, rank, dim))
        lo, hi = lay.row_begin, lay.row_end
        check(f"{model.name}: local block of the sorted list", np.array_equal(np.asarray(basis.states), reps[lo:hi]))
        op = model.operator(basis)
        x = np.random.default_rng(17).standard_normal(dim)
        want = oracle.matvec(ob, off, diag, index, x)[0]
        scale = max(float(np.linalg.norm(want)), 1e-300)
        dop = DistributedOperator(op)
        xl = torch.from_numpy(x[lo:hi].copy()).cuda()
        yl = dop.empty_vector()
        for mode, label in ((ALLGATHER, "all-gather"), (ALLTOALL, "all-to-all")):
            if mode == ALLGATHER and lay.global_index == 0:
                continue
            for repeat in range(2):        # twice: buffers are reused
                yl.fill_(float("nan"))
                dop.matvec(xl, yl, mode)
                dop.sync()
                err = float(np.linalg.norm(yl.cpu().numpy() - want[lo:hi])) / scale
                check(f"{model.name}: {label} product (pass {repeat}) rel err {err:.2e}", err <= 1e-12)
        # the reference-facing call with this rank's host blocks (DistributedMatrixVector.chpl:1060-1088)
        yh = op.apply_to_state_vector(x[lo:hi].copy())
        err = float(np.linalg.norm(yh - want[lo:hi])) / scale
        check(f"{model.name}: ls_chpl_matrix_vector_product on local blocks rel err {err:.2e}", err <= 1e-12)
        # global dot product: local dot + all-reduce on the library's communicator
        d = float(dop.dot(xl, xl).item())
        check(f"{model.name}: all-reduced dot", abs(d - float(x @ x)) <= 1e-9 * float(x @ x))
        if model.name.startswith("heisenberg_chain"):
            res = lanczos_ground_state(dop, max_iters=200, tol=1e-12)
            e_ref, _ = H.oracle_ground_state_energy(oracle, p)
            check(f"{model.name}: Lanczos E0 {res.energy:.12f} vs oracle eigsh {e_ref:.12f}",
                  abs(res.energy - e_ref) <= 1e-9 * abs(e_ref))
        # re-balance by the kernel times of the last product (collective), then everything again on the new blocks
        from lattice_symmetries_b200.distributed import rebalance_distributed
        dop.matvec(xl, yl, ALLGATHER if lay.global_index else ALLTOALL)
        dop.sync()
        lay2 = rebalance_distributed(basis)
        lo, hi = lay2.row_begin, lay2.row_end
        check(f"{model.name}: block after re-balancing", np.array_equal(np.asarray(basis.states), reps[lo:hi]))
        dop = DistributedOperator(op)
        xl = torch.from_numpy(x[lo:hi].copy()).cuda()
        yl = dop.empty_vector()
        for mode, label in ((ALLGATHER, "all-gather"), (ALLTOALL, "all-to-all")):
            if mode == ALLGATHER and lay2.global_index == 0:
                continue
            dop.matvec(xl, yl, mode)
            dop.sync()
            err = float(np.linalg.norm(yl.cpu().numpy() - want[lo:hi])) / scale
            check(f"{model.name}: {label} product after re-balancing rel err {err:.2e}", err <= 1e-12)
        del dop, op, basis

    flag = torch.tensor([len(failures)], dtype=torch.int64, device="cuda")
    dist.all_reduce(flag)
    total = int(flag.item())
    if rank == 0:
        print("NCCL_WORKER_OK" if total == 0 else f"NCCL_WORKER_FAILED ({total})", flush=True)
    ls._lib.lib.ls_b200_comm_finalize()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if total == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
