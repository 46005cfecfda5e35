"""Ground state of a bench workload by on-device Lanczos, on one GPU or sharded over the ranks of a torchrun launch:

    python tools/ground_state.py kagome36
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/ground_state.py kagome42 \
        --time-limit 420 --tol 1e-9 --out gpurun_out/kagome42.json

Under torchrun the representatives and all vectors are sharded (csrc/dist.cu; NCCL inside the library) -- the build
is ``ls_hs_build_representatives`` and every product is the distributed form chosen by --mode.
Literature (S.S convention): kagome-36 E0/N = -0.438377 (Leung & Elser 1993; Waldtmann et al. 1998);
kagome-42 E0/N = -0.438143 (Laeuchli, Sudan, Moessner 2019).  The models here are written with Pauli matrices,
sigma.sigma = 4 S.S, hence E0 / (4 N)."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="kagome36")
    ap.add_argument("--mode", default="auto", choices=["auto", "allgather", "alltoall"])
    ap.add_argument("--flags", type=int, default=0, help="ls_b200_dist_build flags (1 no replicated index, 2 wide index, 4 even rows)")
    ap.add_argument("--max-iters", type=int, default=400)
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--time-limit", type=float, default=None, help="seconds of Lanczos wall time")
    ap.add_argument("--matvecs", type=int, default=3, help="timed products before the Lanczos run")
    ap.add_argument("--no-lanczos", action="store_true")
    ap.add_argument("--rebalance", action="store_true",
                    help="after the timed products, move the row boundaries by the measured kernel times (sets LS_B200_PROFILE)")
    ap.add_argument("--energy-tol", type=float, default=None, help="stop when E0 moved by less than this (relative) over two checks")
    ap.add_argument("--check-every", type=int, default=10)
    ap.add_argument("--checkpoint", default=None, help="path prefix: Lanczos vectors + coefficients are saved there")
    ap.add_argument("--checkpoint-every", type=int, default=0)
    ap.add_argument("--resume", action="store_true", help="continue from the checkpoint (same layout: do not combine with --rebalance unless the checkpoint was written after re-balancing in the same way)")
    ap.add_argument("--also-mode", default=None, choices=["allgather", "alltoall"],
                    help="time the products in a second form as well (same basis, same vector)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    if args.rebalance:
        os.environ["LS_B200_PROFILE"] = "1"
    import torch
    import torch.distributed as dist
    from lattice_symmetries_b200 import _lib
    from lattice_symmetries_b200.distributed import (ALLGATHER, ALLTOALL, AUTO, build_distributed, hashed_vector,
                                                     init_communicator, init_process)
    from lattice_symmetries_b200.lanczos import _wrap, lanczos_ground_state
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    init_process(local_rank)   # select this rank's GPU BEFORE the library touches a device
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        init_communicator()
    mode = {"auto": AUTO, "allgather": ALLGATHER, "alltoall": ALLTOALL}[args.mode]
    lib = _lib.lib

    def say(msg):
        if rank == 0:
            print(msg, flush=True)

    def mem():
        free, total = torch.cuda.mem_get_info()
        return f"{(total - free) / 2**30:.1f} GiB used of {total / 2**30:.0f}"

    def barrier():
        torch.cuda.current_stream().synchronize()
        if world > 1:
            dist.barrier()

    if args.workload.startswith("chaint"):   # translations x spin inversion only: the large-footprint probe
        from lattice_symmetries_b200 import lattices
        n = int(args.workload[6:])
        model = lattices.heisenberg_chain(n, parity_sector=None, spin_inversion=1)
        desc = f"Heisenberg chain N={n}, Sz=0, translations x spin inversion"
    else:
        model, desc = bench.make_model(args.workload)
    basis = model.basis()
    op = model.operator(basis)
    say(f"{world} GPU(s): {desc}; {basis.number_candidates} candidates")
    barrier()
    t0 = time.perf_counter()
    if world > 1:
        lay = build_distributed(basis, balance_for=op, flags=args.flags)
    else:
        basis.build()
    barrier()
    t_build = time.perf_counter() - t0
    sh = _wrap(op)
    L = sh.layout
    say(f"build: dim {L.dim} in {t_build:.2f} s ({basis.number_candidates / t_build:.3e} candidates/s); rank 0 holds rows "
        f"[{L.row_begin}, {L.row_end}); replicated index kind {L.global_index}; {mem()}")
    nnz = op.count_matrix_elements(0, L.rows)
    t = torch.tensor([nnz], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t)
    elements = int(t.item()) + L.dim
    result = {"workload": args.workload, "description": desc, "gpus": world, "dim": L.dim, "candidates": basis.number_candidates,
              "build_s": t_build, "matrix_elements": elements, "mode": args.mode, "bounds": L.bounds}

    x = hashed_vector(L.row_begin, L.row_end, 42)
    y = sh.empty_vector()
    if args.also_mode is not None:
        other = {"allgather": ALLGATHER, "alltoall": ALLTOALL}[args.also_mode]
        for k in range(max(1, args.matvecs - 1)):
            barrier()
            t0 = time.perf_counter()
            sh.matvec(x, y, other)
            sh.sync()
            barrier()
            dt = time.perf_counter() - t0
            say(f"matvec {k} ({args.also_mode}): {dt * 1e3:.1f} ms = {elements / dt:.3e} matrix-elements/s; {mem()}")
        result["also_mode"] = {"mode": args.also_mode, "ms": dt * 1e3, "x_H_x": float(sh.dot(x, y).item())}
    times = []
    for k in range(args.matvecs):
        barrier()
        t0 = time.perf_counter()
        sh.matvec(x, y, mode)
        t_enqueue = time.perf_counter() - t0
        sh.sync()
        t_local = time.perf_counter() - t0
        barrier()
        times.append(time.perf_counter() - t0)
        parts = {k2: round(lib.ls_b200_last_kernel_ms(k2.encode()), 1) for k2 in
                 ("matvec", "count", "orbit", "gather", "combine", "allgather")}
        parts["enqueue"] = round(t_enqueue * 1e3, 1)
        parts["this_rank"] = round(t_local * 1e3, 1)
        say(f"matvec {k}: {times[-1] * 1e3:.1f} ms = {elements / times[-1]:.3e} matrix-elements/s; {mem()}")
        if world > 1:
            every = [None] * world
            dist.all_gather_object(every, parts)
        else:
            every = [parts]
        for r, p in enumerate(every):   # LS_B200_PROFILE: device spans per kernel; allgather = of the previous product
            say(f"    rank {r}: {p}")
    if times:
        result["matvec_ms"] = [v * 1e3 for v in times]
        result["matrix_elements_per_s"] = elements / min(times)
        # <x, H x> is the same number for every product form and every number of ranks
        result["x_H_x"] = float(sh.dot(x, y).item())
        result["x_x"] = float(sh.dot(x, x).item())
        say(f"<x|H|x> / <x|x> = {result['x_H_x'] / result['x_x']:.12f}")
    del x, y

    if args.rebalance and world > 1:
        from lattice_symmetries_b200.distributed import rebalance_distributed
        barrier()
        t0 = time.perf_counter()
        rebalance_distributed(basis)
        barrier()
        sh = _wrap(op)
        L = sh.layout
        say(f"rebalanced by measured kernel time in {time.perf_counter() - t0:.2f} s: bounds {L.bounds}; {mem()}")
        result["rebalanced_bounds"] = L.bounds
        x = hashed_vector(L.row_begin, L.row_end, 42)
        y = sh.empty_vector()
        for k in range(1):
            barrier()
            t0 = time.perf_counter()
            sh.matvec(x, y, mode)
            sh.sync()
            t_local = time.perf_counter() - t0
            barrier()
            dt = time.perf_counter() - t0
            say(f"matvec {k} after re-balancing: {dt * 1e3:.1f} ms = {elements / dt:.3e} matrix-elements/s (rank 0 alone {t_local * 1e3:.1f} ms)")
        result["matvec_ms_rebalanced"] = dt * 1e3
        result["matrix_elements_per_s_rebalanced"] = elements / dt
        result["x_H_x_rebalanced"] = float(sh.dot(x, y).item())
        say(f"<x|H|x> / <x|x> = {result['x_H_x_rebalanced'] / float(sh.dot(x, x).item()):.12f} (must not change)")
        del x, y

    if not args.no_lanczos:
        def progress(k, energy, resid):
            say(f"  iteration {k}: E = {energy:.10f}, residual {resid:.2e}, {time.perf_counter() - t1:.1f} s")
        sh.mode = mode
        t1 = time.perf_counter()
        res = lanczos_ground_state(sh, max_iters=args.max_iters, tol=args.tol, time_limit_s=args.time_limit, progress=progress,
                                   energy_tol=args.energy_tol, check_every=args.check_every, checkpoint=args.checkpoint,
                                   checkpoint_every=args.checkpoint_every, resume=args.resume)
        t_l = time.perf_counter() - t1
        n = model.number_sites
        say(f"Lanczos: {res.iterations} iterations in {t_l:.2f} s ({t_l / max(res.iterations, 1) * 1e3:.1f} ms per iteration)")
        say(f"E0 = {res.energy:.10f}  (converged={res.converged}, residual {res.residual:.2e});  "
            f"E0 / (4 N) = {res.energy / (4 * n):.8f} per site")
        result.update({"E0": res.energy, "E0_per_site_SS": res.energy / (4 * n), "iterations": res.iterations,
                       "converged": bool(res.converged), "residual": res.residual, "lanczos_s": t_l,
                       "ms_per_iteration": t_l / max(res.iterations, 1) * 1e3})
    if rank == 0 and args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(result, indent=1))
    if world > 1:
        dist.barrier()
        lib.ls_b200_comm_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
