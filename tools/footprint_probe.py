"""Single-GPU probe of the large-footprint regime (lookup structure + vector >> L2, as on a kagome-42 rank):
Heisenberg chain of N sites, Sz = 0, translations only (x spin inversion) -- N = 40: dim 1.72e9, 3.4e10 matrix elements.

    python tools/footprint_probe.py 40 [--wide] [--matvecs 3]

--wide drives the product through the distributed code path with one virtual rank and the "wide" replicated index
(64-bit bucket starts + compact keys: what a kagome-42 rank searches).  LS_B200_MV_SORT=0/1 switches the sorted
ranking; <x|H|x> / <x|x> is printed for cross-checking the variants (the vector is a hash of the row number)."""
import argparse
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sites", type=int, nargs="?", default=40)
    ap.add_argument("--wide", action="store_true")
    ap.add_argument("--matvecs", type=int, default=3)
    ap.add_argument("--ballast", type=float, default=0.0, help="GB of device memory to occupy before the products")
    args = ap.parse_args()
    os.environ.setdefault("LS_B200_PROFILE", "1")
    import torch
    import ctypes as C
    from lattice_symmetries_b200 import _lib, lattices
    from lattice_symmetries_b200.distributed import EmulatedRanks, WIDE_INDEX, NO_BALANCE, hashed_vector, init_process
    init_process(0)
    lib = _lib.lib
    model = lattices.heisenberg_chain(args.sites, parity_sector=None, spin_inversion=1 if (args.sites // 2) % 2 == 0 else -1)
    t0 = time.perf_counter()
    if args.wide:
        team = EmulatedRanks(model.basis, model.operator, 1, WIDE_INDEX | NO_BALANCE)
        basis, op = team.bases[0], team.ops[0]
    else:
        basis = model.basis()
        basis.build()
        op = model.operator(basis)
    torch.cuda.synchronize()
    dim = basis.number_states
    print(f"chain {args.sites}: dim {dim}, build {time.perf_counter() - t0:.2f} s, wide={args.wide}, "
          f"sort={os.environ.get('LS_B200_MV_SORT', 'auto')}", flush=True)
    ballast = torch.empty(int(args.ballast * 2**30), dtype=torch.uint8, device="cuda") if args.ballast > 0 else None
    nnz = op.count_matrix_elements(0, dim)
    x = hashed_vector(0, dim, 42)
    y = torch.zeros_like(x)
    for k in range(args.matvecs):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if args.wide:
            xp = (C.c_void_p * 1)(x.data_ptr())
            yp = (C.c_void_p * 1)(y.data_ptr())
            lib.ls_b200_emu_matvec(team._ops_array, 1, xp, yp, 1, 0)
        else:
            op.matvec_device(x.data_ptr(), y.data_ptr())
        lib.ls_b200_matvec_sync()
        _lib.check_error()
        dt = time.perf_counter() - t0
        parts = {k2: round(lib.ls_b200_last_kernel_ms(k2.encode()), 1) for k2 in ("orbit", "gather", "combine", "matvec")}
        print(f"matvec {k}: {dt * 1e3:.1f} ms = {(nnz + dim) / dt:.3e} elements/s; kernels ms {parts}", flush=True)
    e = float(torch.dot(x, y).item()) / float(torch.dot(x, x).item())
    free, total = torch.cuda.mem_get_info()
    print(f"<x|H|x>/<x|x> = {e:.13f}; {nnz} off-diagonal elements; {(total - free) / 2**30:.1f} GiB used")


if __name__ == "__main__":
    main()
