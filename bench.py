#!/usr/bin/env python
"""bench.py -- the headline measurement (BASELINE.json ``metric``):
matvec matrix-elements/s (``value``) and basis-build states/s (``build``) on
N B200s of one node, next to the reference-equivalent CPU path on the host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload kagome36]
    python bench.py --impl reference ...        # CPU arm (oracle port of the reference path)
    torchrun --nproc-per-node N bench.py --gpus N ...

A *step* is one y = H x over the whole basis.  For N > 1 the representatives, x and y are sharded over the ranks
as contiguous row ranges (csrc/dist.cu, NCCL inside the library); a step includes the collective that the product
form needs (all-gather of the pre-scaled vector, or the all-to-all of (representative, coefficient) records).  The
basis build is timed outside the step loop (median of three full builds, every sample listed) and reported under
``build``.  Every line carries full-size parity checks (``checks``): sampled rows of y recomputed by the CPU oracle.

``value``  : x resident in HBM, device-timed with CUDA events on the library stream.
``e2e``    : the same step through the reference-facing call
             (``ls_chpl_matrix_vector_product`` behind the ``ls_chpl_kernels``
             vtable, chapel/src/DistributedMatrixVector.chpl:1090-1105) with
             pinned HOST buffers; H2D of x and D2H of y inside the timed region.
Inputs (x 8 B x dim, representatives 8 B x dim) are larger than the 126 MB L2 for
the default workload; for small workloads an L2 flush buffer is written between steps.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "matvec matrix-elements/s"
UNIT = "matrix-elements/s"


# ---- workloads (BASELINE.json configs) --------------------------------------------------
def make_model(name: str):
    from lattice_symmetries_b200 import lattices as L
    if name == "chain24":          # configs[0]
        return L.heisenberg_chain(24), "Heisenberg chain N=24, Sz=0, T x P x spin inversion (|G|=48 x 2)"
    if name == "kagome36":         # configs[1]
        return (L.kagome_heisenberg(36, spin_inversion=1),
                "36-site kagome Heisenberg, Sz=0, 12 translations x C6v (|G|=144) x spin inversion")
    if name == "kagome36_noinv":
        return L.kagome_heisenberg(36), "36-site kagome Heisenberg, Sz=0, 12 translations x C6v (|G|=144)"
    if name == "kagome30":
        return L.kagome_heisenberg(30, spin_inversion=1), "30-site kagome Heisenberg, Sz=0, translations x C2 x inversion"
    if name == "kagome27":
        return L.kagome_heisenberg(27), "27-site kagome Heisenberg"
    if name == "ladder_dm":        # configs[2]
        return L.ladder_dm(16), "2x16 spin ladder with DM terms, Sz=0, leg translation k=1 (complex characters)"
    if name == "hubbard4x4":       # configs[3]
        return L.hubbard_square(4, 4), "4x4 square-lattice Hubbard at half filling (8 up, 8 down), no projection"
    if name == "kagome42":         # configs[4]
        # (a singlet of 42 spins is ODD under spin inversion: (-1)^(N/2) with N/2 = 21)
        return (L.kagome_heisenberg(42, spin_inversion=-1),
                "42-site kagome Heisenberg, Sz=0, 14 translations x C2 (|G|=28) x spin inversion (odd sector)")
    if name.startswith("chain"):
        return L.heisenberg_chain(int(name[5:])), f"Heisenberg chain N={name[5:]}, Sz=0, T x P x spin inversion"
    raise SystemExit(f"unknown workload {name}")


# ---- clocks -------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device_index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU arm --------------------------------------------------------------------------------
def host_threads() -> int:
    """Cores this process may use (cgroup / affinity aware)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def load_oracle():
    """The CPU checker / baseline with ALL host cores: torchrun exports OMP_NUM_THREADS=1 to its workers, which
    would silently time the CPU arm on one core."""
    os.environ.pop("OMP_NUM_THREADS", None)
    from oracle import ls_oracle as oracle
    oracle.build()
    oracle.set_num_threads(host_threads())
    return oracle


def oracle_problem(model):
    """The workload described to the CPU oracle: (basis, off-diagonal terms, diagonal terms)."""
    oracle = load_oracle()
    from lattice_symmetries_b200.expr import compile_terms
    from lattice_symmetries_b200.symmetry import Symmetries
    if model.particle == "spin-1/2":
        syms = model.symmetries if model.symmetries is not None else Symmetries([])
        group = oracle.Group.from_symmetries(syms, model.number_sites, model.spin_inversion)
        ob = oracle.Basis(model.number_sites, 0, model.number_sites, model.hamming_weight, model.spin_inversion, group)
    else:
        up, down = model.number_particles
        ob = oracle.Basis(model.number_sites, 1, up + down, up, None, None)
    ts = compile_terms(model.expression, model.number_sites)
    off = oracle.Terms([t for t in ts if t.x != 0])
    diag = oracle.Terms([t for t in ts if t.x == 0])
    return oracle, ob, off, diag


def cpu_matvec_sample(model, reps, seconds_per_pass: float, passes: int = 1, warmup: int = 0, prefix: bool = False,
                      uniform: bool = False):
    """Times the reference's CPU path for one y = H x (push form: apply_off_diag -> state_info on betas and alphas ->
    state_index -> atomic add; OpenMP over all host cores) on a bounded sample of the columns.

    apply_off_diag and state_index are the reference's OWN compiled kernels/reference.c and kernels/indexing.c
    (oracle/_ref/libref.so) when that library is present; state_info is the C restatement (the reference's is
    Halide-generated and cannot be built here).  ``prefix`` = False: ``reps`` is the whole basis and the sample is
    UNIFORM (one block of 64 columns every ``stride`` columns); True: ``reps`` is only a sorted prefix of the basis
    (the stand-alone CPU arm cannot enumerate 9e9 candidates), all its columns are processed, and matrix elements
    that leave the prefix are searched for and dropped; with ``uniform`` as well, ``reps`` is a uniform thinning of the
    basis (cpu_build_sample with ``spread``) and its columns are sampled by stride like a whole basis.
    Returns (elements/s, description, cores, s per pass)."""
    oracle, ob, off, diag = oracle_problem(model)
    cores = oracle.num_threads()
    use_ref = oracle.ref_available()
    index = oracle.Index(reps, ob.number_bits, 22)
    dim = reps.shape[0]
    rng = np.random.default_rng(42)
    x = rng.standard_normal(dim)

    def run(stride, rows):
        t0 = time.perf_counter()
        _, n = oracle.matvec(ob, off, diag, index, x, 0, rows, sampling_prefix=prefix, block_stride=stride,
                             reference_kernels=use_ref)
        return time.perf_counter() - t0, n

    # calibrate on ~1024 columns per core, then size the sample for `seconds_per_pass`
    probe_cols = min(dim, 1024 * cores)
    if prefix and not uniform:
        dt, _ = run(64, probe_cols)
        rows, stride = int(min(dim, max(probe_cols, probe_cols * seconds_per_pass / max(dt, 1e-4)))), 64
    else:
        stride0 = max(64, dim // max(1, probe_cols // 64) // 64 * 64)
        dt, _ = run(stride0, dim)
        cols = min(dim, max(probe_cols, int(probe_cols * seconds_per_pass / max(dt, 1e-4))))
        rows, stride = dim, max(64, dim // max(1, cols // 64) // 64 * 64)
    times, nnz = [], 0
    for it in range(warmup + passes):
        dt, nnz = run(stride, rows)
        if it >= warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    columns = rows if stride == 64 else ((rows + stride - 1) // stride) * 64
    elements = nnz + columns  # off-diagonal elements + the diagonal
    if prefix and uniform:
        where = ("%d columns (64 every %d) of a CPU-built uniform thinning of the basis (%d representatives from blocks "
                 "spaced evenly over the candidate range)" % (columns, stride, dim))
    elif prefix:
        where = "columns [0,%d) of a CPU-built sorted prefix of the basis" % rows
    else:
        where = "%d columns sampled uniformly (64 every %d) from the basis (dim %d)" % (columns, stride, dim)
    kernels = ("apply_off_diag + state_index = the reference's compiled kernels/reference.c + indexing.c, state_info = C port"
               if use_ref else "all kernels = C port (oracle/_ref/libref.so absent)")
    return (elements / dt, f"{where}: {nnz} off-diagonal elements per pass, {dt:.2f} s per pass, {len(times)} timed "
            f"pass(es); {kernels}", cores, dt)


def cpu_build_sample(model, seconds_target: float = 8.0, want_reps: int = 0, spread: int = 1):
    """Times the oracle's enumeration (Gosper stepping + is_representative, OpenMP over chunks like
    StatesEnumeration.chpl:392-458) on n candidates; returns (stats, representatives found, sorted).

    ``spread`` = 1: the first n candidates (a sorted PREFIX of the basis -- what the in-line baseline compares with
    the head of the GPU-built basis).  ``spread`` > 1: n candidates in that many equal blocks spaced evenly over the
    whole candidate range, i.e. a uniform thinning of the basis: representatives crowd into the low indices (the
    first 3 % of kagome-36's range hold 90 % of them) and cost differently to find there, so a prefix alone says
    little about the whole scan -- or about the columns of the whole matrix."""
    oracle, ob, _, _ = oracle_problem(model)
    if model.particle != "spin-1/2" or model.hamming_weight is None:
        reps = ob.enumerate()
        return None, reps
    lo, hi = ob.min_state(), ob.max_state()
    lib = oracle.lib()
    r_lo = int(lib.oracle_fixed_hamming_state_to_index(lo))
    r_hi = int(lib.oracle_fixed_hamming_state_to_index(hi))
    total = r_hi - r_lo + 1
    n = min(total, 1 << 21)
    hw = model.hamming_weight
    state_at = lambda k: int(lib.oracle_fixed_hamming_index_to_state(r_lo + k, hw))
    while True:
        blocks = max(1, min(spread, n >> 16)) if n < total else 1
        size = n // blocks
        starts = [0] if blocks == 1 else [(total - size) * k // (blocks - 1) for k in range(blocks)]
        t0 = time.perf_counter()
        parts = [ob.enumerate_range(state_at(a), state_at(a + size - 1)) for a in starts]
        dt = time.perf_counter() - t0
        reps = np.concatenate(parts) if len(parts) > 1 else parts[0]
        scanned = blocks * size
        if n == total or (dt > seconds_target / 3 and reps.shape[0] >= want_reps):
            break
        n = int(min(total, max(2 * n, n * seconds_target / 1.5 / max(dt, 1e-3))))
    where = f"first {scanned}" if blocks == 1 else f"{blocks} blocks of {size} spaced evenly over the range, {scanned}"
    stats = {"candidates_per_s": scanned / dt, "representatives_per_s": reps.shape[0] / dt, "cores": oracle.num_threads(),
             "sample": f"{where} of {total} candidates, {dt:.2f} s"}
    return stats, reps


# ---- full-size parity: sampled rows recomputed by the oracle ------------------------------------------------------
def sampled_rows_check(model, basis, lay, y_local, seed_x: int, complex_vectors: bool, world: int, samples: int = 4096):
    """Rows of y = H x at the benchmark size, recomputed on the CPU by the oracle and compared with the GPU result.

    Row i of a Hermitian H is the conjugate of column i, and column i is what the reference's kernels produce:
    apply_off_diag on |alpha_i>, state_info on every beta (representative, character, norm), index of the
    representative.  So  y_i = d_i x_i + sum_k conj(chi_k c_k) (n_k / n_i) x_{j_k}.  The entries of x are a hash of
    the GLOBAL row (distributed.hashed_values), so any rank can evaluate x_j; representatives are ranked with each
    rank's own index and summed over ranks.  Every rank calls this (collectives inside when world > 1); returns
    {samples, max_rel_err, missing} (max |y_i - want_i| relative to the RMS of the sampled rows)."""
    import torch
    from lattice_symmetries_b200.distributed import hashed_values
    oracle, ob, off, diag = oracle_problem(model)
    dim = lay.dim
    n = int(min(samples, dim))
    rows = np.unique(np.random.default_rng(1234).integers(0, dim, size=n)) if dim > n else np.arange(dim)
    n = rows.shape[0]
    lo, hi = lay.row_begin, lay.row_end
    mine = (rows >= lo) & (rows < hi)
    states = np.asarray(basis.states) if hi > lo else np.zeros(0, np.uint64)
    alphas = np.zeros(n, dtype=np.uint64)
    alphas[mine] = states[rows[mine] - lo]
    width = 2 if complex_vectors else 1
    got = np.zeros((n, width))
    if mine.any():
        sel = torch.as_tensor(rows[mine] - lo, device=y_local.device)
        picked = y_local[sel]
        got[mine] = (torch.view_as_real(picked) if complex_vectors else picked.reshape(-1, 1)).cpu().numpy()

    def allreduce(a, dtype):
        if world == 1:
            return a
        import torch.distributed as dist
        t = torch.as_tensor(a.view(np.int64) if a.dtype == np.uint64 else a, device="cuda").to(dtype)
        dist.all_reduce(t)
        out = t.cpu().numpy()
        return out.view(np.uint64) if a.dtype == np.uint64 else out

    alphas = allreduce(alphas, torch.int64)
    got = allreduce(got, torch.float64)
    betas, coeffs, offsets = oracle.apply_off_diag(off, alphas)
    col = np.repeat(np.arange(n), np.diff(offsets))
    if ob.group is not None and ob.c.has_permutation_symmetries:
        rep_b, chi, n_b = ob.group.state_info(betas)
        n_a = ob.group.state_info(alphas)[2]
    elif model.spin_inversion:
        mask = np.uint64((1 << model.number_sites) - 1)
        flipped = betas ^ mask
        chi = np.where(flipped < betas, float(model.spin_inversion), 1.0).astype(np.complex128)
        rep_b = np.minimum(betas, flipped)
        n_b, n_a = np.ones(betas.shape[0]), np.ones(n)
    else:
        rep_b, chi, n_b, n_a = betas, np.ones(betas.shape[0], np.complex128), np.ones(betas.shape[0]), np.ones(n)
    live = n_b > 0
    # global row of every representative: each rank looks up its own block
    j_local = basis.index(rep_b) if hi > lo else np.full(rep_b.shape[0], -1, dtype=np.int64)
    found = j_local >= 0
    j_glob = allreduce(np.where(found, j_local + lo + 1, 0).astype(np.int64), torch.int64) - 1
    missing = int(np.count_nonzero(live & (j_glob < 0) & (np.abs(coeffs) > 0)))
    ok = live & (j_glob >= 0)
    jt = torch.as_tensor(np.where(ok, j_glob, 0))
    xj = hashed_values(jt, seed_x).numpy().astype(np.complex128)
    xi = hashed_values(torch.as_tensor(rows), seed_x).numpy().astype(np.complex128)
    if complex_vectors:
        xj = xj + 1j * hashed_values(jt, seed_x + 1000003).numpy()
        xi = xi + 1j * hashed_values(torch.as_tensor(rows), seed_x + 1000003).numpy()
    h = np.where(ok, np.conj(chi * coeffs) * n_b / n_a[col], 0.0)
    want = np.zeros(n, dtype=np.complex128)
    np.add.at(want, col, h * xj)
    want += oracle.apply_diag(diag, alphas) * xi if diag.n else 0.0
    got_c = got[:, 0] + (1j * got[:, 1] if complex_vectors else 0.0)
    if not complex_vectors:
        want = want.real
    rms = float(np.sqrt(np.mean(np.abs(want) ** 2))) or 1.0
    return {"samples": int(n), "max_rel_err": float(np.max(np.abs(got_c - want)) / rms), "missing": missing,
            "definition": "max_i |y_i - oracle_i| / rms(oracle rows); row i = conj(column i) from the oracle's "
                          "apply_off_diag + state_info, indices from the GPU basis"}


# ---- main -------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("LS_BENCH_WORKLOAD", "kagome36"))
    ap.add_argument("--mode", default="auto", choices=["auto", "allgather", "alltoall"],
                    help="distributed product form (N > 1)")
    ap.add_argument("--dist-flags", type=int, default=0,
                    help="ls_b200_dist_build flags (N > 1): 1 no replicated index, 2 wide replicated index, 4 even row split")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-checks", action="store_true")
    return ap.parse_args()


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(args):
    """CPU arm: the reference's path for this metric on the host cores.  The reference's orbit kernels are
    Halide-generated and its driver is Chapel -- neither toolchain exists here -- so the loop and state_info are the
    oracle port (oracle/ls_oracle.c, a line-by-line restatement), with the reference's own compiled apply_off_diag
    and state_index (oracle/_ref/libref.so) inside it.  Self-contained on the CPU: it scans blocks of candidates spaced
    evenly over the whole range (a uniform thinning of the basis -- a prefix would hold only the cheap low-index
    columns), then times matvec passes over columns sampled evenly from those.  Under torchrun rank 0 alone runs, on
    ALL host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model, desc = make_model(args.workload)
    passes = max(1, args.steps)
    warm = max(0, min(args.warmup, 1))
    build_stats, reps = cpu_build_sample(model, 10.0, want_reps=20000, spread=16)
    per_pass = min(10.0, 60.0 / (passes + warm))
    value, sample, cores, dt = cpu_matvec_sample(model, reps, per_pass, passes, warm, prefix=True, uniform=True)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "name": args.workload},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "build": build_stats},
        "build": build_stats,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def ncu_capture(kernel: str, workload: str):
    """Numbers of the latest committed ``ncu --set full`` capture of ``kernel`` on ``workload``
    (profiles/rNN/KERNEL_WORKLOAD.summary.txt, written by tools/ncu_summary.py): per-launch DRAM bytes and executed
    alu-pipe warp instructions, plus the launch's matrix elements when the summary records them."""
    profs = sorted((ROOT / "profiles").glob(f"r*/{kernel}_{workload}.summary.txt"))
    if not profs:
        return None
    out = {"file": str(profs[-1].relative_to(ROOT))}
    for l in profs[-1].read_text().splitlines():
        parts = l.split()
        if len(parts) < 2:
            continue
        if l.startswith("Elapsed Cycles"):
            try:
                out["elapsed_cycles"] = float(parts[-1].replace(",", ""))
            except ValueError:
                pass
            continue
        key, val = parts[0], parts[1].replace(",", "")
        try:
            v = float(val)
        except ValueError:
            continue
        if key in ("dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed_pipe_alu.sum",
                   "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
                   "launch_matrix_elements", "gpu__time_duration.sum"):
            out[key] = v
    # alu-pipe warp instructions of the captured launch: counted directly when the capture has the counter, else from
    # the pipe's utilisation: pct x elapsed cycles x 2 warp-instructions per cycle per SM (4 sub-partitions x 16 lanes)
    pct = out.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")
    if "smsp__inst_executed_pipe_alu.sum" not in out and pct is not None and "elapsed_cycles" in out:
        out["smsp__inst_executed_pipe_alu.sum"] = pct / 100.0 * out["elapsed_cycles"] * 2.0 * 148
        out["alu_inst_source"] = "pct_of_peak x elapsed cycles x 2 warp-inst/clk/SM x 148 SMs"
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    os.environ.setdefault("LS_B200_PROFILE", "1")  # library-side CUDA events around every orbit / rank launch
    import torch
    import torch.distributed as dist
    import lattice_symmetries_b200 as ls  # noqa: F401
    from lattice_symmetries_b200 import _lib
    from lattice_symmetries_b200.distributed import (ALLGATHER, ALLTOALL, AUTO, build_distributed, hashed_vector,
                                                     init_communicator, init_process)
    from lattice_symmetries_b200.lanczos import _wrap

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    init_process(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        init_communicator()   # the library's own NCCL communicator; torch.distributed only carries its 128-byte id
    lib = _lib.lib
    hbm_peak, peak_src = load_peaks()
    mode = {"auto": AUTO, "allgather": ALLGATHER, "alltoall": ALLTOALL}[args.mode]

    model, desc = make_model(args.workload)
    basis = model.basis()

    def sync():
        torch.cuda.current_stream().synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(values):
        if world == 1:
            return list(values)
        t = torch.tensor(list(values), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    # ---- basis build ------------------------------------------------------------------------------------------
    total_candidates = basis.number_candidates
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # warm-up build of a small shard (module load, constant upload), then the timed full builds
    wb = model.basis()
    r, n, c = wb.build_shard(0, min(total_candidates, 1 << 22))
    lib.ls_b200_device_free(r)
    if n:
        lib.ls_b200_device_free(n)
    del wb
    sync()
    # The whole build (enumeration, redistribution to contiguous row ranges when N > 1, host view, state -> index
    # structure) from scratch on a fresh basis object; median of three (one for the largest workloads); all samples listed.
    # (N > 1: five samples -- the first build of a process also sets up the NCCL connections, and an allocation that
    # reaches the driver under peer access can take a second; the median of five shrugs off two such outliers)
    number_builds = 1 if total_candidates > (1 << 36) else (5 if world > 1 else 3)
    samples = []
    op = None
    for attempt in range(number_builds):
        if attempt > 0:
            del op, basis
            basis = model.basis()
        op = model.operator(basis)
        sync()
        launches_b0 = lib.ls_b200_kernel_launch_count()
        t0 = time.perf_counter()
        ev0.record()
        if world > 1:
            build_distributed(basis, balance_for=op, flags=args.dist_flags)   # rows split so that every rank holds the same number of elements
        else:
            basis.build()
        ev1.record()
        sync()
        wall = time.perf_counter() - t0
        ms, kernel_ms, wall = max_over_ranks([ev0.elapsed_time(ev1), lib.ls_b200_last_kernel_ms(b"build"), wall])
        samples.append((ms, kernel_ms, wall))
    build_samples_ms = [m for m, _, _ in samples]
    build_ms, build_kernel_ms, build_wall = sorted(samples)[len(samples) // 2]
    build_launches = lib.ls_b200_kernel_launch_count() - launches_b0

    sh = _wrap(op)   # DistributedOperator on a sharded basis, the plain device product on one GPU
    L = sh.layout
    dim = L.dim
    rows_local = L.rows
    nnz_local = op.count_matrix_elements(0, rows_local)
    if world > 1:
        t = torch.tensor([nnz_local], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        nnz_total = int(t.item())
    else:
        nnz_total = nnz_local
    elements_total = nnz_total + dim

    # ---- device-resident steps ------------------------------------------------------------------
    complex_vectors = model.symmetries is not None and not bool(
        np.all(np.abs(model.symmetries.characters()[1]) < 1e-9))
    vdtype = torch.complex128 if complex_vectors else torch.float64
    SEED_X = 42
    x = hashed_vector(L.row_begin, L.row_end, SEED_X)   # entry i depends on (seed, global row i) only
    if complex_vectors:
        x = torch.complex(x, hashed_vector(L.row_begin, L.row_end, SEED_X + 1000003))
    y = sh.empty_vector(vdtype)
    small = (dim * 16) < (256 << 20)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None

    def step():
        if flush is not None:
            flush.fill_(1)
        sh.matvec(x, y, mode)

    for _ in range(args.warmup):
        step()
    sh.sync()
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.ls_b200_kernel_launch_count()
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in events:
        if flush is not None:
            flush.fill_(1)
        a.record()
        sh.matvec(x, y, mode)
        b.record()
    sync()
    launches = lib.ls_b200_kernel_launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in events]
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel device time inside one step, CUDA events on the launching stream (library-side, LS_B200_PROFILE)
    kernel_ms, per_kernel = [], []
    for _ in range(min(3, args.steps)):
        sh.matvec(x, y, mode)
        sh.sync()
        kernel_ms.append(lib.ls_b200_last_kernel_ms(b"matvec"))
        per_kernel.append({k: lib.ls_b200_last_kernel_ms(k.encode()) for k in
                           ("orbit", "orbit_launches", "gather", "gather_launches", "combine")})
    (total_ms,) = max_over_ranks([sum(step_ms)])
    ms_per_step = total_ms / args.steps
    value = elements_total / (ms_per_step * 1e-3)

    # ---- full-size checks, at every N ------------------------------------------------------------------------------
    checks = {}
    if not args.no_checks:
        sh.matvec(x, y, mode)
        sh.sync()
        try:
            checks["sampled_rows"] = sampled_rows_check(model, basis, L, y, SEED_X, complex_vectors, world)
            checks["sampled_rows_max_rel_err"] = checks["sampled_rows"]["max_rel_err"]
        except Exception as e:  # never lose the line over the checker
            checks["sampled_rows"] = {"failed": repr(e)}
        states = np.asarray(basis.states[:min(rows_local, 1 << 22)]) if rows_local else np.zeros(0, np.uint64)
        ok_sorted = bool(np.all(states[1:] > states[:-1])) if states.size > 1 else True
        (bad,) = max_over_ranks([0.0 if ok_sorted else 1.0])
        checks["representatives_sorted_prefix"] = bad == 0.0
        # Hermiticity of the projected operator: <u, H x> == <H u, x>
        u = hashed_vector(L.row_begin, L.row_end, 7)
        if complex_vectors:
            u = torch.complex(u, hashed_vector(L.row_begin, L.row_end, 8))
        hu = sh.empty_vector(vdtype)
        sh.matvec(u, hu, mode)
        sh.sync()
        a = sh.dot(u, y)
        b = sh.dot(hu, x)
        scale = torch.sqrt(sh.dot(u, u).real * sh.dot(y, y).real)
        checks["hermiticity_rel_err"] = float((torch.abs(a - b) / scale).item())
        del u, hu

    # ---- e2e: HOST buffers through the reference-facing vtable call, at every N -------------------------------------
    # ls_chpl_kernels.matrix_vector_product (chapel/src/DistributedMatrixVector.chpl:1090-1105): every rank passes its
    # own blocks of x and y, like the per-locale blocks of the reference (:1060-1088); H2D, the collectives and D2H are
    # all inside the call and inside the timed region.
    e2e = None
    if not args.no_e2e and not complex_vectors:
        nbytes = 8 * rows_local
        hx = lib.ls_b200_host_malloc(max(nbytes, 8))
        hy = lib.ls_b200_host_malloc(max(nbytes, 8))
        if rows_local:
            np.frombuffer((C.c_double * rows_local).from_address(hx), dtype=np.float64)[:] = x.cpu().numpy()
        kernels = lib.ls_hs_internal_get_chpl_kernels()
        mv = kernels.contents.matrix_vector_product
        for _ in range(max(1, args.warmup // 2)):
            mv(C.byref(op._payload), 1, C.cast(hx, _lib.f64_p), C.cast(hy, _lib.f64_p))
        _lib.check_error()
        sync()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            mv(C.byref(op._payload), 1, C.cast(hx, _lib.f64_p), C.cast(hy, _lib.f64_p))
        e2e_s = (time.perf_counter() - t0) / args.steps
        _lib.check_error()
        (e2e_s,) = max_over_ranks([e2e_s])
        if rows_local and not args.no_checks:
            yh = np.frombuffer((C.c_double * rows_local).from_address(hy), dtype=np.float64)
            ref = y.cpu().numpy()
            err = float(np.linalg.norm(yh - ref) / max(np.linalg.norm(ref), 1e-300))
            (err,) = max_over_ranks([err])
            checks["e2e_equals_device_resident_rel_err"] = err
        lib.ls_b200_host_free(hx)
        lib.ls_b200_host_free(hy)
        e2e = {"value": elements_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * dim,
               "d2h_bytes_per_step": 8 * dim, "ms_per_step": e2e_s * 1e3,
               "call": "ls_chpl_kernels.matrix_vector_product (pinned host pointers; every rank passes its row block)"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            lib.ls_b200_comm_finalize()
            dist.destroy_process_group()
        return

    # ---- roofline -------------------------------------------------------------------------------------------
    # One y = H x is three kernels per row chunk: orbit_kernel (canonicalise every matrix element: integer-issue
    # bound), rank_gather_kernel (state -> index, gather n_j x_j: latency / random-access bound), row_sum_kernel.
    # For every kernel: its in-situ device time per step (CUDA events on the launching stream), its share of the step,
    # its algorithmic HBM bytes against the measured copy bandwidth, and -- the roofline that binds this path -- the
    # alu-pipe thread-instructions it executes (from the committed ncu capture, scaled by matrix elements) against the
    # LOP3 peak measured live by csrc/peaks.cu.
    vec_bytes = 16 if complex_vectors else 8
    pk = per_kernel[len(per_kernel) // 2] if per_kernel else {}
    orbit_launches = int(pk.get("orbit_launches", 0) or 0)
    group_size = len(model.symmetries.elements) if model.symmetries is not None else 0
    images = group_size * (2 if model.spin_inversion else 1)
    nbits = model.number_sites * (2 if model.particle != "spin-1/2" else 1)
    depth = 2 * max(1, (nbits - 1).bit_length()) - 1
    # W_m = I (6 depth + 4) + 3 ceil(log2 range) + 8 u64-ops per matrix element (un-pruned reference count)
    w_m = images * (6 * depth + 4) + 3 * 5 + 8
    whole_ms = statistics.median(kernel_ms) if kernel_ms else ms_per_step
    lop3_peak = float(lib.ls_b200_measure_lop3_peak())  # measured LOP3 thread-instructions/s (alu pipe)
    per_el = 8 + 2 + (1 if complex_vectors else 0)      # orbit writes: representative (8) + term/sign (2) [+ character index]
    algo = {  # algorithmic bytes per STEP on this rank
        "orbit_kernel": nnz_local * per_el + rows_local * 12,
        "rank_gather_kernel": nnz_local * (per_el + vec_bytes + vec_bytes),   # reads what orbit wrote, gathers x_j, writes the value
        "row_sum_kernel": nnz_local * vec_bytes + rows_local * (16 + 2 * vec_bytes),
    }
    times = {"orbit_kernel": pk.get("orbit"), "rank_gather_kernel": pk.get("gather"), "row_sum_kernel": pk.get("combine")}
    kernels_block = {}
    for name, t_ms in times.items():
        if not t_ms:
            continue
        cap = ncu_capture(name, args.workload)
        entry = {"ms_per_step": t_ms, "share_of_step": t_ms / whole_ms,
                 "algorithmic_bytes_per_step": algo[name], "hbm_GBps": algo[name] / (t_ms * 1e-3) / 1e9,
                 "hbm_frac": algo[name] / (t_ms * 1e-3) / 1e9 / hbm_peak}
        if cap is not None and "smsp__inst_executed_pipe_alu.sum" in cap:
            # warp instructions -> thread instructions (x 32).  With the captured launch's element count on record:
            # per matrix element x this rank's elements; otherwise per launch x launches (chunks are cut to hold the
            # same number of elements, so the captured launch is representative).
            if cap.get("launch_matrix_elements"):
                alu = cap["smsp__inst_executed_pipe_alu.sum"] * 32.0 / cap["launch_matrix_elements"] * nnz_local
            else:
                alu = cap["smsp__inst_executed_pipe_alu.sum"] * 32.0 * max(1, orbit_launches)
            entry.update({
                "alu_thread_inst_per_step": alu, "alu_Tinst_per_s": alu / (t_ms * 1e-3) / 1e12,
                "int_frac": alu / (t_ms * 1e-3) / lop3_peak,
                "alu_pipe_pct_of_peak_under_ncu": cap.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                "dram_bytes_per_launch_ncu": (cap.get("dram__bytes_read.sum", 0) + cap.get("dram__bytes_write.sum", 0)) * 1e6
                if "dram__bytes_read.sum" in cap else None,
                "ncu_capture": cap["file"]})
        kernels_block[name] = entry
    if orbit_launches > 0 and "orbit_kernel" in kernels_block:
        dom = kernels_block["orbit_kernel"]
        k_ms = pk["orbit"] / orbit_launches
        if "int_frac" in dom:
            roofline = {"bound": "int", "achieved": dom["alu_Tinst_per_s"], "peak": lop3_peak / 1e12,
                        "unit": "T alu-pipe thread-instructions/s", "frac": dom["int_frac"],
                        "traffic": dom.get("dram_bytes_per_launch_ncu"),
                        "peak_source": "LOP3 issue rate measured live (csrc/peaks.cu, ls_b200_measure_lop3_peak)"}
        else:
            roofline = {"bound": "hbm", "achieved": dom["hbm_GBps"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": dom["hbm_frac"], "traffic": None, "peak_source": peak_src}
        roofline.update({"kernel": "orbit_kernel", "kernel_ms": k_ms, "launches_per_step": orbit_launches,
                         "algorithmic_bytes_per_launch": algo["orbit_kernel"] / orbit_launches})
    else:
        # no symmetry group: one gather kernel per step (x[j] gather per element; alpha, norm, x, y per row)
        algo_bytes = nnz_local * vec_bytes + rows_local * (16 + 2 * vec_bytes)
        achieved = algo_bytes / (whole_ms * 1e-3) / 1e9
        cap = ncu_capture("gather_kernel", args.workload)
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": ((cap.get("dram__bytes_read.sum", 0) + cap.get("dram__bytes_write.sum", 0)) * 1e6
                                if cap and "dram__bytes_read.sum" in cap else None),
                    "peak_source": peak_src, "kernel": "gather_kernel", "kernel_ms": whole_ms, "launches_per_step": 1,
                    "algorithmic_bytes_per_launch": algo_bytes}
    roofline.update({
        "hbm_peak_GBps": hbm_peak, "hbm_peak_source": peak_src, "lop3_peak_Tinst_per_s": lop3_peak / 1e12,
        "whole_step_kernels_ms": whole_ms, "kernels": kernels_block,
        "reference_op_count": {"u64_ops_per_element": w_m,
                               "Tops_per_s_if_unpruned": (nnz_local + rows_local) * w_m / (whole_ms * 1e-3) / 1e12,
                               "note": "SURVEY 8(d) un-pruned count of the reference algorithm; the bit-sliced kernels "
                                       "execute far fewer instructions (32 states per LOP3), so this is context, not a fraction"},
        "note": "int_frac = executed alu-pipe thread-instructions (committed ncu capture, per matrix element, x this "
                "step's elements) / in-situ kernel time / measured LOP3 peak; hbm_frac = algorithmic bytes / time / "
                "measured copy bandwidth"})

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1 and not complex_vectors:
        try:
            v, sample, cores, _ = cpu_matvec_sample(model, np.asarray(basis.states), 12.0)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
            cb, cpu_reps = cpu_build_sample(model, 8.0)
            if cb is not None:
                cpu_baseline["build"] = cb
                # full-size parity check for free: the CPU port enumerated a prefix of the candidate range
                # (a complete basis when the range is small) -- it must be the head of the GPU's sorted list
                head = np.asarray(basis.states[:cpu_reps.shape[0]])
                checks["build_prefix_equals_cpu_port"] = bool(
                    cpu_reps.shape[0] <= dim and np.array_equal(head, cpu_reps))
                checks["build_prefix_states"] = int(cpu_reps.shape[0])
        except Exception as e:  # the baseline is informational; never lose the GPU line over it
            cpu_baseline = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": f"failed: {e}"}

    form = {0: "auto", 1: "all-gather", 2: "all-to-all"}[mode]
    if world > 1 and mode == AUTO:
        form = "all-gather" if L.global_index else "all-to-all"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "c128" if complex_vectors else "f64", "data": "synthetic",
        "config": {"workload": desc, "name": args.workload, "dim": dim, "candidates": total_candidates,
                   "off_diag_elements": nnz_total, "l2": "flushed between steps" if small else "inputs larger than L2",
                   "parallelism": (f"representatives, x and y sharded over {world} ranks as contiguous row ranges balanced "
                                   f"by matrix-element count; {form} product; NCCL inside the library"
                                   if world > 1 else "single GPU")},
        "build": {"candidates_per_s": total_candidates / (build_ms * 1e-3), "representatives_per_s": dim / (build_ms * 1e-3),
                  "ms": build_ms, "samples_ms": build_samples_ms, "kernel_ms_slowest_rank": build_kernel_ms,
                  "wall_ms": build_wall * 1e3,
                  "gpu_launches": int(build_launches), "unit": "states/s"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "checks": checks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        lib.ls_b200_comm_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
