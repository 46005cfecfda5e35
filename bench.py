#!/usr/bin/env python
"""bench.py -- the headline measurement (BASELINE.json ``metric``):
matvec matrix-elements/s (``value``) and basis-build states/s (``build``) on
N B200s of one node, next to the reference-equivalent CPU path on the host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload kagome36]
    python bench.py --impl reference ...        # CPU arm (oracle port of the reference path)
    torchrun --nproc-per-node N bench.py --gpus N ...

A *step* is one y = H x over the whole basis (every rank: its contiguous row
shard, preceded for N > 1 by the NCCL all-gather that replicates x).  The basis
build is timed outside the step loop (median of three full builds, every sample listed) and reported under ``build``.

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
        return (L.kagome_heisenberg(42, spin_inversion=1),
                "42-site kagome Heisenberg, Sz=0, 14 translations x C2 (|G|=28) x spin inversion")
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
def oracle_problem(model):
    """The workload described to the CPU oracle: (basis, off-diagonal terms, diagonal terms)."""
    from oracle import ls_oracle as oracle
    from lattice_symmetries_b200.expr import compile_terms
    from lattice_symmetries_b200.symmetry import Symmetries
    oracle.build()
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


def cpu_matvec_sample(model, reps, seconds_per_pass: float, passes: int = 1, warmup: int = 0, prefix: bool = False):
    """Times the oracle's push-form matvec (the CPU restatement of the reference
    path: apply_off_diag -> state_info on betas and alphas -> state_index ->
    atomic add; OpenMP over all host cores) on the columns [0, R) of the
    workload.  ``prefix``: ``reps`` is only a sorted prefix of the basis (see
    oracle_matvec's sampling mode).  Returns (elements/s, description, cores, s per pass)."""
    oracle, ob, off, diag = oracle_problem(model)
    cores = oracle.num_threads()
    index = oracle.Index(reps, ob.number_bits, 22)
    dim = reps.shape[0]
    rng = np.random.default_rng(42)
    x = rng.standard_normal(dim)
    rows = min(dim, 1024 * cores)
    t0 = time.perf_counter()
    oracle.matvec(ob, off, diag, index, x, 0, rows, sampling_prefix=prefix)
    dt = max(time.perf_counter() - t0, 1e-4)
    rows = int(min(dim, max(rows, rows * seconds_per_pass / dt)))
    times, nnz = [], 0
    for it in range(warmup + passes):
        t0 = time.perf_counter()
        _, nnz = oracle.matvec(ob, off, diag, index, x, 0, rows, sampling_prefix=prefix)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    elements = nnz + rows  # off-diagonal elements + the diagonal
    what = "a CPU-built sorted prefix of the basis" if prefix else f"the basis (dim {dim})"
    return (elements / dt, f"columns [0,{rows}) of {what}: {nnz} off-diagonal elements per pass, {dt:.2f} s per pass, "
            f"{len(times)} timed pass(es)", cores, dt)


def cpu_build_sample(model, seconds_target: float = 8.0, want_reps: int = 0):
    """Times the oracle's enumeration (Gosper stepping + is_representative,
    OpenMP over chunks like StatesEnumeration.chpl:392-458) on the first n
    candidates; returns (stats, representatives found)."""
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
    while True:
        upper = int(lib.oracle_fixed_hamming_index_to_state(r_lo + n - 1, hw))
        t0 = time.perf_counter()
        reps = ob.enumerate_range(lo, upper)
        dt = time.perf_counter() - t0
        if n == total or (dt > seconds_target / 3 and reps.shape[0] >= want_reps):
            break
        n = int(min(total, max(2 * n, n * seconds_target / 1.5 / max(dt, 1e-3))))
    stats = {"candidates_per_s": n / dt, "representatives_per_s": reps.shape[0] / dt, "cores": oracle.num_threads(),
             "sample": f"first {n} of {total} candidates, {dt:.2f} s"}
    return stats, reps


# ---- main -------------------------------------------------------------------------------------
def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("LS_BENCH_WORKLOAD", "kagome36"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(args):
    """CPU arm: the reference's path for this metric on the host cores.  The
    reference's orbit kernels are Halide-generated and its driver is Chapel --
    neither toolchain exists here -- so this is the oracle port (kind "port"),
    which restates them line by line (oracle/ls_oracle.c), self-contained on
    the CPU: it builds a prefix of the basis itself, then times matvec passes
    over those columns."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model, desc = make_model(args.workload)
    passes = max(1, args.steps)
    warm = max(0, min(args.warmup, 1))
    build_stats, reps = cpu_build_sample(model, 10.0, want_reps=20000)
    per_pass = min(10.0, 60.0 / (passes + warm))
    value, sample, cores, dt = cpu_matvec_sample(model, reps, per_pass, passes, warm, prefix=True)
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


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    os.environ.setdefault("LS_B200_PROFILE", "1")  # library-side CUDA events around every orbit / rank launch
    import torch
    import torch.distributed as dist
    import lattice_symmetries_b200 as ls
    from lattice_symmetries_b200 import _lib
    from lattice_symmetries_b200.distributed import ShardedOperator, build_sharded, init_process

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    stream = init_process(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.lib
    hbm_peak, peak_src = load_peaks()

    model, desc = make_model(args.workload)
    basis = model.basis()

    # ---- basis build (timed once) ------------------------------------------------------------
    def sync():
        torch.cuda.current_stream().synchronize()
        if world > 1:
            dist.barrier()

    total_candidates = basis.number_candidates
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # warm-up build of a small shard (module load, constant upload), then the timed full build
    wb = model.basis()
    r, n, c = wb.build_shard(0, min(total_candidates, 1 << 22))
    lib.ls_b200_device_free(r)
    if n:
        lib.ls_b200_device_free(n)
    sync()
    # The whole build (enumeration, host view, state -> index structure) three times, each from scratch on a fresh
    # basis object; the reported time is the median.  Its kernels take the same 46.5 ms every time, but the
    # cudaMalloc / cudaMallocManaged calls around them take anything from 3 to 150 ms on these boxes, so a single
    # sample says more about the driver's mood than about the build.  All samples are in the line.
    samples = []
    for attempt in range(3):
        if attempt > 0:
            basis = model.basis()
        sync()
        launches_b0 = lib.ls_b200_kernel_launch_count()
        t0 = time.perf_counter()
        ev0.record()
        if world > 1:
            build_sharded(basis)
        else:
            basis.build()
        ev1.record()
        sync()
        wall = time.perf_counter() - t0
        ms = ev0.elapsed_time(ev1)
        kernel_ms = lib.ls_b200_last_kernel_ms(b"build")
        if world > 1:
            t = torch.tensor([ms, kernel_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, kernel_ms = t.tolist()
        samples.append((ms, kernel_ms, wall))
    build_samples_ms = [m for m, _, _ in samples]
    build_ms, build_kernel_ms, build_wall = sorted(samples)[len(samples) // 2]
    build_launches = lib.ls_b200_kernel_launch_count() - launches_b0
    dim = basis.number_states

    op = model.operator(basis)
    sh = ShardedOperator(op)
    L = sh.layout
    nnz_total = op.count_matrix_elements(0, dim)
    nnz_local = op.count_matrix_elements(L.row_begin, L.row_end)
    elements_total = nnz_total + dim

    # ---- device-resident steps ------------------------------------------------------------------
    complex_vectors = model.symmetries is not None and not bool(
        np.all(np.abs(model.symmetries.characters()[1]) < 1e-9))
    vdtype = torch.complex128 if complex_vectors else torch.float64
    g = torch.Generator(device="cpu")
    g.manual_seed(42)
    x = sh.empty_vector(vdtype)
    host_x = torch.randn(dim, dtype=torch.float64, generator=g)
    if complex_vectors:
        host_x = torch.complex(host_x, torch.randn(dim, dtype=torch.float64, generator=g))
    x[:dim].copy_(host_x)
    y = sh.empty_vector(vdtype)
    small = (dim * 16) < (256 << 20)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None

    def step():
        if flush is not None:
            flush.fill_(1)
        sh.matvec(x, y)

    for _ in range(args.warmup):
        step()
    lib.ls_b200_matvec_sync()
    _lib.check_error()
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.ls_b200_kernel_launch_count()
    kernel_ms = []
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in events:
        if flush is not None:
            flush.fill_(1)
        a.record()
        sh.matvec(x, y)
        b.record()
    sync()
    step_ms = [a.elapsed_time(b) for a, b in events]
    # per-launch kernel time of the dominant kernel, CUDA events on the launching stream (library-side)
    per_kernel = []
    for _ in range(min(3, args.steps)):
        sh.matvec(x, y, gather=False)
        lib.ls_b200_matvec_sync()
        kernel_ms.append(lib.ls_b200_last_kernel_ms(b"matvec"))
        per_kernel.append({k: lib.ls_b200_last_kernel_ms(k.encode()) for k in
                           ("orbit", "orbit_launches", "gather", "gather_launches", "combine")})
    launches = lib.ls_b200_kernel_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(step_ms)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = elements_total / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the reference-facing vtable call ----------------------------------
    e2e = None
    if not args.no_e2e and not complex_vectors:
        if world == 1:
            nbytes = 8 * dim
            hx = lib.ls_b200_host_malloc(nbytes)
            hy = lib.ls_b200_host_malloc(nbytes)
            np.frombuffer((C.c_double * dim).from_address(hx), dtype=np.float64)[:] = host_x.numpy()
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
            yh = np.frombuffer((C.c_double * dim).from_address(hy), dtype=np.float64)
            assert np.isfinite(yh).all()
            lib.ls_b200_host_free(hx)
            lib.ls_b200_host_free(hy)
            e2e = {"value": elements_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": nbytes,
                   "d2h_bytes_per_step": nbytes, "ms_per_step": e2e_s * 1e3,
                   "call": "ls_chpl_kernels.matrix_vector_product (host pointers, pinned)"}
        else:
            rows = L.row_end - L.row_begin
            hx = torch.empty(L.chunk, dtype=torch.float64).pin_memory()
            hy = torch.empty(L.chunk, dtype=torch.float64).pin_memory()
            hx[:rows].copy_(host_x[L.row_begin:L.row_end])

            def e2e_step():
                x[L.row_begin:L.row_end].copy_(hx[:rows], non_blocking=True)
                sh.gather_rows(x)
                sh.matvec(x, y, gather=False)
                hy[:rows].copy_(y[L.row_begin:L.row_end], non_blocking=True)
                torch.cuda.current_stream().synchronize()
            e2e_step()
            sync()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step()
            sync()
            e2e_s = (time.perf_counter() - t0) / args.steps
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
            e2e = {"value": elements_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * L.chunk * world,
                   "d2h_bytes_per_step": 8 * L.chunk * world, "ms_per_step": e2e_s * 1e3,
                   "call": "per-rank pinned x shard -> NCCL all-gather -> ls_b200_matvec_device -> pinned y shard"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline -------------------------------------------------------------------------------------------
    # One y = H x is three kernels per row chunk: orbit_kernel (canonicalise every matrix element, integer
    # bound), rank_gather_kernel (state -> index, gather n_j x_j; latency / random-access bound), row_sum_kernel.
    # The dominant one is orbit_kernel; the bounding roofline of the path is integer issue, so the
    # mandated HBM figure is tiny by construction and the integer figures sit next to it.
    vec_bytes = 16 if complex_vectors else 8
    rows_local = L.row_end - L.row_begin
    pk = per_kernel[len(per_kernel) // 2] if per_kernel else {}
    orbit_launches = int(pk.get("orbit_launches", 0) or 0)
    group_size = len(model.symmetries.elements) if model.symmetries is not None else 0
    images = group_size * (2 if model.spin_inversion else 1)
    nbits = model.number_sites * (2 if model.particle != "spin-1/2" else 1)
    depth = 2 * max(1, (nbits - 1).bit_length()) - 1
    # W_m = I (6 depth + 4) + 3 ceil(log2 range) + 8 u64-ops per matrix element (un-pruned reference count)
    w_m = images * (6 * depth + 4) + 3 * 5 + 8
    step_ms = statistics.median(kernel_ms) if kernel_ms else ms_per_step
    if orbit_launches > 0:
        # orbit_kernel, per launch: reads alpha (8) + CSR offset (4) per row, writes representative (8) +
        # term/sign (2) [+ character index (1)] per matrix element
        per_element = 8 + 2 + (1 if complex_vectors else 0)
        algo_bytes = (nnz_local * per_element + rows_local * 12) / orbit_launches
        k_ms = pk["orbit"] / orbit_launches
        kernel_name = "orbit_kernel"
        ops = nnz_local * w_m / orbit_launches
    else:
        # no symmetry group: one gather kernel per step (x[j] gather per element; alpha, norm, x, y per row)
        algo_bytes = nnz_local * vec_bytes + rows_local * (16 + 2 * vec_bytes)
        k_ms = step_ms
        kernel_name = "gather_kernel"
        ops = (nnz_local + rows_local) * w_m
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    alu_pct = None
    profs = sorted((ROOT / "profiles").glob(f"r*/{kernel_name}_{args.workload}.summary.txt"))
    prof = profs[-1] if profs else None  # the latest committed ncu capture of this kernel on this workload
    if prof is not None:  # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel (MB)
        vals = dict(l.split()[:2] for l in prof.read_text().splitlines() if l.startswith("dram__bytes_"))
        if "dram__bytes_read.sum" in vals and "dram__bytes_write.sum" in vals:
            traffic = (float(vals["dram__bytes_read.sum"]) + float(vals["dram__bytes_write.sum"])) * 1e6
        for l in prof.read_text().splitlines():
            if l.startswith("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"):
                alu_pct = float(l.split()[1])
    lop3_peak = float(lib.ls_b200_measure_lop3_peak())  # measured LOP3 thread-instructions/s (alu pipe)
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
        "traffic": traffic, "peak_source": peak_src, "kernel": kernel_name, "kernel_ms": k_ms,
        "launches_per_step": orbit_launches or 1, "algorithmic_bytes_per_launch": algo_bytes,
        "note": "the path is integer-issue bound (SURVEY 8d): see int; step_kernels_ms is the per-step device time by kernel",
        "step_kernels_ms": {"orbit_kernel": pk.get("orbit"), "rank_gather_kernel": pk.get("gather"),
                            "row_sum_kernel": pk.get("combine"), "whole_step": step_ms},
        "int": {"reference_u64_ops_per_element": w_m, "achieved_Tops": ops / (k_ms * 1e-3) / 1e12,
                "lop3_peak_Tops": lop3_peak / 1e12,
                "alu_pipe_pct_of_peak_ncu": alu_pct,  # executed alu-pipe instructions vs peak, from the committed ncu capture
                "note": "achieved = un-pruned reference op count / orbit_kernel time; the bit-sliced kernel executes "
                        "~3 LOP3 per plane per group element for 32 states, so this exceeds the LOP3 peak"},
    }

    # ---- full-size, size-independent checks (the parity tests proper run at sizes the oracle can follow) ---------
    checks = {}
    if world == 1:
        states = np.asarray(basis.states[:min(dim, 1 << 22)])
        checks["representatives_sorted_prefix"] = bool(np.all(states[1:] > states[:-1]))
        if not complex_vectors:
            # Hermiticity of the projected operator: <u, H v> == <v, H u>
            gen = torch.Generator(device="cpu")
            gen.manual_seed(7)
            u = sh.empty_vector(vdtype)
            u[:dim].copy_(torch.randn(dim, dtype=torch.float64, generator=gen))
            hu, hv = sh.empty_vector(vdtype), sh.empty_vector(vdtype)
            sh.matvec(u, hu)
            sh.matvec(x, hv)
            lib.ls_b200_matvec_sync()
            _lib.check_error()
            a = float(torch.dot(u[:dim], hv[:dim]).item())
            b = float(torch.dot(x[:dim], hu[:dim]).item())
            scale = float(torch.linalg.vector_norm(u[:dim]).item() * torch.linalg.vector_norm(hv[:dim]).item())
            checks["hermiticity_rel_err"] = abs(a - b) / max(scale, 1e-300)
            del u, hu, hv

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
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "c128" if complex_vectors else "f64", "data": "synthetic",
        "config": {"workload": desc, "name": args.workload, "dim": dim, "candidates": total_candidates,
                   "off_diag_elements": nnz_total, "l2": "flushed between steps" if small else "inputs larger than L2",
                   "parallelism": (f"rows sharded over {world} rank(s) by matrix-element count, result all-gathered"
                                   if world > 1 else "single GPU")},
        "build": {"candidates_per_s": total_candidates / (build_ms * 1e-3), "representatives_per_s": dim / (build_ms * 1e-3),
                  "ms": build_ms, "samples_ms": build_samples_ms, "kernel_ms_last_shard": build_kernel_ms,
                  "wall_ms": build_wall * 1e3,
                  "gpu_launches": int(build_launches), "unit": "states/s"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "checks": checks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
