#!/usr/bin/env python
"""Benchmark of the hot path: fused Chebyshev H·X steps on a 10^6-site BdG Hamiltonian.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C5] [--cols 8]

One "step" = one recursion step T_{n+1} = 2 H~ T_n - T_{n-1} (+ the two moment dot products) over one
tile of ``cols`` vectors (default 8) per GPU.  The single-step kernels make one HBM pass over the
matrix and the vectors per step; the default kernel on this workload (``pair``) does TWO steps per
launch, so K steps are K/2 launches (``gpu_launches``, ``roofline.steps_per_launch``).  Workload: BASELINE.json config C5, CubicLattice((1000,1000,1))
altermagnet/superconductor Josephson junction, 10^6 sites, 4,996,000 BSR blocks, synthetic
(SURVEY 8d).  N > 1: one process per GPU (torchrun), a replica of the matrix and its own 8 columns
on every GPU (weak scaling), one NCCL all-reduce of the moments at the end of the timed region.

Prints ONE JSON line (see the task contract): ``value`` = whole-job steps/s with everything
resident in HBM; ``e2e`` = the same metric through the public API with the Hamiltonian terms in
pinned HOST memory (upload + scatter + recursion + moments back to the host inside the timed
region); ``roofline`` = algorithmic bytes per step / measured kernel time vs the measured HBM
peak; ``cpu_baseline`` = scipy's bsr_matvecs recursion on the host cores (oracle port).

``--impl reference`` times only that CPU path (rank 0), on the same config/metric/unit.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "chebyshev_spmm_steps_per_s"
UNIT = "steps/s"
FALLBACK_HBM_GBS = 6650.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C5")
    ap.add_argument("--cols", type=int, default=8, help="vector columns per GPU (weak scaling, the default)")
    ap.add_argument("--total-cols", type=int, default=0,
                    help="strong scaling instead: this many columns in total, split evenly over the GPUs (SURVEY 8e: 64)")
    ap.add_argument("--kernel", default="auto_moments",
                    help="auto_moments = what chebyshev_moments / free_energy / ldos use (only moments are read)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-plain", action="store_true", help="skip the uncompressed-matrix comparison run")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload_key):
    """DRAM bytes per launch of the step kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(workload_key)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi sampled in the background during the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, watts, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            try:
                watts.append(float(r[2]))
            except (ValueError, IndexError):
                pass
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w": statistics.median(watts) if watts else None}


# ------------------------------------------------------------------------------------------
# reference arm / CPU baseline: scipy bsr_matvecs recursion on the host (oracle port)
# ------------------------------------------------------------------------------------------
def cpu_chebyshev(cfg_key, cols, steps, warmup, budget_s):
    from bodge_b200 import workloads
    from oracle import bdg_oracle as orc
    from oracle import cpu_baseline as cb

    cfg = workloads.CONFIGS[cfg_key]
    t0 = time.perf_counter()
    packed = cfg["build"](cfg["shape"])
    ptr, idx, dat = cb.assemble(cfg["shape"], packed)
    t_asm = time.perf_counter() - t0
    scale = 1.01 * orc.norm_inf(ptr, idx, dat)
    x0 = orc.rademacher(1234, 4 * (len(ptr) - 1), np.arange(cols))
    res = cb.time_steps(ptr, idx, dat, scale, x0, steps=steps, warmup=warmup, budget_s=budget_s)
    res["assembly_s"] = t_asm
    res["n_sites"] = len(ptr) - 1
    res["n_blocks"] = len(idx)
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bodge_b200 import workloads

    cfg = workloads.CONFIGS[args.config]
    res = cpu_chebyshev(args.config, args.cols, args.steps, args.warmup, budget_s=120.0)
    value = 1e3 / res["ms_per_step"]
    sample = (f"{args.steps} steps after {args.warmup} warm-up on {res['fraction']:.3f} of the block rows per step "
              f"(time scaled to a full step), k={args.cols}, rows split over {res['cores']} processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["label"], "config": args.config, "n_sites": res["n_sites"], "n_blocks": res["n_blocks"],
                   "cols_per_gpu": args.cols, "what": "scipy bsr_matvecs recursion 2*(H~@T1)-T0 on the host (oracle port of the reference path)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def dict_api_assembly(b, device):
    """BASELINE config C2 through the reference's own surface -- ``with system as (H, Δ)`` and two
    Python loops (README.md:73-86) -- to ``matrix("bsr")``.  The user's loops are interpreter time
    no library can remove; ``library_s`` is everything else (skeleton, dict packing,
    H2D, scatter + symmetry fill + Hermitian check, zero-block compaction, D2H of the BSR arrays)."""
    shape = (100, 100, 1)
    t0 = time.perf_counter()
    lattice = b.CubicLattice(shape)
    system = b.Hamiltonian(lattice, device=device)
    t1 = time.perf_counter()
    with system as (H, D):
        for i in lattice.sites():
            H[i, i] = 3.0 * b.σ0 - 0.05 * b.σ3
            D[i, i] = -0.10 * b.jσ2
        for i, j in lattice.bonds():
            H[i, j] = -1.0 * b.σ0
        t2 = time.perf_counter()
    t3 = time.perf_counter()
    bsr = system.matrix("bsr")
    t4 = time.perf_counter()
    return {"config": "C2 CubicLattice((100,100,1)) README s-wave via the dict API", "n_sites": lattice.size,
            "n_blocks": int(len(bsr.indices)), "sites_per_s": lattice.size / (t4 - t0),
            "library_sites_per_s": lattice.size / ((t1 - t0) + (t3 - t2) + (t4 - t3)),
            "user_loop_s": t2 - t1, "skeleton_s": t1 - t0, "exit_s": t3 - t2, "export_bsr_s": t4 - t3,
            "reference": "6.7-7.1 k sites/s end to end measured for the reference in the build container (SURVEY 6.2)"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import bodge_b200 as b
    from bodge_b200 import workloads

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner to STDOUT at NCCL_DEBUG=VERSION/INFO; stdout carries the JSON line only
        os.environ["NCCL_DEBUG"] = os.environ.get("BDG_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = workloads.CONFIGS[args.config]
    shape, cols, K, W = cfg["shape"], args.cols, args.steps, args.warmup
    scaling = "weak"
    if args.total_cols > 0:
        if args.total_cols % world:
            raise SystemExit(f"--total-cols {args.total_cols} is not divisible by {world} GPUs")
        cols, scaling = args.total_cols // world, "strong"

    # ---- inputs: Hamiltonian terms as packed arrays in pinned host memory --------------------
    packed = cfg["build"](shape)
    pinned = []
    for arr in packed:
        t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
        pinned.append(t)
    host = [t.numpy() for t in pinned]
    h2d_bytes = sum(a.nbytes for a in host)

    # ---- assembly (timed separately: the second half of BASELINE.json's metric) ---------------
    # One untimed warm-up (lazy module loading, first allocations), then the median of three full
    # assemblies: skeleton on the device, H2D of the packed Hamiltonian terms from pinned host memory,
    # scatter + symmetry fill + Hermitian check, zero-block compaction.  Everything a Hamiltonian owns
    # is released in between (bdg_destroy), so every repetition allocates its 1.3 GB again.
    def assemble_once():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sysm = b.Hamiltonian(b.CubicLattice(shape), device=local)
        sysm._sys.sync()
        t1 = time.perf_counter()
        dev = sysm.fill(*host)
        sysm._sys.sync()
        t2 = time.perf_counter()
        nfo = sysm._sys.cheb_info()  # builds the compacted BSR the Chebyshev engine consumes
        sysm._sys.sync()
        t3 = time.perf_counter()
        return sysm, dev, nfo, (t1 - t0, t2 - t1, t3 - t2)

    system, max_dev, info0, cold = assemble_once()
    reps = []
    for _ in range(3):
        del system
        system, max_dev, info0, times = assemble_once()
        reps.append(times)
    reps.sort(key=sum)
    t_skel, t_fill, t_pack = reps[1]
    n_sites = system.lattice.size
    assembly = {
        "sites_per_s": n_sites / (t_skel + t_fill + t_pack), "unit": "sites/s", "n_sites": n_sites,
        "n_blocks": info0["n_blocks"], "skeleton_s": t_skel, "h2d_scatter_check_s": t_fill, "compaction_s": t_pack,
        "cold_first_call_s": sum(cold), "h2d_bytes": h2d_bytes, "hermitian_dev": max_dev,
        "what": "median of 3 after 1 warm-up: bdg_create_cubic + bdg_scatter (pinned host arrays -> device, symmetry "
                "fill, Hermitian check) + zero-block compaction",
    }

    if rank == 0:
        assembly["dict_api"] = dict_api_assembly(b, local)

    scale = system.spectral_bound()
    s = system._sys
    stream = torch.cuda.Stream(device=local)  # the library launches on this stream; events are recorded on it
    torch.cuda.set_stream(stream)
    s.set_stream(stream.cuda_stream)

    # ---- device-resident timing ------------------------------------------------------------------
    def timed_steps(kernel):
        """W warm-up + K timed steps of `kernel`, then the single exchange of the path (moments of all
        column shards).  Returns (total_ms, kernel_ms, launches, info, fmt), max over ranks."""
        s.cheb_begin(n_random=cols, seed=1234, col_offset=rank * cols, scale=scale, kernel=kernel)
        s.cheb_reserve(W + K + 8)
        s.cheb_steps(W)
        info, fmt = s.cheb_info(), s.cheb_format()
        launches0 = info["launches"]
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        s.cheb_steps(K)
        e1.record(stream)
        n_mom = 2 * (W + K + 1)
        s.cheb_read(n_mom, cols, summed=True, device_ptr=mu_dev.data_ptr())
        if world > 1:
            dist.all_reduce(mu_dev[:n_mom])
        e2.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times = torch.tensor([e0.elapsed_time(e2), e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
        total, kern = (float(v) for v in times.cpu())
        return total, kern, s.cheb_info()["launches"] - launches0, info, fmt

    mu_dev = torch.empty(2 * (W + K + 1), dtype=torch.float64, device=f"cuda:{local}")
    with ClockSampler(local) as clocks:
        total_ms, kernel_ms, launches, info, fmt = timed_steps(args.kernel)
    mu0 = float(mu_dev[0].cpu())
    assert abs(mu0 - 4.0 * n_sites * cols * world) < 1e-6 * mu0, "moment 0 must equal the number of vector entries"
    # The same steps on the uncompressed fixed-width matrix copy (every block read from HBM): shows
    # what the block dictionary buys and how close the plain kernel runs to the HBM roofline.
    plain = None
    if (fmt["kernel"].startswith("dict") or fmt["kernel"] in ("pair", "t2")) and not args.no_plain:
        p_total, p_kernel, _, _, p_fmt = timed_steps("ell")
        plain = {"kernel": "cheb_step_ell (every block from HBM)", "kernel_ms_per_launch": p_kernel / K,
                 "steps_per_s": world * K / (p_total * 1e-3), "matrix_bytes_per_launch": p_fmt["matrix_bytes_per_step"]}

    # The literal three-term recursion T_{n+1} = 2 H~ T_n - T_{n-1} (every T_n materialised) for comparison when the
    # headline ran the even-vector form of it: the pair kernel (two steps per launch) and the single-step kernel.
    three_term = None
    if fmt["kernel"] == "t2" and not args.no_plain:
        three_term = {}
        for name in ("pair", "dict_diag"):
            try:
                t_total, t_kernel, t_launches, _, t_fmt = timed_steps(name)
            except (ValueError, RuntimeError):
                continue
            three_term[t_fmt["kernel"]] = {"steps_per_s": world * K / (t_total * 1e-3), "kernel_ms_per_step": t_kernel / K}

    # ---- end to end through the public API with host buffers -------------------------------------
    e2e = None
    if not args.no_e2e:
        s.set_stream(None)
        e2e_steps = 1024  # 2050 moments: the recursion length of the C5 free-energy evaluation
        calls = 2

        def one_call():
            system.fill(*host)  # H2D of all Hamiltonian terms + scatter + Hermitian check
            mu = system.chebyshev_moments(2 * e2e_steps + 2, vectors=cols * world, seed=1234, summed=True, kernel=args.kernel)
            return mu  # host numpy (D2H inside)

        one_call()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(calls):
            mu = one_call()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.cpu())
        e2e = {"value": world * calls * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": h2d_bytes / e2e_steps, "d2h_bytes_per_step": mu.nbytes / e2e_steps,
               "what": f"{calls} x [Hamiltonian.fill(pinned host arrays) + chebyshev_moments({2 * e2e_steps + 2}) -> host]; "
                       f"{e2e_steps} steps per call", "seconds": dt}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel ---------------------------------------------------
    # `achieved` follows the contract: ALGORITHMIC bytes (every 4x4 block + index + three
    # vector passes) / measured kernel time.  With the block-dictionary format the kernel moves fewer
    # bytes than that (codes instead of blocks), so `frac` can exceed 1; `moved_*` is what it really
    # streams, and `plain` the same steps on the uncompressed copy.
    peak, peak_src = measured_peak()
    bytes_step = info["bytes_per_step"]
    achieved = bytes_step * K / (kernel_ms * 1e-3) / 1e9
    # per step: three vector passes; the pair kernel (two steps per launch) moves four per two steps
    # (the even-vector recursion "t2" moves three per two steps)
    moved_step = fmt["matrix_bytes_per_step"] + {"pair": 128, "t2": 96}.get(fmt["kernel"], 192) * n_sites * cols
    moved = moved_step * K / (kernel_ms * 1e-3) / 1e9
    # `launches` also counts the moment read-out kernel (and, for t2, the kernel that normalises its dot rows)
    step_launches = max(launches - (2 if fmt["kernel"] == "t2" else 1), 1)
    steps_per_launch = K / step_launches
    kernel_name = {"t2": "cheb_pair_step<MODE=T2> (two applications of H~ per launch on the even vectors E_j = T_2j x: "
                         "E_{j+1} = 2 T_2(H~) E_j - E_{j-1}; block-dictionary matrix, E_j planes staged in shared memory by TMA "
                         "bulk copies, H~ E_j kept in shared memory, E_{j+1} written over E_{j-1})",
                   "pair": "cheb_pair_step (two steps per launch: block-dictionary matrix, T_n planes staged in shared memory by "
                           "TMA bulk copies, T_{n-1} straight to registers, T_{n+1} kept in shared memory)",
                   "dict": "cheb_step_ell<DICT> (block-dictionary matrix)",
                   "dict_diag": "cheb_step_ell<DICT,DIAG> (block-dictionary matrix, real-diagonal hopping blocks by DFMA)",
                   "ell": "cheb_step_ell",
                   "dmma": "cheb_step_dmma", "fma": "cheb_step_fma"}.get(fmt["kernel"], fmt["kernel"])
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": recorded_traffic(f"{args.config}_k{cols}_{fmt['kernel']}"), "peak_source": peak_src,
                "kernel": kernel_name, "steps_per_launch": steps_per_launch,
                "algorithmic_bytes_per_launch": bytes_step * steps_per_launch,
                "kernel_ms_per_launch": kernel_ms / step_launches, "matrix_format": fmt["kernel"],
                "distinct_blocks": fmt["distinct_blocks"], "moved_bytes_per_launch": moved_step * steps_per_launch,
                "moved_GBps": moved, "moved_frac": moved / peak}
    if three_term:
        roofline["three_term_recursion"] = three_term
    if plain is not None:
        plain["achieved"] = bytes_step / (plain["kernel_ms_per_launch"] * 1e-3) / 1e9
        plain["frac"] = plain["achieved"] / peak
        plain["traffic"] = recorded_traffic(f"{args.config}_k{cols}_ell")
        roofline["plain"] = plain

    cpu = None
    if not args.no_cpu_baseline:
        res = cpu_chebyshev(args.config, cols, steps=3, warmup=1, budget_s=20.0)
        cpu = {"value": 1e3 / res["ms_per_step"], "unit": UNIT, "cores": res["cores"], "kind": "port",
               "sample": f"3 steps after 1 warm-up of the same workload (k={cols}) on {res['fraction']:.3f} of the block rows, "
                         f"scipy bsr_matvecs, rows split over {res['cores']} processes; numpy assembly took {res['assembly_s']:.1f} s",
               "assembly_sites_per_s": res["n_sites"] / res["assembly_s"]}

    line = {
        # weak scaling: every GPU advances its own 8-column job, the job count adds up; strong scaling:
        # one job of --total-cols columns, a step is done when every shard has done it
        "metric": METRIC, "value": (world if scaling == "weak" else 1) * K / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": cfg["label"], "config": args.config, "n_sites": n_sites, "n_blocks": info["n_blocks"],
                   "cols_per_gpu": cols, "parallelism": f"column shards x{world}, matrix replicated",
                   "l2": "inputs (1.3 GB matrix + 1.0 GB vectors) exceed the 126 MB L2; no flush needed",
                   "kernel": fmt["kernel"], "panel_width": info["panel_width"],
                   "recursion": ("even-vector form E_{j+1} = 2 T_2(H~) E_j - E_{j-1}, E_j = T_2j(H~) x: one launch = two applications "
                                 "of H~ = two steps = four moments (roofline.three_term_recursion: the literal recursion)"
                                 if fmt["kernel"] == "t2" else "three-term T_{n+1} = 2 H~ T_n - T_{n-1}")},
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "assembly": assembly,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to
    fd 1 when it creates a communicator, at NCCL_DEBUG=VERSION and WARN alike), so keep a private duplicate of
    the real stdout for the JSON line and point fd 1 at stderr for everything else."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
