#!/usr/bin/env python
"""bench.py — log-likelihood evaluations per second for a walker ensemble (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--solver dense_i8|dense]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (covariance build → Cholesky → solve → lnL) over the whole ensemble:
256 walkers at N=8192 pixels, M=6 eigenspectra, global + 2 local kernels (BASELINE.json configs[2]; the
8-GPU line is configs[3]).  The ensemble is fixed, so multi-GPU runs are STRONG scaling: rank r owns walkers
[r·B/G, (r+1)·B/G) and the only exchange is one NCCL all-gather of the lnL scalars per step, issued by the
library itself (sfb_allgather_lnL).

Prints ONE JSON line (rank 0):
  value        walkers / step time with inputs resident in HBM, dense factorisation with the trailing update on
               the int8 tensor path (solver dense_i8: exact products of 48-bit fixed-point operands, panels in true
               fp64 — parity block below); `fp64_dmma` is the same step with the pure-fp64 DMMA trailing update
  e2e          the same through the host-buffer C-ABI call (H2D of X/A/model_flux/hyper-parameters and D2H of
               lnL/info inside the timed region);  e2e_model = SpectrumModel.log_likelihood_batch on host parameters
  parity       GPU lnL of the first walkers of this very run against the dense CPU oracle (the cpu_baseline sample)
  roofline     the dominant kernel (trailing update) against its tensor peak
  configs      BASELINE.json configs[1] (64 walkers, N=4096, global kernel only) and a configs[4] shard
               (64 walkers per GPU, N=16384, 16 local kernels): value, fraction of the fp64 floor, kernel shares
  frozen_shared  the reference's frozen-kernel mode: all walkers share the kernel hyper-parameters, S factorised once
  structured   the structure-exploiting solver (row f4) on the same inputs
  cpu_baseline the CPU oracle port on this box's host cores, both threading modes, on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = dict(n_pix=8192, n_walkers=256, n_comp=6, n_local=2)
METRIC = "log-likelihood evals/sec, 256 walkers, N=8192 pixels"
DEFAULT_SOLVER = "dense_i8"
# fp64 tensor (DMMA m8n8k4) peak measured on this pool's B200 with tools/fp64_peak.cu
# (profiles/r01_fp64_peak.txt): MEASURED_PEAKS.json carries no fp64 entry.
FP64_DMMA_PEAK_TFLOPS = 37.1
FP64_DFMA_PEAK_TFLOPS = 33.8   # plain DFMA issue peak from the same measurement (the band kernels use DFMA)
# int8 MMAs per fp64 multiply-add in the dense_i8 trailing update (csrc/ozaki.cu: 6 slices, anti-diagonals 0..6)
I8_PRODUCTS = 26
LNL_RTOL = 1e-10


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-pix", type=int, default=WORKLOAD["n_pix"])
    ap.add_argument("--walkers", type=int, default=WORKLOAD["n_walkers"])
    ap.add_argument("--solver", default=DEFAULT_SOLVER, choices=["dense", "dense_i8"],
                    help="trailing update of the dense factorisation: int8 tensor-core restatement or fp64 DMMA")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-model", action="store_true", help="skip the parameter-level (drop-in model) legs")
    ap.add_argument("--no-structured", action="store_true", help="skip the structured-solver (row f4) legs")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[1] / configs[4] legs")
    ap.add_argument("--no-alt", action="store_true", help="skip the leg with the other dense trailing-update mode")
    ap.add_argument("--no-frozen", action="store_true", help="skip the frozen-kernel (shared factor) leg")
    ap.add_argument("--cpu-sample", type=int, default=0, help="walkers in the CPU sample (0 = one per usable core)")
    return ap.parse_args()


def flops_eval(N, M):
    """whole-path fp64 FLOPs per evaluation: N^3/3 + 2MN^2 + 2N^2"""
    return N ** 3 / 3 + 2 * M * N ** 2 + 2 * N ** 2


def solver_note(solver):
    return ("dense Cholesky, trailing update on the int8 tensor path (exact int8 x int8 -> int32 products of 48-bit "
            "fixed-point operands, fp64 recombination; panels, solves and logdet in fp64)" if solver == "dense_i8"
            else "dense fp64 Cholesky (DMMA trailing update)")


def config_dict(args, world):
    B = args.walkers
    return {
        "workload": f"{B}-walker ensemble, N={args.n_pix} px, M={WORKLOAD['n_comp']} eigenspectra, global "
                    f"Matern + {WORKLOAD['n_local']} local kernels, {solver_note(args.solver)} "
                    f"(BASELINE.json configs[{2 if world == 1 else 3}])",
        "n_walkers": B, "n_pix": args.n_pix, "n_comp": WORKLOAD["n_comp"], "n_local": WORKLOAD["n_local"],
        "solver": args.solver,
        "walkers_per_gpu": B // world if B % world == 0 else f"{B // world}-{B // world + 1}",
        "parallelism": f"walker-sharded x{world}, one all-gather of lnL per step (ncclAllGather inside libsfb200)",
        "cache": "inputs + per-walker N^2 factorisation workspace (>= 16 GiB per step) far exceed the 126 MB L2; "
                 "no flush needed",
    }


# ---------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# reference arm / CPU baseline (the only places that touch oracle/)
# ---------------------------------------------------------------------------------------------------
def host_description():
    import scipy

    d = {"os_cpu_count": os.cpu_count(), "numpy": np.__version__, "scipy": scipy.__version__}
    try:
        d["sched_affinity"] = len(os.sched_getaffinity(0))
    except Exception:
        d["sched_affinity"] = None
    try:
        from threadpoolctl import threadpool_info

        d["blas"] = sorted({f"{i.get('internal_api')} {i.get('version')} ({i.get('num_threads')} threads)"
                            for i in threadpool_info() if i.get("user_api") == "blas"})
    except Exception:
        d["blas"] = None
    try:
        import psutil

        d["mem_available_gb"] = round(psutil.virtual_memory().available / 2 ** 30, 1)
    except Exception:
        d["mem_available_gb"] = None
    return d


def cpu_measure(args, sample_walkers=None, serial=True):
    """The oracle port on the host cores, both threading modes of BASELINE.md §4:
    pool = P worker processes x 1 BLAS thread (emcee-pool style), P = every usable core unless host memory
           (~8 N^2 doubles per worker) says fewer;   serial = one walker at a time, BLAS free to use every core.
    Returns the better of the two as `value` and the lnL of the sampled walkers (used for the parity block)."""
    from oracle import cpu_bench
    from starfish_b200 import synth

    host = host_description()
    cores = cpu_bench.host_cores()
    N = args.n_pix
    per_worker_gb = 8 * N * N * 8 / 2 ** 30
    mem_cap = int(0.6 * host["mem_available_gb"] / per_worker_gb) if host.get("mem_available_gb") else cores
    procs = max(1, min(cores, mem_cap))
    n = sample_walkers or args.cpu_sample or procs
    procs = min(procs, n)
    stage = synth.stage_inputs_direct(N, n, n_comp=WORKLOAD["n_comp"], n_local=WORKLOAD["n_local"])
    dt, lnl = cpu_bench.time_pool(stage, range(n), procs)
    out = dict(value=n / dt, seconds=dt, cores=procs, n=n, lnl=lnl, host=host,
               pool={"evals_per_s": n / dt, "processes": procs, "walkers": n, "seconds": dt,
                     "memory_capped": bool(mem_cap < cores)},
               mode="pool")
    if serial:
        ns = 1 if N >= 8192 else 2
        dts, _ = cpu_bench.time_serial(stage, range(ns))
        out["serial"] = {"evals_per_s": ns / dts, "blas_threads": "all", "walkers": ns, "seconds": dts}
        if ns / dts > out["value"]:
            out.update(value=ns / dts, seconds=dts, cores=cores, mode="serial")
    out["sample"] = (f"{n} of {args.walkers} walkers at N={N}: pool mode {procs} worker processes x 1 BLAS thread = "
                     f"{n / dt:.3f} evals/s in {dt:.1f} s" +
                     (f"; serial mode (all-core BLAS) = {out['serial']['evals_per_s']:.3f} evals/s" if serial else "") +
                     f"; quoted: {out['mode']} (numpy/scipy oracle port of the reference path; host has "
                     f"{host['os_cpu_count']} CPUs, {host['sched_affinity']} usable)")
    return out


def cpu_baseline_block(m):
    return {"value": m["value"], "unit": "evals/s", "cores": m["cores"], "kind": "port", "sample": m["sample"],
            "mode": m["mode"], "pool": m["pool"], "serial": m.get("serial"), "host": m["host"]}


def run_reference(args, rank, world):
    if rank != 0:
        return
    t_all = time.perf_counter()
    vals = []
    for i in range(args.warmup + args.steps):
        # every step is a bounded sample of the ensemble (one walker per usable core); warm-up steps use one walker
        m = cpu_measure(args, sample_walkers=1 if i < args.warmup else None, serial=(i == args.warmup))
        if i >= args.warmup:
            vals.append(m)
    best = sorted(vals, key=lambda m: m["seconds"])[len(vals) // 2]
    first = vals[0]
    value = float(np.mean([m["pool"]["evals_per_s"] for m in vals]))
    if first.get("serial") and first["serial"]["evals_per_s"] > value:
        value = first["serial"]["evals_per_s"]
    ms_per_step = float(np.mean([m["pool"]["seconds"] for m in vals])) * 1e3
    cb = cpu_baseline_block(first)
    cb.update(value=value, cores=best["cores"],
              sample=best["sample"] + f" — each reference step is this {best['n']}-walker sample, not the 256-walker "
                                      "ensemble (ms_per_step is per sample; value is the rate)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, max(world, 1)),
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t_all,
    }
    line["config"]["solver"] = "scipy cho_factor / cho_solve (CPU oracle port)"
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------------
class Timer:
    """W warm-up calls, then K timed calls bracketed by barrier + synchronize, CUDA events, max over ranks."""

    def __init__(self, torch, dist, dev, world, eng):
        self.torch, self.dist, self.dev, self.world, self.eng = torch, dist, dev, world, eng

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def __call__(self, fn, steps, warmup, eng=None):
        torch = self.torch
        eng = eng or self.eng
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = eng.launch_count
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item()), eng.launch_count - launches0, out


def kernel_shares(prof):
    tot = sum(v["ms"] for v in prof.values())
    return {k: {"launches": v["launches"], "ms": round(v["ms"], 3), "share": round(v["ms"] / tot, 4) if tot else None}
            for k, v in prof.items() if v["launches"]}


def roofline_block(prof, solver, peaks, issued_frac=1.0):
    sy = prof["syrk"]
    total_ms = sum(v["ms"] for v in prof.values())
    ach = sy["work"] / (sy["ms"] * 1e-3) / 1e12 if sy["ms"] > 0 else 0.0   # fp64(-equivalent) TFLOP/s
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            key = "syrk_i8" if solver == "dense_i8" else "syrk"
            traffic = tj.get(f"{key}_dram_bytes_per_launch")
            traffic_src = tj.get(f"{key}_source")
        except Exception:
            pass
    other = {k: {"launches": v["launches"], "ms": v["ms"],
                 "achieved": (v["work"] / (v["ms"] * 1e-3) / (1e9 if k in ("build", "oz_slice") else 1e12))
                 if v["ms"] > 0 else None,
                 "unit": "GB/s" if k in ("build", "oz_slice") else "TFLOP/s"}
             for k, v in prof.items() if k in ("build", "potrf_diag", "trsm", "oz_slice") and v["launches"]}
    if "build" in other:
        other["build"]["peak_hbm_gbs"] = peaks.get("hbm_gbs")
    common = {"bound": "tensor", "unit": "TFLOP/s", "traffic": traffic,
              "traffic_source": traffic_src or "ncu --set full capture of one launch (see profiles/README.md); a "
                                               "constant from that capture, not a measurement of this run",
              "launches": sy["launches"], "avg_launch_ms": sy["ms"] / max(sy["launches"], 1),
              "flops_per_launch": sy["work"] / max(sy["launches"], 1),
              "share_of_step": sy["ms"] / total_ms if total_ms else None,
              "timing": "one extra profiled step after the timed region: single stream, CUDA events around every "
                        "launch", "other_kernels": other}
    if solver == "dense_i8":
        i8 = ach * I8_PRODUCTS * issued_frac   # int8 TOP/s actually executed on the tensor pipe
        peak = 2.0 * peaks["bf16_tflops_sustained"] if peaks.get("bf16_tflops_sustained") else 2.0 * 1400.0
        common.update({
            "kernel": "syrk_i8_kernel (trailing update A_ij -= L_ik L_jk^T as 26 exact int8 digit-slab products per "
                      "fp64 product term, issued as 9 wide tcgen05.mma kind::i8 per 32-deep chunk (runs of "
                      "consecutive slabs, N = 128..256), 7 int32 TMEM accumulators, fp64 recombination; CTA pairs "
                      "share the A operand by multicast bulk copies)",
            "achieved": i8, "peak": peak, "frac": i8 / peak,
            "achieved_fp64_equivalent_tflops": ach, "fp64_dmma_peak_tflops": FP64_DMMA_PEAK_TFLOPS,
            "int8_mma_issued_frac": issued_frac,
            "speedup_over_fp64_tensor_peak": ach / FP64_DMMA_PEAK_TFLOPS,
            "peak_source": "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (int8 dense rate = 2 x bf16 on sm_100a; the "
                           "file has no int8 entry" + ("" if peaks.get("bf16_tflops_sustained") else "; FALLBACK 1.4 PF") +
                           "); in isolation the kernel's MMA schedule runs at the tensor floor (832 cycles per "
                           "dense chunk, profiles/r3e_umma_i8_wide_run_probe.txt); in the step it is paced by operand "
                           "delivery (36 KB per chunk and SM, ~10 TB/s L2 -> SM) and the per-tile TMEM drain",
            "algorithmic_work": "n(n+1)K fp64 FLOP per trailing update x 26 int8 slab products x int8_mma_issued_frac "
                                "(products with an all-zero digit slab are skipped — exact; the fraction is counted "
                                "by the kernel from the digit-slab flags over the profiled step)"})
    else:
        common.update({
            "kernel": "syrk_kernel (trailing update A_ij -= L_ik L_jk^T, DMMA m8n8k4 fp64)",
            "achieved": ach, "peak": FP64_DMMA_PEAK_TFLOPS, "frac": ach / FP64_DMMA_PEAK_TFLOPS,
            "peak_source": "fp64 DMMA peak measured on this pool's B200 by tools/fp64_peak.cu "
                           "(MEASURED_PEAKS.json has no fp64 entry; nominal 37.2 TFLOP/s)"})
    return common


def run_config_leg(torch, dist, dev, local_rank, world, rank, stage, N, M, K, nb, total_walkers, label, solvers,
                   steps, warmup):
    """One BASELINE config on its own handle: timed value, fraction of the fp64 floor, kernel shares per solver."""
    from starfish_b200.engine import LikelihoodEngine

    eng = LikelihoodEngine(N, M, K, nb, device=local_rank)
    eng.set_data(stage["wave"], stage["sigma"], stage["data_flux"])
    X = torch.from_numpy(stage["X"]).to(dev) if stage["X"] is not None else None
    A = torch.from_numpy(stage["A"]).to(dev) if stage["A"] is not None else None
    F = torch.from_numpy(stage["model_flux"]).to(dev)
    g, n, l = eng.pack_hyper(nb, stage["glob"], stage["nloc"] if K else None, stage["loc"] if K else None, False)
    lnL = torch.empty(nb, dtype=torch.float64, device=dev)
    info = torch.empty(nb, dtype=torch.int32, device=dev)
    timer = Timer(torch, dist, dev, world, eng)
    fl = flops_eval(N, M)
    floor_per_gpu = FP64_DMMA_PEAK_TFLOPS * 1e12 / fl
    out = {"workload": label, "n_pix": N, "n_comp": M, "n_local": K, "walkers": total_walkers,
           "walkers_per_gpu": nb, "workspace_walkers": eng.workspace_walkers,
           "fp64_floor_evals_per_s_per_gpu": floor_per_gpu, "flops_per_eval": fl}
    ref = None
    for solver in solvers:
        eng.set_solver(solver)

        def step():
            eng.log_likelihood_resident(nb, X, A, F, g, n, l, lnL, info)

        ms, launches, _ = timer(step, steps, warmup)
        value = total_walkers / (ms / steps * 1e-3)
        leg = {"value": value, "unit": "evals/s", "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
               "gpu_launches": int(launches), "path_tflops_fp64_equivalent": value * fl / 1e12,
               "frac_of_fp64_floor": value / (floor_per_gpu * world),
               "not_positive_definite": int((info != 0).sum().item())}
        cur = lnL.clone()
        if ref is None:
            ref = cur
        else:
            leg["max_rel_diff_vs_" + solvers[0]] = float(((cur - ref).abs() / ref.abs().clamp(min=1.0)).max().item())
        if rank == 0:
            eng.profile(True)
            eng.i8_mma_counts()
            step()
            leg["kernels"] = kernel_shares(eng.profile_read())
            if solver == "dense_i8":
                iss, den = eng.i8_mma_counts()
                leg["int8_mma_issued_frac"] = iss / den if den else None
            eng.profile(False)
        out[solver] = leg
    eng.close()
    return out


def run_b200(args, rank, world, local_rank, cpu_sample):
    import torch
    import torch.distributed as dist

    from starfish_b200 import synth
    from starfish_b200.dist import gather_lnl, init_engine_comm, shard_range
    from starfish_b200.engine import LikelihoodEngine

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, N, M, K = args.walkers, args.n_pix, WORKLOAD["n_comp"], WORKLOAD["n_local"]
    lo, hi = shard_range(B, rank, world)
    nb = hi - lo
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {}

    # synthetic stage inputs for this rank's shard (seeded per walker -> identical across world sizes)
    stage = synth.stage_inputs_direct(N, nb, n_comp=M, n_local=K, first_walker=lo)
    eng = LikelihoodEngine(N, M, K, max(nb, 1), device=local_rank)
    eng.set_data(stage["wave"], stage["sigma"], stage["data_flux"])
    eng.set_solver(args.solver)
    X = torch.from_numpy(stage["X"]).to(dev)
    A = torch.from_numpy(stage["A"]).to(dev)
    F = torch.from_numpy(stage["model_flux"]).to(dev)
    g, n, l = eng.pack_hyper(nb, stage["glob"], stage["nloc"], stage["loc"], False)
    lnL = torch.empty(nb, dtype=torch.float64, device=dev)
    info = torch.empty(nb, dtype=torch.int32, device=dev)
    # the per-step exchange: ncclAllGather inside the library (equal shards), torch only for ragged shards
    equal = (B % world == 0)
    if world > 1 and equal:
        init_engine_comm(eng)
    all_buf = torch.empty(B, dtype=torch.float64, device=dev)

    def gather(local):
        if world == 1:
            return local
        if equal:
            return eng.allgather_lnl(local, all_buf)
        return gather_lnl(local, B)

    def step_device():
        eng.log_likelihood_resident(nb, X, A, F, g, n, l, lnL, info)
        return gather(lnL)

    timed = Timer(torch, dist, dev, world, eng)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, launches, all_lnl = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = B / (ms_step * 1e-3)
    bad = int((info != 0).sum().item())
    lnl_main = lnL.clone()
    checksum = float(all_lnl.double().sum().item())

    # ---- parity of THIS run: the bench walkers the CPU oracle evaluated for cpu_baseline ---------------
    parity = None
    if rank == 0 and cpu_sample is not None:
        got = lnl_main.cpu().numpy()
        idx = [b for b in sorted(cpu_sample["lnl"]) if b < nb]
        ref = np.array([cpu_sample["lnl"][b] for b in idx])
        rel = np.abs(got[idx] - ref) / np.maximum(1.0, np.abs(ref))
        parity = {"walkers": len(idx), "max_rel_vs_dense_oracle": float(rel.max()), "tolerance": LNL_RTOL,
                  "ok": bool(rel.max() <= LNL_RTOL), "solver": args.solver,
                  "oracle": "oracle/starfish_oracle.py (numpy kernels + scipy cho_factor/cho_solve) on the same "
                            "seeded stage inputs at N=%d" % N}

    # ---- e2e: host buffers through the C-ABI host entry point -------------------------------------
    e2e = None
    aux_steps = max(10, args.steps)
    if not args.no_e2e:
        def pinned(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return t, t.numpy()

        keep = []
        hb = {}
        loc_pad = np.zeros((nb, eng.K, 3))
        loc_pad[:, :stage["loc"].shape[1]] = stage["loc"]
        for name, arr in (("X", stage["X"]), ("A", stage["A"]), ("F", stage["model_flux"]),
                          ("g", stage["glob"]), ("n", stage["nloc"].astype(np.int32)), ("l", loc_pad),
                          ("lnL", np.zeros(nb)), ("info", np.zeros(nb, dtype=np.int32))):
            t, v = pinned(arr)
            keep.append(t)
            hb[name] = v
        lnl_host_t = keep[-2]

        def step_host():
            eng.log_likelihood_host(hb["X"], hb["A"], hb["F"], hb["g"], hb["n"], hb["l"], hb["lnL"], hb["info"])
            if world > 1:
                return gather(lnl_host_t.to(dev, non_blocking=True))
            return lnl_host_t

        ms_e2e, _, _ = timed(step_host, aux_steps, 2)
        h2d = sum(hb[k].nbytes for k in ("X", "A", "F", "g", "n", "l"))
        d2h = hb["lnL"].nbytes + hb["info"].nbytes
        e2e = {"value": B / (ms_e2e / aux_steps * 1e-3), "unit": "evals/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": aux_steps, "warmup": 2,
               "api": "LikelihoodEngine.log_likelihood_host -> sfb_loglike_host (pinned host buffers)"}
        assert np.array_equal(hb["lnL"], lnl_main.cpu().numpy()), "host and device paths disagree"

    # ---- the parameter-level step (rows f1/f2/f3): theta -> transforms + emulator + covariance path ------
    model_leg = None
    model = None
    if not args.no_model:
        import copy
        import warnings

        from starfish_b200.emulator import Emulator
        from starfish_b200.spectrum import Spectrum
        from starfish_b200.spectrum_model import SpectrumModel

        emu = Emulator(**copy.deepcopy(synth.make_emulator_arrays(n_comp=M)))
        emu._trained = True
        grid0, p0 = synth.walker_params(lo, n_local=K)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model = SpectrumModel(emu, Spectrum(stage["wave"], stage["data_flux"], sigmas=stage["sigma"],
                                                name="synthetic"), grid_params=grid0, device=local_rank,
                                  solver=args.solver, **p0)
        model._engine = eng   # share the workspace of this process' handle
        model._static_sig = None
        labels = list(model.labels)
        P = np.empty((nb, len(labels)))
        for i in range(nb):
            grid_i, p_i = synth.walker_params(lo + i, n_local=K)
            flat = dict(zip(emu.param_names, grid_i))
            flat.update({k: v for k, v in p_i.items() if k not in ("cheb", "global_cov", "local_cov")})
            flat.update({f"cheb:{j + 1}": c for j, c in enumerate(p_i["cheb"])})
            flat.update({f"global_cov:{k}": v for k, v in p_i["global_cov"].items()})
            for j, kern in enumerate(p_i["local_cov"]):
                flat.update({f"local_cov:{j}:{k}": v for k, v in kern.items()})
            P[i] = [flat[k] for k in labels]
        lnl_model = [None]

        def step_model():
            lnl_model[0] = model.log_likelihood_batch(P)
            if world > 1:
                return gather(torch.from_numpy(lnl_model[0]).to(dev))
            return lnl_model[0]

        ms_model, launches_model, _ = timed(step_model, aux_steps, 2)
        nf = len(model.min_dv_wave)
        model_leg = {"value": B / (ms_model / aux_steps * 1e-3), "unit": "evals/s", "steps": aux_steps, "warmup": 2,
                     "h2d_bytes_per_step": int(P.nbytes + nb * (2 + 3 * eng.K) * 8 + nb * 4),
                     "d2h_bytes_per_step": int(nb * 12),
                     "api": "SpectrumModel.log_likelihood_batch(P[B,ndim]) -> sfb_loglike_params_host: emulator GP, "
                            "rotational broadening, Doppler shift, spline resampling, Chebyshev, covariance build, "
                            "Cholesky, solve — all on the device; only parameters and lnL cross PCIe",
                     "n_fine": nf, "ndim": len(labels), "gpu_launches": int(launches_model),
                     "not_finite": int((~np.isfinite(lnl_model[0])).sum())}
        if rank == 0:
            eng.profile(True)
            model.log_likelihood_batch(P)
            up = eng.profile_read()["upstream"]
            eng.profile(False)
            model_leg["upstream_ms_per_step"] = up["ms"]
            model_leg["upstream_gbs"] = up["work"] / (up["ms"] * 1e-3) / 1e9 if up["ms"] > 0 else None

    # ---- roofline of the dominant kernel: one extra profiled step (single stream, events per launch)
    roof = None
    shares = None
    if rank == 0:
        eng.profile(True)
        eng.i8_mma_counts()                      # reset
        eng.log_likelihood_resident(nb, X, A, F, g, n, l, lnL, info)
        prof = eng.profile_read()
        issued, dense_cnt = eng.i8_mma_counts()
        eng.profile(False)
        roof = roofline_block(prof, args.solver, peaks, issued / dense_cnt if dense_cnt else 1.0)
        shares = kernel_shares(prof)

    # ---- the other dense trailing-update mode on the same inputs -------------------------------------
    alt = None
    if not args.no_alt:
        other = "dense" if args.solver == "dense_i8" else "dense_i8"
        eng.set_solver(other)
        ms_a, launches_a, _ = timed(step_device, args.steps, 2)
        v_a = B / (ms_a / args.steps * 1e-3)
        alt = {"solver": other, "note": solver_note(other), "value": v_a, "unit": "evals/s",
               "ms_per_step": ms_a / args.steps, "steps": args.steps, "gpu_launches": int(launches_a),
               "path_tflops_fp64_equivalent": v_a * flops_eval(N, M) / 1e12,
               "frac_of_fp64_floor": v_a * flops_eval(N, M) / 1e12 / (FP64_DMMA_PEAK_TFLOPS * world),
               "max_rel_diff_vs_headline_lnL":
                   float(((lnL - lnl_main).abs() / lnl_main.abs().clamp(min=1.0)).max().item()),
               "not_positive_definite": int((info != 0).sum().item())}
        if rank == 0:
            eng.profile(True)
            eng.i8_mma_counts()
            eng.log_likelihood_resident(nb, X, A, F, g, n, l, lnL, info)
            pa = eng.profile_read()
            issued_a, dense_a = eng.i8_mma_counts()
            eng.profile(False)
            alt["roofline"] = roofline_block(pa, other, peaks, issued_a / dense_a if dense_a else 1.0)
            alt["kernels"] = kernel_shares(pa)
        eng.set_solver(args.solver)

    # ---- frozen kernel groups: every walker shares the hyper-parameter row -> S factorised once per step ------
    frozen = None
    if not args.no_frozen:
        g1, n1, l1 = eng.pack_hyper(nb, stage["glob"][:1], stage["nloc"][:1], stage["loc"][:1], True)

        def step_frozen():
            eng.log_likelihood_resident(nb, X, A, F, g1, n1, l1, lnL, info, shared_hyper=True)
            return gather(lnL)

        f_steps = max(10, 2 * args.steps)
        ms_f, launches_f, _ = timed(step_frozen, f_steps, 2)
        lnl_frozen = lnL.clone()
        eng.set_shared_factor(False)      # the same call, every walker's full covariance factorised: the check
        step_frozen()
        torch.cuda.synchronize(dev)
        diff = float(((lnl_frozen - lnL).abs() / lnL.abs().clamp(min=1.0)).max().item())
        eng.set_shared_factor(True)
        fl_f = N ** 3 / 3 + nb * float(N) ** 2 * (M + 1)   # per step and GPU: one factorisation + the forward solves
        v_f = B / (ms_f / f_steps * 1e-3)
        frozen = {"value": v_f, "unit": "evals/s", "ms_per_step": ms_f / f_steps, "steps": f_steps,
                  "gpu_launches": int(launches_f), "max_rel_diff_vs_per_walker_factorisation": diff,
                  "not_positive_definite": int((info != 0).sum().item()),
                  "flop_model": "N^3/3 (one factorisation of S) + B*N^2*(M+1) (forward solves of [R | X^T]) per step",
                  "path_tflops": fl_f / (ms_f / f_steps * 1e-3) / 1e12,
                  "frac_of_fp64_peak": fl_f / (ms_f / f_steps * 1e-3) / 1e12 / FP64_DMMA_PEAK_TFLOPS,
                  "api": "sfb_loglike(shared_hyper=1): the reference's frozen global_cov/local_cov mode "
                         "(spectrum_model.py:341-363); S built and factorised once, all walkers' right-hand sides "
                         "solved together, M x M capacitance system per walker"}
        if rank == 0:   # rank 0 alone: the local call only — step_frozen() ends in the all-gather and would wait for the others
            eng.profile(True)
            eng.log_likelihood_resident(nb, X, A, F, g1, n1, l1, lnL, info, shared_hyper=True)
            frozen["kernels"] = kernel_shares(eng.profile_read())
            eng.profile(False)

    # ---- the structure-exploiting solver (row f4): same stage boundary, banded Cholesky + capacitance ------
    structured = None
    if not args.no_structured:
        eng.set_solver("structured")
        s_steps = max(10, 4 * args.steps)
        ms_s, launches_s, _ = timed(step_device, s_steps, args.warmup)
        diff = float(((lnL - lnl_main).abs() / lnl_main.abs().clamp(min=1.0)).max().item())
        structured = {"value": B / (ms_s / s_steps * 1e-3), "unit": "evals/s", "steps": s_steps,
                      "ms_per_step": ms_s / s_steps, "gpu_launches": int(launches_s),
                      "max_rel_diff_vs_dense_lnL": diff, "not_positive_definite": int((info != 0).sum().item()),
                      "window_classes": {str(k): v for k, v in eng.band_classes().items()},
                      "api": "sfb_set_solver(SFB_SOLVER_STRUCTURED) + sfb_loglike (inputs resident)"}
        if model is not None:
            model.solver = "structured"
            ms_m, launches_m, _ = timed(step_model, s_steps, 2)
            structured["e2e_model"] = {"value": B / (ms_m / s_steps * 1e-3), "unit": "evals/s",
                                       "ms_per_step": ms_m / s_steps, "gpu_launches": int(launches_m),
                                       "api": "SpectrumModel(solver='structured').log_likelihood_batch(P)"}
            model.solver = args.solver
        if rank == 0:
            eng.profile(True)
            eng.log_likelihood_resident(nb, X, A, F, g, n, l, lnL, info)
            pr = eng.profile_read()
            eng.profile(False)
            bc, bb = pr["band_chol"], pr["band_build"]
            ach = bc["work"] / (bc["ms"] * 1e-3) / 1e12 if bc["ms"] > 0 else 0.0
            structured["roofline"] = {
                "kernel": "band_mma_kernel (rank-4 DMMA) — register-resident sliding-window banded Cholesky + "
                          "forward solves",
                "bound": "serial pivot chain (panel solve + barriers), then fp64 issue", "achieved": ach,
                "peak": FP64_DFMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP64_DFMA_PEAK_TFLOPS,
                "peak_source": "DFMA peak measured by tools/fp64_peak.cu (profiles/r01_fp64_peak.txt)",
                "work": "algorithmic N*(b^2 + 2b(M+1)) FLOP per walker, b = its exact half-bandwidth",
                "launches": bc["launches"], "ms": bc["ms"],
                "band_build": {"launches": bb["launches"], "ms": bb["ms"],
                               "achieved_gbs": bb["work"] / (bb["ms"] * 1e-3) / 1e9 if bb["ms"] > 0 else None}}
        eng.set_solver(args.solver)
    if model is not None:
        model._engine = None
    ws_walkers = eng.workspace_walkers
    eng.close()
    del X, A, F
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, each on its own handle ------------------------------------------
    configs = None
    if not args.no_configs:
        configs = {}
        both = [args.solver, "dense" if args.solver == "dense_i8" else "dense_i8"]
        if world == 1:
            st2 = synth.stage_inputs_direct(4096, 64, n_comp=0, n_local=0)
            configs["configs[1]"] = run_config_leg(
                torch, dist, dev, local_rank, world, rank, st2, 4096, 0, 0, 64, 64,
                "BASELINE.json configs[1]: 1xB200, 64-walker batch, N=4096, global kernel only (M=0, K=0)", both,
                steps=10, warmup=3)
        st5 = synth.stage_inputs_orders(64, first_walker=64 * rank)
        configs["configs[4]"] = run_config_leg(
            torch, dist, dev, local_rank, world, rank, st5, 16384, 6, 16, 64, 64 * world,
            f"BASELINE.json configs[4]: 64 walkers per GPU x {world} GPU(s) (512 on 8), N=16384 = 8 concatenated "
            "orders, 2 local kernels per order (16), M=6", both, steps=2, warmup=1)

    fl = flops_eval(N, M)
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f64" if args.solver == "dense" else
                     "f64 results; trailing update as exact int8 tensor-core products of 48-bit fixed-point operands",
            "data": "synthetic",
            "config": config_dict(args, world),
            "clocks": clocks, "e2e": e2e, "e2e_model": model_leg, "parity": parity,
            "gpu_launches": int(launches),
            "roofline": roof, "kernels": shares,
            "path_tflops_fp64_equivalent": value * fl / 1e12,
            "path_frac_of_fp64_peak": value * fl / 1e12 / (FP64_DMMA_PEAK_TFLOPS * world),
            ("fp64_dmma" if args.solver == "dense_i8" else "int8_tensor"): alt,
            "frozen_shared": frozen, "structured": structured, "configs": configs,
            "not_positive_definite": bad,
            "lnL_checksum": checksum,
            "workspace_walkers": ws_walkers,
        }
    return line


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    # CPU baseline first: its worker pool is forked before this process creates a CUDA context
    cpu_sample = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_sample = cpu_measure(args)
    import torch
    import torch.distributed as dist

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            # NCCL prints its "NCCL version ..." banner to stdout at these two levels; keep stdout to the JSON line
            del os.environ["NCCL_DEBUG"]
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    try:
        line = run_b200(args, rank, world, local_rank, cpu_sample)
        if rank == 0:
            line["cpu_baseline"] = cpu_baseline_block(cpu_sample) if cpu_sample else None
            print(json.dumps(line), flush=True)
    finally:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
