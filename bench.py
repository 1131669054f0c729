#!/usr/bin/env python
"""bench.py — log-likelihood evaluations per second for a walker ensemble (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (covariance build → Cholesky → solve → lnL) over the whole ensemble:
256 walkers at N=8192 pixels, M=6 eigenspectra, global + 2 local kernels (BASELINE.json configs[2]; the
8-GPU line is configs[3]).  The ensemble is fixed, so multi-GPU runs are STRONG scaling: rank r owns walkers
[r·B/G, (r+1)·B/G) and the only exchange is one NCCL all-gather of the lnL scalars per step.

Prints ONE JSON line (rank 0).  `value` = walkers / step time with inputs resident in HBM; `e2e` = the same
through the host-buffer C-ABI call (H2D of X/A/model_flux/hyper-parameters and D2H of lnL/info inside the
timed region); `roofline` = the dominant kernel (syrk trailing update, DMMA fp64) against the measured fp64
tensor peak; `cpu_baseline` = the CPU oracle on this box's host cores on a bounded sample.
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
# fp64 tensor (DMMA m8n8k4) peak measured on this pool's B200 with tools/fp64_peak.cu
# (profiles/r01_fp64_peak.txt): MEASURED_PEAKS.json carries no fp64 entry.
FP64_DMMA_PEAK_TFLOPS = 37.1
FP64_DFMA_PEAK_TFLOPS = 33.8   # plain DFMA issue peak from the same measurement (the band kernels use DFMA)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-pix", type=int, default=WORKLOAD["n_pix"])
    ap.add_argument("--walkers", type=int, default=WORKLOAD["n_walkers"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-model", action="store_true", help="skip the parameter-level (drop-in model) legs")
    ap.add_argument("--no-structured", action="store_true", help="skip the structured-solver (row f4) legs")
    ap.add_argument("--cpu-sample", type=int, default=0, help="walkers in the CPU sample (0 = auto)")
    ap.add_argument("--solver", default="dense", choices=["dense", "dense_i8"],
                    help="trailing update of the dense factorisation: fp64 DMMA or the int8 tensor-core restatement")
    return ap.parse_args()


def config_dict(args, world):
    B = args.walkers
    return {
        "workload": f"{B}-walker ensemble, N={args.n_pix} px, M={WORKLOAD['n_comp']} eigenspectra, global "
                    f"Matern + {WORKLOAD['n_local']} local kernels, dense fp64 Cholesky "
                    f"(BASELINE.json configs[{2 if world == 1 else 3}])",
        "n_walkers": B, "n_pix": args.n_pix, "n_comp": WORKLOAD["n_comp"], "n_local": WORKLOAD["n_local"],
        "walkers_per_gpu": B // world if B % world == 0 else f"{B // world}-{B // world + 1}",
        "parallelism": f"walker-sharded x{world}, one all-gather of lnL per step",
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
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# reference arm / CPU baseline (the only places that touch oracle/)
# ---------------------------------------------------------------------------------------------------
def cpu_measure(args, sample_walkers=None):
    from oracle import cpu_bench
    from starfish_b200 import synth

    cores = cpu_bench.host_cores()
    n = sample_walkers or args.cpu_sample or max(1, min(cores, 16))
    stage = synth.stage_inputs_direct(args.n_pix, n, n_comp=WORKLOAD["n_comp"], n_local=WORKLOAD["n_local"])
    procs = min(cores, n)
    dt, lnl = cpu_bench.time_pool(stage, range(n), procs)
    return dict(value=n / dt, seconds=dt, cores=procs, n=n, lnl=lnl,
                sample=f"{n} of {args.walkers} walkers at N={args.n_pix}, {procs} worker processes x 1 BLAS "
                       f"thread (numpy/scipy oracle port of the reference path), {dt:.1f} s")


def run_reference(args, rank, world):
    if rank != 0:
        return
    t_all = time.perf_counter()
    vals = []
    for i in range(args.warmup + args.steps):
        # every step is a bounded sample of the ensemble; warm-up steps use a 1-walker sample
        m = cpu_measure(args, sample_walkers=1 if i < args.warmup else None)
        if i >= args.warmup:
            vals.append(m)
    best = sorted(vals, key=lambda m: m["seconds"])[len(vals) // 2]
    value = float(np.mean([m["value"] for m in vals]))
    ms_per_step = float(np.mean([m["seconds"] for m in vals])) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, max(world, 1)),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": best["cores"], "kind": "port",
                         "sample": best["sample"] + " per step"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from starfish_b200 import synth
    from starfish_b200.dist import gather_lnl, shard_range
    from starfish_b200.engine import LikelihoodEngine

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, N, M, K = args.walkers, args.n_pix, WORKLOAD["n_comp"], WORKLOAD["n_local"]
    lo, hi = shard_range(B, rank, world)
    nb = hi - lo

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

    def step_device():
        eng.log_likelihood_resident(nb, X, A, F, g, n, l, lnL, info)
        return gather_lnl(lnL, B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = eng.launch_count
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), eng.launch_count - launches0, out

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, launches, all_lnl = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = B / (ms_step * 1e-3)
    bad = int((info != 0).sum().item())

    # ---- e2e: host buffers through the C-ABI host entry point -------------------------------------
    e2e = None
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
                return gather_lnl(lnl_host_t.to(dev, non_blocking=True), B)
            return lnl_host_t

        e2e_steps = max(2, min(args.steps, 3))
        ms_e2e, _, _ = timed(step_host, e2e_steps, 1)
        h2d = sum(hb[k].nbytes for k in ("X", "A", "F", "g", "n", "l"))
        d2h = hb["lnL"].nbytes + hb["info"].nbytes
        e2e = {"value": B / (ms_e2e / e2e_steps * 1e-3), "unit": "evals/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
               "api": "LikelihoodEngine.log_likelihood_host -> sfb_loglike_host (pinned host buffers)"}
        assert np.allclose(hb["lnL"], lnL.cpu().numpy(), rtol=0, atol=0), "host and device paths disagree"

    # ---- the parameter-level step (rows f1/f2/f3): theta -> transforms + emulator + covariance path ------
    model_leg = None
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
                                                name="synthetic"), grid_params=grid0, device=local_rank, **p0)
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
                return gather_lnl(torch.from_numpy(lnl_model[0]).to(dev), B)
            return lnl_model[0]

        m_steps = max(2, min(args.steps, 3))
        ms_model, launches_model, _ = timed(step_model, m_steps, 1)
        nf = len(model.min_dv_wave)
        model_leg = {"value": B / (ms_model / m_steps * 1e-3), "unit": "evals/s", "steps": m_steps,
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
    prof = None
    if rank == 0:
        eng.profile(True)
        eng.log_likelihood_resident(nb, X, A, F, g, n, l, lnL, info)
        prof = eng.profile_read()
        eng.profile(False)
        sy = prof["syrk"]
        total_ms = sum(v["ms"] for v in prof.values())
        ach = sy["work"] / (sy["ms"] * 1e-3) / 1e12 if sy["ms"] > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("syrk_dram_bytes_per_launch")
            except Exception:
                traffic = None
        roof = {"kernel": "syrk_kernel (trailing update A_ij -= L_ik L_jk^T, DMMA m8n8k4 fp64)",
                "bound": "tensor", "achieved": ach, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": ach / FP64_DMMA_PEAK_TFLOPS, "traffic": traffic,
                "peak_source": "fp64 DMMA peak measured on this pool's B200 by tools/fp64_peak.cu "
                               "(MEASURED_PEAKS.json has no fp64 entry; nominal 37.2 TFLOP/s)",
                "launches": sy["launches"], "avg_launch_ms": sy["ms"] / max(sy["launches"], 1),
                "flops_per_launch": sy["work"] / max(sy["launches"], 1),
                "share_of_step": sy["ms"] / total_ms if total_ms else None,
                "timing": "one extra profiled step after the timed region: single stream, CUDA events around "
                          "every launch",
                "other_kernels": {k: {"launches": v["launches"], "ms": v["ms"],
                                      "achieved": (v["work"] / (v["ms"] * 1e-3) / (1e9 if k == "build" else 1e12))
                                      if v["ms"] > 0 else None,
                                      "unit": "GB/s" if k == "build" else "TFLOP/s"}
                                  for k, v in prof.items() if k in ("build", "potrf_diag", "trsm")}}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            roof["other_kernels"]["build"]["peak_hbm_gbs"] = peaks.get("hbm_gbs")
        except Exception:
            pass

    # ---- the structure-exploiting solver (row f4): same stage boundary, banded Cholesky + capacitance ------
    structured = None
    if not args.no_structured:
        lnl_dense = lnL.clone()
        eng.set_solver("structured")
        s_steps = max(10, 4 * args.steps)
        ms_s, launches_s, _ = timed(step_device, s_steps, args.warmup)
        diff = float(((lnL - lnl_dense).abs() / lnl_dense.abs().clamp(min=1.0)).max().item())
        structured = {"value": B / (ms_s / s_steps * 1e-3), "unit": "evals/s", "steps": s_steps,
                      "ms_per_step": ms_s / s_steps, "gpu_launches": int(launches_s),
                      "max_rel_diff_vs_dense_lnL": diff, "not_positive_definite": int((info != 0).sum().item()),
                      "window_classes": {str(k): v for k, v in eng.band_classes().items()},
                      "api": "sfb_set_solver(SFB_SOLVER_STRUCTURED) + sfb_loglike (inputs resident)"}
        if not args.no_model:
            model.solver = "structured"
            ms_m, launches_m, _ = timed(step_model, s_steps, 2)
            structured["e2e_model"] = {"value": B / (ms_m / s_steps * 1e-3), "unit": "evals/s",
                                       "ms_per_step": ms_m / s_steps, "gpu_launches": int(launches_m),
                                       "api": "SpectrumModel(solver='structured').log_likelihood_batch(P)"}
            model.solver = "dense"
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
                "bound": "serial pivot chain (panel solve + barriers), then fp64 issue", "achieved": ach, "peak": FP64_DFMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": ach / FP64_DFMA_PEAK_TFLOPS,
                "peak_source": "DFMA peak measured by tools/fp64_peak.cu (profiles/r01_fp64_peak.txt)",
                "work": "algorithmic N*(b^2 + 2b(M+1)) FLOP per walker, b = its exact half-bandwidth",
                "launches": bc["launches"], "ms": bc["ms"],
                "band_build": {"launches": bb["launches"], "ms": bb["ms"],
                               "achieved_gbs": bb["work"] / (bb["ms"] * 1e-3) / 1e9 if bb["ms"] > 0 else None}}
        eng.set_solver(args.solver)
    if not args.no_model:
        model._engine = None

    # whole-path fp64 FLOPs per evaluation: N^3/3 + 2MN^2 + 2N^2
    flops_eval = N ** 3 / 3 + 2 * M * N ** 2 + 2 * N ** 2
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, world),
            "clocks": clocks, "e2e": e2e, "e2e_model": model_leg, "structured": structured,
            "gpu_launches": int(launches),
            "roofline": roof,
            "path_tflops": value * flops_eval / 1e12,
            "path_frac_of_fp64_peak": value * flops_eval / 1e12 / (FP64_DMMA_PEAK_TFLOPS * world),
            "not_positive_definite": bad,
            "lnL_checksum": float(all_lnl.double().sum().item()),
            "workspace_walkers": eng.workspace_walkers,
        }
    eng.close()
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
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        m = cpu_measure(args)
        cpu_base = {"value": m["value"], "unit": "evals/s", "cores": m["cores"], "kind": "port",
                    "sample": m["sample"]}
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
        line = run_b200(args, rank, world, local_rank)
        if rank == 0:
            line["cpu_baseline"] = cpu_base
        if rank == 0:
            print(json.dumps(line), flush=True)
    finally:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
