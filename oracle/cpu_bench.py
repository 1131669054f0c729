"""Timing of the CPU oracle on the host cores — used only by bench.py (`cpu_baseline` leg and
`--impl reference`).  TEST/BENCH INFRASTRUCTURE, never imported by starfish_b200.

Two threading modes are tried: one walker at a time with all-core BLAS, and P worker
processes × 1 BLAS thread (emcee-pool style).  The numpy element-wise kernel builders of the reference
are single-threaded, so the process-parallel mode is the one that uses all the host threads.
"""
import os
import time

import numpy as np

from oracle import starfish_oracle as O  # imported in the parent so forked workers start warm

_STATE = {}


def _worker_init(stage, blas_threads):
    try:
        from threadpoolctl import threadpool_limits

        _STATE["limit"] = threadpool_limits(limits=blas_threads)
    except Exception:
        pass
    _STATE["stage"] = stage


def _eval_one(b):
    d = _STATE["stage"]
    t0 = time.perf_counter()
    X = d["X"][b] if d["X"] is not None else None
    wcov = np.linalg.inv(d["A"][b]) if X is not None else None  # oracle takes Σ_w; A = Σ_w⁻¹
    loc = d["loc"][b][: d["nloc"][b]]
    glob = d["glob"][b] if d["glob"][b][0] > 0 else None
    lnl = O.stage_log_likelihood(d["wave"], d["sigma"], d["data_flux"], X, wcov, d["model_flux"][b], glob, loc)
    return b, lnl, time.perf_counter() - t0


def time_pool(stage, walkers, procs):
    """Evaluate `walkers` (indices) on `procs` forked workers, 1 BLAS thread each. -> (seconds, lnL dict)"""
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    with ctx.Pool(processes=procs, initializer=_worker_init, initargs=(stage, 1)) as pool:
        pool.map(_noop, range(procs))  # spin the workers up outside the timed region
        t0 = time.perf_counter()
        res = pool.map(_eval_one, list(walkers), chunksize=1)
        dt = time.perf_counter() - t0
    return dt, {b: l for b, l, _ in res}


def _noop(_):
    return 0


def time_serial(stage, walkers):
    """One walker at a time, BLAS free to use every core. -> (seconds, lnL dict)"""
    _STATE["stage"] = stage
    t0 = time.perf_counter()
    res = [_eval_one(b) for b in walkers]
    return time.perf_counter() - t0, {b: l for b, l, _ in res}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1
