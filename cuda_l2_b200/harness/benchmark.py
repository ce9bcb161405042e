"""Timing loops of the harness (offline and server mode).

The metric is the reference's: one timed sample = host wall clock around ONE call bracketed by device
synchronisation (reference benchmarking_utils.py:12-33); TFLOP/s = 2*M*N*K / t with the UNPADDED dims
(:66); operands are fresh N(0,1) fp16 draws per sample, prepared outside the timed region (:35-58);
the score of a run is the mean of per-sample TFLOP/s (benchmarking_offline.py:156-161).  Offline mode runs
samples back to back; server mode sleeps Exp(1/target_qps) between samples (benchmarking_server.py:127-128,
144-145).  On top of the reference's mean we also record p50/p99 latency and CUDA-event device time.
"""
from __future__ import annotations

import json
import os
import random
import time

import numpy as np
import torch

from tools.utils import as_col_major

from .common import RESULT_VERSION, Padding, result_dir


@torch.no_grad()
def run_benchmark(*, perf_func, a, b, b_col_major, out):
    """One timed call; returns (out, elapsed milliseconds of host wall clock)."""
    is_matmul = perf_func.__name__ == "matmul"
    out.fill_(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if is_matmul:
        perf_func(a, b, out=out)
    else:
        perf_func(a, b, b_col_major, out)
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) * 1e3


@torch.no_grad()
def run_all_perf_funcs_once(*, perf_func_list, m, n, k, acc_precise, device_type, padding_m, padding_k, padding_n):
    """Draw one (A, B), give every function its own operand copies (padded for the kernel under test when its
    source asks for padding), time each once. Returns {name: TFLOP/s, name_ms: milliseconds}."""
    a = torch.randn((m, k), dtype=torch.half, device="cuda")
    b = torch.randn((k, n), dtype=torch.half, device="cuda")
    under_test = f"cuda_l2_{device_type}_{acc_precise}"
    prepared = []
    for func in perf_func_list:
        if func.__name__ == under_test and (padding_m or padding_k or padding_n):
            a_use = torch.zeros((m + padding_m, k + padding_k), dtype=torch.half, device="cuda")
            b_use = torch.zeros((k + padding_k, n + padding_n), dtype=torch.half, device="cuda")
            a_use[:m, :k] = a
            b_use[:k, :n] = b
            c_use = torch.randn((m + padding_m, n + padding_n), dtype=torch.half, device="cuda")
        else:
            a_use, b_use = a.clone(), b.clone()
            c_use = torch.randn((m, n), dtype=torch.half, device="cuda")
        prepared.append((func, a_use, b_use, as_col_major(b_use), c_use))
    torch.cuda.synchronize()
    flops = 2.0 * m * n * k
    record = {}
    for func, a_use, b_use, bt_use, c_use in prepared:
        _, ms = run_benchmark(perf_func=func, a=a_use, b=b_use, b_col_major=bt_use, out=c_use)
        record[func.__name__] = flops * 1e-12 * 1e3 / ms
        record[func.__name__ + "_ms"] = ms
    return record


def timed_loop(*, perf_func_list, m, n, k, acc_precise, device_type, padding: Padding, warmup_seconds: float,
               benchmark_seconds: float, target_qps: float | None = None):
    """Warm up for ``warmup_seconds`` then sample for ``benchmark_seconds``; the two functions are called in a
    random order each sample. ``target_qps`` switches to server mode."""
    kw = dict(m=m, n=n, k=k, acc_precise=acc_precise, device_type=device_type,
              padding_m=padding.m, padding_k=padding.k, padding_n=padding.n)
    order = list(perf_func_list)

    def pause():
        if target_qps:
            time.sleep(np.random.exponential(1.0 / target_qps))

    t0, warm = time.time(), 0
    while time.time() - t0 < warmup_seconds:
        run_all_perf_funcs_once(perf_func_list=order, **kw)
        warm += 1
        pause()
    records = []
    t0 = time.time()
    while time.time() - t0 < benchmark_seconds:
        random.shuffle(order)
        rec = run_all_perf_funcs_once(perf_func_list=order, **kw)
        rec["idx"] = len(records)
        records.append(rec)
        pause()
    return warm, records


def summarise_records(records, func_names):
    """Mean TFLOP/s per function (the reference's score) plus latency percentiles."""
    out = {}
    for nm in func_names:
        tf = np.array([r[nm] for r in records], dtype=np.float64)
        ms = np.array([r[nm + "_ms"] for r in records], dtype=np.float64)
        out[nm] = float(tf.mean()) if len(tf) else float("nan")
        out[nm + "_ms_mean"] = float(ms.mean()) if len(ms) else float("nan")
        out[nm + "_ms_p50"] = float(np.percentile(ms, 50)) if len(ms) else float("nan")
        out[nm + "_ms_p99"] = float(np.percentile(ms, 99)) if len(ms) else float("nan")
    out["samples"] = len(records)
    out["version"] = RESULT_VERSION
    return out


def write_benchmark_result(base_dir, perf_func_name, merged) -> str:
    path = os.path.join(str(result_dir(base_dir)), f"benchmark_result_{perf_func_name}.json")
    with open(path, "w") as f:
        json.dump({"records": merged}, f)
    return path
