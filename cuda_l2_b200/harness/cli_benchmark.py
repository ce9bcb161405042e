"""Shared main() of benchmarking_offline.py and benchmarking_server.py."""
from __future__ import annotations

import argparse
import gc
import time

import pandas
import torch

from . import benchmark as bm
from .common import (BASELINE_FUNCS, LibraryHandles, add_common_args, baseline_table, kernel_func_name, load_extension,
                     padding_for, parse_mnk, seed_everything)


def build_parser(server: bool) -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description=f"HGEMM benchmark, {'server' if server else 'offline'} mode")
    add_common_args(p)
    p.add_argument("--warmup_seconds", type=float, required=True)
    p.add_argument("--benchmark_seconds", type=float, required=True)
    p.add_argument("--perf_func", type=str, required=True, choices=list(BASELINE_FUNCS),
                   help="the baseline timed against the kernel in this process")
    if server:
        p.add_argument("--target_qps", type=float, required=True)
    return p


def main(server: bool, argv=None) -> int:
    mode = "Server" if server else "Offline"
    print(f"=====================Benchmarking Script -- {mode} Mode======================")
    args = build_parser(server).parse_args(argv)
    torch.set_grad_enabled(False)
    seed_everything(args.seed)
    torch.cuda.set_device(args.gpu_device_id)   # before the first context-creating call

    t0 = time.time()
    hgemm, kernel = load_extension(args)
    print(f"Load hgemm module time: {time.time() - t0:.2f} seconds")
    under_test = kernel_func_name(args.device_type, args.acc_precise)
    m, n, k = parse_mnk(args.mnk)
    print(f"m={m}, n={n}, k={k}, Warmup={args.warmup_seconds}s, Benchmark={args.benchmark_seconds}s")

    start = time.time()
    with LibraryHandles(hgemm):
        if args.perf_func.startswith("hgemm_cublaslt_auto_tuning"):
            print(f"Finding best algo for {args.perf_func}...")
            t1 = time.time()
            finder = hgemm.find_best_algo_tn_v2_torch if args.perf_func.endswith("_tn") else hgemm.find_best_algo_nn_v2_torch
            finder(m, n, k)
            torch.cuda.synchronize()
            print(f"Find best algo time: {time.time() - t1:.2f} seconds")
        baseline = baseline_table(hgemm)[args.perf_func]
        pad = padding_for(args.mnk, args.acc_precise, args.device_type)
        print(f"Using padding_m={pad.m}, padding_k={pad.k}, padding_n={pad.n}")
        print("Warmup...")
        warm, records = bm.timed_loop(
            perf_func_list=[baseline, kernel], m=m, n=n, k=k, acc_precise=args.acc_precise,
            device_type=args.device_type, padding=pad, warmup_seconds=args.warmup_seconds,
            benchmark_seconds=args.benchmark_seconds, target_qps=args.target_qps if server else None)
        print(f"Warmup done: {warm} iterations. Benchmarking done: {len(records)} records.")
    gc.collect()
    torch.cuda.empty_cache()
    print(f"Total time: {time.time() - start:.2f} seconds, {len(records)} records collected.")
    if not records:
        print("no sample fit into --benchmark_seconds")
        return 1

    base_name = baseline.__name__
    cols = [base_name, under_test, base_name + "_ms", under_test + "_ms"]
    df = pandas.DataFrame.from_records(records, columns=["idx"] + cols)
    print(df.head().to_markdown())
    print(df.tail().to_markdown())
    merged = bm.summarise_records(records, [base_name, under_test])
    merged["mode"] = mode.lower()
    merged["seed"] = args.seed
    if server:
        merged["target_qps"] = args.target_qps
    print(merged)
    print(f"speedup over {args.perf_func}: {merged[under_test] / merged[base_name]:.2f}x")
    # key under --perf_func (the reference keys on func.__name__, which is the same string)
    merged[args.perf_func] = merged[base_name]
    bm.write_benchmark_result(args.base_dir, args.perf_func, merged)
    return 0
