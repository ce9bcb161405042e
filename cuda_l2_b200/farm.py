"""Farm the (M,N,K) sweep over the GPUs of one box — one problem per GPU at a time, no collective on the GEMM path.

The reference evaluates one shape per invocation of eval_one_file.sh and has no multi-GPU notion beyond
``--gpu_device_id`` (benchmarking_offline.py:27,53). Every shape is an independent unit, so the sweep
shards trivially: shapes are sorted longest-first by an analytic cost and dealt round-robin to ranks (LPT order),
each rank pins one GPU and evaluates its share, results are gathered on rank 0 (``all_gather_object`` over
gloo/nccl when launched under torchrun, or through per-rank JSONL files when the orchestrator spawns workers itself).

Engines:
  wall     ``dev_check wall``: the harness metric (wall clock around one call + device sync, mean TFLOP/s) in C++,
           our dispatcher vs cuBLAS / cuBLASLt-heuristic / cuBLASLt-auto-tuning, both layouts. Seconds per shape.
  wallgrid the same, one process per GPU walking its whole share (start-up paid once) — the default for the full grid.
  harness  the reference-style ``eval_one_file.sh`` (JIT build + 1 correctness + 7 benchmark processes). Minutes per shape.
"""
from __future__ import annotations

import csv
import itertools
import json
import os
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
GRID = (64, 128, 256, 512, 1024, 2048, 4096, 8192, 12288, 16384)
EXTRA_SHAPES = ((2048, 11008, 4096),)
CSV_COLUMNS = ["mnk", "torch.matmul", "cuBLAS-tn", "cuBLAS-nn", "cuBLAS-max", "cuBLASLt-heuristic-tn",
               "cuBLASLt-heuristic-nn", "cuBLASLt-heuristic-max", "cuBLASLt-auto-tuning-tn", "cuBLASLt-auto-tuning-nn",
               "cuBLASLt-auto-tuning-max"]     # header of the reference's eval_results/*.csv


def grid_shapes() -> list[tuple[int, int, int]]:
    return list(itertools.product(GRID, GRID, GRID)) + list(EXTRA_SHAPES)


def estimated_cost(shape, tflops: float = 1400.0, gbs: float = 6000.0, fixed_s: float = 2.5) -> float:
    """Seconds a worker spends on a shape: fixed per-shape overhead (tuning, process start) + ~3000 GEMMs."""
    m, n, k = shape
    t = max(2.0 * m * n * k / (tflops * 1e12), 2.0 * (m * k + n * k + m * n) / (gbs * 1e9), 4e-6)
    return fixed_s + 3000 * t


def partition(shapes, world: int) -> list[list[tuple[int, int, int]]]:
    """Longest-processing-time-first onto the least loaded rank: deterministic, every shape exactly once."""
    loads = [0.0] * world
    parts: list[list] = [[] for _ in range(world)]
    for s in sorted(shapes, key=lambda s: (-estimated_cost(s), s)):
        r = min(range(world), key=lambda i: (loads[i], i))
        parts[r].append(s)
        loads[r] += estimated_cost(s)
    return parts


def _flops(mnk: str) -> float:
    m, n, k = (int(x) for x in mnk.split("_"))
    return 2.0 * m * n * k


def speedup_row(mnk: str, tf: dict) -> dict:
    """One row of the reference's CSV schema from absolute TFLOP/s (ours + baselines); "-max" is the harder
    of the two layouts, i.e. the smaller speed-up (summarize_result.py:43-53)."""
    ours = tf["ours"]
    row = {"mnk": mnk, "torch.matmul": (ours / tf["matmul"]) if tf.get("matmul") else ""}
    for fam, key in (("cuBLAS", "cublas"), ("cuBLASLt-heuristic", "lt_heur"), ("cuBLASLt-auto-tuning", "lt_auto")):
        if tf.get(f"{key}_tn") and tf.get(f"{key}_nn"):
            # a pair-protocol record carries the speed-up measured inside each (baseline, ours) pair — the harness's own
            # number; older records only have absolute rates
            tn = tf.get(f"{key}_tn_speedup") or ours / tf[f"{key}_tn"]
            nn = tf.get(f"{key}_nn_speedup") or ours / tf[f"{key}_nn"]
            row[f"{fam}-tn"], row[f"{fam}-nn"], row[f"{fam}-max"] = tn, nn, min(tn, nn)
        else:                                   # a baseline the run did not time (eval_one_file.sh --perf_funcs)
            row[f"{fam}-tn"] = row[f"{fam}-nn"] = row[f"{fam}-max"] = ""
    return row


def parse_wall_line(line: str) -> dict:
    f = line.strip().split(",")
    assert f[0] == "WALL"
    out = {"acc": int(f[1]), "m": int(f[2]), "n": int(f[3]), "k": int(f[4])}
    for tok in f[5:]:
        key, val = tok.split("=", 1)
        try:
            out[key] = float(val)
        except ValueError:
            out[key] = val
    return out


def run_wall_engine(shape, acc_bits: int, seconds: float, tune_rounds: tuple[int, int], gpu: int | None) -> dict:
    exe = REPO / "cuda_l2_b200" / "lib" / "dev_check"
    if not exe.exists():
        raise RuntimeError(f"{exe} missing: run __graft_entry__.build() first (no fallback)")
    env = dict(os.environ)
    if gpu is not None:
        env["CUDA_VISIBLE_DEVICES"] = str(gpu)
    m, n, k = shape
    cmd = [str(exe), "wall", str(acc_bits), str(m), str(n), str(k), str(seconds), str(tune_rounds[0]), str(tune_rounds[1])]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    for line in r.stdout.splitlines():
        if line.startswith("WALL,"):
            return parse_wall_line(line)
    raise RuntimeError(f"dev_check wall failed for {shape}: rc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")


HARNESS_KEYS = {"torch.matmul": "matmul", "cuBLAS-tn": "cublas_tn", "cuBLAS-nn": "cublas_nn",
                "cuBLASLt-heuristic-tn": "lt_heur_tn", "cuBLASLt-heuristic-nn": "lt_heur_nn",
                "cuBLASLt-auto-tuning-tn": "lt_auto_tn", "cuBLASLt-auto-tuning-nn": "lt_auto_nn"}


def record_from_harness_summary(summary: dict) -> dict:
    """{base_dir}/summary.json of one eval_one_file.sh run (written by summarize_result.py) -> a sweep record.
    Every baseline process timed the kernel again; `ours` is their mean, each baseline keeps its own pairing through
    the stored speed-ups, so the "-max" columns are exactly the harness's. Baselines that were not run (eval_one_file.sh
    --perf_funcs) are simply absent; the two cuBLASLt-auto-tuning layouts are required."""
    rec, ours = {}, []
    for name, key in HARNESS_KEYS.items():
        if name not in summary:
            continue
        row = summary[name]
        ours.append(row["CUDA-L2 TFLOPS"])
        rec[key + "_speedup"] = row["Speedup"]
    rec["ours"] = sum(ours) / len(ours)
    for key in HARNESS_KEYS.values():          # baselines re-expressed against the common `ours`
        if key + "_speedup" in rec:
            rec[key] = rec["ours"] / rec[key + "_speedup"]
    rec["speedup_vs_lt_auto_max"] = min(rec["lt_auto_tn_speedup"], rec["lt_auto_nn_speedup"])
    return rec


AUTO_TUNING_PAIR = "hgemm_cublaslt_auto_tuning_tn,hgemm_cublaslt_auto_tuning_nn"


def run_harness_engine(shape, acc_precise: str, warmup_s: float, bench_s: float, gpu: int | None, base_dir: Path,
                       mode: str = "offline", target_qps: float | None = None, perf_funcs: str | None = None,
                       shared_build_dir: bool = True) -> dict:
    """The reference-style flow for one shape: eval_one_file.sh (JIT build, 0/1 check, one fresh process per baseline,
    summary). ``perf_funcs`` restricts the baselines (None = all seven). With ``shared_build_dir`` every shape of a
    worker builds in the same directory, so ninja recompiles only kernels/.../<mnk>.cu (the three library sources and
    the binding are shape-independent); each shape's summary is copied to ``<base_dir>/summaries/<mnk>.json``."""
    m, n, k = shape
    mnk = f"{m}_{n}_{k}"
    env = dict(os.environ)
    if gpu is not None:
        env["CUDA_VISIBLE_DEVICES"] = str(gpu)      # the worker then addresses its GPU as device 0
    base_dir = Path(base_dir)
    out = base_dir / (f"build_{acc_precise}_{gpu if gpu is not None else 0}" if shared_build_dir else mnk)
    cmd = [str(REPO / "eval_one_file.sh"), "--mnk", mnk, "--acc_precise", acc_precise, "--device_type", "b200",
           "--warmup_seconds", str(warmup_s), "--benchmark_seconds", str(bench_s), "--base_dir", str(out),
           "--gpu_device_id", "0", "--mode", mode]
    if mode == "server":
        cmd += ["--target_qps", str(target_qps or 100)]
    if perf_funcs:
        cmd += ["--perf_funcs", perf_funcs]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=3600)
    if r.returncode != 0:
        raise RuntimeError(f"eval_one_file.sh failed for {shape}: rc={r.returncode}\n{r.stdout[-1500:]}\n{r.stderr[-1500:]}")
    summary = json.loads((out / "summary.json").read_text())
    keep = base_dir / "summaries"
    keep.mkdir(parents=True, exist_ok=True)
    (keep / f"{mnk}_{acc_precise}_{mode}.json").write_text(json.dumps(summary, indent=1))
    return record_from_harness_summary(summary)


def run_wallgrid_worker(rank: int, world: int, acc_bits: int, seconds: float, tune_rounds: tuple[int, int],
                        gpu: int | None, out_path: Path, limit: int = 0, shapes_file: str | None = None) -> list[dict]:
    """One process per GPU walks its share of the cost-sorted grid inside ``dev_check wallgrid`` (CUDA/cuBLAS start-up
    and the 20 GB auto-tuning workspace are paid once per GPU instead of once per shape). The share is the
    round-robin deal ``index % world == rank`` of the cost-sorted list — the same rule ``dev_check`` applies."""
    exe = REPO / "cuda_l2_b200" / "lib" / "dev_check"
    if not exe.exists():
        raise RuntimeError(f"{exe} missing: run __graft_entry__.build() first (no fallback)")
    env = dict(os.environ)
    if gpu is not None:
        env["CUDA_VISIBLE_DEVICES"] = str(gpu)
    if shapes_file:
        env["B200_WALLGRID_SHAPES"] = str(shapes_file)      # "M N K" lines replace the grid (partial re-sweeps, samples)
    cmd = [str(exe), "wallgrid", str(acc_bits), str(rank), str(world), str(seconds), str(tune_rounds[0]),
           str(tune_rounds[1]), str(limit)]
    results = []
    with subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) as proc, \
            open(out_path, "a") as out:
        for line in proc.stdout:
            if line.startswith("WALL,"):
                rec = parse_wall_line(line)
                rec.update(mnk=f"{rec['m']}_{rec['n']}_{rec['k']}", rank=rank, ok=True)
            elif line.startswith("WALLFAIL,"):
                f = line.strip().split(",")
                rec = {"mnk": "_".join(f[2:5]), "rank": rank, "ok": False, "error": ",".join(f[5:])}
            else:
                continue
            results.append(rec)
            out.write(json.dumps(rec) + "\n")
            out.flush()
    return results


PY_FUNCS = {"matmul": "matmul", "hgemm_cublas_tn": "cublas_tn", "hgemm_cublas_nn": "cublas_nn",
            "hgemm_cublaslt_heuristic_tn": "lt_heur_tn", "hgemm_cublaslt_heuristic_nn": "lt_heur_nn",
            "hgemm_cublaslt_auto_tuning_tn": "lt_auto_tn", "hgemm_cublaslt_auto_tuning_nn": "lt_auto_nn"}


def ctypes_harness_functions(acc_precise: str):
    """The harness's function table (cuda_l2_b200/harness/common.py baseline_table) with the C-ABI libraries behind the
    names instead of a JIT-built torch extension: ``cuda_l2_b200_<acc>`` -> libb200_hgemm.so, the six library baselines
    -> libb200_baselines.so, ``matmul`` -> torch.matmul itself. Same signatures, same ``__name__``s, so the harness's own
    timing code (harness/benchmark.py) runs on them unchanged."""
    import torch

    from . import capi

    bl = capi.Baselines(acc_precise)

    def named(name, fn):
        fn.__name__ = name
        return fn
    table = {"matmul": torch.matmul}
    for fam, method in (("hgemm_cublas", bl.cublas), ("hgemm_cublaslt_heuristic", bl.lt_heuristic), ("hgemm_cublaslt_auto_tuning", bl.lt_autotune)):
        table[f"{fam}_tn"] = named(f"{fam}_tn", lambda a, b, b_col_major, out, _m=method: _m(capi.Baselines.TN, a, b_col_major, out))
        table[f"{fam}_nn"] = named(f"{fam}_nn", lambda a, b, b_col_major, out, _m=method: _m(capi.Baselines.NN, a, b, out))
    kernel = named(f"cuda_l2_b200_{acc_precise}", lambda a, b, b_col_major, out: capi.hgemm(a, b_col_major, out, acc_precise))
    return table, kernel, bl


def run_pyharness_worker(rank: int, world: int, acc_precise: str, shapes, warmup_s: float, bench_s: float, gpu: int | None,
                         out_path: Path, perf_funcs=("matmul",), mode: str = "offline", target_qps: float | None = None,
                         seed: int = 0) -> list[dict]:
    """The harness's Python timing loop (harness/benchmark.py timed_loop — fresh torch.randn operands per sample, one copy
    per function, zero-filled output, wall clock around one synchronised call, offline or server pacing) over this rank's
    share of ``shapes``, every requested baseline paired with the kernel in turn, all in ONE process: no JIT build and
    one process start per GPU instead of eight per shape. It is what fills the ``torch.matmul`` column (dev_check
    cannot call torch) and the server-mode tables. Call inside a process whose CUDA_VISIBLE_DEVICES selects the GPU."""
    import random

    import numpy as np
    import torch

    from .harness import benchmark as bm
    from .harness.common import Padding

    torch.cuda.set_device(0)
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    table, kernel, bl = ctypes_harness_functions(acc_precise)
    mine = partition(shapes, world)[rank]
    results = []
    # (no global torch.set_grad_enabled(False) here: the timing functions are @torch.no_grad themselves, and a caller
    #  in the same process — a test session — must get its autograd state back)
    with open(out_path, "a") as out, torch.no_grad():
        for (m, n, k) in mine:
            rec = {"mnk": f"{m}_{n}_{k}", "rank": rank, "ok": True, "engine": "pyharness", "mode": mode}
            try:
                ours = []
                for name in perf_funcs:
                    if name.startswith("hgemm_cublaslt_auto_tuning"):
                        bl.lt_autotune_find(bl.TN if name.endswith("_tn") else bl.NN, m, n, k)
                    # long kernels: at least three samples even when they do not fit the time budget
                    est = max(2.0 * m * n * k / 1.2e15, 8e-6)
                    b_s = max(bench_s, 3 * (2 * est + 12 * (m * k + k * n + m * n) * 2 / 4e12))
                    _, records = bm.timed_loop(perf_func_list=[table[name], kernel], m=m, n=n, k=k, acc_precise=acc_precise,
                                               device_type="b200", padding=Padding(), warmup_seconds=warmup_s,
                                               benchmark_seconds=b_s, target_qps=target_qps if mode == "server" else None)
                    merged = bm.summarise_records(records, [table[name].__name__, kernel.__name__])
                    key = PY_FUNCS[name]
                    rec[key] = merged[table[name].__name__]
                    rec[key + "_speedup"] = merged[kernel.__name__] / merged[table[name].__name__]
                    rec[key + "_n"] = merged["samples"]
                    ours.append(merged[kernel.__name__])
                rec["ours"] = sum(ours) / len(ours)
                if "lt_auto_tn_speedup" in rec and "lt_auto_nn_speedup" in rec:
                    rec["speedup_vs_lt_auto_max"] = min(rec["lt_auto_tn_speedup"], rec["lt_auto_nn_speedup"])
            except Exception as e:  # keep the sweep alive
                rec = {"mnk": f"{m}_{n}_{k}", "rank": rank, "ok": False, "error": str(e)[:500]}
            results.append(rec)
            out.write(json.dumps(rec) + "\n")
            out.flush()
    bl.close()
    return results


def run_partition(rank: int, shapes, engine, out_path: Path | None = None, done: set | None = None) -> list[dict]:
    """Evaluate this rank's share with ``engine(shape) -> dict``; append each result to ``out_path`` (JSONL) so an
    interrupted sweep resumes where it stopped. A failing shape is recorded, not fatal (per-shape isolation)."""
    results = []
    for s in shapes:
        key = "_".join(map(str, s))
        if done and key in done:
            continue
        try:
            rec = dict(engine(s), mnk=key, rank=rank, ok=True)
        except Exception as e:  # keep the sweep alive; the failure list is part of the report
            rec = {"mnk": key, "rank": rank, "ok": False, "error": str(e)[:500]}
        results.append(rec)
        if out_path is not None:
            with open(out_path, "a") as f:
                f.write(json.dumps(rec) + "\n")
    return results


def load_done(paths) -> dict[str, dict]:
    done = {}
    for p in paths:
        if Path(p).exists():
            for line in Path(p).read_text().splitlines():
                try:
                    rec = json.loads(line)
                except json.JSONDecodeError:
                    continue
                if rec.get("ok"):
                    done[rec["mnk"]] = rec
    return done


def gather(results: list[dict]) -> list[list[dict]] | None:
    """All ranks' result lists on rank 0 (None elsewhere); a plain list when not running under torch.distributed."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return [results]
    bucket = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(results, bucket, dst=0)
    return bucket


def write_reports(records: list[dict], out_csv: Path, peak_tflops: float, peak_gbs: float) -> dict:
    """The reference-schema speed-up CSV plus an extended CSV with absolute TFLOP/s and roofline fractions."""
    ok = sorted((r for r in records if r.get("ok")), key=lambda r: tuple(int(x) for x in r["mnk"].split("_")))
    out_csv.parent.mkdir(parents=True, exist_ok=True)
    with open(out_csv, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=CSV_COLUMNS)
        w.writeheader()
        for r in ok:
            w.writerow(speedup_row(r["mnk"], r))
    ext = out_csv.with_name(out_csv.stem + "_absolute.csv")
    cols = ["mnk", "ours", "cublas_tn", "cublas_nn", "lt_heur_tn", "lt_heur_nn", "lt_auto_tn", "lt_auto_nn",
            "speedup_vs_lt_auto_max", "roofline_bound", "roofline_frac", "cfg", "gm", "splits"]
    wins = 0
    with open(ext, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        for r in ok:
            m, n, k = (int(x) for x in r["mnk"].split("_"))
            flops, byts = 2.0 * m * n * k, 2.0 * (m * k + n * k + m * n)
            t_tensor, t_hbm = flops / (peak_tflops * 1e12), byts / (peak_gbs * 1e9)
            t_ours = flops / (r["ours"] * 1e12)
            row = {c: r.get(c, "") for c in cols}
            row["roofline_bound"] = "tensor" if t_tensor >= t_hbm else "hbm"
            row["roofline_frac"] = max(t_tensor, t_hbm) / t_ours
            w.writerow(row)
            wins += r["speedup_vs_lt_auto_max"] >= 1.0
    n = len(ok)
    mean = sum(r["speedup_vs_lt_auto_max"] for r in ok) / n if n else float("nan")
    # breakdown by roofline class at the measured peaks: launch-bound (< 1 GFLOP), tensor-bound, HBM-bound
    classes: dict[str, list[float]] = {"launch_bound_lt_1gflop": [], "tensor_bound": [], "hbm_bound": []}
    for r in ok:
        m, nn_, k = (int(x) for x in r["mnk"].split("_"))
        flops, byts = 2.0 * m * nn_ * k, 2.0 * (m * k + nn_ * k + m * nn_)
        key = ("launch_bound_lt_1gflop" if flops < 1e9 else
               "tensor_bound" if flops / (peak_tflops * 1e12) >= byts / (peak_gbs * 1e9) else "hbm_bound")
        classes[key].append(r["speedup_vs_lt_auto_max"])
    by_class = {k: {"shapes": len(v), "won": sum(x >= 1.0 for x in v), "mean_speedup": (sum(v) / len(v)) if v else None}
                for k, v in classes.items()}
    baseline_rows = {r["mnk"]: {"ours_tflops": r["ours"], "speedup_vs_lt_auto_max": r["speedup_vs_lt_auto_max"],
                                "speedup_vs_cublas_max": (r["ours"] / max(r["cublas_tn"], r["cublas_nn"])) if r.get("cublas_tn") and r.get("cublas_nn") else None,
                                "cfg": r.get("cfg"), "gm": r.get("gm"), "splits": r.get("splits")}
                     for r in ok if r["mnk"] in ("64_4096_64", "4096_4096_4096", "8192_8192_8192", "2048_11008_4096")}
    return {"shapes": n, "failed": [r["mnk"] for r in records if not r.get("ok")], "by_class": by_class,
            "baseline_config_shapes": baseline_rows,
            "won_vs_cublas_max": sum(r["ours"] >= max(r["cublas_tn"], r["cublas_nn"]) for r in ok if r.get("cublas_tn") and r.get("cublas_nn")),
            "won_vs_lt_heuristic_max": sum(r["ours"] >= max(r["lt_heur_tn"], r["lt_heur_nn"]) for r in ok if r.get("lt_heur_tn") and r.get("lt_heur_nn")),
            "won_vs_lt_auto_max": wins, "win_fraction": wins / n if n else float("nan"), "mean_speedup_vs_lt_auto_max": mean,
            "aggregate_tflops": (sum(_flops(r["mnk"]) for r in ok) /
                                 sum(_flops(r["mnk"]) / (r["ours"] * 1e12) for r in ok) * 1e-12) if n else 0.0}
