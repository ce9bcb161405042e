"""Merge {base_dir}/benchmark_result_*.json into the speed-up table of one shape.

Same CLI and naming as the reference's summarize_result.py (:10-13, :26-31, :43-53): per baseline the
speed-up is kernel TFLOP/s / baseline TFLOP/s, and "<name>-max" is the HARDER of the two layouts
(the one with the smaller speed-up). Also writes {base_dir}/summary.json for the multi-GPU farm.
"""
import argparse
import json
import sys
from pathlib import Path

import pandas

from cuda_l2_b200.harness.common import DEVICE_CHOICES, kernel_func_name

ORDER = ["torch.matmul", "cuBLAS-tn", "cuBLAS-nn", "cuBLAS-max", "cuBLASLt-heuristic-tn", "cuBLASLt-heuristic-nn",
         "cuBLASLt-heuristic-max", "cuBLASLt-auto-tuning-tn", "cuBLASLt-auto-tuning-nn", "cuBLASLt-auto-tuning-max"]


def display_name(method: str) -> str:
    if method == "matmul":
        return "torch.matmul"
    return method.replace("hgemm_", "").replace("cublaslt", "cuBLASLt").replace("cublas", "cuBLAS").replace("_", "-")


def summarize(base_dir: str, acc_precise: str, device_type: str) -> dict:
    ours = kernel_func_name(device_type, acc_precise)
    rows = {}
    for f in sorted(Path(base_dir).glob("benchmark_result_*.json")):
        method = f.stem[len("benchmark_result_"):]
        rec = json.loads(f.read_text())["records"]
        rows[display_name(method)] = {
            "Baseline Method Name": display_name(method),
            "Baseline TFLOPS": rec[method],
            "CUDA-L2 TFLOPS": rec[ours],
            "Speedup": rec[ours] / rec[method],
        }
    for fam in ("cuBLAS", "cuBLASLt-heuristic", "cuBLASLt-auto-tuning"):
        pair = [rows[f"{fam}-{s}"] for s in ("tn", "nn") if f"{fam}-{s}" in rows]
        if pair:
            harder = min(pair, key=lambda r: r["Speedup"])
            rows[f"{fam}-max"] = dict(harder, **{"Baseline Method Name": f"{fam}-max"})
    return rows


def main(argv=None) -> int:
    p = argparse.ArgumentParser()
    p.add_argument("--base_dir", type=str, required=True)
    p.add_argument("--acc_precise", type=str, required=True, choices=["fp16", "fp32"])
    p.add_argument("--device_type", type=str, required=True, choices=DEVICE_CHOICES)
    a = p.parse_args(argv)
    rows = summarize(a.base_dir, a.acc_precise, a.device_type)
    if not rows:
        print(f"no benchmark_result_*.json under {a.base_dir}")
        return 1
    df = pandas.DataFrame.from_records([rows[nm] for nm in ORDER if nm in rows])
    print("Summary of Benchmark Results:")
    print(df.to_markdown(floatfmt=".3f", missingval="-"))
    (Path(a.base_dir) / "summary.json").write_text(json.dumps(rows, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
