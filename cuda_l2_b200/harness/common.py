"""Pieces shared by the harness entry points (zero_one_correctness_check.py, benchmarking_offline.py,
benchmarking_server.py, compile.py, summarize_result.py).

The CLI surface is the reference's (same flag names and meaning: reference eval_one_file.sh:14-59,
benchmarking_offline.py:20-29), with ``b200`` as the device type and two additions every script accepts:
``--seed`` (the reference is unseeded; we default to 0 so runs are reproducible) and, where a GPU is not
needed to make sense of the run, ``--device cpu``.
"""
from __future__ import annotations

import argparse
import math
import os
import random
from dataclasses import dataclass
from pathlib import Path

import numpy as np
import torch

from tools.utils import PROJECT_DIR, acc_dir_name, extract_bm_bk_bn, kernel_source_path

DEVICE_CHOICES = ["b200"]
RESULT_VERSION = "b200-r1"

BASELINE_FUNCS = (
    "hgemm_cublas_tn",
    "hgemm_cublas_nn",
    "hgemm_cublaslt_heuristic_tn",
    "hgemm_cublaslt_heuristic_nn",
    "hgemm_cublaslt_auto_tuning_tn",
    "hgemm_cublaslt_auto_tuning_nn",
    "matmul",
)


def add_common_args(p: argparse.ArgumentParser, *, need_gpu_id: bool = True) -> None:
    p.add_argument("--mnk", type=str, required=True, help="problem size as M_N_K, e.g. 4096_4096_4096")
    p.add_argument("--acc_precise", type=str, required=True, choices=["fp16", "fp32"])
    p.add_argument("--device_type", type=str, required=True, choices=DEVICE_CHOICES)
    p.add_argument("--base_dir", type=str, required=True, help="build + result directory (kept between runs)")
    if need_gpu_id:
        p.add_argument("--gpu_device_id", type=int, required=True)
    p.add_argument("--seed", type=int, default=0)


def parse_mnk(mnk: str) -> tuple[int, int, int]:
    parts = mnk.split("_")
    if len(parts) != 3:
        raise ValueError(f"--mnk must look like M_N_K, got {mnk!r}")
    m, n, k = (int(x) for x in parts)
    if min(m, n, k) <= 0:
        raise ValueError("M, N and K must be positive")
    return m, n, k


def seed_everything(seed: int) -> None:
    random.seed(seed)
    np.random.seed(seed % (2**32))
    torch.manual_seed(seed)


def kernel_func_name(device_type: str, acc_precise: str) -> str:
    acc_dir_name(acc_precise)
    return f"cuda_l2_{device_type}_{acc_precise}"


@dataclass(frozen=True)
class Padding:
    m: int = 0
    k: int = 0
    n: int = 0

    @property
    def any(self) -> bool:
        return bool(self.m or self.k or self.n)


def padding_for(mnk: str, acc_precise: str, device_type: str) -> Padding:
    """Padding the harness must apply for this kernel: derived from the tile sizes its SOURCE TEXT declares
    (reference benchmarking_offline.py:102-113). b200 kernels declare none, so this is all zeros for them."""
    m, n, k = parse_mnk(mnk)
    src = PROJECT_DIR / kernel_source_path(mnk, acc_precise, device_type)
    bm, bk, bn = extract_bm_bk_bn(src.read_text())
    if min(bm, bk, bn) <= 0:
        return Padding()
    up = lambda x, b: math.ceil(x / b) * b - x
    return Padding(m=up(m, bm), k=up(k, bk), n=up(n, bn))


def load_extension(args):
    """JIT-build/import hgemm_lib for the run and return (module, kernel function)."""
    from tools.utils import build_from_sources

    mod = build_from_sources(mnk=args.mnk, acc_precise=args.acc_precise, device_type=args.device_type,
                             base_dir=args.base_dir, verbose=False)
    return mod, getattr(mod, kernel_func_name(args.device_type, args.acc_precise))


def baseline_table(hgemm) -> dict:
    return {
        "hgemm_cublas_tn": hgemm.hgemm_cublas_tn,
        "hgemm_cublas_nn": hgemm.hgemm_cublas_nn,
        "hgemm_cublaslt_heuristic_tn": hgemm.hgemm_cublaslt_heuristic_tn,
        "hgemm_cublaslt_heuristic_nn": hgemm.hgemm_cublaslt_heuristic_nn,
        "hgemm_cublaslt_auto_tuning_tn": hgemm.hgemm_cublaslt_auto_tuning_tn,
        "hgemm_cublaslt_auto_tuning_nn": hgemm.hgemm_cublaslt_auto_tuning_nn,
        "matmul": torch.matmul,
    }


class LibraryHandles:
    """init/destroy of the three comparator handle sets around a run (reference benchmarking_offline.py:66-68,141-143)."""

    def __init__(self, hgemm):
        self.hgemm = hgemm

    def __enter__(self):
        self.hgemm.init_cublas_handle()
        self.hgemm.init_cublaslt_handle_v1()
        self.hgemm.init_cublaslt_handle_v2()
        torch.cuda.synchronize()
        return self

    def __exit__(self, *exc):
        self.hgemm.destroy_cublas_handle()
        self.hgemm.destroy_cublaslt_handle_v1()
        self.hgemm.destroy_cublaslt_handle_v2()
        torch.cuda.synchronize()
        return False


def result_dir(base_dir: str) -> Path:
    p = Path(base_dir)
    p.mkdir(parents=True, exist_ok=True)
    return p


def host_info() -> dict:
    return {"cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads()}
