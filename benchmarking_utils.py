"""Timing primitives under the reference's module name (benchmarking_utils.py:12-69):
``run_benchmark`` (one call between two device synchronisations) and ``run_all_perf_funcs_once``.
Implementation: cuda_l2_b200/harness/benchmark.py."""
from cuda_l2_b200.harness.benchmark import run_all_perf_funcs_once, run_benchmark

__all__ = ["run_benchmark", "run_all_perf_funcs_once"]
