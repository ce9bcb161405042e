"""Offline-mode HGEMM benchmark: back-to-back timed calls of one baseline and the b200 kernel.

Same CLI as the reference's benchmarking_offline.py (:20-29) with --device_type b200:
    python benchmarking_offline.py --mnk 4096_4096_4096 --acc_precise fp32 --device_type b200 \
        --warmup_seconds 5 --benchmark_seconds 10 --base_dir ./results --gpu_device_id 0 \
        --perf_func hgemm_cublaslt_auto_tuning_tn
Writes {base_dir}/benchmark_result_{perf_func}.json. Logic: cuda_l2_b200/harness/cli_benchmark.py.
"""
from cuda_l2_b200.harness.cli_benchmark import main

if __name__ == "__main__":
    raise SystemExit(main(server=False))
