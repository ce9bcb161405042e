"""Server-mode HGEMM benchmark: as benchmarking_offline.py, but with an exponentially distributed pause
(mean 1/--target_qps seconds) between samples, so every call starts from an idle, clocked-down GPU
(reference benchmarking_server.py:127-128,144-145). Also reports p50/p99 latency.
"""
from cuda_l2_b200.harness.cli_benchmark import main

if __name__ == "__main__":
    raise SystemExit(main(server=True))
