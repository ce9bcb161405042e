"""Pre-build hgemm_lib for one (M,N,K) — same flags as the reference's compile.py (:13-17), with --device_type b200.
Logic: cuda_l2_b200/harness/cli_compile.py."""
from cuda_l2_b200.harness.cli_compile import main

if __name__ == "__main__":
    raise SystemExit(main())
