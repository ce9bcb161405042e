"""Exactness check of cuda_l2_b200_<acc> on 0/1 matrices against the CPU fp32 truth.

Same CLI as the reference's zero_one_correctness_check.py (:19-25) with --device_type b200, plus:
  --seed N          reproducible inputs (default 0; the reference is unseeded)
  --device cpu      run the whole procedure without a GPU ("plumbing" mode, BASELINE config 1): the kernel slot
                    is filled by a CPU stand-in that reads the same operands, so generator, K-major layout,
                    guard bands, mask and verdict are exercised end to end
  --iterations N    cap on iterations (default 100, as in the reference)
  --attest          also run the five timing-integrity checks of the reference's defense.py on the kernel (no hidden
                    streams or threads, real fp16 output, genuine timers) and store the verdict in the result file
Exit status is 1 when the check fails (the reference always exits 0). Logic: cuda_l2_b200/harness/correctness.py.
"""
import argparse
import sys
import time

import torch

from cuda_l2_b200.harness import correctness as zc
from cuda_l2_b200.harness.common import (LibraryHandles, Padding, add_common_args, baseline_table, kernel_func_name,
                                         load_extension, padding_for, parse_mnk, seed_everything)


def main(argv=None) -> int:
    print("======================Correctness Check======================")
    p = argparse.ArgumentParser()
    add_common_args(p)
    p.add_argument("--device", choices=["cuda", "cpu"], default="cuda")
    p.add_argument("--iterations", type=int, default=100)
    p.add_argument("--attest", action="store_true")
    args = p.parse_args(argv)
    torch.set_grad_enabled(False)
    seed_everything(args.seed)
    m, n, k = parse_mnk(args.mnk)
    under_test = kernel_func_name(args.device_type, args.acc_precise)

    if args.device == "cpu":
        pad = Padding()
        funcs = [torch.matmul, zc.cpu_stand_in(under_test)]
        print(f"Running correctness plumbing on the CPU for m={m}, n={n}, k={k} ...")
        res = zc.run_zero_one_check(kernel_funcs=funcs, kernel_under_test_name=under_test, m=m, n=n, k=k, padding=pad,
                                    device="cpu", num_iterations=args.iterations)
    else:
        torch.cuda.set_device(args.gpu_device_id)
        t0 = time.time()
        hgemm, kernel = load_extension(args)
        print(f"Load hgemm module time: {time.time() - t0:.2f} seconds")
        pad = padding_for(args.mnk, args.acc_precise, args.device_type)
        print(f"Running correctness check for m={m}, n={n}, k={k} ...")
        print(f"Padding: padding_m={pad.m}, padding_k={pad.k}, padding_n={pad.n}")
        with LibraryHandles(hgemm):
            hgemm.find_best_algo_tn_v2_torch(m, n, k)
            hgemm.find_best_algo_nn_v2_torch(m, n, k)
            print("Initialize Done.")
            table = baseline_table(hgemm)
            funcs = [table[nm] for nm in table] + [kernel]
            try:
                res = zc.run_zero_one_check(kernel_funcs=funcs, kernel_under_test_name=under_test, m=m, n=n, k=k,
                                            padding=pad, device="cuda", num_iterations=args.iterations)
            except Exception as e:  # an asynchronous CUDA fault: report and fail
                import traceback
                traceback.print_exc()
                res = zc.CheckResult(False, str(e), {})
    if args.attest and args.device == "cuda" and res.success:
        from cuda_l2_b200.harness.attestation import attest
        from tools.utils import as_col_major
        a = torch.randn((m, k), device="cuda").half()
        b = torch.randn((k, n), device="cuda").half()
        verdict = attest(kernel, a, b, as_col_major(b), torch.empty((m, n), dtype=torch.half, device="cuda"))
        res.result["attestation"] = verdict.to_json()
        print("Attestation:", verdict.to_json())
        if not verdict.passed:
            res = zc.CheckResult(False, "timing-integrity attestation failed", res.result)
    print(res.result)
    zc.write_result(args.base_dir, res)
    print("Correctness Check PASSED:" if res.success else "Correctness Check FAILED:", res.message)
    return 0 if res.success else 1


if __name__ == "__main__":
    sys.exit(main())
