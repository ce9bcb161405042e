"""Build hgemm_lib for one shape without running anything (pre-warms the ninja cache; needs no GPU).
Same flags as the reference's compile.py (:13-17) with --device_type b200."""
import argparse
import time

import torch

from cuda_l2_b200.harness.common import DEVICE_CHOICES
from tools.utils import build_from_sources

if __name__ == "__main__":
    torch.set_grad_enabled(False)
    p = argparse.ArgumentParser()
    p.add_argument("--base_dir", type=str, required=True)
    p.add_argument("--mnk", type=str, required=True)
    p.add_argument("--acc_precise", type=str, required=True, choices=["fp16", "fp32"])
    p.add_argument("--device_type", type=str, required=True, choices=DEVICE_CHOICES)
    a = p.parse_args()
    t0 = time.time()
    build_from_sources(mnk=a.mnk, acc_precise=a.acc_precise, device_type=a.device_type, base_dir=a.base_dir, verbose=False)
    print(f"Compile hgemm module time: {time.time() - t0:.2f} seconds")
