"""`compile.py`: build hgemm_lib for one shape without running anything (pre-warms the ninja cache; needs no GPU)."""
from __future__ import annotations

import argparse
import time

from .common import DEVICE_CHOICES, parse_mnk


def main(argv=None) -> int:
    import torch

    from tools.utils import build_from_sources, kernel_source_path

    ap = argparse.ArgumentParser(description=__doc__)
    for flag, extra in (("--base_dir", {}), ("--mnk", {}), ("--acc_precise", {"choices": ["fp16", "fp32"]}),
                        ("--device_type", {"choices": DEVICE_CHOICES})):
        ap.add_argument(flag, type=str, required=True, **extra)
    ns = ap.parse_args(argv)
    parse_mnk(ns.mnk)
    torch.set_grad_enabled(False)
    began = time.time()
    build_from_sources(mnk=ns.mnk, acc_precise=ns.acc_precise, device_type=ns.device_type, base_dir=ns.base_dir, verbose=False)
    print(f"Compile hgemm module time: {time.time() - began:.2f} seconds "
          f"({kernel_source_path(ns.mnk, ns.acc_precise, ns.device_type)})")
    return 0
