#!/usr/bin/env python
"""A stratified sample of the (M,N,K) grid for runs that are too slow for all 1001 shapes (the real
eval_one_file.sh harness, server mode).

    python tools/sample_shapes.py 48 > shapes.txt            # comma-separated M_N_K, BASELINE shapes always included
    python tools/sample_shapes.py 48 --lines                 # one "M N K" per line (B200_WALLGRID_SHAPES format)

Strata = roofline class at the measured peaks (launch-bound < 1 GFLOP, tensor-bound, HBM-bound) x K band
(<= 512, 1024-2048, >= 4096); the sample takes the same fraction of every stratum (at least one shape), spread evenly
over the stratum's cost-sorted list, so it is deterministic and covers small and large problems of every kind.
"""
import itertools
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
GRID = (64, 128, 256, 512, 1024, 2048, 4096, 8192, 12288, 16384)
ALWAYS = [(64, 4096, 64), (4096, 4096, 4096), (8192, 8192, 8192), (2048, 11008, 4096)]


def peaks():
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return float(d["bf16_tflops"]), float(d["hbm_gbs"])
    return 1590.0, 6650.0


def roofline_class(m, n, k, peak_tf, peak_gbs):
    flops, byts = 2.0 * m * n * k, 2.0 * (m * k + n * k + m * n)
    if flops < 1e9:
        return "launch"
    return "tensor" if flops / (peak_tf * 1e12) >= byts / (peak_gbs * 1e9) else "hbm"


def k_band(k):
    return "k<=512" if k <= 512 else "k1024-2048" if k <= 2048 else "k>=4096"


def sample(count: int):
    peak_tf, peak_gbs = peaks()
    shapes = list(itertools.product(GRID, GRID, GRID)) + [(2048, 11008, 4096)]
    strata = {}
    for s in shapes:
        strata.setdefault((roofline_class(*s, peak_tf, peak_gbs), k_band(s[2])), []).append(s)
    picked = list(ALWAYS)
    frac = max(0.0, (count - len(ALWAYS))) / len(shapes)
    for key in sorted(strata):
        members = sorted(strata[key], key=lambda s: (s[0] * s[1] * s[2], s))
        take = max(1, round(frac * len(members)))
        step = len(members) / take
        for i in range(take):
            s = members[int((i + 0.5) * step)]
            if s not in picked:
                picked.append(s)
    return picked


def main(argv):
    count = int(argv[1]) if len(argv) > 1 and argv[1].isdigit() else 48
    picked = sample(count)
    if "--lines" in argv:
        print("\n".join(f"{m} {n} {k}" for m, n, k in picked))
    else:
        print(",".join(f"{m}_{n}_{k}" for m, n, k in picked))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
