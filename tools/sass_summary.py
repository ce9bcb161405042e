#!/usr/bin/env python
"""SASS evidence that libb200_hgemm.so is a Blackwell-native kernel family (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): counts of tcgen05 / TMEM / TMA mnemonics, and of the legacy tensor-core paths that must be 0.

    python tools/sass_summary.py > profiles/sass_summary.txt      # no GPU needed (cuobjdump reads the cubin)
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
LIB = REPO / "cuda_l2_b200" / "lib" / "libb200_hgemm.so"
WANT = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR", "ACQBULK", "PREEXIT",
        "HMMA", "HGMMA", "LDGSTS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs = re.findall(r"Function : (\S+)", sass)
    ops = collections.Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", sass, re.M):
        ops[m.group(1)] += 1
    print(f"# cuobjdump -sass {LIB.relative_to(REPO)}   ({len(funcs)} kernels, {sum(ops.values())} instructions)")
    print("# kernels by K-mode (last template argument): " +
          ", ".join(f"{k}: {v}" for k, v in sorted(collections.Counter(re.search(r"ELi(\d)EEEv", f).group(1) for f in funcs
                                                                       if re.search(r"ELi(\d)EEEv", f)).items())) +
          "   (0 plain, 1 workspace split-K, 2 cluster split-K, 3 stream-K)")
    print("# operand types: " + ", ".join(f"{k}: {v}" for k, v in sorted(collections.Counter(
        ("bf16" if re.search(r"Lb1EEELi\d", f) else "fp16 in, fp32 acc" if re.search(r"ELi[12]ELb1E", f) else "fp16 in, fp16 acc")
        for f in funcs).items())))
    print()
    for w in WANT:
        match = lambda k: k == w or k.startswith(w + ".") or k.startswith(w + "_")
        exact = sum(v for k, v in ops.items() if match(k))
        variants = sorted(k for k in ops if match(k))
        note = {"HMMA": "   <- legacy mma.sync path: must be 0", "HGMMA": "   <- Hopper wgmma: must be 0",
                "LDGSTS": "   <- cp.async (not used: every bulk load is TMA)"}.get(w, "")
        print(f"{w:14s} {exact:6d}   {' '.join(variants[:8])}{note}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
