#!/usr/bin/env python
"""Generate the committed golden fixtures in tests/golden/ — run HERE, where /root/reference is mounted.

Everything is produced by the REFERENCE's own code or expressions, never by this repository's oracle:

* zero_one_cases.npz   0/1 operands drawn with the reference's density rule and the ground truth
                       ``torch.matmul(a.cpu().float(), b.cpu().float()).half()`` (zero_one_correctness_check.py:65-92)
* randn_cases.npz      N(0,1) fp16 operands and the same fp32 truth expression (tolerance tests; the
                       reference defines no pass rule here — see oracle/hgemm_oracle.c header)
* helpers.json         outputs of the reference's importable helpers tools/utils.py:
                       extract_bm_bk_bn on synthetic source snippets (ours), as_col_major on small tensors

* bf16_cases.npz       the bf16 variant (this repository's extension; the reference has no bf16 kernel, so there is
                       no reference code to run): bf16 operands as uint16 bit patterns and the truth
                       ``torch.matmul(a.float(), b.float()).bfloat16()`` — the reference's truth expression with the output
                       type swapped. Integer-valued cases are exact (|c| <= 256 fits bf16's 8 significant bits).

    python tests/golden/make_golden.py          # needs /root/reference; the GPU box only reads the outputs
    python tests/golden/make_golden.py --bf16   # only (re)writes bf16_cases.npz (needs torch only)
"""
import sys
import importlib.util
import json
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")


def ref_utils():
    spec = importlib.util.spec_from_file_location("ref_tools_utils", REF / "tools" / "utils.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_truth(a, b):
    return torch.matmul(a.cpu().float(), b.cpu().float()).half()      # zero_one_correctness_check.py:87-90


def draw01(shape, levels, gen):
    values = torch.tensor([0.0] * (levels - 1) + [1.0], dtype=torch.half)   # :65-73: [0,1] or [0,0,1]
    idx = torch.randint(0, len(values), shape, generator=gen)
    return values[idx].contiguous()


ZERO_ONE_CASES = [
    # (m, n, k, levels, seed)        levels: 2 if max(m,n,k) <= 8192 else 3
    (64, 256, 64, 2, 1),             # BASELINE config 1 family (64_4096_64), narrowed
    (128, 64, 64, 2, 2),             # one tile, one k-block
    (200, 328, 72, 2, 3),            # ragged in all three dims (edges handled in-kernel)
    (256, 512, 1024, 2, 4),
    (8, 16, 8192, 2, 5),             # E[c] = 2048: about half the entries exceed 2047 -> mask path
    (8, 8, 12288, 3, 6),             # > 8192 -> {0,0,1} density
    (1, 8, 8, 2, 7),                 # degenerate M
    (384, 264, 136, 2, 8),
]
RANDN_CASES = [(64, 128, 64, 11), (200, 328, 72, 12), (128, 128, 1024, 13)]

SNIPPETS = {
    "plain": "using BM = Int<128>;\nusing BN = Int<256>;\nusing BK = Int<32>;\n",
    "spaces": "  static constexpr auto BM   =   Int< 160 >{};\n auto BN=Int<128>{}; \n auto BK = Int<32>{};",
    "last_wins": "auto BM = Int<64>{};\nauto BM = Int<96>{};\nauto BN = Int<128>{};\nauto BK = Int<16>{};",
    "missing_bk": "auto BM = Int<64>{};\nauto BN = Int<128>{};\n",
    "none": "__global__ void k() {}\n// tile 256 x 128 x 64\n",
    "two_on_a_line": "auto BM = Int<64>{}; auto BN = Int<32>{};\nauto BN = Int<48>{};\nauto BK = Int<8>{};",
    "prefixed_name": "auto kBM = Int<32>{};\nauto BN = Int<64>{};\nauto BK = Int<64>{};",
    "zero": "auto BM = Int<0>{};\nauto BN = Int<64>{};\nauto BK = Int<64>{};",
    "comment": "// BM = Int<512>\nauto BN = Int<64>{};\nauto BK = Int<64>{};",
    "b200_generated": "// tile 256 x 256 x 64 per CTA pair\nB200_HGEMM_SHAPE_ENTRY(true, 256, 6, 2, 8)\n",
}


BF16_CASES = [
    # (m, n, k, kind, seed)   kind: "01" 0/1 operands (sums <= k <= 256: exact in bf16), "int" integers in [-2, 2]
    # with k small enough that |c| <= 256 always, "randn" N(0,1) rounded to bf16 (tolerance tests)
    (64, 256, 64, "01", 21), (200, 328, 72, "01", 22), (128, 64, 256, "01", 23), (384, 264, 136, "01", 24),
    (256, 512, 64, "int", 25), (1, 8, 8, "01", 26),
    (64, 128, 64, "randn", 31), (200, 328, 72, "randn", 32), (128, 128, 1024, "randn", 33),
]


def bits(t):
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


def make_bf16():
    out = {}
    for i, (m, n, k, kind, seed) in enumerate(BF16_CASES):
        gen = torch.Generator().manual_seed(seed)
        if kind == "01":
            a = torch.randint(0, 2, (m, k), generator=gen).bfloat16()
            b = torch.randint(0, 2, (k, n), generator=gen).bfloat16()
        elif kind == "int":
            a = (torch.randint(0, 5, (m, k), generator=gen) - 2).bfloat16()
            b = (torch.randint(0, 5, (k, n), generator=gen) - 2).bfloat16()
        else:
            a = torch.randn((m, k), generator=gen).bfloat16()
            b = torch.randn((k, n), generator=gen).bfloat16()
        truth = torch.matmul(a.float(), b.float()).bfloat16()     # the reference's truth expression, bf16 output
        out[f"a{i}"], out[f"b{i}"], out[f"truth{i}"] = bits(a), bits(b), bits(truth)
        out[f"meta{i}"] = np.array([m, n, k, {"01": 0, "int": 1, "randn": 2}[kind], seed])
    np.savez_compressed(HERE / "bf16_cases.npz", **out)
    print("bf16 fixtures written")


def main():
    if "--bf16" in sys.argv:
        make_bf16()
        return
    utils = ref_utils()
    zo = {}
    for i, (m, n, k, levels, seed) in enumerate(ZERO_ONE_CASES):
        gen = torch.Generator().manual_seed(seed)
        a, b = draw01((m, k), levels, gen), draw01((k, n), levels, gen)
        truth = ref_truth(a, b)
        zo[f"a{i}"] = np.packbits(a.numpy().astype(np.uint8))
        zo[f"b{i}"] = np.packbits(b.numpy().astype(np.uint8))
        zo[f"truth{i}"] = truth.numpy()
        zo[f"meta{i}"] = np.array([m, n, k, levels, seed])
    np.savez_compressed(HERE / "zero_one_cases.npz", **zo)

    rn = {}
    for i, (m, n, k, seed) in enumerate(RANDN_CASES):
        gen = torch.Generator().manual_seed(seed)
        a = torch.randn((m, k), generator=gen).half()
        b = torch.randn((k, n), generator=gen).half()
        rn[f"a{i}"], rn[f"b{i}"] = a.numpy(), b.numpy()
        rn[f"truth{i}"] = ref_truth(a, b).numpy()
        rn[f"meta{i}"] = np.array([m, n, k, seed])
    np.savez_compressed(HERE / "randn_cases.npz", **rn)

    helpers = {"extract_bm_bk_bn": {name: list(utils.extract_bm_bk_bn(text)) for name, text in SNIPPETS.items()},
               "snippets": SNIPPETS, "as_col_major": []}
    for rows, cols, seed in [(3, 5, 0), (8, 8, 1), (16, 24, 2), (1, 7, 3)]:
        x = torch.arange(rows * cols, dtype=torch.float32).reshape(rows, cols).half()
        y = utils.as_col_major(x)
        helpers["as_col_major"].append({"rows": rows, "cols": cols, "flat_out": y.flatten().tolist(),
                                        "shape_out": list(y.shape), "contiguous": bool(y.is_contiguous())})
    (HERE / "helpers.json").write_text(json.dumps(helpers, indent=1))
    make_bf16()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
