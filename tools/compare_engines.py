#!/usr/bin/env python
"""Agreement between the C++ restatement of the harness protocol (`dev_check wallgrid`, every shape) and the real
reference-style harness (`eval_one_file.sh`, a stratified sample): per roofline class, how close the two speed-ups over
cuBLASLt-auto-tuning-max are, whether they agree on win/lose, and the harness sample's win fraction with a Wilson 95 %
interval.

    python tools/compare_engines.py <wallgrid worker_*.jsonl dir> <harness worker_*.jsonl dir> [out.md]
"""
import json
import math
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
from sample_shapes import peaks, roofline_class  # noqa: E402


def load(dirpath):
    recs = {}
    for f in sorted(Path(dirpath).glob("worker_*.jsonl")):
        for line in f.read_text().splitlines():
            try:
                r = json.loads(line)
            except json.JSONDecodeError:
                continue
            if r.get("ok"):
                recs[r["mnk"]] = r
    return recs


def wilson(wins: int, n: int, z: float = 1.96):
    if n == 0:
        return float("nan"), float("nan")
    p = wins / n
    d = 1 + z * z / n
    c = (p + z * z / (2 * n)) / d
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / d
    return c - h, c + h


def main(argv):
    if len(argv) < 3:
        print(__doc__)
        return 2
    proxy, harness = load(argv[1]), load(argv[2])
    peak_tf, peak_gbs = peaks()
    rows = {}
    for mnk, h in harness.items():
        if mnk not in proxy:
            continue
        m, n, k = (int(x) for x in mnk.split("_"))
        cls = roofline_class(m, n, k, peak_tf, peak_gbs)
        rows.setdefault(cls, []).append((mnk, proxy[mnk]["speedup_vs_lt_auto_max"], h["speedup_vs_lt_auto_max"]))
    out = ["# wallgrid (C++ restatement of the harness protocol) vs the real eval_one_file.sh harness", "",
           "Speed-up over cuBLASLt-auto-tuning-max of the same shapes, measured both ways.", "",
           "| class | shapes | mean speed-up (harness) | mean speed-up (wallgrid) | mean |diff| | same win/lose verdict | harness wins | Wilson 95 % |",
           "|---|---|---|---|---|---|---|---|"]
    allrows = []
    for cls in ("launch", "hbm", "tensor"):
        r = rows.get(cls, [])
        allrows += r
        if not r:
            continue
        wins = sum(h >= 1.0 for _, _, h in r)
        lo, hi = wilson(wins, len(r))
        out.append(f"| {cls}-bound | {len(r)} | {sum(h for _, _, h in r) / len(r):.3f} | {sum(p for _, p, _ in r) / len(r):.3f} | "
                   f"{sum(abs(p - h) for _, p, h in r) / len(r):.3f} | {sum((p >= 1.0) == (h >= 1.0) for _, p, h in r)}/{len(r)} | "
                   f"{wins}/{len(r)} | {lo:.2f}-{hi:.2f} |")
    if allrows:
        wins = sum(h >= 1.0 for _, _, h in allrows)
        lo, hi = wilson(wins, len(allrows))
        out.append(f"| all | {len(allrows)} | {sum(h for _, _, h in allrows) / len(allrows):.3f} | {sum(p for _, p, _ in allrows) / len(allrows):.3f} | "
                   f"{sum(abs(p - h) for _, p, h in allrows) / len(allrows):.3f} | {sum((p >= 1.0) == (h >= 1.0) for _, p, h in allrows)}/{len(allrows)} | "
                   f"{wins}/{len(allrows)} | {lo:.2f}-{hi:.2f} |")
    out += ["", "| shape | class | wallgrid | harness |", "|---|---|---|---|"]
    for cls in ("launch", "hbm", "tensor"):
        for mnk, p, h in sorted(rows.get(cls, []), key=lambda t: tuple(int(x) for x in t[0].split("_"))):
            out.append(f"| {mnk} | {cls} | {p:.3f} | {h:.3f} |")
    text = "\n".join(out) + "\n"
    if len(argv) > 3:
        Path(argv[3]).write_text(text)
    print(text)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
