#!/usr/bin/env python
"""Let harness-protocol sweeps overrule the rotation-based tuner where they disagree.

    python tools/select_from_sweeps.py <acc 32|16> <threshold> <incumbent.jsonl> <challenger.jsonl> [...] [--out merged.jsonl]

Every record file holds `dev_check wallgrid` results (one JSON per shape, with the configuration / group_m / splits that
ran and `speedup_vs_lt_auto_max`, the harness's own score). The first file is the incumbent: a full sweep with the
tuner's table. A later file (a partial re-sweep with another table, or a trial with one forced configuration) replaces the
incumbent's choice for a shape only when its score is higher by more than `threshold` (0.02 = two points, about the
run-to-run scatter of one shape) — so the table moves on direct evidence in the target metric, not on noise. Writes the
selected (cfg, group_m, splits) into cuda_l2_b200/csrc/hgemm_tuned_table.inc for that accumulator and, with --out, the
per-shape records of the selected runs (each shape's number is the measurement of the entry that is in the table).
Being a selection among noisy measurements it flatters the selected shapes slightly; a confirmation sweep with the final
table is the number to quote.
"""
import json
import re
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
TABLE = REPO / "cuda_l2_b200" / "csrc" / "hgemm_tuned_table.inc"
ROW = re.compile(r"\{(\d+), (\d+), (\d+), (-?\d+), (-?\d+), (-?\d+), (-?\d+), (-?\d+), (-?\d+)\}")


def load(path):
    recs = {}
    for line in Path(path).read_text().splitlines():
        try:
            r = json.loads(line)
        except json.JSONDecodeError:
            continue
        if r.get("ok", True) and "speedup_vs_lt_auto_max" in r:
            recs[r["mnk"]] = r
    return recs


def main(argv):
    out = None
    if "--out" in argv:
        i = argv.index("--out")
        out = Path(argv[i + 1])
        argv = argv[:i] + argv[i + 2:]
    if len(argv) < 5:
        print(__doc__)
        return 2
    acc, threshold = int(argv[1]), float(argv[2])
    best = {k: dict(r, source=Path(argv[3]).name) for k, r in load(argv[3]).items()}
    switched = 0
    for path in argv[4:]:
        for k, r in load(path).items():
            if k in best and r["speedup_vs_lt_auto_max"] > best[k]["speedup_vs_lt_auto_max"] + threshold:
                best[k] = dict(r, source=Path(path).name)
                switched += 1
    col = 3 if acc == 32 else 6
    text, changed = TABLE.read_text(), 0

    def repl(m):
        nonlocal changed
        v = [int(x) for x in m.groups()]
        key = f"{v[0]}_{v[1]}_{v[2]}"
        if key in best:
            b = best[key]
            want = [int(b["cfg"]), int(b["gm"]), int(b["splits"])]
            if v[col:col + 3] != want:
                v[col:col + 3] = want
                changed += 1
        return "{" + ", ".join(str(x) for x in v) + "}"
    text = ROW.sub(repl, text)
    note = ("// fp%d-accumulate entries overruled by harness-protocol sweeps where those beat the tuner's choice by > %.0f %% "
            "(tools/select_from_sweeps.py: %s)\n" % (acc, threshold * 100, ", ".join(Path(p).name for p in argv[3:])))
    if note not in text:
        text = text.replace("static const TunedEntry kTuned[] = {", note + "static const TunedEntry kTuned[] = {", 1)
    TABLE.write_text(text)
    wins = sum(b["speedup_vs_lt_auto_max"] >= 1.0 for b in best.values())
    print(f"{len(best)} shapes, {switched} switches, {changed} table entries rewritten; selected records: {wins} at or above 1.0 "
          f"(mean {sum(b['speedup_vs_lt_auto_max'] for b in best.values()) / len(best):.4f})")
    if out:
        out.write_text("".join(json.dumps(b) + "\n" for _, b in sorted(best.items())))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
