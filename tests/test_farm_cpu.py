"""Sharding + gathering logic of the multi-GPU sweep, covered on the CPU (gloo, world_size 2)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

from conftest import REPO
from cuda_l2_b200 import farm


def test_grid_and_partition_cover_every_shape_exactly_once():
    shapes = farm.grid_shapes()
    assert len(shapes) == 1001 and len(set(shapes)) == 1001 and (2048, 11008, 4096) in shapes
    for world in (1, 2, 4, 8):
        parts = farm.partition(shapes, world)
        flat = [s for p in parts for s in p]
        assert sorted(flat) == sorted(shapes)
        loads = [sum(farm.estimated_cost(s) for s in p) for p in parts]
        assert max(loads) / (sum(loads) / world) < 1.05          # LPT keeps ranks balanced
        assert parts[0][0] == (16384, 16384, 16384)               # longest first
    assert farm.partition(shapes, 8) == farm.partition(shapes, 8)  # deterministic


def test_wall_line_and_speedup_row():
    line = ("WALL,32,4096,4096,4096,samples=100,cfg=3,gm=8,splits=1,lt_candidates=7/9,ours=1200.000,cublas_tn=1000.000,"
            "cublas_nn=1100.000,lt_heur_tn=900.000,lt_heur_nn=1000.000,lt_auto_tn=1150.000,lt_auto_nn=1250.000,"
            "speedup_vs_lt_auto_max=0.960,ours_us=114.5,lt_auto_tn_us=119.5")
    rec = farm.parse_wall_line(line)
    assert (rec["m"], rec["n"], rec["k"], rec["cfg"]) == (4096, 4096, 4096, 3.0) and rec["lt_candidates"] == "7/9"
    row = farm.speedup_row("4096_4096_4096", rec)
    assert row["cuBLAS-max"] == pytest.approx(1200 / 1100) and row["cuBLASLt-auto-tuning-max"] == pytest.approx(0.96)
    assert row["cuBLASLt-heuristic-tn"] == pytest.approx(1200 / 900) and row["torch.matmul"] == ""


def test_run_partition_isolates_failures_and_resumes(tmp_path):
    calls = []

    def engine(s):
        calls.append(s)
        if s == (128, 128, 128):
            raise RuntimeError("boom")
        return {"ours": 10.0, "cublas_tn": 5.0, "cublas_nn": 5.0, "lt_heur_tn": 5.0, "lt_heur_nn": 5.0, "lt_auto_tn": 8.0,
                "lt_auto_nn": 9.0, "speedup_vs_lt_auto_max": 10 / 9}
    shapes = [(64, 64, 64), (128, 128, 128), (256, 256, 256)]
    out = tmp_path / "w.jsonl"
    res = farm.run_partition(0, shapes, engine, out)
    assert [r["ok"] for r in res] == [True, False, True]
    done = set(farm.load_done([out]))
    assert done == {"64_64_64", "256_256_256"}
    calls.clear()
    farm.run_partition(0, shapes, engine, out, done)
    assert calls == [(128, 128, 128)]           # only the failed shape is retried
    summary = farm.write_reports(list(farm.load_done([out]).values()), tmp_path / "r.csv", 1600.0, 6500.0)
    assert summary["shapes"] == 2 and summary["won_vs_lt_auto_max"] == 2
    header = (tmp_path / "r.csv").read_text().splitlines()[0]
    assert header == ",".join(farm.CSV_COLUMNS)


WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from cuda_l2_b200 import farm
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
shapes = [s for s in farm.grid_shapes() if max(s) <= 256]           # 27 small shapes
mine = farm.partition(shapes, world)[rank]
res = farm.run_partition(rank, mine, lambda s: {"ours": float(sum(s)), "rank_seen": rank})
bucket = farm.gather(res)
if rank == 0:
    flat = [r for part in bucket for r in part]
    json.dump({"n": len(flat), "keys": sorted(r["mnk"] for r in flat), "ranks": sorted({r["rank"] for r in flat})},
              open(sys.argv[2], "w"))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gloo_sweep_gathers_everything_on_rank0(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    out = tmp_path / "out.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", str(script), str(REPO), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=REPO)
    assert r.returncode == 0, r.stdout + r.stderr
    got = json.loads(out.read_text())
    assert got["n"] == 27 and got["ranks"] == [0, 1]
    assert got["keys"] == sorted("_".join(map(str, s)) for s in farm.grid_shapes() if max(s) <= 256)


def test_harness_summary_becomes_a_sweep_record():
    summary = {name: {"Baseline Method Name": name, "Baseline TFLOPS": base, "CUDA-L2 TFLOPS": ours, "Speedup": ours / base}
               for name, base, ours in [("torch.matmul", 80.0, 120.0), ("cuBLAS-tn", 100.0, 120.0), ("cuBLAS-nn", 90.0, 117.0),
                                        ("cuBLASLt-heuristic-tn", 100.0, 121.0), ("cuBLASLt-heuristic-nn", 100.0, 119.0),
                                        ("cuBLASLt-auto-tuning-tn", 110.0, 121.0), ("cuBLASLt-auto-tuning-nn", 125.0, 120.0)]}
    rec = farm.record_from_harness_summary(summary)
    assert rec["speedup_vs_lt_auto_max"] == pytest.approx(120.0 / 125.0)
    row = farm.speedup_row("64_64_64", rec)
    assert row["cuBLAS-tn"] == pytest.approx(1.2) and row["cuBLAS-max"] == pytest.approx(1.2)     # tn is the harder layout
    assert row["cuBLASLt-auto-tuning-max"] == pytest.approx(0.96) and row["torch.matmul"] == pytest.approx(1.5)
