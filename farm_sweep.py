#!/usr/bin/env python
"""Run the (M,N,K) sweep across the GPUs of one box (BASELINE config 5) and write the eval_results CSVs.

    python farm_sweep.py --gpus 8 --acc_precise fp32 --seconds 0.5                     # spawns one worker per GPU
    python -m torch.distributed.run --nproc-per-node 8 farm_sweep.py --acc_precise fp32  # same, under torchrun
    python farm_sweep.py --gpus 1 --shapes 64_4096_64,4096_4096_4096 --seconds 1         # a few shapes

One problem per GPU at a time, no NCCL on the GEMM path; results are gathered at the end. Interrupted runs resume
from {base_dir}/worker_*.jsonl. Output: eval_results/cuda_l2_b200_<ACC>_speedup_<mode>.csv in the reference's schema
plus *_absolute.csv (TFLOP/s, roofline fraction) and *_summary.json (win fraction vs cuBLASLt-auto-tuning-max).
"""
import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

from cuda_l2_b200 import farm  # noqa: E402


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=0, help="GPUs to use (0 = all visible; ignored under torchrun)")
    p.add_argument("--acc_precise", default="fp32", choices=["fp32", "fp16"])
    p.add_argument("--seconds", type=float, default=0.5, help="sampling time per shape (wall engine)")
    p.add_argument("--tune_rounds", default="10,30", help="cuBLASLt auto-tuning warm-up,timed rounds (reference: 50,100)")
    p.add_argument("--shapes", default="grid", help="'grid' (1000 + 2048_11008_4096) or a comma list of M_N_K")
    p.add_argument("--limit", type=int, default=0, help="evaluate only the first N shapes (per GPU for the wallgrid engine)")
    p.add_argument("--engine", default="auto", choices=["auto", "wallgrid", "wall", "harness", "pyharness"],
                   help="wallgrid: one dev_check process per GPU walks its share (default for --shapes grid); "
                        "wall: one dev_check process per shape (resumable, any shape list); "
                        "harness: eval_one_file.sh per shape (--seconds = benchmark seconds, warm-up = a third of it); "
                        "pyharness: the harness's Python timing loop in one process per GPU on the C-ABI libraries "
                        "(--perf_funcs, default matmul: fills the torch.matmul column; server mode supported)")
    p.add_argument("--base_dir", default=str(REPO / "gpurun_out" / "farm"))
    p.add_argument("--out_dir", default=str(REPO / "eval_results"))
    p.add_argument("--mode", default="offline", choices=["offline", "server"], help="server: harness engine only (eval_one_file.sh --mode server)")
    p.add_argument("--target_qps", type=float, default=100.0)
    p.add_argument("--perf_funcs", default="", help="harness engine: comma list of baselines to time (default all seven; "
                   "'auto' = the cuBLASLt-auto-tuning pair, which is all the sweep's target needs)")
    p.add_argument("--shapes_file", default="", help="wallgrid engine: a file of 'M N K' lines instead of the grid")
    p.add_argument("--merge_from", default="", help="base_dir of an earlier run of the same engine: its records are kept for every "
                   "shape this run did not measure again (partial re-sweep after a table change)")
    p.add_argument("--merge_matmul", default="", help="base_dir of a pyharness run whose torch.matmul pair fills that column of this report")
    p.add_argument("--finish_only", action="store_true", help="write the reports from existing worker_*.jsonl files")
    p.add_argument("--tag", default="", help="suffix of the report files, e.g. _harness_sample")
    p.add_argument("--import_wallgrid", default="", help="turn the stdout of a `dev_check wallgrid` run into the CSV reports")
    p.add_argument("--worker", type=int, default=-1, help=argparse.SUPPRESS)
    p.add_argument("--world", type=int, default=0, help=argparse.SUPPRESS)
    return p.parse_args(argv)


def shape_list(args):
    if args.shapes == "grid":
        shapes = farm.grid_shapes()
    else:
        shapes = [tuple(int(x) for x in s.split("_")) for s in args.shapes.split(",")]
    if args.limit:
        shapes = sorted(shapes, key=lambda s: -farm.estimated_cost(s))[: args.limit]
    return shapes


def worker(args, rank, world, gpu):
    bits = 32 if args.acc_precise == "fp32" else 16
    warm, bench = (int(x) for x in args.tune_rounds.split(",")) if "," in args.tune_rounds else (-25, 0)
    mine = farm.partition(shape_list(args), world)[rank]
    base = Path(args.base_dir)
    base.mkdir(parents=True, exist_ok=True)
    out = base / f"worker_{args.acc_precise}_{rank}.jsonl"
    engine_name = args.engine if args.engine != "auto" else ("wallgrid" if args.shapes == "grid" else "wall")
    if engine_name == "wallgrid":
        return farm.run_wallgrid_worker(rank, world, bits, args.seconds, (warm, bench), gpu, out, args.limit, args.shapes_file or None)
    done = set(farm.load_done([out]))
    if engine_name == "pyharness":
        names = [x for x in (args.perf_funcs or "matmul").split(",") if x]
        if args.perf_funcs == "auto":
            names = farm.AUTO_TUNING_PAIR.split(",")
        if gpu is not None:
            os.environ["CUDA_VISIBLE_DEVICES"] = str(gpu)       # before the first CUDA call of this worker process
        out = base / f"worker_{args.acc_precise}_{rank}.jsonl"
        return farm.run_pyharness_worker(rank, world, args.acc_precise, shape_list(args), args.seconds / 4, args.seconds, gpu, out,
                                         perf_funcs=names, mode=args.mode, target_qps=args.target_qps)
    if engine_name == "harness":
        funcs = farm.AUTO_TUNING_PAIR if args.perf_funcs == "auto" else (args.perf_funcs or None)
        engine = lambda s: farm.run_harness_engine(s, args.acc_precise, args.seconds / 3, args.seconds, gpu, base / "harness",
                                                   mode=args.mode, target_qps=args.target_qps, perf_funcs=funcs)
    else:
        engine = lambda s: farm.run_wall_engine(s, bits, args.seconds, (warm, bench), gpu)
    return farm.run_partition(rank, mine, engine, out, done)


def finish(args, world):
    import bench
    base = Path(args.base_dir)
    done = {}
    if args.merge_from:
        done.update(farm.load_done(sorted(Path(args.merge_from).glob(f"worker_{args.acc_precise}_*.jsonl"))))
    done.update(farm.load_done(sorted(base.glob(f"worker_{args.acc_precise}_*.jsonl"))))     # this run's records win
    recs = list(done.values())
    wanted = {"_".join(map(str, s)) for s in (farm.grid_shapes() if args.shapes == "grid" else shape_list(args))}
    if args.shapes_file and not args.merge_from:
        wanted = {"_".join(line.split()) for line in Path(args.shapes_file).read_text().splitlines() if line.strip()}
    recs = [r for r in recs if r["mnk"] in wanted]
    if args.merge_matmul:      # the torch.matmul column comes from a pyharness run (dev_check cannot call torch)
        extra = farm.load_done(sorted(Path(args.merge_matmul).glob(f"worker_{args.acc_precise}_*.jsonl")))
        for r in recs:
            if r["mnk"] in extra and extra[r["mnk"]].get("matmul_speedup"):
                r["matmul"] = r["ours"] / extra[r["mnk"]]["matmul_speedup"]     # the pair's own speed-up, on this record's scale
    recs = [r for r in recs if "speedup_vs_lt_auto_max" in r]
    peak_tf, peak_hbm, src, _ = bench.peaks()
    acc_dir = "F32F16F16F32" if args.acc_precise == "fp32" else "F16F16F16F16"
    out_csv = Path(args.out_dir) / f"cuda_l2_b200_{acc_dir}_speedup_{args.mode}{args.tag}.csv"
    summary = farm.write_reports(recs, out_csv, peak_tf, peak_hbm)
    summary.update({"n_gpus": world, "peak_source": src, "seconds_per_shape": args.seconds, "tune_rounds": args.tune_rounds,
                    "engine": args.engine, "mode": args.mode,
                    "missing": sorted(wanted - {r["mnk"] for r in recs})[:20]})
    out_csv.with_name(out_csv.stem + "_summary.json").write_text(json.dumps(summary, indent=1))
    print(json.dumps(summary))


def main(argv=None):
    args = parse_args(argv)
    t0 = time.time()
    if args.import_wallgrid:
        base = Path(args.base_dir)
        base.mkdir(parents=True, exist_ok=True)
        out = base / f"worker_{args.acc_precise}_0.jsonl"
        with open(out, "w") as f:
            for line in Path(args.import_wallgrid).read_text().splitlines():
                if line.startswith("WALL,"):
                    rec = farm.parse_wall_line(line)
                    rec.update(mnk=f"{rec['m']}_{rec['n']}_{rec['k']}", rank=0, ok=True)
                    f.write(json.dumps(rec) + "\n")
        finish(args, args.gpus or 1)
        return 0
    if args.finish_only:
        finish(args, args.gpus or 1)
        return 0
    if "RANK" in os.environ and args.worker < 0:          # under torchrun: one rank per GPU
        import torch.distributed as dist
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        dist.init_process_group("gloo")                    # results are tiny Python objects; no GPU collective needed
        worker(args, rank, world, int(os.environ.get("LOCAL_RANK", rank)))
        dist.barrier()
        if rank == 0:
            finish(args, world)
        dist.destroy_process_group()
        return 0
    if args.worker >= 0:                                   # child of the self-spawning orchestrator
        worker(args, args.worker, args.world, args.worker)
        return 0
    n = args.gpus
    if n <= 0:
        import torch
        n = max(1, torch.cuda.device_count())
    procs = [subprocess.Popen([sys.executable, __file__, *sys.argv[1:], "--worker", str(i), "--world", str(n)]) for i in range(n)]
    rc = [p.wait() for p in procs]
    finish(args, n)
    print(f"sweep wall time {time.time() - t0:.1f} s on {n} GPU(s); worker exit codes {rc}")
    return 0 if all(r == 0 for r in rc) else 1


if __name__ == "__main__":
    sys.exit(main())
