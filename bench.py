#!/usr/bin/env python
"""bench.py — the measured headline of this repository.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mnk M_N_K] [--acc fp32|fp16] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): HGEMM TFLOP/s, offline mode (back-to-back calls), per (M,N,K). One "step" is ONE
GEMM of the workload shape through the C ABI (libb200_hgemm.so). Default workload = BASELINE config[1]:
4096_4096_4096, fp32 accumulate.  Synthetic N(0,1) fp16 operands, as the reference harness draws them
(benchmarking_utils.py:36-37).

What is timed
  value   K steps with operands resident in HBM, CUDA events on the launching (legacy default) stream, bracketed
          by barrier + device synchronize, MAX over ranks; FLOPs = 2*M*N*K per step per rank (unpadded).
          Successive steps rotate over operand sets whose total size exceeds the 126 MB L2 several times.
  e2e     the same GEMM through b200_hgemm_host(): pinned HOST buffers, H2D of A and B + kernel + D2H of C inside
          the timed region every step.
  roofline  achieved = 2MNK / (average kernel duration from the same CUDA events); peak = measured cuBLAS bf16
          burst TFLOP/s from MEASURED_PEAKS.json (fallback 1590, said so in `peak_source`).
  cpu_baseline  the reference's CPU path, verbatim: torch.matmul(a.float(), b.float()).half() on the host cores
          (zero_one_correctness_check.py:87-90), a bounded ~10 s sample of the same shape, rank 0, N=1 only.
  --impl reference   times that same CPU path as the whole job (the reference has no other CPU implementation
          and its GPU kernels target sm_80/sm_90 — see DESIGN.md), K bounded steps, rank 0 only.

Multi-GPU: the path shards by problem (one GEMM per GPU, no collective on the data path), so
at N > 1 every rank runs the same per-GPU work ("scaling": "weak") and value is the sum over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

import torch  # noqa: E402

L2_BYTES = 126 * 1024 * 1024
FALLBACK_TFLOPS, FALLBACK_HBM = 1590.0, 6650.0


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2000)
    p.add_argument("--warmup", type=int, default=20)
    p.add_argument("--mnk", type=str, default="4096_4096_4096")
    p.add_argument("--acc", type=str, default="fp32", choices=["fp32", "fp16"])
    p.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    p.add_argument("--e2e_steps", type=int, default=0, help="steps of the host-buffer leg (default min(steps, 50))")
    p.add_argument("--cpu_seconds", type=float, default=10.0, help="bound on the cpu_baseline sample")
    return p.parse_args()


def peaks():
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            d = json.loads(f.read_text())
            return (float(d["bf16_tflops"]), float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, cuBLAS bf16 burst)",
                    float(d.get("bf16_tflops_sustained", 0.0)) or None)
        except Exception:
            pass
    return FALLBACK_TFLOPS, FALLBACK_HBM, "fallback (B200_PROFILING.md)", 1400.0


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed regions, every few milliseconds through NVML
    (nvidia-ml-py) from a background thread; falls back to `nvidia-smi -lms` when NVML cannot be loaded."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.gpu, self.samples, self.thread, self.stop_flag, self.proc, self.path = gpu_index, [], None, False, None, None
        self.max_mhz = None

    def _physical_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.gpu])
            except (ValueError, IndexError):
                pass
        return self.gpu

    def start(self):
        try:
            import threading

            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))

            def loop():
                while not self.stop_flag:
                    try:
                        self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                             nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                             int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
                    except Exception:
                        pass
                    time.sleep(0.004)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.proc = subprocess.Popen([exe, f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20", "-i",
                                      str(self._physical_index())], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def stop(self) -> dict:
        sm, power, reasons = [], [], set()
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            for mhz, watts, mask in self.samples:
                sm.append(mhz); power.append(watts)
                reasons.update(nm for bit, nm in self.REASONS.items() if mask & bit)
            source = "nvml"
        elif self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in Path(self.path).read_text().splitlines():
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0])); self.max_mhz = float(f[1]); power.append(float(f[2]))
                except (ValueError, IndexError):
                    continue
                reasons.update(nm for nm, v in zip(names, f[3:7]) if v.lower().startswith("active"))
            os.unlink(self.path)
            source = "nvidia-smi"
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "neither NVML nor nvidia-smi available"}
        # "under load" = samples taken while the board drew more than half of the highest power seen
        load = [c for c, w in zip(sm, power) if w >= 0.5 * max(power)] if power else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(load),
                "power_w_max": max(power) if power else None, "source": source}


def reference_cpu_step(a32, b32):
    return torch.matmul(a32, b32).half()          # the reference's truth path (zero_one_correctness_check.py:87-90)


def cpu_baseline(m, n, k, seconds: float) -> dict:
    import oracle  # the one place bench.py may use oracle/: the reported CPU baseline (never the measured path)

    g = torch.Generator().manual_seed(0)
    a = torch.randn((m, k), generator=g).half()
    b = torch.randn((k, n), generator=g).half()
    reference_cpu_step(a.float(), b.float())      # warm-up
    reps, t0 = 0, time.perf_counter()
    while True:
        oracle.reference_truth(a, b)
        reps += 1
        el = time.perf_counter() - t0
        if (reps >= 3 and el >= seconds) or el >= 3 * seconds:
            break
    tf = 2.0 * m * n * k * reps / el * 1e-12
    return {"value": tf, "unit": "TFLOP/s", "cores": oracle.cpu_threads(), "host_cpus": os.cpu_count(),
            "kind": "reference",
            "sample": f"{reps} x torch.matmul(a.cpu().float(), b.cpu().float()).half() at {m}x{n}x{k} in {el:.1f} s"}


def run_reference(args, m, n, k, rank, world):
    """--impl reference: the reference's own CPU implementation of the path as the whole job (rank 0 only)."""
    if rank != 0:
        return
    import oracle

    g = torch.Generator().manual_seed(0)
    a = torch.randn((m, k), generator=g).half()
    b = torch.randn((k, n), generator=g).half()
    per_step = 2.0 * m * n * k
    # bounded sample: a step is one full GEMM of the workload; cap the step count so the run stays in minutes
    for _ in range(max(1, min(args.warmup, 2))):
        oracle.reference_truth(a, b)
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        oracle.reference_truth(a, b)
        done += 1
        if time.perf_counter() - t0 > 120.0:
            break
    el = time.perf_counter() - t0
    tf = per_step * done / el * 1e-12
    cores = oracle.cpu_threads()
    sample = f"{done} of {steps} requested steps, each one {m}x{n}x{k} fp32 torch.matmul + .half() on {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "HGEMM TFLOP/s, offline mode, per (M,N,K)", "value": tf, "unit": "TFLOP/s",
        "n_gpus": world, "steps": done, "warmup": args.warmup, "ms_per_step": el / done * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (fp16 operands widened)",
        "data": "synthetic N(0,1) fp16",
        "config": {"workload": f"{m}_{n}_{k} --acc_precise {args.acc} --mode offline", "parallelism": "cpu threads",
                   "note": "reference CPU path = its ground-truth expression; its GPU kernels target sm_80/sm_90"},
        "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "host_cpus": os.cpu_count(),
                         "kind": "reference", "sample": sample},
        "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    m, n, k = (int(x) for x in args.mnk.split("_"))
    if args.impl == "reference":
        run_reference(args, m, n, k, rank, world)
        return 0

    from cuda_l2_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the HGEMM path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    capi.hgemm_lib()   # fail loudly now if the library is missing

    # operand sets: enough of them that a step never finds its inputs in L2
    set_bytes = 2 * (m * k + n * k + m * n)
    nsets = max(2, min(16, -(-4 * L2_BYTES // set_bytes)))
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    sets = []
    for _ in range(nsets):
        a = torch.randn((m, k), device="cuda", generator=g).half()
        bt = torch.randn((n, k), device="cuda", generator=g).half()      # K-major B (the harness's b_col_major)
        c = torch.empty((m, n), dtype=torch.half, device="cuda")
        sets.append((a, bt.view(k, n), c))

    def step(i):
        a, b_col_major, c = sets[i % nsets]
        capi.hgemm(a, b_col_major, c, args.acc)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, 3)
    for i in range(warm):
        step(i)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.1)
    launches0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    launches = capi.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    flops_step = 2.0 * m * n * k
    value = flops_step * args.steps * world / (ms_total * 1e-3) * 1e-12

    # ---- end-to-end leg: host buffers through the C ABI, copies inside the timed region
    ha = torch.randn((m, k)).half().pin_memory()
    hbt = torch.randn((n, k)).half().pin_memory()
    hc = torch.empty((m, n), dtype=torch.half).pin_memory()
    e2e_steps = args.e2e_steps or max(1, min(args.steps, 50))
    for _ in range(3):
        capi.hgemm_host(ha, hbt.view(k, n), hc, args.acc)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        capi.hgemm_host(ha, hbt.view(k, n), hc, args.acc)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = flops_step * e2e_steps * world / e2e_s * 1e-12
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peak_tf, peak_hbm, peak_src, peak_sustained = peaks()
        achieved = flops_step / (ms_total / args.steps * 1e-3) * 1e-12      # per launch, from the CUDA events above
        cfg_id, group_m, splits = capi.select(args.acc, m, n, k)
        cfg = capi.configs()[cfg_id]
        traffic = None
        tf = REPO / "profiles" / "dram_traffic.json"
        if tf.exists():
            traffic = json.loads(tf.read_text()).get(f"{args.mnk}_{args.acc}")
        out = {
            "metric": "HGEMM TFLOP/s, offline mode, per (M,N,K)",
            "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 x f16 -> " + ("f32" if args.acc == "fp32" else "f16") + " accumulate -> f16",
            "data": "synthetic N(0,1) fp16",
            "config": {"workload": f"{args.mnk} --acc_precise {args.acc} --mode offline", "parallelism": f"1 GEMM per GPU x {world}",
                       "l2_policy": f"rotating {nsets} operand sets ({nsets * set_bytes >> 20} MiB > 126 MiB L2)",
                       "kernel_config": {"tile": f"{128 * cfg['cta_group'] * cfg.get('m_rep', 1)}x{cfg['bn']}x64", "stages": cfg["stages"],
                                         "cta_group": cfg["cta_group"], "cluster": f"{cfg['cluster_m']}x{cfg['cluster_n']}", "group_m": group_m,
                                         "split_k": splits}},
            "e2e": {"value": e2e_value, "unit": "TFLOP/s", "h2d_bytes_per_step": 2 * (m * k + n * k),
                    "d2h_bytes_per_step": 2 * m * n, "steps": e2e_steps, "api": "b200_hgemm_host (pinned host buffers)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf, "traffic": traffic, "peak_source": peak_src,
                         "peak_sustained": peak_sustained, "frac_of_sustained": (achieved / peak_sustained) if peak_sustained else None,
                         "algorithmic_flops_per_launch": flops_step, "algorithmic_bytes_per_launch": set_bytes},
        }
        if world == 1:
            out["cpu_baseline"] = cpu_baseline(m, n, k, args.cpu_seconds)
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
