"""Exactness check on 0/1 matrices — the parity contract of the harness.

Procedure and pass rule follow the reference's zero_one_correctness_check.py (cited per step below); the
code is restructured, seeded, able to run its plumbing without a GPU (``--device cpu``), and it EXITS
NON-ZERO on a failed check (the reference prints FAILED but exits 0, zero_one_correctness_check.py:302-305).

  inputs      i.i.d. uniform over {0,1}, or {0,0,1} when max(M,N,K) > 8192           (ref :65-73)
  truth       torch.matmul(a.cpu().float(), b.cpu().float()).half()                    (ref :87-90)
  mask        entries with |truth| > 2047 are ignored (fp16 integers are exact < 2048) (ref :92,:170)
  guard bands the kernel under test gets a, b, b_col_major and c carved out of 1-D buffers with 16384 random
              elements on both sides; all eight bands must be bit-identical afterwards  (ref :101-150)
  verdict     mean over iterations of max|out - truth| must be exactly 0, finite, no band touched
                                                                                        (ref :169-172,:253-268)
  budget      at most 100 iterations or 60 s                                            (ref :60,:77-79,:227)
"""
from __future__ import annotations

import json
import math
import time
from dataclasses import dataclass, field

import torch

from tools.utils import as_col_major

from .common import Padding

GUARD = 16384
MASK_ABOVE = 2047.0


def zero_one_levels(m: int, n: int, k: int) -> int:
    return 2 if max(m, n, k) <= 8192 else 3


def draw_zero_one(shape, levels: int, device, generator=None) -> torch.Tensor:
    """Each element is 1 with probability 1/levels, else 0 (fp16)."""
    idx = torch.randint(0, levels, shape, device=device, generator=generator)
    return (idx == levels - 1).to(torch.half).contiguous()


def reference_truth(a: torch.Tensor, b: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    truth = torch.matmul(a.detach().cpu().float(), b.detach().cpu().float()).half()
    return truth, truth.abs() > MASK_ABOVE


class GuardedOperand:
    """A [rows, cols] fp16 matrix living inside a 1-D buffer with GUARD random elements before and after."""

    def __init__(self, rows: int, cols: int, device):
        self.n = rows * cols
        self.flat = torch.randn(self.n + 2 * GUARD, dtype=torch.half, device=device)
        self.snapshot = self.flat.clone()
        self.view = self.flat[GUARD:GUARD + self.n].view(rows, cols)
        self.view.zero_()

    def bands_intact(self) -> bool:
        head = torch.equal(self.flat[:GUARD], self.snapshot[:GUARD])
        tail = torch.equal(self.flat[GUARD + self.n:], self.snapshot[GUARD + self.n:])
        return head and tail


@dataclass
class CheckResult:
    success: bool
    message: str
    result: dict = field(default_factory=dict)

    def to_json(self) -> dict:
        return {"success": self.success, "message": self.message, "result": self.result}


def _sync(device):
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()


@torch.no_grad()
def run_zero_one_check(*, kernel_funcs: list, kernel_under_test_name: str, m: int, n: int, k: int, padding: Padding,
                       device, num_iterations: int = 100, max_seconds: float = 60.0, generator=None) -> CheckResult:
    """``kernel_funcs``: callables ``f(a, b, b_col_major, out)`` (or ``torch.matmul``, called as
    ``matmul(a, b, out=out)``), identified by ``__name__``; exactly one of them is the kernel under test."""
    names = [f.__name__ for f in kernel_funcs]
    if kernel_under_test_name not in names:
        raise ValueError(f"{kernel_under_test_name} is not among the functions to run: {names}")
    diffs: dict[str, list[float]] = {nm: [] for nm in names}
    levels = zero_one_levels(m, n, k)
    pm, pk, pn = padding.m, padding.k, padding.n
    bands_ok = True
    t_start = time.time()
    done = 0
    for _ in range(num_iterations):
        if time.time() - t_start > max_seconds:
            break
        a = draw_zero_one((m, k), levels, device, generator)
        b = draw_zero_one((k, n), levels, device, generator)
        _sync(device)
        truth, mask = reference_truth(a, b)
        for func in kernel_funcs:
            tag = func.__name__
            if tag == kernel_under_test_name:
                ga = GuardedOperand(m + pm, k + pk, device)
                gb = GuardedOperand(k + pk, n + pn, device)
                gbt = GuardedOperand(k + pk, n + pn, device)
                gc = GuardedOperand(m + pm, n + pn, device)
                ga.view[:m, :k] = a
                gb.view[:k, :n] = b
                gbt.view.copy_(as_col_major(gb.view))
                for g in (ga, gb, gbt, gc):
                    assert g.view.is_contiguous()
                _sync(device)
                func(ga.view, gb.view, gbt.view, gc.view)
                _sync(device)   # an asynchronous fault surfaces here and aborts the check (ref :161-165)
                bands_ok &= all(g.bands_intact() for g in (ga, gb, gbt, gc))
                out = gc.view
            else:
                a_use, b_use = a.clone(), b.clone()
                out = torch.zeros((m, n), dtype=torch.half, device=device)
                _sync(device)
                if tag == "matmul":
                    func(a_use, b_use, out=out)
                else:
                    func(a_use, b_use, as_col_major(b_use), out)
                _sync(device)
            got = out[:m, :n].cpu()
            diff = (got - truth).abs()
            diff[mask] = 0.0
            diffs[tag].append(float(diff.max().item()))
        done += 1

    summary: dict = {"if_success": True, "m": m, "n": n, "k": k, "num_iterations": num_iterations,
                     "iterations_run": done, "levels": levels}
    for tag, vals in diffs.items():
        summary[f"avg_{tag}_diff"] = round(sum(vals) / len(vals), 6) if vals else float("nan")
    finite = {t: v for t, v in ((t, summary[f"avg_{t}_diff"]) for t in names) if math.isfinite(v)}
    summary["best_kernel"] = min(finite, key=finite.get) if finite else None

    if done == 0:
        return CheckResult(False, "no iteration completed inside the time budget", summary)
    if not bands_ok:
        return CheckResult(False, "memory overflow detected.", summary)
    mine = summary[f"avg_{kernel_under_test_name}_diff"]
    if not math.isfinite(mine):
        return CheckResult(False, f"{kernel_under_test_name} has nan or Inf value: {mine}", summary)
    others = [v for t, v in finite.items() if t != kernel_under_test_name]
    worst_other = max(others) if others else 0.0
    if mine > 0.0:
        return CheckResult(False, f"{kernel_under_test_name} diff ({mine:.6f}) exceeds 0 (max_other: {worst_other:.6f})",
                           summary)
    return CheckResult(True, f"Precise Correctness check passed: v2_diff={mine:.6f}, max_other={worst_other:.6f}", summary)


def cpu_stand_in(name: str):
    """A CPU function with the kernel calling convention, used ONLY by ``--device cpu`` plumbing runs: it
    reads ``a`` and the K-major ``b_col_major`` like the GPU kernel does and writes ``c`` through the
    reference's truth expression. It exercises generator, layout, guard bands, mask and verdict without a GPU."""

    def fn(a, b, b_col_major, c):
        kk, nn = b_col_major.shape
        bt = b_col_major.reshape(nn, kk)          # storage is [N,K]
        c.copy_(torch.matmul(a.float(), bt.float().t()).half())

    fn.__name__ = name
    return fn


def write_result(base_dir, result: CheckResult) -> None:
    from .common import result_dir

    with open(result_dir(base_dir) / "zero_one_correctness_check_result.json", "w") as f:
        json.dump(result.to_json(), f, indent=4, ensure_ascii=False)
