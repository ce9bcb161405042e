"""Developer tools that shape the shipped dispatcher: the tuner's reading of `dev_check grid` output."""
import importlib.util

from conftest import REPO


def _tuner():
    spec = importlib.util.spec_from_file_location("tune_b200", REPO / "tools" / "tune_b200.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_tuner_merges_repeated_runs_and_keeps_the_simple_schedule_on_near_ties(tmp_path):
    t = _tuner()
    # GRID,acc,M,N,K,cublas_us,best_cfg,best_gm,best_splits,best_us,cfg:gm:splits:us...
    run1 = tmp_path / "a.csv"
    run2 = tmp_path / "b.csv"
    run1.write_text("GRID,32,256,256,4096,10.0,1,0,-4,8.0,1:0:-4:8.0,2:0:1:9.0,0:0:1:12.0\n"
                    "GRID,32,4096,4096,4096,95.0,0,4,1,94.0,0:4:1:94.0,3:8:1:94.9,6:8:1:96.0\n"
                    "noise line\n")
    run2.write_text("GRID,32,256,256,4096,10.0,2,0,1,8.5,1:0:-4:10.0,2:0:1:8.5,0:0:1:12.0\n"
                    "GRID,16,256,256,4096,10.0,2,0,1,7.0,2:0:1:7.0\n")
    one = t.parse([str(run1)])
    assert one[(256, 256, 4096)][32][:3] == (1, 0, -4)          # a clear winner may be a split-K schedule
    assert one[(4096, 4096, 4096)][32][:3] == (3, 8, 1)         # within 1.5 %: the default-raster CTA-pair schedule is kept
    both = t.parse([str(run1), str(run2)])
    # config 1 / cluster split-K: mean rate of (8.0, 10.0) us = 8.89 us; config 2: (9.0, 8.5) -> 8.74 us: config 2 wins over two runs
    assert both[(256, 256, 4096)][32][:3] == (2, 0, 1)
    assert abs(both[(256, 256, 4096)][32][3] - 2.0 / (1 / 9.0 + 1 / 8.5)) < 1e-9
    assert both[(256, 256, 4096)][16][:3] == (2, 0, 1)
    assert both[(4096, 4096, 4096)][32][:3] == (3, 8, 1)        # measured in one run only: still judged on that run
