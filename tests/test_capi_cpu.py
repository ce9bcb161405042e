"""The C-ABI libraries load and export every symbol include/*.h declares (no compute without a GPU)."""
import ctypes
import re

import pytest

from conftest import REPO
from cuda_l2_b200 import capi

DECL = re.compile(r"^\s*(?:const\s+)?(?:unsigned\s+long\s+long|int|void|char\s*\*|const\s+char\s*\*)\s*\*?\s*(b200_\w+)\s*\(", re.M)


def declared(header: str) -> list[str]:
    return sorted(set(DECL.findall((REPO / "include" / header).read_text())))


def test_headers_and_python_binding_agree():
    assert declared("b200_hgemm.h") == sorted(capi.exported_symbols()["libb200_hgemm.so"])
    assert declared("b200_baselines.h") == sorted(capi.exported_symbols()["libb200_baselines.so"])


def test_libraries_export_every_declared_symbol(built_libs):
    for header, path in (("b200_hgemm.h", built_libs["capi"]), ("b200_baselines.h", built_libs["baselines"])):
        lib = ctypes.CDLL(str(path))
        for sym in declared(header):
            assert hasattr(lib, sym), f"{path.name} does not export {sym}"


def test_config_table_and_dispatch(built_libs):
    cfgs = capi.configs()
    assert len(cfgs) >= 5
    for c in cfgs:
        assert (c["bn"] == 32 or c["bn"] % 64 == 0) and c["bn"] <= 256 and c["cta_group"] in (1, 2) and c["stages"] >= 2
        assert c["cta_group"] * c["cluster_m"] * c["cluster_n"] <= 8
        assert c["m_rep"] in (1, 2)
        smem = 1024 + c["stages"] * (128 * c["m_rep"] * 64 * 2 + (c["bn"] // c["cta_group"]) * 64 * 2) + 32768 + 256
        assert smem + 256 <= 232448
    ids = {c["id"] for c in cfgs}
    for acc in ("fp32", "fp16"):
        for mnk in ((64, 4096, 64), (4096, 4096, 4096), (8192, 8192, 8192), (2048, 11008, 4096), (64, 64, 64),
                    (16384, 16384, 16384), (200, 328, 72)):
            cid, gm, sp = capi.select(acc, *mnk)
            assert cid in ids and gm >= 0 and (sp >= 1 or sp in (-2, -4, -8))
            if mnk[0] <= 128:
                assert cfgs[cid]["cta_group"] == 1     # a CTA pair would waste its second half on padding


def test_argument_validation_happens_before_any_cuda_call(built_libs):
    lib = capi.hgemm_lib()
    assert lib.b200_hgemm_f32acc(None, None, None, None, 64, 64, 64, None) == -5       # null pointers
    buf = ctypes.create_string_buffer(1 << 16)
    p = ctypes.addressof(buf)
    p = (p + 15) & ~15
    assert lib.b200_hgemm_f32acc(p, None, p, p, 0, 64, 64, None) == -1                  # bad shape
    assert lib.b200_hgemm_f16acc(p, None, p, p, 64, 64, 60, None) == -2                 # K % 8 != 0
    assert lib.b200_hgemm_f16acc(p, None, p, p, 64, 60, 64, None) == -2                 # N % 8 != 0
    assert lib.b200_hgemm_f32acc(p + 2, None, p, p, 64, 64, 64, None) == -2             # misaligned A
    assert lib.b200_hgemm_run_config(32, 99, p, p, p, 64, 64, 64, 0, 0, 1, None) == -6     # unknown config
    assert lib.b200_hgemm_run_config(8, 0, p, p, p, 64, 64, 64, 0, 0, 1, None) == -6       # unknown accumulator
    assert "16-byte" in capi.strerror(-2)
    assert capi.launch_count() == 0


def test_python_binding_rejects_cpu_tensors(built_libs):
    import torch
    a = torch.zeros(64, 64, dtype=torch.half)
    with pytest.raises(capi.B200HgemmError):
        capi.hgemm(a, a, a)          # CPU tensors: there is no CPU fallback
    with pytest.raises(capi.B200HgemmError):
        capi.hgemm(a.float(), a, a)


def test_library_keeps_its_template_statics_private(built_libs):
    """libb200_hgemm.so and a JIT-built hgemm_lib.so instantiate the same templates; a process-wide (STB_GNU_UNIQUE)
    static would let one library skip the per-kernel setup of the other (seen on the B200 as `invalid argument` on the
    first launch after the harness ran). The build uses -fno-gnu-unique: no 'u' symbols may remain."""
    import shutil
    import subprocess
    nm = shutil.which("nm")
    if nm is None:
        pytest.skip("nm not available")
    out = subprocess.run([nm, "-D", str(built_libs["capi"])], capture_output=True, text=True, check=True).stdout
    assert not [ln for ln in out.splitlines() if " u " in ln]


def test_tuned_table_entries_are_launchable_for_every_grid_shape(built_libs):
    """Every tuned entry must name an existing configuration whose cluster is not wider than the problem, and
    split-K (either flavour) only on plain single-CTA configurations with tiles of at least 64 columns."""
    from cuda_l2_b200 import farm
    cfgs = {c["id"]: c for c in capi.configs()}
    seen_cluster_split = seen_mcast = 0
    for (m, n, k) in farm.grid_shapes():
        for acc in ("fp32", "fp16"):
            cid, gm, sp = capi.select(acc, m, n, k)
            c = cfgs[cid]
            assert 0 <= gm <= 64
            assert -(-m // 128) >= c["cta_group"] * c["cluster_m"] * c["m_rep"], (m, n, k, acc, c)
            assert -(-n // c["bn"]) >= c["cluster_n"], (m, n, k, acc, c)
            if sp != 1:
                assert sp in (-2, -4, -8) or 2 <= sp <= 64
                assert c["cta_group"] == 1 and c["cluster_m"] * c["cluster_n"] == 1 and c["bn"] >= 64, (m, n, k, acc, sp, c)
                seen_cluster_split += sp < 0
            seen_mcast += c["cluster_m"] * c["cluster_n"] > 1
    assert seen_cluster_split > 50 and seen_mcast > 20          # the table really uses both mechanisms


def test_off_grid_shapes_borrow_the_nearest_tuned_entry(built_libs):
    assert capi.select("fp32", 2048, 11008, 4096) == capi.select("fp32", 2048, 11008, 4096)
    assert capi.select("fp32", 4000, 4100, 4090) == capi.select("fp32", 4096, 4096, 4096)
    assert capi.select("fp16", 70, 60 * 8, 16000) == capi.select("fp16", 64, 512, 16384)
    cid, _, _ = capi.select("fp32", 8, 16, 8192)               # far off the grid in M and N: still a valid choice
    cfg = capi.configs()[cid]
    assert cfg["cta_group"] == 1 and cfg["cluster_m"] * cfg["cluster_n"] == 1
