"""The CPU oracle against the reference-generated golden vectors (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import oracle


def test_binary16_conversions_match_numpy_exhaustively():
    L = oracle.lib()
    h = np.arange(65536, dtype=np.uint16)
    want = h.view(np.float16).astype(np.float32)
    got = np.array([L.oracle_f16_to_f32(int(x)) for x in h], dtype=np.float32)
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan)
    assert np.array_equal(got[~nan], want[~nan])
    # every half converts back to itself
    back = np.array([L.oracle_f32_to_f16(float(v)) for v in want[~nan]], dtype=np.uint16)
    assert np.array_equal(back, h[~nan])


def test_f32_to_f16_rounding_matches_numpy():
    L = oracle.lib()
    rng = np.random.default_rng(7)
    scales = np.array([1e-8, 6e-8, 1e-5, 6.1e-5, 1e-3, 1.0, 100.0, 2048.0, 6.5e4], dtype=np.float32)
    xs = (rng.standard_normal(60000).astype(np.float32) * rng.choice(scales, 60000)).astype(np.float32)
    edge = np.array([65504, 65519.996, 65520, 65536, 2**-25, np.nextafter(np.float32(2**-25), np.float32(1)), 2**-24,
                     0.0, -0.0, np.inf, -np.inf, 2047.5, 2048.5, 2049.0, 1e10, -1e10], dtype=np.float32)
    with np.errstate(over="ignore"):
        for x in np.concatenate([xs, edge]):
            assert L.oracle_f32_to_f16(float(x)) == int(np.float32(x).astype(np.float16).view(np.uint16))


def test_zero_one_golden_truth_is_reproduced_bit_exactly(zero_one_cases):
    for c in zero_one_cases:
        bt = oracle.as_col_major(c["b"])
        assert np.array_equal(bt, c["b"].T)
        truth = c["truth"]
        unmasked = np.abs(truth.astype(np.float32)) <= 2047
        for got in (oracle.hgemm_f32acc(c["a"], bt), oracle.hgemm_f32acc(c["a"], bt, fast=True)):
            # fp32 accumulation of 0/1 products is exact below 2^24, so even the masked entries agree
            assert np.array_equal(got.view(np.uint16), truth.view(np.uint16)), (c["m"], c["n"], c["k"])
        got16 = oracle.hgemm_f16acc(c["a"], bt)
        assert np.array_equal(got16[unmasked], truth[unmasked])
        d, n_masked, n_bad = oracle.zero_one_max_diff(got16, truth)
        assert d == 0.0 and n_bad == 0 and n_masked == int((~unmasked).sum())


def test_mask_case_really_masks(zero_one_cases):
    c = next(c for c in zero_one_cases if c["k"] == 8192)
    frac = float((np.abs(c["truth"].astype(np.float32)) > 2047).mean())
    assert 0.2 < frac < 0.8      # E[c] = K/4 = 2048: about half the entries are masked


def test_randn_golden_within_tolerance(randn_cases):
    # Outside the 0/1 domain the reference pins nothing; our stated tolerance vs its fp32 truth expression:
    # |err| <= 2^-10 * |truth| + 1e-3 (one fp16 rounding of the result plus fp32 summation-order noise).
    for c in randn_cases:
        bt = oracle.as_col_major(c["b"])
        for got in (oracle.hgemm_f32acc(c["a"], bt), oracle.hgemm_f32acc(c["a"], bt, fast=True)):
            err = np.abs(got.astype(np.float32) - c["truth"].astype(np.float32))
            tol = 2.0**-10 * np.abs(c["truth"].astype(np.float32)) + 1e-3
            assert (err <= tol).all()


def test_reference_truth_expression_matches_golden(zero_one_cases):
    c = zero_one_cases[3]
    t = oracle.reference_truth(torch.from_numpy(c["a"]), torch.from_numpy(c["b"])).numpy()
    assert np.array_equal(t, c["truth"])


def test_density_rule():
    assert oracle.zero_one_levels(8192, 64, 64) == 2
    assert oracle.zero_one_levels(64, 12288, 64) == 3
    x = oracle.fill_zero_one((512, 512), 3, 5)
    assert set(np.unique(x).tolist()) <= {0.0, 1.0}
    assert 0.28 < float(x.mean()) < 0.39


def test_max_diff_flags_errors():
    truth = np.array([[1, 2, 3000, 4]], dtype=np.float16)
    out = np.array([[1, 2, 7, 5]], dtype=np.float16)       # error at a masked entry AND at an unmasked one
    d, n_masked, n_bad = oracle.zero_one_max_diff(out, truth)
    assert d == 1.0 and n_masked == 1 and n_bad == 0
    out[0, 0] = np.inf
    assert oracle.zero_one_max_diff(out, truth)[2] == 1


def test_bf16_oracle_matches_the_torch_truth_expression(bf16_cases):
    """bf16 variant (this repository's extension): the C restatement against fixtures generated with
    torch.matmul(a.float(), b.float()).bfloat16() — bit-exact on the integer cases (every sum fits bf16's 8 bits),
    within one bf16 ulp + summation noise on N(0,1) operands."""
    for c in bf16_cases:
        bt = np.ascontiguousarray(c["b"].T)
        got = oracle.bgemm_f32acc(c["a"], bt)
        if c["kind"] in ("01", "int"):
            assert np.abs(oracle.bf16_bits_to_f32(c["truth"])).max() <= 256
            assert np.array_equal(got, c["truth"]), (c["m"], c["n"], c["k"], c["kind"])
        else:
            g, t = oracle.bf16_bits_to_f32(got), oracle.bf16_bits_to_f32(c["truth"])
            assert (np.abs(g - t) <= 2.0**-7 * np.abs(t) + 1e-2).all()


def test_bf16_conversions_round_to_nearest_even():
    lib = oracle.lib()
    assert lib.oracle_f32_to_bf16(1.0) == 0x3F80 and lib.oracle_bf16_to_f32(0x3F80) == 1.0
    assert lib.oracle_f32_to_bf16(1.00390625) == 0x3F80          # 1 + 2^-8: tie -> even (down)
    assert lib.oracle_f32_to_bf16(1.01171875) == 0x3F82          # 1 + 3*2^-8: tie -> even (up)
    assert lib.oracle_f32_to_bf16(257.0) == 0x4380               # 257 -> 256 (ties to even)
    assert lib.oracle_f32_to_bf16(float("inf")) == 0x7F80
    assert (lib.oracle_f32_to_bf16(float("nan")) & 0x7FC0) == 0x7FC0
    x = np.random.default_rng(0).standard_normal(4096).astype(np.float32) * 100
    want = torch.from_numpy(x).bfloat16().view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(oracle.f32_to_bf16_bits(x), want)
    assert all(lib.oracle_f32_to_bf16(float(v)) == int(w) for v, w in zip(x[:256], want[:256]))
