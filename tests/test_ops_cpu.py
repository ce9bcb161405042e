"""The deployable face (torch.library op + nn.Linear drop-in), as far as it can be checked without a GPU:
schema, shape inference on meta tensors, loud failure instead of a CPU fallback, module swapping."""
import pytest
import torch
from torch import nn

from cuda_l2_b200 import capi, ops


def test_op_is_registered_with_the_documented_schema():
    schema = str(torch.ops.cuda_l2_b200.hgemm.default._schema)
    assert schema.startswith("cuda_l2_b200::hgemm(Tensor a, Tensor b_kmajor, str acc") and schema.endswith("-> Tensor")


def test_shape_inference_on_meta_tensors_and_operand_checks():
    for dt in (torch.float16, torch.bfloat16):
        a, w = torch.empty(5, 64, dtype=dt, device="meta"), torch.empty(128, 64, dtype=dt, device="meta")
        out = torch.ops.cuda_l2_b200.hgemm(a, w, "fp32")
        assert out.shape == (5, 128) and out.dtype == dt
    a = torch.empty(5, 64, dtype=torch.half, device="meta")
    for bad_w, acc in ((torch.empty(128, 72, dtype=torch.half, device="meta"), "fp32"),        # K mismatch
                       (torch.empty(12, 64, dtype=torch.half, device="meta"), "fp32"),         # N % 8
                       (torch.empty(128, 64, dtype=torch.bfloat16, device="meta"), "fp32"),    # mixed dtypes
                       (torch.empty(128, 64, dtype=torch.half, device="meta"), "tf32")):       # unknown accumulator
        with pytest.raises(capi.B200HgemmError):
            torch.ops.cuda_l2_b200.hgemm(a, bad_w, acc)
    with pytest.raises(capi.B200HgemmError):      # bf16 accumulates in fp32 only
        torch.ops.cuda_l2_b200.hgemm(torch.empty(8, 64, dtype=torch.bfloat16, device="meta"),
                                     torch.empty(8, 64, dtype=torch.bfloat16, device="meta"), "fp16")


def test_no_cpu_fallback():
    with pytest.raises(capi.B200HgemmError, match="no CPU implementation"):
        ops.hgemm(torch.zeros(8, 64, dtype=torch.half), torch.zeros(16, 64, dtype=torch.half))
    lin = ops.B200Linear(64, 32)
    with pytest.raises(capi.B200HgemmError):
        lin(torch.zeros(4, 64, dtype=torch.half))


def test_replace_linear_modules_swaps_only_eligible_layers_and_shares_parameters():
    model = nn.Sequential(nn.Linear(64, 128, dtype=torch.half), nn.ReLU(),
                          nn.Linear(128, 20, dtype=torch.half),        # 20 % 8 != 0: left alone
                          nn.Linear(20, 8),                            # fp32 weights: left alone
                          nn.Sequential(nn.Linear(8, 16, dtype=torch.bfloat16, bias=False)))
    w0 = model[0].weight
    done = ops.replace_linear_modules(model)
    assert done == ["0", "4.0"]
    assert isinstance(model[0], ops.B200Linear) and model[0].weight is w0 and model[0].bias is not None
    assert isinstance(model[4][0], ops.B200Linear) and model[4][0].bias is None
    assert type(model[2]) is nn.Linear and type(model[3]) is nn.Linear
    assert set(dict(model.named_parameters())) == {"0.weight", "0.bias", "2.weight", "2.bias", "3.weight", "3.bias", "4.0.weight"}
    assert ops.replace_linear_modules(model) == []                     # idempotent
    with pytest.raises(capi.B200HgemmError):
        ops.B200Linear(60, 64)
