"""First hardware run of two kernel instantiations that only the full-size shapes reach (bench.py --workload
deepseek-block at the public V3 shape, Llama-3-70B under TP 4).  Sorted after the model-stack files: under the driver's
`-x` a failure here must not hide their results."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n1,n2", [(64, 112), (128, 128), (32, 48), (32, 64)])
def test_v3_shape_transform_configurations(n1, n2):
    """The Kronecker pairs of the public V3 shape that bench.py --workload deepseek-block reaches: dim 7168 -> (64, 112)
    (HMMA kernel, KT2 = 7 -- an instantiation the verified suite never ran; also Llama-70B's 28672 / 4), wo 16384 ->
    (128, 128) (tcgen05 <128, 128, 1>, likewise new), q_lora 1536 -> (32, 48) and moe_inter 2048 -> (32, 64) (both in the
    verified suite; repeated here so that the block's whole set is in one place).  Same checks as
    tests/test_gpu_parity.py::test_transform_quant_codes."""
    import test_gpu_parity as P
    from flatquant_b200 import ops
    P.test_transform_quant_codes(ops, torch.device("cuda:0"), n1, n2)
