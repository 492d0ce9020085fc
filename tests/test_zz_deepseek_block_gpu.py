"""BASELINE.json configs[4] on the B200: the real-quant DeepSeek-V3 block (MLA + MoE expert FlatLinears) on the CUDA
kernels against the reference's fake-quant block (fixture: tests/golden/make_golden_deepseek.py).

Sorts last on purpose (first hardware run is the round-end run; see tests/test_zz_llama_model_gpu.py).  The host logic
is checked exactly (fp32, oracle-backed ops) in tests/test_deepseek_block_host.py."""
import pytest
import torch

from test_deepseek_block_host import check_stage_errors, load_fixture, stage_errors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from flatquant_b200 import _lib
    from flatquant_b200.deploy.transformers import modeling_deepseek as MD
    _lib.load()
    g, args, params = load_fixture()
    dev = torch.device("cuda:0")
    return g, args, MD.FlatQuantDeepSeekBlock(args, params, dev), dev, MD


def test_block_stages_teacher_forced(setup):
    g, args, block, dev, MD = setup
    check_stage_errors(stage_errors(block, g, dev))


def test_block_prefill_and_decode(setup):
    g, args, block, dev, MD = setup
    x = torch.from_numpy(g["x"]).to(dev).half()
    s = x.shape[1]
    cos, sin = MD.precompute_freqs(args, s, dev)
    with torch.no_grad():
        y = block(x, 0, cos, sin)
        block.kv_cache.zero_()
        block.pe_cache.zero_()
        block(x[:, :-1], 0, cos[:-1], sin[:-1])
        last = block(x[:, -1:], s - 1, cos[-1:], sin[-1:])
    torch.cuda.synchronize()
    assert y.dtype == torch.float16 and torch.isfinite(y).all()
    ref = torch.from_numpy(g["y"]).to(dev)
    c = torch.nn.functional.cosine_similarity(y.float().flatten(), ref.flatten(), dim=0)
    print("block output cosine vs reference fake-quant", float(c))
    # the CPU model of the kernels' rounding gives 0.979 (tests/test_deepseek_block_host.py); another realisation of the
    # same W4A4 code-flip noise moves it by a few 1e-3 -- the stage-by-stage bounds above are the parity check
    assert c > 0.95
    # same kernels, same inputs up to the batch shape: the decode step reproduces the prefill's last position closely
    # (different M takes different GEMM kernels, which are bit-identical; the transform kernels differ in token grouping)
    c2 = torch.nn.functional.cosine_similarity(last.float().flatten(), y[:, -1:].float().flatten(), dim=0)
    print("decode vs prefill last position cosine", float(c2))
    assert c2 > 0.97
