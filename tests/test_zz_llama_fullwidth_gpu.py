"""A full-width decoder layer on the B200 against the reference's fake-quant forward ("Tighten and widen
parity"): ONE Llama-3-8B-shaped layer -- hidden 4096 -> (64,64) Kronecker pair, intermediate 14336 -> (112,128), 32 query /
8 kv heads x 128, 16 tokens -- built, calibrated-parameter-jittered, reparameterised, RTN-quantised and exported by the
reference's own code (tests/golden/make_golden_llama.py --full-width), loaded through from_pretrained and run on the CUDA
kernels: every quantised stage teacher-forced on the golden input of that stage (tcgen05 transform kernels at the real
factor sizes, left-only 32-head o_proj transform, right-only 128 x 128 K/Q transform, paged INT4 KV), then the whole
layer.  The 110 MB checkpoint is a git-ignored build artefact (tests/golden/llama_fullwidth_ckpt/, written by
__graft_entry__.build() in the build container, carried to the GPU box with the working tree); the expected activations
(tests/golden/llama_fullwidth.npz) are committed."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from test_llama_model_host import teacher_forced_stage_errors

pytestmark = pytest.mark.gpu
CKPT = os.path.join(GOLDEN, "llama_fullwidth_ckpt")

# vs the reference's fp32 run: measured on the B200 AND on the CPU model of the kernels' fp16 rounding points
# (tests/cpu_ops.py) max <= 2.8 % / mean <= 1.75 % of max|y| -- the price of fp16 activations at this width (the same
# layer in fp32 equals the reference to 2e-5: test_llama_model_host.py); bounds = 1.5x measured
FULLWIDTH_STAGE_TOL = (0.042, 0.026)
# vs that CPU rounding-point model, stage by stage on the same inputs: only the summation order inside the tensor cores
# differs (measured: max <= 0.9 %, mean <= 0.012 % of max|y| -- a handful of +-1 code flips per stage)
MODEL_TOL = (0.02, 0.0005)
# o_proj alone sits behind the causal attention, which is library code on both sides (torch SDPA: a flash kernel on the GPU,
# the math path on the CPU) and is not modelled rounding point by rounding point: its fp16-ulp differences move a few more
# activation codes of the o_proj input.  Measured 1.06 % / 0.070 % -- the same figures with the legacy HMMA transform, with
# PDL off, with the tcgen05 split-K GEMM instead of the weight-streaming one and with the torch glue (FLATQ_KRON_HMMA,
# FLATQ_NO_PDL, FLATQ_NO_SKINNY, FLATQ_FUSED_GLUE=0), i.e. independent of this repo's kernel variants.
MODEL_TOL_BEHIND_SDPA = (0.02, 0.0015)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def model(dev):
    if not os.path.exists(os.path.join(CKPT, "model.safetensors")):
        pytest.skip("tests/golden/llama_fullwidth_ckpt/ not built (python tests/golden/make_golden_llama.py --full-width, "
                    "build container only)")
    from flatquant_b200 import _lib
    from flatquant_b200.deploy.transformers import FlatQuantLlamaForCausalLM
    _lib.load()
    return FlatQuantLlamaForCausalLM.from_pretrained(CKPT, device=dev)


@pytest.fixture(scope="module")
def g():
    z = np.load(os.path.join(GOLDEN, "llama_fullwidth.npz"))
    return {k: (z[k].astype(np.float32) if z[k].dtype == np.float16 else z[k]) for k in z.files}


def test_full_width_layer_every_stage_teacher_forced(model, g, dev):
    cfg = model.config
    assert (cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads, cfg.num_key_value_heads) == (4096, 14336, 32, 8)
    got = {}
    errs = teacher_forced_stage_errors(model, g, dev, outputs=got)
    for key, (mx, mean) in errs.items():
        print(key, f"vs reference fp32: max {mx:.4f} mean {mean:.5f}")
    for key, (mx, mean) in errs.items():
        assert mx < FULLWIDTH_STAGE_TOL[0] and mean < FULLWIDTH_STAGE_TOL[1], (key, mx, mean)
    # the same stages on the CPU model of the kernels' arithmetic (oracle-backed ops, fp16 rounding points)
    import cpu_ops
    from flatquant_b200.deploy.transformers import FlatQuantLlamaForCausalLM
    mp = pytest.MonkeyPatch()
    try:
        cpu_ops.install(mp)
        cpu_model = FlatQuantLlamaForCausalLM.from_pretrained(CKPT)
        want = {}
        teacher_forced_stage_errors(cpu_model, g, torch.device("cpu"), outputs=want)
    finally:
        mp.undo()
    for key in got:
        d = np.abs(got[key] - want[key])
        mx, mean = float(d.max() / np.abs(want[key]).max()), float(d.mean() / np.abs(want[key]).max())
        print(key, f"vs CPU rounding-point model: max {mx:.4f} mean {mean:.6f}")
        tol = MODEL_TOL_BEHIND_SDPA if key[1] == "o_proj" else MODEL_TOL
        assert mx < tol[0] and mean < tol[1], (key, mx, mean)


def test_full_width_layer_end_to_end(model, g, dev):
    """The whole layer + lm_head from the token ids: seven chained W4A4 stages (code flips of one stage move the inputs of
    the next), compared with the reference's logits by cosine and relative error."""
    ids = torch.from_numpy(g["ids"]).to(dev)
    with torch.no_grad():
        logits, cache = model(ids)
    torch.cuda.synchronize()
    ref = torch.from_numpy(g["prefill_logits"]).to(dev)
    cos = float(torch.nn.functional.cosine_similarity(logits.float().flatten(), ref.flatten(), dim=0))
    rel = float((logits.float() - ref).abs().max() / ref.abs().max())
    print("full-width logits: cosine", cos, "max rel err", rel)
    assert cos > 0.97 and rel < 0.3   # measured 0.981 / 0.17: seven chained W4A4 stages in fp16 vs the fp32 reference
    assert cache.get_seq_length() == ids.shape[1]
