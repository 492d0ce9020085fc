"""Run the BODIES of the model-stack `-m gpu` tests on the CPU: flatquant_b200.ops is replaced by the oracle-backed
stand-ins of tests/cpu_ops.py (STRICT: the C ABI's own argument checks) and device = cpu.  This does not test a kernel;
it proves that the test code and the host code it drives are executable and only make calls the CUDA library accepts,
before GPU minutes are spent on them.  (Graph capture and the kernels' numerics are what remains for the hardware run.)

    python tests/tools/dry_run_gpu_tests_on_cpu.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import cpu_ops  # noqa: E402
from flatquant_b200 import ops  # noqa: E402

for name in cpu_ops.FUNCTIONS:
    setattr(ops, name, getattr(cpu_ops, name))
torch.cuda.synchronize = lambda *a, **k: None
dev = torch.device("cpu")

m = importlib.import_module("test_zz_llama_model_gpu")
from flatquant_b200.deploy.transformers import FlatQuantLlamaForCausalLM  # noqa: E402
from flatquant_b200.deploy.transformers import modeling_deepseek as MD  # noqa: E402

model = FlatQuantLlamaForCausalLM.from_pretrained(m.CKPT, device=dev)
g = np.load(os.path.join(m.GOLDEN, "llama_model.npz"))
m.test_right_only_transform_and_matmul_kv_cache(dev)
m.test_every_quantised_stage_teacher_forced(model, g, dev)
m.test_prefill_then_decode_end_to_end(model, g, dev)
m.test_decode_attention_over_paged_int4_cache(model, g, dev)
print("[dry run] tests/test_zz_llama_model_gpu.py (graph test excluded): ok")

d = importlib.import_module("test_zz_deepseek_block_gpu")
from test_deepseek_block_host import load_fixture  # noqa: E402

gg, args, params = load_fixture()
setup = (gg, args, MD.FlatQuantDeepSeekBlock(args, params, dev), dev, MD)
d.test_block_stages_teacher_forced(setup)
d.test_block_prefill_and_decode(setup)
print("[dry run] tests/test_zz_deepseek_block_gpu.py: ok")

z = importlib.import_module("test_zzz_grouped_gemm_gpu")
z.test_grouped_moe_equals_loop_on_gpu(dev)
z._case(dev, [12, 0, 40, 7], 256, 512, out_dtype=torch.bfloat16, with_bias=True, seed=3)
print("[dry run] tests/test_zzz_grouped_gemm_gpu.py (equal-to-loop parts): ok")

gl = importlib.import_module("test_zzz_glue_kernels_gpu")
gl.test_rmsnorm_matches_hf_expression(dev, 5, 512)
gl.test_rope_is_bit_identical_to_the_torch_expression(dev, 2, 9, 4)
gl.test_silu_mul_matches_the_torch_expression(dev, 3, 64)
print("[dry run] tests/test_zzz_glue_kernels_gpu.py: ok")
