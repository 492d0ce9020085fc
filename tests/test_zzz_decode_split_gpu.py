"""Split-KV decode attention (flatq_kv_decode_i4_split) on the B200 against the verified single-pass kernel
(flatq_kv_decode_i4) and the oracle, over the same paged INT4 cache.

First hardware run: round 2 (6 passed, profiles/r02_first_gpu_call.md); the split-KV kernel is the default of
ops.kv_decode_i4 since.  The algorithm is also restated in numpy (tests/test_decode_split_model.py)."""
import os

import pytest
import torch

from oracle import flatq_oracle as O

pytestmark = pytest.mark.gpu


def _cache(dev, lens, hkv, ps, n_layers, layer, seed):
    from flatquant_b200 import ops
    torch.manual_seed(seed)
    D = 128
    pages_per = [(n + ps - 1) // ps for n in lens]
    n_pages = sum(pages_per) + 2
    kv_data = torch.zeros(n_pages, n_layers, 2, hkv, ps, D // 2, dtype=torch.uint8, device=dev)
    kv_param = torch.zeros(n_pages, n_layers, 2, hkv, ps, 2, dtype=torch.float16, device=dev)
    perm = torch.randperm(n_pages)[:sum(pages_per)].tolist()
    indptr, indices, slots, last = [0], [], [], []
    for b, n in enumerate(lens):
        pg = perm[indptr[-1]:indptr[-1] + pages_per[b]]
        indices += pg
        indptr.append(indptr[-1] + pages_per[b])
        last.append(n - (pages_per[b] - 1) * ps)
        slots += [pg[t // ps] * ps + t % ps for t in range(n)]
    k = (torch.randn(sum(lens), hkv, D, device=dev) * 1.5).half()
    v = torch.randn(sum(lens), hkv, D, device=dev).half()
    slot_t = torch.tensor(slots, dtype=torch.int32, device=dev)
    ops.kv_quant(k, kv_data, kv_param, slot_t, layer, 0)
    ops.kv_quant(v, kv_data, kv_param, slot_t, layer, 1)
    t32 = lambda x: torch.tensor(x, dtype=torch.int32, device=dev)  # noqa: E731
    return kv_data, kv_param, t32(indptr), t32(indices), t32(last), k, v


@pytest.mark.parametrize("lens,hq,hkv,ps,hint", [
    ([37, 16, 5], 8, 2, 16, 0),            # GQA 4 -> 4 heads per CTA; ragged lengths, scattered pages, no length bound
    ([37, 16, 5], 8, 8, 16, 40),           # no GQA -> 1 head per CTA
    ([1], 6, 3, 4, 1),                     # GQA 2 -> 2 heads per CTA; a single token
    ([2048], 32, 8, 2048, 2048),           # Llama-3-8B decode, batch 1, one page per sequence
    ([2051] * 16, 32, 8, 2064, 2064),      # batch 16
    ([300, 1, 77, 300], 64, 8, 64, 300),   # Llama-3-70B head counts: GQA 8 -> two CTAs of 4 heads per kv head
])
def test_split_decode_equals_single_pass_kernel_and_oracle(lens, hq, hkv, ps, hint):
    from flatquant_b200 import ops
    dev = torch.device("cuda:0")
    layer = 1
    kv_data, kv_param, ip, ix, lp, k, v = _cache(dev, lens, hkv, ps, 2, layer, seed=len(lens) + hq)
    q = torch.randn(len(lens), hq, 128, device=dev).half()
    want = ops.kv_decode_i4(q, kv_data, kv_param, ip, ix, lp, layer=layer, split=False)
    got = ops.kv_decode_i4(q, kv_data, kv_param, ip, ix, lp, layer=layer, split=True, max_tokens_hint=hint)
    again = ops.kv_decode_i4(q, kv_data, kv_param, ip, ix, lp, layer=layer, split=True, max_tokens_hint=hint)
    torch.cuda.synchronize()
    assert torch.equal(got, again)   # no atomics, fixed merge order: run-to-run deterministic
    scale = max(1.0, float(want.float().abs().max()))
    err = float((got.float() - want.float()).abs().max())
    print(f"split vs single-pass kernel: max |diff| {err:.2e}")
    assert err < 2e-3 * scale        # same fp32 arithmetic, different summation order, fp16 outputs
    pos, rep = 0, hq // hkv
    for b, n in enumerate(lens[:3]):  # the oracle on the first sequences
        kp, ks, kz = O.kv_asym_quant(k[pos:pos + n].cpu())
        vp, vs, vz = O.kv_asym_quant(v[pos:pos + n].cpu())
        r = lambda t: t.repeat_interleave(rep, 1)[None]  # noqa: E731
        ref = O.decode_attention_i4(q[b:b + 1].cpu(), r(kp), r(ks), r(kz), r(vp), r(vs), r(vz))[0]
        assert (got[b].float().cpu() - ref).abs().max() < 5e-3
        pos += n
