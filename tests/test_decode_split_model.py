"""The split-KV decode attention (csrc/qattention.cuh: decode_attn_i4_split_kernel + decode_attn_merge_kernel) restated
step for step in numpy -- the same partition of the sequence into `splits` ranges, 16 token groups per range each with
its own online-softmax state, the in-CTA merge and the cross-split merge, including empty splits and groups (-inf / 0
states) -- against the oracle's direct softmax.  This checks the ALGORITHM the kernel implements (the CUDA code itself
first runs in tests/test_zzz_decode_split_gpu.py)."""
import numpy as np
import pytest
import torch

from oracle import flatq_oracle as O

GROUPS = 16


def _online(scores, values):
    """online softmax over a token stream, as one 8-lane group does it: returns (m, l, acc)"""
    m, l, acc = -np.inf, np.float32(0), np.zeros(values.shape[1], np.float32)
    for s, v in zip(scores, values):
        m_new = max(m, s)
        corr, p = np.exp(np.float32(m - m_new)), np.exp(np.float32(s - m_new))
        l = l * corr + p
        acc = acc * corr + p * v
        m = m_new
    return np.float32(m), np.float32(l), acc


def _merge(states):
    """(m, l, acc) list -> merged state; a state with m = -inf contributes nothing (the kernel's `c = 0`)"""
    mm = max(s[0] for s in states)
    ll, out = np.float32(0), np.zeros_like(states[0][2])
    for m, l, acc in states:
        c = np.float32(0) if m == -np.inf else np.exp(np.float32(m - mm))
        ll += l * c
        out += acc * c
    return np.float32(mm), ll, out


def split_decode_model(q, k_hat, v_hat, splits, sm_scale):
    """q [D], k_hat / v_hat [T, D] (dequantised) -> o [D], following the kernel's schedule."""
    n_tok = k_hat.shape[0]
    per = (n_tok + splits - 1) // splits
    scores = (k_hat @ (q * sm_scale)).astype(np.float32)
    parts = []
    for s in range(splits):
        t0, t1 = s * per, min(n_tok, s * per + per)
        groups = []
        for r in range(GROUPS):
            idx = np.arange(t0 + r, t1, GROUPS) if t1 > t0 else np.arange(0)
            groups.append(_online(scores[idx], v_hat[idx]))
        parts.append(_merge(groups))
    mm, ll, out = _merge(parts)
    return out / ll if ll > 0 else np.zeros_like(out)


@pytest.mark.parametrize("n_tok,splits", [(1, 1), (5, 4), (16, 1), (17, 2), (37, 64), (300, 7), (2048, 32)])
def test_split_schedule_equals_direct_softmax(n_tok, splits):
    torch.manual_seed(n_tok + splits)
    D = 128
    k = (torch.randn(1, n_tok, 1, D) * 1.5).half()
    v = torch.randn(1, n_tok, 1, D).half()
    q = torch.randn(1, 1, D).half()
    kp, ks, kz = O.kv_asym_quant(k)
    vp, vs, vz = O.kv_asym_quant(v)
    want = O.decode_attention_i4(q, kp, ks, kz, vp, vs, vz)[0, 0].numpy()
    nib = lambda p: O.unpack_u4(p.numpy()) if hasattr(O, "unpack_u4") else np.stack(  # noqa: E731
        [p.numpy() & 0xF, p.numpy() >> 4], -1).reshape(*p.shape[:-1], -1)
    k_hat = (nib(kp)[0, :, 0].astype(np.float32) * ks[0, :, 0].float().numpy() - kz[0, :, 0].float().numpy())
    v_hat = (nib(vp)[0, :, 0].astype(np.float32) * vs[0, :, 0].float().numpy() - vz[0, :, 0].float().numpy())
    got = split_decode_model(q[0, 0].float().numpy(), k_hat, v_hat, splits, D ** -0.5)
    # (1) the schedule is exact: direct softmax over the same fp32-dequantised rows, in float64
    qd, kd, vd = q[0, 0].double().numpy(), k_hat.astype(np.float64), v_hat.astype(np.float64)
    logits = kd @ qd / np.sqrt(D)
    pr = np.exp(logits - logits.max())
    direct = (pr / pr.sum()) @ vd
    assert np.abs(got - direct).max() < 2e-5 * max(1.0, np.abs(direct).max())
    # (2) and within the GPU tests' tolerance of the oracle, which dequantises in half precision like the reference
    assert np.abs(got - want).max() < 5e-3 * max(1.0, np.abs(want).max())
