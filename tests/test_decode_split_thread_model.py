"""Thread-level transcription of decode_attn_i4_split_kernel / decode_attn_merge_kernel (csrc/qattention.cuh) in
numpy: every thread's byte addresses into the paged cache (uint2 loads of 8 packed bytes per lane, nibble order), its
16-dim slice of q, the 8-lane shuffle reduction, the [d][lane8] shared-memory layout with its per-group padding, the
`jj` read index of the in-CTA merge and the partial-result layout read by the merge kernel -- transcribed line by line
from the CUDA source, on a real paged cache written in the reference's layout, against the oracle.  The kernel has not
run on hardware yet (tests/test_zzz_decode_split_gpu.py); this pins its index arithmetic."""
import numpy as np
import torch

from oracle import flatq_oracle as O

GROUPS, STRIDE = 16, 130


def emulate_split_kernel(q, kv_data, kv_param, indptr, indices, last, layer, q_heads, kv_heads, page_size, n_layers,
                         splits, GH, sm_scale):
    """q fp16 [B, HQ, 128] (numpy), kv_data u8 flat, kv_param fp16 flat (scale, zero pairs) -> o fp16 [B, HQ, 128]"""
    B, D = q.shape[0], 128
    G = q_heads // kv_heads
    head_groups = G // GH
    grp_stride = GH * 128 + 8
    part = np.full((B * q_heads * splits * STRIDE,), np.nan, np.float32)
    qflat = q.reshape(-1)
    for b in range(B):                                        # blockIdx.z
        for by in range(kv_heads * head_groups):              # blockIdx.y
            hk, hq0 = by // head_groups, (by // head_groups) * G + (by % head_groups) * GH
            for split in range(splits):                       # blockIdx.x
                pg0, pg1 = int(indptr[b]), int(indptr[b + 1])
                n_tok = (pg1 - pg0 - 1) * page_size + int(last[b])
                per = (n_tok + splits - 1) // splits
                t0, t1 = split * per, min(n_tok, split * per + per)
                n_iter = (t1 - t0 + GROUPS - 1) // GROUPS if t1 > t0 else 0
                s_m = np.zeros((GROUPS, GH), np.float32)
                s_l = np.zeros((GROUPS, GH), np.float32)
                s_acc = np.full((GROUPS * grp_stride,), np.nan, np.float32)
                # ---- per-thread state, threads 0..127: lane8 = tid & 7, grp = tid >> 3
                m = np.full((128, GH), -np.inf, np.float32)
                l = np.zeros((128, GH), np.float32)
                acc = np.zeros((128, GH, 16), np.float32)
                qv = np.zeros((128, GH, 16), np.float32)
                for tid in range(128):
                    lane8 = tid & 7
                    for g in range(GH):
                        base = (b * q_heads + hq0 + g) * D + 16 * lane8       # halves; qp[0] = 8 halves, qp[1] = next 8
                        lo, hi = qflat[base:base + 8], qflat[base + 8:base + 16]
                        for j in range(4):
                            qv[tid, g, 2 * j], qv[tid, g, 2 * j + 1] = lo[2 * j], lo[2 * j + 1]
                            qv[tid, g, 8 + 2 * j], qv[tid, g, 8 + 2 * j + 1] = hi[2 * j], hi[2 * j + 1]
                qv = (qv * np.float32(sm_scale)).astype(np.float32)

                def load(t, lane8):
                    page, off = int(indices[pg0 + t // page_size]), t % page_size
                    kidx = ((((page * n_layers + layer) * 2 + 0) * kv_heads + hk) * page_size + off)
                    vidx = kidx + kv_heads * page_size
                    kb = kv_data[kidx * 64 + 8 * lane8: kidx * 64 + 8 * lane8 + 8]       # uint2 at [lane8]
                    vb = kv_data[vidx * 64 + 8 * lane8: vidx * 64 + 8 * lane8 + 8]
                    return kb, vb, kv_param[2 * kidx: 2 * kidx + 2].astype(np.float32), kv_param[2 * vidx: 2 * vidx + 2].astype(np.float32)

                def dequant16(raw8, par):
                    x = np.zeros(16, np.float32)
                    rx = int.from_bytes(bytes(raw8[:4]), "little")
                    ry = int.from_bytes(bytes(raw8[4:]), "little")
                    for j in range(8):
                        x[j] = np.float32((rx >> (4 * j)) & 0xF) * par[0] - par[1]
                        x[8 + j] = np.float32((ry >> (4 * j)) & 0xF) * par[0] - par[1]
                    return x
                for it in range(n_iter):
                    dots = np.zeros((128, GH), np.float32)
                    vhat = np.zeros((128, 16), np.float32)
                    valid = np.zeros(128, bool)
                    for tid in range(128):
                        lane8, grp = tid & 7, tid >> 3
                        t = t0 + it * GROUPS + grp
                        valid[tid] = t < t1
                        if not valid[tid]:
                            continue   # (the kernel computes on stale registers here and discards the result)
                        kb, vb, kpar, vpar = load(t, lane8)
                        xh = dequant16(kb, kpar)
                        for g in range(GH):
                            dots[tid, g] = np.dot(qv[tid, g], xh)
                        vhat[tid] = dequant16(vb, vpar)
                    # __shfl_xor 4, 2, 1: every lane of an aligned 8-lane group ends with the group's sum
                    dots = dots.reshape(16, 8, GH).sum(1, keepdims=True).repeat(8, 1).reshape(128, GH)
                    for tid in range(128):
                        if not valid[tid]:
                            continue
                        for g in range(GH):
                            m_new = max(m[tid, g], dots[tid, g])
                            corr, pexp = np.exp(np.float32(m[tid, g] - m_new)), np.exp(np.float32(dots[tid, g] - m_new))
                            l[tid, g] = l[tid, g] * corr + pexp
                            acc[tid, g] = acc[tid, g] * corr + pexp * vhat[tid]
                            m[tid, g] = m_new
                for tid in range(128):
                    lane8, grp = tid & 7, tid >> 3
                    for g in range(GH):
                        if lane8 == 0:
                            s_m[grp, g], s_l[grp, g] = m[tid, g], l[tid, g]
                        for d in range(16):
                            s_acc[grp * grp_stride + g * 128 + d * 8 + lane8] = acc[tid, g, d]
                for j in range(128):                                   # thread j owns output dim j
                    jj = (j & 15) * 8 + (j >> 4)
                    for g in range(GH):
                        mm = s_m[:, g].max()
                        ll, out = np.float32(0), np.float32(0)
                        for r in range(GROUPS):
                            c = np.float32(0) if s_m[r, g] == -np.inf else np.exp(np.float32(s_m[r, g] - mm))
                            ll += s_l[r, g] * c
                            out += s_acc[r * grp_stride + g * 128 + jj] * c
                        dst = ((b * q_heads + hq0 + g) * splits + split) * STRIDE
                        part[dst + j] = out
                        if j == 0:
                            part[dst + 128], part[dst + 129] = mm, ll
    o = np.zeros((B * q_heads, 128), np.float16)
    for blk in range(B * q_heads):                                     # merge kernel
        src = part[blk * splits * STRIDE:(blk + 1) * splits * STRIDE].reshape(splits, STRIDE)
        mm = src[:, 128].max()
        c = np.where(src[:, 128] == -np.inf, 0, np.exp(src[:, 128] - mm)).astype(np.float32)
        ll = (src[:, 129] * c).sum()
        o[blk] = ((src[:, :128] * c[:, None]).sum(0) / ll).astype(np.float16)
    assert not np.isnan(part).any()     # every (query head, split) slot was written exactly by its CTA
    return o.reshape(B, q_heads, 128)


def test_thread_level_model_of_the_split_kernel_equals_the_oracle():
    torch.manual_seed(0)
    B, HQ, HKV, PS, L, layer, D = 2, 4, 2, 8, 2, 1, 128
    lens = [21, 5]
    pages_per = [(n + PS - 1) // PS for n in lens]
    n_pages = sum(pages_per) + 1
    kv_data = np.zeros((n_pages, L, 2, HKV, PS, D // 2), np.uint8)
    kv_param = np.zeros((n_pages, L, 2, HKV, PS, 2), np.float16)
    perm = torch.randperm(n_pages)[:sum(pages_per)].tolist()
    indptr, indices, last, ks, vs = [0], [], [], [], []
    for b, n in enumerate(lens):
        pg = perm[indptr[-1]:indptr[-1] + pages_per[b]]
        indices += pg
        indptr.append(indptr[-1] + pages_per[b])
        last.append(n - (pages_per[b] - 1) * PS)
        k, v = (torch.randn(n, HKV, D) * 1.5).half(), torch.randn(n, HKV, D).half()
        ks.append(k), vs.append(v)
        for sel, x in ((0, k), (1, v)):
            packed, s, z = O.kv_asym_quant(x)
            for t in range(n):
                kv_data[pg[t // PS], layer, sel, :, t % PS] = packed[t].numpy()
                kv_param[pg[t // PS], layer, sel, :, t % PS, 0] = s[t, :, 0].numpy()
                kv_param[pg[t // PS], layer, sel, :, t % PS, 1] = z[t, :, 0].numpy()
    q = torch.randn(B, HQ, D).half()
    for splits, GH in ((1, 2), (3, 2), (2, 1)):
        got = emulate_split_kernel(q.numpy(), kv_data.reshape(-1), kv_param.reshape(-1), indptr, indices, last, layer, HQ,
                                   HKV, PS, L, splits, GH, D ** -0.5)
        for b, n in enumerate(lens):
            kp, ksc, kz = O.kv_asym_quant(ks[b])
            vp, vsc, vz = O.kv_asym_quant(vs[b])
            r = lambda t: t.repeat_interleave(HQ // HKV, 1)[None]  # noqa: E731
            ref = O.decode_attention_i4(q[b:b + 1], r(kp), r(ksc), r(kz), r(vp), r(vsc), r(vz))[0].numpy()
            assert np.abs(got[b].astype(np.float32) - ref).max() < 5e-3, (splits, GH, b)
