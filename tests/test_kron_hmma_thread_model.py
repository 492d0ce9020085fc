"""Warp-level transcription of kron_quant_mma_kernel's transform (csrc/kron_quant.cuh) in numpy: the zero-padded
shared-memory tiles, every lane's ldmatrix address, the m16n8k16 fragment layouts (A row-major / B column-major / C),
the fp16 rounding of the intermediate and of the result, and the z_out store addresses -- transcribed line by line from
the CUDA source.  Configurations the GPU parity suite has run (two-factor incl. padded n1, left-only) validate the
transcription itself; the others are what it is for: the explicit-identity forms the model stacks use ((4,128) + I for
o_proj, I + (g,128) for the K/Q cache transform) and the right-only configuration, which has not run on hardware yet."""
import numpy as np
import pytest
import torch


def _ldsm_x4(addr, trans):
    """addr(lane) -> (array2d, row, col): lane 8i+j supplies row j of 8x8 matrix i.  Returns regs[lane][i] = 2 halves."""
    mats = []
    for i in range(4):
        rows = []
        for j in range(8):
            arr, r, c = addr(8 * i + j)
            rows.append(arr[r, c:c + 8])
        mats.append(np.stack(rows))
    regs = np.zeros((32, 4, 2), np.float16)
    for lane in range(32):
        g, tig = lane >> 2, lane & 3
        for i in range(4):
            regs[lane, i] = (mats[i][2 * tig, g], mats[i][2 * tig + 1, g]) if trans else mats[i][g, 2 * tig:2 * tig + 2]
    return regs


def _mma(c, a, b0, b1):
    """mma.sync.m16n8k16.row.col: c [32][4] fp32 += A(16x16 from a [32][4][2]) . B(16x8 from b0, b1 [32][2])"""
    A, B = np.zeros((16, 16), np.float32), np.zeros((16, 8), np.float32)
    for lane in range(32):
        g, tig = lane >> 2, lane & 3
        A[g, 2 * tig:2 * tig + 2], A[g + 8, 2 * tig:2 * tig + 2] = a[lane, 0], a[lane, 1]
        A[g, 2 * tig + 8:2 * tig + 10], A[g + 8, 2 * tig + 8:2 * tig + 10] = a[lane, 2], a[lane, 3]
        B[2 * tig:2 * tig + 2, g], B[2 * tig + 8:2 * tig + 10, g] = b0[lane], b1[lane]
    C = A @ B
    for lane in range(32):
        g, tig = lane >> 2, lane & 3
        c[lane, 0:2] += C[g, 2 * tig:2 * tig + 2]
        c[lane, 2:4] += C[g + 8, 2 * tig:2 * tig + 2]
    return c


def emulate_kron_transform(x, left, right, n1, n2):
    """x fp16 [tokens, n1*n2] -> z fp16, one token per CTA iteration (tokens_per_cta only batches independent tokens)."""
    tokens, K = x.shape
    P2, KT2 = n2, n2 // 16
    P1 = (n1 + 15) // 16 * 16
    LDL, LDX, wpt = P1 + 8, P2 + 8, P1 // 16
    z = np.full((tokens, K), np.nan, np.float16)
    Ls = Rs = None
    if left is not None:
        Ls = np.zeros((P1, LDL), np.float16)
        Ls[:n1, :n1] = left
    if right is not None:
        Rs = np.zeros((P2, LDX), np.float16)
        Rs[:, :P2] = right
    for token in range(tokens):
        Xt = np.full((P1, LDX), np.nan, np.float16)          # pad columns stay unwritten; pad rows are zeroed
        Xt[:, :P2] = 0
        Xt[:n1, :P2] = x[token].reshape(n1, P2)
        for slab in range(wpt):                               # one warp per 16-row slab
            m0 = slab * 16
            mi, r8 = np.arange(32) >> 3, np.arange(32) & 7
            ufrag = []
            for kt2 in range(KT2):
                if left is not None:
                    c0, c1 = np.zeros((32, 4), np.float32), np.zeros((32, 4), np.float32)
                    for k0 in range(0, P1, 16):
                        a = _ldsm_x4(lambda l: (Ls, k0 + (mi[l] >> 1) * 8 + r8[l], m0 + (mi[l] & 1) * 8), True)
                        b = _ldsm_x4(lambda l: (Xt, k0 + (mi[l] & 1) * 8 + r8[l], kt2 * 16 + (mi[l] >> 1) * 8), True)
                        c0 = _mma(c0, a, b[:, 0], b[:, 1])
                        c1 = _mma(c1, a, b[:, 2], b[:, 3])
                    u = np.zeros((32, 4, 2), np.float16)
                    u[:, 0], u[:, 1], u[:, 2], u[:, 3] = c0[:, 0:2], c0[:, 2:4], c1[:, 0:2], c1[:, 2:4]   # pack_h2: fp16 rounding
                else:
                    u = _ldsm_x4(lambda l: (Xt, m0 + (mi[l] & 1) * 8 + r8[l], kt2 * 16 + (mi[l] >> 1) * 8), False)
                ufrag.append(u)
            zf = np.zeros((2 * KT2, 32, 2, 2), np.float16)     # z[nt][lane][h] = pair of halves
            if right is not None:
                for nt2 in range(KT2):
                    c0, c1 = np.zeros((32, 4), np.float32), np.zeros((32, 4), np.float32)
                    for kt in range(KT2):
                        b = _ldsm_x4(lambda l: (Rs, kt * 16 + (mi[l] & 1) * 8 + r8[l], nt2 * 16 + (mi[l] >> 1) * 8), True)
                        c0 = _mma(c0, ufrag[kt], b[:, 0], b[:, 1])
                        c1 = _mma(c1, ufrag[kt], b[:, 2], b[:, 3])
                    zf[2 * nt2, :, 0], zf[2 * nt2, :, 1] = c0[:, 0:2], c0[:, 2:4]
                    zf[2 * nt2 + 1, :, 0], zf[2 * nt2 + 1, :, 1] = c1[:, 0:2], c1[:, 2:4]
            else:
                for kt in range(KT2):
                    zf[2 * kt, :, 0], zf[2 * kt, :, 1] = ufrag[kt][:, 0], ufrag[kt][:, 1]
                    zf[2 * kt + 1, :, 0], zf[2 * kt + 1, :, 1] = ufrag[kt][:, 2], ufrag[kt][:, 3]
            for lane in range(32):
                g, tig = lane >> 2, lane & 3
                for h, row in ((0, m0 + g), (1, m0 + g + 8)):
                    if row < n1:
                        for nt in range(2 * KT2):
                            o = row * P2 + nt * 8 + 2 * tig
                            z[token, o:o + 2] = zf[nt, lane, h]
    assert not np.isnan(z.astype(np.float32)).any()            # every output element was stored
    return z


def _expected(x, left, right, n1, n2):
    xm = x.astype(np.float32).reshape(-1, n1, n2)
    u = xm if left is None else np.einsum("ij,tik->tjk", left.astype(np.float32), xm).astype(np.float16).astype(np.float32)
    zz = u if right is None else (u @ right.astype(np.float32))
    return zz.reshape(x.shape[0], -1).astype(np.float16) if right is not None else u.reshape(x.shape[0], -1).astype(np.float16)


@pytest.mark.parametrize("n1,n2,has_l,has_r,note", [
    (8, 16, True, True, "two-factor, padded n1: run on hardware (validates the transcription)"),
    (32, 48, True, True, "two-factor, two slabs: run on hardware"),
    (32, 128, True, False, "left-only: run on hardware (block_matmul)"),
    (4, 128, True, True, "o_proj of a 4-head model with an explicit identity right factor (zz Llama test)"),
    (16, 128, True, True, "K / Q cache transform with an explicit identity left factor (zz Llama test)"),
    (8, 128, True, True, "same, 8 rows per group (decode)"),
    (16, 128, False, True, "right-only: not yet run on hardware"),
    (1, 128, False, True, "right-only, one row per group"),
])
def test_hmma_transform_transcription(n1, n2, has_l, has_r, note):
    torch.manual_seed(n1 * 1000 + n2)
    tokens = 2
    x = torch.randn(tokens, n1 * n2).half().numpy()
    left = (torch.randn(n1, n1) / n1 ** 0.5).half().numpy() if has_l else None
    right = (torch.randn(n2, n2) / n2 ** 0.5).half().numpy() if has_r else None
    if "identity right" in note:
        right = np.eye(n2, dtype=np.float16)
    if "identity left" in note or "8 rows per group" in note:
        left = np.eye(n1, dtype=np.float16)
    got = emulate_kron_transform(x, left, right, n1, n2)
    want = _expected(x, left, right, n1, n2)
    d = np.abs(got.astype(np.float32) - want.astype(np.float32))
    # identical rounding points; only the fp32 summation order inside a product differs -> at most one fp16 ulp, rarely
    assert d.max() <= 2.0 ** -9 * max(1.0, np.abs(want.astype(np.float32)).max()), note
    assert (d > 0).mean() < 0.02, note
