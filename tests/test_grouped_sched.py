"""The tile schedule of the grouped (mixture-of-experts) GEMM, csrc/grouped_sched.h -- the same header the CUDA kernel
compiles -- built for the host with g++ and checked against a Python model: for random routings (empty experts, one
expert taking everything, ragged counts) the tiles cover every (expert, token row, feature block) exactly once, never
cross an expert boundary in their valid rows, and their number equals the device-built prefix."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = tmp_path_factory.mktemp("grouped") / "libgrouped_sched_shim.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(out),
                           os.path.join(ROOT, "tests", "native", "grouped_sched_shim.cpp")])
    lib = ctypes.CDLL(str(out))
    ip = ctypes.POINTER(ctypes.c_int)
    lib.shim_build_prefix.argtypes = [ip, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip]
    lib.shim_tile.argtypes = [ip, ip, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip]
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _check(shim, counts, tn, n, block_w=128):
    counts = np.asarray(counts, np.int32)
    e = len(counts)
    off = np.zeros(e + 1, np.int32)
    off[1:] = np.cumsum(counts)
    nb_w = (n + block_w - 1) // block_w
    prefix = np.full(e + 1, -7, np.int32)
    shim.shim_build_prefix(_ptr(off), e, tn, nb_w, _ptr(prefix))
    want = np.concatenate([[0], np.cumsum(nb_w * ((counts + tn - 1) // tn))]).astype(np.int32)
    assert np.array_equal(prefix, want)
    total = int(prefix[-1])
    assert total <= nb_w * ((int(off[-1]) + tn - 1) // tn + e)  # the host-side bound used for the grid / t_upper
    covered = np.zeros((int(off[-1]), nb_w), np.int32)
    out = np.zeros(4, np.int32)
    for t in range(total):
        shim.shim_tile(_ptr(off), _ptr(prefix), e, tn, block_w, t, _ptr(out))
        g, m0, m_end, n0 = (int(v) for v in out)
        assert 0 <= g < e and counts[g] > 0
        assert off[g] <= m0 < m_end == off[g + 1] and (m0 - off[g]) % tn == 0
        assert n0 % block_w == 0 and 0 <= n0 < n
        covered[m0:min(m0 + tn, m_end), n0 // block_w] += 1
    assert (covered == 1).all()


def test_grouped_schedule_covers_every_row_once(shim):
    rng = np.random.default_rng(0)
    _check(shim, [5], 16, 256)
    _check(shim, [0, 0, 7, 0], 16, 100)
    _check(shim, [300, 0, 0, 1], 64, 4096)
    _check(shim, [1] * 9, 16, 130)
    for _ in range(40):
        e = int(rng.integers(1, 40))
        counts = rng.integers(0, 200, e) * (rng.random(e) < 0.7)
        if counts.sum() == 0:
            counts[0] = 3
        _check(shim, counts, int(rng.choice([16, 32, 64, 128, 256])), int(rng.choice([64, 96, 128, 200, 1024, 4096])))


def test_grouped_schedule_v3_shape(shim):
    """DeepSeek-V3: 256 experts, 2048 tokens x top-8 skewed routing, w13 N = 4096 and w2 N = 7168."""
    rng = np.random.default_rng(1)
    p = rng.dirichlet(np.full(256, 0.3))
    counts = np.bincount(rng.choice(256, 2048 * 8, p=p), minlength=256)
    _check(shim, counts, 64, 4096)
    _check(shim, counts, 128, 7168)


def test_grouped_gemm_tile_level_model_equals_per_expert_gemms(shim):
    """Tile-level transcription of the GROUPED branches of gemm_w4a8_kernel (csrc/gemm_w4a8.cuh): per output tile the X
    TMA box rows [m0, m0 + TN) of the sorted activation (rows of the NEXT expert included, rows past M zero-filled), the
    W TMA box rows [group * N + n0, + 128) of the stacked weight (rows of the next expert included), the epilogue's
    masks (n < N, m < m_end) and scale offsets (ch0 = group * N) -- against one plain GEMM per expert."""
    rng = np.random.default_rng(7)
    for counts, n, k, tn in (([5, 0, 19, 2], 100, 64, 16), ([40, 3, 0, 0, 70], 200, 32, 32), ([1] * 6, 130, 64, 16)):
        counts = np.asarray(counts, np.int32)
        e, m = len(counts), int(counts.sum())
        off = np.zeros(e + 1, np.int32)
        off[1:] = np.cumsum(counts)
        nb_w = (n + 127) // 128
        prefix = np.zeros(e + 1, np.int32)
        shim.shim_build_prefix(_ptr(off), e, tn, nb_w, _ptr(prefix))
        x = rng.integers(-8, 8, (m, k)).astype(np.int64)
        w = rng.integers(-8, 8, (e * n, k)).astype(np.int64)          # stacked [E * N, K]
        sx, sw = rng.random(m) + 0.5, rng.random(e * n) + 0.5
        y = np.full((m, n), np.nan)
        out = np.zeros(4, np.int32)
        for t in range(int(prefix[-1])):
            shim.shim_tile(_ptr(off), _ptr(prefix), e, tn, 128, t, _ptr(out))
            g, m0, m_end, n0 = (int(v) for v in out)
            m_end = min(m_end, m)
            w_row0, ch0 = g * n + n0, g * n
            xt = np.zeros((tn, k), np.int64)
            rows = x[m0:m0 + tn]                                        # TMA box; out-of-bounds rows are zero-filled
            xt[:len(rows)] = rows
            wt = np.zeros((128, k), np.int64)
            wrows = w[w_row0:w_row0 + 128]
            wt[:len(wrows)] = wrows
            acc = wt @ xt.T                                             # D[n, m] of the tile
            for r in range(128):
                nn = n0 + r
                if nn >= n:
                    continue                                            # n_ok
                for j in range(tn):
                    mm = m0 + j
                    if mm < m_end:
                        assert np.isnan(y[mm, nn])                      # written exactly once
                        y[mm, nn] = acc[r, j] * sx[mm] * sw[ch0 + nn]
        want = np.zeros((m, n))
        for g in range(e):
            r = slice(off[g], off[g + 1])
            want[r] = (x[r] @ w[g * n:(g + 1) * n].T) * sx[r, None] * sw[None, g * n:(g + 1) * n]
        assert not np.isnan(y).any() and np.array_equal(y, want)
