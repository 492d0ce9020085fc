"""Tensor-parallel host logic on CPU: world_size 2 over gloo; the per-rank compute is the oracle (tests may use it)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flatquant_b200 import tp as TP
from oracle import flatq_oracle as O


def _oracle_linear(x, left, right, n1, n2, cmax_raw, cmin_raw, w_codes, s_w, bias, out_dtype):
    """Same signature as ops.linear_fused, but clip factors are the raw (pre-sigmoid) ones and W is unpacked codes."""
    r = O.flat_linear_int(x.float(), left, right, cmax_raw, cmin_raw, w_codes.numpy(), s_w.numpy(),
                          None if bias is None else bias.numpy())
    return torch.from_numpy(r["y"]).to(out_dtype)


def _problem(K, N, M, tp, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g)
    w = torch.randint(-8, 8, (N, K), generator=g, dtype=torch.int8)
    s_w = torch.rand(N, generator=g) * 0.02 + 0.01
    bias = torch.randn(N, generator=g)
    facs = {}
    for name, kk in (("full", K), ("shard", K // tp)):
        n1, n2 = TP.get_decompose_dim(kk)
        facs[name] = [(torch.randn(n1, n1, generator=g) / n1 ** 0.5, torch.randn(n2, n2, generator=g) / n2 ** 0.5)
                      for _ in range(tp)]
    return x, w, s_w, bias, facs


def _worker(rank, world, port, K, N, M, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, w, s_w, bias, facs = _problem(K, N, M, world)
        # row parallel: own K shard, own factor pair, own clip factors; one all-reduce; bias after the reduction
        sh = TP.row_shard(K, N, world, rank)
        L, R = facs["shard"][rank]
        clip = (4.0 - 0.5 * rank, 3.0 + 0.25 * rank)
        row = TP.RowParallelFlatLinear(sh, L, R, clip, w[:, sh.k0:sh.k1].contiguous(), s_w, bias, linear_fn=_oracle_linear)
        y_row = row(x[:, sh.k0:sh.k1].contiguous(), out_dtype=torch.float32)
        # column parallel: replicated transform, own rows of W, all-gather
        shc = TP.column_shard(K, N, world, rank)
        Lf, Rf = facs["full"][0]
        col = TP.ColumnParallelFlatLinear(shc, Lf, Rf, (4.0, 4.0), w[shc.n0:shc.n1].contiguous(), s_w[shc.n0:shc.n1],
                                          bias[shc.n0:shc.n1], linear_fn=_oracle_linear)
        y_col = col(x, out_dtype=torch.float32, gather=True)
        if rank == 0:
            ret["row"], ret["col"] = y_row.numpy(), y_col.numpy()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_plans():
    s = TP.row_shard(28672, 8192, 8, 3)
    assert (s.k0, s.k1) == (3 * 3584, 4 * 3584) and s.factors == (56, 64)  # 28672/8 -> (56, 64)
    c = TP.column_shard(8192, 8192 + 2 * 1024, 8, 7)
    assert (c.n0, c.n1) == (7 * 1280, 8 * 1280) and c.factors == (64, 128)
    with pytest.raises(ValueError):
        TP.row_shard(4096 + 64, 4096, 8, 0)
    w = torch.arange(16 * 64, dtype=torch.uint8).reshape(16, 64)  # packed [N=16, K/2=64]
    assert torch.equal(TP.split_packed_rows(w, TP.row_shard(128, 16, 2, 1)), w[:, 32:])
    assert torch.equal(TP.split_packed_rows(w, TP.column_shard(128, 16, 2, 1)), w[8:])


def test_row_and_column_parallel_world2():
    K, N, M, world = 512, 96, 9, 2
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), K, N, M, ret), nprocs=world, join=True)
    x, w, s_w, bias, facs = _problem(K, N, M, world)
    # serial restatement of the same sharded arithmetic
    want = np.zeros((M, N), np.float32)
    for r in range(world):
        sh = TP.row_shard(K, N, world, r)
        L, R = facs["shard"][r]
        want += O.flat_linear_int(x[:, sh.k0:sh.k1].contiguous(), L, R, 4.0 - 0.5 * r, 3.0 + 0.25 * r,
                                  w[:, sh.k0:sh.k1].numpy(), s_w.numpy())["y"]
    want += bias.numpy()[None, :]
    assert np.array_equal(ret["row"], want)  # fp32 a+b of two partials is order independent
    Lf, Rf = facs["full"][0]
    full = O.flat_linear_int(x, Lf, Rf, 4.0, 4.0, w.numpy(), s_w.numpy(), bias.numpy())["y"]
    assert np.array_equal(ret["col"], full)


def _reducer_worker(rank, world, port, ret):
    """The reducer interface bench.py and the model stack use (tp.make_reducer): on CPU / gloo it is the NCCL-style
    reducer (the peer-memory one needs CUDA symmetric memory), driving the oracle-backed GEMM stand-in."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import cpu_ops
    from flatquant_b200 import ops
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for name in cpu_ops.FUNCTIONS:
            setattr(ops, name, getattr(cpu_ops, name))
        M, N, K = 8, 32, 128
        g = torch.Generator().manual_seed(10 + rank)
        xq = torch.randint(-8, 8, (M, K), generator=g, dtype=torch.int8)
        sx = torch.rand(M, generator=g) + 0.5
        codes = torch.randint(-8, 8, (N, K), generator=g, dtype=torch.int8)
        w = torch.from_numpy(O.pack_i4(codes.numpy()))
        sw = torch.rand(N, generator=g) * 0.01
        bias = torch.arange(N).half()
        red = TP.make_reducer(world, rank, torch.device("cpu"), (M, N), torch.float16)
        assert isinstance(red, TP.NcclReducer) and red.partial_buffer().shape == (M, N)
        y = red.rowparallel_gemm(xq, sx, w, sw, bias)
        part = torch.from_numpy(O.sym_dequant(O.gemm_i4(xq.numpy(), codes.numpy()), sx.numpy(), sw.numpy())).half()
        ret[rank] = (y.float().numpy(), part.float().numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_reducer_interface_world2():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_reducer_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    want = (torch.from_numpy(ret[0][1]).half() + torch.from_numpy(ret[1][1]).half()) + torch.arange(32).half()
    for r in range(world):
        assert np.array_equal(ret[r][0], want.float().numpy()), r   # same sum, bias once, on every rank


def _capacity_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flatquant_b200.deploy.transformers.modeling_llama import TensorParallel
        tp = TensorParallel(world, rank)
        dev = torch.device("cpu")
        r1 = tp.reducer(5, 64, dev)                 # first use: sized for max(m, 2048) rows
        r2 = tp.reducer(2048, 64, dev)              # fits: the same object
        r3 = tp.reducer(3000, 64, dev)              # grows (a collective: both ranks get here with the same shapes)
        r4 = tp.reducer(7, 32, dev)                 # another width: another reducer
        caps = (r1.partial_buffer().shape[0], r3.partial_buffer().shape[0], r4.partial_buffer().shape[0])
        # the reducer sums a smaller call's rows only
        buf = r3.partial_buffer()
        buf.zero_()
        part = buf[:4]
        part.fill_(float(rank + 1))
        out = r3.all_reduce(part)
        ret[rank] = (r1 is r2, r3 is not r1, r4 is not r3, caps, out.shape[0], float(out[0, 0]), float(buf[5, 0]))
    finally:
        dist.destroy_process_group()


def test_one_reducer_per_width_serves_every_token_count():
    """TensorParallel.reducer: one reduction object per output width, sized once and reused for every smaller token count
    (creating one is a collective), grown when a larger count arrives; a call reduces its own rows only."""
    import torch.multiprocessing as mp
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_capacity_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for r in (0, 1):
        same, grown, other, caps, rows, val, untouched = ret[r]
        assert same and grown and other and caps == (2048, 4096, 2048)
        assert rows == 4 and val == 3.0 and untouched == 0.0
