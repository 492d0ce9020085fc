"""Multi-GPU check of the peer-memory reduction path (run under torchrun on >= 2 B200s; not a pytest file because it needs
several processes with one GPU each):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/tp_peer_check.py

Per rank: a row-parallel FlatLinear shard (own K range, own Kronecker pair) at (M, N) = (2048, 8192) and at a decode size;
the PeerReducer result (GEMM epilogue peer stores + owner sum + broadcast; in-switch all-reduce for the decode size) must
equal the NCCL all-reduce of the same partial products up to the fp16 rounding of the partial sums, on every rank, and be
identical on all ranks.  Prints one JSON line per check on rank 0 and timings of both forms."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from flatquant_b200 import ops, tp as TP  # noqa: E402
from flatquant_b200.peer_reduce import PeerReducer  # noqa: E402


def timed(fn, reps):
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for (M, N, K) in ((2048, 8192, 28672), (16, 8192, 8192), (2048, 8192, 8192)):
        sh = TP.row_shard(K, N, world, rank)
        kk = sh.k1 - sh.k0
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        xq = torch.randint(-8, 8, (M, kk), dtype=torch.int8, device=dev, generator=g)
        sx = torch.rand(M, device=dev, generator=g) * 0.1 + 0.05
        w = torch.randint(0, 256, (N, kk // 2), dtype=torch.uint8, device=dev, generator=g)
        sw = torch.rand(N, device=dev, generator=g) * 0.01 + 0.005
        gb = torch.Generator(device=dev).manual_seed(7)
        bias = torch.randn(N, device=dev, generator=gb).half()
        red = PeerReducer(world, rank, dev, (M, N))
        nccl = TP.NcclReducer((M, N), torch.float16, dev)
        y_peer = red.rowparallel_gemm(xq, sx, w, sw, bias).clone()
        y_nccl = nccl.rowparallel_gemm(xq, sx, w, sw, bias).clone()
        torch.cuda.synchronize()
        # the exact sum of the ranks' fp16 partial products, in fp32
        part = ops.gemm_w4a8(xq, sx, w, sw, None, torch.float16).float()
        dist.all_reduce(part)
        exact = part + bias.float()
        err_peer = float((y_peer.float() - exact).abs().max() / exact.abs().max())
        err_nccl = float((y_nccl.float() - exact).abs().max() / exact.abs().max())
        same = y_peer.clone()
        dist.broadcast(same, 0)
        identical = bool(torch.equal(same, y_peer))
        flags = torch.tensor([int(identical), int(err_peer < 2e-3)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        us_peer = timed(lambda: red.rowparallel_gemm(xq, sx, w, sw), 20)
        us_nccl = timed(lambda: nccl.rowparallel_gemm(xq, sx, w, sw), 20)
        us_gemm = timed(lambda: ops.gemm_w4a8(xq, sx, w, sw, None, torch.float16, out=nccl.buf), 20)
        if rank == 0:
            print(json.dumps({"M": M, "N": N, "K_per_rank": kk, "world": world, "path": red.name, "multicast": bool(red.out_mc),
                              "max_rel_err_vs_fp32_sum": {"peer": err_peer, "nccl": err_nccl},
                              "identical_on_all_ranks": bool(flags[0]), "within_tol_on_all_ranks": bool(flags[1]),
                              "us": {"gemm_alone": us_gemm, "gemm+peer_reduce": us_peer, "gemm+nccl_allreduce": us_nccl}}), flush=True)
        ok &= bool(flags[0]) and bool(flags[1])
        del red
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
