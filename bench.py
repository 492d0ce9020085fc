#!/usr/bin/env python
"""Benchmark of the FlatQuant W4A4 FlatLinear hot path on B200.

Workload (BASELINE.json configs[1], "qlinear_benchmark.py: all Llama-3-8B linear shapes, W4A4, seq=2048"):
one STEP = one prefill pass of a 2048-token batch through all 224 FlatLinears of Llama-3-8B
(32 layers x {q,k,v,o,gate,up,down}); per layer that is 4 transform+quantise launches (q/k/v and
gate/up share one quantised activation, as the fused design intends) and 4 tcgen05 GEMM launches
(qkv N=6144, o N=4096, gate_up N=28672, down K=14336->4096).  Every layer has its own random packed
weights (3.5 GB in total, so consecutive GEMMs never find their weights in the 126 MB L2).  Synthetic
activations; attention / SiLU between the linears are outside the hot path and are not run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for the definition of every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN, INTER, KV = 4096, 14336, 1024
N_LAYERS, SEQ = 32, 2048
# (name, K, N, (n1, n2)) -- fused column groups of one decoder layer
LINEARS = [("qkv", HIDDEN, HIDDEN + 2 * KV, (64, 64)), ("o", HIDDEN, HIDDEN, (64, 64)),
           ("gate_up", HIDDEN, 2 * INTER, (64, 64)), ("down", INTER, HIDDEN, (112, 128))]
OPS_PER_TOKEN_LAYER = sum(2 * k * n for _, k, n, _ in LINEARS)  # GEMM ops only
METRIC = "Llama-3-8B W4A4 prefill tokens/sec/GPU (FlatLinear path: 224 linears, seq 2048)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed summary of this round's
    `ncu --set full` capture (profiles/ncu_traffic.json: {kernel: {"bytes_per_launch": ..., "source": ...}}); None when
    the kernel has not been captured."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return d[kernel]["bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout_s=5.0):
        """nvidia-smi takes a while to come up on a fresh box: block until it delivers (warm-up runs meanwhile)."""
        t_end = time.time() + timeout_s
        while self.proc is not None and not self.rows and time.time() < t_end:
            time.sleep(0.01)

    def mark(self):
        return time.time()

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 is None or (t0 <= t <= t1 + 0.03)]
        window = "timed region"
        if not rows:  # region shorter than one sampling period: fall back to everything seen under the same load
            rows, window = [r for (_, r) in self.rows], "warm-up + timed region"
        self.rows = rows
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        # "under load": the upper half of the samples (idle samples at the edges pull the median down)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window,
                "power_w": (statistics.median(pw) if pw else None)}


def make_layers(dev, n_layers, gen):
    """Random packed INT4 weights, per-channel scales and Kronecker factors for every layer."""
    layers = []
    for _ in range(n_layers):
        layer = {}
        for name, k, n, (n1, n2) in LINEARS:
            layer[name] = dict(
                w=torch.randint(0, 256, (n, k // 2), dtype=torch.uint8, device=dev, generator=gen),
                s_w=(torch.rand(n, device=dev, generator=gen) * 0.01 + 0.005),
                left=(torch.randn(n1, n1, device=dev, generator=gen) / n1 ** 0.5).half(),
                right=(torch.randn(n2, n2, device=dev, generator=gen) / n2 ** 0.5).half())
        layers.append(layer)
    return layers


class Runner:
    """Holds device buffers so the timed region allocates nothing."""

    def __init__(self, dev, layers):
        from flatquant_b200 import ops
        self.ops, self.dev, self.layers = ops, dev, layers
        self.cm = float(torch.sigmoid(torch.tensor(4.0)))
        self.x = {k: torch.empty(SEQ, k, dtype=torch.float16, device=dev) for k in (HIDDEN, INTER)}
        self.xq = {k: (torch.empty(SEQ, k, dtype=torch.int8, device=dev), torch.empty(SEQ, dtype=torch.float32, device=dev))
                   for k in (HIDDEN, INTER)}
        self.y = {name: torch.empty(SEQ, n, dtype=torch.float16, device=dev) for name, _, n, _ in LINEARS}
        self.launches = 0

    def set_input(self, x_hidden, x_inter):
        self.x[HIDDEN].copy_(x_hidden, non_blocking=True)
        self.x[INTER].copy_(x_inter, non_blocking=True)

    def step(self, x=None, y_last=None):
        """One pass over all layers.  x: input buffers to read (default: the resident ones); y_last: where the last
        layer's down projection goes (the step's result)."""
        ops, cm = self.ops, self.cm
        x = x or self.x
        last = self.layers[-1]
        for layer in self.layers:
            for name, k, n, (n1, n2) in LINEARS:
                p = layer[name]
                out = y_last if (y_last is not None and layer is last and name == "down") else self.y[name]
                ops.act_quant(x[k], p["left"], p["right"], n1, n2, cm, cm, out=self.xq[k])
                ops.gemm_w4a8(self.xq[k][0], self.xq[k][1], p["w"], p["s_w"], out=out)
                self.launches += 2
        return y_last if y_last is not None else self.y["down"]

    def fused_per_shape(self, reps, barrier, peak_tops):
        """BASELINE's second metric ("FlatLinear INT4 TFLOPS vs B200 peak", qlinear_benchmark.py protocol: one shape at a
        time): GPU time of ONE fused FlatLinear -- transform + quantise + W4A4 GEMM + dequant through ops.linear_fused --
        per Llama-3-8B shape at 2048 tokens, cycling over the layers' distinct weights.  The 4096 -> 4096 row is the
        shape the north star quotes its 70 % target on.  Timed as a kernel "timed alone" (the burst peak's definition in
        MEASURED_PEAKS.json): the pass over the 32 layers' weights is replayed from a CUDA graph after a short idle gap,
        so neither the host's per-call cost (ctypes + tensor-map encoding, ~10-20 us) nor the power state left behind
        by the preceding sustained passes is in the number; `us_host_loop` is the same pass issued call by call."""
        out = {}
        for name, k, n, (n1, n2) in LINEARS:
            ps = [layer[name] for layer in self.layers]

            def one_pass():
                for p in ps:
                    self.ops.linear_fused(self.x[k], p["left"], p["right"], n1, n2, self.cm, self.cm, p["w"], p["s_w"])
            one_pass()
            ms_loop = timed(one_pass, reps, barrier)
            row = {"us_host_loop": ms_loop * 1e3 / (reps * len(ps))}
            try:
                side = torch.cuda.Stream(device=self.dev)
                side.wait_stream(torch.cuda.current_stream(self.dev))
                with torch.cuda.stream(side):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        one_pass()
                torch.cuda.current_stream(self.dev).wait_stream(side)
                g.replay()
                torch.cuda.synchronize()
                time.sleep(0.25)
                us = timed(g.replay, reps, barrier) * 1e3 / (reps * len(ps))
                del g
            except Exception as e:  # noqa: BLE001 -- no graph support (the CPU mock run): keep the host-loop figure
                row["graph_error"] = f"{type(e).__name__}: {e}"[:120]
                us = row["us_host_loop"]
            tops = 2.0 * SEQ * k * n / us / 1e6
            row.update({"us": us, "tops": tops, "frac_of_peak": tops / peak_tops})
            out[f"{name} {k}->{n}"] = row
        return out

    def transform_only(self):
        """The second kernel of the step alone: online Kronecker transform + per-token quantise of every linear's input."""
        cm = self.cm
        for layer in self.layers:
            for name, k, n, (n1, n2) in LINEARS:
                p = layer[name]
                self.ops.act_quant(self.x[k], p["left"], p["right"], n1, n2, cm, cm, out=self.xq[k])

    def gemm_only(self):
        for layer in self.layers:
            for name, k, n, _ in LINEARS:
                p = layer[name]
                self.ops.gemm_w4a8(self.xq[k][0], self.xq[k][1], p["w"], p["s_w"], out=self.y[name])


# ---------------------------------------------------------------- configs[3]: Llama-3-70B tensor parallel
H70, I70, KV70, L70 = 8192, 28672, 1024, 80


def measure_tp70(args, rank, world, dev, dist, barrier, steps, n_layers=None):
    """BASELINE configs[3]: Llama-3-70B FlatLinears sharded row/column parallel over `world` GPUs (flatquant_b200/tp.py):
    per layer qkv (column), o (row + all-reduce), gate_up (column), down (row + all-reduce); seq 2048; strong scaling.
    Timed three ways in the same process group: the full step, the same step with the reductions left out (compute
    only), and the reductions alone (bus bandwidth) -> the exposed-communication share of the step."""
    from flatquant_b200 import ops, tp as TP
    gen = torch.Generator(device=dev).manual_seed(77 + rank)
    n_layers = n_layers or L70
    cm = float(torch.sigmoid(torch.tensor(4.0)))
    specs = [("qkv", "col", H70, H70 + 2 * KV70), ("o", "row", H70, H70), ("gate_up", "col", H70, 2 * I70), ("down", "row", I70, H70)]
    layers = []
    for _ in range(n_layers):
        layer = []
        for name, kind, K, N in specs:
            sh = TP.column_shard(K, N, world, rank) if kind == "col" else TP.row_shard(K, N, world, rank)
            kk, nn = sh.k1 - sh.k0, sh.n1 - sh.n0
            n1, n2 = sh.factors
            layer.append(dict(kind=kind, sh=sh, kk=kk, nn=nn, n1=n1, n2=n2,
                              w=torch.randint(0, 256, (nn, kk // 2), dtype=torch.uint8, device=dev, generator=gen),
                              s_w=torch.rand(nn, device=dev, generator=gen) * 0.01 + 0.005,
                              left=(torch.randn(n1, n1, device=dev, generator=gen) / n1 ** 0.5).half(),
                              right=(torch.randn(n2, n2, device=dev, generator=gen) / n2 ** 0.5).half()))
        layers.append(layer)
    xs = {kk: torch.randn(SEQ, kk, device=dev, generator=gen).half() for kk in {l["kk"] for l in layers[0]}}
    xq = {kk: (torch.empty(SEQ, kk, dtype=torch.int8, device=dev), torch.empty(SEQ, dtype=torch.float32, device=dev)) for kk in xs}
    ys = {i: torch.empty(SEQ, l["nn"], dtype=torch.float16, device=dev) for i, l in enumerate(layers[0])}
    launches = [0]
    reducer = TP.make_reducer(world, rank, dev, (SEQ, H70), torch.float16) if world > 1 else None
    how = reducer.name if reducer is not None else "none (tp1)"
    # what the reductions alone cost, for the bus-bandwidth figure: the in-switch all-reduce kernel of the same library
    # when the box has multicast, else NCCL
    comm_reducer = reducer
    if reducer is not None and getattr(reducer, "fused", False) and not getattr(reducer, "par_mc", 0):
        comm_reducer = TP.NcclReducer((SEQ, H70), torch.float16, dev)

    def step(comm=True):
        for layer in layers:
            for i, l in enumerate(layer):
                row = l["kind"] == "row" and world > 1
                ops.act_quant(xs[l["kk"]], l["left"], l["right"], l["n1"], l["n2"], cm, cm, out=xq[l["kk"]])
                if row and comm:   # GEMM + reduction: one call into the reducer (fused epilogue reduce-scatter, or NCCL)
                    reducer.rowparallel_gemm(xq[l["kk"]][0], xq[l["kk"]][1], l["w"], l["s_w"])
                    launches[0] += 3
                else:
                    ops.gemm_w4a8(xq[l["kk"]][0], xq[l["kk"]][1], l["w"], l["s_w"], out=ys[i])
                    launches[0] += 2

    def comm_only():   # the same number of all-reduces of the same [2048, 8192] fp16 payload, nothing else
        for layer in layers:
            for i, l in enumerate(layer):
                if l["kind"] == "row":
                    comm_reducer.all_reduce(comm_reducer.partial_buffer())
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    launches[0] = 0
    ms = timed(step, steps, barrier)
    n_launch = launches[0]
    ms_nocomm = timed(lambda: step(False), steps, barrier) if world > 1 else ms
    ms_comm = timed(comm_only, steps, barrier) if world > 1 else 0.0
    if world > 1:
        t = torch.tensor([ms, ms_nocomm, ms_comm], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_nocomm, ms_comm = t.tolist()
    scale = L70 / n_layers
    ops_layer = sum(2 * K * N for _, _, K, N in specs) * SEQ
    ar_bytes = SEQ * H70 * 2
    n_ar = 2 * n_layers * steps
    res = {"tokens_per_s": SEQ * steps / (ms * 1e-3) / scale, "ms_per_step": ms / steps * scale,
           "ms_per_step_compute_only": ms_nocomm / steps * scale, "exposed_comm_share": max(0.0, 1.0 - ms_nocomm / ms),
           "tops_gemm_per_gpu": ops_layer * n_layers * steps / world / (ms * 1e-3) / 1e12,
           "parallelism": f"tp{world}", "layers_timed": n_layers, "gpu_launches": n_launch, "reduction": how,
           "allreduce_bytes": ar_bytes, "allreduces_per_step": 2 * L70}
    if world > 1:
        us_ar = ms_comm * 1e3 / n_ar
        res.update({"allreduce_us_alone": us_ar, "allreduce_bus_GBps": 2 * (world - 1) / world * ar_bytes / us_ar / 1e3,
                    "limiter": "all-reduce" if ms - ms_nocomm > 0.5 * ms_nocomm else "FlatLinear GEMMs (per-launch fixed cost at the sharded shapes)"})
    if n_layers != L70:
        res["note"] = f"{n_layers} of 80 identical layers timed, scaled"
    del layers
    torch.cuda.empty_cache()
    return res


def run_tp70(args, rank, world, local):
    """`--workload llama70b-tp`: configs[3] as its own bench line."""
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist, barrier = _init_dist(world, dev)
    r = measure_tp70(args, rank, world, dev, dist, barrier, args.steps, args.layers if args.layers != N_LAYERS else L70)
    if rank == 0:
        line = {"metric": "Llama-3-70B W4A4 tensor-parallel prefill tokens/sec (FlatLinear path, 80 layers, seq 2048)",
                "value": r["tokens_per_s"], "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
                "warmup": 3, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "s4 x s4 -> s32, f16 in/out", "data": "synthetic",
                "config": {"workload": "llama3-70b FlatLinear prefill, row/column parallel, 2 all-reduces of [2048, 8192] f16 per layer",
                           "parallelism": f"tp{world}", "layers": r["layers_timed"]},
                "gpu_launches": r["gpu_launches"], "tp70b": r}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------- configs[2]: Llama-3-8B end to end (model stack)
def _init_dist(world, dev):
    if world <= 1:
        return None, (lambda: None)
    import torch.distributed as dist
    if not dist.is_initialized():
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL's version banner goes to stdout, where only the JSON line belongs
        dist.init_process_group("nccl", device_id=dev)
    return dist, (lambda: dist.barrier())


def _count_ops(ops, names, counted):
    """Wrap the named flatquant_b200.ops entry points with a launch counter; returns the undo function.
    linear_fused / linear_fused_silu are two launches (transform + GEMM) except for the one or two tokens of a decode
    step with a two-sided or one-sided Kronecker pair, where they are one (csrc/flatlinear_decode.cuh)."""
    saved = {n: getattr(ops, n) for n in names}
    for name, fn in saved.items():
        def wrap(f, two=name.startswith("linear_fused")):
            def inner(*a, **k):
                counted["n"] += 2 if two and a[0].numel() // a[0].shape[-1] > 2 else 1
                return f(*a, **k)
            return inner
        setattr(ops, name, wrap(fn))

    def undo():
        for n, f in saved.items():
            setattr(ops, n, f)
    return undo


def measure_model(args, rank, world, dev, dist, barrier, steps, sample_clocks=True):
    """BASELINE configs[2]: the whole random-init Llama-3-8B W4A4 model (flatquant_b200.deploy.transformers) -- embedding,
    32 decoder layers (FlatLinears + RoPE + KV transform/quantise + attention), final norm, fp16 lm_head -- prefill at
    seq 2048 and CUDA-event-timed decode steps at bsz 1 and 16 over the paged INT4 cache, eager and replayed from one
    CUDA graph (benchmarks/benchmark_model.py:264-305 protocol).  Returns the measurements as a dict."""
    from flatquant_b200 import ops
    from flatquant_b200.deploy.transformers import FlatQuantLlamaConfig, FlatQuantLlamaForCausalLM
    cfg = FlatQuantLlamaConfig.llama3_8b(num_hidden_layers=args.layers)
    model = FlatQuantLlamaForCausalLM.random_init(cfg, dev, seed=1234 + rank)
    counted = {"n": 0}
    undo = _count_ops(ops, ("act_quant", "gemm_w4a8", "gemm_w4a8_silu", "linear_fused", "linear_fused_silu", "kron_transform", "kv_quant",
                            "kv_fused", "kv_decode_i4",
                            "hadamard", "rmsnorm", "add_rmsnorm", "rope_", "silu_mul"), counted)
    try:
        gen = torch.Generator().manual_seed(rank)
        ids_host = torch.randint(100, 200, (1, SEQ), generator=gen).pin_memory()  # benchmark_model.py:281
        ids = ids_host.to(dev)
        decode_len = 64
        cache = model.build_cache(1, page_size=2048, max_length=SEQ + decode_len)
        out_host = torch.empty(1, cfg.vocab_size, dtype=torch.float16).pin_memory()

        def prefill():
            cache.reset()
            with torch.no_grad():
                return model(ids, past_key_values=cache)[0]

        def prefill_e2e():
            cache.reset()
            with torch.no_grad():
                logits = model(ids_host.to(dev, non_blocking=True), past_key_values=cache)[0]
            out_host.copy_(logits[:, -1], non_blocking=True)

        sampler = ClockSampler(dev.index or 0)
        if rank == 0 and sample_clocks:
            sampler.start()
        for _ in range(3):
            prefill()
        torch.cuda.synchronize()
        if rank == 0 and sample_clocks:
            sampler.wait_first()
        counted["n"] = 0
        t0 = sampler.mark()
        ms = timed(prefill, steps, barrier)
        t1 = sampler.mark()
        launches = counted["n"]
        clocks = sampler.stop(t0, t1) if (rank == 0 and sample_clocks) else None
        ms_e2e = timed(prefill_e2e, steps, barrier)

        def decode_ms(bsz):
            c = model.build_cache(bsz, page_size=2048, max_length=SEQ + decode_len)
            tok = torch.randint(100, 200, (bsz, 1), device=dev)
            with torch.no_grad():
                model(ids.expand(bsz, SEQ).contiguous(), past_key_values=c, logits_to_keep=1)
                for _ in range(4):
                    model(tok, past_key_values=c)

                def run_steps():
                    for _ in range(decode_len - 8):
                        model(tok, past_key_values=c)
                return timed(run_steps, 1, barrier) / (decode_len - 8)
        dec = {}
        for b in (1, 16):
            try:
                dec[f"bsz{b}_ms_per_token"] = decode_ms(b)
            except Exception as e:  # noqa: BLE001
                dec[f"bsz{b}_error"] = f"{type(e).__name__}: {e}"[:200]

        def decode_graph_ms(bsz):
            """The same steps replayed from one CUDA graph (position on the device, graphs.GraphedDecoder)."""
            from flatquant_b200.deploy.transformers.graphs import GraphedDecoder
            c = model.build_cache(bsz, page_size=SEQ + decode_len, max_length=SEQ + decode_len)
            tok = torch.randint(100, 200, (bsz, 1), device=dev)
            with torch.no_grad():
                model(ids.expand(bsz, SEQ).contiguous(), past_key_values=c, logits_to_keep=1)
            g = GraphedDecoder(model, c)
            for _ in range(4):
                g.step(tok)

            def run_steps():
                for _ in range(decode_len - 8):
                    g.step(tok)
            return timed(run_steps, 1, barrier) / (decode_len - 8)
        for b in (1, 16):
            try:
                dec[f"bsz{b}_graph_ms_per_token"] = decode_graph_ms(b)
            except Exception as e:  # a capture problem must not lose the eager numbers of the same run
                dec[f"bsz{b}_graph_error"] = f"{type(e).__name__}: {e}"[:200]
        # HBM floor of one decode step: every packed weight + the fp16 lm_head streamed once (bsz-independent), plus the
        # INT4 KV cache of the attended 2048 tokens
        w_bytes = args.layers * sum(k * n // 2 for _, k, n, _ in LINEARS) + cfg.vocab_size * HIDDEN * 2
        hbm = peaks()[0]
        for b in (1, 16):
            kv_bytes = args.layers * b * cfg.num_key_value_heads * SEQ * 2 * (64 + 4)
            floor_ms = (w_bytes + kv_bytes) / (hbm * 1e9) * 1e3
            dec[f"bsz{b}_hbm_floor_ms"] = floor_ms
            best = min([v for k_, v in dec.items() if k_.startswith(f"bsz{b}_") and k_.endswith("ms_per_token")], default=None)
            if best:
                dec[f"bsz{b}_frac_of_hbm_floor"] = floor_ms / best
        if world > 1:
            t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, ms_e2e = t.tolist()
        return {"prefill_tokens_per_s": world * SEQ * steps / (ms * 1e-3), "prefill_ms": ms / steps,
                "e2e_prefill_tokens_per_s": world * SEQ * steps / (ms_e2e * 1e-3),
                "h2d_bytes_per_step": ids_host.numel() * 8, "d2h_bytes_per_step": out_host.numel() * 2,
                "gpu_launches": launches, "clocks": clocks, "decode": dec, "steps": steps, "layers": args.layers,
                "decode_protocol": "prefill 2048 then 56 timed single-token steps (benchmarks/benchmark_model.py:285-297); "
                                   "graph = the whole step replayed from one CUDA graph"}
    finally:
        undo()
        del model
        torch.cuda.empty_cache()


def run_model(args, rank, world, local):
    """`--workload llama8b-model`: configs[2] as its own bench line.  Data-parallel replicas for N > 1."""
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist, barrier = _init_dist(world, dev)
    r = measure_model(args, rank, world, dev, dist, barrier, args.steps)
    if rank == 0:
        line = {"metric": "Llama-3-8B W4A4 end-to-end prefill tokens/sec/GPU (whole model, seq 2048, bsz 1)",
                "value": r["prefill_tokens_per_s"], "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
                "warmup": 3, "ms_per_step": r["prefill_ms"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "s4 x s4 -> s32 linears, f16 attention / lm_head", "data": "synthetic",
                "config": {"workload": "llama3-8b whole-model prefill seq 2048 + decode bsz 1/16, random-init weights",
                           "l2": "inputs larger than L2: 3.5 GB of packed weights + 1 GB lm_head per step",
                           "parallelism": f"dp{world} (independent replicas)", "layers": args.layers},
                "clocks": r["clocks"], "gpu_launches": r["gpu_launches"],
                "e2e": {"value": r["e2e_prefill_tokens_per_s"], "unit": "tokens/s",
                        "h2d_bytes_per_step": r["h2d_bytes_per_step"], "d2h_bytes_per_step": r["d2h_bytes_per_step"]},
                "decode": r["decode"]}
        if args.layers != N_LAYERS:
            line["INVALID"] = "debug run with fewer layers"
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------- configs[4]: DeepSeek-V3 single block
def measure_deepseek(args, rank, world, dev, dist, barrier, steps, sample_clocks=True):
    """BASELINE configs[4]: one random-init DeepSeek-V3 MoE block at the public V3 shape (dim 7168, 128 heads, MLA
    ranks 1536 / 512, 256 routed experts top-8 + 1 shared, moe_inter 2048; 5.8 GB of packed expert weights), prefill of
    2048 tokens.  Returns the measurements as a dict."""
    from types import SimpleNamespace

    from flatquant_b200 import ops
    from flatquant_b200.deploy.transformers import modeling_deepseek as MD
    a = SimpleNamespace(**MD.V3_ARGS)
    block = MD.FlatQuantDeepSeekBlock(a, MD.random_block_params(a, dev, seed=99 + rank), dev)
    counted = {"n": 0}
    undo = _count_ops(ops, ("act_quant", "gemm_w4a8", "gemm_w4a8_grouped"), counted)
    try:
        x_host = torch.randn(1, SEQ, a.dim, generator=torch.Generator().manual_seed(rank)).half().pin_memory()
        x = x_host.to(dev)
        y_host = torch.empty(1, SEQ, a.dim, dtype=torch.float16).pin_memory()
        cos, sin = MD.precompute_freqs(a, SEQ, dev)

        def step():
            with torch.no_grad():
                return block(x, 0, cos, sin)

        def step_e2e():
            with torch.no_grad():
                y_host.copy_(block(x_host.to(dev, non_blocking=True), 0, cos, sin), non_blocking=True)

        sampler = ClockSampler(dev.index or 0)
        if rank == 0 and sample_clocks:
            sampler.start()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if rank == 0 and sample_clocks:
            sampler.wait_first()
        counted["n"] = 0
        t0 = sampler.mark()
        ms = timed(step, steps, barrier)
        t1 = sampler.mark()
        launches = counted["n"]
        clocks = sampler.stop(t0, t1) if (rank == 0 and sample_clocks) else None
        ms_e2e = timed(step_e2e, steps, barrier)
        if world > 1:
            t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, ms_e2e = t.tolist()
        qk = a.qk_nope_head_dim + a.qk_rope_head_dim
        ops_tok = 2 * (a.dim * (a.q_lora_rank + a.kv_lora_rank + a.qk_rope_head_dim) + a.q_lora_rank * a.n_heads * qk +
                       a.n_heads * a.v_head_dim * a.dim + 3 * a.dim * a.moe_inter_dim * (a.n_activated_experts + a.n_shared_experts))
        return {"tokens_per_s": world * SEQ * steps / (ms * 1e-3), "ms_per_step": ms / steps,
                "e2e_tokens_per_s": world * SEQ * steps / (ms_e2e * 1e-3),
                "h2d_bytes_per_step": x_host.numel() * 2, "d2h_bytes_per_step": y_host.numel() * 2,
                "tops_flatlinear": ops_tok * SEQ * steps / (ms * 1e-3) / 1e12, "gpu_launches": launches, "clocks": clocks,
                "steps": steps,
                "workload": "deepseek-v3 MoE block (public 671B shape), MLA + 256 routed experts top-8 (grouped GEMM) + shared, seq 2048"}
    finally:
        undo()
        del block
        torch.cuda.empty_cache()


def run_deepseek(args, rank, world, local):
    """`--workload deepseek-block`: configs[4] as its own bench line.  Data-parallel replicas for N > 1."""
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist, barrier = _init_dist(world, dev)
    r = measure_deepseek(args, rank, world, dev, dist, barrier, args.steps)
    if rank == 0:
        line = {"metric": "DeepSeek-V3 single MoE block W4A4 prefill tokens/sec/GPU (seq 2048)",
                "value": r["tokens_per_s"], "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
                "warmup": 3, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "s4 x s4 -> s32 FlatLinears, f16 attention / gate", "data": "synthetic",
                "config": {"workload": r["workload"], "l2": "inputs larger than L2: 5.8 GB of packed expert weights",
                           "parallelism": f"dp{world} (independent replicas)"},
                "tops_flatlinear": r["tops_flatlinear"], "clocks": r["clocks"], "gpu_launches": r["gpu_launches"],
                "e2e": {"value": r["e2e_tokens_per_s"], "unit": "tokens/s",
                        "h2d_bytes_per_step": r["h2d_bytes_per_step"], "d2h_bytes_per_step": r["d2h_bytes_per_step"]}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------- decode and qattention rooflines (HBM-bound side kernels)
def measure_decode_linears(dev, barrier):
    """Decode-shaped FlatLinears (1 and 16 tokens): GPU time of one fused FlatLinear (transform + quantise + weight-stream
    GEMM + dequant through ops.linear_fused) per Llama-3-8B shape, replayed from a CUDA graph over 16 distinct weights
    (every launch streams its packed weight from HBM).  Bound: HBM; algorithmic bytes per launch = K*N/2 (the 4-bit
    weight stream) -- activations, scales and outputs are < 1 % at these shapes."""
    from flatquant_b200 import ops
    hbm = peaks()[0]
    cm = float(torch.sigmoid(torch.tensor(4.0)))
    out = {}
    for m in (1, 16):
        for name, k, n, (n1, n2) in LINEARS:
            x = torch.randn(m, k, device=dev).half()
            left = (torch.randn(n1, n1, device=dev) / n1 ** 0.5).half()
            right = (torch.randn(n2, n2, device=dev) / n2 ** 0.5).half()
            ws = [torch.randint(0, 256, (n, k // 2), dtype=torch.uint8, device=dev) for _ in range(16)]
            sw = torch.rand(n, device=dev) * 0.01 + 0.005

            def run(i):
                return ops.linear_fused(x, left, right, n1, n2, cm, cm, ws[i % 16], sw)
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                run(0)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for i in range(16):
                        run(i)
            torch.cuda.current_stream(dev).wait_stream(side)
            g.replay()
            us = timed(g.replay, 5, barrier) * 1e3 / (5 * 16)
            gbs = (k * n / 2) / us / 1e3
            out[f"m{m} {name} {k}->{n}"] = {"us": us, "weight_stream_GBps": gbs, "frac_of_hbm": gbs / hbm,
                                            "algorithmic_bytes": k * n // 2, "traffic": ncu_traffic(f"decode m{m} {name}")}
            del ws, g
    return out


def measure_qattention(dev, barrier):
    """The qattention side path (shapes of benchmarks/qattention_benchmark.py:10-14: heads in
    {32, 40, 64} x 128, seq 2048): online Hadamard (bytes = 4 * rows * n), KV quantise + pack + page scatter (256 B in,
    64 + 4 B out per (token, head) row) and single-query attention over the packed INT4 cache (2 * bsz * kv_heads * T *
    (64 + 4) bytes).  CUDA events over 20 launches each; bound: HBM."""
    from flatquant_b200 import ops
    hbm = peaks()[0]
    D = 128
    out = {}

    def t_us(fn, reps=20):
        # best of three timed regions: the first region after an allocation / an idle gap has come out 10-20x slow on some
        # boxes (clock ramp after idling; seen on whichever sub-measurement ran first), a property of the box, not the kernel
        for _ in range(5):
            fn()
        return min(timed(lambda: [fn() for _ in range(reps)], 1, barrier) for _ in range(3)) * 1e3 / reps
    spin = torch.randn(4096, 4096, device=dev, dtype=torch.float16)   # the sub-benchmarks before this one end with frees and
    for _ in range(50):                                                # idle gaps: bring the clocks back up before timing
        spin = (spin @ spin).clamp_(-1, 1)                             # microsecond-scale kernels
    torch.cuda.synchronize()
    del spin
    for n in (128, 4096, 14336):   # K/Q head transform; hidden; Llama-3-8B down_proj input (28 x 512, fused K x 2^m kernel)
        rows = 2048 * 32 if n == 128 else 2048
        x = torch.randn(rows, n, device=dev).half()
        if n & (n - 1) == 0:
            us = t_us(lambda: ops.hadamard(x, 1.0 / n ** 0.5))
        else:
            from flatquant_b200.deploy.functional.hadamard import get_hadK
            hk, k = get_hadK(n)
            hk = hk.to(dev).half()
            us = t_us(lambda: ops.hadamard_kn(x, hk, k, 1.0 / n ** 0.5))
        gbs = 4.0 * rows * n / us / 1e3
        out[f"hadamard rows={rows} n={n}"] = {"us": us, "GBps": gbs, "frac_of_hbm": gbs / hbm}
    for heads in (32, 40, 64):
        kvh = 8 if heads in (32, 64) else heads   # Llama-3 GQA (8 kv heads); llama-13b is MHA
        for bsz in (1, 16):
            T = SEQ
            pages = torch.zeros(bsz, 1, 2, kvh, T, D // 2, dtype=torch.uint8, device=dev)
            params = torch.zeros(bsz, 1, 2, kvh, T, 2, dtype=torch.float16, device=dev)
            slot = (torch.arange(bsz, device=dev)[:, None] * T + torch.arange(T, device=dev)[None]).reshape(-1).int()
            kx = torch.randn(bsz * T, kvh, D, device=dev).half()
            us_q = t_us(lambda: ops.kv_quant(kx, pages, params, slot, 0, 0))
            ops.kv_quant(kx, pages, params, slot, 0, 1)
            gbs_q = bsz * T * kvh * (256 + 68) / us_q / 1e3
            ip = torch.arange(bsz + 1, dtype=torch.int32, device=dev)
            ix = torch.arange(bsz, dtype=torch.int32, device=dev)
            lp = torch.full((bsz,), T, dtype=torch.int32, device=dev)
            q = torch.randn(bsz, heads, D, device=dev).half()
            us_d = t_us(lambda: ops.kv_decode_i4(q, pages, params, ip, ix, lp, layer=0, max_tokens_hint=T))
            gbs_d = 2.0 * bsz * kvh * T * 68 / us_d / 1e3
            out[f"heads={heads} kv_heads={kvh} bsz={bsz}"] = {
                "kv_quant_us": us_q, "kv_quant_GBps": gbs_q, "kv_quant_frac_of_hbm": gbs_q / hbm,
                "decode_attn_us": us_d, "decode_attn_GBps": gbs_d, "decode_attn_frac_of_hbm": gbs_d / hbm,
                "decode_attn_traffic": ncu_traffic(f"decode_attn heads={heads} kv_heads={kvh} bsz={bsz}")}
            del pages, params, kx
    return out


def timed(fn, steps, dist_barrier):
    dist_barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    dist_barrier()
    return e0.elapsed_time(e1)  # ms on the launching (current) stream


CPU_SAMPLE_TOKENS = 256


def cpu_reference_pass(threads, reps, warm, tokens=CPU_SAMPLE_TOKENS, n_layers=None):
    """The oracle's fake-quant path (the reference's own CPU arithmetic: trans -> fake-quant -> F.linear), fp32, on the
    host cores.  One step = `tokens` tokens (a bounded sample of the 2048-token batch; tokens are independent on this
    path) through ALL 224 linears -- 32 passes over one layer's weights (same shapes; 872 MB of fp32 weights per layer is
    far beyond any last-level cache, so every pass streams them from DRAM).  Nothing is extrapolated: tokens/s =
    tokens / measured step time.  Returns (tokens/s, seconds per step)."""
    from oracle import flatq_oracle as O
    torch.set_num_threads(threads)
    n_layers = n_layers or N_LAYERS
    g = torch.Generator().manual_seed(0)
    lin = []
    for name, k, n, (n1, n2) in LINEARS:
        codes = torch.randint(-8, 8, (n, k), generator=g).float()
        lin.append((torch.randn(tokens, k, generator=g), torch.randn(n1, n1, generator=g) / n1 ** 0.5,
                    torch.randn(n2, n2, generator=g) / n2 ** 0.5, codes * 0.01))

    def one_step():
        for _ in range(n_layers):
            for x, L, R, w in lin:
                O.flat_linear_fake(x, L, R, 4.0, 4.0, w)
    for _ in range(warm):
        one_step()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        one_step()
        times.append(time.perf_counter() - t0)
    t_step = statistics.mean(times)
    return tokens / t_step, t_step


def reference_gpu_kernels(device_index):
    """Reference arm only: the reference's own CUTLASS s4 GEMM + dequant kernels (oracle/_ref, built from
    /root/reference for sm_100a) over the work of our arm's GEMM-only roofline pass, in a SUBPROCESS
    (tests/tools/ref_kernels_bench.py) so that a fault inside a reference kernel cannot touch this process."""
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "ref_kernels_bench.py"), "--device",
                              str(device_index)], capture_output=True, text=True, timeout=300)
        last = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        res = json.loads(last[-1]) if last else {"unavailable": (out.stderr or "no output")[-300:]}
    except Exception as e:  # noqa: BLE001
        res = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", type=int, default=N_LAYERS, help="debug only: fewer layers => INVALID as a bench value")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="default arm only: skip the extra sub-measurements (model_e2e, deepseek_block, roofline_decode, "
                         "roofline_qattention at N = 1; tp70b at N > 1)")
    ap.add_argument("--workload", default="llama8b", choices=["llama8b", "llama70b-tp", "llama8b-model", "deepseek-block"],
                    help="llama8b = BASELINE configs[1] (default, data-parallel replicas); llama70b-tp = configs[3]; "
                         "llama8b-model = configs[2] (whole model: prefill + decode); deepseek-block = configs[4]")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    threads = os.cpu_count() or 1

    if args.impl == "reference":
        # Tier rule: the reference arm = the reference's own CPU implementation of the path (oracle port of the
        # fake-quant code; /root/reference is not on the GPU box) on all host threads.  Rank 0 only.
        if rank != 0:
            return
        # every step = CPU_SAMPLE_TOKENS tokens through all 224 linears (nothing extrapolated); the step count is bounded
        # so that the whole arm ends within a few minutes whatever --steps the driver passes
        steps_run = max(1, min(args.steps, 8))
        warm_run = max(1, min(args.warmup, 2))
        tok_s, t_step = cpu_reference_pass(threads, steps_run, warm_run)
        sample = (f"{CPU_SAMPLE_TOKENS} of the 2048 tokens per step through all 32 layers x 7 linears (oracle port of the "
                  f"reference's CPU fake-quant path, fp32, {threads} threads); {warm_run} warm-up + {steps_run} timed steps")
        line = {"impl": "reference", "metric": METRIC, "value": tok_s, "unit": "tokens/s", "n_gpus": args.gpus,
                "steps": steps_run, "warmup": warm_run, "ms_per_step": t_step * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": "llama3-8b FlatLinear prefill: 32 layers x (qkv, o, gate_up, down) W4A4, seq 2048, bsz 1",
                                                "sample": sample,
                                                "n_gpus_note": "one CPU process on rank 0 whatever N: compare at N = 1"},
                "cpu_baseline": {"value": tok_s, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": tok_s, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        # extra, GPU boxes only: the reference's own native GEMM + dequant kernels on this GPU (compare with the other
        # arm's roofline.avg_launch_us x launches / steps, i.e. its GEMM-only pass)
        try:
            if torch.cuda.is_available():
                line["reference_gpu_kernels"] = reference_gpu_kernels(local)
        except Exception as e:  # noqa: BLE001
            line["reference_gpu_kernels"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        print(json.dumps(line), flush=True)
        return

    if args.workload == "llama70b-tp":
        return run_tp70(args, rank, world, local)
    if args.workload == "llama8b-model":
        return run_model(args, rank, world, local)
    if args.workload == "deepseek-block":
        return run_deepseek(args, rank, world, local)

    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL's version banner goes to stdout, where only the JSON line belongs
        dist.init_process_group("nccl", device_id=dev)
        barrier = lambda: dist.barrier()  # noqa: E731
    else:
        dist, barrier = None, (lambda: None)

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    layers = make_layers(dev, args.layers, gen)
    run = Runner(dev, layers)
    hx = torch.randn(SEQ, HIDDEN, generator=torch.Generator().manual_seed(rank)).half().pin_memory()
    hi = torch.randn(SEQ, INTER, generator=torch.Generator().manual_seed(rank + 1)).half().pin_memory()
    hy = torch.empty(SEQ, HIDDEN, dtype=torch.float16).pin_memory()
    run.set_input(hx, hi)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        run.step()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.wait_first()

    # ---- device-resident timing (value)
    run.launches = 0
    t_mark0 = sampler.mark()
    ms = timed(run.step, args.steps, barrier)
    t_mark1 = sampler.mark()
    launches = run.launches
    clocks = sampler.stop(t_mark0, t_mark1) if rank == 0 else None

    # ---- end-to-end through the public API with host buffers (e2e): every step copies ITS inputs from pinned host
    # memory and its result back, inside the timed region.  Serving-style double buffering: while step i computes, the
    # copy stream brings in step i+1's inputs and takes step i-1's result out; the first copy-in and the last copy-out
    # have nothing to hide behind and are exposed.
    class E2E:
        def __init__(self, steps):
            self.steps, self.i = steps, 0
            self.copy = torch.cuda.Stream(device=dev)
            self.xb = [run.x, {k: torch.empty_like(v) for k, v in run.x.items()}]
            self.yb = [torch.empty(SEQ, HIDDEN, dtype=torch.float16, device=dev) for _ in range(2)]
            self.in_ready = [torch.cuda.Event() for _ in range(2)]
            self.done = [torch.cuda.Event() for _ in range(2)]      # compute of the step that used buffer pair b
            self.out_done = [torch.cuda.Event() for _ in range(2)]  # D2H of that step's result

        def h2d(self, j):
            b = j % 2
            with torch.cuda.stream(self.copy):
                if j >= 2:
                    self.copy.wait_event(self.done[b])   # step j-2 no longer reads these buffers
                self.xb[b][HIDDEN].copy_(hx, non_blocking=True)
                self.xb[b][INTER].copy_(hi, non_blocking=True)
                self.in_ready[b].record(self.copy)

        def __call__(self):
            i, b = self.i, self.i % 2
            main = torch.cuda.current_stream()
            if i == 0:
                self.h2d(0)
            main.wait_event(self.in_ready[b])
            if i + 1 < self.steps:
                self.h2d(i + 1)
            if i >= 2:
                main.wait_event(self.out_done[b])        # result buffer b has been read back
            y = run.step(self.xb[b], self.yb[b])
            self.done[b].record(main)
            with torch.cuda.stream(self.copy):
                self.copy.wait_event(self.done[b])
                hy.copy_(y, non_blocking=True)
                self.out_done[b].record(self.copy)
            if i + 1 == self.steps:
                main.wait_event(self.out_done[b])        # the timed region ends after the last result is on the host
            self.i += 1

    warm = E2E(2)
    for _ in range(2):
        warm()
    torch.cuda.synchronize()
    ms_e2e = timed(E2E(args.steps), args.steps, barrier)

    # ---- roofline pass: the dominant kernel (tcgen05 GEMM) alone, same weights / shapes
    ms_gemm = timed(run.gemm_only, args.steps, barrier)
    ms_tr = None
    try:
        ms_tr = timed(run.transform_only, args.steps, barrier)
    except Exception:  # noqa: BLE001 -- an extra figure must never cost the headline line
        ms_tr = None
    per_shape = None
    if world == 1:
        try:
            per_shape = run.fused_per_shape(3, barrier, 2 * peaks()[1])
        except Exception as e:  # noqa: BLE001 -- an extra table must never cost the headline line
            per_shape = {"error": f"{type(e).__name__}: {e}"[:200]}

    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_gemm], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_gemm = t.tolist()

    if rank == 0:
        scale = N_LAYERS / args.layers
        tok_s = world * SEQ * args.steps / (ms * 1e-3) / scale
        tok_s_e2e = world * SEQ * args.steps / (ms_e2e * 1e-3) / scale
        hbm, bf16, src = peaks()
        peak_i8 = 2 * bf16  # kind::i8 issues at twice the bf16 rate (same as fp8); no faster exact-integer path exists
        n_gemm = args.steps * args.layers * len(LINEARS)
        ops_total = OPS_PER_TOKEN_LAYER * SEQ * args.layers * args.steps
        achieved = ops_total / (ms_gemm * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": tok_s, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "s4 x s4 -> s32 (int8 tensor-core path), f16 in/out",
            "data": "synthetic",
            "config": {"workload": "llama3-8b FlatLinear prefill: 32 layers x (qkv, o, gate_up, down) W4A4, seq 2048, bsz 1",
                       "l2": "inputs larger than L2: 3.5 GB of distinct per-layer packed weights per step",
                       "parallelism": f"dp{world} (independent replicas, no collective on the path)",
                       "layers": args.layers},
            "tops_gemm": OPS_PER_TOKEN_LAYER * SEQ * args.layers * args.steps / (ms * 1e-3) / 1e12,
            "clocks": clocks,
            "e2e": {"value": tok_s_e2e, "unit": "tokens/s",
                    "h2d_bytes_per_step": hx.numel() * 2 + hi.numel() * 2, "d2h_bytes_per_step": hy.numel() * 2,
                    "pipeline": "double-buffered: step i+1's H2D and step i-1's D2H run on a copy stream under step i"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "gemm_w4a8_2sm_kernel (tcgen05.mma.cta_group::2 kind::i8)",
                         "achieved": achieved, "peak": peak_i8, "unit": "TOP/s", "frac": achieved / peak_i8,
                         "peak_source": f"2 x bf16_tflops of MEASURED_PEAKS.json ({src})",
                         "avg_launch_us": ms_gemm * 1e3 / n_gemm, "launches": n_gemm,
                         # dram__bytes_read+write per GEMM launch, mean over the 4 shapes of a layer, from this round's
                         # ncu --set full capture (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the
                         # .ncu-rep; algorithmic bytes 85.0 MB per launch -- the difference is output still dirty in L2)
                         "traffic": ncu_traffic("gemm_w4a8_2sm_kernel"), "traffic_unit": "bytes/launch",
                         "algorithmic_bytes_per_launch": 85.0e6,
                         "issue_peak_note": "tcgen05 kind::i8 issue peak measured by tools/mma_rate2.cu: 8189 MAC/cycle/SM"},
        }
        if ms_tr:
            # HBM-bound companion kernel: algorithmic bytes = fp16 activations in (2 B) + int8 codes out (1 B) + fp32 scale
            tr_bytes = sum(SEQ * k * 3 + SEQ * 4 for _, k, _, _ in LINEARS) * args.layers * args.steps
            gbs = tr_bytes / (ms_tr * 1e-3) / 1e9
            line["roofline_transform"] = {"bound": "hbm", "kernel": "kron_quant_ts_kernel (tcgen05 kind::f16 transform + quantise, intermediate in TMEM)",
                                          "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                          "avg_launch_us": ms_tr * 1e3 / n_gemm, "launches": n_gemm,
                                          "note": "inputs are re-read from L2 when they fit (25-88 MB per launch vs 126 MB L2): "
                                                  "an L2-resident pass can exceed the HBM figure"}
        if per_shape is not None:
            line["fused_linear_per_shape"] = per_shape
        # sustained figure of the same file: what a kernel timed inside a long step can be held against
        try:
            sus = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained")
        except (OSError, ValueError):
            sus = None
        if sus:
            line["roofline"]["peak_sustained"] = 2 * sus
            line["roofline"]["frac_of_sustained"] = achieved / (2 * sus)
        try:   # third denominator: the tensor pipe's own issue rate (tools/mma_rate2.cu, profiles/r02_mma_rate2.log) at the clock of this run
            mhz = float((clocks or {}).get("sm_mhz") or 0)
            if mhz > 0:
                sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
                pk = 2 * 8189 * sms * mhz * 1e6 / 1e12
                line["roofline"]["peak_mma_issue"] = pk
                line["roofline"]["frac_of_mma_issue"] = achieved / pk
                line["roofline"]["peak_mma_issue_source"] = "measured kind::i8 issue rate 8189 MAC/cycle/SM x SMs x median SM clock of the timed region"
        except Exception:  # noqa: BLE001 -- optional evidence only
            pass
        if not args.no_cpu_baseline and world == 1:
            v, t_step = cpu_reference_pass(threads, 3, 1)
            line["cpu_baseline"] = {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port",
                                    "sample": f"oracle fake-quant path, {CPU_SAMPLE_TOKENS} of the 2048 tokens through all 32 layers x 7 "
                                              f"linears per step ({t_step:.2f} s/step), 1 warm-up + 3 timed steps, nothing extrapolated"}
        if args.layers != N_LAYERS:
            line["INVALID"] = "debug run with fewer layers"
    else:
        line = None

    # ---- extra sub-measurements in the same line: the other BASELINE configs and the HBM-bound side
    # kernels.  Each is optional evidence: a failure is recorded under its key and never costs the headline.
    if not args.no_extras:
        del run, layers
        torch.cuda.empty_cache()
        extras = {}

        def attempt(key, fn):
            try:
                extras[key] = fn()
            except Exception as e:  # noqa: BLE001
                extras[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()
        sub_steps = max(2, min(args.steps, 5))
        if world == 1:
            attempt("model_e2e", lambda: measure_model(args, rank, world, dev, dist, barrier, sub_steps, sample_clocks=False))
            attempt("deepseek_block", lambda: measure_deepseek(args, rank, world, dev, dist, barrier, sub_steps, sample_clocks=False))
            attempt("roofline_decode", lambda: measure_decode_linears(dev, barrier))
            attempt("roofline_qattention", lambda: measure_qattention(dev, barrier))
        else:
            attempt("tp70b", lambda: measure_tp70(args, rank, world, dev, dist, barrier, sub_steps))
        if line is not None:
            line.update(extras)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
