"""One graphed decode step of the Llama-3-8B-shaped stack (few layers, real widths) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --csv --log-file out.csv python tools/decode_step_probe.py
Prints the measured time per step (CUDA events, graph replay) when run without a profiler."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flatquant_b200.deploy.transformers import FlatQuantLlamaConfig, FlatQuantLlamaForCausalLM
from flatquant_b200.deploy.transformers.graphs import GraphedDecoder

layers = int(os.environ.get("PROBE_LAYERS", "4"))
bsz = int(os.environ.get("PROBE_BSZ", "1"))
dev = torch.device("cuda:0")
cfg = FlatQuantLlamaConfig.llama3_8b(num_hidden_layers=layers)
model = FlatQuantLlamaForCausalLM.random_init(cfg, dev, seed=1)
ids = torch.randint(100, 200, (bsz, 2048), device=dev)
cache = model.build_cache(bsz, page_size=2048 + 64, max_length=2048 + 64)
with torch.no_grad():
    model(ids, past_key_values=cache, logits_to_keep=1)
g = GraphedDecoder(model, cache)
tok = torch.randint(100, 200, (bsz, 1), device=dev)
for _ in range(3):
    g.step(tok)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.step(tok)
e1.record()
torch.cuda.synchronize()
print(f"layers={layers} bsz={bsz}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per graphed decode step")
