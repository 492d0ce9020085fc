"""One prefill pass (2048 tokens) of the Llama-3-8B-shaped stack with a few layers, bracketed by cudaProfilerStart/Stop:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/prefill_probe.py
Prints the measured time per pass (CUDA events) when run without a profiler."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flatquant_b200.deploy.transformers import FlatQuantLlamaConfig, FlatQuantLlamaForCausalLM

layers = int(os.environ.get("PROBE_LAYERS", "2"))
dev = torch.device("cuda:0")
cfg = FlatQuantLlamaConfig.llama3_8b(num_hidden_layers=layers)
model = FlatQuantLlamaForCausalLM.random_init(cfg, dev, seed=1)
ids = torch.randint(100, 200, (1, 2048), device=dev)
cache = model.build_cache(1, page_size=2048, max_length=2048 + 64)


def prefill():
    cache.reset()
    with torch.no_grad():
        return model(ids, past_key_values=cache, logits_to_keep=1)[0]


for _ in range(3):
    prefill()
torch.cuda.synchronize()
torch.cuda.profiler.start()
prefill()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    prefill()
e1.record()
torch.cuda.synchronize()
print(f"layers={layers}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us per prefill pass")
