"""TEST INFRASTRUCTURE: run bench.py's default main() end to end on the CPU -- flatquant_b200.ops replaced by the
oracle-backed stand-ins (tests/cpu_ops.py), the CUDA surface bench.py touches (events, streams, pinned memory, device
selection) replaced by host stubs, shapes shrunk.  Times mean nothing here; the point is that every line of the
benchmark's control flow and JSON assembly executes (tests/test_bench_logic_host.py asserts the keys)."""
import sys, time, types, json
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, cpu_ops
from flatquant_b200 import ops
for n in cpu_ops.FUNCTIONS: setattr(ops, n, getattr(cpu_ops, n))
# ---- fake CUDA surface
class Ev:
    def __init__(self, enable_timing=False): self.t=None
    def record(self, stream=None): self.t=time.perf_counter()
    def elapsed_time(self, other): return (other.t-self.t)*1e3
class St:
    def __init__(self, device=None): pass
    def wait_event(self, e): pass
    def wait_stream(self, s): pass
    cuda_stream = 0
class Ctx:
    def __init__(self, s): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False
torch.cuda.Event = Ev; torch.cuda.Stream = St; torch.cuda.stream = Ctx
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.set_device = lambda d: None
torch.cuda.current_stream = lambda *a, **k: St()
torch.Tensor.pin_memory = lambda self: self
_real_device = torch.device
import bench
_orig_gen = torch.Generator
bench.torch.device = lambda *a, **k: _real_device("cpu")
bench.torch.Generator = lambda device=None: _orig_gen()
bench.SEQ = 8
bench.LINEARS = [("qkv", 256, 384, (16, 16)), ("o", 256, 256, (16, 16)), ("gate_up", 256, 1024, (16, 16)), ("down", 512, 256, (16, 32))]
bench.HIDDEN, bench.INTER, bench.N_LAYERS = 256, 512, 2
bench.OPS_PER_TOKEN_LAYER = sum(2*k*n for _,k,n,_ in bench.LINEARS)
bench.cpu_reference_pass = lambda *a, **k: (1.0, 1.0)
sys.argv = ["bench.py", "--steps", "2", "--warmup", "1", "--layers", "2", "--no-extras"]
bench.main()
