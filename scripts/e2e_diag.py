"""Where does the end-to-end (host buffers in, host buffers out) step time go?"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from quickvc_official_b200 import SynthesizerTrn
from quickvc_official_b200.pipeline import PipelinedConverter

cfg = bench.model_cfg(); sd = bench.random_init_state_dict(cfg)
dev = torch.device("cuda:0")
net = SynthesizerTrn(641, 32, **cfg, precision=os.environ.get("PREC", "tf32")).eval(); net.load_state_dict(sd); net = net.to(dev)
B, T = 64, 500
g = torch.Generator().manual_seed(1)
unit_h = torch.randn(B, 256, T, generator=g).pin_memory(); mel_h = (torch.randn(1, 80, T, generator=g) * 2 - 5).pin_memory()
unit, mel = unit_h.to(dev), mel_h.to(dev)
noise = torch.randn(B, 192, T, device=dev)

def ev_time(fn, n):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); s.record(); fn(n); e.record(); host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n, host * 1e3 / n

for _ in range(3): net.infer(unit, mel, noise=noise)
print("infer (noise given)      ms/step dev, host:", ev_time(lambda n: [net.infer(unit, mel, noise=noise) for _ in range(n)], 10))
print("infer (noise drawn)      ms/step dev, host:", ev_time(lambda n: [net.infer(unit, mel) for _ in range(n)], 10))
w = net.infer(unit, mel, noise=noise); wave_h = torch.empty(w.shape).pin_memory()
print("H2D unit+mel             ms:", ev_time(lambda n: [(unit.copy_(unit_h, non_blocking=True), mel.copy_(mel_h, non_blocking=True)) for _ in range(n)], 10))
print("D2H wave                 ms:", ev_time(lambda n: [wave_h.copy_(w, non_blocking=True) for _ in range(n)], 10))
conv = PipelinedConverter(net, B, T, T, device=dev)
def run(n):
    for _ in range(n): conv.submit(unit_h, mel_h)
    conv.drain()
run(3)
for n in (5, 10, 20):
    print(f"pipelined e2e, {n} steps  ms/step dev, host:", ev_time(run, n))
def run_noise(n):
    for _ in range(n): conv.submit(unit_h, mel_h, noise=noise)
    conv.drain()
print("pipelined e2e, noise given, 10 steps:", ev_time(run_noise, 10))
