"""Does the clock sampler perturb the end-to-end region?  Times each NVML query and the pipelined e2e with / without it."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from quickvc_official_b200 import SynthesizerTrn
from quickvc_official_b200.pipeline import PipelinedConverter
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
for name, fn in (("clock", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                 ("maxclock", lambda: pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)),
                 ("reasons", lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)),
                 ("power", lambda: pynvml.nvmlDeviceGetPowerUsage(h))):
    fn(); t0 = time.perf_counter()
    for _ in range(20): fn()
    print(f"nvml {name}: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per call")
cfg = bench.model_cfg(); sd = bench.random_init_state_dict(cfg)
dev = torch.device("cuda:0")
net = SynthesizerTrn(641, 32, **cfg).eval(); net.load_state_dict(sd); net = net.to(dev)
B, T = 64, 500
g = torch.Generator().manual_seed(1)
unit_h = torch.randn(B, 256, T, generator=g).pin_memory(); mel_h = (torch.randn(1, 80, T, generator=g) * 2 - 5).pin_memory()
conv = PipelinedConverter(net, B, T, T, device=dev)
def run(n):
    for _ in range(n): conv.submit(unit_h, mel_h)
    conv.drain()
def ev_time(n):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(n); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
run(3)
print("e2e no sampler:", [round(ev_time(10), 2) for _ in range(3)])
for label, patch in (("nvml full", None), ("nvml no power", "nopower")):
    smp = bench.ClockSampler(0)
    if patch == "nopower":
        import types
        smp._nvml = types.SimpleNamespace(**{k: getattr(pynvml, k) for k in dir(pynvml) if k.startswith("nvml") or k.startswith("NVML")})
        smp._nvml.nvmlDeviceGetPowerUsage = lambda h: 0
    smp.start(); time.sleep(0.3)
    print(f"e2e with sampler ({label}):", [round(ev_time(10), 2) for _ in range(3)], smp.stop())
