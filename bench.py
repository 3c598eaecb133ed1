#!/usr/bin/env python
"""Benchmark of the QuickVC conversion hot path (SynthesizerTrn.infer) on B200.

    python bench.py --gpus N --steps K --warmup W                (our arm)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU arm: the reference's own PyTorch
                                                                  infer from baseline/_ref on all host threads;
                                                                  the oracle port when that is not staged)

One "step" = one `infer` over one batch of synthetic inputs.  Workload at every N: BASELINE.json
configs[1], batch 64 x 10 s utterances (T = 500 unit frames, one 10 s target mel), fp32 mode
(fp32 storage + accumulation, TF32 tensor-core operands), per GPU -- weak scaling, utterances are
independent so ranks share nothing.  `value` = audio-seconds converted per second by the whole job
with inputs resident in HBM; `e2e` = the same through the public Python API with pinned host buffers
(H2D of unit+mel and D2H of the waveform inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FLOP_PER_FRAME = 207.2e6          # conv/linear FLOPs per unit frame per utterance
FLOP_PER_WINDOW = 356.5e6         # speaker-encoder LSTM FLOPs per 128-frame mel window
TAIL_BYTES_PER_POST_FRAME = 128 * 4 + 16 * 4  # fused definition: 128-channel MRF output in (fp32), 16 fp32 samples out = 576,000 B per audio-second
DECODER_FLOP_PER_UTT_10S = 89.36e9            # decoder-only conv FLOPs per 10 s utterance


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="tf32", choices=["tf32", "bf16", "fp16", "fp32"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--mel-frames", type=int, default=500)
    ap.add_argument("--chunk-utts", type=int, default=0)
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="--impl reference: CPU seconds the whole run may take")
    ap.add_argument("--cpu-baseline-budget-s", type=float, default=25.0, help="our arm: CPU seconds of the cpu_baseline leg")
    ap.add_argument("--e2e-reps", type=int, default=3, help="timed end-to-end regions of --steps steps each (median reported)")
    ap.add_argument("--sweep-utts", type=int, default=4096, help="utterances of the configs[4] strong-scaling leg (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the bf16 / latency / tail side measurements")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def model_cfg():
    with open(os.path.join(ROOT, "tests", "golden", "quickvc_model_config.json")) as f:
        return json.load(f)


def n_windows(tm: int) -> int:
    return (tm - 128 + 63) // 64 + 1 if tm > 128 else 1


def random_init_state_dict(cfg):
    """Random-init weights of the reference architecture: the drop-in module's seeded constructor (equal to
    the reference's own under the same seed, tests/test_boundary.py) with the zero-initialised flow `post`
    layers (modules.py:196-197) re-drawn so the flow is not an identity."""
    import torch
    from quickvc_official_b200 import SynthesizerTrn
    torch.manual_seed(0)
    net = SynthesizerTrn(641, 32, **cfg)
    sd = net.state_dict()
    g = torch.Generator().manual_seed(7)
    for k in sd:
        if ".post." in k and k.startswith("flow."):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.05
    return sd


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled every 100 ms DURING the timed regions.  Reads NVML in-process (the library
    nvidia-smi itself reads; `nvidia_ml_py`): spawning an nvidia-smi process every 200 ms was measured to stall the
    launching thread for tens of milliseconds on some boxes -- a 5-step end-to-end region once read 2x too slow.  Falls
    back to nvidia-smi when NVML cannot be loaded."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes the visible devices; NVML enumerates all of them
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self._handle, n.NVML_CLOCK_SM)
        try:
            reasons = n.nvmlDeviceGetCurrentClocksEventReasons(self._handle)
        except Exception:
            reasons = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle)
        try:
            power = n.nvmlDeviceGetPowerUsage(self._handle) / 1000.0
        except Exception:
            power = 0.0
        flags = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        return [str(sm), str(mx), f"{power:.1f}"] + ["Active" if reasons & flags[k] else "Not Active"
                                                      for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")]

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_evt.is_set():
            try:
                if self._nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.1 if self._nvml is not None else 0.5)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = max((int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()), default=None)
        pw = max((float(r[2]) for r in self.rows if len(r) > 2), default=None)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(self.rows),
                    power_w_max=pw, source="nvml" if self._nvml is not None else "nvidia-smi")


# ------------------------------------------------------------------------------------------------
# CPU arm / baseline: the reference's own SynthesizerTrn.infer (baseline/_ref, staged by oracle/stage_reference.py;
# kind "reference"), else the oracle port of it (kind "port"); PyTorch fp32 on all host threads
# ------------------------------------------------------------------------------------------------
def cpu_infer_fn(sd, cfg):
    """Returns (callable(unit, mel, noise) -> waveform, kind, description)."""
    import torch
    from oracle import qvc_oracle, stage_reference
    if stage_reference.staged():
        models = stage_reference.load_reference_models()
        import contextlib
        import warnings
        with contextlib.redirect_stdout(sys.stderr), warnings.catch_warnings():   # the ctor prints; stdout carries the JSON line
            warnings.simplefilter("ignore")
            net = models.SynthesizerTrn(641, 32, **cfg).eval()
        net.load_state_dict(sd)                  # strict: the drop-in module's state_dict IS the reference's

        def run(unit, mel, noise):
            with torch.no_grad():
                return net.infer(unit, mel)      # the call convert.py:81 makes; it draws its own noise (models.py:94)
        return run, "reference", "the reference's own SynthesizerTrn.infer (unmodified files staged under baseline/_ref)"
    return (lambda unit, mel, noise: qvc_oracle.infer(sd, unit, mel, noise)), "port", \
        "oracle port of the reference's PyTorch infer (baseline/_ref not staged)"


def cpu_infer_rate(sd, cfg, frames, mel_frames, warmup, steps, budget_s, max_batch=64):
    """Times `steps` calls of the CPU implementation on ONE batch size chosen up front -- the largest of 64, 32, 16, 8, 4,
    2, 1 utterances for which warmup + steps calls fit `budget_s` seconds (probed with one 2-utterance call) -- so every
    step of a run does the same work."""
    import torch
    import synth
    torch.set_num_threads(os.cpu_count() or 1)
    run, kind, desc = cpu_infer_fn(sd, cfg)
    unit, mel, noise = synth.synthetic_inputs(2, frames, 1, mel_frames, 3)
    run(unit[:1], mel, noise[:1])                                   # first-call costs (thread pool, oneDNN primitives)
    t0 = time.perf_counter()
    run(unit, mel, noise)
    per_utt = (time.perf_counter() - t0) / 2
    batch = 1
    for b in (64, 32, 16, 8, 4, 2):
        if b <= max_batch and (warmup + steps) * b * per_utt * 0.8 <= budget_s:   # larger batches run a little faster per utterance
            batch = b
            break
    unit, mel, noise = synth.synthetic_inputs(batch, frames, 1, mel_frames, 3)
    for _ in range(warmup):
        run(unit, mel, noise)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        run(unit, mel, noise)
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch * frames / 50.0 / sec, sec, batch, kind, desc


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args, rank):
    if rank != 0:
        return
    cfg = model_cfg()
    sd = random_init_state_dict(cfg)
    rate, sec, sample, kind, desc = cpu_infer_rate(sd, cfg, args.frames, args.mel_frames, args.warmup, args.steps,
                                                   budget_s=args.cpu_budget_s, max_batch=args.batch)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": "audio-sec/sec", "value": rate, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE.json configs[1]: QuickVC SynthesizerTrn.infer, batch 64 x 10 s utterances, fp32, "
                               f"random-init weights, one 10 s target mel; each CPU step converts {sample} x "
                               f"{args.frames / 50:.0f} s utterances of it" + (" (the whole batch)" if sample == args.batch else
                                                                              f" (bounded sample: {args.cpu_budget_s:.0f} s budget)"),
                   "batch": sample, "frames": args.frames, "mel_frames": args.mel_frames, "same_config": sample == args.batch},
        "cpu_baseline": {"value": rate, "unit": "audio-s/s", "cores": cores, "kind": kind,
                         "sample": f"{sample} x {args.frames / 50:.0f} s utterances per step, {desc}, "
                                   f"torch.set_num_threads({cores}), CPU {cpu_model_name()}"},
        "e2e": {"value": rate, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Deadline:
    """Safety net of the measurement itself: if a section does not finish in its time (a hung kernel or collective on any
    rank), rank 0 prints the line with what was measured so far -- marked `incomplete` -- and every rank leaves.  A partial
    line beats a ten-minute collective timeout that ends with none."""

    current = None               # the instance of this process (main()'s exception path reads its partial line)

    def __init__(self, rank):
        import threading
        Deadline.current = self
        self._threading = threading
        self.rank = rank
        self.partial = None          # rank 0: the line so far; other ranks: {} once they have something to wait for
        self._timer = None

    def arm(self, seconds, what):
        seconds = seconds * float(os.environ.get("QVC_BENCH_DEADLINE_SCALE", "1"))     # tests shorten the deadlines
        self.disarm()
        self._timer = self._threading.Timer(seconds, self._fire, (seconds, what))
        self._timer.daemon = True
        self._timer.start()

    def disarm(self):
        if self._timer is not None:
            self._timer.cancel()
            self._timer = None

    def _fire(self, seconds, what):
        sys.stderr.write(f"bench.py: rank {self.rank}: '{what}' did not finish in {seconds} s -- giving up\n")
        sys.stderr.flush()
        if self.rank == 0 and self.partial:
            line = dict(self.partial)
            line["incomplete"] = f"'{what}' did not finish in {seconds} s; keys measured after it are absent"
            sys.stdout.write(json.dumps(line) + "\n")
            sys.stdout.flush()
        os._exit(0 if self.partial is not None else 3)


def run_sweep(args, make_net, dev, rank, world, dist, barrier, T, TM, slice_utts=64, reps=2):
    """configs[4]; every rank calls this.  Returns the result dict on rank 0, None elsewhere."""
    import torch
    from quickvc_official_b200.shard import convert_sharded, shard_range
    n = args.sweep_utts
    net = make_net("bf16")
    gen = torch.Generator(device=dev).manual_seed(5)                      # the same full batch on every rank
    unit = torch.randn(n, 256, T, device=dev, generator=gen)
    mel = torch.randn(1, 80, TM, device=dev, generator=gen) * 2 - 5
    lo, hi = shard_range(n, world, rank)
    calls = [0]

    def infer(u, m):                                                       # a rank's slice, 64 utterances per infer call
        outs = []
        for i in range(0, u.shape[0], slice_utts):
            outs.append(net.infer(u[i:i + slice_utts], m))
            calls[0] += 1
        return torch.cat(outs, 0) if len(outs) > 1 else outs[0]

    from quickvc_official_b200.shard import HostGather
    hg = HostGather(n, 320 * T)

    def once():
        # every rank converts its shard 64 utterances per call and copies each call's waveforms into ITS rows of one host
        # buffer shared by the ranks (page-locked in each), on a side stream, while the next call runs: the "final host
        # gather" with no device-to-device traffic and no rank-0 copy bottleneck
        hg.convert(lambda u, m: (calls.__setitem__(0, calls[0] + 1), net.infer(u, m))[1], unit, mel, chunk=slice_utts)

    def timed_sweep(fn):
        out = []
        for _ in range(reps):
            s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            s_ev.record()
            fn()
            e_ev.record()
            barrier()
            t = torch.tensor([s_ev.elapsed_time(e_ev)], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out.append(float(t.item()))
        return out

    convert_sharded(infer, unit[: world * slice_utts], mel)               # warm-up: folds the weights, sizes the workspace
    once()
    barrier()
    per = timed_sweep(once)
    hg.finish(dev)
    check = float(hg.host[n - 1].abs().sum()) if rank == 0 else 0.0       # the last rank's rows arrived in rank 0's view
    hg.close()
    # the round-1 form for comparison: torch.distributed gather of every waveform to rank 0's device, then ONE copy to
    # pinned host memory there
    host = torch.empty(n, 1, 320 * T).pin_memory() if rank == 0 else None

    def once_device_gather():
        out = convert_sharded(infer, unit, mel)
        if rank == 0:
            host.copy_(out, non_blocking=True)

    per_dev = timed_sweep(once_device_gather)
    del host
    ms = sum(per) / len(per)
    del unit, net
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    audio_s = n * T / 50.0
    ms_dev = sum(per_dev) / len(per_dev)
    return {"value": audio_s / (ms * 1e-3), "unit": "audio-s/s", "scaling": "strong", "n_gpus": world, "utterances": n,
            "ms_per_sweep": ms, "ms_min_max": [min(per), max(per)], "precision": "bf16 operands, f32 accumulate",
            "utterances_per_gpu": hi - lo, "infer_calls_per_gpu": (hi - lo + slice_utts - 1) // slice_utts,
            "d2h_bytes_per_rank": (hi - lo) * 320 * T * 4, "host_bytes_total": n * 320 * T * 4, "last_row_abs_sum": check,
            "device_gather_form": {"value": audio_s / (ms_dev * 1e-3), "ms_per_sweep": ms_dev,
                                   "gather_bytes_to_rank0": (n - (hi - lo)) * 320 * T * 4, "d2h_bytes_rank0": n * 320 * T * 4,
                                   "note": "shard.convert_sharded: torch.distributed gather to rank 0's device, then one copy "
                                           "to pinned host memory there (the round-1 form)"},
            "note": "BASELINE.json configs[4]: shard.HostGather.convert(infer, unit, mel): contiguous utterance shards, one "
                    "infer call per 64 utterances, no hot-path collective; the final host gather -- every rank's copy of its "
                    "waveforms into its rows of ONE host buffer shared by the ranks (page-locked, /dev/shm), overlapped with "
                    "the next call -- is inside the timed region (CUDA events on each rank, max over ranks)"}


def run_ours(args, rank, local_rank, world):
    import torch
    import synth  # noqa: F401
    from quickvc_official_b200 import SynthesizerTrn, capi

    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the image exports NCCL_DEBUG=VERSION, which makes NCCL print "NCCL version ..." on stdout; stdout carries the one
        # JSON line, so drop that level (an explicit INFO / TRACE request is left alone, with its log sent to stderr)
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    pk = peaks()
    cfg = model_cfg()
    sd = random_init_state_dict(cfg)

    def make_net(precision):
        net = SynthesizerTrn(641, 32, **cfg, precision=precision, chunk_utts=args.chunk_utts).eval()
        net.load_state_dict(sd)
        return net.to(dev)

    net = make_net(args.precision)
    B, T, TM = args.batch, args.frames, args.mel_frames
    g = torch.Generator().manual_seed(100 + rank)
    unit_h = torch.randn(B, 256, T, generator=g).pin_memory()
    mel_h = (torch.randn(1, 80, TM, generator=g) * 2 - 5).pin_memory()
    noise = torch.randn(B, 192, T, generator=g).to(dev)
    unit, mel = unit_h.to(dev), mel_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def barrier():
        # the device is drained BEFORE the collective is enqueued: a stuck kernel then shows up here (and in the Deadline
        # below), not as a collective that every rank waits ten minutes for
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()                 # device: the one bound at init_process_group(device_id=...)
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            flush.zero_()                       # L2 flush between timed iterations, outside the events
            starts[i].record()
            fn()
            ends[i].record()
        barrier()
        per = [s.elapsed_time(e) for s, e in zip(starts, ends)]
        tot = torch.tensor([sum(per)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()) / steps, per

    audio_s = B * T / 50.0
    sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get("QVC_BENCH_NO_SAMPLER") else None
    if sampler:
        sampler.start()
    deadline = Deadline(rank)
    deadline.arm(200, "device-resident timing")

    # ---- device-resident throughput ----
    l0 = capi.launch_count()
    ms, per = timed(lambda: net.infer(unit, mel, noise=noise), args.steps, args.warmup)
    launches = (capi.launch_count() - l0) // (args.steps + args.warmup)
    value = world * audio_s / (ms * 1e-3)
    workload = ("BASELINE.json configs[1]: QuickVC SynthesizerTrn.infer, batch 64 x 10 s utterances, fp32 mode, "
                "random-init weights, one 10 s target mel")
    dtype_name = {"tf32": "tf32 operands, f32 accumulate/storage", "bf16": "bf16 operands, f32 accumulate",
                  "fp16": "fp16 operands (10-bit mantissa, as TF32), f32 accumulate", "fp32": "f32"}[args.precision]
    deadline.partial = {} if rank != 0 else {
        "metric": "audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype_name, "data": "synthetic",
        "config": {"workload": workload, "batch_per_gpu": B, "frames": T, "mel_frames": TM, "precision": args.precision,
                   "l2": "flushed (256 MB memset) between timed steps", "audio_seconds_per_step_per_gpu": audio_s},
        "clocks": None, "e2e": None, "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches),
        "x_realtime_per_gpu": value / world, "step_ms_min_max": [min(per), max(per)]}
    deadline.arm(180, "end-to-end timing (PipelinedConverter)")

    # ---- end to end through the public API: pinned host in, pinned host out, every step ----
    # quickvc_official_b200.pipeline.PipelinedConverter = H2D of unit + mel, net.infer(unit, mel), D2H of the waveform
    # on three streams with double buffers (the copies of neighbouring steps overlap the kernels).  Timed as one
    # region: start event, K submits, drain (all D2H done), end event.  No L2 flush inside the region: the 4.6 GB
    # working set of a step is 36x the L2.
    from quickvc_official_b200.pipeline import PipelinedConverter
    conv = PipelinedConverter(net, B, T, TM, device=dev)

    def e2e_run(n):
        for _ in range(n):
            conv.submit(unit_h, mel_h)
        conv.drain()

    e2e_run(args.warmup)
    e2e_all = []
    for _ in range(max(1, args.e2e_reps)):          # the region is repeated; the MEDIAN region is reported (all are listed)
        barrier()
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_ev.record()
        e2e_run(args.steps)
        e_ev.record()
        barrier()
        tot = torch.tensor([s_ev.elapsed_time(e_ev)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        e2e_all.append(float(tot.item()) / args.steps)
    ms_e2e = sorted(e2e_all)[len(e2e_all) // 2]
    e2e_value = world * audio_s / (ms_e2e * 1e-3)
    wave_h = conv._wave_h[0]
    clocks = sampler.stop() if sampler else None

    # ---- BASELINE.json configs[4]: throughput sweep, 4096 utterances x 10 s sharded over the N GPUs (STRONG scaling: the
    # job is fixed), bf16 operands with fp32 accumulation, through quickvc_official_b200.shard.convert_sharded with the
    # gather of all waveforms to rank 0 (and their copy to pinned host memory there) inside the timed region.
    if rank == 0:
        deadline.partial["clocks"] = clocks
        deadline.partial["e2e"] = {"value": e2e_value, "unit": "audio-s/s", "ms_per_step": ms_e2e, "ms_per_step_regions": e2e_all,
                                   "h2d_bytes_per_step": unit_h.numel() * 4 + mel_h.numel() * 4,
                                   "d2h_bytes_per_step": wave_h.numel() * 4}
    sweep = None
    if args.sweep_utts > 0:
        deadline.arm(300, "configs[4] sweep")
        if world > 1:
            sweep = run_sweep(args, make_net, dev, rank, world, dist, barrier, T, TM)
        else:
            try:                                   # one rank: nobody waits in a collective, so a failure costs this key only
                sweep = run_sweep(args, make_net, dev, rank, world, dist, barrier, T, TM)
            except Exception as exc:               # noqa: BLE001
                import traceback
                traceback.print_exc(file=sys.stderr)
                sweep = {"failed": f"{type(exc).__name__}: {exc}"[:400]}
                torch.cuda.empty_cache()
    deadline.arm(900, "roofline / side measurements / CPU baseline")

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        deadline.disarm()
        return

    # ---- roofline of the dominant kernel: conv_tc_kernel, the tcgen05 series convolution (99.9 % of the
    # FLOPs, 113 of the 127 launches of a step).  Its launches are timed live with CUDA events on the
    # launching stream (qvc_profile, include/qvc_b200.h) over `prof_steps` further steps of the same
    # workload; achieved = algorithmic conv FLOPs of a step / summed duration of its conv launches.
    import ctypes as C
    lib = capi.load()
    prof_steps = 3
    flush.zero_()
    torch.cuda.synchronize()
    capi.check(lib.qvc_profile(1), "qvc_profile")
    for _ in range(prof_steps):
        net.infer(unit, mel, noise=noise)
    torch.cuda.synchronize()
    capi.check(lib.qvc_profile(0), "qvc_profile")
    conv_ms, conv_n = C.c_double(0.0), C.c_uint64(0)
    capi.check(lib.qvc_profile_read(C.byref(conv_ms), C.byref(conv_n)), "qvc_profile_read")
    conv_ms_step = conv_ms.value / prof_steps
    conv_launches = conv_n.value // prof_steps
    conv_flops = B * T * FLOP_PER_FRAME
    flops = conv_flops + n_windows(TM) * FLOP_PER_WINDOW
    achieved = conv_flops / (conv_ms_step * 1e-3) / 1e12
    tf32 = args.precision not in ("bf16", "fp16")       # 2-byte operands run at the bf16 tensor rate
    # Denominator: the driver's MEASURED_PEAKS.json, sustained figure (the kernels are timed inside a long step under
    # the power cap); TF32 operands run at half the bf16 tensor rate.  The builder's own cuBLAS-TF32 measurement on this
    # pool (profiles/r01_tf32_peak.json) is reported beside it as a secondary reading.
    peak = pk["bf16_sustained"] * (0.5 if tf32 else 1.0)
    peak_src = pk["source"] + ": bf16_tflops_sustained (kernels timed inside a long step)" + \
        (" x 0.5 -- TF32 operands run at half the bf16 tensor rate" if tf32 else "")
    cublas_tf32 = None
    tf32_path = os.path.join(ROOT, "profiles", "r01_tf32_peak.json")
    if tf32 and os.path.exists(tf32_path):
        with open(tf32_path) as f:
            cublas_tf32 = json.load(f)["tf32_tflops_sustained"]
    traffic, traffic_note, tj = None, None, None
    tpath = next((q for q in (os.path.join(ROOT, "profiles", f"r02_traffic_{args.precision}.json"),
                              os.path.join(ROOT, "profiles", "r01_traffic.json")) if os.path.exists(q)), None)
    if tpath and (args.precision == "tf32" or "r02" in tpath) and B == 64 and T == 500:
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj["conv_tc"]["dram_bytes_per_launch"]
        traffic_note = ("DRAM bytes (read + write) per launch, averaged over the kernel's launches of one step: " + tj["source"])
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
        "traffic_note": traffic_note,
        "kernel": "conv_tcr_kernel + conv_tc2_kernel + conv_tc_kernel + post_tail_kernel (tcgen05 implicit-GEMM series convolution: "
                  "CTA pairs with frames on the accumulator rows, CTA pairs with channels on them, cta_group::1, and the post-net "
                  "convolution whose epilogue is the iSTFT / overlap-add / synthesis tail)",
        "launches_per_step": int(conv_launches), "avg_launch_us": conv_ms_step * 1e3 / max(1, conv_launches),
        "kernel_ms_per_step": conv_ms_step, "kernel_share_of_step": conv_ms_step / ms,
        "flops_per_step": conv_flops,
        "peak_source": peak_src,
        "frac_of_cublas_tf32_sustained": (achieved / cublas_tf32) if cublas_tf32 else None,
        "cublas_tf32_sustained": cublas_tf32,
        "frac_of_bf16_sustained": achieved / pk["bf16_sustained"],
        "whole_step": {"achieved": flops / (ms * 1e-3) / 1e12, "frac": flops / (ms * 1e-3) / 1e12 / peak, "flops": flops},
        "note": "algorithmic FLOPs (207.2 MFLOP per unit frame per utterance, SURVEY.md section 8d) over CUDA-event launch "
                "durations summed across the kernel's launches of one step; zero-padded polyphase taps and padded output "
                "channels are not counted as work; the event pairs between launches switch programmatic dependent launch off, so "
                "kernel_ms_per_step is an upper bound of the kernels' time inside a plain step (kernel_share_of_step can pass 1 in "
                "the 16-bit modes); per-layer ncu captures: profiles/",
    }

    line = {
        "metric": "audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"tf32": "tf32 operands, f32 accumulate/storage", "bf16": "bf16 operands, f32 accumulate",
                  "fp16": "fp16 operands (10-bit mantissa, as TF32), f32 accumulate", "fp32": "f32"}[args.precision],
        "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[1]: QuickVC SynthesizerTrn.infer, batch 64 x 10 s utterances, fp32 mode, "
                               "random-init weights, one 10 s target mel", "batch_per_gpu": B, "frames": T, "mel_frames": TM,
                   "precision": args.precision, "l2": "flushed (256 MB memset) between timed steps",
                   "audio_seconds_per_step_per_gpu": audio_s},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "ms_per_step": ms_e2e, "ms_per_step_regions": e2e_all,
                "h2d_bytes_per_step": unit_h.numel() * 4 + mel_h.numel() * 4, "d2h_bytes_per_step": wave_h.numel() * 4,
                "api": "quickvc_official_b200.pipeline.PipelinedConverter.submit(unit_host, mel_host) -> net.infer(unit, mel); "
                       "pinned host buffers, H2D / kernels / D2H of consecutive steps overlapped on three streams"},
        "gpu_launches": int(launches) * args.steps,
        "gpu_launches_per_step": int(launches),
        "roofline": roofline,
        "x_realtime_per_gpu": value / world,
        "step_ms_min_max": [min(per), max(per)],
    }
    if sweep is not None:
        line["sweep_4096_bf16"] = sweep
    deadline.partial = line                       # from here on the safety net prints the full line with whatever keys it has

    def section(name, fn):
        # a side measurement that fails (or a box that cannot run it) costs its own key, never the headline line
        try:
            fn()
        except Exception as exc:      # noqa: BLE001
            import traceback
            traceback.print_exc(file=sys.stderr)
            line.setdefault("failed_sections", {})[name] = f"{type(exc).__name__}: {exc}"[:400]
            try:
                torch.cuda.synchronize()
            except Exception:         # noqa: BLE001
                pass

    def extra_tail():
        # the fused post-net convolution + iSTFT / overlap-add / synthesis kernel (post_tail.cu) alone, on the HBM roofline of
        # SURVEY.md section 8d's fused definition: the 128-channel operand series in, the waveform out
        frames_post = 20 * T + 1
        model = net._engine._ensure_model(dev)
        stream = torch.cuda.current_stream().cuda_stream
        esz = 4 if args.precision in ("tf32", "fp32") else 2
        odt = {"tf32": torch.float32, "fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}[args.precision]
        xin = (torch.randn(B, frames_post, 128, device=dev) * 0.3).to(odt)
        wave = torch.empty(B, 1, 320 * T, device=dev)
        L = model.layers[capi.QVC_NUM_LAYERS - 1]
        pa = capi.ConvArgs()
        pa.x = capi.Tensor(xin.data_ptr(), frames_post * 128, 128, 0)
        pa.batch, pa.x_rows, pa.out_rows, pa.cin = B, frames_post, frames_post, 128
        pa.w, pa.bias, pa.bias_bstride = L.w, L.bias, 0
        pa.cout, pa.k, pa.dil, pa.pad_left = L.cout, L.k, L.dil, L.pad_left
        pa.epilogue, pa.nseg = 0, 1
        pa.seg[0].col0, pa.seg[0].ncols, pa.seg[0].alpha, pa.seg[0].beta, pa.seg[0].slope = 0, 72, 1.0, 1.0, 1.0
        pa.opformat, pa.backend = model.opformat, model.backend
        tail_bytes = B * frames_post * TAIL_BYTES_PER_POST_FRAME
        if lib.qvc_post_tail(C.byref(pa), C.byref(model.tail), None, 0, wave.data_ptr(), None, stream) == 0:
            ms_tail, _ = timed(lambda: capi.check(lib.qvc_post_tail(C.byref(pa), C.byref(model.tail), None, 0, wave.data_ptr(),
                                                                    None, stream), "qvc_post_tail"), 20, 3)
            line["tail_roofline"] = {
                "bound": "hbm", "achieved": tail_bytes / (ms_tail * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                "frac": tail_bytes / (ms_tail * 1e-3) / 1e9 / pk["hbm"],
                "traffic": (tj["tail"]["dram_bytes_per_launch"] if tj is not None and "tail" in tj else None), "ms": ms_tail,
                "algorithmic_bytes": tail_bytes, "actual_bytes": B * frames_post * (128 * esz + 64),
                "kernel": "post_tail_kernel: subband_conv_post (tcgen05, frames on the accumulator rows) with the magnitude exp / "
                          "phase sin-cos, 16-point inverse DFT, windowed overlap-add and synthesis FIR as its epilogue",
                "note": "measured issue-bound (polar + DFT + 272 FMAs of synthesis FIR per sub-band sample on 8 epilogue warps), "
                        "not HBM-bound: profiles/r02_summary.md",
                "peak_source": pk["source"]}
    def extra_modes():
        # the separately reported bf16 / fp16 / strict-fp32 modes
        if args.precision == "tf32":
            nb = make_net("bf16")
            ms_b, _ = timed(lambda: nb.infer(unit, mel, noise=noise), max(3, args.steps // 2), 2)
            line["bf16_mode"] = {"value": audio_s / (ms_b * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_b,
                                 "roofline_frac_of_bf16_sustained": flops / (ms_b * 1e-3) / 1e12 / pk["bf16_sustained"],
                                 "tolerance": "waveform max-abs 2e-3, per-stage rel-L2 1.5e-2 (tests/test_gpu_infer.py)"}
            del nb
            # fp16 operands: TF32's mantissa in two bytes -- the fp32-mode tolerance at the bf16 tensor rate
            nh = make_net("fp16")
            ms_h, _ = timed(lambda: nh.infer(unit, mel, noise=noise), max(3, args.steps // 2), 2)
            line["fp16_mode"] = {"value": audio_s / (ms_h * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_h,
                                 "roofline_frac_of_bf16_sustained": flops / (ms_h * 1e-3) / 1e12 / pk["bf16_sustained"],
                                 "tolerance": "the fp32-mode bound: waveform max-abs 1e-4, per-stage rel-L2 1e-3 "
                                              "(tests/test_gpu_infer.py; measured 5.9e-5 / 5.9e-4)"}
            del nh
            # strict fp32 (CUDA-core FMA back end, no operand rounding at all): the cross-check mode, for scale
            nf = make_net("fp32")
            ms_f, _ = timed(lambda: nf.infer(unit, mel, noise=noise), 2, 1)
            line["fp32_strict_mode"] = {"value": audio_s / (ms_f * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_f,
                                        "note": "exact fp32 FMA kernels (waveform max-abs 1.5e-7 vs the reference); not the "
                                                "product path, kept to validate the tensor-core one"}
            del nf
    def extra_decoder():
        # ---- BASELINE.json configs[2]: the decoder alone (Multistream_iSTFT_Generator: conv_pre, ConvTranspose / MRF
        # ResBlocks, conv_post, fused iSTFT / OLA / sub-band synthesis) at batch 256 x 10 s, through net.decode(z, g)
        DEC_B = 256
        dec_flops = DEC_B * DECODER_FLOP_PER_UTT_10S * (T / 500.0)
        gz = torch.Generator(device=dev).manual_seed(9)
        z256 = torch.randn(DEC_B, 192, T, device=dev, generator=gz)
        g256 = torch.nn.functional.normalize(torch.randn(1, 256, device=dev, generator=gz), dim=1)
        dec = {}
        for prec in ("tf32", "fp16", "bf16"):
            nd = net if prec == args.precision else make_net(prec)
            ms_d, _ = timed(lambda: nd.decode(z256, g256), 5, 3)
            pk_d = pk["bf16_sustained"] * (0.5 if prec == "tf32" else 1.0)
            dec[prec] = {"value": DEC_B * T / 50.0 / (ms_d * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_d,
                         "roofline": {"bound": "tensor", "achieved": dec_flops / (ms_d * 1e-3) / 1e12, "peak": pk_d,
                                      "unit": "TFLOP/s", "frac": dec_flops / (ms_d * 1e-3) / 1e12 / pk_d}}
            if nd is not net:
                del nd
        line["decoder_only_b256"] = {
            "workload": f"BASELINE.json configs[2]: net.decode(z (256,192,{T}), g (1,256)) -- decoder only, 256 x 10 s",
            "flops_per_step": dec_flops, "modes": dec,
            "note": "whole-call CUDA-event time (L2 flushed between steps); 89.36 GFLOP per 10 s utterance (SURVEY.md section "
                    "8d); peak = MEASURED_PEAKS.json bf16_tflops_sustained (x 0.5 for TF32 operands)"}
        del z256
        torch.cuda.empty_cache()

    def extra_latency():
        # ---- single-call latency: the 5 s clip of the metric, and BASELINE.json configs[3] (0.5 s chunks, batch 1) in the
        # fp32 mode (tf32), fp16 and bf16: p50 / p99 over 1000 calls each, through infer(unit, mel) as convert.py calls it
        # and with the target-speaker embedding cached (the reference recomputes it on every call, models.py:635)
        def latency(fn, n=1000, skip=20):
            lat = []
            for i in range(n + skip):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn()
                e.record()
                torch.cuda.synchronize()
                if i >= skip:
                    lat.append(s.elapsed_time(e))
            lat.sort()
            return {"p50": lat[len(lat) // 2], "p99": lat[min(len(lat) - 1, int(len(lat) * 0.99))], "calls": len(lat)}

        from quickvc_official_b200.pipeline import GraphedInfer
        m5 = mel[:, :, :250].contiguous()
        for name, frames in (("latency_5s_clip_ms", 250), ("latency_0p5s_chunk_ms", 25)):
            u1, n1 = unit[:1, :, :frames].contiguous(), noise[:1, :, :frames].contiguous()
            entry = {}
            for prec in ("tf32", "fp16", "bf16"):
                nl = net if prec == args.precision else make_net(prec)
                emb = nl.embed_speaker(m5)
                r = latency(lambda: nl.infer(u1, m5, noise=n1))
                r["cached_speaker"] = latency(lambda: nl.infer_with_embedding(u1, emb, noise=n1))
                gi = GraphedInfer(nl, 1, frames, mel_frames=0, device=dev)
                gi(u1, emb, n1)
                r["cached_speaker_cuda_graph"] = latency(lambda: gi(u1, emb, n1))
                entry[prec] = r
                del gi
                if nl is not net:
                    del nl
            entry.update(entry[args.precision if args.precision in entry else "tf32"])      # top-level p50 / p99: this run's mode
            line[name] = entry

    def extra_ragged():
        # SURVEY.md section 8f "next" #3: ragged batches -- the same B utterances with lengths drawn from 5 .. 10 s, sorted
        # as the conversion driver does, padded to the longest; the rate counts live audio only
        lens = torch.sort(torch.randint(T // 2, T + 1, (B,), generator=torch.Generator().manual_seed(3))).values
        lens[-1] = T
        lens_dev = lens.to(dev, torch.int32)
        ms_rag, _ = timed(lambda: net.infer(unit, mel, noise=noise, lengths=lens_dev), max(3, args.steps // 2), 3)
        live_s = float(lens.sum()) / 50.0
        line["ragged_batch"] = {"value": live_s / (ms_rag * 1e-3), "unit": "audio-s/s", "ms_per_step": ms_rag,
                                "live_audio_s": live_s, "padded_audio_s": B * T / 50.0,
                                "note": f"infer(unit, mel, lengths=...) on {B} utterances of 5..10 s padded to 10 s: each gets "
                                        "exactly the samples of its own single-utterance call (tests/test_gpu_ragged.py); "
                                        "tiles wholly past an utterance's end are skipped by every warp role"}

    def extra_mel():
        # SURVEY.md section 8f "next" #1: the target-mel front end (wave_to_mel, convert.py:75-77) for one 10 s target
        from quickvc_official_b200 import mel as qmel
        from oracle import mel_oracle
        margs = (1280, 80, 16000, 320, 1280, 0.0, None)
        wav_h = (torch.rand(1, 160000) * 2 - 1) * 0.5
        wav = wav_h.to(dev)
        ms_mel, _ = timed(lambda: qmel.wave_to_mel(wav, *margs), 20, 3)
        t0 = time.perf_counter()
        for _ in range(5):
            mel_oracle.wave_to_mel(wav_h, *margs)
        cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
        stft_flops = 2.0 * 500 * 1280 * 1296
        line["mel_frontend"] = {"ms": ms_mel, "audio_s_per_s": 10.0 / (ms_mel * 1e-3), "cpu_oracle_ms": cpu_ms,
                                "stft_gemm_tflops": stft_flops / (ms_mel * 1e-3) / 1e12,
                                "note": "wave_to_mel of one 10 s target utterance (B=1): reflect pad + STFT as an exact-fp32 FMA "
                                        "GEMM (1.66 GFLOP) + magnitude + Slaney mel + log; latency bound (500 frames); CPU = "
                                        "torch.stft oracle on the host threads"}

    if not args.no_extras and world == 1:
        for name, fn in (("tail_roofline", extra_tail), ("precision_modes", extra_modes), ("decoder_only_b256", extra_decoder),
                         ("latency", extra_latency), ("ragged_batch", extra_ragged), ("mel_frontend", extra_mel)):
            section(name, fn)

    def cpu_baseline():
        cores = os.cpu_count() or 1
        rate, sec, sample, kind, desc = cpu_infer_rate(sd, cfg, T, TM, 1, 3, budget_s=args.cpu_baseline_budget_s, max_batch=B)
        line["cpu_baseline"] = {"value": rate, "unit": "audio-s/s", "cores": cores, "kind": kind,
                                "sample": f"{sample} x {T / 50:.0f} s utterances per call, {desc} (fp32, "
                                          f"torch.set_num_threads({cores})), 1 warm-up + 3 timed calls of {sec:.2f} s, "
                                          f"CPU {cpu_model_name()}"}

    if not args.no_cpu_baseline and world == 1:
        section("cpu_baseline", cpu_baseline)
    deadline.disarm()
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    try:
        run_ours(args, rank, local_rank, world)
    except Exception:      # noqa: BLE001
        # a failure after the headline was measured still ends with that line on stdout (marked `incomplete`) and rc 0; the
        # other ranks of a multi-GPU run then leave through their own deadline.  Before it, the failure is the result.
        import traceback
        traceback.print_exc(file=sys.stderr)
        sys.stderr.flush()
        d = Deadline.current
        if d is not None and d.partial is not None:
            d.disarm()
            if rank == 0 and d.partial:
                line = dict(d.partial)
                line["incomplete"] = "an exception ended the run after these keys were measured (traceback on stderr)"
                sys.stdout.write(json.dumps(line) + "\n")
                sys.stdout.flush()
            os._exit(0)
        raise


if __name__ == "__main__":
    main()
