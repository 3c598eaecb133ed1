"""Fakes that let bench.py's control flow run on a CPU-only host (TEST ONLY, tests/test_bench_dryrun.py).

Nothing here measures anything: the device is the CPU, CUDA events return a constant, the module's entry points return
zeros of the right shapes, and the C library is a stub.  What the dry run checks is the Python around the measurements --
every section runs, every key the driver and the judge read is present and well formed, a failing section costs its own key
only -- so that a slip in bench.py shows up here and not as a missing line at the end of a round."""
import ctypes as C
import os
import types

import torch

_real_device = torch.device
CPU = _real_device("cpu")


class FakeEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return 2.0                                       # ms


class FakeStream:
    cuda_stream = 0

    def __init__(self, device=None):
        pass

    def wait_event(self, e):
        pass

    def wait_stream(self, s):
        pass


class FakeLib:
    """The entry points bench.py calls directly."""

    def qvc_profile(self, on):
        return 0

    def qvc_profile_read(self, ms_ref, n_ref):
        ms_ref._obj.value = 3 * 1.5                      # 3 profiled steps of 1.5 ms of convolution launches
        n_ref._obj.value = 3 * 110
        return 0

    def qvc_post_tail(self, *a):
        if os.environ.get("QVC_FAKE_FAIL") == "tail":
            raise RuntimeError("injected failure in the tail section")
        return 0


def install():
    import quickvc_official_b200 as pkg
    from quickvc_official_b200 import capi, engine, mel as qmel, pipeline

    torch.device = lambda *a, **k: CPU                   # "cuda:N" -> cpu: every .to(dev) / device=dev stays on the host
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda d=None: None
    torch.cuda.empty_cache = lambda: None
    torch.cuda.Event = FakeEvent
    torch.cuda.Stream = FakeStream
    torch.cuda.current_stream = lambda d=None: FakeStream()

    # multi-rank runs: the same collectives over gloo (NCCL needs GPUs)
    import torch.distributed as dist
    real_init, real_barrier = dist.init_process_group, dist.barrier
    dist.init_process_group = lambda backend=None, **kw: real_init("gloo")
    dist.barrier = lambda *a, **kw: real_barrier()

    launches = [0]
    lib = FakeLib()
    capi.load = lambda: lib
    capi.launch_count = lambda: launches[0]
    capi.check = lambda status, what: None

    def infer(self, unit, mel, *, noise=None, taps=None, lengths=None):
        launches[0] += 123
        return torch.zeros(unit.shape[0], 1, 320 * unit.shape[2])

    def decode(self, z, g, **kw):
        return torch.zeros(z.shape[0], 1, 320 * z.shape[2])

    pkg.SynthesizerTrn.infer = infer
    pkg.SynthesizerTrn.decode = decode
    pkg.SynthesizerTrn.embed_speaker = lambda self, mel: torch.zeros(1, 256)
    pkg.SynthesizerTrn.infer_with_embedding = lambda self, unit, g, **kw: torch.zeros(unit.shape[0], 1, 320 * unit.shape[2])

    def ensure_model(self, device):
        m = capi.Model()
        m.opformat, m.backend = capi.OPF_TF32, capi.BACKEND_TCGEN05
        return m

    engine.InferEngine._ensure_model = ensure_model

    class Conv:
        def __init__(self, net, batch, frames, mel_frames, device=None, depth=2):
            self.net = net
            self._wave_h = [torch.zeros(batch, 1, 320 * frames)]

        def submit(self, unit_h, mel_h, noise=None):
            if os.environ.get("QVC_FAKE_FAIL") == "e2e_rank1" and os.environ.get("RANK") == "1":
                raise RuntimeError("injected failure on rank 1")
            if os.environ.get("QVC_FAKE_FAIL") == "hang_rank1" and os.environ.get("RANK") == "1":
                import time
                time.sleep(3600)                                   # a kernel that never finishes
            return self.net.infer(unit_h, mel_h), FakeEvent()

        def drain(self):
            pass

    class Graphed:
        def __init__(self, net, batch, frames, mel_frames=0, device=None):
            self.out = torch.zeros(batch, 1, 320 * frames)

        def __call__(self, *a):
            return self.out

    pipeline.PipelinedConverter = Conv
    pipeline.GraphedInfer = Graphed
    qmel.wave_to_mel = lambda wav, *a: torch.zeros(wav.shape[0], 80, wav.shape[1] // 320)
