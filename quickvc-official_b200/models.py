"""Drop-in mirror of the reference generator class for the conversion path.

`SynthesizerTrn` here has the reference's constructor signature (/root/reference/models.py:551-569),
the reference's 467-entry `state_dict` layout (old-style `weight_g` / `weight_v` names, the
`dec.updown_filter` and `dec.stft.window` buffers, and all training-only `enc_q.*` entries) and the
reference's `infer(unit, mel)` (/root/reference/models.py:625-642).  The compute is not PyTorch:
`infer` hands raw device pointers to libqvc_b200.so (include/qvc_b200.h), whose sm_100a kernels run
the whole path.  There is no CPU or eager fallback; without the library or a B200 `infer` raises.

The submodules below are parameter containers only (same names, shapes, registration order and
default initialisation draws as the reference's torch modules) -- they have no forward.
"""
from __future__ import annotations

import logging
import math
from typing import Dict, List, Optional

import torch
from torch import Tensor, nn
from torch.nn import init

from . import capi
from .engine import InferEngine

_log = logging.getLogger("quickvc_b200")


# --------------------------------------------------------------------------------------------
# parameter containers
# --------------------------------------------------------------------------------------------
class _ConvParams(nn.Module):
    """Parameters of a (transposed) 1-d convolution, optionally in old-style weight-norm form.

    Registration order (bias, weight_g, weight_v) and the random draws (kaiming-uniform weight,
    then uniform bias) follow nn.Conv1d + torch.nn.utils.weight_norm, so a seeded construction
    reproduces the reference's initial weights and `state_dict()` key order.
    """

    def __init__(self, cin: int, cout: int, k: int, *, weight_norm: bool, bias: bool = True,
                 transposed: bool = False, extra_init_draw: bool = False) -> None:
        super().__init__()
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        w = torch.empty(shape)
        init.kaiming_uniform_(w, a=math.sqrt(5))
        b = None
        if bias:
            fan_in = shape[1] * k
            bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
            b = torch.empty(cout)
            init.uniform_(b, -bound, bound)
        if weight_norm:
            if b is not None:
                self.bias = nn.Parameter(b)
            g = w.reshape(shape[0], -1).norm(dim=1).reshape(shape[0], 1, 1)
            self.weight_g = nn.Parameter(g)
            self.weight_v = nn.Parameter(w)
        else:
            self.weight = nn.Parameter(w)
            if b is not None:
                self.bias = nn.Parameter(b)
        if extra_init_draw:
            # the reference calls init_weights (commons.py:8-11) on weight-normed convs: with the
            # old-style hook that only overwrites the derived `.weight`, i.e. it consumes
            # numel normal draws and changes nothing.  Consume the same draws.
            torch.empty(shape).normal_(0.0, 0.01)


class _Empty(nn.Module):
    """Parameter-free placeholder (Flip, modules.py:165)."""


class _WN(nn.Module):
    """modules.py:37-67."""

    def __init__(self, hidden: int, k: int, n_layers: int, gin: int) -> None:
        super().__init__()
        if gin != 0:
            self.cond_layer = _ConvParams(gin, 2 * hidden * n_layers, 1, weight_norm=True)
        self.in_layers = nn.ModuleList()
        self.res_skip_layers = nn.ModuleList()
        for i in range(n_layers):
            self.in_layers.append(_ConvParams(hidden, 2 * hidden, k, weight_norm=True))
            rs = 2 * hidden if i < n_layers - 1 else hidden
            self.res_skip_layers.append(_ConvParams(hidden, rs, 1, weight_norm=True))


class _CondNormalWN(nn.Module):
    """models.py:54-73."""

    def __init__(self, cin: int, cout: int, hidden: int, k: int, n_layers: int, gin: int) -> None:
        super().__init__()
        self.pre = _ConvParams(cin, hidden, 1, weight_norm=False)
        self.enc = _WN(hidden, k, n_layers, gin)
        self.proj = _ConvParams(hidden, 2 * cout, 1, weight_norm=False)


class _Coupling(nn.Module):
    """modules.py:175-197 (post is zero-initialised)."""

    def __init__(self, channels: int, hidden: int, k: int, n_layers: int, gin: int) -> None:
        super().__init__()
        half = channels // 2
        self.pre = _ConvParams(half, hidden, 1, weight_norm=False)
        self.enc = _WN(hidden, k, n_layers, gin)
        self.post = _ConvParams(hidden, half, 1, weight_norm=False)
        with torch.no_grad():
            self.post.weight.zero_()
            self.post.bias.zero_()


class _Flow(nn.Module):
    """models.py:17-37."""

    def __init__(self, channels: int, hidden: int, k: int, n_layers: int, n_flows: int, gin: int) -> None:
        super().__init__()
        self.flows = nn.ModuleList()
        for _ in range(n_flows):
            self.flows.append(_Coupling(channels, hidden, k, n_layers, gin))
            self.flows.append(_Empty())


class _LSTMParams(nn.Module):
    """nn.LSTM(80, 256, 3) parameters (models.py:510)."""

    def __init__(self, cin: int, hidden: int, layers: int) -> None:
        super().__init__()
        bound = 1.0 / math.sqrt(hidden)
        for l in range(layers):
            for name, shape in ((f"weight_ih_l{l}", (4 * hidden, cin if l == 0 else hidden)),
                                (f"weight_hh_l{l}", (4 * hidden, hidden)),
                                (f"bias_ih_l{l}", (4 * hidden,)), (f"bias_hh_l{l}", (4 * hidden,))):
                self.register_parameter(name, nn.Parameter(torch.empty(shape)))
        for p in self.parameters():
            init.uniform_(p, -bound, bound)


class _LinearParams(nn.Module):
    def __init__(self, cin: int, cout: int) -> None:
        super().__init__()
        w = torch.empty(cout, cin)
        init.kaiming_uniform_(w, a=math.sqrt(5))
        b = torch.empty(cout)
        init.uniform_(b, -1.0 / math.sqrt(cin), 1.0 / math.sqrt(cin))
        self.weight = nn.Parameter(w)
        self.bias = nn.Parameter(b)


class _SpeakerEncoder(nn.Module):
    """models.py:507-512."""

    def __init__(self, mel_n_channels: int = 80, model_num_layers: int = 3, model_hidden_size: int = 256,
                 model_embedding_size: int = 256) -> None:
        super().__init__()
        self.lstm = _LSTMParams(mel_n_channels, model_hidden_size, model_num_layers)
        self.linear = _LinearParams(model_hidden_size, model_embedding_size)


class _ResBlock1(nn.Module):
    """modules.py:128-145."""

    def __init__(self, ch: int, k: int) -> None:
        super().__init__()
        self.convs1 = nn.ModuleList([_ConvParams(ch, ch, k, weight_norm=True) for _ in range(3)])
        for _ in range(3):
            torch.empty(ch, ch, k).normal_(0.0, 0.01)      # convs1.apply(init_weights): draws only
        self.convs2 = nn.ModuleList([_ConvParams(ch, ch, k, weight_norm=True) for _ in range(3)])
        for _ in range(3):
            torch.empty(ch, ch, k).normal_(0.0, 0.01)


class _Window(nn.Module):
    """InverseSpectrogram(16, 16, 4) keeps only its window buffer (models.py:350)."""

    def __init__(self, n_fft: int) -> None:
        super().__init__()
        self.register_buffer("window", torch.hann_window(n_fft))


class _MultistreamGenerator(nn.Module):
    """models.py:304-358."""

    def __init__(self, initial_channel, resblock_kernel_sizes, resblock_dilation_sizes, upsample_rates,
                 upsample_initial_channel, upsample_kernel_sizes, n_fft, hop_istft, subbands, gin_channels) -> None:
        super().__init__()
        self.conv_pre = _ConvParams(initial_channel, upsample_initial_channel, 7, weight_norm=True)
        self.cond = _ConvParams(gin_channels, upsample_initial_channel, 1, weight_norm=False)
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)):
            self.ups.append(_ConvParams(upsample_initial_channel // (2 ** i), upsample_initial_channel // (2 ** (i + 1)),
                                        k, weight_norm=True, transposed=True))
        for i, k in enumerate(upsample_kernel_sizes):      # self.ups.apply(init_weights): draws only
            torch.empty(upsample_initial_channel // (2 ** i), upsample_initial_channel // (2 ** (i + 1)), k).normal_(0.0, 0.01)
        self.resblocks = nn.ModuleList()
        ch = 0
        for i in range(len(upsample_rates)):
            ch = upsample_initial_channel // (2 ** (i + 1))
            for k in resblock_kernel_sizes:
                self.resblocks.append(_ResBlock1(ch, k))
        n_freq = n_fft // 2 + 1
        self.subband_conv_post = _ConvParams(ch, subbands * 2 * n_freq, 7, weight_norm=True, extra_init_draw=True)
        self.stft = _Window(n_fft)
        updown = torch.zeros(subbands, subbands, subbands)
        for k in range(subbands):
            updown[k, k, 0] = 1.0
        self.register_buffer("updown_filter", updown)
        self.multistream_conv_post = _ConvParams(4, 1, 63, weight_norm=True, bias=False, extra_init_draw=True)


# --------------------------------------------------------------------------------------------
# the public class
# --------------------------------------------------------------------------------------------
_SUPPORTED = dict(inter_channels=192, hidden_channels=192, resblock_kernel_sizes=[3, 7, 11],
                  resblock_dilation_sizes=[[1, 3, 5]] * 3, upsample_rates=[5, 4], upsample_initial_channel=512,
                  upsample_kernel_sizes=[16, 16], gen_istft_n_fft=16, gen_istft_hop_size=4, subbands=4,
                  gin_channels=256)


class SynthesizerTrn(nn.Module):
    """QuickVC generator; same constructor and `infer` as the reference class of this name.

    Extra, keyword-only knobs (not in the reference):
      precision  "tf32" (default; the north-star's "fp32 mode": fp32 storage and accumulation, TF32
                 tensor-core operands), "fp16" (IEEE-half operands -- TF32's 10-bit mantissa in two bytes, fp32
                 accumulation: the fp32-mode tolerance at the bf16 tensor rate, values saturate at 65504),
                 "bf16" (bf16 operands, fp32 accumulation), or "fp32" (exact fp32 CUDA-core FMA kernels, the
                 strict mode).
      backend    "tcgen05" or "fma"; default follows precision.
      chunk_utts decoder sub-batch size (0 = auto).
    """

    def __init__(self, spec_channels: int, segment_size: int, inter_channels: int, hidden_channels: int,
                 resblock_kernel_sizes: List[int], resblock_dilation_sizes: List[List[int]],
                 upsample_rates: List[int], upsample_initial_channel: int, upsample_kernel_sizes: List[int],
                 gen_istft_n_fft: int, gen_istft_hop_size: int, istft_vits: bool = False,
                 ms_istft_vits: bool = False, mb_istft_vits: bool = False, subbands=False,
                 gin_channels: int = 0, *, precision: str = "tf32", backend: Optional[str] = None,
                 chunk_utts: int = 0, **kwargs) -> None:
        super().__init__()
        _log.info("Loaded but not used: %s", kwargs)                       # models.py:573
        if kwargs.get("resblock"):
            assert kwargs["resblock"] == "1", "ResBlock2 support is droped."   # models.py:574-575
        self.segment_size = segment_size
        unit_channels = 256                                                 # models.py:579

        # decoder selection (models.py:588-589): only the configured MS-iSTFT decoder has kernels
        if not (mb_istft_vits or ms_istft_vits or istft_vits):
            raise RuntimeError(f"Not-supported decoder flag: {mb_istft_vits}/{ms_istft_vits}/{istft_vits}")
        if mb_istft_vits or not ms_istft_vits:
            raise NotImplementedError("only the Multistream_iSTFT_Generator decoder (ms_istft_vits=True, the "
                                      "shipped configs/quickvc.json) is implemented on B200")
        got = dict(inter_channels=inter_channels, hidden_channels=hidden_channels,
                   resblock_kernel_sizes=list(resblock_kernel_sizes),
                   resblock_dilation_sizes=[list(d) for d in resblock_dilation_sizes],
                   upsample_rates=list(upsample_rates), upsample_initial_channel=upsample_initial_channel,
                   upsample_kernel_sizes=list(upsample_kernel_sizes), gen_istft_n_fft=gen_istft_n_fft,
                   gen_istft_hop_size=gen_istft_hop_size, subbands=int(subbands), gin_channels=gin_channels)
        bad = {k: v for k, v in got.items() if v != _SUPPORTED[k]}
        if bad:
            raise NotImplementedError(f"the sm_100a kernels are built for configs/quickvc.json; unsupported: {bad}")

        # same construction order as the reference (models.py:582-591) so seeded inits coincide
        self.enc_q = _CondNormalWN(spec_channels, inter_channels, hidden_channels, 5, 16, gin_channels)
        self.enc_p = _CondNormalWN(unit_channels, inter_channels, hidden_channels, 5, 16, 0)
        self.flow = _Flow(inter_channels, hidden_channels, 5, 4, 4, gin_channels)
        self.enc_spk = _SpeakerEncoder(model_hidden_size=gin_channels, model_embedding_size=gin_channels)
        _log.info("Decoder type: Multistream_iSTFT_Generator")              # models.py:590
        self.dec = _MultistreamGenerator(inter_channels, resblock_kernel_sizes, resblock_dilation_sizes,
                                         upsample_rates, upsample_initial_channel, upsample_kernel_sizes,
                                         gen_istft_n_fft, gen_istft_hop_size, subbands, gin_channels)

        self._engine = InferEngine(self, precision=precision, backend=backend, chunk_utts=chunk_utts)

    # ---- weight lifecycle: any change of the parameters invalidates the folded copies ----
    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._engine.invalidate()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._engine.invalidate()
        return out

    def refold(self) -> None:
        """Call after mutating parameters in place (the folded copies cannot see that)."""
        self._engine.invalidate()

    def set_precision(self, precision: str, backend: Optional[str] = None) -> None:
        self._engine.configure(precision, backend)

    def forward(self, unit: Tensor, spec: Tensor, mel: Tensor):
        raise NotImplementedError("the training forward (models.py:593-623) is outside this library's scope; "
                                  "only `infer` is implemented")

    @torch.no_grad()
    def infer(self, unit: Tensor, mel: Tensor, *, noise: Optional[Tensor] = None,
              taps: Optional[Dict[str, Tensor]] = None, lengths: Optional[Tensor] = None) -> Tensor:
        """unit (B,256,T) fp32, mel (1,80,Tm) fp32 [or (B,80,Tm<=128)] -> waveform (B,1,320 T) fp32.

        `noise` (B,192,T) replaces the `torch.randn_like` draw of models.py:94 (default: drawn here
        with torch's generator, so seeded runs are reproducible).  `taps`, if a dict, is filled with
        the per-stage tensors of SURVEY.md section 8a in the reference layout.
        `lengths` (B,) integer frames per utterance, 1 <= lengths[b] <= T, declares a ragged batch padded to T:
        wave[b, 0, :320*lengths[b]] is then exactly what `infer(unit[b:b+1, :, :lengths[b]], ...)` returns for
        that utterance alone (the reference converts one utterance per call, convert.py:59-86), the rest zeros.
        """
        return self._engine.infer(unit, mel, noise=noise, taps=taps, lengths=lengths)

    @torch.no_grad()
    def embed_speaker(self, mel: Tensor) -> Tensor:
        """enc_spk.embed_utterance(mel.transpose(1,2)) (models.py:635) -> (1|Bm, 256); cache it and pass
        it to `infer_with_embedding` when converting many utterances to one target speaker."""
        return self._engine.embed(mel)

    @torch.no_grad()
    def infer_with_embedding(self, unit: Tensor, g: Tensor, *, noise: Optional[Tensor] = None,
                             lengths: Optional[Tensor] = None) -> Tensor:
        """`infer` with the speaker embedding(s) `g` (1|B, 256) of `embed_speaker` instead of a mel."""
        return self._engine.infer(unit, None, noise=noise, g=g, lengths=lengths)

    @torch.no_grad()
    def decode(self, z: Tensor, g: Tensor, *, taps: Optional[Dict[str, Tensor]] = None,
               lengths: Optional[Tensor] = None) -> Tensor:
        """dec(z, g=g)[0] (models.py:640): z (B,192,T), g (1|B,256,1) -> (B,1,320 T); `lengths` as for `infer`."""
        return self._engine.decode(z, g, taps=taps, lengths=lengths)
