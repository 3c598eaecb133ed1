"""CPU oracle for QuickVC `SynthesizerTrn.infer` -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file is a from-scratch restatement, in plain torch functional ops on the CPU, of the
algorithm the reference runs in `SynthesizerTrn.infer` (/root/reference/models.py:625-642) and
its callees.  It consumes a reference-layout `state_dict` (467 keys) and
produces the waveform plus the per-stage taps of SURVEY.md section 8a.

Who may import it: `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` -- only ever as the checker or as the timed CPU baseline.  The product package
(`quickvc-official_b200/`) must never import, call or fall back to anything in this directory.

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the reference itself run in the authoring container:
`tests/golden/make_golden.py` imports /root/reference, loads the same synthetic state_dict,
injects the same noise and dumps taps to `tests/golden/*.npz`; `tests/test_oracle.py` checks this
file against those fixtures (and, when /root/reference is present, against the live reference).

Each function cites the reference lines it restates.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

LRELU_SLOPE = 0.1          # modules.py:11
N_FFT, HOP, SUBBANDS = 16, 4, 4   # configs/quickvc.json:38-40
PARTIAL_FRAMES, PARTIAL_HOP = 128, 64   # models.py:528


# --------------------------------------------------------------------------------------------
# parameter access
# --------------------------------------------------------------------------------------------
class _P:
    """state_dict view with dtype cast and old-style weight-norm resolution."""

    def __init__(self, sd: Mapping[str, Tensor], dtype: torch.dtype):
        self.sd, self.dtype = sd, dtype

    def raw(self, key: str) -> Tensor:
        return self.sd[key].detach().to("cpu", self.dtype)

    def has(self, key: str) -> bool:
        return key in self.sd

    def weight(self, prefix: str) -> Tensor:
        """`w = g * v / ||v||` with the norm over every dim but 0 (torch weight_norm, dim=0);
        plain `.weight` when the layer is not weight-normed."""
        if self.has(prefix + ".weight_v"):
            v, g = self.raw(prefix + ".weight_v"), self.raw(prefix + ".weight_g")
            nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
            return v * (g / nrm)
        return self.raw(prefix + ".weight")

    def bias(self, prefix: str) -> Optional[Tensor]:
        return self.raw(prefix + ".bias") if self.has(prefix + ".bias") else None


def _same_conv(x: Tensor, w: Tensor, b: Optional[Tensor], dilation: int = 1) -> Tensor:
    """Conv1d with odd kernel and zero 'same' padding = get_padding (commons.py:14-15)."""
    k = w.shape[-1]
    return F.conv1d(x, w, b, padding=(k * dilation - dilation) // 2, dilation=dilation)


# --------------------------------------------------------------------------------------------
# speaker encoder  (models.py:507-546)
# --------------------------------------------------------------------------------------------
def lstm_last_hidden(p: _P, x: Tensor) -> Tensor:
    """3-layer LSTM, batch_first, gate order i,f,g,o, zero initial state; returns hidden[-1].
    Restates nn.LSTM(80, 256, 3) as used at models.py:510,515-516.  x :: (N, steps, 80)."""
    n, steps, _ = x.shape
    seq = x
    h = None
    for layer in range(3):
        w_ih, w_hh = p.raw(f"enc_spk.lstm.weight_ih_l{layer}"), p.raw(f"enc_spk.lstm.weight_hh_l{layer}")
        b = p.raw(f"enc_spk.lstm.bias_ih_l{layer}") + p.raw(f"enc_spk.lstm.bias_hh_l{layer}")
        hid = w_hh.shape[1]
        gx = seq @ w_ih.t() + b                      # (N, steps, 4H): input projection for all steps
        h = torch.zeros(n, hid, dtype=x.dtype)
        c = torch.zeros(n, hid, dtype=x.dtype)
        outs = []
        for t in range(steps):
            gates = gx[:, t] + h @ w_hh.t()
            i, f, g, o = gates.split(hid, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        seq = torch.stack(outs, dim=1)
    return h


def speaker_forward(p: _P, mels: Tensor) -> Tensor:
    """SpeakerEncoder.forward (models.py:514-518): relu(linear(h_last)) / ||.||_2, no eps."""
    h = lstm_last_hidden(p, mels)
    e = torch.relu(h @ p.raw("enc_spk.linear.weight").t() + p.raw("enc_spk.linear.bias"))
    return e / e.norm(dim=1, keepdim=True)


def window_starts(total_frames: int) -> list:
    """compute_partial_slices (models.py:520-526) plus the always-appended last window (:530,535)."""
    starts = list(range(0, total_frames - PARTIAL_FRAMES, PARTIAL_HOP))
    starts.append(total_frames - PARTIAL_FRAMES)
    return starts


def embed_utterance(p: _P, mel: Tensor) -> Tensor:
    """embed_utterance (models.py:528-546).  mel :: (Bm, 80, Tm) -> (1|Bm, 256)."""
    frames = mel.transpose(1, 2)                       # (Bm, Tm, 80), models.py:635
    tm = frames.shape[1]
    if tm > PARTIAL_FRAMES:
        if frames.shape[0] != 1:
            # the reference stacks (W, Bm, 128, 80).squeeze(1): a 4-D LSTM input for Bm > 1
            raise ValueError("mel longer than 128 frames must have batch 1 (models.py:536)")
        wins = torch.stack([frames[0, s:s + PARTIAL_FRAMES] for s in window_starts(tm)], 0)
        return speaker_forward(p, wins).mean(dim=0, keepdim=True)      # not re-normalised (:540-541)
    return speaker_forward(p, frames)


# --------------------------------------------------------------------------------------------
# WaveNet stack  (modules.py:37-114)
# --------------------------------------------------------------------------------------------
def wn(p: _P, prefix: str, x: Tensor, n_layers: int, g: Optional[Tensor]) -> Tensor:
    hidden = x.shape[1]
    out = torch.zeros_like(x)
    cond = None
    if g is not None:                                    # modules.py:83-84, k=1 conv on (B,256,1)
        cond = F.conv1d(g, p.weight(prefix + ".cond_layer"), p.bias(prefix + ".cond_layer"))
    for i in range(n_layers):
        pre = _same_conv(x, p.weight(f"{prefix}.in_layers.{i}"), p.bias(f"{prefix}.in_layers.{i}"))
        if cond is not None:                             # modules.py:94-96
            pre = pre + cond[:, 2 * hidden * i: 2 * hidden * (i + 1)]
        acts = torch.tanh(pre[:, :hidden]) * torch.sigmoid(pre[:, hidden:])    # modules.py:14-34
        rs = F.conv1d(acts, p.weight(f"{prefix}.res_skip_layers.{i}"), p.bias(f"{prefix}.res_skip_layers.{i}"))
        if i < n_layers - 1:                             # modules.py:106-112
            x = x + rs[:, :hidden]
            out = out + rs[:, hidden:]
        else:
            out = out + rs
    return out


def prior_encoder(p: _P, unit: Tensor, noise: Tensor, taps: Optional[Dict[str, Tensor]]) -> Tensor:
    """CondNormalWN.forward for enc_p (models.py:75-95, instance :583): z = m + noise * exp(logs)."""
    h = F.conv1d(unit, p.weight("enc_p.pre"), p.bias("enc_p.pre"))
    h = wn(p, "enc_p.enc", h, 16, None)
    stats = F.conv1d(h, p.weight("enc_p.proj"), p.bias("enc_p.proj"))
    half = stats.shape[1] // 2
    m, logs = stats[:, :half], stats[:, half:]
    z = m + noise * torch.exp(logs)
    if taps is not None:
        taps["m_p"], taps["logs_p"], taps["z_p"] = m, logs, z
    return z


def flow_reverse(p: _P, z: Tensor, g: Tensor, taps: Optional[Dict[str, Tensor]]) -> Tensor:
    """ResidualCouplingBlock.forward(reverse=True) (models.py:39-51): Flip, RCL3, Flip, RCL2, ...
    with each coupling the mean-only reverse update x1 -= post(WN(pre(x0))) (modules.py:199-224)."""
    half = z.shape[1] // 2
    for idx in (6, 4, 2, 0):
        z = torch.flip(z, [1])                                          # modules.py:165-170
        pre = f"flow.flows.{idx}"
        x0, x1 = z[:, :half], z[:, half:]
        h = F.conv1d(x0, p.weight(pre + ".pre"), p.bias(pre + ".pre"))
        h = wn(p, pre + ".enc", h, 4, g)
        m = F.conv1d(h, p.weight(pre + ".post"), p.bias(pre + ".post"))
        z = torch.cat([x0, x1 - m], 1)
        if taps is not None:
            taps[f"flow_{idx}"] = z
    return z


# --------------------------------------------------------------------------------------------
# decoder  (models.py:304-408)
# --------------------------------------------------------------------------------------------
def resblock1(p: _P, prefix: str, x: Tensor) -> Tensor:
    """ResBlock1.forward (modules.py:147-154): three [lrelu, dilated conv, lrelu, conv, +x]."""
    for j, d in enumerate((1, 3, 5)):
        t = F.leaky_relu(x, LRELU_SLOPE)
        t = _same_conv(t, p.weight(f"{prefix}.convs1.{j}"), p.bias(f"{prefix}.convs1.{j}"), d)
        t = F.leaky_relu(t, LRELU_SLOPE)
        t = _same_conv(t, p.weight(f"{prefix}.convs2.{j}"), p.bias(f"{prefix}.convs2.{j}"), 1)
        x = t + x
    return x


def istft_closed_form(spec_log: Tensor, phase_raw: Tensor, window: Tensor) -> Tensor:
    """models.py:399-401 -> torchaudio InverseSpectrogram(16,16,4) -> torch.istft(center=True,
    onesided, periodic-Hann window buffer, length=None), restated without FFT calls:
    real inverse DFT of each frame (Im of DC/Nyquist ignored), window, hop-4 overlap-add,
    division by the overlap-added squared window, 8 samples trimmed at both ends.
    spec_log / phase_raw :: (N, 9, L) -> (N, 4 (L - 1))."""
    n, bins, frames = spec_log.shape
    mag = torch.exp(spec_log)
    ph = math.pi * torch.sin(phase_raw)
    re, im = mag * torch.cos(ph), mag * torch.sin(ph)
    k = torch.arange(bins, dtype=spec_log.dtype).unsqueeze(1)
    t = torch.arange(N_FFT, dtype=spec_log.dtype).unsqueeze(0)
    ang = 2.0 * math.pi * k * t / N_FFT                                   # (9, 16)
    wk = torch.full((bins, 1), 2.0, dtype=spec_log.dtype)
    wk[0, 0] = 1.0
    wk[-1, 0] = 1.0
    c = (wk * torch.cos(ang)) / N_FFT                                     # Re weights
    s = (wk * torch.sin(ang)) / N_FFT                                     # Im weights (row 0, 8 are 0)
    x = torch.einsum("nkl,kt->ntl", re, c) - torch.einsum("nkl,kt->ntl", im, s)    # (N, 16, L)
    x = x * window.reshape(1, N_FFT, 1)
    full = N_FFT + HOP * (frames - 1)
    y = F.fold(x, (1, full), (1, N_FFT), stride=(1, HOP)).reshape(n, full)
    env = F.fold((window * window).reshape(1, N_FFT, 1).expand(1, N_FFT, frames).contiguous(),
                 (1, full), (1, N_FFT), stride=(1, HOP)).reshape(1, full)
    return (y / env)[:, N_FFT // 2: full - N_FFT // 2]


def decoder(p: _P, z: Tensor, g: Tensor, taps: Optional[Dict[str, Tensor]]) -> Tensor:
    """Multistream_iSTFT_Generator.forward (models.py:360-408)."""
    x = _same_conv(z, p.weight("dec.conv_pre"), p.bias("dec.conv_pre")) \
        + F.conv1d(g, p.weight("dec.cond"), p.bias("dec.cond"))                     # :372
    if taps is not None:
        taps["conv_pre"] = x
    ups = ((5, 16, 6, 1), (4, 16, 6, 0))      # stride, kernel, padding, output_padding (models.py:335)
    for i, (u, k, pad, opad) in enumerate(ups):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, p.weight(f"dec.ups.{i}"), p.bias(f"dec.ups.{i}"),
                               stride=u, padding=pad, output_padding=opad)
        if taps is not None:
            taps[f"ups_{i}"] = x
        acc = None
        for j in range(3):                                                         # :378-384
            r = resblock1(p, f"dec.resblocks.{3 * i + j}", x)
            acc = r if acc is None else acc + r
        x = acc / 3
        if taps is not None:
            taps[f"mrf_{i}"] = x
    x = F.leaky_relu(x)                       # default slope 0.01 (models.py:385)
    x = F.pad(x, (1, 0), mode="reflect")      # ReflectionPad1d((1, 0)) (:345,388)
    x = _same_conv(x, p.weight("dec.subband_conv_post"), p.bias("dec.subband_conv_post"))
    if taps is not None:
        taps["conv_post"] = x
    b, _, frames = x.shape
    x = x.reshape(b, SUBBANDS, 2 * (N_FFT // 2 + 1), frames)                      # :390
    spec = x[:, :, :N_FFT // 2 + 1].reshape(b * SUBBANDS, N_FFT // 2 + 1, frames)
    phase = x[:, :, N_FFT // 2 + 1:].reshape(b * SUBBANDS, N_FFT // 2 + 1, frames)
    y_mb = istft_closed_form(spec, phase, p.raw("dec.stft.window")).reshape(b, SUBBANDS, -1)
    if taps is not None:
        taps["y_mb"] = y_mb
    up = F.conv_transpose1d(y_mb, p.raw("dec.updown_filter") * SUBBANDS, stride=SUBBANDS)      # :405
    w_syn = p.weight("dec.multistream_conv_post")
    return F.conv1d(up, w_syn, None, padding=(w_syn.shape[-1] - 1) // 2)          # :406


# --------------------------------------------------------------------------------------------
# top level  (models.py:625-642)
# --------------------------------------------------------------------------------------------
@torch.no_grad()
def infer(sd: Mapping[str, Tensor], unit: Tensor, mel: Tensor, noise: Tensor,
          dtype: torch.dtype = torch.float32,
          taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """unit (B,256,T), mel (Bm,80,Tm), noise (B,192,T) standing in for `torch.randn_like(mu)`
    at models.py:94  ->  waveform (B,1,320 T).  `taps`, if given, receives the stage tensors."""
    p = _P(sd, dtype)
    unit, mel, noise = (t.detach().to("cpu", dtype) for t in (unit, mel, noise))
    g = embed_utterance(p, mel).unsqueeze(-1)                                    # :635
    if taps is not None:
        taps["g"] = g
    z_p = prior_encoder(p, unit, noise, taps)                                     # :638
    z = flow_reverse(p, z_p, g, taps)                                             # :639
    o = decoder(p, z, g, taps)                                                    # :640
    if taps is not None:
        taps["wave"] = o
    return o


@torch.no_grad()
def decode_only(sd: Mapping[str, Tensor], z: Tensor, g: Tensor, dtype: torch.dtype = torch.float32,
                taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """BASELINE.json config 3: the decoder alone on a latent z (B,192,T) and embedding g (1|B,256,1)."""
    p = _P(sd, dtype)
    return decoder(p, z.detach().to("cpu", dtype), g.detach().to("cpu", dtype), taps)


TAP_NAMES = ("g", "m_p", "logs_p", "z_p", "flow_6", "flow_4", "flow_2", "flow_0", "conv_pre",
             "ups_0", "mrf_0", "ups_1", "mrf_1", "conv_post", "y_mb", "wave")
