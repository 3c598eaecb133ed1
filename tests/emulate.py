"""CPU emulation of the launch sequence of csrc/engine.cu on the folded weights (TEST ONLY).

It runs the same series convolutions, epilogues and buffer roles as the CUDA engine, in torch on the
CPU, so that the host-side logic (fold.py: weight-norm, Flip folding, zero-embedded couplings,
cond-bias matrix, polyphase ConvTranspose, polyphase synthesis filter; and the engine's sequencing)
can be checked against the oracle without a GPU.  Operand rounding (TF32 / bf16) is emulated so the
expected numerical error of each precision mode can be measured here as well.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from quickvc_official_b200 import capi, fold
from oracle import qvc_oracle

HID = 192


def rnd(x: torch.Tensor, opf: int) -> torch.Tensor:
    if opf == capi.OPF_F32:
        return x
    if opf == capi.OPF_TF32:
        return fold.round_tf32(x.float())
    if opf == capi.OPF_F16:
        return x.to(torch.float16).float()
    return x.to(torch.bfloat16).float()


def lrelu(x, slope):
    return torch.where(x > 0, x, x * slope)


def conv(x: torch.Tensor, L: Dict, out_rows: Optional[int] = None) -> torch.Tensor:
    """x [B][rows][cin] -> acc [B][out_rows][cout]; zero rows outside [0, rows) (qvc_b200.h)."""
    w = L["w"].float()                                   # [cout][k][cin]
    k, dil, pl = L["k"], L["dil"], L["pad_left"]
    rows = x.shape[1]
    out_rows = rows if out_rows is None else out_rows
    right = (out_rows - 1) + (k - 1) * dil - pl - (rows - 1)
    xt = F.pad(x.transpose(1, 2), (pl, max(right, 0)))
    y = F.conv1d(xt.double(), w.permute(0, 2, 1).double(), dilation=dil)[:, :, :out_rows]
    return y.transpose(1, 2).float()


class Emu:
    def __init__(self, sd, opf: int):
        self.opf = opf
        self.f = fold.fold_state_dict(sd, opf)
        self.L = self.f.layers
        self.t = self.f.tensors
        self.lens = None            # ragged batches: unit frames per utterance (qvc_infer `lengths`)

    def op(self, x: torch.Tensor, rate: int) -> torch.Tensor:
        """The operand copy an epilogue writes: rounded to the operand format and, in a ragged batch, ZERO for rows at
        or past lengths[b] * rate -- the only thing the kernels do about unequal lengths (qvc_conv_args.live_units)."""
        y = rnd(x, self.opf)
        if self.lens is not None:
            rows = torch.arange(y.shape[1]).reshape(1, -1, 1)
            y = torch.where(rows < (self.lens.reshape(-1, 1, 1) * rate), y, torch.zeros_like(y))
        return y

    def wn(self, xR, xO, l_in, l_rs, n_layers, gate_bias):
        skipR = None
        for i in range(n_layers):
            acc = conv(xO, self.L[l_in + i])
            b = gate_bias[i] if gate_bias is not None else self.L[l_in + i]["bias"]
            a = acc + b.reshape(-1, 1, b.shape[-1]) if b.dim() == 2 else acc + b
            acts = self.op(torch.tanh(a[..., :HID]) * torch.sigmoid(a[..., HID:]), 1)
            rs = conv(acts, self.L[l_rs + i]) + self.L[l_rs + i]["bias"]
            if i < n_layers - 1:
                xR = xR + rs[..., :HID]
                xO = self.op(xR, 1)
                skipR = rs[..., HID:2 * HID] if skipR is None else skipR + rs[..., HID:2 * HID]
            else:
                skipR = skipR + rs[..., :HID]
        return self.op(skipR, 1)

    def mrf(self, x1R, x1O, l0, final_slope, rate):
        total = None
        for r in range(3):
            srcR, srcO = x1R, x1O
            for j in range(3):
                c1, c2 = self.L[l0 + 6 * r + j], self.L[l0 + 6 * r + 3 + j]
                tO = self.op(lrelu(conv(srcO, c1) + c1["bias"], 0.1), rate)
                v = conv(tO, c2) + c2["bias"] + srcR
                if j < 2:
                    srcR, srcO = v, self.op(lrelu(v, 0.1), rate)
                else:
                    total = v / 3 if total is None else total + v / 3
        return total, self.op(lrelu(total, final_slope), rate)

    def decoder(self, zO, condvec, taps):
        L, opf = self.L, self.opf
        B, T, _ = zO.shape
        cp_bias = condvec[:, -512:].reshape(-1, 1, 512)
        pre = conv(zO, L[74]) + cp_bias
        taps["conv_pre"] = pre.transpose(1, 2)
        aO = self.op(lrelu(pre, 0.1), 1)
        u0 = (conv(aO, L[75]) + L[75]["bias"]).reshape(B, 5 * T, 256)
        taps["ups_0"] = u0.transpose(1, 2)
        m0, uO = self.mrf(u0, self.op(lrelu(u0, 0.1), 5), 77, 0.1, 5)
        taps["mrf_0"] = m0.transpose(1, 2)
        u1 = (conv(uO, L[76]) + L[76]["bias"]).reshape(B, 20 * T, 128)
        taps["ups_1"] = u1.transpose(1, 2)
        m1, pO = self.mrf(u1, self.op(lrelu(u1, 0.1), 20), 95, 0.01, 20)
        taps["mrf_1"] = m1.transpose(1, 2)
        pO = torch.cat([pO[:, 1:2], pO], 1)                     # reflect_row: row 0 <- x[1]
        cp = (conv(pO, L[113]) + L[113]["bias"])[..., :72]
        taps["conv_post"] = cp.transpose(1, 2)
        if self.lens is not None:
            # the tail sees, per utterance, only its own 20 len + 1 post-net frames (qvc_tail `live_units`); zeros after
            out = torch.zeros(B, 1, 320 * T)
            for b in range(B):
                n = int(self.lens[b])
                out[b, :, :320 * n] = self.tail(cp[b:b + 1, :20 * n + 1], {})
            return out
        return self.tail(cp, taps)

    def tail(self, cp, taps):
        B = cp.shape[0]
        x = cp.transpose(1, 2).reshape(B, 4, 18, -1)
        y = qvc_oracle.istft_closed_form(x[:, :, :9].reshape(B * 4, 9, -1), x[:, :, 9:].reshape(B * 4, 9, -1),
                                         self.t["tail.window"]).reshape(B, 4, -1)
        taps["y_mb"] = y
        E = self.t["tail.synth"]                               # [s][r][e]
        wt = torch.flip(E.permute(1, 0, 2), [2]).contiguous()  # [r][s][k], k = 16 - e
        o = F.conv1d(F.pad(y, (8, 8)), wt)                     # (B, 4 phases, Ny)
        return o.transpose(1, 2).reshape(B, 1, -1)

    def infer(self, unit, mel, noise, taps: Optional[Dict] = None, lengths=None):
        taps = {} if taps is None else taps
        self.lens = None if lengths is None else torch.as_tensor(lengths, dtype=torch.int64)
        L, opf, t = self.L, self.opf, self.t
        # speaker encoder: fp32 in every mode; the kernel path is checked on the GPU
        sd_spk = {f"enc_spk.lstm.weight_ih_l{l}": t[f"spk.w_ih.{l}"] for l in range(3)}
        sd_spk.update({f"enc_spk.lstm.weight_hh_l{l}": t[f"spk.w_hh.{l}"] for l in range(3)})
        sd_spk.update({f"enc_spk.lstm.bias_ih_l{l}": t[f"spk.bias.{l}"] for l in range(3)})
        sd_spk.update({f"enc_spk.lstm.bias_hh_l{l}": torch.zeros(1024) for l in range(3)})
        sd_spk["enc_spk.linear.weight"], sd_spk["enc_spk.linear.bias"] = t["spk.lin_w"], t["spk.lin_b"]
        g = qvc_oracle.embed_utterance(qvc_oracle._P(sd_spk, torch.float32), mel)          # (n_embed, 256)
        taps["g"] = g.unsqueeze(-1)
        condvec = g @ t["cond_w"].t() + t["cond_b"]                                       # (n_embed, 6656)

        unitO = self.op(unit.transpose(1, 2), 1)                 # to_series zeroes the padding frames of unit and noise
        noiseT = noise.transpose(1, 2)
        if self.lens is not None:
            noiseT = torch.where(torch.arange(noiseT.shape[1]).reshape(1, -1, 1) < self.lens.reshape(-1, 1, 1), noiseT,
                                 torch.zeros_like(noiseT))
        xR = conv(unitO, L[0]) + L[0]["bias"]
        skipO = self.wn(xR, self.op(xR, 1), 1, 17, 16, None)
        st = conv(skipO, L[33]) + L[33]["bias"]
        m, logs = st[..., :HID], st[..., HID:]
        zR = m + noiseT * torch.exp(logs)
        taps["m_p"], taps["logs_p"], taps["z_p"] = m.transpose(1, 2), logs.transpose(1, 2), zR.transpose(1, 2)
        zO = self.op(zR, 1)
        for c, idx in enumerate((6, 4, 2, 0)):
            lb = 34 + 10 * c
            hR = conv(zO, L[lb]) + L[lb]["bias"]
            gb = [condvec[:, c * 1536 + i * 384: c * 1536 + (i + 1) * 384] for i in range(4)]
            skipO = self.wn(hR, self.op(hR, 1), lb + 1, lb + 5, 4, gb)
            zR = zR - (conv(skipO, L[lb + 9]) + L[lb + 9]["bias"])
            zO = self.op(zR, 1)
            tap = zR.transpose(1, 2)
            taps[f"flow_{idx}"] = torch.flip(tap, [1]) if c % 2 == 0 else tap
        wave = self.decoder(zO, condvec, taps)
        taps["wave"] = wave
        return wave
