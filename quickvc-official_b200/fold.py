"""Weight folding: reference-layout state_dict -> kernel-ready tensors (run once per load).

Everything that the reference recomputes on every forward but that depends only on the weights is
done here, once:
  * old-style weight-norm `w = g * v / ||v||` (every `weight_norm(...)` site);
  * `Flip` (modules.py:165-170) folded into channel permutations of each coupling's pre / post;
  * the 96-channel coupling halves (modules.py:209-222) zero-embedded into 192-channel filters so
    the flow state never has to be split or concatenated;
  * `cond_layer(g)` / `dec.cond(g)` (modules.py:83-96, models.py:372) turned into one matrix whose
    product with the speaker embedding yields all per-utterance bias vectors;
  * ConvTranspose1d (models.py:333-335) re-expressed as a stride-1 series convolution with
    phase-major output columns ([T][s*Cout] is bit-identical memory to [s*T][Cout]);
  * `updown_filter` zero-stuffing + the 63-tap synthesis Conv1d (models.py:353-357,405-406) folded
    into a 4-phase, 17-tap-per-band polyphase filter;
  * operands rounded to the GEMM operand format (TF32 round-to-nearest, or bf16).
All filters end up as [cout][taps][cin], the K-major layout both convolution back ends consume.
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Tuple

import torch

from . import capi

Tensor = torch.Tensor

HID = 192
N_WN_ENC, N_WN_FLOW, N_COUPLINGS = 16, 4, 4
COND_ROWS = N_COUPLINGS * N_WN_FLOW * 2 * HID + 512


def _weight(sd: Mapping[str, Tensor], prefix: str) -> Tensor:
    """Resolve old-style weight norm (norm over all dims but 0) or return the plain weight."""
    if prefix + ".weight_v" in sd:
        v = sd[prefix + ".weight_v"].detach().double()
        g = sd[prefix + ".weight_g"].detach().double()
        nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
        return v * (g / nrm)
    return sd[prefix + ".weight"].detach().double()


def _bias(sd: Mapping[str, Tensor], prefix: str, n: int) -> Tensor:
    key = prefix + ".bias"
    return sd[key].detach().double() if key in sd else torch.zeros(n, dtype=torch.float64, device=_dev(sd))


def _dev(sd: Mapping[str, Tensor]):
    return next(iter(sd.values())).device


def round_tf32(w: Tensor) -> Tensor:
    """fp32 -> nearest TF32 (10-bit mantissa), ties away from zero: what `cvt.rna.tf32.f32` does."""
    bits = w.contiguous().view(torch.int32)
    bits = (bits + 0x1000) & ~0x1FFF
    return bits.view(torch.float32)


def to_operand(w: Tensor, opformat: int) -> Tensor:
    w = w.to(torch.float32).contiguous()
    if opformat == capi.OPF_F32:
        return w
    if opformat == capi.OPF_TF32:
        return round_tf32(w)
    if opformat == capi.OPF_BF16:
        return w.to(torch.bfloat16).contiguous()
    if opformat == capi.OPF_F16:
        return w.to(torch.float16).contiguous()
    raise ValueError(f"bad operand format {opformat}")


def conv_filter(w: Tensor) -> Tensor:
    """Conv1d weight (Cout, Cin, k) -> [Cout][k][Cin]."""
    return w.permute(0, 2, 1).contiguous()


def polyphase_transpose_filter(w: Tensor, stride: int, padding: int) -> Tuple[Tensor, int]:
    """ConvTranspose1d weight (Cin, Cout, k) -> ([stride*Cout][taps][Cin], pad_left).

    y[co, s*q + r] = sum_{ci, j : (r + p - j) % s == 0} w[ci, co, j] * x[ci, q + (r + p - j)/s],
    i.e. output phase r is an ordinary convolution over input offsets o = (r + p - j)/s.  All
    phases share one tap window [o_min, o_max]; absent (phase, offset) pairs hold zeros.
    """
    cin, cout, k = w.shape
    offs = [(r + padding - j) // stride for r in range(stride) for j in range(k) if (r + padding - j) % stride == 0]
    o_min, o_max = min(offs), max(offs)
    taps = o_max - o_min + 1
    f = torch.zeros(stride * cout, taps, cin, dtype=w.dtype, device=w.device)
    for r in range(stride):
        for j in range(k):
            if (r + padding - j) % stride:
                continue
            o = (r + padding - j) // stride
            f[r * cout:(r + 1) * cout, o - o_min, :] = w[:, :, j].t()
    return f, -o_min


def synthesis_polyphase(updown: Tensor, w_syn: Tensor) -> Tensor:
    """E[s][r][e] with wave[4q + r] = sum_s sum_e E[s][r][e] * y[s][q + 8 - e]  (tail.cu).

    Derivation: up[s'][4m + tau] = 4 * sum_s F[s][s'][tau] * y[s][m]  (conv_transpose1d, models.py:405),
    wave[n] = sum_{s', j} w_syn[s'][j] * up[s'][n + j - 31]  (Conv1d k=63 pad 31, models.py:406); with
    n = 4q + r and m = q + 8 - e the tap index is j = 63 - 4e + tau - r.
    """
    nb = updown.shape[0]
    k = w_syn.shape[-1]
    assert nb == 4 and tuple(updown.shape) == (4, 4, 4) and k == 63, "tail kernel is built for 4 bands, 63 taps"
    e_f = torch.zeros(nb, 4, 17, dtype=torch.float64, device=w_syn.device)
    for e in range(17):
        for r in range(4):
            for tau in range(4):
                j = 63 - 4 * e + tau - r
                if 0 <= j < k:
                    # sum over s' of F[s][s'][tau] * w_syn[s'][j]
                    e_f[:, r, e] += 4.0 * (updown[:, :, tau].double() @ w_syn[0, :, j].double())
    return e_f


def frame_pair_filter(w: Tensor, pad_left: int) -> Tuple[Tensor, int]:
    """(C, k, cin) dilation-1 filter -> the (2C, k', 2cin) filter of the same convolution on frame PAIRS, and its left
    padding in pair rows (include/qvc_b200.h, qvc_model.paired):
        out[2n + p][c] = sum_j sum_ci w[c][j][ci] x[2n + p + j - pad][ci],   2n + p + j - pad = 2 (n + a) + q
        =>  w'[p C + c][a - a_min][q cin + ci] = w[c][2a + q - p + pad][ci]."""
    C, k, cin = w.shape
    a_min = (0 - pad_left) // 2
    a_max = (1 + (k - 1) - pad_left) // 2
    wp = torch.zeros(2 * C, a_max - a_min + 1, 2 * cin, dtype=w.dtype, device=w.device)
    for p in (0, 1):
        for j in range(k):
            a, q = divmod(p + j - pad_left, 2)
            wp[p * C:(p + 1) * C, a - a_min, q * cin:(q + 1) * cin] = w[:, j, :]
    return wp, -a_min


class Folded:
    """Device tensors plus the metadata `capi.Model` needs.  Keeps every tensor alive."""

    def __init__(self) -> None:
        self.layers: List[Dict] = []
        self.paired: Dict[int, Dict] = {}        # layer index -> frame-paired form (frame_pair_filter)
        self.wn_skip: List[Dict] = []            # per WN stack: deferred skip sum (qvc_model.wn_skip)
        self.tensors: Dict[str, Tensor] = {}

    def to(self, device) -> "Folded":
        """Move every device-side tensor (the `*_host` copies stay on the host); the layer tables follow."""
        moved: Dict[int, Tensor] = {}

        def mv(t):
            if t is None:
                return None
            if id(t) not in moved:
                moved[id(t)] = t.to(device)
            return moved[id(t)]

        for name in list(self.tensors):
            if not name.endswith("_host"):
                self.tensors[name] = mv(self.tensors[name])
        for L in list(self.layers) + list(self.paired.values()) + list(self.wn_skip):
            L["w"], L["bias"] = mv(L["w"]), mv(L["bias"])
        return self

    def add_layer(self, name: str, w: Tensor, bias, dil: int, pad_left: int, opformat: int) -> None:
        cout, k, cin = w.shape
        pad_rows = (-cout) % 16
        if pad_rows:
            w = torch.cat([w, torch.zeros(pad_rows, k, cin, dtype=w.dtype, device=w.device)], 0)
            if bias is not None:
                bias = torch.cat([bias, torch.zeros(pad_rows, dtype=bias.dtype, device=bias.device)], 0)
        wt = to_operand(w, opformat)
        bt = bias.to(torch.float32).contiguous() if bias is not None else None
        self.tensors[name + ".w"] = wt
        if bt is not None:
            self.tensors[name + ".b"] = bt
        self.layers.append(dict(name=name, w=wt, bias=bt, cin=cin, cout=cout + pad_rows, k=k, dil=dil, pad_left=pad_left))
        # dilation-1 layers with 128 channels in and out (MRF-2): also keep the frame-paired form
        if dil == 1 and cin == 128 and cout == 128 and k % 2 == 1 and k > 1 and bias is not None:
            wp, pad_p = frame_pair_filter(w, pad_left)
            wpt = to_operand(wp, opformat)
            bpt = torch.cat([bt, bt]).contiguous()
            self.tensors[name + ".w2"] = wpt
            self.tensors[name + ".b2"] = bpt
            self.paired[len(self.layers) - 1] = dict(w=wpt, bias=bpt, cin=2 * cin, cout=2 * cout, k=wp.shape[1], dil=1,
                                                     pad_left=pad_p)


def fold_state_dict(sd: Mapping[str, Tensor], opformat: int) -> Folded:
    """Build the canonical 114-layer table (order documented in include/qvc_b200.h)."""
    f = Folded()
    dev = _dev(sd)

    def plain(name: str, prefix: str, dil: int = 1, with_bias: bool = True) -> None:
        w = _weight(sd, prefix)
        k = w.shape[-1]
        f.add_layer(name, conv_filter(w), _bias(sd, prefix, w.shape[0]) if with_bias else None, dil,
                    (k - 1) * dil // 2, opformat)

    # ---- prior encoder (models.py:71-73, modules.py:64-67) ----
    plain("enc_p.pre", "enc_p.pre")
    for i in range(N_WN_ENC):
        plain(f"enc_p.in.{i}", f"enc_p.enc.in_layers.{i}")
    for i in range(N_WN_ENC):
        plain(f"enc_p.rs.{i}", f"enc_p.enc.res_skip_layers.{i}")
    plain("enc_p.proj", "enc_p.proj")

    def skip_sum(prefix: str, n_layers: int) -> None:
        """qvc_model.wn_skip: the skip halves of a stack's res_skip layers side by side along the input channels."""
        ws, b = [], torch.zeros(HID, dtype=torch.float64, device=dev)
        for i in range(n_layers):
            w = _weight(sd, f"{prefix}{i}")[:, :, 0]                  # (2H | H, H)
            bi = _bias(sd, f"{prefix}{i}", w.shape[0])
            ws.append(w[-HID:])
            b = b + bi[-HID:].float().double()
        w_all = torch.cat(ws, dim=1).reshape(HID, 1, n_layers * HID)
        f.wn_skip.append(dict(w=to_operand(w_all, opformat).contiguous(), bias=b.float().contiguous(), cin=n_layers * HID,
                              cout=HID, k=1, dil=1, pad_left=0))

    skip_sum("enc_p.enc.res_skip_layers.", N_WN_ENC)

    # ---- flow, execution order of reverse=True: flows 6, 4, 2, 0 (models.py:48) ----
    # The reference applies Flip before each coupling; two flips cancel, so couplings 6 and 2 see the
    # channel-reversed state and couplings 4 and 0 the original one.  We keep the state in the original
    # orientation and reverse the channel indexing of pre (inputs) and post (outputs) instead.
    cond_w = torch.zeros(COND_ROWS, 256, dtype=torch.float64, device=dev)
    cond_b = torch.zeros(COND_ROWS, dtype=torch.float64, device=dev)
    for c, idx in enumerate((6, 4, 2, 0)):
        flipped = c % 2 == 0
        p = f"flow.flows.{idx}"
        w_pre = _weight(sd, p + ".pre")[:, :, 0]            # (192, 96)
        w_post = _weight(sd, p + ".post")[:, :, 0]          # (96, 192)
        b_post = _bias(sd, p + ".post", 96)
        half = w_pre.shape[1]
        wp = torch.zeros(HID, 1, HID, dtype=torch.float64, device=dev)
        wq = torch.zeros(HID, 1, HID, dtype=torch.float64, device=dev)
        bq = torch.zeros(HID, dtype=torch.float64, device=dev)
        if flipped:
            # x0 of the flipped state = reversed upper half; the update lands on the reversed lower half
            wp[:, 0, HID - half:] = torch.flip(w_pre, [1])
            wq[:half, 0, :] = torch.flip(w_post, [0])
            bq[:half] = torch.flip(b_post, [0])
        else:
            wp[:, 0, :half] = w_pre
            wq[half:, 0, :] = w_post
            bq[half:] = b_post
        f.add_layer(f"flow.{c}.pre", wp, _bias(sd, p + ".pre", HID), 1, 0, opformat)
        for i in range(N_WN_FLOW):
            plain(f"flow.{c}.in.{i}", f"{p}.enc.in_layers.{i}", with_bias=False)
        for i in range(N_WN_FLOW):
            plain(f"flow.{c}.rs.{i}", f"{p}.enc.res_skip_layers.{i}")
        f.add_layer(f"flow.{c}.post", wq, bq, 1, 0, opformat)
        skip_sum(f"{p}.enc.res_skip_layers.", N_WN_FLOW)
        rows = slice(c * N_WN_FLOW * 2 * HID, (c + 1) * N_WN_FLOW * 2 * HID)
        cond_w[rows] = _weight(sd, p + ".enc.cond_layer")[:, :, 0]
        cond_b[rows] = _bias(sd, p + ".enc.cond_layer", N_WN_FLOW * 2 * HID) + torch.cat(
            [_bias(sd, f"{p}.enc.in_layers.{i}", 2 * HID) for i in range(N_WN_FLOW)])

    # ---- decoder (models.py:327-346) ----
    plain("dec.conv_pre", "dec.conv_pre", with_bias=False)
    cond_w[-512:] = _weight(sd, "dec.cond")[:, :, 0]
    cond_b[-512:] = _bias(sd, "dec.cond", 512) + _bias(sd, "dec.conv_pre", 512)
    for i, (stride, k, pad) in enumerate(((5, 16, 6), (4, 16, 6))):      # models.py:335
        w = _weight(sd, f"dec.ups.{i}")                                  # (Cin, Cout, k), norm per Cin
        assert w.shape[-1] == k
        wf, pad_left = polyphase_transpose_filter(w, stride, pad)
        f.add_layer(f"dec.ups.{i}", wf, _bias(sd, f"dec.ups.{i}", w.shape[1]).repeat(stride), 1, pad_left, opformat)
    for r in range(6):
        for j, d in enumerate((1, 3, 5)):
            plain(f"dec.res.{r}.c1.{j}", f"dec.resblocks.{r}.convs1.{j}", dil=d)
        for j in range(3):
            plain(f"dec.res.{r}.c2.{j}", f"dec.resblocks.{r}.convs2.{j}")
    plain("dec.post", "dec.subband_conv_post")
    assert len(f.layers) == capi.QVC_NUM_LAYERS, len(f.layers)

    f.tensors["cond_w"] = cond_w.to(torch.float32).contiguous()
    f.tensors["cond_b"] = cond_b.to(torch.float32).contiguous()

    # ---- speaker encoder (models.py:510-511) ----
    for l in range(3):
        f.tensors[f"spk.w_ih.{l}"] = sd[f"enc_spk.lstm.weight_ih_l{l}"].detach().float().contiguous()
        f.tensors[f"spk.w_hh.{l}"] = sd[f"enc_spk.lstm.weight_hh_l{l}"].detach().float().contiguous()
        f.tensors[f"spk.bias.{l}"] = (sd[f"enc_spk.lstm.bias_ih_l{l}"].detach().double()
                                      + sd[f"enc_spk.lstm.bias_hh_l{l}"].detach().double()).float().contiguous()
    f.tensors["spk.lin_w"] = sd["enc_spk.linear.weight"].detach().float().contiguous()
    f.tensors["spk.lin_b"] = sd["enc_spk.linear.bias"].detach().float().contiguous()

    # ---- tail (models.py:350-357) ----
    f.tensors["tail.window"] = sd["dec.stft.window"].detach().float().contiguous()
    f.tensors["tail.synth"] = synthesis_polyphase(sd["dec.updown_filter"].detach(),
                                                  _weight(sd, "dec.multistream_conv_post")).float().contiguous()
    # host copies: the tail kernel takes its 16 + 272 coefficients as kernel parameters (constant bank)
    f.tensors["tail.window_host"] = f.tensors["tail.window"].cpu().contiguous()
    f.tensors["tail.synth_host"] = f.tensors["tail.synth"].cpu().contiguous()
    return f


def build_model_struct(f: Folded, opformat: int, backend: int, chunk_utts: int) -> capi.Model:
    m = capi.Model()
    m.abi_version = capi.QVC_ABI_VERSION
    m.opformat, m.backend, m.chunk_utts = opformat, backend, chunk_utts
    for i, L in enumerate(f.layers):
        m.layers[i].w = L["w"].data_ptr()
        m.layers[i].bias = L["bias"].data_ptr() if L["bias"] is not None else None
        m.layers[i].cin, m.layers[i].cout = L["cin"], L["cout"]
        m.layers[i].k, m.layers[i].dil, m.layers[i].pad_left = L["k"], L["dil"], L["pad_left"]
    for i, L in f.paired.items():
        m.paired[i].w, m.paired[i].bias = L["w"].data_ptr(), L["bias"].data_ptr()
        m.paired[i].cin, m.paired[i].cout = L["cin"], L["cout"]
        m.paired[i].k, m.paired[i].dil, m.paired[i].pad_left = L["k"], L["dil"], L["pad_left"]
    for i, L in enumerate(f.wn_skip):
        m.wn_skip[i].w, m.wn_skip[i].bias = L["w"].data_ptr(), L["bias"].data_ptr()
        m.wn_skip[i].cin, m.wn_skip[i].cout = L["cin"], L["cout"]
        m.wn_skip[i].k, m.wn_skip[i].dil, m.wn_skip[i].pad_left = L["k"], L["dil"], L["pad_left"]
    t = f.tensors
    m.cond_w, m.cond_b, m.cond_rows = t["cond_w"].data_ptr(), t["cond_b"].data_ptr(), COND_ROWS
    for l in range(3):
        m.spk.w_ih[l] = t[f"spk.w_ih.{l}"].data_ptr()
        m.spk.w_hh[l] = t[f"spk.w_hh.{l}"].data_ptr()
        m.spk.bias[l] = t[f"spk.bias.{l}"].data_ptr()
    m.spk.lin_w, m.spk.lin_b = t["spk.lin_w"].data_ptr(), t["spk.lin_b"].data_ptr()
    m.tail.window, m.tail.synth = t["tail.window"].data_ptr(), t["tail.synth"].data_ptr()
    m.tail.window_host, m.tail.synth_host = t["tail.window_host"].data_ptr(), t["tail.synth_host"].data_ptr()
    return m
