#!/usr/bin/env python
"""Per-launch roofline table of one B = 64 x 10 s step, from the committed ncu passes (no GPU needed).

    python scripts/layer_roofline.py profiles/r02_traffic_fp16.csv fp16 > profiles/r02_layer_roofline_fp16.md

Input: an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none` pass over
one step (scripts/profile_step.py; caches cold and launches serialised, so times are upper bounds of the in-step ones).
For every launch: the layer it serves (the engine's launch order, csrc/engine.cu), its algorithmic FLOPs (SURVEY.md
section 8a: multiply-accumulates per unit frame x 2 x B x T), the measured duration and DRAM bytes, and both roofline
fractions -- tensor: FLOP/s over MEASURED_PEAKS.json bf16_tflops_sustained (x 0.5 for TF32 operands); HBM: DRAM bytes/s over
hbm_gbs.  `bound` names the larger one; a launch far from both is latency- or issue-bound and says so."""
import collections
import csv
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, T = 64, 500


def layer_sequence():
    """(name, kernel family, MAC per unit frame) in launch order for the tensor-core part of a step."""
    seq = [("enc_p.pre 256->192 k1", "tcr", 49152)]
    for i in range(16):
        seq.append((f"enc_p.wn.{i} in_layer + gate 192->384 k5", "tcr", 368640))
        if i < 15:
            seq.append((f"enc_p.wn.{i} res_skip (residual half) 192->192 k1", "tcr", 36864))
    seq.append(("enc_p.wn deferred skip sum [16x192]->192 k1", "tcr", 16 * 36864))
    seq.append(("enc_p.proj + sample 192->384 k1", "tc", 73728))
    for c in (6, 4, 2, 0):
        seq.append((f"flow.{c}.pre 96->192 k1", "tcr", 18432))
        for i in range(4):
            seq.append((f"flow.{c}.wn.{i} in_layer + gate 192->384 k5", "tcr", 368640))
            if i < 3:
                seq.append((f"flow.{c}.wn.{i} res_skip (residual half) 192->192 k1", "tcr", 36864))
        seq.append((f"flow.{c}.wn deferred skip sum [4x192]->192 k1", "tcr", 4 * 36864))
        seq.append((f"flow.{c}.post 192->96 k1, x1 - m", "tcr", 18432))
    seq.append(("dec.conv_pre 192->512 k7 (+ cond)", "tc2", 688128))
    seq.append(("dec.ups.0 ConvT 512->256 k16 s5 (polyphase)", "tc2", 2097152))
    for stage, fam, name in ((0, "tc2", "256 ch @ 5T"), (1, "tcr", "128 ch @ 20T")):
        per_k = 327680
        for k in (3, 7, 11):
            for j in range(3):
                seq.append((f"MRF-{stage + 1} k{k} c1.{j} ({name}, dil {(1, 3, 5)[j]})", fam, per_k * k))
                if j < 2:
                    seq.append((f"MRF-{stage + 1} k{k} c2.{j} + residual ({name})", fam, per_k * k))
        seq.append((f"MRF-{stage + 1} sum of the three last c2 (k3 + k7 + k11) -> mean ({name})", "tc2", per_k * 21))
        if stage == 0:
            seq.append(("dec.ups.1 ConvT 256->128 k16 s4 (polyphase)", "tc2", 2621440))
    seq.append(("subband_conv_post 128->72 k7 + iSTFT / OLA / synthesis tail", "post_tail", 1290240))
    return seq


def read(path):
    txt = open(path).read().splitlines()
    i = next(n for n, l in enumerate(txt) if l.startswith('"ID"'))
    per = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("\n".join(txt[i:]))):
        d = per.setdefault(int(r["ID"]), {"kernel": r["Kernel Name"], "grid": r["Grid Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"])
    return list(per.values())


def main():
    path, mode = sys.argv[1], sys.argv[2]
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        pk = json.load(f)
    tensor_peak = pk["bf16_tflops_sustained"] * (0.5 if mode == "tf32" else 1.0)
    hbm_peak = pk["hbm_gbs"]
    launches = read(path)
    fam_of = lambda k: ("post_tail" if "post_tail_kernel" in k else "tcr" if "conv_tcr_kernel" in k else
                        "tc2" if "conv_tc2_kernel" in k else "tc" if "conv_tc_kernel" in k else None)
    tc = [l for l in launches if fam_of(l["kernel"])]
    seq = layer_sequence()
    assert len(tc) == len(seq), (len(tc), len(seq))
    total_mac = sum(m for _, _, m in seq)
    assert abs(total_mac - 103.6e6) / 103.6e6 < 2e-3, total_mac            # 103.6 M MAC per unit frame
    print(f"# Per-launch roofline of one step, B = {B} x {T / 50:.0f} s, {mode} mode\n")
    print(f"Source: `{os.path.relpath(path, ROOT)}` (ncu, caches cold, launches serialised); peaks: MEASURED_PEAKS.json "
          f"bf16_tflops_sustained {pk['bf16_tflops_sustained']}" + (" x 0.5 (TF32 operands)" if mode == "tf32" else "") +
          f" = {tensor_peak:.1f} TFLOP/s, hbm_gbs {hbm_peak:.0f} GB/s.  FLOPs are algorithmic (SURVEY.md section 8a); DRAM bytes are "
          "measured (read + write).  Made by `scripts/layer_roofline.py`.\n")
    print("| # | layer | kernel | GFLOP | us | TFLOP/s | of tensor peak | DRAM MB | GB/s | of HBM peak | bound |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    groups = collections.OrderedDict()
    tot_t = tot_f = tot_b = 0.0
    for n, (l, (name, fam, mac)) in enumerate(zip(tc, seq)):
        assert fam_of(l["kernel"]) == fam, (n, name, l["kernel"])
        flop = 2.0 * mac * B * T
        us = l["gpu__time_duration.sum"] / 1e3
        byt = l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"]
        tf, gb = flop / us / 1e6, byt / us / 1e3
        ft, fh = tf / tensor_peak, gb / hbm_peak
        bound = "tensor" if ft >= fh and ft >= 0.5 else "hbm" if fh > ft and fh >= 0.5 else \
            ("tensor (partly)" if ft >= fh else "hbm (partly)") if max(ft, fh) >= 0.3 else "latency / issue"
        print(f"| {n} | {name} | {fam} | {flop / 1e9:.1f} | {us:.1f} | {tf:.0f} | {ft:.2f} | {byt / 1e6:.0f} | {gb:.0f} | {fh:.2f} | {bound} |")
        key = name.split(" ")[0].split(".")[0] if not name.startswith("MRF") else name[:5]
        key = {"enc_p": "enc_p (prior encoder)", "flow": "flow (4 couplings)", "dec": "decoder pre / upsamplers",
               "subband_conv_post": "post-net + tail"}.get(key, key)
        g = groups.setdefault(key, [0.0, 0.0, 0.0, 0])
        g[0] += us; g[1] += flop; g[2] += byt; g[3] += 1
        tot_t += us; tot_f += flop; tot_b += byt
    print("\n| block | launches | us | share of conv time | TFLOP/s | of tensor peak | DRAM GB | GB/s | of HBM peak |")
    print("|---|---|---|---|---|---|---|---|---|")
    for k, (us, flop, byt, cnt) in groups.items():
        print(f"| {k} | {cnt} | {us:.0f} | {100 * us / tot_t:.1f} % | {flop / us / 1e6:.0f} | {flop / us / 1e6 / tensor_peak:.2f} | "
              f"{byt / 1e9:.2f} | {byt / us / 1e3:.0f} | {byt / us / 1e3 / hbm_peak:.2f} |")
    print(f"| **all tensor-core launches** | {len(tc)} | {tot_t:.0f} | 100 % | {tot_f / tot_t / 1e6:.0f} | "
          f"**{tot_f / tot_t / 1e6 / tensor_peak:.2f}** | {tot_b / 1e9:.2f} | {tot_b / tot_t / 1e3:.0f} | {tot_b / tot_t / 1e3 / hbm_peak:.2f} |")
    burst = pk["bf16_tflops"] * (0.5 if mode == "tf32" else 1.0)
    print(f"\nncu times each launch alone, at full boost clocks, while the sustained peak is what a long step holds under the power "
          f"cap: single launches can exceed 1.00 of it.  Against the burst figure (bf16_tflops {pk['bf16_tflops']}"
          + (" x 0.5" if mode == "tf32" else "") + f" = {burst:.1f} TFLOP/s) all tensor-core launches together run at "
          f"{tot_f / tot_t / 1e6 / burst:.2f}, the best block at "
          f"{max(f / u / 1e6 for u, f, _, _ in groups.values()) / burst:.2f}.")
    other = [l for l in launches if not fam_of(l["kernel"])]
    print(f"\nOther launches of the step ({len(other)}: layout, conditioning, speaker encoder on its side stream): "
          f"{sum(l['gpu__time_duration.sum'] for l in other) / 1e3:.0f} us serialised, of which the speaker encoder overlaps enc_p in a real step.")


if __name__ == "__main__":
    main()
