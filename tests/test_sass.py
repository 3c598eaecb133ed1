"""What the shipped library is made of (no GPU needed): `cuobjdump -sass` of libqvc_b200.so must show, per hot kernel, the
Blackwell instructions the design rests on -- tcgen05 MMAs (UTCHMMA, the .2CTA form for the CTA-pair kernels), TMA tensor
loads (UTMALDG), TMEM loads (LDTM), transaction mbarriers (SYNCS) -- sm_100a code only, and no dependency on cuBLAS, cuDNN
or any other compute library.  A build that silently fell back to CUDA-core code would still pass the parity tests; it does
not pass this one."""
import collections
import os
import re
import shutil
import subprocess

import pytest

from quickvc_official_b200 import capi

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not installed")


KERNELS = ("conv_tcr_kernel", "conv_tc2_kernel", "conv_tc_kernel", "conv_wn_kernel", "post_tail_kernel", "tail_kernel",
           "lstm_recurrent_kernel", "conv_fma_kernel", "spk_project_kernel", "spk_mean_kernel", "to_series_kernel",
           "from_series_kernel", "cond_kernel", "reflect_row_kernel", "mel_log_kernel", "reflect_pad_kernel")


@pytest.fixture(scope="module")
def sass():
    """kernel base name -> Counter of SASS mnemonics over all of its template instances"""
    out = subprocess.run([CUOBJDUMP, "-sass", capi.LIB_PATH], capture_output=True, text=True, check=True, timeout=600).stdout
    arch = set(re.findall(r"arch = (sm_\w+)", out))
    per = collections.defaultdict(collections.Counter)
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            mangled = m.group(1)
            # Itanium mangling: <length><identifier>; the anonymous-namespace prefix carries the file name, so match exactly
            name = next((k for k in KERNELS if f"{len(k)}{k}" in mangled), mangled)
            per[name]["__instances__"] += 1
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            per[name][m.group(1)] += 1
    return arch, per


def _count(counter, prefix):
    return sum(n for op, n in counter.items() if op.startswith(prefix))


def test_only_sm_100a_code(sass):
    arch, _ = sass
    assert arch == {"sm_100a"}, arch


@pytest.mark.parametrize("kernel,pair", [("conv_tcr_kernel", True), ("conv_tc2_kernel", True), ("conv_wn_kernel", True),
                                         ("post_tail_kernel", True), ("conv_tc_kernel", False)])
def test_tensor_core_kernels_are_tcgen05_fed_by_tma(sass, kernel, pair):
    _, per = sass
    c = per[kernel]
    assert c["__instances__"] >= 1, f"{kernel} not in the library"
    assert _count(c, "UTCHMMA") > 0, f"{kernel}: no tcgen05.mma"
    assert _count(c, "UTMALDG") > 0, f"{kernel}: no TMA tensor load"
    assert _count(c, "LDTM") > 0, f"{kernel}: no tcgen05.ld"
    assert _count(c, "SYNCS") > 0, f"{kernel}: no mbarrier"
    assert _count(c, "HMMA") == 0 and _count(c, "IMMA") == 0, f"{kernel}: legacy mma.sync"
    if pair:
        assert _count(c, "UTCHMMA.2CTA") > 0 and _count(c, "UTCBAR.2CTA") > 0, f"{kernel}: not a cta_group::2 kernel"
        assert _count(c, "UCGABAR") > 0, f"{kernel}: no cluster barrier"
    print(kernel, c["__instances__"], "instances;", {k: _count(c, k) for k in ("UTCHMMA", "UTMALDG", "LDTM", "SYNCS", "STG", "LDG")})


def test_persistent_rnn_uses_clusters_and_async_stores(sass):
    _, per = sass
    c = per["lstm_recurrent_kernel"]
    assert c["__instances__"] >= 1
    assert _count(c, "STAS") > 0, "no st.async into the cluster"           # csrc/lstm.cu: a step closes by data arrival
    assert _count(c, "SYNCS") > 0 and _count(c, "UCGABAR") > 0


def test_no_compute_library_is_linked():
    out = subprocess.run(["readelf", "-d", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    needed = re.findall(r"\(NEEDED\)\s+Shared library: \[(.+?)\]", out)
    for lib in needed:
        assert not re.search(r"cublas|cudnn|cufft|cutlass|nccl|torch|c10|triton", lib, re.I), needed
