"""ctypes binding of libqvc_b200.so (include/qvc_b200.h).

The library is the product: there is no Python / PyTorch fallback for any entry point.  Loading fails
loudly when the shared object is missing, and every call raises `QvcError` on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libqvc_b200.so"
LIB_PATH = os.path.join(_HERE, LIB_NAME)

QVC_ABI_VERSION = 6
QVC_NUM_LAYERS = 114

OPF_F32, OPF_TF32, OPF_BF16, OPF_F16 = 0, 1, 2, 3
BACKEND_FMA, BACKEND_TCGEN05 = 0, 1
EPI_LINEAR, EPI_GATE, EPI_SAMPLE = 0, 1, 2


class QvcError(RuntimeError):
    pass


class Tensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("bstride", C.c_int64), ("ld", C.c_int32), ("_pad", C.c_int32)]


class EpiSegment(C.Structure):
    _fields_ = [("col0", C.c_int32), ("ncols", C.c_int32), ("alpha", C.c_float), ("beta", C.c_float),
                ("slope", C.c_float), ("_pad", C.c_int32),
                ("res", Tensor), ("res_op", Tensor), ("res_inv_slope", C.c_float), ("_pad2", C.c_int32),
                ("accin", Tensor), ("raw", Tensor), ("op", Tensor)]


class ConvArgs(C.Structure):
    _fields_ = [("x", Tensor), ("batch", C.c_int32), ("x_rows", C.c_int32), ("out_rows", C.c_int32),
                ("cin", C.c_int32),
                ("w", C.c_void_p), ("bias", C.c_void_p), ("bias_bstride", C.c_int64),
                ("cout", C.c_int32), ("k", C.c_int32), ("dil", C.c_int32), ("pad_left", C.c_int32),
                ("epilogue", C.c_int32), ("nseg", C.c_int32),
                ("seg", EpiSegment * 2),
                ("noise", Tensor), ("aux0", Tensor), ("aux1", Tensor),
                ("opformat", C.c_int32), ("backend", C.c_int32),
                ("live_units", C.c_void_p), ("live_mul", C.c_int32), ("tap_split", C.c_int32),
                ("tap_lo", (C.c_int32 * 2) * 2), ("tap_hi", (C.c_int32 * 2) * 2)]


class SpkWeights(C.Structure):
    _fields_ = [("w_ih", C.c_void_p * 3), ("w_hh", C.c_void_p * 3), ("bias", C.c_void_p * 3),
                ("lin_w", C.c_void_p), ("lin_b", C.c_void_p)]


class MelWeights(C.Structure):
    _fields_ = [("basis", C.c_void_p), ("fbank_t", C.c_void_p), ("n_fft", C.c_int32), ("hop", C.c_int32),
                ("n_mels", C.c_int32), ("_pad", C.c_int32)]


class TailWeights(C.Structure):
    _fields_ = [("window", C.c_void_p), ("synth", C.c_void_p), ("window_host", C.c_void_p), ("synth_host", C.c_void_p)]


class Layer(C.Structure):
    _fields_ = [("w", C.c_void_p), ("bias", C.c_void_p), ("cin", C.c_int32), ("cout", C.c_int32),
                ("k", C.c_int32), ("dil", C.c_int32), ("pad_left", C.c_int32), ("_pad", C.c_int32)]


class Model(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("opformat", C.c_int32), ("backend", C.c_int32),
                ("chunk_utts", C.c_int32),
                ("layers", Layer * QVC_NUM_LAYERS), ("paired", Layer * QVC_NUM_LAYERS), ("wn_skip", Layer * 5),
                ("cond_w", C.c_void_p), ("cond_b", C.c_void_p), ("cond_rows", C.c_int32), ("_pad", C.c_int32),
                ("spk", SpkWeights), ("tail", TailWeights)]


class StateEntry(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64)]


class Taps(C.Structure):
    _fields_ = [("g", C.c_void_p), ("m_p", C.c_void_p), ("logs_p", C.c_void_p), ("z_p", C.c_void_p),
                ("flow", C.c_void_p * 4), ("conv_pre", C.c_void_p), ("ups0", C.c_void_p),
                ("mrf0", C.c_void_p), ("ups1", C.c_void_p), ("mrf1", C.c_void_p),
                ("conv_post", C.c_void_p), ("y_mb", C.c_void_p)]


# every symbol include/qvc_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "qvc_conv1d": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "qvc_conv1d_sum": (C.c_int, [C.POINTER(C.POINTER(ConvArgs)), C.c_int, C.c_void_p]),
    "qvc_wn_layer": (C.c_int, [C.POINTER(ConvArgs), C.POINTER(ConvArgs), C.c_void_p]),
    "qvc_to_series_major": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "qvc_from_series_major": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "qvc_spk_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "qvc_spk_embed": (C.c_int, [C.POINTER(SpkWeights), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_size_t, C.c_void_p]),
    "qvc_mel_frames": (C.c_int, [C.POINTER(MelWeights), C.c_int]),
    "qvc_mel_workspace_bytes": (C.c_size_t, [C.POINTER(MelWeights), C.c_int, C.c_int]),
    "qvc_wave_to_mel": (C.c_int, [C.POINTER(MelWeights), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_size_t, C.c_void_p]),
    "qvc_tail": (C.c_int, [C.POINTER(TailWeights), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "qvc_post_tail": (C.c_int, [C.POINTER(ConvArgs), C.POINTER(TailWeights), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "qvc_infer_workspace_bytes": (C.c_size_t, [C.POINTER(Model), C.c_int, C.c_int, C.c_int, C.c_int]),
    "qvc_infer": (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                            C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(Taps), C.c_void_p, C.c_size_t, C.c_void_p]),
    "qvc_decode": (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                             C.c_void_p, C.POINTER(Taps), C.c_void_p, C.c_size_t, C.c_void_p]),
    "qvc_prepared_bytes": (C.c_size_t, [C.c_int]),
    "qvc_prepare_weights": (C.c_int, [C.POINTER(StateEntry), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                      C.POINTER(Model), C.c_void_p]),
    "qvc_fold_host": (C.c_int, [C.POINTER(StateEntry), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                C.POINTER(Model)]),
    "qvc_last_error": (C.c_char_p, []),
    "qvc_abi_version": (C.c_int, []),
    "qvc_launch_count": (C.c_uint64, []),
    "qvc_last_kernel": (C.c_char_p, []),
    "qvc_check_device": (C.c_int, [C.c_int]),
    "qvc_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "qvc_host_unregister": (C.c_int, [C.c_void_p]),
    "qvc_profile": (C.c_int, [C.c_int]),
    "qvc_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the in-tree library and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QvcError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
            f"(or `make -C {os.path.join(_HERE, 'csrc')}`). There is no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.qvc_abi_version() != QVC_ABI_VERSION:
        raise QvcError(f"{LIB_NAME} has ABI {lib.qvc_abi_version()}, binding expects {QVC_ABI_VERSION}")
    _lib = lib
    return lib


def state_entries(sd):
    """state_dict (CPU tensors) -> (array of qvc_state_entry, keep-alive list): fp32, contiguous, enc_q.* left out."""
    import torch
    keep, rows = [], []
    for k, v in sd.items():
        if k.startswith("enc_q."):
            continue
        t = v.detach().to("cpu", torch.float32).contiguous()
        name = k.encode()
        keep.append((t, name))
        rows.append((name, t.data_ptr(), t.numel()))
    arr = (StateEntry * len(rows))()
    for i, (name, ptr, n) in enumerate(rows):
        arr[i].name, arr[i].data, arr[i].numel = name, ptr, n
    return arr, keep


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().qvc_last_error()
        raise QvcError(f"{what} failed with status {status}: {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().qvc_launch_count())


def last_kernel() -> str:
    return load().qvc_last_kernel().decode()
