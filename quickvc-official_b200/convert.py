"""Batch conversion driver: the reference's `convert.py` (convert.py:19-86) around the B200 path.

Same command line (`--hpfile --ptfile --txtpath --outdir --use_timestamp`), same list format (`title|src|tgt` per
line, convert.py:48-54), same per-utterance arithmetic -- target wave -> trim -> `wave_to_mel` -> speaker embedding;
source wave -> content units; `infer`; float32 WAV at `hps.data.sampling_rate` (convert.py:62-86) -- with the parts the
reference takes from libraries that are not dependencies here restated below:

  * `load_wave`      librosa.load(path, sr=...): mono float32 in [-1, 1]; files already at `sr` are read exactly,
                     others are resampled with a polyphase filter (librosa uses soxr_hq: not bit-identical).
  * `trim_silence`   librosa.effects.trim(y, top_db=20): frame RMS (2048 / 512, centred, zero padded) in dB below the
                     loudest frame.  Restated from the published algorithm (librosa 0.10); librosa is absent from
                     this image and from /root/reference, so this piece is pinned by hand-computed cases only.
  * content units    `hubert_soft.units(wav)` comes from torch.hub `bshall/hubert:main` (convert.py:44,79), which is
                     neither vendored nor reachable offline: the driver takes any callable `wav (1,1,S) -> (1,T,256)`,
                     or reads precomputed units from `--units-dir/<source stem>.{pt,npy}`.

What is B200-specific: the reference converts one utterance at a time and leaves the GPU idle while the host reads,
trims and writes files.  Here every utterance still gets exactly the samples of its own single-utterance call, and
throughput comes from (1) one speaker embedding per distinct target file, (2) ragged batches: utterances sorted by
length go through `infer(..., lengths=...)` up to `max_batch` at a time -- the kernels write zeros past each
utterance's own end, which is the zero padding that utterance's convolutions see alone (include/qvc_b200.h,
qvc_infer `lengths`) -- (3) `streams` CUDA streams in round-robin, each with its own workspace, so that small batches
and single clips overlap on the 148 SMs, and (4) pinned device->host copies that overlap the next batch's kernels, with
the WAV files written once their copy event has fired.  `ragged=False` restricts batches to equal lengths and targets.
"""
from __future__ import annotations

import argparse
import json
import os
import time
from dataclasses import dataclass
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

ContentEncoder = Callable[[Tensor], Tensor]          # wav (1, 1, S) on the device -> units (1, T, 256)


# ---------------------------------------------------------------------------------------------- hyper-parameters
class HParams:
    """Attribute view of the JSON config, nested (utils.py:114-140)."""

    def __init__(self, **kwargs) -> None:
        for k, v in kwargs.items():
            setattr(self, k, HParams(**v) if isinstance(v, dict) else v)

    def keys(self):
        return self.__dict__.keys()

    def __getitem__(self, key):
        return getattr(self, key)

    def __contains__(self, key) -> bool:
        return key in self.__dict__

    def as_dict(self) -> dict:
        return {k: (v.as_dict() if isinstance(v, HParams) else v) for k, v in self.__dict__.items()}


def get_hparams_from_file(config_path: str) -> HParams:
    """utils.py:105-111."""
    with open(config_path, "r") as f:
        return HParams(**json.load(f))


def load_checkpoint(checkpoint_path: str, model: torch.nn.Module, trust: bool = False) -> int:
    """utils.py:148-178 without the optimizer: key-wise load of `checkpoint['model']`, keys missing from the file keep
    the module's values; returns the stored iteration.  The reference checkpoint holds tensors, optimizer state and
    scalars only, so it is unpickled with `weights_only=True`; `trust` (CLI: --trust-checkpoint) allows arbitrary
    pickled objects for files from a source the caller vouches for."""
    if not os.path.isfile(checkpoint_path):
        raise FileNotFoundError(checkpoint_path)
    ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=not trust)
    saved = ckpt["model"]
    state = model.state_dict()
    model.load_state_dict({k: saved.get(k, v) for k, v in state.items()})
    return int(ckpt.get("iteration", 0))


# ---------------------------------------------------------------------------------------------- list + audio files
def read_list(txtpath: str) -> List[Tuple[str, str, str]]:
    """`title|src|tgt` per line (convert.py:48-54; convert.txt:1-12).  Blank lines are skipped."""
    items = []
    with open(txtpath, "r") as f:
        for n, raw in enumerate(f, 1):
            line = raw.strip()
            if not line:
                continue
            parts = line.split("|")
            if len(parts) != 3:
                raise ValueError(f"{txtpath}:{n}: expected title|src|tgt, got {line!r}")   # the reference's unpack raises too
            items.append((parts[0], parts[1], parts[2]))
    return items


def shard_items(items: Sequence, rank: int, world: int) -> List:
    """The share of rank `rank` out of `world` processes (one per GPU): every world-th line, so that long and short
    utterances spread evenly whatever the order of the list.  Utterances are independent: no collective, every rank
    writes its own files."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    return list(items[rank::world])


def load_wave(path: str, sr: int) -> np.ndarray:
    """Mono float32 waveform at `sr` (the role of librosa.load at convert.py:62,65)."""
    from scipy.io import wavfile
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", wavfile.WavFileWarning)
        file_sr, data = wavfile.read(path)
    if data.dtype == np.int16:
        wav = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        wav = (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    elif data.dtype == np.uint8:
        wav = (data.astype(np.float32) - 128.0) / 128.0
    elif data.dtype in (np.float32, np.float64):
        wav = data.astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported sample type {data.dtype}")
    if wav.ndim == 2:
        wav = wav.mean(axis=1, dtype=np.float32)                       # librosa's to_mono
    if file_sr != sr:
        from math import gcd
        from scipy.signal import resample_poly

        g = gcd(int(file_sr), int(sr))
        wav = resample_poly(wav.astype(np.float64), sr // g, file_sr // g).astype(np.float32)
    return np.ascontiguousarray(wav)


def trim_silence(y: np.ndarray, top_db: float = 20.0, frame_length: int = 2048, hop_length: int = 512,
                 pad_mode: str = "constant") -> np.ndarray:
    """Leading / trailing silence removed as librosa.effects.trim(y, top_db=top_db)[0] does (convert.py:63).

    Frame t covers samples [t*hop - frame/2, t*hop + frame/2) of the padded signal; a frame is non-silent when its RMS
    is less than `top_db` dB below the loudest frame's (power floor 1e-10).  Keeps [first*hop, min(len, (last+1)*hop)).
    """
    n = y.shape[-1]
    if n == 0:
        return y
    half = frame_length // 2
    x = np.pad(y.astype(np.float64), (half, half), mode=pad_mode)
    n_frames = 1 + n // hop_length
    # mean square per frame through a prefix sum of squares
    csum = np.concatenate(([0.0], np.cumsum(x * x)))
    starts = np.arange(n_frames) * hop_length
    power = (csum[starts + frame_length] - csum[starts]) / frame_length
    amin = 1e-10
    db = 10.0 * np.log10(np.maximum(amin, power)) - 10.0 * np.log10(max(amin, float(power.max())))
    nz = np.flatnonzero(db > -top_db)
    if nz.size == 0:
        return y[0:0]
    start = int(nz[0]) * hop_length
    end = min(n, (int(nz[-1]) + 1) * hop_length)
    return y[start:end]


def write_wave(path: str, sr: int, audio: np.ndarray) -> None:
    """float32 WAV, as `scipy.io.wavfile.write` at convert.py:86."""
    from scipy.io import wavfile

    wavfile.write(path, sr, np.ascontiguousarray(audio, dtype=np.float32))


# ---------------------------------------------------------------------------------------------- content units
class UnitsDirectory:
    """Content encoder that reads precomputed soft units: `<dir>/<source file stem>.pt` (tensor) or `.npy`,
    shaped (T, 256) or (1, T, 256) -- what `hubert_soft.units(wav)` returned for that source (convert.py:79)."""

    def __init__(self, directory: str) -> None:
        self.directory = directory

    def for_source(self, src: str, device: torch.device) -> Tensor:
        stem = os.path.splitext(os.path.basename(src))[0]
        for ext in (".pt", ".npy"):
            p = os.path.join(self.directory, stem + ext)
            if os.path.isfile(p):
                u = torch.load(p, map_location="cpu", weights_only=True) if ext == ".pt" else torch.from_numpy(np.load(p))
                u = u.to(torch.float32)
                if u.dim() == 2:
                    u = u.unsqueeze(0)
                if u.dim() != 3 or u.shape[0] != 1 or u.shape[2] != 256:
                    raise ValueError(f"{p}: units must be (T, 256) or (1, T, 256), got {tuple(u.shape)}")
                return u.to(device, non_blocking=True)
        raise FileNotFoundError(f"no units for {src!r} under {self.directory} (looked for {stem}.pt / {stem}.npy)")


def load_hubert_soft(device: torch.device) -> ContentEncoder:
    """The reference's content encoder (convert.py:44).  torch.hub needs the network or a populated hub cache."""
    try:
        model = torch.hub.load("bshall/hubert:main", "hubert_soft", trust_repo=True).to(device).eval()
    except Exception as ex:  # noqa: BLE001
        raise RuntimeError("hubert_soft is not available offline (torch.hub bshall/hubert:main); pass --units-dir with "
                           "precomputed soft units, or a content_encoder callable") from ex
    return lambda wav: model.units(wav)


# ---------------------------------------------------------------------------------------------- the driver
@dataclass
class _Item:
    title: str
    tgt: str
    unit: Tensor                      # (1, 256, frames) on the device
    noise: Optional[Tensor]           # (1, 192, frames) or None
    frames: int


@dataclass
class _Pending:
    titles: List[str]
    frames: List[int]
    host: Tensor                      # pinned (B, 1, 320 * max frames)
    done: torch.cuda.Event


class Converter:
    """`convert.py`'s synthesis loop.  `net` is a `quickvc_official_b200.SynthesizerTrn` on a CUDA device."""

    def __init__(self, net, hps: HParams, *, content_encoder: Optional[ContentEncoder] = None,
                 units_dir: Optional[str] = None, streams: int = 2, max_batch: int = 64, ragged: bool = True,
                 max_padding: float = 0.1) -> None:
        self.net = net
        self.hps = hps
        self.device = next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("Converter needs the module on a CUDA device (there is no CPU path)")
        if content_encoder is None and units_dir is None:
            raise ValueError("need a content_encoder callable or a units_dir")
        self.content_encoder = content_encoder
        self.units = UnitsDirectory(units_dir) if units_dir is not None else None
        self.max_batch = max(1, int(max_batch))
        self.ragged = bool(ragged)
        self.max_padding = float(max_padding)
        self._streams = [torch.cuda.Stream(self.device) for _ in range(max(1, int(streams)))]
        self._embeddings: Dict[str, Tensor] = {}
        self.stats = {"utterances": 0, "calls": 0, "audio_seconds": 0.0, "padded_audio_seconds": 0.0, "targets": 0}

    # -- preprocessing, per file -------------------------------------------------------------------
    def speaker_embedding(self, tgt: str) -> Tensor:
        """convert.py:62-64,75-77 + models.py:632: one embedding per distinct target file."""
        g = self._embeddings.get(tgt)
        if g is None:
            from .mel import wave_to_mel

            d = self.hps.data
            wav = trim_silence(load_wave(tgt, d.sampling_rate), top_db=20)
            wav_t = torch.from_numpy(wav).unsqueeze(0).to(self.device)
            mel = wave_to_mel(wav_t, d.filter_length, d.n_mel_channels, d.sampling_rate, d.hop_length, d.win_length,
                              d.mel_fmin, d.mel_fmax)
            g = self.net.embed_speaker(mel)
            self._embeddings[tgt] = g
            self.stats["targets"] += 1
        return g

    def source_units(self, src) -> Tensor:
        """(1, 256, T) on the device (convert.py:65-66,79).  `src` is a file path, or the soft units themselves as a
        (T, 256) / (1, T, 256) tensor (an in-memory caller that already ran its content encoder)."""
        if isinstance(src, Tensor):
            u = src.to(self.device, torch.float32, non_blocking=True)
            u = u.unsqueeze(0) if u.dim() == 2 else u
            if u.dim() != 3 or u.shape[0] != 1 or u.shape[2] != 256:
                raise ValueError(f"units must be (T, 256) or (1, T, 256), got {tuple(src.shape)}")
        elif self.units is not None:
            u = self.units.for_source(src, self.device)
        else:
            wav = load_wave(src, self.hps.data.sampling_rate)
            u = self.content_encoder(torch.from_numpy(wav).unsqueeze(0).unsqueeze(0).to(self.device))
        return u.transpose(2, 1).contiguous()

    # -- conversion ----------------------------------------------------------------------------------
    def _plan(self, loaded: List["_Item"]) -> List[List["_Item"]]:
        """Batches of utterances that go through one `infer` call.

        ragged (default): utterances sorted by length and cut into batches of at most `max_batch`, closed early when
        padding everything to the longest member would waste more than `max_padding` of the batch's frames; each batch
        is one call with `lengths` (per-utterance speaker embeddings, so targets mix freely).
        Otherwise: only utterances with the same target and the same number of frames share a call.
        """
        if not self.ragged:
            groups: Dict[Tuple[str, int], List[_Item]] = {}
            for it in loaded:
                groups.setdefault((it.tgt, it.frames), []).append(it)
            return [m[i:i + self.max_batch] for m in groups.values() for i in range(0, len(m), self.max_batch)]
        batches: List[List[_Item]] = []
        cur: List[_Item] = []
        total = 0
        for it in sorted(loaded, key=lambda x: x.frames):
            if cur and (len(cur) >= self.max_batch or
                        1.0 - (total + it.frames) / float(it.frames * (len(cur) + 1)) > self.max_padding):
                batches.append(cur)
                cur, total = [], 0
            cur.append(it)
            total += it.frames
        if cur:
            batches.append(cur)
        return batches

    @torch.no_grad()
    def convert(self, items: Sequence[Tuple[str, str, str]], *, noise_seed: Optional[int] = None
                ) -> Iterable[Tuple[str, np.ndarray]]:
        """Yields `(title, waveform float32 (320 * frames,))` for every item; the order follows the batches, not the
        list.  `noise_seed` makes the prior's draw (models.py:94) reproducible: item i of the list uses a generator
        seeded with `noise_seed + i`, whatever batch it lands in."""
        sr = self.hps.data.sampling_rate
        main = torch.cuda.current_stream(self.device)
        # 1. speaker embeddings and units on the caller's stream (file reading dominates)
        loaded: List[_Item] = []
        for index, (title, src, tgt) in enumerate(items):
            self.speaker_embedding(tgt)
            u = self.source_units(src)
            noise = None
            if noise_seed is not None:
                gen = torch.Generator(device=self.device)
                gen.manual_seed(int(noise_seed) + index)
                noise = torch.randn((1, 192, u.shape[2]), device=self.device, generator=gen)
            loaded.append(_Item(title, tgt, u, noise, u.shape[2]))
        ready = torch.cuda.Event()
        ready.record(main)
        # 2. one infer per batch, streams in round-robin
        pending: List[_Pending] = []
        for n_call, part in enumerate(self._plan(loaded)):
            s = self._streams[n_call % len(self._streams)]
            frames = [it.frames for it in part]
            t_max = max(frames)
            with torch.cuda.stream(s):
                s.wait_event(ready)
                same_len = min(frames) == t_max
                same_tgt = all(it.tgt == part[0].tgt for it in part)
                if len(part) == 1:
                    unit, noise = part[0].unit, part[0].noise
                elif same_len:
                    unit = torch.cat([it.unit for it in part], dim=0)
                    noise = torch.cat([it.noise for it in part], dim=0) if part[0].noise is not None else None
                else:
                    unit = torch.zeros((len(part), 256, t_max), device=self.device)
                    noise = torch.zeros((len(part), 192, t_max), device=self.device) if part[0].noise is not None else None
                    for b, it in enumerate(part):
                        unit[b, :, :it.frames] = it.unit[0]
                        if noise is not None:
                            noise[b, :, :it.frames] = it.noise[0]
                g = self._embeddings[part[0].tgt] if same_tgt else torch.cat([self._embeddings[it.tgt] for it in part], dim=0)
                lengths = None if same_len else torch.tensor(frames, dtype=torch.int32)
                wave = self.net.infer_with_embedding(unit, g, noise=noise, lengths=lengths)
                host = torch.empty(wave.shape, dtype=torch.float32, pin_memory=True)
                host.copy_(wave, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s)
            pending.append(_Pending([it.title for it in part], frames, host, ev))
            self.stats["calls"] += 1
            self.stats["utterances"] += len(part)
            self.stats["audio_seconds"] += sum(frames) * 320 / sr
            self.stats["padded_audio_seconds"] += len(part) * t_max * 320 / sr
            # hand finished clips back while later ones are still running
            while pending and pending[0].done.query():
                yield from self._emit(pending.pop(0))
        for p in pending:
            p.done.synchronize()
            yield from self._emit(p)
        for s in self._streams:
            main.wait_stream(s)

    @staticmethod
    def _emit(p: "_Pending") -> Iterable[Tuple[str, np.ndarray]]:
        arr = p.host.numpy()
        for b, title in enumerate(p.titles):
            yield title, arr[b, 0, :320 * p.frames[b]]

    def convert_list(self, txtpath: str, outdir: str, *, use_timestamp: bool = False,
                     noise_seed: Optional[int] = None, rank: int = 0, world: int = 1) -> List[str]:
        """convert.py:47-86: reads the list, converts, writes `<outdir>/<title>.wav`; returns the written paths.
        With `world` > 1 (one process per GPU) this rank converts every world-th line of the list."""
        os.makedirs(outdir, exist_ok=True)
        written = []
        for title, audio in self.convert(shard_items(read_list(txtpath), rank, world), noise_seed=noise_seed):
            name = f"{time.strftime('%m-%d_%H-%M', time.localtime())}_{title}.wav" if use_timestamp else f"{title}.wav"
            path = os.path.join(outdir, name)
            write_wave(path, self.hps.data.sampling_rate, audio)
            written.append(path)
        return written


def build_net(hps: HParams, ptfile: Optional[str], device: torch.device, precision: str = "tf32",
              trust_checkpoint: bool = False):
    """convert.py:35-41."""
    from .models import SynthesizerTrn

    net = SynthesizerTrn(hps.data.filter_length // 2 + 1, hps.train.segment_size // hps.data.hop_length,
                         precision=precision, **hps.model.as_dict()).to(device)
    net.eval()
    if ptfile is not None:
        load_checkpoint(ptfile, net, trust=trust_checkpoint)
    return net


def main(argv: Optional[Sequence[str]] = None) -> int:
    parser = argparse.ArgumentParser(description="QuickVC conversion on B200 (the reference's convert.py command line)")
    parser.add_argument("--hpfile", type=str, default="logs/quickvc/config.json", help="path to json config file")
    parser.add_argument("--ptfile", type=str, default="logs/quickvc/quickvc.pth", help="path to pth file")
    parser.add_argument("--txtpath", type=str, default="convert.txt", help="path to txt file")
    parser.add_argument("--outdir", type=str, default="output/quickvc", help="path to output dir")
    parser.add_argument("--use_timestamp", default=False, action="store_true")
    parser.add_argument("--units-dir", type=str, default=None, help="precomputed soft units (<source stem>.pt/.npy) instead of torch.hub hubert_soft")
    parser.add_argument("--precision", type=str, default="tf32", choices=("tf32", "fp16", "bf16", "fp32"))
    parser.add_argument("--streams", type=int, default=2, help="CUDA streams the infer calls rotate over")
    parser.add_argument("--max-batch", type=int, default=64, help="utterances per infer call")
    parser.add_argument("--no-ragged", action="store_true", help="only batch utterances of equal length and target")
    parser.add_argument("--seed", type=int, default=None, help="seed of the prior's noise draw (default: unseeded, as the reference)")
    parser.add_argument("--trust-checkpoint", action="store_true",
                        help="unpickle --ptfile without weights_only (it may then run arbitrary code: trusted files only)")
    args = parser.parse_args(argv)

    # one process per GPU under torchrun: rank r converts lines r, r + world, ... on GPU LOCAL_RANK
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if "LOCAL_RANK" in os.environ:
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    device = torch.device("cuda", torch.cuda.current_device())
    hps = get_hparams_from_file(args.hpfile)
    print("Loading model...")
    net = build_net(hps, args.ptfile, device, args.precision, args.trust_checkpoint)
    print("Number of parameter: %.2fM" % (sum(p.nelement() for p in net.parameters()) / 1e6))
    encoder = None if args.units_dir else load_hubert_soft(device)
    conv = Converter(net, hps, content_encoder=encoder, units_dir=args.units_dir, streams=args.streams,
                     max_batch=args.max_batch, ragged=not args.no_ragged)
    print("Synthesizing...")
    t0 = time.perf_counter()
    written = conv.convert_list(args.txtpath, args.outdir, use_timestamp=args.use_timestamp, noise_seed=args.seed,
                                rank=rank, world=world)
    dt = time.perf_counter() - t0
    print(f"{len(written)} files, {conv.stats['audio_seconds']:.1f} s of audio in {dt:.2f} s "
          f"({conv.stats['calls']} infer calls, {conv.stats['targets']} target speakers)")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
