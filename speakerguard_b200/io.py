"""Caller I/O around ``attacker.attack``, with the reference's names and semantics:

* ``save_audio``        attackMain.py:154-166  float batch -> PCM16 wav files under ``root/<spk_id>/<name>.wav``
* ``WavBatchLoader``    dataset/Dataset.py:20-87 + ``DataLoader(dataset, batch_size, num_workers=0)``
                        (attackMain.py:186-190): directory walk, label lookup, crop / zero-pad to ``wav_length``

The float -> int16 conversion runs on the GPU (``sg_pcm16_quantize``), the files are written / read by a pool of
host threads in libsgb200 (``sg_wav_write_batch`` / ``sg_wav_read_batch``) through one pinned batch buffer, and the
loader prefetches the next batch on a background thread while the current one is attacked.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check
from .engine import Engine, _ptr, default_engine


def _c_paths(paths: Sequence[str]):
    arr = (C.c_char_p * len(paths))()
    arr[:] = [os.fsencode(p) for p in paths]
    return arr


def quantize_pcm16(advers: torch.Tensor, engine: Optional[Engine] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """advers [B,N] or [B,1,N] float (CUDA) -> (pcm int16 [B,N], scaled bool [B]) on the device."""
    if not advers.is_cuda:
        raise _lib.SgError("quantize_pcm16 needs a CUDA tensor (no CPU fallback)")
    a = advers[:, 0, :] if advers.dim() == 3 else advers
    a = a.detach().to(torch.float32).contiguous()
    eng = engine or default_engine(a.device)
    B, N = a.shape
    pcm = torch.empty(B, N, device=a.device, dtype=torch.int16)
    scaled = torch.empty(B, device=a.device, dtype=torch.int32)
    check(eng.lib.sg_pcm16_quantize(eng._h, _ptr(a), B, N, _ptr(pcm), _ptr(scaled), eng.stream), "sg_pcm16_quantize")
    return pcm, scaled.bool()


def write_wav_batch(paths: Sequence[str], pcm: torch.Tensor, fs: int = 16000, nthreads: int = 0) -> None:
    """pcm: host int16 [B,N] (contiguous; pinned or not)."""
    assert pcm.dtype == torch.int16 and not pcm.is_cuda and pcm.is_contiguous()
    B, N = pcm.shape
    assert len(paths) == B
    check(_lib.load().sg_wav_write_batch(_c_paths(paths), pcm.data_ptr(), B, N, int(fs), int(nthreads)), "sg_wav_write_batch")


def read_wav_batch(paths: Sequence[str], wav_length: int, starts: Optional[np.ndarray] = None, normalize: bool = True,
                   out: Optional[torch.Tensor] = None, nthreads: int = 0) -> Tuple[torch.Tensor, np.ndarray]:
    """-> (audio [B,wav_length] float32 host tensor (``out`` if given), lens [B])."""
    B = len(paths)
    if out is None:
        out = torch.empty(B, wav_length, dtype=torch.float32)
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, wav_length)
    lens = np.zeros(B, dtype=np.int32)
    st = None if starts is None else np.ascontiguousarray(starts, dtype=np.int64)
    check(_lib.load().sg_wav_read_batch(_c_paths(paths), B, int(wav_length), None if st is None else st.ctypes.data,
                                        int(bool(normalize)), out.data_ptr(), lens.ctypes.data, int(nthreads)), "sg_wav_read_batch")
    return out, lens


_PINNED = {}


def _pinned_i16(B: int, N: int) -> torch.Tensor:
    buf = _PINNED.get((B, N))
    if buf is None:
        buf = _PINNED[(B, N)] = torch.empty(B, N, dtype=torch.int16).pin_memory()
    return buf


def save_audio(advers, names, root, fs=16000, engine: Optional[Engine] = None, nthreads: int = 0):
    """Drop-in for attackMain.save_audio: advers (B,1,N) in [-1,1] (or int16 range), names like 'spk-utt'."""
    pcm_dev, _ = quantize_pcm16(advers, engine)
    host = _pinned_i16(*pcm_dev.shape)
    host.copy_(pcm_dev, non_blocking=True)
    torch.cuda.current_stream(pcm_dev.device).synchronize()
    paths = [os.path.join(root, n.split("-")[0], n + ".wav") for n in names]
    write_wav_batch(paths, host, fs, nthreads)
    return paths


class WavBatchLoader:
    """``for origin, true, file_name in loader`` as in attackMain.py:306 (origin [B,1,wav_length] on ``device``,
    true int64 [B], file_name list of stems).  ``spk_ids`` maps directory names to labels, unknown speakers get -1
    (dataset/Dataset.py:66-70).  Files longer than ``wav_length`` are cropped at a random offset drawn from
    ``numpy.random`` like the reference (:78), shorter ones zero-padded."""

    def __init__(self, spk_ids, root, name, normalize=True, wav_length=48000, batch_size=1024, device="cuda",
                 nthreads: int = 0, prefetch: bool = True):
        self.spk_ids = list(spk_ids)
        self.root = os.path.join(root, name)
        if not os.path.isdir(self.root):
            raise FileNotFoundError(f"{self.root} does not exist (the engine does not download datasets)")
        self.audio_paths: List[Tuple[str, str]] = []
        for spk_id in os.listdir(self.root):
            for audio_name in os.listdir(os.path.join(self.root, spk_id)):
                self.audio_paths.append((spk_id, audio_name))
        self.normalize, self.wav_length, self.batch_size = normalize, int(wav_length), int(batch_size)
        self.device, self.nthreads, self.prefetch = torch.device(device), nthreads, prefetch
        self._bufs = [torch.empty(self.batch_size, self.wav_length, dtype=torch.float32) for _ in range(2)]
        if self.device.type == "cuda":
            self._bufs = [b.pin_memory() for b in self._bufs]

    def __len__(self):
        return (len(self.audio_paths) + self.batch_size - 1) // self.batch_size

    def _load(self, b: int, slot: int):
        items = self.audio_paths[b * self.batch_size:(b + 1) * self.batch_size]
        paths = [os.path.join(self.root, s, a) for s, a in items]
        buf = self._bufs[slot][:len(items)]
        # first pass reads the lengths for the random crop offsets; the files stay in the page cache for the second
        _, lens = read_wav_batch(paths, self.wav_length, None, self.normalize, buf, self.nthreads)
        if (lens > self.wav_length).any():
            starts = np.array([np.random.choice(n - self.wav_length + 1) if n > self.wav_length else 0 for n in lens])
            read_wav_batch(paths, self.wav_length, starts, self.normalize, buf, self.nthreads)
        labels = torch.tensor([self.spk_ids.index(s) if s in self.spk_ids else -1 for s, _ in items], dtype=torch.long)
        names = [os.path.splitext(a)[0] for _, a in items]
        return buf, labels, names

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, List[str]]]:
        n = len(self)
        nxt = [None]

        def fetch(b, slot):
            nxt[0] = self._load(b, slot)

        th = None
        if n:
            fetch(0, 0)
        for b in range(n):
            buf, labels, names = nxt[0]
            origin = buf.to(self.device, non_blocking=True).unsqueeze(1)
            true = labels.to(self.device, non_blocking=True)
            if self.device.type == "cuda":
                torch.cuda.current_stream(self.device).synchronize()     # the pinned slot is re-used two batches later
            if b + 1 < n:
                if self.prefetch:
                    th = threading.Thread(target=fetch, args=(b + 1, (b + 1) & 1))
                    th.start()
                else:
                    fetch(b + 1, (b + 1) & 1)
            yield origin, true, names
            if th is not None:
                th.join()
                th = None


def attack_stream(attacker, batches, device=None, out: Optional[List[torch.Tensor]] = None):
    """Run ``attacker.attack`` over an iterable of HOST batches ``(x [B,1,N] float32, y [B] int64)`` with the copies off
    the critical path (the attackMain.py:306-333 loop: load batch -> attack -> save).

    The host->device copy of batch k+1 and the device->host copy of the adversarial batch k-1 run on a side stream while
    batch k is attacked on the current stream; pinned host tensors make both copies asynchronous (pageable ones still
    work, they just serialise).  Yields ``(adv_host [B,1,N], success list)`` per batch, in order.  ``out``: optional list
    of pinned host tensors to receive the adversarial batches (cycled; at least 2), else pinned buffers are allocated.
    """
    it = iter(batches)
    first = next(it, None)
    if first is None:
        return
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    side = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)

    def upload(b):
        x, y = b
        with torch.cuda.stream(side):
            xd, yd = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
        return xd, yd, ev

    nxt = upload(first)
    pending = None                                  # (adv_host, success, event) of the previous batch
    k = 0
    while nxt is not None:
        xd, yd, ev = nxt
        b = next(it, None)
        nxt = upload(b) if b is not None else None  # overlaps the attack below
        main.wait_event(ev)
        xd.record_stream(main); yd.record_stream(main)
        adv, success = attacker.attack(xd, yd)      # synchronises on the decisions at its end
        if out is not None:
            host = out[k % len(out)]
        else:
            host = torch.empty(adv.shape, dtype=adv.dtype, pin_memory=True)
        done = torch.cuda.Event()
        with torch.cuda.stream(side):
            side.wait_event(_record(main))
            adv.record_stream(side)
            host.copy_(adv, non_blocking=True)
            done.record(side)
        if pending is not None:
            pending[2].synchronize()
            yield pending[0], pending[1]
        pending = (host, success, done)
        k += 1
    pending[2].synchronize()
    yield pending[0], pending[1]


def _record(stream):
    ev = torch.cuda.Event()
    ev.record(stream)
    return ev
