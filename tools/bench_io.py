#!/usr/bin/env python
"""Caller-I/O throughput: saving a batch of adversarial utterances the reference's way
(attackMain.save_audio: per-utterance torch max/min, x 2^15, .cpu().numpy().astype(int16), scipy.io.wavfile.write in a
Python loop) vs speakerguard_b200.io.save_audio (device PCM16 kernel, one pinned D2H copy, threaded writer), and loading a
batch the reference's way (scipy read per file standing in for torchaudio.load, crop / pad, stack) vs WavBatchLoader's reader.
Files go to /dev/shm so the file system is not the variable.  Prints one JSON line."""
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from scipy.io import wavfile  # noqa: E402


def reference_save(advers, names, root, fs=16000):
    for adver, name in zip(advers[:, 0, :], names):
        if 0.9 * adver.max() <= 1 and 0.9 * adver.min() >= -1:
            adver = adver * (2 ** 15)
        adver = adver.detach().cpu().numpy().astype(np.int16)
        d = os.path.join(root, name.split("-")[0])
        if not os.path.exists(d):
            os.makedirs(d)
        wavfile.write(os.path.join(d, name + ".wav"), fs, adver)


def reference_load(paths, L):
    out = []
    for p in paths:
        a = torch.from_numpy(wavfile.read(p)[1].astype(np.float32) / 32768.0).unsqueeze(0)
        n = a.shape[1]
        if L < n:
            a = a[..., :L]
        elif L > n:
            a = torch.cat((a, torch.zeros(1, L - n)), 1)
        out.append(a)
    return torch.stack(out)


def main():
    from speakerguard_b200.io import read_wav_batch, save_audio
    B, N = int(os.environ.get("SGB200_IO_B", "1024")), 48000
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="sgb200_io_", dir=base)
    try:
        x = ((torch.rand(B, 1, N) * 2 - 1) * 0.5).cuda()
        names = [f"spk{i % 10}-utt{i}" for i in range(B)]
        save_audio(x, names, os.path.join(tmp, "warm"))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        paths = save_audio(x, names, os.path.join(tmp, "ours"))
        t_ours = time.perf_counter() - t0
        t0 = time.perf_counter()
        reference_save(x, names, os.path.join(tmp, "ref"))
        t_ref = time.perf_counter() - t0
        same = all(open(p, "rb").read() == open(p.replace("/ours/", "/ref/"), "rb").read() for p in paths[:64])
        buf = torch.empty(B, N).pin_memory()
        read_wav_batch(paths, N, None, True, buf)
        t0 = time.perf_counter()
        read_wav_batch(paths, N, None, True, buf)
        xd = buf.cuda(non_blocking=True)
        torch.cuda.synchronize()
        t_load = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref = reference_load(paths, N).cuda()
        torch.cuda.synchronize()
        t_load_ref = time.perf_counter() - t0
        print(json.dumps({"batch": B, "samples": N, "cores": os.cpu_count(),
                          "save_utt_per_s": B / t_ours, "save_utt_per_s_reference_loop": B / t_ref, "save_files_identical": same,
                          "load_utt_per_s": B / t_load, "load_utt_per_s_reference_loop": B / t_load_ref,
                          "load_identical": bool(torch.equal(xd.unsqueeze(1), ref))}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
