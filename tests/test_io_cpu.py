"""Host-side caller I/O of libsgb200 (no GPU work): the threaded wav writer must produce the bytes
scipy.io.wavfile.write produces (attackMain.py:166) and the reader must return what Dataset.__getitem__
returns (dataset/Dataset.py:72-84: torchaudio.load scale, crop, zero-pad)."""
import os

import numpy as np
import pytest
import torch
from scipy.io import wavfile


def test_writer_is_byte_identical_to_scipy(tmp_path):
    from speakerguard_b200.io import write_wav_batch
    rng = np.random.default_rng(0)
    B, N = 7, 16001
    pcm = torch.from_numpy(rng.integers(-32768, 32768, size=(B, N), dtype=np.int16))
    paths = [str(tmp_path / "ours" / f"spk{i % 3}" / f"spk{i % 3}-utt{i}.wav") for i in range(B)]
    write_wav_batch(paths, pcm, 16000, nthreads=3)
    for i, p in enumerate(paths):
        ref = str(tmp_path / f"ref{i}.wav")
        wavfile.write(ref, 16000, pcm[i].numpy())
        assert open(p, "rb").read() == open(ref, "rb").read()


def test_reader_crop_pad_and_scale(tmp_path):
    from speakerguard_b200.io import read_wav_batch
    rng = np.random.default_rng(1)
    lens = [12000, 16000, 20000, 1]
    paths = []
    for i, n in enumerate(lens):
        p = str(tmp_path / f"u{i}.wav")
        wavfile.write(p, 16000, rng.integers(-32768, 32768, size=n, dtype=np.int16))
        paths.append(p)
    L = 16000
    starts = np.array([0, 0, 1234, 0])
    out, got_lens = read_wav_batch(paths, L, starts, normalize=True, nthreads=2)
    assert got_lens.tolist() == lens
    for i, p in enumerate(paths):
        ref = wavfile.read(p)[1].astype(np.float32) / 32768.0          # torchaudio.load's normalisation of PCM16
        if len(ref) > L:
            ref = ref[starts[i]:starts[i] + L]
        else:
            ref = np.concatenate([ref, np.zeros(L - len(ref), np.float32)])
        assert np.array_equal(out[i].numpy(), ref)
    raw, _ = read_wav_batch(paths, L, starts, normalize=False)
    assert np.array_equal(raw.numpy(), out.numpy() * 32768.0)
    centred, _ = read_wav_batch(paths[2:3], L, None)
    assert np.array_equal(centred[0].numpy(), wavfile.read(paths[2])[1][2000:18000].astype(np.float32) / 32768.0)


def test_reader_matches_torchaudio_when_available(tmp_path):
    torchaudio = pytest.importorskip("torchaudio")
    from speakerguard_b200.io import read_wav_batch
    p = str(tmp_path / "a.wav")
    wavfile.write(p, 16000, np.random.default_rng(2).integers(-32768, 32768, size=8000, dtype=np.int16))
    try:
        ref, _ = torchaudio.load(p)
    except Exception as e:                                                # no decoding backend in this image
        pytest.skip(f"torchaudio.load unavailable: {e}")
    out, _ = read_wav_batch([p], 8000, None)
    assert torch.equal(out, ref)


def test_io_errors(tmp_path):
    from speakerguard_b200 import _lib
    from speakerguard_b200.io import read_wav_batch, write_wav_batch
    with pytest.raises(_lib.SgError, match="cannot read"):
        read_wav_batch([str(tmp_path / "missing.wav")], 100)
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"not a riff file at all")
    with pytest.raises(_lib.SgError, match="cannot read"):
        read_wav_batch([str(bad)], 100)
    ro = tmp_path / "file_not_dir"
    ro.write_bytes(b"x")
    with pytest.raises(_lib.SgError, match="cannot write"):
        write_wav_batch([str(ro / "sub" / "a.wav")], torch.zeros(1, 10, dtype=torch.int16))


def test_loader_walks_tree_like_the_reference_dataset(tmp_path):
    from speakerguard_b200.io import WavBatchLoader
    rng = np.random.default_rng(3)
    root = tmp_path / "data" / "Spk_test"
    names = {}
    for s, spk in enumerate(["1001", "1002", "9999"]):
        os.makedirs(root / spk)
        for u in range(3):
            n = [8000, 16000, 20000][u]
            pcm = rng.integers(-32768, 32768, size=n, dtype=np.int16)
            wavfile.write(str(root / spk / f"{spk}-u{u}.wav"), 16000, pcm)
            names[f"{spk}-u{u}"] = pcm
    np.random.seed(0)
    loader = WavBatchLoader(["1001", "1002"], str(tmp_path / "data"), "Spk_test", wav_length=16000, batch_size=4, device="cpu",
                            prefetch=True)
    assert len(loader) == 3
    seen = {}
    for origin, true, file_name in loader:
        assert origin.shape[1:] == (1, 16000) and origin.dtype == torch.float32 and true.dtype == torch.long
        for o, t, n in zip(origin, true, file_name):
            seen[n] = (o[0].clone(), int(t))
    assert sorted(seen) == sorted(names)
    for n, (o, t) in seen.items():
        spk = n.split("-")[0]
        assert t == {"1001": 0, "1002": 1}.get(spk, -1)
        ref = names[n].astype(np.float32) / 32768.0
        if len(ref) <= 16000:
            assert np.array_equal(o.numpy()[:len(ref)], ref) and not o.numpy()[len(ref):].any()
        else:                                                           # random crop: must be a contiguous window
            w = o.numpy()
            assert any(np.array_equal(w, ref[s:s + 16000]) for s in range(len(ref) - 16000 + 1))
