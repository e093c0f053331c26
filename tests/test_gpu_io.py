"""GPU tests of the caller-I/O pieces: device PCM16 conversion with save_audio's range heuristic and numpy's
cast semantics (attackMain.py:154-160), and the end-to-end drop-in ``save_audio``."""
import os
import warnings

import numpy as np
import pytest
import torch
from scipy.io import wavfile

pytestmark = pytest.mark.gpu


def reference_int16(advers: np.ndarray) -> np.ndarray:
    """attackMain.save_audio's conversion, restated with numpy (bits = 16)."""
    out = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                                   # numpy warns on the out-of-range casts it wraps
        for adver in advers:
            if 0.9 * adver.max() <= 1 and 0.9 * adver.min() >= -1:
                adver = adver * (2 ** 15)
            out.append(adver.astype(np.int16))
    return np.stack(out)


def batch():
    g = torch.Generator().manual_seed(0)
    N = 48000
    rows = [
        (torch.rand(N, generator=g) * 2 - 1) * 0.5,                        # ordinary [-1,1] audio
        torch.rand(N, generator=g) * 2 - 1,                                # full scale
        (torch.rand(N, generator=g) * 2 - 1) * 1.1,                        # 0.9*max <= 1 still true: scaled, wraps above 1.0
        (torch.rand(N, generator=g) * 2 - 1) * 20000.0,                    # already int16 range: not scaled
        (torch.rand(N, generator=g) * 2 - 1) * 1.2,                        # between the two regimes: not scaled -> 0 / +-1
        torch.zeros(N),
    ]
    rows[1][7], rows[1][8] = 1.0, -1.0                                     # +1.0 * 32768 wraps to -32768 in the reference
    x = torch.stack(rows)
    return x


def test_quantize_matches_numpy_semantics():
    from speakerguard_b200.io import quantize_pcm16
    x = batch()
    pcm, scaled = quantize_pcm16(x.cuda().unsqueeze(1))
    ref = reference_int16(x.numpy())
    assert scaled.cpu().tolist() == [True, True, True, False, False, True]
    assert np.array_equal(pcm.cpu().numpy(), ref)
    assert pcm.cpu().numpy()[1, 7] == -32768                                # the reference's wrap-around quirk is kept


def test_save_audio_drop_in(tmp_path):
    from speakerguard_b200.io import save_audio
    x = batch()
    names = [f"spk{i % 2}-utt{i}" for i in range(x.shape[0])]
    paths = save_audio(x.cuda().unsqueeze(1), names, str(tmp_path / "adv"))
    ref = reference_int16(x.numpy())
    for i, (p, n) in enumerate(zip(paths, names)):
        assert p == os.path.join(str(tmp_path / "adv"), n.split("-")[0], n + ".wav")
        fs, data = wavfile.read(p)
        assert fs == 16000 and np.array_equal(data, ref[i])
        r = str(tmp_path / f"ref{i}.wav")
        wavfile.write(r, 16000, ref[i])
        assert open(p, "rb").read() == open(r, "rb").read()


def test_loader_to_attack_to_save_round_trip(tmp_path):
    """attackMain's loop (attackMain.py:306-333) on a synthetic corpus: load -> FGSM -> save -> reload."""
    from oracle import sg_oracle as O
    from speakerguard_b200.attack.FGSM import FGSM
    from speakerguard_b200.io import WavBatchLoader, save_audio
    from speakerguard_b200.model.xv_plda import xv_plda
    from speakerguard_b200.synthetic import state_dict_of, write_xv_model_files
    p = O.make_xv_params(seed=0)
    f = write_xv_model_files(p, str(tmp_path / "model"))
    model = xv_plda(state_dict_of(p), f["plda.txt"], f["mean.vec"], f["transform.txt"], model_file=f["speaker_model"],
                    device="cuda:0", dither="off")
    rng = np.random.default_rng(5)
    root = tmp_path / "data" / "Spk10_test"
    for spk in model.spk_ids[:3]:
        os.makedirs(root / spk)
        for u in range(2):
            wavfile.write(str(root / spk / f"{spk}-u{u}.wav"), 16000, (rng.standard_normal(20000) * 3000).astype(np.int16))
    np.random.seed(1)
    loader = WavBatchLoader(model.spk_ids, str(tmp_path / "data"), "Spk10_test", normalize=True, wav_length=16000, batch_size=4)
    attacker = FGSM(model, epsilon=0.002, batch_size=4, verbose=0)
    total = 0
    for origin, true, file_name in loader:
        assert origin.is_cuda and origin.shape[1:] == (1, 16000)
        adver, success = attacker.attack(origin, true)
        paths = save_audio(adver, file_name, str(tmp_path / "adv"))
        for pth, a in zip(paths, adver):
            fs, data = wavfile.read(pth)
            assert np.array_equal(data, reference_int16(a.cpu().numpy())[0])
            assert np.abs(data.astype(np.float32) / 32768.0 - a[0].cpu().numpy()).max() <= 1.0 / 32768.0
        total += len(paths)
    assert total == 6
