"""Oracle for the audio front-end (reference whisper/audio.py).  TEST INFRASTRUCTURE ONLY.

numpy, float64 internally, one rounding to fp32 at the end.
"""
from __future__ import annotations

import os

import numpy as np

SAMPLE_RATE = 16000      # audio.py:13
N_FFT = 400              # audio.py:14
HOP_LENGTH = 160         # audio.py:15
N_SAMPLES = 480000       # audio.py:17
N_FRAMES = 3000          # audio.py:18

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def mel_filters(n_mels: int) -> np.ndarray:
    """The reference's filterbank asset (audio.py:91-107), read from the golden copy of
    whisper/assets/mel_filters.npz made by oracle/make_golden.py."""
    assert n_mels in (80, 128), f"Unsupported n_mels: {n_mels}"  # audio.py:103
    with np.load(os.path.join(_GOLDEN, "mel_filters.npz")) as f:
        return f[f"mel_{n_mels}"].astype(np.float32)


def pad_or_trim(array: np.ndarray, length: int = N_SAMPLES, axis: int = -1) -> np.ndarray:
    """audio.py:65-88 (numpy branch)."""
    if array.shape[axis] > length:
        array = array.take(indices=range(length), axis=axis)
    if array.shape[axis] < length:
        pad = [(0, 0)] * array.ndim
        pad[axis] = (0, length - array.shape[axis])
        array = np.pad(array, pad)
    return array


def log_mel_spectrogram(audio: np.ndarray, n_mels: int = 80, padding: int = 0) -> np.ndarray:
    """audio.py:110-157: right-pad, periodic Hann(400), STFT(n_fft=400, hop=160, center=True with
    reflect padding), drop the last frame, |.|^2, mel projection, log10(clamp 1e-10),
    max(., global_max - 8) where the max is over the WHOLE array passed in, (.+4)/4."""
    x = np.asarray(audio, dtype=np.float64)
    if padding > 0:
        x = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(0, padding)])                 # audio.py:145-146
    lead = x.shape[:-1]
    x = x.reshape(-1, x.shape[-1])
    n = np.arange(N_FFT, dtype=np.float64)
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)                        # torch.hann_window (periodic)
    xp = np.pad(x, [(0, 0), (N_FFT // 2, N_FFT // 2)], mode="reflect")          # torch.stft center=True
    n_frames = 1 + (xp.shape[-1] - N_FFT) // HOP_LENGTH
    idx = np.arange(N_FFT)[None, :] + HOP_LENGTH * np.arange(n_frames)[:, None]
    out = []
    filt = mel_filters(n_mels).astype(np.float64)
    for row in xp:
        frames = row[idx] * window[None, :]
        spec = np.fft.rfft(frames, axis=-1)                                       # (frames, 201)
        power = (spec.real ** 2 + spec.imag ** 2)[:-1]                            # audio.py:149 drop last frame
        out.append(filt @ power.T)                                                # audio.py:151-152
    mel = np.stack(out)
    log_spec = np.log10(np.maximum(mel, 1e-10))                                   # audio.py:154
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)                         # audio.py:155
    log_spec = (log_spec + 4.0) / 4.0                                             # audio.py:156
    return log_spec.reshape(*lead, n_mels, -1).astype(np.float32)
