"""Audio front-end with the reference's surface (whisper/audio.py): constants, `load_audio`,
`pad_or_trim`, `mel_filters`, `log_mel_spectrogram` - the last one running as the fused sm_100a
kernel of csrc/mel.cu through the C ABI (wb200_log_mel)."""
from __future__ import annotations

from ctypes import c_int, c_int64, c_size_t
from functools import lru_cache
from subprocess import CalledProcessError, run
from typing import Optional, Union

import numpy as np
import torch
import torch.nn.functional as F

from ._lib import check, lib, ptr, stream_ptr

# hard-coded audio hyperparameters (reference audio.py:13-22)
SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE  # 480000 samples in a 30-second chunk
N_FRAMES = N_SAMPLES // HOP_LENGTH  # 3000 frames in a mel spectrogram input
N_SAMPLES_PER_TOKEN = HOP_LENGTH * 2  # the initial convolutions have stride 2
FRAMES_PER_SECOND = SAMPLE_RATE // HOP_LENGTH  # 10 ms per audio frame
TOKENS_PER_SECOND = SAMPLE_RATE // N_SAMPLES_PER_TOKEN  # 20 ms per audio token


def load_audio(file: str, sr: int = SAMPLE_RATE):
    """Decode + down-mix + resample a file to a mono float32 waveform with the ffmpeg CLI
    (reference audio.py:25-62; host-side plumbing, unchanged in spirit)."""
    cmd = ["ffmpeg", "-nostdin", "-threads", "0", "-i", file, "-f", "s16le", "-ac", "1",
           "-acodec", "pcm_s16le", "-ar", str(sr), "-"]
    try:
        out = run(cmd, capture_output=True, check=True).stdout
    except CalledProcessError as e:
        raise RuntimeError(f"Failed to load audio: {e.stderr.decode()}") from e
    except FileNotFoundError as e:
        raise RuntimeError("Failed to load audio: the ffmpeg executable was not found") from e
    return np.frombuffer(out, np.int16).flatten().astype(np.float32) / 32768.0


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """Pad with zeros or trim `array` to `length` along `axis` (reference audio.py:65-88)."""
    if torch.is_tensor(array):
        if array.shape[axis] > length:
            array = array.narrow(axis if axis >= 0 else array.dim() + axis, 0, length)
        if array.shape[axis] < length:
            pad = [0, 0] * array.dim()
            ax = axis if axis >= 0 else array.dim() + axis
            pad[2 * (array.dim() - 1 - ax) + 1] = length - array.shape[axis]
            array = F.pad(array, pad)
    else:
        if array.shape[axis] > length:
            array = array.take(indices=range(length), axis=axis)
        if array.shape[axis] < length:
            widths = [(0, 0)] * array.ndim
            widths[axis] = (0, length - array.shape[axis])
            array = np.pad(array, widths)
    return array


def _slaney_mel_filterbank(n_mels: int, sr: int = SAMPLE_RATE, n_fft: int = N_FFT) -> np.ndarray:
    """The Slaney-style, area-normalised triangular filterbank the reference ships as
    assets/mel_filters.npz (audio.py:91-107 documents how it was produced); regenerated here from
    the published formula and checked bit-for-bit against that asset in tests/test_host_logic.py."""
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fft_freqs = np.linspace(0, sr / 2, 1 + n_fft // 2)
    mel_pts = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2), n_mels + 2))
    widths = np.diff(mel_pts)
    ramps = np.subtract.outer(mel_pts, fft_freqs)
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    for i in range(n_mels):
        weights[i] = np.maximum(0, np.minimum(-ramps[i] / widths[i], ramps[i + 2] / widths[i + 1]))
    weights *= (2.0 / (mel_pts[2: n_mels + 2] - mel_pts[:n_mels]))[:, np.newaxis]
    return weights


@lru_cache(maxsize=None)
def mel_filters(device, n_mels: int) -> torch.Tensor:
    """Mel filterbank matrix (n_mels, 201) on `device` (reference audio.py:91-107)."""
    assert n_mels in {80, 128}, f"Unsupported n_mels: {n_mels}"
    return torch.from_numpy(_slaney_mel_filterbank(n_mels)).to(device)


def log_mel_spectrogram(
    audio: Union[str, np.ndarray, torch.Tensor],
    n_mels: int = 80,
    padding: int = 0,
    device: Optional[Union[str, torch.device]] = None,
    per_waveform_max: bool = False,
):
    """Log-mel spectrogram, shape (*, n_mels, n_frames) (reference audio.py:110-157).

    The computation always runs on the GPU (there is no CPU path): a CPU input is moved to
    `device` (default "cuda") and the result stays there.  `per_waveform_max=True` clamps each
    waveform of a batch against its own maximum, i.e. what the reference returns when called once
    per waveform; the default reproduces the reference's batched call (one global maximum).
    """
    if not torch.is_tensor(audio):
        if isinstance(audio, str):
            audio = load_audio(audio)
        audio = torch.from_numpy(np.ascontiguousarray(audio))
    if device is not None:
        audio = audio.to(device)
    if not audio.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("whisper_b200.log_mel_spectrogram needs a CUDA device (no CPU path)")
        audio = audio.cuda()
    audio = audio.to(torch.float32)
    lead = audio.shape[:-1]
    x = audio.reshape(-1, audio.shape[-1])
    if padding > 0:
        x = F.pad(x, (0, padding))
    x = x.contiguous()
    n_audio, n_samples = x.shape
    n_frames = n_samples // HOP_LENGTH
    out = torch.empty((n_audio, n_mels, n_frames), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        ws_bytes = int(lib().wb200_log_mel_workspace_bytes(c_int(n_audio)))
        ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
        check(lib().wb200_log_mel(ptr(x), c_int(n_audio), c_int64(n_samples), c_int(n_mels),
                                  ptr(mel_filters(x.device, n_mels)), ptr(out), ptr(ws), c_size_t(ws_bytes),
                                  c_int(int(per_waveform_max)), stream_ptr()), "wb200_log_mel")
    return out.reshape(*lead, n_mels, n_frames)
