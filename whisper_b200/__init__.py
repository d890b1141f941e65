"""whisper_b200 - B200-native Whisper inference hot path behind the openai/whisper Python surface.

    import whisper_b200 as whisper
    model = whisper.load_model("large-v3", synthetic=True)      # or a .pt checkpoint path
    result = model.transcribe(audio)                            # audio: 16 kHz float32 waveform

Same entry points as the reference package (whisper/__init__.py:11-15): load_model,
available_models, load_audio, log_mel_spectrogram, pad_or_trim, DecodingOptions, DecodingResult,
decode, detect_language, ModelDimensions, Whisper, transcribe.
"""
from __future__ import annotations

import os
from typing import List, Optional, Union

import torch

from .audio import load_audio, log_mel_spectrogram, pad_or_trim
from .decoding import DecodingOptions, DecodingResult, decode, detect_language
from .model import ModelDimensions, Whisper
from .synthetic import MODEL_DIMS, dims_dict, synthetic_state_dict
from .transcribe import transcribe, transcribe_batch

__version__ = "0.1.0"


def available_models() -> List[str]:
    """Names of the official architectures (reference __init__.py:98-100)."""
    return [k for k in MODEL_DIMS if not k.startswith("test-")]


def load_model(name: str, device: Optional[Union[str, torch.device]] = None, download_root: str = None,
               in_memory: bool = False, *, dtype: torch.dtype = torch.float16, synthetic: bool = False,
               seed: int = 0) -> Whisper:
    """Load a Whisper model (reference __init__.py:103-161).

    `name` is either a path to a checkpoint in the reference's format
    (`{"dims": {...}, "model_state_dict": {...}}`, __init__.py:147-156) or an official model name.
    This build has no network access, so official names resolve to a local file
    `<download_root or ~/.cache/whisper>/<name>.pt` if one exists; with `synthetic=True` they
    resolve to deterministic random weights of that architecture (whisper_b200.synthetic), which
    is what the benchmarks use.
    """
    if device is None:
        device = "cuda"
    if torch.device(device).type != "cuda":
        raise RuntimeError("whisper_b200 runs on CUDA devices only (no CPU path)")
    path = None
    if os.path.isfile(name):
        path = name
    elif name in MODEL_DIMS:
        root = download_root or os.path.join(os.getenv("XDG_CACHE_HOME", os.path.join(os.path.expanduser("~"), ".cache")), "whisper")
        cand = os.path.join(root, f"{name}.pt")
        if os.path.isfile(cand):
            path = cand
        elif not synthetic:
            raise RuntimeError(f"Model {name} not found at {cand}; there is no network here to download it. "
                               f"Pass synthetic=True for random weights of that architecture.")
    else:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    if path is not None:
        checkpoint = torch.load(path, map_location="cpu", weights_only=True)
        dims = ModelDimensions(**checkpoint["dims"])
        return Whisper(dims, checkpoint["model_state_dict"], device=device, dtype=dtype)
    dims = ModelDimensions(**dims_dict(name))
    return Whisper(dims, synthetic_state_dict(dims_dict(name), seed=seed), device=device, dtype=dtype)
