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


def _checkpoint_table() -> dict:
    """Per official model name: file name and SHA-256 digest the reference expects in its cache directory (both taken
    from its download URLs, whisper/__init__.py:17-32) and the dump of its alignment heads (:36-51); generated from the
    reference's tables by tools/make_checkpoint_table.py."""
    import json

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "checkpoints.json")) as f:
        return json.load(f)


def _sha256_file(path: str) -> str:
    import hashlib

    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(64 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def _resolve_official(name: str, root: str) -> Optional[str]:
    """The verified local file of an official model, or None if there is none.  The reference's _download
    (whisper/__init__.py:54-95) accepts a cached file only if its SHA-256 matches and otherwise downloads again; there is
    no network here, so a mismatch is an error instead of a silent use of the wrong weights."""
    entry = _checkpoint_table().get(name)
    candidates = ([os.path.join(root, entry["file"])] if entry else []) + [os.path.join(root, f"{name}.pt")]
    for cand in candidates:
        if os.path.exists(cand) and not os.path.isfile(cand):
            raise RuntimeError(f"{cand} exists and is not a regular file")
        if os.path.isfile(cand):
            if entry and _sha256_file(cand) != entry["sha256"]:
                raise RuntimeError(f"{cand} exists, but the SHA256 checksum does not match the official {name} checkpoint "
                                   f"({entry['sha256'][:16]}...); there is no network here to download it again. "
                                   f"Pass the path itself to load it as an unofficial checkpoint.")
            return cand
    return None


def load_model(name: str, device: Optional[Union[str, torch.device]] = None, download_root: str = None,
               in_memory: bool = False, *, dtype: torch.dtype = torch.float16, synthetic: bool = False,
               seed: int = 0) -> Whisper:
    """Load a Whisper model (reference __init__.py:103-161).

    `name` is either a path to a checkpoint in the reference's format
    (`{"dims": {...}, "model_state_dict": {...}}`, __init__.py:147-156) or an official model name.
    This build has no network access, so official names resolve to the file the reference would have cached -
    `<download_root or ~/.cache/whisper>/<file>.pt`, accepted only if its SHA-256 matches the official digest
    (__init__.py:63-71) - and get the reference's alignment heads for word timing (__init__.py:158-159); with
    `synthetic=True` they resolve to deterministic random weights of that architecture (whisper_b200.synthetic), which
    is what the benchmarks use.  `in_memory` preloads the file's bytes like the reference does (:141-145).
    """
    if device is None:
        device = "cuda"
    if torch.device(device).type != "cuda":
        raise RuntimeError("whisper_b200 runs on CUDA devices only (no CPU path)")
    path = None
    official = None
    if os.path.isfile(name):
        path = name
    elif name in MODEL_DIMS:
        root = download_root or os.path.join(os.getenv("XDG_CACHE_HOME", os.path.join(os.path.expanduser("~"), ".cache")), "whisper")
        path = _resolve_official(name, root)
        official = name if path is not None else None
        if path is None and not synthetic:
            raise RuntimeError(f"Model {name} not found under {root}; there is no network here to download it. "
                               f"Pass synthetic=True for random weights of that architecture.")
    else:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    if path is not None:
        if in_memory:
            import io

            with open(path, "rb") as f:
                checkpoint = torch.load(io.BytesIO(f.read()), map_location="cpu", weights_only=True)
        else:
            checkpoint = torch.load(path, map_location="cpu", weights_only=True)
        dims = ModelDimensions(**checkpoint["dims"])
        model = Whisper(dims, checkpoint["model_state_dict"], device=device, dtype=dtype)
        heads = _checkpoint_table().get(official or "", {}).get("alignment_heads")
        if heads:
            model.set_alignment_heads(heads.encode("ascii"))
        return model
    dims = ModelDimensions(**dims_dict(name))
    return Whisper(dims, synthetic_state_dict(dims_dict(name), seed=seed), device=device, dtype=dtype)
