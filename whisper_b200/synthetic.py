"""Deterministic synthetic checkpoints and audio for parity tests and benchmarks.

There is no network in the build or bench containers, so no released Whisper checkpoint can be
downloaded; BASELINE.md's measurement plan calls for random-init weights of the real architecture.
The generator below is independent of torch's RNG and of the reference package: every tensor is
drawn from a numpy PCG64 stream seeded by (seed, tensor name), so the GPU box, this container and
the reference (`load_state_dict`) all see bit-identical weights.

All Linear / Conv / Embedding weights are rounded so that they are exactly representable in BOTH
bf16 and fp16 (8-bit significand, magnitudes below 2^-14 flushed to zero): the fp32 oracle and the
16-bit GPU path then hold identical weight values and differ only in activation rounding.

State-dict key names follow the reference's module tree (whisper/model.py:142-275) so that the
dict loads into the reference `Whisper` unchanged.
"""
from __future__ import annotations

import zlib
from typing import Dict, List, Tuple

import numpy as np

# name -> (n_mels, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer,
#          n_vocab, n_text_ctx, n_text_state, n_text_head, n_text_layer)   [SURVEY.md dims table]
MODEL_DIMS = {
    "tiny.en": (80, 1500, 384, 6, 4, 51864, 448, 384, 6, 4),
    "tiny": (80, 1500, 384, 6, 4, 51865, 448, 384, 6, 4),
    "base.en": (80, 1500, 512, 8, 6, 51864, 448, 512, 8, 6),
    "base": (80, 1500, 512, 8, 6, 51865, 448, 512, 8, 6),
    "small.en": (80, 1500, 768, 12, 12, 51864, 448, 768, 12, 12),
    "small": (80, 1500, 768, 12, 12, 51865, 448, 768, 12, 12),
    "medium.en": (80, 1500, 1024, 16, 24, 51864, 448, 1024, 16, 24),
    "medium": (80, 1500, 1024, 16, 24, 51865, 448, 1024, 16, 24),
    "large-v1": (80, 1500, 1280, 20, 32, 51865, 448, 1280, 20, 32),
    "large-v2": (80, 1500, 1280, 20, 32, 51865, 448, 1280, 20, 32),
    "large-v3": (128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 32),
    "large": (128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 32),
    "large-v3-turbo": (128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 4),
    "turbo": (128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 4),
    # small test-only shapes (not released models): multilingual ids, 128 mels, 2+2 layers
    "test-multi": (128, 1500, 256, 4, 2, 51866, 448, 256, 4, 2),
    "test-en": (80, 1500, 128, 2, 2, 51864, 448, 128, 2, 2),
}
DIM_FIELDS = ("n_mels", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer",
              "n_vocab", "n_text_ctx", "n_text_state", "n_text_head", "n_text_layer")


def dims_dict(name: str) -> Dict[str, int]:
    return dict(zip(DIM_FIELDS, MODEL_DIMS[name]))


def round_to_16bit_common(x: np.ndarray) -> np.ndarray:
    """Round fp32 to an 8-bit significand (bf16, round-to-nearest-even) and flush |x| < 2^-14 to 0,
    so the value is exact in both bf16 and fp16."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).copy()                 # finite inputs only: the +0x7FFF cannot carry out of 32 bits
    lsb = (u >> np.uint32(16)) & np.uint32(1)
    u += np.uint32(0x7FFF)
    u += lsb
    u &= np.uint32(0xFFFF0000)
    y = u.view(np.float32)
    y[np.abs(y) < 2.0 ** -14] = 0.0
    np.clip(y, -60000.0, 60000.0, out=y)
    return y


def sinusoid_table(length: int, channels: int, max_timescale: float = 10000.0) -> np.ndarray:
    """Encoder positional table; same formula as whisper/model.py:62-68, evaluated in fp32."""
    inc = np.float32(np.log(max_timescale) / (channels // 2 - 1))
    inv = np.exp(-inc * np.arange(channels // 2, dtype=np.float32)).astype(np.float32)
    t = np.arange(length, dtype=np.float32)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)


def state_dict_spec(dims: Dict[str, int]) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) for every persistent tensor of the reference state dict."""
    da, dt = dims["n_audio_state"], dims["n_text_state"]
    spec: List[Tuple[str, Tuple[int, ...], str]] = [
        ("encoder.conv1.weight", (da, dims["n_mels"], 3), "conv"),
        ("encoder.conv1.bias", (da,), "bias"),
        ("encoder.conv2.weight", (da, da, 3), "conv"),
        ("encoder.conv2.bias", (da,), "bias"),
        ("encoder.positional_embedding", (dims["n_audio_ctx"], da), "sinusoid"),
    ]

    def block(prefix: str, d: int, cross: bool):
        out = []
        atts = ["attn"] + (["cross_attn"] if cross else [])
        for att in atts:
            out += [
                (f"{prefix}.{att}.query.weight", (d, d), "linear"),
                (f"{prefix}.{att}.query.bias", (d,), "bias"),
                (f"{prefix}.{att}.key.weight", (d, d), "linear"),
                (f"{prefix}.{att}.value.weight", (d, d), "linear"),
                (f"{prefix}.{att}.value.bias", (d,), "bias"),
                (f"{prefix}.{att}.out.weight", (d, d), "linear"),
                (f"{prefix}.{att}.out.bias", (d,), "bias"),
                (f"{prefix}.{att}_ln.weight", (d,), "ln_w"),
                (f"{prefix}.{att}_ln.bias", (d,), "ln_b"),
            ]
        out += [
            (f"{prefix}.mlp.0.weight", (4 * d, d), "linear"),
            (f"{prefix}.mlp.0.bias", (4 * d,), "bias"),
            (f"{prefix}.mlp.2.weight", (d, 4 * d), "linear"),
            (f"{prefix}.mlp.2.bias", (d,), "bias"),
            (f"{prefix}.mlp_ln.weight", (d,), "ln_w"),
            (f"{prefix}.mlp_ln.bias", (d,), "ln_b"),
        ]
        return out

    for i in range(dims["n_audio_layer"]):
        spec += block(f"encoder.blocks.{i}", da, False)
    spec += [("encoder.ln_post.weight", (da,), "ln_w"), ("encoder.ln_post.bias", (da,), "ln_b")]
    spec += [
        ("decoder.token_embedding.weight", (dims["n_vocab"], dt), "embed"),
        ("decoder.positional_embedding", (dims["n_text_ctx"], dt), "pos"),
    ]
    for i in range(dims["n_text_layer"]):
        spec += block(f"decoder.blocks.{i}", dt, True)
    spec += [("decoder.ln.weight", (dt,), "ln_w"), ("decoder.ln.bias", (dt,), "ln_b")]
    return spec


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def synthetic_state_dict(dims: Dict[str, int], seed: int = 0, linear_gain: float = 2.0,
                         logit_std: float = 8.0, eot_scale: float = 8.0,
                         timestamp_scale: float = 3.5, row_sigma: float = 0.5) -> Dict[str, np.ndarray]:
    """Random weights at the given dims, shaped so that decoding is NOT degenerate:

    * Linear layers have gain `linear_gain` (> 1), so the residual stream is dominated by what the
      blocks compute (attention over the audio, MLPs) rather than by the input token embedding -
      with tied embeddings a weak network would just echo its last token for ever.
    * The token embedding has std logit_std / sqrt(d): logits are ~N(0, logit_std^2), peaky enough
      that greedy / beam decisions usually have margins far above 16-bit rounding noise.
    * The <|endoftext|> row is scaled by `eot_scale` and the 1501 timestamp rows by
      `timestamp_scale`, so sequences end naturally after tens of tokens and the timestamp rules
      (pairs, monotonicity, "timestamp mass beats best text token") actually fire.
    """
    sd: Dict[str, np.ndarray] = {}
    n_vocab = dims["n_vocab"]
    multilingual = n_vocab >= 51865
    eot = 50257 if multilingual else 50256
    ts_begin = n_vocab - 1501
    embed_std = logit_std / np.sqrt(dims["n_text_state"])
    for name, shape, kind in state_dict_spec(dims):
        g = _rng(seed, name)
        if kind == "sinusoid":
            w = sinusoid_table(*shape)
        elif kind == "linear":
            w = g.standard_normal(shape, dtype=np.float32) * np.float32(linear_gain / np.sqrt(shape[1]))
            w = round_to_16bit_common(w)
        elif kind == "conv":
            w = g.standard_normal(shape, dtype=np.float32) * np.float32(1.0 / np.sqrt(shape[1] * shape[2]))
            w = round_to_16bit_common(w)
        elif kind == "bias":
            w = round_to_16bit_common(g.standard_normal(shape, dtype=np.float32) * np.float32(0.05))
        elif kind == "ln_w":
            w = (1.0 + 0.1 * g.standard_normal(shape, dtype=np.float32)).astype(np.float32)
            if name == "decoder.ln.weight":
                # random signs: with tied embeddings, an all-positive final gain makes the logit of
                # the token just fed in systematically the largest (the model echoes itself)
                w = w * np.where(g.random(shape) < 0.5, -1.0, 1.0).astype(np.float32)
        elif kind == "ln_b":
            w = (0.05 * g.standard_normal(shape, dtype=np.float32)).astype(np.float32)
        elif kind == "embed":
            w = g.standard_normal(shape, dtype=np.float32) * np.float32(embed_std)
            # log-normal row norms: a heavy-tailed logit distribution, like a trained LM's (the
            # top-1 / top-2 gap is then O(1) relative to the top logit instead of O(1/ln V))
            w *= np.exp(row_sigma * g.standard_normal((shape[0], 1), dtype=np.float32))
            w[eot] *= np.float32(eot_scale)
            w[ts_begin:] *= np.float32(timestamp_scale)
            w = round_to_16bit_common(w)
        elif kind == "pos":
            w = round_to_16bit_common(g.standard_normal(shape, dtype=np.float32) * np.float32(1.0))
        else:  # pragma: no cover
            raise ValueError(kind)
        sd[name] = w
    return sd


def synthetic_audio(n_audio: int, n_samples: int = 480000, seed: int = 1234, kind: str = "noise") -> np.ndarray:
    """Synthetic 16 kHz waveforms in [-1, 1], fp32, shape (n_audio, n_samples).

    kind="noise":  0.1 * N(0,1) clipped.
    kind="speechlike": sum of five sinusoids 100-4000 Hz under a 4 Hz envelope plus weak noise, which
    exercises the max-8 dynamic-range clamp of log_mel_spectrogram non-trivially.
    """
    out = np.empty((n_audio, n_samples), dtype=np.float32)
    t = np.arange(n_samples, dtype=np.float64) / 16000.0
    for i in range(n_audio):
        g = np.random.Generator(np.random.PCG64([seed, i]))
        if kind == "noise":
            x = 0.1 * g.standard_normal(n_samples)
        elif kind == "speechlike":
            f = g.uniform(100.0, 4000.0, size=5)
            ph = g.uniform(0, 2 * np.pi, size=5)
            amp = g.uniform(0.02, 0.15, size=5)
            x = sum(a * np.sin(2 * np.pi * fi * t + p) for a, fi, p in zip(amp, f, ph))
            x = x * (0.55 + 0.45 * np.sin(2 * np.pi * 4.0 * t + g.uniform(0, 6.28)))
            x = x + 1e-3 * g.standard_normal(n_samples)
        else:
            raise ValueError(kind)
        out[i] = np.clip(x, -1.0, 1.0).astype(np.float32)
    return out
