"""Shared helpers for the parity tests (fixtures, synthetic inputs, oracle plumbing)."""
import json
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_model_fixture(name):
    with open(os.path.join(GOLD, f"model_{name}.json")) as f:
        meta = json.load(f)
    arrays = dict(np.load(os.path.join(GOLD, f"model_{name}.npz")))
    return meta, arrays


def fixture_inputs(meta):
    """(dims, numpy state dict, audio) exactly as oracle/make_golden.py built them."""
    from whisper_b200 import synthetic

    dims = meta["dims"]
    sd = synthetic.synthetic_state_dict(dims, seed=meta["seed"], **meta["synth_kwargs"])
    audio = synthetic.synthetic_audio(2, 480000, seed=meta["audio_seed"], kind=meta["audio_kind"])
    return dims, sd, audio


def oracle_options(opts):
    from oracle import decoding as OD

    o = dict(opts)
    if o.get("suppress_tokens") == "":
        o["suppress_tokens"] = ()
    return OD.Options(**o)


_ORACLE_CACHE = {}


def oracle_features(name):
    """(meta, arrays, dims, W, mel, feats) with the oracle encoder output cached per test session."""
    if name not in _ORACLE_CACHE:
        from oracle import audio as OA
        from oracle import model as OM

        meta, arrays = load_model_fixture(name)
        dims, sd, audio = fixture_inputs(meta)
        W = OM.to_weights(sd)
        mel = torch.from_numpy(np.stack([OA.log_mel_spectrogram(a, dims["n_mels"]) for a in audio]))
        feats = OM.encoder_forward(W, dims, mel)
        _ORACLE_CACHE[name] = (meta, arrays, dims, W, mel, feats)
    return _ORACLE_CACHE[name]
