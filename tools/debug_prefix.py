"""Debug: tiny.en, one audio, greedy with a 5-token prefix, bf16 - teacher-forced logit error per step for the few-rows chains
and the one-launch stack, and the free-running tokens of both against the reference tokens with the oracle's margins."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from helpers import fixture_inputs, load_model_fixture, oracle_features, oracle_options
import whisper_b200 as wb
from whisper_b200 import _lib
from whisper_b200.decoding import DecodingOptions
from oracle import decoding as OD
from oracle import parity

name, case = "tiny.en", "greedy_prefix"
meta, arrays = load_model_fixture(name)
c = meta["decode"][case]
_, _, dims, W, _, feats = oracle_features(name)
rec = {}
o_res = OD.decode(W, dims, feats[:1], oracle_options(c["options"]), record=rec)
ref = c["results"][0]["tokens"]
margins = o_res[0].step_margins
print("ref tokens", ref)
print("margins", [round(float(m), 2) for m in margins])
dims_, sd, audio = fixture_inputs(meta)
for dtype in (torch.float16, torch.bfloat16):
    model = wb.Whisper(wb.ModelDimensions(**dims_), sd, device="cuda", dtype=dtype)
    mel = torch.stack([wb.log_mel_spectrogram(torch.from_numpy(a).cuda(), dims_["n_mels"]) for a in audio])[:1]
    g_feats = model.embed_audio(mel)
    opts = dict(c["options"])
    for mode, label in ((0, "chains"), (1, "stack")):
        _lib.lib().wb200_set_fused_decoder_stack(mode)
        model.clear_sessions()
        task, sess = parity.open_session(model, opts, 1, g_feats)
        errs, scales = [], []
        try:
            for i in range(len(rec["raw_logits"])):
                if i > 0:
                    sess.step()
                lg = sess.get_logits(1).float().cpu()
                r = rec["raw_logits"][i]
                errs.append(float((lg - r).abs().max()))
                scales.append(float(r.abs().max()))
                sess.set_logits(r)
                sess.select()
        finally:
            sess.close()
        got = model.decode(mel, DecodingOptions(language="en", **c["options"]))[0]
        first = next((i for i, (a, b) in enumerate(zip(got.tokens, ref)) if a != b), None)
        print(f"{dtype} {label}: max|logit| {max(scales):.1f}; abs logit err per step {[round(e, 2) for e in errs]}")
        print(f"   free-running == ref: {got.tokens == ref}, first difference at {first}", (got.tokens[first], ref[first], round(float(margins[first]), 3)) if first is not None else "")
    _lib.lib().wb200_set_fused_decoder_stack(1)
    del model
