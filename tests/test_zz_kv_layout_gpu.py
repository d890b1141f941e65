"""GPU: the head-major kv-cache layout (wb200_set_kv_head_major) must be invisible in the results.

The layout only changes WHERE the two cp.async decode-attention kernels (and the cross-K/V projection's epilogue) put
the same 16-bit values, and the order of every reduction is untouched, so with those kernels a decode in either layout
must be equal bit for bit.  (The TMA attention kernels and the one-launch decoder stack, which exist for the head-major
layout only and reduce in another order, are switched off for this comparison; tests/test_xattn_tma_gpu.py and
tests/test_fused_layer_gpu.py cover them.)
"""
import numpy as np
import pytest
import torch

from helpers import fixture_inputs, load_model_fixture

pytestmark = pytest.mark.gpu


def _setup(name):
    import whisper_b200 as wb

    meta, _ = load_model_fixture(name)
    dims, sd, audio = fixture_inputs(meta)
    model = wb.Whisper(wb.ModelDimensions(**dims), sd, device="cuda", dtype=torch.float16)
    mel = torch.stack([wb.log_mel_spectrogram(torch.from_numpy(a).cuda(), dims["n_mels"]) for a in audio])
    return wb, model, mel


@pytest.mark.parametrize("name", ["test-en", "tiny.en"])
def test_head_major_layout_is_bit_identical(name):
    from whisper_b200 import _lib

    wb, model, mel = _setup(name)
    feats = model.embed_audio(mel)
    cases = [dict(beam_size=5, sample_len=24), dict(sample_len=24), dict(beam_size=3, sample_len=16, prompt=[1000, 1001, 1002]),
             dict(temperature=0.8, best_of=3, seed=5, sample_len=16)]
    toks = torch.tensor([[50257, 50362, 1000, 2000, 3000, 50256]] * feats.shape[0])
    heads = [(model.dims.n_text_layer - 1, 0), (model.dims.n_text_layer - 1, 1)]

    def run_all():
        out = []
        for o in cases:
            res = model.decode(feats, wb.DecodingOptions(language="en", **o))
            out.append([(r.tokens, r.avg_logprob, r.no_speech_prob) for r in res])
        logits, qk = model.logits(toks, feats, alignment_heads=heads)
        return out, logits.clone(), qk.clone()

    try:
        _lib.lib().wb200_set_cross_attention_tma(0)
        _lib.lib().wb200_set_self_attention_tma(0)
        _lib.lib().wb200_set_fused_decoder_stack(0)
        _lib.lib().wb200_set_kv_head_major(0)
        model.clear_sessions()
        base, base_logits, base_qk = run_all()
        _lib.lib().wb200_set_kv_head_major(1)
        model.clear_sessions()
        hm, hm_logits, hm_qk = run_all()
    finally:
        _lib.lib().wb200_set_kv_head_major(1)
        _lib.lib().wb200_set_cross_attention_tma(1)
        _lib.lib().wb200_set_fused_decoder_stack(1)
        model.clear_sessions()                       # (the beam-window self-attention stays off: that is its default)
    assert hm == base
    assert torch.equal(hm_logits, base_logits)
    assert torch.equal(hm_qk, base_qk)
