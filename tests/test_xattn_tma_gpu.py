"""GPU: the persistent TMA attention kernels of the decoder step (csrc/dec_attention.cu: cross_attention_tma_kernel,
self_attention_tma_kernel) against the cp.async kernels they replace and, through the model tests that run with them
by default, against the oracle.  Old and new kernels split and merge the keys differently (and round P to 16 bits per
16-key slice), so they agree to that noise, not bit for bit."""
import pytest
import torch

from helpers import fixture_inputs, load_model_fixture, oracle_features

pytestmark = pytest.mark.gpu


def _logits_run(model, g_feats, rec, opts, n_audio):
    from oracle import parity

    task, sess = parity.open_session(model, opts, n_audio, g_feats)
    G = task.n_group
    out = []
    try:
        for i in range(len(rec["raw_logits"])):
            if i > 0:
                sess.step()
            out.append(sess.get_logits(n_audio if i == 0 else n_audio * G).float().cpu())
            ref = rec["raw_logits"][i]
            sess.set_logits(ref[::G] if i == 0 else ref)
            sess.select()
    finally:
        sess.close()
    return out


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("name,opts", [("test-en", dict(beam_size=5, sample_len=12)), ("tiny.en", dict(sample_len=12)),
                                       ("test-multi", dict(beam_size=3, sample_len=10))])
@pytest.mark.parametrize("which", ["cross", "self"])
def test_tma_attention_matches_cp_async(which, name, opts, dtype):
    import whisper_b200 as wb
    from oracle import parity
    from whisper_b200 import _lib

    meta, arrays, dims, W, mel, feats = oracle_features(name)
    rec = parity.oracle_record(W, dims, feats, opts, 2)
    _, sd, audio = fixture_inputs(meta)
    model = wb.Whisper(wb.ModelDimensions(**dims), sd, device="cuda", dtype=dtype)
    g_mel = torch.stack([wb.log_mel_spectrogram(torch.from_numpy(a).cuda(), dims["n_mels"]) for a in audio])
    g_feats = model.embed_audio(g_mel)
    switch = _lib.lib().wb200_set_cross_attention_tma if which == "cross" else _lib.lib().wb200_set_self_attention_tma
    n_audio = 2 if which == "cross" else 40      # the beam-window kernel needs enough (audio, head) items to fill the SMs
    if which == "self":
        if not opts.get("beam_size"):
            pytest.skip("greedy decoding always uses the warp-per-row kernel")
        from oracle import audio as OA
        from oracle import model as OM
        from whisper_b200 import synthetic
        audio = synthetic.synthetic_audio(n_audio, 480000, seed=7, kind="speechlike")
        mel_o = torch.from_numpy(__import__("numpy").stack([OA.log_mel_spectrogram(a, dims["n_mels"]) for a in audio]))
        rec = parity.oracle_record(W, dims, OM.encoder_forward(W, dims, mel_o), dict(opts, sample_len=6), n_audio)
        opts = dict(opts, sample_len=6)
        g_mel = torch.stack([wb.log_mel_spectrogram(torch.from_numpy(a).cuda(), dims["n_mels"]) for a in audio])
        g_feats = model.embed_audio(g_mel)
    try:
        # few-rows sessions run their attention inside the one-launch decoder stack: keep the stand-alone kernels in play
        _lib.lib().wb200_set_fused_decoder_stack(0)
        switch(0)
        model.clear_sessions()
        old = _logits_run(model, g_feats, rec, opts, n_audio)
        switch(1)
        model.clear_sessions()
        new = _logits_run(model, g_feats, rec, opts, n_audio)
    finally:
        switch(1 if which == "cross" else 0)         # the defaults: TMA cross-attention on, beam-window self-attention off
        _lib.lib().wb200_set_fused_decoder_stack(1)
        model.clear_sessions()
    worst, worst_ora = 0.0, 0.0
    G = opts.get("beam_size") or 1
    for i, (a, b) in enumerate(zip(old, new)):
        assert bool(torch.isfinite(b).all()), f"step {i}: non-finite logits"
        worst = max(worst, float((a - b).abs().max() / a.abs().max()))
        ref = rec["raw_logits"][i]
        ref = ref[::G] if i == 0 else ref
        worst_ora = max(worst_ora, float((b - ref).abs().max() / ref.abs().max()))
    assert worst_ora < (2.5e-3 if dtype == torch.float16 else 2.5e-2), f"TMA {which} attention vs the oracle: {worst_ora}"
    tol = 2e-3 if dtype == torch.float16 else 1.6e-2        # P is rounded to 16 bits per 16-key slice in either kernel
    print(f"{name} {dtype}: TMA vs cp.async {which} attention, worst |dlogit| / max|logit| = {worst:.6f} over {len(new)} steps")
    assert 0.0 < worst < tol, "kernels must differ by rounding only (and must not be the same kernel)"
