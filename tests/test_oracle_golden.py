"""CPU: the oracle (oracle/) against the golden vectors produced by running the reference
(oracle/make_golden.py).  This is what pins the oracle: every stage of the hot path, on the same
synthetic inputs, must reproduce the reference's outputs - token ids exactly, floating-point
values to the stated tolerances."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLD, load_model_fixture, oracle_features, oracle_options

MEL_TOL = 1e-4          # fp32 mel, post-scaling range <= 2.0


@pytest.mark.parametrize("kind", ["noise", "speechlike"])
@pytest.mark.parametrize("n_mels", [80, 128])
def test_log_mel(kind, n_mels):
    from oracle import audio as OA
    from whisper_b200 import synthetic

    g = np.load(os.path.join(GOLD, f"mel_{kind}.npz"))
    audio = synthetic.synthetic_audio(2, 480000, seed=1234, kind=kind)
    mel = OA.log_mel_spectrogram(audio, n_mels)                  # one global max over the batch
    assert mel.shape == (2, n_mels, 3000)
    assert np.abs(mel[:, :, ::8] - g[f"batch_{n_mels}"]).max() < MEL_TOL
    assert np.abs(mel[:, :, :64] - g[f"batch_{n_mels}_head"]).max() < MEL_TOL
    assert np.abs(mel[:, :, -64:] - g[f"batch_{n_mels}_tail"]).max() < MEL_TOL
    s, mx, mn = g[f"batch_{n_mels}_sum"]
    assert abs(mel.astype(np.float64).sum() - s) < 1e-6 * abs(s) + 5.0
    assert abs(mel.max() - mx) < MEL_TOL and abs(mel.min() - mn) < MEL_TOL
    single = OA.log_mel_spectrogram(audio[1, :160000], n_mels, padding=480000)   # transcribe.py:139
    assert tuple(single.shape) == tuple(g[f"single_{n_mels}_shape"])
    assert np.abs(single[:, ::8] - g[f"single_{n_mels}"]).max() < MEL_TOL


def test_mel_dynamic_range_property():
    """tests/test_audio.py:19 of the reference: mel.max() - mel.min() <= 2.0."""
    from oracle import audio as OA
    from whisper_b200 import synthetic

    mel = OA.log_mel_spectrogram(synthetic.synthetic_audio(1, 176000, seed=5, kind="speechlike")[0], 80)
    assert mel.max() - mel.min() <= 2.0


def test_timing_golden():
    from oracle import timing as OT

    g = np.load(os.path.join(GOLD, "timing.npz"))
    for i in range(4):
        x = g[f"med_in_{i}"]
        for w in (3, 5, 7, 13):
            assert np.array_equal(OT.median_filter(x, w), g[f"med_out_{i}_{w}"]), (i, w)
        assert np.array_equal(OT.dtw(g[f"dtw_in_{i}"]), g[f"dtw_out_{i}"]), i


def test_token_ids_table():
    import json

    from oracle import decoding as OD

    with open(os.path.join(GOLD, "token_ids.json")) as f:
        table = json.load(f)
    for n_vocab, spec in table["specials"].items():
        ids = OD.token_ids(int(n_vocab))
        for k in ("eot", "sot", "translate", "transcribe", "sot_lm", "sot_prev", "no_speech",
                  "no_timestamps", "timestamp_begin"):
            assert getattr(ids, k) == spec[k], (n_vocab, k)
        assert list(ids.sot_sequence("en", "transcribe")) == spec["sot_sequence"]
        # the reference enumerates a Python set (tokenizer.py:222-228): order is arbitrary, compare as sets
        assert sorted(ids.all_language_tokens) == sorted(spec["all_language_tokens"])
        assert len(ids.non_speech) == spec["n_non_speech"]


@pytest.mark.parametrize("name", ["test-en", "test-multi"])
def test_encoder_and_prefill(name):
    from oracle import decoding as OD
    from oracle import model as OM

    meta, arrays, dims, W, mel, feats = oracle_features(name)
    assert np.abs(feats[:, ::25].numpy() - arrays["feats_sub"]).max() < 2e-4
    ids = OD.token_ids(dims["n_vocab"])
    init = torch.tensor([list(ids.sot_sequence("en", "transcribe"))] * 2)
    logits = OM.decoder_forward(W, dims, init, feats)
    assert np.array_equal(logits[:, -1].topk(16).indices.numpy(), arrays["logits0_last_top_idx"])
    assert np.abs(logits[:, -1, ::97].numpy() - arrays["logits0_last_sub"]).max() < 2e-3
    assert np.abs(logits[:, 0, ::97].numpy() - arrays["logits0_sot_sub"]).max() < 2e-3


def _decode_cases(name):
    meta, _ = load_model_fixture(name)
    return sorted(meta["decode"].keys())


@pytest.mark.parametrize("name,case", [(n, c) for n in ("test-en", "test-multi") for c in _decode_cases(n)])
def test_decode_matches_reference(name, case):
    """Token ids bit-identical to the reference's decode(); avg_logprob / no_speech_prob to 1e-5."""
    from oracle import decoding as OD

    meta, arrays, dims, W, mel, feats = oracle_features(name)
    c = meta["decode"][case]
    res = OD.decode(W, dims, feats[: c["n_audio"]], oracle_options(c["options"]))
    for r, g in zip(res, c["results"]):
        assert r.tokens == g["tokens"]
        assert abs(r.avg_logprob - g["avg_logprob"]) < 1e-5
        assert abs(r.no_speech_prob - g["no_speech_prob"]) <= 1e-5 * max(g["no_speech_prob"], 1e-30) + 1e-12


@pytest.mark.parametrize("case", ["translate", "translate_de_beam", "auto_language", "auto_language_beam", "lang_id",
                                  "french_prompt"])
def test_decode_task_and_language_options(case):
    """decode() with task="translate", other language tokens, language=None (detect_language inside decode,
    decoding.py:666-678) and task="lang_id" (:722-727) against the reference's results."""
    from oracle import decoding as OD

    with open(os.path.join(GOLD, "decode_extra_test-multi.json")) as f:
        gold = json.load(f)
    meta, arrays, dims, W, mel, feats = oracle_features("test-multi")
    c = gold["cases"][case]
    res = OD.decode(W, dims, feats[: c["n_audio"]], OD.Options(**c["options"]))
    assert len(res) == len(c["results"])
    for r, g in zip(res, c["results"]):
        assert r.language == g["language"]
        assert r.tokens == g["tokens"]
        if g["avg_logprob"] is not None:
            assert abs(r.avg_logprob - g["avg_logprob"]) < 1e-5
            assert abs(r.no_speech_prob - g["no_speech_prob"]) <= 1e-5 * max(g["no_speech_prob"], 1e-30) + 1e-12
        if g["top_language_prob"] is not None:
            assert abs(r.top_language_prob - g["top_language_prob"]) < 1e-5


def test_detect_language():
    from oracle import decoding as OD

    meta, arrays, dims, W, mel, feats = oracle_features("test-multi")
    toks, probs = OD.detect_language(W, dims, feats)
    assert toks.tolist() == meta["detect_language"]["tokens"]
    assert np.allclose(probs.max(dim=-1).values.numpy(), meta["detect_language"]["top_prob"], atol=1e-5)


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (kat_vectors of the reference implementation): pins the
    generator the sampling contract (include/whisper_b200.h: wb200_decoder_set_sampling) is stated on."""
    from oracle.decoding import philox4x32_10

    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(int(x) for x in philox4x32_10(np.array(ctr), key)) == want
    batch = philox4x32_10(np.array([k[0] for k in kat[:1]] * 3), (0, 0))
    assert batch.shape == (3, 4) and (batch == batch[0]).all()


def test_gumbel_max_is_categorical():
    """The Gumbel-max draw of the sampling contract is distributed as Categorical(softmax(logits / T))
    (what decoding.py:283 samples) and never picks a filtered (-inf) token."""
    import torch
    from oracle.decoding import gumbel_noise, sample_update

    logits = np.full(12, -np.inf, dtype=np.float32)
    live = [1, 4, 5, 9, 11]
    logits[live] = [0.3, 1.7, -0.5, 2.2, 0.9]
    T = 0.8
    p = np.exp(logits[live] / T)
    p /= p.sum()
    n = 4000
    counts = np.zeros(12)
    for step in range(n):
        key = np.where(np.isneginf(logits), -np.inf, logits / np.float32(T) + gumbel_noise(77, 3, step, 12))
        counts[int(np.argmax(key))] += 1
    assert counts[[i for i in range(12) if i not in live]].sum() == 0
    chi2 = float((((counts[live] - n * p) ** 2) / (n * p)).sum())
    assert chi2 < 18.5, (chi2, counts[live], n * p)          # 4 dof, p ~ 1e-3
    # noise differs across rows, steps and seeds, and is reproducible
    assert not np.array_equal(gumbel_noise(1, 0, 0, 64), gumbel_noise(1, 1, 0, 64))
    assert not np.array_equal(gumbel_noise(1, 0, 0, 64), gumbel_noise(1, 0, 1, 64))
    assert not np.array_equal(gumbel_noise(1, 0, 0, 64), gumbel_noise(2, 0, 0, 64))
    assert np.array_equal(gumbel_noise(1 << 40, 5, 9, 64), gumbel_noise(1 << 40, 5, 9, 64))
    # sample_update: un-tempered log-probability accumulated, rows that ended keep emitting eot
    lg = torch.from_numpy(np.stack([logits, logits]))
    sums = torch.zeros(2)
    toks, done, gaps = sample_update([[7, 8], [7, 0]], lg, sums, eot=0, temperature=T, seed=5)
    assert toks[1][-1] == 0 and float(sums[1]) == 0.0 and toks[0][-1] in live and len(gaps) == 2
    want = float(torch.log_softmax(lg[0], -1)[toks[0][-1]])
    assert abs(float(sums[0]) - want) < 1e-6


def test_margins_include_the_timestamp_rule_distance():
    """The diagnostic decision margins the GPU parity tests gate on must see a near-tie of ApplyTimestampRules' "timestamps
    outweigh every text token" comparison (reference decoding.py:498-505): a flip there masks the whole text vocabulary
    while the top-2 gap of the filtered logits stays large (found on hardware: step margin 47.8, decision flipped)."""
    from oracle import decoding as OD

    ids = OD.token_ids(51864)
    tb = ids.timestamp_begin
    sample_begin = 3
    tokens = [[ids.sot, 50362, 50363, 1000, 1001]]          # two text tokens sampled: text and timestamps both allowed
    base = torch.full((1, 51864), -30.0)
    base[0, 2000] = 5.0                                      # best text token
    for eps, flips in ((+0.02, True), (-0.02, False)):
        logits = base.clone()
        logits[0, tb + 10] = 5.0 + eps                       # one dominant timestamp: lse(ts) ~ 5 + eps vs text max 5
        gaps = []
        OD.timestamp_rules(logits, tokens, ids, sample_begin, None, gaps)
        assert len(gaps) == 1 and gaps[0] < 0.05             # the decision was close ...
        assert bool(torch.isinf(logits[0, 2000])) == flips   # ... and it decides whether any text token survives
        top2 = logits.topk(2, dim=-1).values[0]
        assert float(top2[0] - top2[1]) > 10.0 or not flips  # while the filtered top-2 gap looks perfectly safe
    far = base.clone()
    far[0, tb + 10] = -5.0
    gaps = []
    OD.timestamp_rules(far, tokens, ids, sample_begin, None, gaps)
    assert gaps[0] > 5.0
