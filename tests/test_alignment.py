"""find_alignment tensor part (reference timing.py:176-216): oracle vs the reference's golden (CPU) and
the CUDA path vs the oracle (GPU)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLD, fixture_inputs, load_model_fixture


def _setup():
    from oracle import audio as OA
    from oracle import decoding as OD
    from oracle import model as OM
    from whisper_b200 import synthetic

    g = np.load(os.path.join(GOLD, "alignment_test-en.npz"))
    meta, _ = load_model_fixture("test-en")
    dims, sd, _ = fixture_inputs(meta)
    audio = synthetic.synthetic_audio(1, 480000, seed=4321, kind="speechlike")
    ids = OD.token_ids(dims["n_vocab"])
    text_tokens = g["text_tokens"].tolist()
    tokens = [*ids.sot_sequence("en"), ids.no_timestamps, *text_tokens, ids.eot]
    return g, dims, sd, audio, ids, text_tokens, tokens, OA, OM


def test_oracle_alignment_matches_reference():
    from oracle import timing as OT

    g, dims, sd, audio, ids, text_tokens, tokens, OA, OM = _setup()
    W = OM.to_weights(sd)
    mel = torch.from_numpy(OA.log_mel_spectrogram(audio, dims["n_mels"]))
    feats = OM.encoder_forward(W, dims, mel)
    qks = []
    logits = OM.decoder_forward(W, dims, torch.tensor([tokens]), feats, collect_qk=qks)[0]
    heads = g["heads"]                      # rows are (layer, head) pairs
    qk = np.stack([qks[l][0, h].numpy() for l, h in heads])
    n_frames = int(g["num_frames"]) // 2
    matrix = OT.alignment_matrix(qk, n_frames)[len(ids.sot_sequence("en")): -1]
    assert np.abs(matrix - g["matrix"]).max() < 2e-4
    path = OT.dtw(-matrix)
    assert np.array_equal(path[0], g["text_indices"]) and np.array_equal(path[1], g["time_indices"])
    probs = torch.softmax(logits[1:, : ids.eot], dim=-1)[np.arange(len(text_tokens)), text_tokens].numpy()
    assert np.allclose(probs, g["token_probs"], rtol=1e-3, atol=1e-7)


@pytest.mark.gpu
def test_gpu_alignment_vs_oracle():
    import whisper_b200 as wb
    from oracle import timing as OT
    from whisper_b200 import timing as WT

    g, dims, sd, audio, ids, text_tokens, tokens, OA, OM = _setup()
    W = OM.to_weights(sd)
    o_mel = torch.from_numpy(OA.log_mel_spectrogram(audio, dims["n_mels"]))
    o_feats = OM.encoder_forward(W, dims, o_mel)
    qks = []
    o_logits = OM.decoder_forward(W, dims, torch.tensor([tokens]), o_feats, collect_qk=qks)[0]
    heads = [tuple(x) for x in g["heads"].tolist()]
    o_qk = np.stack([qks[l][0, h].numpy() for l, h in heads])

    model = wb.Whisper(wb.ModelDimensions(**dims), sd, device="cuda", dtype=torch.float16)
    assert model.alignment_heads.indices().T.tolist() == [list(h) for h in heads]     # default heads, model.py:268-276
    feats = model.embed_audio(wb.log_mel_spectrogram(torch.from_numpy(audio).cuda(), dims["n_mels"]))
    logits, qk = model.logits(torch.tensor([tokens]), feats, alignment_heads=heads)
    # un-cached full-sequence forward: logits at every position, scores of the alignment heads
    lerr = float((logits[0].cpu() - o_logits).abs().max() / o_logits.abs().max())
    assert lerr < 2.5e-3, f"full-forward logits rel err {lerr}"
    qerr = float(np.abs(qk.cpu().numpy() - o_qk).max())
    assert qerr < 0.05 * max(1.0, float(np.abs(o_qk).max()) / 10), f"qk max err {qerr} (max |qk| {np.abs(o_qk).max():.2f})"

    n_frames = int(g["num_frames"]) // 2
    # the post-processing kernels on the ORACLE's scores: tight tolerance, then bit-exact DTW on that matrix
    m_gpu = WT.alignment_matrix(torch.from_numpy(o_qk).cuda().contiguous(), n_frames).cpu().numpy()
    assert np.abs(m_gpu - OT.alignment_matrix(o_qk, n_frames)).max() < 2e-4
    sub = np.ascontiguousarray(-m_gpu[1:-1])
    assert np.array_equal(WT.dtw(torch.from_numpy(sub).cuda()), OT.dtw_gpu_tiebreak(sub))
    # negate flag == negating afterwards
    m_neg = WT.alignment_matrix(torch.from_numpy(o_qk).cuda().contiguous(), n_frames, negate=True).cpu().numpy()
    assert np.array_equal(m_neg, -m_gpu)

    # end to end through find_alignment (ids-only tokenizer: word splitting degenerates, timings still come out)
    tk = wb.tokenizer.get_tokenizer(False)
    words = WT.find_alignment(model, tk, text_tokens, None, int(g["num_frames"]), audio_features=feats)
    assert len(words) >= 1 and all(w.end >= w.start for w in words)
    assert 1 <= sum(len(w.tokens) for w in words) <= len(text_tokens) + 1


@pytest.mark.gpu
def test_gpu_logits_method_matches_prefill_session():
    """Whisper.logits (all positions) agrees with the last-position logits the decode session produces."""
    import whisper_b200 as wb
    from whisper_b200.decoding import DecodingOptions, DecodingTask

    g, dims, sd, audio, ids, text_tokens, tokens, OA, OM = _setup()
    model = wb.Whisper(wb.ModelDimensions(**dims), sd, device="cuda", dtype=torch.float16)
    feats = model.embed_audio(wb.log_mel_spectrogram(torch.from_numpy(audio).cuda(), dims["n_mels"]))
    task = DecodingTask(model, DecodingOptions(language="en", prompt=text_tokens[:9]))
    sess = task.open_session(1)
    try:
        sess.set_audio(feats)
        sess.prefill(np.asarray([task.initial_tokens], dtype=np.int32))
        last = sess.get_logits(1)[0].clone()
    finally:
        sess.close()
    full = model.logits(torch.tensor([list(task.initial_tokens)]), feats)[0]
    assert torch.equal(full[-1], last)


def test_word_spans_match_the_reference_formula():
    """timing.word_spans against the reference's own arithmetic (whisper/timing.py:227-241, restated here as the checker:
    word_boundaries = pad(cumsum(len(word_tokens[:-1]))), jumps = pad(diff(text_indices), 1) != 0,
    jump_times = time_indices[jumps] / 50, start / end = jump_times[boundaries[:-1] / [1:]]) on random monotone paths."""
    from whisper_b200.timing import word_spans

    rng = np.random.default_rng(5)
    for trial in range(50):
        n_tok = int(rng.integers(2, 40))
        # a monotone DTW path over n_tok + 1 text positions (the last is <|endoftext|>) and ~3 frames per token
        steps = []
        t = f = 0
        path_t, path_f = [0], [0]
        while t < n_tok:
            move = rng.integers(0, 3)
            if move == 0:
                t += 1
                f += 1
            elif move == 1:
                t += 1
            else:
                f += 1
            path_t.append(t)
            path_f.append(f)
        text_indices, time_indices = np.array(path_t), np.array(path_f)
        cuts = sorted(set(rng.integers(1, n_tok, size=int(rng.integers(0, 6))).tolist())) if n_tok > 1 else []
        bounds = [0] + cuts + [n_tok]
        word_tokens = [list(range(a, b)) for a, b in zip(bounds, bounds[1:])] + [[99999]]
        words = [f"w{i}" for i in range(len(word_tokens))]
        probs = rng.random(n_tok)
        got = word_spans(words, word_tokens, text_indices, time_indices, probs)
        word_boundaries = np.pad(np.cumsum([len(t) for t in word_tokens[:-1]]), (1, 0))
        jumps = np.pad(np.diff(text_indices), (1, 0), constant_values=1).astype(bool)
        jump_times = time_indices[jumps] / 50.0
        start, end = jump_times[word_boundaries[:-1]], jump_times[word_boundaries[1:]]
        want_p = [np.mean(probs[i:j]) for i, j in zip(word_boundaries[:-1], word_boundaries[1:])]
        assert len(got) == len(word_tokens) - 1
        for k, w in enumerate(got):
            assert (w.word, w.tokens) == (words[k], word_tokens[k])
            assert w.start == start[k] and w.end == end[k] and abs(w.probability - want_p[k]) < 1e-12
