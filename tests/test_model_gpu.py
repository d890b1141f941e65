"""GPU: the full hot path (encoder -> decoder session -> selection) through the C ABI, against
the reference's golden results (tests/golden, made by oracle/make_golden.py) and the oracle.

Parity protocol (SURVEY.md section 7 "hard parts", 8c):
 * integer work - logit filters, top-k, greedy / beam bookkeeping, kv-cache parent table - is checked
   BIT-EXACTLY by feeding the oracle's own fp32 logits into the device selection kernels step by
   step (test_selection_kernels_exact);
 * floating-point work - encoder features, decoder logits - is checked to a stated tolerance against
   the fp32 oracle under teacher forcing (test_decoder_logits_teacher_forced);
 * free-running decodes must reproduce the reference's token ids exactly whenever every decision
   the oracle took had a margin larger than the 16-bit noise bound TAU, and otherwise up to the first
   such low-margin decision (test_decode_end_to_end).
"""
import numpy as np
import pytest
import torch

from helpers import fixture_inputs, load_model_fixture, oracle_features, oracle_options

pytestmark = pytest.mark.gpu

DTYPES = [torch.float16, torch.bfloat16]
# max |logit error| allowed under teacher forcing, and the decision-margin gate for e2e equality.
# logits of the synthetic models have std ~8-40; 16-bit activations give ~2^-11 (fp16) / 2^-8 (bf16)
# relative error per layer.
# LOGIT_TOL is relative to the largest |logit| of the step (the synthetic models have heavy-tailed logits)
LOGIT_TOL = {torch.float16: 2.5e-3, torch.bfloat16: 2.5e-2}
# TAU (absolute logit units): a decision is only asserted where the oracle's margin exceeds twice the worst logit error a
# correct 16-bit pipeline shows under teacher forcing - 1.5e-3 (fp16) / 1.5e-2 (bf16) of the largest |logit| (~40 for the
# synthetic models), measured across the kernel forms (tile / few-rows / one-launch stack differ in summation order only)
TAU = {torch.float16: 0.12, torch.bfloat16: 1.5}
# ... and never less than twice the measured relative error times the model's own logit scale (tiny.en: max |logit| 258,
# measured per-step error up to 0.5 in fp16 and 3.0 in bf16 under teacher forcing, profiles/r2_summary.md section 6)
REL_ERR = {torch.float16: 2.0e-3, torch.bfloat16: 1.3e-2}


def decision_gate(rec, dtype):
    scale = max(float(l.abs().max()) for l in rec["raw_logits"])
    return max(TAU[dtype], 2.0 * REL_ERR[dtype] * scale)
FEAT_TOL = {torch.float16: 0.02, torch.bfloat16: 0.12}

_MODELS = {}


def gpu_model(name, dtype):
    key = (name, dtype)
    if key not in _MODELS:
        import whisper_b200 as wb

        meta, _ = load_model_fixture(name)
        dims, sd, audio = fixture_inputs(meta)
        _MODELS[key] = wb.Whisper(wb.ModelDimensions(**dims), sd, device="cuda", dtype=dtype)
    return _MODELS[key]


def gpu_mel(name):
    import whisper_b200 as wb

    meta, _ = load_model_fixture(name)
    dims, sd, audio = fixture_inputs(meta)
    return torch.stack([wb.log_mel_spectrogram(torch.from_numpy(a).cuda(), dims["n_mels"]) for a in audio])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["test-en", "test-multi", "tiny.en"])
def test_encoder_features(name, dtype):
    meta, arrays = load_model_fixture(name)
    model = gpu_model(name, dtype)
    feats = model.embed_audio(gpu_mel(name)).float().cpu().numpy()
    ref = arrays["feats_sub"]
    err = np.abs(feats[:, ::25] - ref)
    print(f"{name} {dtype}: feature max err {err.max():.4f} mean err {err.mean():.5f} (ref std {ref.std():.3f})")
    assert err.max() < FEAT_TOL[dtype] * 4 and err.mean() < FEAT_TOL[dtype] / 4


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["test-en", "test-multi"])
def test_decoder_logits_teacher_forced(name, dtype):
    """Prefill + 24 cached steps fed with the ORACLE's tokens: logits within LOGIT_TOL of the fp32
    oracle at every step, arg-max identical wherever the oracle's margin exceeds TAU."""
    from oracle import decoding as OD
    from whisper_b200.decoding import DecodingOptions, DecodingTask

    meta, arrays, dims, W, mel, feats = oracle_features(name)
    rec = {}
    OD.decode(W, dims, feats, OD.Options(sample_len=25), record=rec)
    model = gpu_model(name, dtype)
    g_feats = model.embed_audio(gpu_mel(name))
    task = DecodingTask(model, DecodingOptions(language="en", sample_len=25))
    sess = task.open_session(2)
    try:
        sess.set_audio(g_feats)
        sess.prefill(np.tile(np.asarray(task.initial_tokens, dtype=np.int32), (2, 1)))
        worst = 0.0
        for i in range(len(rec["raw_logits"])):
            got = sess.get_logits(2).float().cpu()
            ref = rec["raw_logits"][i]
            assert bool(torch.isfinite(got).all()), f"step {i}: non-finite logits"
            err = float((got - ref).abs().max() / ref.abs().max())
            worst = max(worst, err)
            top2 = ref.topk(2, dim=-1)
            margin = top2.values[:, 0] - top2.values[:, 1]
            same = got.argmax(-1) == top2.indices[:, 0]
            assert bool((same | (margin < TAU[dtype])).all()), f"step {i}: argmax differs with margin {margin.tolist()}"
            if i + 1 < len(rec["raw_logits"]):
                nxt = [t[-1] for t in rec["tokens_out"][i]]
                sess.force_tokens(nxt)
                sess.step()
        print(f"{name} {dtype}: teacher-forced max |logit err| / max|logit| = {worst:.5f} "
              f"(logit std {float(ref.std()):.2f}, max {float(ref.abs().max()):.1f})")
        assert worst < LOGIT_TOL[dtype]
        ns = sess.get("no_speech").cpu().numpy()
        ref_ns = np.array([r["no_speech_prob"] for r in meta["decode"]["greedy"]["results"]])
        assert np.allclose(ns, ref_ns, rtol=0.2, atol=1e-12)
    finally:
        sess.close()


CASES = ["greedy", "greedy_notimestamps", "greedy_prompt", "greedy_prefix", "greedy_nosuppress", "beam5",
         "beam5_patience2", "beam3_lenpen", "beam2_notimestamps"]


def _cases_of(name):
    meta, _ = load_model_fixture(name)
    return [c for c in CASES if c in meta["decode"]]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,case", [(n, c) for n in ("test-en", "test-multi", "test-peak") for c in _cases_of(n)])
def test_selection_kernels_exact(name, case, dtype):
    """Feed the oracle's fp32 logits to the device filters / top-k / greedy / beam kernels at every step: chosen
    tokens, beam parents, completion and finished hypotheses must be IDENTICAL; the fp32 log-probability sums agree
    to 1e-4 (different reduction order in logsumexp).  Before the oracle's logits overwrite them, the logits the
    DEVICE computed for the step are compared with the oracle's (LOGIT_TOL): because the device is forced along the
    oracle's trajectory this checks the kv-cache parent-table indirection under REAL beam reorders
    (reference decoding.py:172-176), which greedy decoding never exercises."""
    from oracle import parity

    meta, arrays, dims, W, mel, feats = oracle_features(name)
    c = meta["decode"][case]
    n_audio = c["n_audio"]
    rec = parity.oracle_record(W, dims, feats, c["options"], n_audio)
    model = gpu_model(name, dtype)
    g_feats = model.embed_audio(gpu_mel(name))
    out = parity.teacher_forced(model, c["options"], n_audio, g_feats, rec, LOGIT_TOL[dtype])
    print(f"{name}/{case}/{dtype}: {out['steps']} steps, {out['reorders']} with a non-identity beam reorder, "
          f"worst |logit err| / max|logit| = {out['worst_rel_logit_err']:.5f}")
    if c["options"].get("beam_size"):
        assert out["reorders"] > 0, "the case never reorders beams: the parent table was not exercised"
    assert out["done"] == int(out["steps"] < (c["options"].get("sample_len") or 224))     # completion flag


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,case", [(n, c) for n in ("test-en", "test-multi", "test-peak", "tiny.en") for c in _cases_of(n)
                                       if c.startswith("beam") or n == "test-peak"])
def test_free_running_stepwise(name, case, dtype):
    """FREE-RUNNING decode on the device's own logits, asserted step by step (tokens, beam parents) for as long as
    the oracle's decision gap exceeds the error bound measured on the spot (oracle/parity.py).  On the `test-peak`
    fixture (wide candidate gaps by construction) at least the first steps MUST be asserted; the other fixtures report
    how many steps their 16-bit noise allows."""
    from oracle import parity

    meta, arrays, dims, W, mel, feats = oracle_features(name)
    c = meta["decode"][case]
    n_audio = c["n_audio"]
    rec = parity.oracle_record(W, dims, feats, c["options"], n_audio)
    model = gpu_model(name, dtype)
    g_feats = model.embed_audio(gpu_mel(name))
    out = parity.free_running(model, c["options"], n_audio, g_feats, rec, dims)
    print(f"{name}/{case}/{dtype}: free-running asserted {out['asserted_steps']} of {out['steps']} steps "
          f"(first gap {out['first_gap']:.4f}, first bound {out['first_bound']:.4f})")
    if name == "test-peak":
        assert out["asserted_steps"] >= 1, out


SAMPLING_CASES = [dict(temperature=0.7, best_of=3, seed=11), dict(temperature=1.0, seed=(1 << 40) + 5),
                  dict(temperature=0.3, best_of=2, seed=3, without_timestamps=True)]


@pytest.mark.parametrize("opts", SAMPLING_CASES)
@pytest.mark.parametrize("name", ["test-en", "test-multi"])
def test_sampling_kernel_matches_contract(name, opts):
    """GreedyDecoder with a temperature (decoding.py:283): feed the oracle's fp32 logits to the device filter /
    Gumbel-max kernel step by step.  The drawn tokens must equal the oracle's restatement of the RNG contract
    (Philox4x32-10 counters, include/whisper_b200.h) wherever the oracle's best and second-best perturbed
    scores are further apart than fp32 log round-off, and the accumulated UN-tempered log-probabilities agree."""
    from oracle import decoding as OD
    from whisper_b200.decoding import DecodingOptions, DecodingTask

    meta, arrays, dims, W, mel, feats = oracle_features(name)
    n_audio = 2
    rec = {}
    OD.decode(W, dims, feats[:n_audio], OD.Options(sample_len=12, **opts), record=rec)
    model = gpu_model(name, torch.float16)
    g_feats = model.embed_audio(gpu_mel(name))[:n_audio].contiguous()
    task = DecodingTask(model, DecodingOptions(language="en", sample_len=12, **opts))
    G = task.n_group
    assert G == (opts.get("best_of") or 1)
    sess = task.open_session(n_audio)
    try:
        sess.set_audio(g_feats)
        sess.set_sampling(opts["temperature"], opts["seed"])
        sess.prefill(np.tile(np.asarray(task.initial_tokens, dtype=np.int32), (n_audio, 1)))
        checked = 0
        for i in range(len(rec["raw_logits"])):
            logits = rec["raw_logits"][i]
            if i == 0:
                sess.set_logits(logits[::G])
            else:
                sess.step()
                sess.set_logits(logits)
            sess.select()
            if min(rec["sample_gaps"][i]) < 1e-4:      # a near-tie: fp32 log() round-off may flip it; stop comparing
                break
            L = int(sess.get("length").item())
            toks = sess.get("tokens")[:, :L].cpu().numpy().tolist()
            assert toks == rec["tokens_out"][i], f"step {i}: sampled tokens differ from the contract"
            lp = sess.get("sum_logprobs").cpu()
            assert torch.allclose(lp, rec["sum_logprobs_out"][i], atol=1e-4, rtol=1e-5), f"step {i}: sum_logprobs differ"
            checked += 1
        assert checked >= 3
    finally:
        sess.close()


def test_sampling_distribution_on_device():
    """Chi-square of the device sampler on a 5-token distribution: 1024 rows x 6 steps of injected logits."""
    from whisper_b200.decoding import DecodingOptions, DecodingTask

    model = gpu_model("test-en", torch.float16)
    n_audio, G = 64, 16
    task = DecodingTask(model, DecodingOptions(language="en", temperature=0.8, best_of=G, without_timestamps=True,
                                               suppress_tokens="", suppress_blank=False, sample_len=8))
    V = model.dims.n_vocab
    live = [101, 2049, 2050, 7000, V - 1]
    vals = torch.tensor([0.3, 1.7, -0.5, 2.2, 0.9])
    p = torch.softmax(vals / 0.8, 0).numpy()
    logits = torch.full((n_audio * G, V), float("-inf"))
    logits[:, live] = vals
    feats = model.embed_audio(gpu_mel("test-en"))[:1].expand(n_audio, -1, -1).contiguous()
    sess = task.open_session(n_audio)
    try:
        sess.set_audio(feats)
        sess.set_sampling(0.8, 2024)
        sess.prefill(np.tile(np.asarray(task.initial_tokens, dtype=np.int32), (n_audio, 1)))
        counts = np.zeros(len(live))
        n = 0
        for i in range(6):
            if i == 0:
                sess.set_logits(logits[::G])
            else:
                sess.step()
                sess.set_logits(logits)
            sess.select()
            L = int(sess.get("length").item())
            drawn = sess.get("tokens")[:, L - 1].cpu().numpy()
            assert set(drawn.tolist()) <= set(live), "a filtered (-inf) token was drawn"
            for k, t in enumerate(live):
                counts[k] += int((drawn == t).sum())
            n += len(drawn)
        chi2 = float((((counts - n * p) ** 2) / (n * p)).sum())
        assert chi2 < 18.5, (chi2, counts, n * p)          # 4 dof, p ~ 1e-3
        lp = sess.get("sum_logprobs").cpu().numpy()
        assert np.isfinite(lp).all() and (lp < 0).all()
    finally:
        sess.close()


def test_decode_with_temperature_end_to_end():
    """decode() with temperature > 0 / best_of (decoding.py:524-526, 283): repeatable for a fixed seed (also
    through torch.manual_seed), with finite negative avg_logprob; the transcribe() ladder reaches the sampled rungs."""
    import whisper_b200 as wb

    model = gpu_model("test-en", torch.float16)
    mel = gpu_mel("test-en")[:2]
    o = dict(language="en", temperature=0.9, best_of=3, sample_len=16)
    a = model.decode(mel, wb.DecodingOptions(seed=7, **o))
    b = model.decode(mel, wb.DecodingOptions(seed=7, **o))
    assert [r.tokens for r in a] == [r.tokens for r in b]
    assert all(r.temperature == 0.9 and np.isfinite(r.avg_logprob) and r.avg_logprob < 0 for r in a)
    torch.manual_seed(99)
    d = model.decode(mel, wb.DecodingOptions(**o))
    torch.manual_seed(99)
    e = model.decode(mel, wb.DecodingOptions(**o))
    assert [r.tokens for r in d] == [r.tokens for r in e]
    greedy = model.decode(mel, wb.DecodingOptions(language="en", sample_len=16))
    assert all(r.temperature == 0.0 for r in greedy)
    # the fallback ladder of transcribe() reaches the sampling rungs on noise-like input without raising
    audio = np.random.RandomState(0).randn(16000 * 3).astype(np.float32) * 0.1
    out = model.transcribe(audio, temperature=(0.0, 0.4, 0.8), sample_len=8, compression_ratio_threshold=0.01)
    assert "segments" in out


def _first_risky_step(margins, tau):
    for i, m in enumerate(margins):
        if m < tau:
            return i
    return None


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CASES + ["greedy_full"])
@pytest.mark.parametrize("name", ["test-en", "test-multi", "tiny.en"])
def test_decode_end_to_end(name, case, dtype):
    """model.decode() on the GPU vs the reference's decode() results stored in the golden files."""
    from oracle import decoding as OD
    from whisper_b200.decoding import DecodingOptions

    meta, arrays = load_model_fixture(name)
    if case not in meta["decode"]:
        pytest.skip("case not generated for this model")
    c = meta["decode"][case]
    n_audio = c["n_audio"]
    model = gpu_model(name, dtype)
    mel = gpu_mel(name)[:n_audio]
    got = model.decode(mel, DecodingOptions(language="en", **c["options"]))
    # decision margins from the oracle (cheap for the test models; tiny.en takes a few seconds)
    _, _, dims, W, _, feats = oracle_features(name)
    rec = {}
    o_res = OD.decode(W, dims, feats[:n_audio], oracle_options(c["options"]), record=rec)
    beam = c["options"].get("beam_size")
    for a, (g, ref) in enumerate(zip(got, c["results"])):
        assert o_res[a].tokens == ref["tokens"]                      # the oracle itself is pinned
        assert np.isfinite(g.avg_logprob), f"audio {a}: non-finite avg_logprob, tokens {g.tokens[:8]}"
        if beam:
            risky = rec["beam_min_gap"] < decision_gate(rec, dtype)
            if not risky:
                assert g.tokens == ref["tokens"], f"audio {a}: beam tokens differ although min gap {rec['beam_min_gap']:.3f}"
            elif g.tokens != ref["tokens"]:
                print(f"{name}/{case}/{dtype} audio {a}: beam result differs, oracle min gap {rec['beam_min_gap']:.4f} < TAU")
        else:
            margins = o_res[a].step_margins
            k = _first_risky_step(margins, decision_gate(rec, dtype))
            if k is None:
                assert g.tokens == ref["tokens"], f"audio {a}: tokens differ with all margins >= TAU"
                assert abs(g.avg_logprob - ref["avg_logprob"]) < 0.02 * max(1.0, abs(ref["avg_logprob"]))
            else:
                assert g.tokens[:k] == ref["tokens"][:k], f"audio {a}: prefix before first low-margin step {k} differs"
                if g.tokens != ref["tokens"]:
                    print(f"{name}/{case}/{dtype} audio {a}: diverged after step {k} (margin {margins[k]:.4f} < TAU)")
        assert abs(g.no_speech_prob - ref["no_speech_prob"]) <= 0.25 * ref["no_speech_prob"] + 1e-12


def test_batched_beam_equals_per_audio():
    """The reference cannot run beam search on a batch (decoding.py:734,740); our batched result must
    equal running each audio alone.  Tokens exactly; the scores to 16-bit noise only - the number of rows selects the
    form of the fused decoder layer (few-rows / 64-row tiles), which differ in the order of their fp32 sums."""
    from whisper_b200.decoding import DecodingOptions

    model = gpu_model("test-en", torch.float16)
    mel = gpu_mel("test-en")
    opt = DecodingOptions(language="en", beam_size=5, sample_len=40)
    both = model.decode(mel, opt)
    for a in range(2):
        alone = model.decode(mel[a], opt)
        assert alone.tokens == both[a].tokens and abs(alone.avg_logprob - both[a].avg_logprob) < 2e-3


def test_checkpoint_file_round_trip(tmp_path):
    """A checkpoint file in the reference's format ({"dims", "model_state_dict"}, whisper/__init__.py:147-156) through
    load_model(path): same decode, bit for bit, as the model built from the same state dict in memory - the fp32 -> 16-bit
    repack (pack_weights, LayerNorm-folded decoder weights included) sees the same values either way."""
    import whisper_b200 as wb
    from whisper_b200.decoding import DecodingOptions

    meta, _ = load_model_fixture("test-en")
    dims, sd, _ = fixture_inputs(meta)
    path = tmp_path / "test-en.pt"
    torch.save({"dims": dict(dims), "model_state_dict": {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}}, path)
    loaded = wb.load_model(str(path), dtype=torch.float16)
    assert loaded.dims == wb.ModelDimensions(**dims) and not loaded.is_multilingual
    built = gpu_model("test-en", torch.float16)
    mel = gpu_mel("test-en")
    for opt in (DecodingOptions(language="en", sample_len=24), DecodingOptions(language="en", beam_size=5, sample_len=24)):
        a, b = loaded.decode(mel, opt), built.decode(mel, opt)
        assert [r.tokens for r in a] == [r.tokens for r in b]
        assert [r.avg_logprob for r in a] == [r.avg_logprob for r in b]
    assert torch.equal(loaded.embed_audio(mel), built.embed_audio(mel))


def test_detect_language():
    meta, arrays = load_model_fixture("test-multi")
    model = gpu_model("test-multi", torch.float16)
    feats = model.embed_audio(gpu_mel("test-multi"))
    toks, probs = model.detect_language(feats)
    assert toks.cpu().tolist() == meta["detect_language"]["tokens"]
    assert [max(p, key=p.get) for p in probs] == meta["detect_language"]["top"]
    assert np.allclose([max(p.values()) for p in probs], meta["detect_language"]["top_prob"], atol=0.03)


def test_transcribe_runs_windows():
    """transcribe() over 70 s of synthetic audio: windows advance, segments carry tokens, and the
    first window equals decode() of the first 30 s under the file-global mel clamp."""
    import whisper_b200 as wb
    from whisper_b200 import synthetic

    model = gpu_model("test-en", torch.float16)
    audio = synthetic.synthetic_audio(1, 70 * 16000, seed=77, kind="speechlike")[0]
    out = model.transcribe(audio, temperature=0.0, condition_on_previous_text=True, sample_len=32,
                           no_speech_threshold=None, logprob_threshold=None, compression_ratio_threshold=None)
    assert out["language"] == "en" and len(out["segments"]) >= 2
    seeks = [s["seek"] for s in out["segments"]]
    assert seeks == sorted(seeks) and seeks[0] == 0 and seeks[-1] > 0
    mel = wb.log_mel_spectrogram(audio, 80, padding=480000)
    first = model.decode(wb.pad_or_trim(mel[:, :3000], 3000), wb.DecodingOptions(language="en", sample_len=32))
    first_tokens = [t for s in out["segments"] if s["seek"] == 0 for t in s["tokens"]]
    assert first_tokens == first.tokens[: len(first_tokens)] or len(first_tokens) == 0
