"""GPU: the fused decoder-layer kernel (csrc/dec_layer.cu: LayerNorm folded into the Linears, six Linears of a layer
in three persistent launches with grid-wide phase barriers) against the unfused round-1 kernels and the fp32 oracle.

The two device paths compute the same reference math (whisper/model.py:142-171) with different roundings - the fused
path never rounds LN(x) to 16 bits but rounds W (.) gamma once - so they agree to 16-bit activation noise, not bit for
bit; both must sit inside the oracle tolerance of tests/test_model_gpu.py."""
import numpy as np
import pytest
import torch

from helpers import fixture_inputs, load_model_fixture, oracle_features

pytestmark = pytest.mark.gpu

LOGIT_TOL = {torch.float16: 2.5e-3, torch.bfloat16: 2.5e-2}


def _set_fused(on: bool):
    from whisper_b200 import _lib

    assert _lib.lib().wb200_set_fused_decoder_layer(int(on)) == 0


def _set_rows(on: bool):
    from whisper_b200 import _lib

    assert _lib.lib().wb200_set_fused_decoder_rows(int(on)) == 0


def _set_stack(mode):
    """0: per-layer launches, 1 / True: one launch per iteration (default), 2: ... ending with the final LayerNorm + logits."""
    from whisper_b200 import _lib

    assert _lib.lib().wb200_set_fused_decoder_stack(int(mode)) == 0


def _teacher_forced_logits(model, g_feats, rec, n_audio, opts):
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
@pytest.mark.parametrize("name,opts", [("test-en", dict(beam_size=5, sample_len=16)), ("test-multi", dict(sample_len=16)),
                                       ("tiny.en", dict(beam_size=3, sample_len=12))])
def test_fused_layer_matches_unfused_and_oracle(name, opts, dtype):
    import whisper_b200 as wb
    from oracle import parity

    meta, arrays, dims, W, mel, feats = oracle_features(name)
    rec = parity.oracle_record(W, dims, feats, opts, 2)
    _, sd, audio = fixture_inputs(meta)
    model = wb.Whisper(wb.ModelDimensions(**dims), sd, device="cuda", dtype=dtype)
    g_mel = torch.stack([wb.log_mel_spectrogram(torch.from_numpy(a).cuda(), dims["n_mels"]) for a in audio])
    g_feats = model.embed_audio(g_mel)
    try:
        _set_fused(False)
        model.clear_sessions()           # the switch applies to sessions created after it: drop the parked ones
        plain = _teacher_forced_logits(model, g_feats, rec, 2, opts)
        _set_fused(True)
        _set_rows(False)                 # the 64-row tile form, whatever the row count
        model.clear_sessions()
        fused = _teacher_forced_logits(model, g_feats, rec, 2, opts)
    finally:
        _set_fused(True)
        _set_rows(True)
        model.clear_sessions()
    worst_pair, worst_ora = 0.0, 0.0
    for i, (a, b) in enumerate(zip(plain, fused)):
        assert bool(torch.isfinite(b).all()), f"step {i}: non-finite logits from the fused path"
        ref = rec["raw_logits"][i]
        ref = ref[::(opts.get("beam_size") or 1)] if i == 0 else ref
        scale = float(ref.abs().max())
        worst_pair = max(worst_pair, float((a - b).abs().max()) / scale)
        worst_ora = max(worst_ora, float((b - ref).abs().max()) / scale)
    print(f"{name} {dtype}: fused vs unfused {worst_pair:.5f}, fused vs oracle {worst_ora:.5f} (of max |logit|), {len(fused)} steps")
    assert worst_ora < LOGIT_TOL[dtype]
    assert 0.0 < worst_pair < LOGIT_TOL[dtype], "the two paths must differ by rounding only (and must not be the same path)"


def test_fused_layer_is_deterministic():
    """Grid barriers, LN partial merges and the column split must not make results depend on scheduling."""
    import whisper_b200 as wb

    meta, _ = load_model_fixture("test-en")
    dims, sd, audio = fixture_inputs(meta)
    model = wb.Whisper(wb.ModelDimensions(**dims), sd, device="cuda", dtype=torch.float16)
    mel = torch.stack([wb.log_mel_spectrogram(torch.from_numpy(a).cuda(), dims["n_mels"]) for a in audio])
    opt = wb.DecodingOptions(language="en", beam_size=5, sample_len=32)
    a = model.decode(mel, opt)
    for _ in range(3):
        b = model.decode(mel, opt)
        assert [r.tokens for r in a] == [r.tokens for r in b]
        assert [r.avg_logprob for r in a] == [r.avg_logprob for r in b]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_fused_layer_more_than_one_row_block(dtype):
    """R = 14 audios x 5 beams = 70 rows: two 64-row blocks in the narrow phases, one 128-row block (UMMA M = 128) in the
    wide ones (QKV, fc1), LN partials written by one split and read by the other."""
    import whisper_b200 as wb
    from oracle import audio as OA
    from oracle import model as OM
    from oracle import parity
    from whisper_b200 import synthetic

    meta, _ = load_model_fixture("test-multi")
    dims, sd, _ = fixture_inputs(meta)
    n_audio, opts = 14, dict(beam_size=5, sample_len=6)
    audio = synthetic.synthetic_audio(n_audio, 480000, seed=99, kind="speechlike")
    W = OM.to_weights(sd)
    mel = torch.from_numpy(np.stack([OA.log_mel_spectrogram(a, dims["n_mels"]) for a in audio]))
    feats = OM.encoder_forward(W, dims, mel)
    rec = parity.oracle_record(W, dims, feats, opts, n_audio)
    model = wb.Whisper(wb.ModelDimensions(**dims), sd, device="cuda", dtype=dtype)
    g_mel = torch.stack([wb.log_mel_spectrogram(torch.from_numpy(a).cuda(), dims["n_mels"]) for a in audio])
    g_feats = model.embed_audio(g_mel)
    try:
        _set_fused(False)
        model.clear_sessions()
        plain = _teacher_forced_logits(model, g_feats, rec, n_audio, opts)
        _set_fused(True)
        model.clear_sessions()
        fused = _teacher_forced_logits(model, g_feats, rec, n_audio, opts)
    finally:
        _set_fused(True)
        model.clear_sessions()
    worst_pair = worst_ora = 0.0
    for i, (a, b) in enumerate(zip(plain, fused)):
        assert bool(torch.isfinite(b).all())
        ref = rec["raw_logits"][i]
        ref = ref[::5] if i == 0 else ref
        scale = float(ref.abs().max())
        worst_pair = max(worst_pair, float((a - b).abs().max()) / scale)
        worst_ora = max(worst_ora, float((b - ref).abs().max()) / scale)
    print(f"70 rows {dtype}: fused vs unfused {worst_pair:.5f}, fused vs oracle {worst_ora:.5f}")
    assert worst_ora < LOGIT_TOL[dtype] and 0.0 < worst_pair < LOGIT_TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("name,n_audio,opts", [("test-multi", 1, dict(sample_len=16)),          # R = 1
                                               ("test-en", 1, dict(beam_size=5, sample_len=16)),  # R = 5
                                               ("tiny.en", 2, dict(beam_size=3, sample_len=12)),  # R = 6
                                               ("tiny.en", 2, dict(beam_size=4, sample_len=10)),  # R = 8
                                               ("test-multi", 7, dict(sample_len=8)),             # R = 7
                                               ("test-multi", 12, dict(sample_len=6)),            # R = 12: two operand tiles
                                               ("tiny.en", 4, dict(beam_size=5, sample_len=6)),   # R = 20: four
                                               ("test-en", 32, dict(sample_len=6))])              # R = 32
def test_few_rows_form_matches_tile_form_and_oracle(name, n_audio, opts, dtype):
    """Sessions with few rows (<= 8 for the large models, <= 32 where the weight slab and the input rows fit in shared
    memory) run the weight-stationary form of the fused layer (dec_rows_kernel: bulk-copied weight slabs, mma.sync with the
    weight rows as the M operand, K split over eight warps); same math as the tile form up to the order of the fp32 sums."""
    import whisper_b200 as wb
    from oracle import audio as OA
    from oracle import model as OM
    from oracle import parity
    from whisper_b200 import synthetic

    meta, _ = load_model_fixture(name)
    dims, sd, _ = fixture_inputs(meta)
    audio = synthetic.synthetic_audio(n_audio, 480000, seed=7 + n_audio, kind="speechlike")
    W = OM.to_weights(sd)
    mel = torch.from_numpy(np.stack([OA.log_mel_spectrogram(a, dims["n_mels"]) for a in audio]))
    feats = OM.encoder_forward(W, dims, mel)
    rec = parity.oracle_record(W, dims, feats, opts, n_audio)
    G = opts.get("beam_size") or 1
    assert n_audio * G <= 32
    model = wb.Whisper(wb.ModelDimensions(**dims), sd, device="cuda", dtype=dtype)
    g_mel = torch.stack([wb.log_mel_spectrogram(torch.from_numpy(a).cuda(), dims["n_mels"]) for a in audio])
    g_feats = model.embed_audio(g_mel)
    try:
        _set_rows(False)
        model.clear_sessions()
        tile = _teacher_forced_logits(model, g_feats, rec, n_audio, opts)
        _set_rows(True)
        _set_stack(False)                # three few-rows launches per layer around the two attention kernels
        model.clear_sessions()
        rows = _teacher_forced_logits(model, g_feats, rec, n_audio, opts)
        _set_stack(True)                 # the whole stack, attention included, as one launch per iteration
        model.clear_sessions()
        stack = _teacher_forced_logits(model, g_feats, rec, n_audio, opts)
        again = _teacher_forced_logits(model, g_feats, rec, n_audio, opts)
    finally:
        _set_rows(True)
        _set_stack(True)
        model.clear_sessions()
    worst_pair = worst_stack = worst_ora = 0.0
    for i, (a, b, c, e) in enumerate(zip(tile, rows, stack, again)):
        assert bool(torch.isfinite(b).all()) and bool(torch.isfinite(c).all()), f"step {i}: non-finite logits from the few-rows form"
        assert torch.equal(c, e), f"step {i}: the one-launch stack is not deterministic"
        ref = rec["raw_logits"][i]
        ref = ref[::G] if i == 0 else ref
        scale = float(ref.abs().max())
        worst_pair = max(worst_pair, float((a - b).abs().max()) / scale)
        worst_stack = max(worst_stack, float((b - c).abs().max()) / scale)
        worst_ora = max(worst_ora, float((b - ref).abs().max()) / scale, float((c - ref).abs().max()) / scale)
    print(f"{name} R={n_audio * G} {dtype}: few-rows vs tile form {worst_pair:.5f}, one-launch stack vs per-layer launches "
          f"{worst_stack:.5f}, vs oracle {worst_ora:.5f} (of max |logit|), {len(rows)} steps")
    # the forms differ by the order of fp32 sums only (Linears: K split over warps; attention: keys split over warps / slices)
    assert worst_ora < LOGIT_TOL[dtype] and worst_pair < LOGIT_TOL[dtype] and worst_stack < LOGIT_TOL[dtype]


@pytest.mark.parametrize("name,n_audio,opts", [("test-multi", 1, dict(sample_len=12)), ("tiny.en", 1, dict(beam_size=5, sample_len=10))])
def test_stack_closing_with_layernorm_and_logits(name, n_audio, opts):
    """Mode 2 of the one-launch stack: the table ends with the decoder's final LayerNorm (a warp per row) and the logits (a
    Linear whose 1/148 share of the vocabulary passes through the slab buffer in several slabs, fp32 output) - against
    the default (separate layernorm + tcgen05 GEMM launches) and the oracle."""
    import whisper_b200 as wb
    from oracle import audio as OA
    from oracle import model as OM
    from oracle import parity
    from whisper_b200 import synthetic

    dtype = torch.float16
    meta, _ = load_model_fixture(name)
    dims, sd, _ = fixture_inputs(meta)
    audio = synthetic.synthetic_audio(n_audio, 480000, seed=31, kind="speechlike")
    W = OM.to_weights(sd)
    mel = torch.from_numpy(np.stack([OA.log_mel_spectrogram(a, dims["n_mels"]) for a in audio]))
    rec = parity.oracle_record(W, dims, OM.encoder_forward(W, dims, mel), opts, n_audio)
    G = opts.get("beam_size") or 1
    model = wb.Whisper(wb.ModelDimensions(**dims), sd, device="cuda", dtype=dtype)
    g_feats = model.embed_audio(torch.stack([wb.log_mel_spectrogram(torch.from_numpy(a).cuda(), dims["n_mels"]) for a in audio]))
    try:
        _set_stack(1)
        model.clear_sessions()
        base = _teacher_forced_logits(model, g_feats, rec, n_audio, opts)
        _set_stack(2)
        model.clear_sessions()
        full = _teacher_forced_logits(model, g_feats, rec, n_audio, opts)
    finally:
        _set_stack(1)
        model.clear_sessions()
    worst_pair = worst_ora = 0.0
    for i, (a, b) in enumerate(zip(base, full)):
        assert bool(torch.isfinite(b).all())
        ref = rec["raw_logits"][i]
        ref = ref[::G] if i == 0 else ref
        scale = float(ref.abs().max())
        worst_pair = max(worst_pair, float((a - b).abs().max()) / scale)
        worst_ora = max(worst_ora, float((b - ref).abs().max()) / scale)
    print(f"{name} R={n_audio * G}: stack with LayerNorm + logits vs default {worst_pair:.6f}, vs oracle {worst_ora:.5f}")
    assert worst_ora < LOGIT_TOL[dtype] and worst_pair < 1e-3      # same 16-bit inputs, fp32 sums in another order
