"""GPU parity at the dimensions BASELINE.json's configs are quoted on: base.en (configs[1]),
large-v3 (configs[2]) and large-v3-turbo (configs[3]) - d = 512 / 1280, 8 / 20 heads, 6 / 32 / 4 decoder layers,
80 / 128 mels - in fp16 and bf16, two audios, beam 5.

The fp32 oracle runs the same synthetic checkpoint on the host CPU (encoder ~5 s per segment, ~1 s per decode
iteration at large-v3 on the GPU box), so the comparison is bounded to: encoder features of both audios, and the
teacher-forced decoder - prefill + 8 beam-search iterations along the oracle's trajectory with the oracle's logits
injected into the device selection kernels (oracle/parity.py), which checks the device logits of every iteration
(cross-attention over 2 x 5 rows x 20 heads, self-attention through the beam parent table after real reorders, the
R x 51866 fp32 logits GEMM, 32 layers of 16-bit accumulation) and the exact tokens / parents / score sums.

Tolerances: the 16-bit emulation of the oracle (oracle/model.py ACT_DTYPE) puts the logit error of a correct
16-bit pipeline at 2.4e-3 (fp16) / 1.9e-2 (bf16) of the largest |logit| at large-v3, and the feature error at
0.017 / 0.12 max, 0.001 / 0.008 mean; the gates below are twice that.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOGIT_TOL = {torch.float16: 5e-3, torch.bfloat16: 4e-2}
FEAT_MAX = {torch.float16: 0.04, torch.bfloat16: 0.25}
FEAT_MEAN = {torch.float16: 0.0025, torch.bfloat16: 0.017}
N_AUDIO, BEAM, STEPS = 2, 5, 9          # prefill + 8 cached iterations

_STATE = {}


def _state_dict(name):
    """Synthetic checkpoints are seeded per (seed, tensor name): turbo's tensors are exactly the matching subset of
    large-v3's, so one generation (and one oracle encoder pass) serves both."""
    from whisper_b200 import synthetic

    base = "large-v3" if name == "turbo" else name
    key = ("sd", base)
    if key not in _STATE:
        _STATE[key] = synthetic.synthetic_state_dict(synthetic.dims_dict(base), seed=0)
    sd = _STATE[key]
    dims = synthetic.dims_dict(name)
    if name == "turbo":
        keep = {n for n, _, _ in synthetic.state_dict_spec(dims)}
        sd = {k: v for k, v in sd.items() if k in keep}
    return dims, sd


def _oracle(name):
    from oracle import audio as OA
    from oracle import model as OM
    from oracle import parity
    from whisper_b200 import synthetic

    if ("rec", name) in _STATE:
        return _STATE[("rec", name)]
    torch.set_num_threads(min(32, len(__import__("os").sched_getaffinity(0))))
    dims, sd = _state_dict(name)
    W = OM.to_weights(sd)
    enc_key = ("enc", "large-v3" if name == "turbo" else name)
    if enc_key not in _STATE:
        audio = synthetic.synthetic_audio(N_AUDIO, 480000, seed=4321, kind="speechlike")
        mel = torch.from_numpy(np.stack([OA.log_mel_spectrogram(a, dims["n_mels"]) for a in audio]))
        with torch.no_grad():
            _STATE[enc_key] = (audio, OM.encoder_forward(W, dims, mel))
    audio, feats = _STATE[enc_key]
    opts = dict(beam_size=BEAM, sample_len=STEPS)
    with torch.no_grad():
        rec = parity.oracle_record(W, dims, feats, opts, N_AUDIO)
    _STATE[("rec", name)] = (dims, sd, audio, feats, opts, rec)
    return _STATE[("rec", name)]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("name", ["base.en", "large-v3", "turbo"])
def test_baseline_dims_against_oracle(name, dtype):
    import whisper_b200 as wb
    from oracle import parity

    dims, sd, audio, feats, opts, rec = _oracle(name)
    model = wb.Whisper(wb.ModelDimensions(**dims), sd, device="cuda", dtype=dtype)
    try:
        mel = torch.stack([wb.log_mel_spectrogram(torch.from_numpy(a).cuda(), dims["n_mels"]) for a in audio])
        g_feats = model.embed_audio(mel)
        err = (g_feats.float().cpu() - feats).abs()
        print(f"{name} {dtype}: encoder feature err max {float(err.max()):.4f} mean {float(err.mean()):.5f} "
              f"(oracle std {float(feats.std()):.3f})")
        assert float(err.max()) < FEAT_MAX[dtype] and float(err.mean()) < FEAT_MEAN[dtype]
        out = parity.teacher_forced(model, opts, N_AUDIO, g_feats, rec, LOGIT_TOL[dtype])
        print(f"{name} {dtype}: teacher-forced beam-{BEAM}, {out['steps']} iterations, {out['reorders']} non-identity "
              f"reorders, worst |logit err| / max|logit| = {out['worst_rel_logit_err']:.5f}")
        assert out["steps"] >= 6 and out["reorders"] > 0          # the oracle may complete a step or two early
        free = parity.free_running(model, opts, N_AUDIO, g_feats, rec, dims)
        print(f"{name} {dtype}: free-running asserted {free['asserted_steps']} of {free['steps']} iterations "
              f"(first gap {free['first_gap']:.4f}, first bound {free['first_bound']:.4f})")
    finally:
        del model
        torch.cuda.empty_cache()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("name,beam", [("turbo", None), ("large-v3", 5)])
def test_few_rows_form_at_baseline_dims(name, beam, dtype):
    """One audio (the shape of BASELINE configs[3]: turbo, greedy, batch 1 - and large-v3 with 5 beams): <= 8 rows, so the
    step runs the weight-stationary form of the fused layer (dec_rows_kernel) with 8-35 weight rows of K = 1280 / 5120 per
    SM.  Device logits along the oracle's trajectory, and the same run through the 64-row tile form."""
    import whisper_b200 as wb
    from oracle import model as OM
    from oracle import parity
    from whisper_b200 import _lib

    dims, sd, audio, feats, _, _ = _oracle(name)
    opts = dict(sample_len=7) if beam is None else dict(beam_size=beam, sample_len=7)
    G = beam or 1
    with torch.no_grad():
        rec = parity.oracle_record(OM.to_weights(sd), dims, feats[:1], opts, 1)
    model = wb.Whisper(wb.ModelDimensions(**dims), sd, device="cuda", dtype=dtype)

    def run():
        task, sess = parity.open_session(model, opts, 1, g_feats)
        out = []
        try:
            for i in range(len(rec["raw_logits"])):
                if i > 0:
                    sess.step()
                out.append(sess.get_logits(1 if i == 0 else G).float().cpu())
                ref = rec["raw_logits"][i]
                sess.set_logits(ref[::G] if i == 0 else ref)
                sess.select()
        finally:
            sess.close()
        return out

    try:
        mel = wb.log_mel_spectrogram(torch.from_numpy(audio[0]).cuda(), dims["n_mels"])[None]
        g_feats = model.embed_audio(mel)
        stack = run()                                   # default: the whole stack as one launch per iteration
        _lib.lib().wb200_set_fused_decoder_stack(0)
        model.clear_sessions()
        rows = run()                                    # few-rows chains around the separate attention kernels
        _lib.lib().wb200_set_fused_decoder_rows(0)
        model.clear_sessions()
        tile = run()
        worst_pair = worst_stack = worst_ora = 0.0
        for i, (a, b, c) in enumerate(zip(tile, rows, stack)):
            ref = rec["raw_logits"][i]
            ref = ref[::G] if i == 0 else ref
            scale = float(ref.abs().max())
            worst_pair = max(worst_pair, float((a - b).abs().max()) / scale)
            worst_stack = max(worst_stack, float((b - c).abs().max()) / scale)
            worst_ora = max(worst_ora, float((b - ref).abs().max()) / scale, float((c - ref).abs().max()) / scale)
        print(f"{name} beam={beam} {dtype}: few-rows vs tile form {worst_pair:.5f}, one-launch stack vs per-layer launches "
              f"{worst_stack:.5f}, vs oracle {worst_ora:.5f} over {len(rows)} iterations")
        assert worst_ora < LOGIT_TOL[dtype]
        assert 0.0 < worst_pair < LOGIT_TOL[dtype]       # two different kernels, the same math
        assert 0.0 < worst_stack < LOGIT_TOL[dtype]
    finally:
        _lib.lib().wb200_set_fused_decoder_stack(1)
        _lib.lib().wb200_set_fused_decoder_rows(1)
        model.clear_sessions()
        del model
        torch.cuda.empty_cache()
