"""One audio, turbo, greedy: the one-launch decoder stack without (mode 1) and with (mode 2) the closing LayerNorm + logits
phases - same tokens?  time per decoder iteration?   python tools/time_stack_modes.py [--model turbo] [--steps 224]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import whisper_b200 as wb  # noqa: E402
from whisper_b200 import _lib, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="turbo")
ap.add_argument("--steps", type=int, default=224)
ap.add_argument("--batch", type=int, default=1)
args = ap.parse_args()
dims = synthetic.dims_dict(args.model)
model = wb.Whisper(wb.ModelDimensions(**dims), synthetic.synthetic_state_dict(dims, seed=0), device="cuda", dtype=torch.float16)
audio = torch.from_numpy(synthetic.synthetic_audio(args.batch, 480000, seed=1234, kind="speechlike")).cuda()
mel = wb.log_mel_spectrogram(audio, dims["n_mels"])
feats = model.embed_audio(mel)
tok = wb.tokenizer.get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language="en", task="transcribe")
opt = wb.DecodingOptions(language="en", sample_len=args.steps, suppress_tokens=[-1, tok.eot])
out = {}
for mode in (1, 2, 1, 2):
    _lib.lib().wb200_set_fused_decoder_stack(mode)
    model.clear_sessions()
    res = model.decode(feats, opt)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count()
    t0.record()
    for _ in range(3):
        res = model.decode(feats, opt)
    t1.record()
    torch.cuda.synchronize()
    us = t0.elapsed_time(t1) * 1000.0 / 3 / args.steps
    print(f"mode {mode}: {us:.1f} us per iteration (decode of {args.steps} tokens incl. prefill), {(_lib.launch_count() - n0) // 3} launches per decode, "
          f"avg_logprob {res[0].avg_logprob:.6f}")
    out[mode] = res
_lib.lib().wb200_set_fused_decoder_stack(1)
a, b = out[1], out[2]
for i in range(args.batch):
    same = a[i].tokens == b[i].tokens
    first = next((k for k, (x, y) in enumerate(zip(a[i].tokens, b[i].tokens)) if x != y), None)
    print(f"audio {i}: tokens identical {same} (first difference {first}), avg_logprob {a[i].avg_logprob:.6f} vs {b[i].avg_logprob:.6f}")
