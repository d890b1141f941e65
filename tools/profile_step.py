#!/usr/bin/env python
"""Workload driver for ncu: the bench.py hot path (large-v3, beam 5, batch 64, bf16) with a short
decode, bracketed by cudaProfilerStart/Stop so `ncu --profile-from-start off` sees exactly one step.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py --decode-steps 6
  ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:cross_attention -c 3 -o gpurun_out/cross_attn python tools/profile_step.py --decode-steps 4
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["WB200_NO_GRAPH"] = "1"          # ncu attributes graph-replayed kernels poorly; use plain launches

import whisper_b200 as wb  # noqa: E402
from whisper_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="large-v3")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--beam", type=int, default=5)
ap.add_argument("--decode-steps", type=int, default=6)
ap.add_argument("--dtype", default="bf16")
args = ap.parse_args()

dims = synthetic.dims_dict(args.model)
dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float16
model = wb.Whisper(wb.ModelDimensions(**dims), synthetic.synthetic_state_dict(dims, seed=0), device="cuda", dtype=dtype)
audio = torch.from_numpy(synthetic.synthetic_audio(args.batch, 480000, seed=1234, kind="noise")).cuda()
tok = wb.tokenizer.get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language="en", task="transcribe")
opt = wb.DecodingOptions(language="en", beam_size=args.beam if args.beam > 1 else None, sample_len=args.decode_steps,
                         suppress_tokens=[-1, tok.eot])


def step():
    mel = wb.log_mel_spectrogram(audio, dims["n_mels"], per_waveform_max=True)
    return model.decode(mel, opt)


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
