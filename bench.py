#!/usr/bin/env python
"""Benchmark of the Whisper hot path on B200:  RTFx (audio-seconds / wall-second).

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference algorithm on the host CPU

One "step" = one pass of the whole hot path over one batch of synthetic 30-second segments:
log-mel -> AudioEncoder -> cross-K/V -> prefill -> (decoder step, logit filters, beam update) x 224
-> finalise / rank.  Default workload = BASELINE.json configs[2]: large-v3, beam 5, batch 64 per GPU,
bf16, kv-cache.  EOT is suppressed so every segment decodes the full 224 tokens (the fixed-length
mode of SURVEY.md 8d; with random weights the natural length would be arbitrary).

Multi-GPU (torchrun, one rank per GPU): replicated weights (rank 0 builds them, NCCL broadcast),
each rank decodes its own 64 segments (weak scaling), results all-gathered at the end of every step.

Prints ONE JSON line on rank 0 (see the keys in `main`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RTFx (audio-s/wall-s) large-v3 beam=5"       # BASELINE.json's metric (the default workload)


def metric_name(args):
    """BASELINE.json's metric string for the default workload; other --model / --beam values are labelled as such."""
    return METRIC if (args.model, args.beam) == ("large-v3", 5) else f"RTFx (audio-s/wall-s) {args.model} beam={args.beam}"
CHUNK_S = 30.0
N_SAMPLES = 480000
DECODE_STEPS = 224


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="large-v3")
    ap.add_argument("--batch", type=int, default=64, help="segments per GPU per step")
    ap.add_argument("--beam", type=int, default=5)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--decode-steps", type=int, default=DECODE_STEPS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true",
                    help="after the timed region, run one extra plain-launch step per kernel class with per-launch "
                         "CUDA events and report each class's total device time (diagnostic, not part of `value`)")
    ap.add_argument("--breakdown-ids", default="1,2,3,4,5,6,7", help="kernel classes for --breakdown")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# workload pieces
# ------------------------------------------------------------------------------------------------
def algorithmic_numbers(dims, B, G, L_avg):
    """SURVEY.md 8(d) / BASELINE.md section 3 figures for the named workload."""
    d, NL, V, T = dims["n_text_state"], dims["n_text_layer"], dims["n_vocab"], dims["n_audio_ctx"]
    R = B * G
    enc_flops = (2 * 3000 * dims["n_mels"] * d * 3 + 2 * 1500 * d * d * 3
                 + dims["n_audio_layer"] * (2 * 1500 * 12 * d * d + 4 * 1500 * 1500 * d))
    return {
        "cross_attn_bytes_per_launch": B * T * 2 * d * 2 + 2 * R * d * 2,          # K+V of one layer + q/out
        "decoder_weight_bytes_per_step": 2 * (NL * 14 * d * d + V * d),
        "cross_kv_bytes_per_step": B * NL * 2 * T * d * 2,
        "self_kv_bytes_per_step_avg": R * NL * 2 * L_avg * d * 2,
        "encoder_flops_per_segment": enc_flops,
    }


def host_threads() -> int:
    """CPU threads this process may really use: the smallest of os.cpu_count(), the scheduler affinity
    mask and the cgroup CPU quota (the GPU boxes expose 128 logical CPUs to containers with far smaller
    quotas; 128 torch threads on such a box run ~20x slower than 16)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                quota = int(txt[0])
                period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if quota > 0:
                    n = min(n, max(1, int(quota / period + 0.5)))
        except Exception:
            pass
    return max(1, min(n, 32))     # torch CPU matmuls stop scaling well before 32 threads on these hosts


def oracle_sample(model_name, beam, n_decode_iters, threads):
    """One bounded CPU sample of the reference algorithm (the oracle port): encoder on ONE segment +
    prefill + a few beam-search steps, extrapolated linearly to the full 224-token window."""
    from oracle import audio as OA
    from oracle import decoding as OD
    from oracle import model as OM
    from whisper_b200 import synthetic

    torch.set_num_threads(threads)
    st = oracle_sample.state
    if st.get("name") != model_name:
        dims = synthetic.dims_dict(model_name)
        st.update(name=model_name, dims=dims, W=OM.to_weights(synthetic.synthetic_state_dict(dims, seed=0)),
                  audio=synthetic.synthetic_audio(1, N_SAMPLES, seed=1234, kind="noise"))
    dims, W = st["dims"], st["W"]
    t0 = time.perf_counter()
    mel = torch.from_numpy(OA.log_mel_spectrogram(st["audio"], dims["n_mels"]))
    t_mel = time.perf_counter() - t0
    t0 = time.perf_counter()
    with torch.no_grad():
        feats = OM.encoder_forward(W, dims, mel)
    t_enc = time.perf_counter() - t0
    ids = OD.token_ids(dims["n_vocab"])
    timings = []
    with torch.no_grad():
        OD.decode(W, dims, feats, OD.Options(beam_size=beam if beam > 1 else None, sample_len=DECODE_STEPS,
                                             suppress_tokens=(-1, ids.eot)), max_steps=1 + n_decode_iters,
                  timings=timings)
    t_prefill = timings[0]
    t_step = float(np.mean(timings[1:])) if len(timings) > 1 else timings[0]
    total = t_mel + t_enc + t_prefill + (DECODE_STEPS - 1) * t_step
    return {"rtfx": CHUNK_S / total, "t_mel": t_mel, "t_enc": t_enc, "t_prefill": t_prefill, "t_step": t_step,
            "wall": t_mel + t_enc + sum(timings)}


oracle_sample.state = {}


def run_reference_arm(args, rank):
    """The reference's algorithm on the host CPU (oracle port; the Python reference itself cannot travel
    to the GPU box).  Rank 0 only."""
    if rank != 0:
        return
    threads = host_threads()
    n_iters = 3
    for _ in range(args.warmup):
        oracle_sample(args.model, args.beam, n_iters, threads)
    t0 = time.perf_counter()
    vals = [oracle_sample(args.model, args.beam, n_iters, threads) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    v = float(np.mean([x["rtfx"] for x in vals]))
    sample = (f"per step: log-mel + encoder on 1 of {args.batch} segments + prefill + {n_iters} beam-{args.beam} decode "
              f"iterations of {DECODE_STEPS}, extrapolated linearly (enc {vals[-1]['t_enc']:.2f}s, prefill "
              f"{vals[-1]['t_prefill']:.2f}s, {vals[-1]['t_step']:.3f}s/iter); fp32, torch CPU threads={threads}")
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args), "value": v, "unit": "x realtime", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} beam={args.beam} batch={args.batch} synthetic 30 s @16 kHz, "
                               f"{DECODE_STEPS} decode steps (EOT suppressed)", "impl_detail": "oracle port on host CPU"},
        "cpu_baseline": {"value": v, "unit": "x realtime", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


NCU_CROSS_ATTN_TRAFFIC = 492.06e6 + 11.95e6   # bytes per launch, C3 shape


def main():
    args = parse()
    from whisper_b200 import parallel

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference_arm(args, rank)
        return

    import torch.distributed as dist

    import whisper_b200 as wb
    from whisper_b200 import _lib, synthetic

    rank, world, local = parallel.init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    dims = synthetic.dims_dict(args.model)
    B, G = args.batch, args.beam

    # ---- weights: rank 0 builds the synthetic checkpoint, NCCL broadcast, every rank packs its replica
    spec = [(n, s) for n, s, _ in synthetic.state_dict_spec(dims)]
    sd = synthetic.synthetic_state_dict(dims, seed=0) if rank == 0 else None
    sd_dev = parallel.broadcast_state_dict(sd, spec, dev)
    model = wb.Whisper(wb.ModelDimensions(**dims), sd_dev, device=dev, dtype=dtype)
    del sd, sd_dev
    torch.cuda.empty_cache()

    # ---- inputs: B x 30 s of synthetic 16 kHz audio per rank (pinned host copy + device copy)
    audio_host = torch.from_numpy(synthetic.synthetic_audio(B, N_SAMPLES, seed=1234 + rank, kind="noise")).pin_memory()
    audio_dev = audio_host.to(dev)
    tok = wb.tokenizer.get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language="en",
                                     task="transcribe")
    options = wb.DecodingOptions(language="en", beam_size=G if G > 1 else None, sample_len=args.decode_steps,
                                 suppress_tokens=[-1, tok.eot])

    def hot_path(audio):
        mel = wb.log_mel_spectrogram(audio, dims["n_mels"], per_waveform_max=True)   # each segment = its own file
        res = model.decode(mel, options)
        toks, lps, nss = parallel.gather_results([r.tokens for r in res], [r.avg_logprob for r in res],
                                                 [r.no_speech_prob for r in res], dev)
        return toks

    def step_resident():
        return hot_path(audio_dev)

    def step_e2e():
        return hot_path(audio_host.to(dev, non_blocking=True))

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(k):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    for _ in range(max(3, args.warmup)):
        out = step_resident()
    n_tokens = [len(t) for t in out]

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib = _lib.lib()
    launches0 = _lib.launch_count()
    ms, out = timed(step_resident, args.steps)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # roofline of the dominant kernel (decoder-step cross-attention): one more step with every launch of
    # that kernel bracketed by CUDA events on its stream.  This pass runs the decode loop as plain launches
    # (the timed steps above replay it from CUDA graphs, where per-launch events cannot be interleaved).
    import ctypes

    lib.wb200_profile_enable(1)
    ms_prof, _ = timed(step_resident, 1)
    prof_ms, prof_n = ctypes.c_double(0), ctypes.c_int64(0)
    lib.wb200_profile_read(ctypes.byref(prof_ms), ctypes.byref(prof_n))
    lib.wb200_profile_enable(0)

    breakdown = None
    if args.breakdown:
        names = {1: "cross_attention", 2: "self_attention", 3: "gemm", 4: "encoder_attention", 5: "layernorm",
                 6: "select", 7: "log_mel"}
        breakdown = {}
        for kid, name in names.items():
            if str(kid) not in args.breakdown_ids.split(","):
                continue
            lib.wb200_profile_enable(kid)
            ms_k, _ = timed(step_resident, 1)
            t_k, n_k = ctypes.c_double(0), ctypes.c_int64(0)
            lib.wb200_profile_read(ctypes.byref(t_k), ctypes.byref(n_k))
            lib.wb200_profile_enable(0)
            breakdown[name] = {"ms": t_k.value, "launches": int(n_k.value), "step_ms": ms_k}

    e2e_steps = max(1, min(args.steps, 3))
    step_e2e()
    ms_e2e, _ = timed(step_e2e, e2e_steps)

    if rank != 0:
        return
    audio_s = world * B * CHUNK_S
    value = audio_s * args.steps / (ms / 1000.0)
    e2e_value = audio_s * e2e_steps / (ms_e2e / 1000.0)
    L_avg = len(tok.sot_sequence) + args.decode_steps / 2
    alg = algorithmic_numbers(dims, B, G, L_avg)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    avg_launch_ms = prof_ms.value / max(1, prof_n.value)
    achieved = alg["cross_attn_bytes_per_launch"] / (avg_launch_ms / 1000.0) / 1e9 if avg_launch_ms > 0 else 0.0
    R = B * G
    ctx = dims["n_text_ctx"]
    d2h = R * (len(tok.sot_sequence) + args.decode_steps) * 4 + R * 4 + B * 4 + 4 + \
        (B * G * ctx * 4 + 3 * B * G * 4 if G > 1 else 0)
    line = {
        "metric": metric_name(args), "value": value, "unit": "x realtime", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {
            "workload": f"{args.model} beam={G} batch={B}/GPU synthetic 30 s @16 kHz, kv-cache, "
                        f"{args.decode_steps} decode steps per segment (EOT suppressed), random-init weights",
            "baseline_config": "BASELINE.json configs[2]" if (args.model, G, B) == ("large-v3", 5, 64) else "custom",
            "parallelism": f"dp{world} (replicated weights, segments sharded, no per-step collective)",
            "l2": "per-step working set (cross-K/V 15.7 GB + self-K/V) far exceeds the 126 MB L2; no flush needed",
            "tokens_per_segment": int(np.mean(n_tokens)),
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "x realtime", "h2d_bytes_per_step": B * N_SAMPLES * 4,
                "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps},
        "gpu_launches": int(launches),
        "roofline": {
            "kernel": "cross_attention_kernel (decoder step, one launch per layer per step)",
            "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
            "frac": achieved / peak_gbs if peak_gbs else None,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
            "algorithmic_bytes_per_launch": alg["cross_attn_bytes_per_launch"],
            "avg_launch_ms": avg_launch_ms, "launches_timed": int(prof_n.value),
            "share_of_step": (prof_ms.value / ms_prof) if ms_prof > 0 else None,
            "profiled_step_ms": ms_prof,
            "how": "CUDA events around every launch of the kernel on its stream, one extra (non-graph) step",
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed `ncu --set full` capture
            # (profiles/r1_dec_attn_v3_ncu_full_selected.csv); only valid for the shape it was captured on
            "traffic": NCU_CROSS_ATTN_TRAFFIC if (args.model, G, B, args.dtype) == ("large-v3", 5, 64, "bf16") else None,
            "traffic_source": "profiles/r1_dec_attn_v3_ncu_full_selected.csv (ncu --set full, one launch)",
        },
        "algorithmic": alg,
    }
    if breakdown is not None:
        line["breakdown"] = breakdown
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        s = oracle_sample(args.model, G, 3, threads)
        line["cpu_baseline"] = {
            "value": s["rtfx"], "unit": "x realtime", "cores": threads, "kind": "port",
            "sample": (f"oracle port (fp32, torch CPU, {threads} threads): log-mel + encoder on 1 segment + prefill + 3 "
                       f"beam-{G} iterations, extrapolated to {DECODE_STEPS} (enc {s['t_enc']:.2f}s, prefill "
                       f"{s['t_prefill']:.2f}s, {s['t_step']:.3f}s/iter; {s['wall']:.1f}s of CPU work)")}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
