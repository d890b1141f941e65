"""Long-form transcription driver with the reference's signature (whisper/transcribe.py:38-514).

The host still owns the sequential 30-second window loop - window n+1 starts at the last timestamp
decoded in window n and is prompted with window n's text (transcribe.py:288-295, 369-377) - but every
tensor operation inside it runs in libwhisper_b200.so: one fused log-mel pass over the whole file
(with the file-global dynamic-range clamp, transcribe.py:139), then per window one encoder pass and
one device-resident decode.
"""
from __future__ import annotations

from typing import TYPE_CHECKING, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .audio import FRAMES_PER_SECOND, HOP_LENGTH, N_FRAMES, N_SAMPLES, SAMPLE_RATE, log_mel_spectrogram, pad_or_trim
from .decoding import DecodingOptions, DecodingResult
from .tokenizer import get_tokenizer

if TYPE_CHECKING:
    from .model import Whisper


class _WindowLoop:
    """State of one file's transcription: seek position, accumulated tokens and segments."""

    def __init__(self, model, mel, tokenizer, *, temperature, compression_ratio_threshold, logprob_threshold,
                 no_speech_threshold, condition_on_previous_text, initial_prompt, carry_initial_prompt,
                 clip_timestamps, verbose, decode_options):
        self.model = model
        self.mel = mel
        self.tokenizer = tokenizer
        self.temperatures = [temperature] if isinstance(temperature, (int, float)) else list(temperature)
        self.cr_threshold = compression_ratio_threshold
        self.lp_threshold = logprob_threshold
        self.ns_threshold = no_speech_threshold
        self.condition = condition_on_previous_text
        self.carry_initial_prompt = carry_initial_prompt
        self.verbose = verbose
        self.decode_options = dict(decode_options)
        self.content_frames = mel.shape[-1] - N_FRAMES
        self.input_stride = N_FRAMES // model.dims.n_audio_ctx            # mel frames per output token: 2
        self.time_precision = self.input_stride * HOP_LENGTH / SAMPLE_RATE  # 0.02 s
        if isinstance(clip_timestamps, str):
            clip_timestamps = [float(ts) for ts in (clip_timestamps.split(",") if clip_timestamps else [])]
        points = [round(ts * FRAMES_PER_SECOND) for ts in clip_timestamps]
        if len(points) == 0:
            points.append(0)
        if len(points) % 2 == 1:
            points.append(self.content_frames)
        self.clips: List[Tuple[int, int]] = list(zip(points[::2], points[1::2]))
        self.all_tokens: List[int] = []
        self.all_segments: List[dict] = []
        self.prompt_reset_since = 0
        self.remaining_prompt_length = model.dims.n_text_ctx // 2 - 1
        if initial_prompt is not None:
            self.initial_prompt_tokens = tokenizer.encode(" " + initial_prompt.strip())
            self.all_tokens.extend(self.initial_prompt_tokens)
            self.remaining_prompt_length -= len(self.initial_prompt_tokens)
        else:
            self.initial_prompt_tokens = []

    # temperature ladder (transcribe.py:184-224), written as a generator: it YIELDS (segment, DecodingOptions)
    # requests and is sent the DecodingResult, so one file can be driven by model.decode directly (run) or many
    # files can be advanced in lock-step with their requests batched (transcribe_batch)
    def fallback_steps(self, segment: torch.Tensor):
        result = None
        for t in self.temperatures:
            kwargs = {**self.decode_options}
            if t > 0:
                kwargs.pop("beam_size", None)
                kwargs.pop("patience", None)
            else:
                kwargs.pop("best_of", None)
            result = yield segment, DecodingOptions(**kwargs, temperature=t)
            needs_fallback = False
            if self.cr_threshold is not None and result.compression_ratio > self.cr_threshold:
                needs_fallback = True
            if self.lp_threshold is not None and result.avg_logprob < self.lp_threshold:
                needs_fallback = True
            if (self.ns_threshold is not None and result.no_speech_prob > self.ns_threshold
                    and self.lp_threshold is not None and result.avg_logprob < self.lp_threshold):
                needs_fallback = False
            if not needs_fallback:
                break
        return result

    def drive(self, gen):
        """Run a request generator to completion against self.model.decode; returns the generator's value."""
        try:
            req = next(gen)
            while True:
                req = gen.send(self.model.decode(req[0], req[1]))
        except StopIteration as stop:
            return stop.value

    def decode_with_fallback(self, segment: torch.Tensor) -> DecodingResult:
        return self.drive(self.fallback_steps(segment))

    def make_segment(self, seek, start, end, tokens: List[int], result: DecodingResult) -> dict:
        text_tokens = [t for t in tokens if t < self.tokenizer.eot]
        return {"seek": seek, "start": start, "end": end, "text": self.tokenizer.decode(text_tokens),
                "tokens": tokens, "temperature": result.temperature, "avg_logprob": result.avg_logprob,
                "compression_ratio": result.compression_ratio, "no_speech_prob": result.no_speech_prob}

    def split_window(self, seek, segment_size, segment_duration, result):
        """Cut one window's tokens into segments at consecutive timestamp pairs and decide how far to
        advance (transcribe.py:339-399).  Returns (segments, new_seek)."""
        tb = self.tokenizer.timestamp_begin
        tokens = list(result.tokens)
        is_ts = [t >= tb for t in tokens]
        time_offset = float(seek * HOP_LENGTH / SAMPLE_RATE)
        single_timestamp_ending = is_ts[-2:] == [False, True]
        consecutive = [i + 1 for i in range(len(tokens) - 1) if is_ts[i] and is_ts[i + 1]]
        segments = []
        if consecutive:
            slices = list(consecutive)
            if single_timestamp_ending:
                slices.append(len(tokens))
            last = 0
            for cur in slices:
                piece = tokens[last:cur]
                segments.append(self.make_segment(seek, time_offset + (piece[0] - tb) * self.time_precision,
                                                  time_offset + (piece[-1] - tb) * self.time_precision, piece, result))
                last = cur
            if single_timestamp_ending:
                seek += segment_size
            else:
                seek += (tokens[last - 1] - tb) * self.input_stride
        else:
            duration = segment_duration
            stamps = [t for t in tokens if t >= tb]
            if stamps and stamps[-1] != tb:
                duration = (stamps[-1] - tb) * self.time_precision
            segments.append(self.make_segment(seek, time_offset, time_offset + duration, tokens, result))
            seek += segment_size
        return segments, seek

    def run(self):
        self.drive(self.steps())
        return self

    def steps(self):
        """The window loop of transcribe.py:272-508 as a request generator (see fallback_steps)."""
        clip_idx = 0
        seek = self.clips[clip_idx][0]
        while clip_idx < len(self.clips):
            clip_start, clip_end = self.clips[clip_idx]
            if seek < clip_start:
                seek = clip_start
            if seek >= clip_end:
                clip_idx += 1
                if clip_idx < len(self.clips):
                    seek = self.clips[clip_idx][0]
                continue
            segment_size = min(N_FRAMES, self.content_frames - seek, clip_end - seek)
            mel_segment = pad_or_trim(self.mel[:, seek: seek + segment_size], N_FRAMES)
            segment_duration = segment_size * HOP_LENGTH / SAMPLE_RATE
            if self.carry_initial_prompt:
                nignored = max(len(self.initial_prompt_tokens), self.prompt_reset_since)
                remaining = self.all_tokens[nignored:][-self.remaining_prompt_length:]
                self.decode_options["prompt"] = self.initial_prompt_tokens + remaining
            else:
                self.decode_options["prompt"] = self.all_tokens[self.prompt_reset_since:]
            result = yield from self.fallback_steps(mel_segment)
            if self.ns_threshold is not None:
                should_skip = result.no_speech_prob > self.ns_threshold
                if self.lp_threshold is not None and result.avg_logprob > self.lp_threshold:
                    should_skip = False
                if should_skip:
                    seek += segment_size
                    continue
            segments, seek = self.split_window(seek, segment_size, segment_duration, result)
            if self.verbose:
                for s in segments:
                    print(f"[{s['start']:.3f} --> {s['end']:.3f}] {s['text']}")
            for s in segments:                                        # transcribe.py:484-489
                if s["start"] == s["end"] or s["text"].strip() == "":
                    s["text"] = ""
                    s["tokens"] = []
            self.all_segments.extend({"id": i, **s} for i, s in enumerate(segments, start=len(self.all_segments)))
            self.all_tokens.extend(t for s in segments for t in s["tokens"])
            if not self.condition or result.temperature > 0.5:
                self.prompt_reset_since = len(self.all_tokens)


def transcribe(
    model: "Whisper",
    audio: Union[str, np.ndarray, torch.Tensor],
    *,
    verbose: Optional[bool] = None,
    temperature: Union[float, Tuple[float, ...]] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0),
    compression_ratio_threshold: Optional[float] = 2.4,
    logprob_threshold: Optional[float] = -1.0,
    no_speech_threshold: Optional[float] = 0.6,
    condition_on_previous_text: bool = True,
    initial_prompt: Optional[str] = None,
    carry_initial_prompt: bool = False,
    word_timestamps: bool = False,
    prepend_punctuations: str = "\"'“¿([{-",
    append_punctuations: str = "\"'.。,，!！?？:：”)]}、",
    clip_timestamps: Union[str, List[float]] = "0",
    hallucination_silence_threshold: Optional[float] = None,
    **decode_options,
):
    """Transcribe an audio file / waveform; returns {"text", "segments", "language"} like the reference
    (transcribe.py:38-126 documents every parameter).

    Differences from the reference, all explicit: the model always computes in its own 16-bit type
    (`fp16=` is accepted and ignored); rungs of the `temperature` ladder above 0 sample with the library's
    counter-based generator (seeded from torch's global generator, so `torch.manual_seed` makes a run
    repeatable, but the draws are not the reference's); `word_timestamps=True` raises.
    """
    loop, tokenizer, language = _prepare(model, audio, verbose=verbose, temperature=temperature,
                                         compression_ratio_threshold=compression_ratio_threshold,
                                         logprob_threshold=logprob_threshold, no_speech_threshold=no_speech_threshold,
                                         condition_on_previous_text=condition_on_previous_text,
                                         initial_prompt=initial_prompt, carry_initial_prompt=carry_initial_prompt,
                                         word_timestamps=word_timestamps, clip_timestamps=clip_timestamps,
                                         decode_options=decode_options)
    loop.run()
    return _result(loop, tokenizer, language)


def _prepare(model, audio, *, verbose, temperature, compression_ratio_threshold, logprob_threshold, no_speech_threshold,
             condition_on_previous_text, initial_prompt, carry_initial_prompt, word_timestamps, clip_timestamps,
             decode_options):
    """Everything transcribe() does before its window loop (transcribe.py:128-183): log-mel of the whole file,
    language detection, tokenizer, loop state."""
    if word_timestamps:
        raise NotImplementedError(
            "word_timestamps=True is not wired into transcribe(): the tensor part of word timing is available as "
            "whisper_b200.timing.find_alignment (cross-attention export + median filter + DTW on the GPU), but the "
            "punctuation-merging / segment-clamping text heuristics of timing.py:245-388 are out of scope (SURVEY.md 2)")
    decode_options = dict(decode_options)
    decode_options.pop("fp16", None)
    temps = [temperature] if isinstance(temperature, (int, float)) else list(temperature)

    mel = log_mel_spectrogram(audio, model.dims.n_mels, padding=N_SAMPLES, device=model.device)   # transcribe.py:139
    if decode_options.get("language", None) is None:
        if not model.is_multilingual:
            decode_options["language"] = "en"
        else:
            if verbose:
                print("Detecting language using up to the first 30 seconds. Use `--language` to specify the language")
            _, probs = model.detect_language(pad_or_trim(mel, N_FRAMES))
            decode_options["language"] = max(probs, key=probs.get)
    language = decode_options["language"]
    task = decode_options.get("task", "transcribe")
    tokenizer = get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language=language, task=task)
    loop = _WindowLoop(model, mel, tokenizer, temperature=temps,
                       compression_ratio_threshold=compression_ratio_threshold,
                       logprob_threshold=logprob_threshold, no_speech_threshold=no_speech_threshold,
                       condition_on_previous_text=condition_on_previous_text, initial_prompt=initial_prompt,
                       carry_initial_prompt=carry_initial_prompt, clip_timestamps=clip_timestamps, verbose=verbose,
                       decode_options=decode_options)
    return loop, tokenizer, language


def _result(loop: _WindowLoop, tokenizer, language: str) -> dict:
    return dict(text=tokenizer.decode(loop.all_tokens[len(loop.initial_prompt_tokens):]),
                segments=loop.all_segments, language=language)


def transcribe_batch(
    model: "Whisper",
    audios: Sequence[Union[str, np.ndarray, torch.Tensor]],
    *,
    max_batch: int = 64,
    verbose: Optional[bool] = None,
    temperature: Union[float, Tuple[float, ...]] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0),
    compression_ratio_threshold: Optional[float] = 2.4,
    logprob_threshold: Optional[float] = -1.0,
    no_speech_threshold: Optional[float] = 0.6,
    condition_on_previous_text: bool = True,
    initial_prompt: Optional[str] = None,
    carry_initial_prompt: bool = False,
    clip_timestamps: Union[str, List[float]] = "0",
    **decode_options,
) -> List[dict]:
    """transcribe() for MANY files at once.  One file's window loop is sequential by construction -
    window n+1 starts at the last timestamp decoded in window n and is prompted with its text (transcribe.py:288-295,
    369-377) - so a single long file keeps the GPU at one decoder row.  Here every file keeps its own loop state and
    the loops advance in LOCK-STEP: each round gathers the next decode request of every unfinished file (a window,
    or the next rung of that file's temperature ladder), decoding.decode_requests() batches the requests that can
    share a session (same options, prompts of the same length - in steady state every prompt has the maximum
    223 tokens), and each result is handed back to its file.  Per-file results are those of transcribe(); the order
    of `audios` is kept.  The parameters are transcribe()'s and apply to every file."""
    from .decoding import decode_requests

    prepared = [_prepare(model, a, verbose=verbose, temperature=temperature,
                         compression_ratio_threshold=compression_ratio_threshold, logprob_threshold=logprob_threshold,
                         no_speech_threshold=no_speech_threshold, condition_on_previous_text=condition_on_previous_text,
                         initial_prompt=initial_prompt, carry_initial_prompt=carry_initial_prompt, word_timestamps=False,
                         clip_timestamps=clip_timestamps, decode_options=decode_options) for a in audios]
    gens = [loop.steps() for loop, _, _ in prepared]
    pending = {}
    for i, g in enumerate(gens):
        try:
            pending[i] = next(g)
        except StopIteration:
            pass
    rounds = 0
    while pending:
        order = sorted(pending)
        results = decode_requests(model, [pending[i] for i in order], max_batch=max_batch)
        rounds += 1
        for i, r in zip(order, results):
            try:
                pending[i] = gens[i].send(r)
            except StopIteration:
                del pending[i]
    out = [_result(loop, tok, lang) for loop, tok, lang in prepared]
    for o in out:
        o["rounds"] = rounds            # diagnostic: lock-step rounds the whole batch needed
    return out
