"""Segment decoding with the reference's surface (whisper/decoding.py): `DecodingOptions`,
`DecodingResult`, `decode`, `detect_language`.

The reference drives one PyTorch forward + several host syncs per generated token from Python
(decoding.py:680-710) and does beam search in Python dicts (decoding.py:323-382).  Here the whole
loop lives on the GPU inside a decoder session of libwhisper_b200.so (kv-cache, logit filters,
log-softmax, top-k, greedy / beam bookkeeping); the host configures the session, launches it and
reads the final token rows back once.  Option handling, initial-token construction, hypothesis
finalisation and ranking stay on the host and follow the reference line by line in behaviour.
"""
from __future__ import annotations

import ctypes
import zlib
from ctypes import POINTER, Structure, c_int, c_int32, c_size_t, c_void_p
from dataclasses import dataclass, field, replace
from typing import TYPE_CHECKING, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from ._lib import WhisperB200Error, check, lib, ptr, stream_ptr
from .audio import CHUNK_LENGTH
from .tokenizer import Tokenizer, get_tokenizer

if TYPE_CHECKING:
    from .model import Whisper


def compression_ratio(text) -> float:
    """Reference whisper/utils.py:45-47."""
    text_bytes = text.encode("utf-8")
    return len(text_bytes) / len(zlib.compress(text_bytes))


@dataclass(frozen=True)
class DecodingOptions:
    """Reference decoding.py:80-114 (same fields, same defaults)."""
    task: str = "transcribe"
    language: Optional[str] = None
    temperature: float = 0.0
    sample_len: Optional[int] = None
    best_of: Optional[int] = None
    beam_size: Optional[int] = None
    patience: Optional[float] = None
    length_penalty: Optional[float] = None
    prompt: Optional[Union[str, List[int]]] = None
    prefix: Optional[Union[str, List[int]]] = None
    suppress_tokens: Optional[Union[str, Iterable[int]]] = "-1"
    suppress_blank: bool = True
    without_timestamps: bool = False
    max_initial_timestamp: Optional[float] = 1.0
    fp16: bool = True     # kept for signature compatibility; the compute type is the model's
    # extension: seed of the counter-based sampler used when temperature > 0 (include/whisper_b200.h,
    # wb200_decoder_set_sampling).  None draws one from torch's global generator, so torch.manual_seed()
    # makes sampled decodes repeatable the way it does for the reference (decoding.py:283).
    seed: Optional[int] = None


@dataclass(frozen=True)
class DecodingResult:
    """Reference decoding.py:117-127."""
    audio_features: torch.Tensor
    language: str
    language_probs: Optional[Dict[str, float]] = None
    tokens: List[int] = field(default_factory=list)
    text: str = ""
    avg_logprob: float = np.nan
    no_speech_prob: float = np.nan
    temperature: float = np.nan
    compression_ratio: float = np.nan


# ------------------------------------------------------------------------------------------------
# ctypes view of wb200_decode_config (include/whisper_b200.h)
# ------------------------------------------------------------------------------------------------
class _DecodeConfig(Structure):
    _fields_ = [(n, c_int32) for n in (
        "n_audio", "n_group", "beam_search", "max_candidates", "n_init", "sample_begin", "sot_index",
        "eot", "no_speech", "no_timestamps", "timestamp_begin", "suppress_blank", "timestamp_rules",
        "max_initial_timestamp_index", "n_suppress", "n_blank", "all_logits")] + [
        ("suppress_ids", POINTER(c_int32)), ("blank_ids", POINTER(c_int32))]


STATE = dict(tokens=0, length=1, sum_logprobs=2, no_speech=3, logits=4, top_val=5, top_idx=6, sources=7,
             fin_tokens=8, fin_len=9, fin_score=10, fin_count=11, done=12)


class DecoderSession:
    """One device-resident decode (kv-cache + selection state) for a fixed (n_audio, n_group, n_init)."""

    def __init__(self, model: "Whisper", cfg: dict, suppress: Sequence[int], blank: Sequence[int]):
        self.model = model
        self.cfg = dict(cfg)
        self._sup = (c_int32 * max(1, len(suppress)))(*suppress)
        self._blank = (c_int32 * max(1, len(blank)))(*blank)
        cfg = dict(cfg)
        cfg.setdefault("all_logits", 0)
        c = _DecodeConfig(**cfg, n_suppress=len(suppress), n_blank=len(blank))
        c.suppress_ids = ctypes.cast(self._sup, POINTER(c_int32))
        c.blank_ids = ctypes.cast(self._blank, POINTER(c_int32))
        self._c = c
        self._cache_key = None
        self._dirty = False          # sampling / alignment settings a reuse has to clear
        self.R = cfg["n_audio"] * cfg["n_group"]
        self.K = cfg["n_group"] + 1 if cfg["beam_search"] else 1
        self.ctx = model.dims.n_text_ctx
        with torch.cuda.device(model.device):
            nbytes = int(lib().wb200_decoder_workspace_bytes(model._handle, ctypes.byref(c)))
            self.workspace = torch.empty(nbytes, device=model.device, dtype=torch.uint8)
            h = c_void_p(0)
            check(lib().wb200_decoder_create(model._handle, ctypes.byref(c), ptr(self.workspace), c_size_t(nbytes),
                                             ctypes.byref(h), stream_ptr()), "wb200_decoder_create")
        self._h = h

    def destroy(self):
        if self._h:
            lib().wb200_decoder_destroy(self._h)
            self._h = c_void_p(0)
        self.workspace = None

    def close(self):
        """Give the session back.  Sessions opened through DecodingTask.open_session are parked on the model and reused by
        the next decode() with the same shape and filters (kv arenas, tensor maps and the instantiated CUDA graph of the
        decode loop survive; wb200_decoder_prefill resets all per-decode state), others are destroyed."""
        if self._h and self._cache_key is not None and self.model._park_session(self):
            return
        self.destroy()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def _call(self, name, *args):
        with torch.cuda.device(self.model.device):
            check(getattr(lib(), name)(self._h, *args, stream_ptr()), name)

    def set_audio(self, features: torch.Tensor):
        assert features.is_cuda and features.dtype == self.model.dtype and features.is_contiguous()
        self._features = features
        self._call("wb200_decoder_set_audio", ptr(features))

    def prefill(self, initial_tokens: np.ndarray):
        toks = np.ascontiguousarray(initial_tokens, dtype=np.int32)
        assert toks.shape == (self.cfg["n_audio"], self.cfg["n_init"])
        self._init_host = toks
        self._call("wb200_decoder_prefill", toks.ctypes.data_as(POINTER(c_int32)))

    def set_sampling(self, temperature: float, seed: int):
        """GreedyDecoder temperature sampling (decoding.py:283) with the library's counter-based generator."""
        from ctypes import c_float, c_uint64
        self._dirty = True
        with torch.cuda.device(self.model.device):
            check(lib().wb200_decoder_set_sampling(self._h, c_float(float(temperature)), c_uint64(int(seed) & (2 ** 64 - 1))),
                  "wb200_decoder_set_sampling")

    def select(self):
        self._call("wb200_decoder_select")

    def step(self):
        self._call("wb200_decoder_step")

    def run(self, max_steps: int) -> int:
        n = c_int32(0)
        self._call("wb200_decoder_run", c_int(max_steps), ctypes.byref(n))
        return int(n.value)

    def set_alignment(self, heads: Sequence[Tuple[int, int]]) -> torch.Tensor:
        """Ask the next prefill to export the pre-softmax cross-attention scores of `heads`
        ((layer, head) pairs) for audio 0; returns the fp32 tensor [n_heads, n_init, n_audio_ctx]
        that will receive them (reference timing.py:185-197)."""
        n = len(heads)
        self._dirty = True
        flat = np.ascontiguousarray(np.asarray(heads, dtype=np.int32).reshape(-1))
        self._align_heads = flat
        self._align_qk = torch.empty((n, self.cfg["n_init"], self.model.dims.n_audio_ctx), device=self.model.device,
                                     dtype=torch.float32)
        with torch.cuda.device(self.model.device):
            check(lib().wb200_decoder_set_alignment(self._h, flat.ctypes.data_as(POINTER(c_int32)), c_int(n),
                                                    ptr(self._align_qk)), "wb200_decoder_set_alignment")
        return self._align_qk

    def reset_for_reuse(self):
        if self._dirty:
            from ctypes import c_float, c_uint64
            with torch.cuda.device(self.model.device):
                check(lib().wb200_decoder_set_sampling(self._h, c_float(0.0), c_uint64(0)), "wb200_decoder_set_sampling")
                check(lib().wb200_decoder_set_alignment(self._h, None, c_int(0), None), "wb200_decoder_set_alignment")
            self._dirty = False

    def force_tokens(self, next_tokens: Sequence[int]):
        arr = np.ascontiguousarray(next_tokens, dtype=np.int32)
        assert arr.shape == (self.R,)
        self._forced = arr
        self._call("wb200_decoder_force_tokens", arr.ctypes.data_as(POINTER(c_int32)))

    def logits_ld(self) -> int:
        return int(lib().wb200_decoder_logits_ld(self._h))

    def get(self, what: str) -> torch.Tensor:
        """Copy a piece of session state into a fresh device tensor."""
        B, G = self.cfg["n_audio"], self.cfg["n_group"]
        mc = max(1, self.cfg["max_candidates"])
        f32 = what in ("sum_logprobs", "no_speech", "logits", "top_val", "fin_score")
        shapes = dict(tokens=(self.R, self.ctx), length=(1,), sum_logprobs=(self.R,), no_speech=(B,),
                      top_val=(self.R, self.K), top_idx=(self.R, self.K), sources=(self.R,),
                      fin_tokens=(B, mc, self.ctx), fin_len=(B, mc), fin_score=(B, mc), fin_count=(B,), done=(1,))
        if what == "logits":
            raise ValueError("use get_logits()")
        out = torch.empty(shapes[what], device=self.model.device, dtype=torch.float32 if f32 else torch.int32)
        self._call("wb200_decoder_get_state", c_int(STATE[what]), ptr(out), c_size_t(out.numel() * 4))
        return out

    def get_logits(self, rows: int) -> torch.Tensor:
        """Current logits [rows, n_vocab] (rows = n_audio right after prefill, R afterwards)."""
        ld = self.logits_ld()
        buf = torch.empty((rows, ld), device=self.model.device, dtype=torch.float32)
        self._call("wb200_decoder_get_state", c_int(STATE["logits"]), ptr(buf), c_size_t(buf.numel() * 4))
        return buf[:, : self.model.dims.n_vocab]

    def set(self, what: str, value: torch.Tensor):
        value = value.contiguous()
        self._call("wb200_decoder_set_state", c_int(STATE[what]), ptr(value), c_size_t(value.numel() * value.element_size()))

    def set_logits(self, logits: torch.Tensor):
        """Overwrite the current logits rows with fp32 [rows, n_vocab] (parity tests)."""
        ld = self.logits_ld()
        buf = torch.zeros((logits.shape[0], ld), device=self.model.device, dtype=torch.float32)
        buf[:, : logits.shape[1]] = logits.to(self.model.device, torch.float32)
        self._call("wb200_decoder_set_state", c_int(STATE["logits"]), ptr(buf), c_size_t(buf.numel() * 4))


# ------------------------------------------------------------------------------------------------
# language detection (reference decoding.py:19-77)
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def detect_language(model: "Whisper", mel: torch.Tensor, tokenizer: Tokenizer = None):
    if tokenizer is None:
        tokenizer = get_tokenizer(model.is_multilingual, num_languages=model.num_languages)
    if tokenizer.language is None or tokenizer.language_token not in tokenizer.sot_sequence:
        raise ValueError("This model doesn't have language tokens so it can't perform lang id")
    single = mel.ndim == 2
    if single:
        mel = mel.unsqueeze(0)
    if tuple(mel.shape[-2:]) != (model.dims.n_audio_ctx, model.dims.n_audio_state):
        mel = model.encoder(mel)
    feats = mel.to(model.dtype).contiguous()
    n_audio = feats.shape[0]
    cfg = dict(n_audio=n_audio, n_group=1, beam_search=0, max_candidates=1, n_init=1, sample_begin=1, sot_index=0,
               eot=tokenizer.eot, no_speech=-1, no_timestamps=tokenizer.no_timestamps,
               timestamp_begin=tokenizer.timestamp_begin, suppress_blank=0, timestamp_rules=0,
               max_initial_timestamp_index=-1)
    sess = DecoderSession(model, cfg, (), ())
    try:
        sess.set_audio(feats)
        sess.prefill(np.full((n_audio, 1), tokenizer.sot, dtype=np.int32))     # decoding.py:56-57
        logits = sess.get_logits(n_audio)      # a view of [n_audio, ld] rows
    finally:
        sess.close()
    # mask everything but the language tokens, argmax, softmax (decoding.py:60-66) in one kernel: the language ids are
    # the contiguous range right after <|startoftranscript|> (tokenizer.py:340-355)
    lang_ids = list(tokenizer.all_language_tokens)
    first, n_lang = lang_ids[0], len(lang_ids)
    assert lang_ids == list(range(first, first + n_lang))
    probs = torch.empty((n_audio, n_lang), device=logits.device, dtype=torch.float32)
    language_tokens = torch.empty((n_audio,), device=logits.device, dtype=torch.int32)
    with torch.cuda.device(model.device):
        check(lib().wb200_range_softmax(ptr(logits), ctypes.c_int64(logits.stride(0)), c_int(first), c_int(n_lang),
                                        c_int(n_audio), ptr(probs), ptr(language_tokens), None, None, stream_ptr()),
              "wb200_range_softmax")
    language_tokens = language_tokens.long()
    language_token_probs = probs.cpu()
    language_probs = [
        {c: language_token_probs[i, j].item() for j, c in enumerate(tokenizer.all_language_codes)}
        for i in range(n_audio)]
    if single:
        language_tokens = language_tokens[0]
        language_probs = language_probs[0]
    return language_tokens, language_probs


# ------------------------------------------------------------------------------------------------
# the decoding task (reference decoding.py:508-789)
# ------------------------------------------------------------------------------------------------
class DecodingTask:
    def __init__(self, model: "Whisper", options: DecodingOptions):
        self.model = model
        language = options.language or "en"
        tokenizer = get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language=language,
                                  task=options.task)
        self.tokenizer = tokenizer
        self.options = self._verify_options(options)
        self.n_group: int = options.beam_size or options.best_of or 1
        self.n_ctx: int = model.dims.n_text_ctx
        self.sample_len: int = options.sample_len or model.dims.n_text_ctx // 2
        self.sot_sequence = tokenizer.sot_sequence
        if options.without_timestamps:
            self.sot_sequence = tokenizer.sot_sequence_including_notimestamps
        self.initial_tokens = self._get_initial_tokens()
        self.sample_begin = len(self.initial_tokens)
        self.sot_index = self.initial_tokens.index(tokenizer.sot)
        self.suppress = self._get_suppress_tokens() if options.suppress_tokens else ()
        self.max_initial_timestamp_index = -1
        if not options.without_timestamps and options.max_initial_timestamp:
            precision = CHUNK_LENGTH / model.dims.n_audio_ctx                   # decoding.py:560
            self.max_initial_timestamp_index = round(options.max_initial_timestamp / precision)

    def _verify_options(self, options: DecodingOptions) -> DecodingOptions:
        """decoding.py:572-585, plus the parts of the surface this build does not cover yet."""
        if options.beam_size is not None and options.best_of is not None:
            raise ValueError("beam_size and best_of can't be given together")
        if options.temperature == 0:
            if options.best_of is not None:
                raise ValueError("best_of with greedy sampling (T=0) is not compatible")
        if options.patience is not None and options.beam_size is None:
            raise ValueError("patience requires beam_size to be given")
        if options.length_penalty is not None and not (0 <= options.length_penalty <= 1):
            raise ValueError("length_penalty (alpha) should be a value between 0 and 1")
        if options.temperature < 0:
            raise ValueError("temperature must be >= 0")
        if options.beam_size is not None and options.beam_size > 16:
            raise ValueError("beam_size > 16 is not supported by the device beam kernel")
        return options

    def _get_initial_tokens(self) -> Tuple[int, ...]:
        """decoding.py:587-613."""
        tokens = list(self.sot_sequence)
        if prefix := self.options.prefix:
            prefix_tokens = (self.tokenizer.encode(" " + prefix.strip()) if isinstance(prefix, str) else list(prefix))
            if self.sample_len is not None:
                max_prefix_len = self.n_ctx // 2 - self.sample_len
                prefix_tokens = prefix_tokens[-max_prefix_len:]
            tokens = tokens + prefix_tokens
        if prompt := self.options.prompt:
            prompt_tokens = (self.tokenizer.encode(" " + prompt.strip()) if isinstance(prompt, str) else list(prompt))
            tokens = [self.tokenizer.sot_prev] + prompt_tokens[-(self.n_ctx // 2 - 1):] + tokens
        return tuple(tokens)

    def _get_suppress_tokens(self) -> Tuple[int, ...]:
        """decoding.py:615-642."""
        suppress_tokens = self.options.suppress_tokens
        if isinstance(suppress_tokens, str):
            suppress_tokens = [int(t) for t in suppress_tokens.split(",")]
        else:
            suppress_tokens = list(suppress_tokens)
        if -1 in suppress_tokens:
            suppress_tokens = [t for t in suppress_tokens if t >= 0]
            suppress_tokens.extend(self.tokenizer.non_speech_tokens)
        tk = self.tokenizer
        suppress_tokens.extend([tk.transcribe, tk.translate, tk.sot, tk.sot_prev, tk.sot_lm])
        if tk.no_speech is not None:
            suppress_tokens.append(tk.no_speech)
        return tuple(sorted(set(suppress_tokens)))

    def _get_audio_features(self, mel: torch.Tensor) -> torch.Tensor:
        """decoding.py:644-664: accept pre-encoded features."""
        if tuple(mel.shape[-2:]) == (self.model.dims.n_audio_ctx, self.model.dims.n_audio_state):
            return mel.to(device=self.model.device, dtype=self.model.dtype).contiguous()
        timing = getattr(self.model, "timing", None)
        if timing is None:
            return self.model.encoder(mel)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = self.model.encoder(mel)
        e1.record()
        timing.setdefault("marks", []).extend([("encoder", e0), ("encoder_end", e1)])
        return out

    def session_config(self, n_audio: int) -> dict:
        tk, o = self.tokenizer, self.options
        beam = o.beam_size is not None
        return dict(
            n_audio=n_audio, n_group=self.n_group, beam_search=int(beam),
            max_candidates=round(o.beam_size * (o.patience or 1.0)) if beam else 1,   # decoding.py:312-313
            n_init=len(self.initial_tokens), sample_begin=self.sample_begin, sot_index=self.sot_index,
            eot=tk.eot, no_speech=tk.no_speech if tk.no_speech is not None else -1,
            no_timestamps=tk.no_timestamps, timestamp_begin=tk.timestamp_begin,
            suppress_blank=int(o.suppress_blank), timestamp_rules=int(not o.without_timestamps),
            max_initial_timestamp_index=self.max_initial_timestamp_index)

    def open_session(self, n_audio: int) -> DecoderSession:
        cfg = self.session_config(n_audio)
        if cfg["beam_search"] and cfg["max_candidates"] <= 0:
            raise AssertionError(f"Invalid beam size ({self.options.beam_size}) or patience ({self.options.patience})")
        key = (tuple(sorted(cfg.items())), tuple(self.suppress), tuple(self.tokenizer.blank_tokens))
        sess = self.model._take_session(key)
        if sess is None:
            sess = DecoderSession(self.model, cfg, self.suppress, self.tokenizer.blank_tokens)
            sess._cache_key = key
        return sess

    @torch.no_grad()
    def run(self, mel: torch.Tensor, initial_tokens: Optional[np.ndarray] = None) -> List[DecodingResult]:
        """DecodingTask.run (decoding.py:713-789).  `initial_tokens` (int [n_audio, n_init], optional) gives every
        audio its OWN prompt of the common length n_init - what the lock-step multi-file scheduler
        (transcribe.transcribe_batch) needs; the reference tiles one prompt over the batch (decoding.py:734)."""
        tokenizer = self.tokenizer
        n_audio = mel.shape[0]
        audio_features = self._get_audio_features(mel)
        if initial_tokens is None:
            init = np.tile(np.asarray(self.initial_tokens, dtype=np.int32), (n_audio, 1))
        else:
            init = np.ascontiguousarray(initial_tokens, dtype=np.int32).copy()
            if init.shape != (n_audio, len(self.initial_tokens)):
                raise ValueError(f"initial_tokens must be [{n_audio}, {len(self.initial_tokens)}], got {init.shape}")
            if not (init[:, self.sot_index] == tokenizer.sot).all():
                raise ValueError("every row of initial_tokens needs <|startoftranscript|> at the task's sot_index")

        # language detection overwrites the language token (decoding.py:666-678)
        languages = [self.options.language] * n_audio
        language_probs = None
        if self.options.language is None or self.options.task == "lang_id":
            lang_tokens, language_probs = self.model.detect_language(audio_features, tokenizer)
            languages = [max(p, key=p.get) for p in language_probs]
            if self.options.language is None:
                init[:, self.sot_index + 1] = lang_tokens.cpu().numpy()
        if self.options.task == "lang_id":
            return [DecodingResult(audio_features=f, language=l, language_probs=p)
                    for f, l, p in zip(audio_features, languages, language_probs)]

        seed = None
        if self.options.temperature > 0 and self.options.beam_size is None:   # beam search ignores the temperature (decoding.py:548-552)
            seed = self.options.seed
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        tokens, sum_logprobs, no_speech, finished = self._run_session(audio_features, init, seed)

        G = self.n_group
        tokens = tokens.reshape(n_audio, G, -1)
        sum_logprobs = sum_logprobs.reshape(n_audio, G)
        candidates, cand_logprobs = self._finalize(tokens, sum_logprobs, finished)
        eot = tokenizer.eot
        candidates = [[s[self.sample_begin: s.index(eot)] for s in group] for group in candidates]  # decoding.py:749-752
        selected = self._rank(candidates, cand_logprobs)
        out_tokens = [group[i] for i, group in zip(selected, candidates)]
        texts = [tokenizer.decode(t).strip() for t in out_tokens]
        sel_logprobs = [lp[i] for i, lp in zip(selected, cand_logprobs)]
        avg_logprobs = [lp / (len(t) + 1) for t, lp in zip(out_tokens, sel_logprobs)]            # decoding.py:760-762
        fields = (texts, languages, out_tokens, audio_features, avg_logprobs, no_speech)
        if len(set(map(len, fields))) != 1:
            raise RuntimeError(f"inconsistent result lengths: {list(map(len, fields))}")
        return [DecodingResult(audio_features=f, language=l, tokens=t, text=x, avg_logprob=a, no_speech_prob=n,
                               temperature=self.options.temperature, compression_ratio=compression_ratio(x))
                for x, l, t, f, a, n in zip(*fields)]

    def _run_session(self, audio_features: torch.Tensor, init: np.ndarray, seed: Optional[int]):
        """One device-resident decode of `init.shape[0]` audios on the current stream: (tokens [R, L], sum_logprobs [R],
        no_speech [n_audio], finished-hypothesis store or None).  If `model.timing` is a dict, CUDA events around the
        cross-K/V build, the prefill and the decode loop are appended to it (bench.py's per-phase rooflines)."""
        n_audio = init.shape[0]
        sess = self.open_session(n_audio)
        timing = getattr(self.model, "timing", None)

        def mark(name):
            if timing is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                timing.setdefault("marks", []).append((name, ev))

        try:
            mark("cross_kv")
            sess.set_audio(audio_features)
            if seed is not None:
                sess.set_sampling(self.options.temperature, seed)
            mark("prefill")
            sess.prefill(init)                       # i == 0 forward + no_speech probabilities
            sess.select()                            # filters + first update
            mark("decode_loop")
            steps = 1
            if self.sample_len > 1:
                steps += sess.run(self.sample_len - 1)        # i = 1 .. sample_len-1, stops on completion
            mark("end")
            if timing is not None:
                timing.setdefault("loop_steps", []).append(steps - 1)
            length = int(sess.get("length").item())
            tokens = sess.get("tokens")[:, :length].cpu().numpy()
            sum_logprobs = sess.get("sum_logprobs").cpu().numpy()
            no_speech = (sess.get("no_speech").cpu().tolist() if self.tokenizer.no_speech is not None
                         else [np.nan] * n_audio)
            finished = None
            if self.options.beam_size is not None:
                finished = (sess.get("fin_tokens").cpu().numpy(), sess.get("fin_len").cpu().numpy(),
                            sess.get("fin_score").cpu().numpy(), sess.get("fin_count").cpu().numpy())
        finally:
            sess.close()
        return tokens, sum_logprobs, no_speech, finished

    def _finalize(self, tokens: np.ndarray, sum_logprobs: np.ndarray, finished):
        """GreedyDecoder.finalize (decoding.py:295-298) / BeamSearchDecoder.finalize (:384-404)."""
        eot = self.tokenizer.eot
        n_audio, G, _ = tokens.shape
        if finished is None:
            return ([[row.tolist() + [eot] for row in tokens[a]] for a in range(n_audio)],
                    [[float(v) for v in sum_logprobs[a]] for a in range(n_audio)])
        fin_tokens, fin_len, fin_score, fin_count = finished
        cands, scores = [], []
        for a in range(n_audio):
            seqs: Dict[tuple, float] = {}
            for k in range(int(fin_count[a])):
                seqs[tuple(fin_tokens[a, k, : fin_len[a, k]].tolist())] = float(fin_score[a, k])
            if len(seqs) < self.options.beam_size:       # not enough finished: top up with live beams
                for j in list(np.argsort(sum_logprobs[a]))[::-1]:
                    seqs[tuple(tokens[a, j].tolist() + [eot])] = float(sum_logprobs[a][j])
                    if len(seqs) >= self.options.beam_size:
                        break
            cands.append([list(s) for s in seqs.keys()])
            scores.append(list(seqs.values()))
        return cands, scores

    def _rank(self, candidates, sum_logprobs) -> List[int]:
        """MaximumLikelihoodRanker.rank (decoding.py:199-213)."""
        alpha = self.options.length_penalty
        picks = []
        for group, lps in zip(candidates, sum_logprobs):
            scored = []
            for seq, lp in zip(group, lps):
                n = len(seq)
                scored.append(lp / (n if alpha is None else ((5 + n) / 6) ** alpha))
            picks.append(int(np.argmax(scored)))
        return picks


@torch.no_grad()
def decode(model: "Whisper", mel: torch.Tensor, options: DecodingOptions = DecodingOptions(),
           **kwargs) -> Union[DecodingResult, List[DecodingResult]]:
    """Decode 30-second segment(s) given as mel spectrogram(s) (n_mels, 3000) / (B, n_mels, 3000) or as
    already-encoded audio features (reference decoding.py:793-826).  Unlike the reference
    (decoding.py:734,740), beam search works for batches: each audio runs its own beam search."""
    if single := mel.ndim == 2:
        mel = mel.unsqueeze(0)
    if kwargs:
        options = replace(options, **kwargs)
    result = DecodingTask(model, options).run(mel)
    return result[0] if single else result


@torch.no_grad()
def decode_requests(model: "Whisper", requests: Sequence[Tuple[torch.Tensor, DecodingOptions]],
                    max_batch: int = 64) -> List[DecodingResult]:
    """Decode many (segment, options) requests whose options may differ (different prompts, temperatures ...) in as
    few device sessions as possible: requests that agree on everything except the CONTENT of equally long
    prompts / prefixes share one batched session, each row prefilled with its own prompt.  Results come back in
    request order.  This is the batching primitive of transcribe_batch; a single request is
    exactly decode(model, segment, options)."""
    tasks = [DecodingTask(model, opt) for _, opt in requests]
    groups: Dict[tuple, List[int]] = {}
    for i, (task, (_, opt)) in enumerate(zip(tasks, requests)):
        key = (repr(replace(opt, prompt=None, prefix=None)), len(task.initial_tokens), task.sot_index,
               task.sample_begin)
        groups.setdefault(key, []).append(i)
    out: List[Optional[DecodingResult]] = [None] * len(requests)
    for idxs in groups.values():
        for lo in range(0, len(idxs), max_batch):
            chunk = idxs[lo: lo + max_batch]
            mel = torch.stack([requests[i][0] for i in chunk])
            init = np.asarray([tasks[i].initial_tokens for i in chunk], dtype=np.int32)
            for i, r in zip(chunk, tasks[chunk[0]].run(mel, initial_tokens=init)):
                out[i] = r
    return out
