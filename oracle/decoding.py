"""Oracle for segment decoding (reference whisper/decoding.py).  TEST INFRASTRUCTURE ONLY.

Token selection is integer / fp32-scalar work: plain Python lists and dicts plus torch CPU fp32 for
log-softmax, exactly the arithmetic the reference does (its beam search is itself pure Python over
`.item()` floats).  The batched beam-search semantics - which the reference cannot execute for
n_audio > 1 (decoding.py:734,740 raise) - are defined as "independent beam search per audio";
`decode()` below runs all audios in one batch and tests check it equals the reference looped over
audios.
"""
from __future__ import annotations

import json
import os
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import model as M

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
NEG_INF = float("-inf")


# ------------------------------------------------------------------------------------------------
# token ids (integers only; the BPE itself is out of scope - SURVEY.md section 2)
# ------------------------------------------------------------------------------------------------
@dataclass
class TokenIds:
    """Special-token ids of tokenizer.py:340-355 for a given vocabulary size."""
    n_vocab: int
    eot: int
    sot: int
    translate: int
    transcribe: int
    sot_lm: int
    sot_prev: int
    no_speech: int
    no_timestamps: int
    timestamp_begin: int
    num_languages: int
    multilingual: bool
    non_speech: Tuple[int, ...]          # tokenizer.py:241-276, from the golden id table
    blank: Tuple[int, ...]               # encode(" ") (decoding.py:430), from the golden id table
    language_codes: Tuple[str, ...] = ()

    def sot_sequence(self, language: Optional[str] = "en", task: str = "transcribe") -> Tuple[int, ...]:
        if not self.multilingual:
            return (self.sot,)                                        # tokenizer.py:381-384
        lang = language or "en"
        return (self.sot, self.sot + 1 + self.language_codes.index(lang),
                self.transcribe if task == "transcribe" else self.translate)  # tokenizer.py:171-181

    @property
    def all_language_tokens(self) -> Tuple[int, ...]:
        return tuple(range(self.sot + 1, self.sot + 1 + self.num_languages))


def token_ids(n_vocab: int) -> TokenIds:
    with open(os.path.join(_GOLDEN, "token_ids.json")) as f:
        table = json.load(f)
    multilingual = n_vocab >= 51865                                    # model.py:302-304
    num_languages = n_vocab - 51765 - int(multilingual)                # model.py:306-308
    base = 50257 if multilingual else 50256                            # size of the BPE rank table
    eot, sot = base, base + 1
    t = sot + 1 + num_languages
    key = "multilingual" if multilingual else "gpt2"
    return TokenIds(n_vocab=n_vocab, eot=eot, sot=sot, translate=t, transcribe=t + 1, sot_lm=t + 2,
                    sot_prev=t + 3, no_speech=t + 4, no_timestamps=t + 5, timestamp_begin=t + 6,
                    num_languages=num_languages, multilingual=multilingual,
                    non_speech=tuple(table[key]["non_speech_tokens"]),
                    blank=tuple(table[key]["blank"]),
                    language_codes=tuple(table["languages"][:num_languages]))


# ------------------------------------------------------------------------------------------------
# options / result (decoding.py:80-127)
# ------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Options:
    task: str = "transcribe"
    language: Optional[str] = "en"
    temperature: float = 0.0
    sample_len: Optional[int] = None
    beam_size: Optional[int] = None
    patience: Optional[float] = None
    length_penalty: Optional[float] = None
    prompt: Optional[List[int]] = None
    prefix: Optional[List[int]] = None
    suppress_tokens: Optional[Sequence[int]] = (-1,)
    suppress_blank: bool = True
    without_timestamps: bool = False
    max_initial_timestamp: Optional[float] = 1.0
    best_of: Optional[int] = None       # independent samples per audio when temperature > 0 (decoding.py:524-526)
    seed: int = 0                       # key of the counter-based sampler (sample_update)


@dataclass
class Result:
    tokens: List[int] = field(default_factory=list)
    avg_logprob: float = float("nan")
    no_speech_prob: float = float("nan")
    sum_logprob: float = float("nan")
    audio_features: Optional[torch.Tensor] = None
    language: Optional[str] = None
    top_language_prob: float = float("nan")
    # diagnostics for margin-gated comparisons
    step_margins: List[float] = field(default_factory=list)


# ------------------------------------------------------------------------------------------------
# logit filters (decoding.py:423-505); logits: fp32 (R, V) modified in place, tokens: list of lists
# ------------------------------------------------------------------------------------------------
def suppress_blank(logits: torch.Tensor, tokens: List[List[int]], ids: TokenIds, sample_begin: int) -> None:
    if len(tokens[0]) == sample_begin:                                  # decoding.py:429
        logits[:, list(ids.blank) + [ids.eot]] = NEG_INF


def suppress_tokens(logits: torch.Tensor, suppress: Sequence[int]) -> None:
    logits[:, list(suppress)] = NEG_INF                                 # decoding.py:438


def timestamp_rules(logits: torch.Tensor, tokens: List[List[int]], ids: TokenIds, sample_begin: int,
                    max_initial_timestamp_index: Optional[int], rule_gaps: Optional[List[float]] = None) -> None:
    """ApplyTimestampRules (decoding.py:441-505).  rule_gaps (diagnostics for the margin-gated parity tests): per row, how
    far the "timestamps outweigh every text token" decision of :498-505 was from flipping, in log-probability units."""
    tb = ids.timestamp_begin
    logits[:, ids.no_timestamps] = NEG_INF                              # decoding.py:454-455
    for k, row in enumerate(tokens):
        seq = row[sample_begin:]
        last_ts = len(seq) >= 1 and seq[-1] >= tb                       # decoding.py:461-463
        penult_ts = len(seq) < 2 or seq[-2] >= tb                       # decoding.py:464-466
        if last_ts:
            if penult_ts:
                logits[k, tb:] = NEG_INF                                # has to be non-timestamp
            else:
                logits[k, : ids.eot] = NEG_INF                          # cannot be normal text
        stamps = [t for t in seq if t >= tb]                            # decoding.py:474-476
        if stamps:
            floor = stamps[-1] if (last_ts and not penult_ts) else stamps[-1] + 1   # :480-483
            logits[k, tb:floor] = NEG_INF
    if len(tokens[0]) == sample_begin:                                  # decoding.py:486-495
        logits[:, :tb] = NEG_INF
        if max_initial_timestamp_index is not None:
            logits[:, tb + max_initial_timestamp_index + 1:] = NEG_INF
    logprobs = torch.log_softmax(logits.float(), dim=-1)                # decoding.py:498-505
    for k in range(len(tokens)):
        ts_lp = torch.logsumexp(logprobs[k, tb:], dim=-1)
        text_max = logprobs[k, :tb].max()
        if rule_gaps is not None:
            both = bool(torch.isfinite(ts_lp)) and bool(torch.isfinite(text_max))
            rule_gaps.append(abs(float(ts_lp - text_max)) if both else float("inf"))
        if ts_lp > text_max:
            logits[k, :tb] = NEG_INF


# ------------------------------------------------------------------------------------------------
# token selection (decoding.py:272-404)
# ------------------------------------------------------------------------------------------------
def greedy_update(tokens: List[List[int]], logits: torch.Tensor, sum_logprobs: torch.Tensor,
                  eot: int) -> Tuple[List[List[int]], bool]:
    """GreedyDecoder.update with temperature 0 (decoding.py:277-293)."""
    nxt = logits.argmax(dim=-1)
    logprobs = torch.log_softmax(logits.float(), dim=-1)
    cur = logprobs[torch.arange(len(tokens), device=logits.device), nxt]
    last = torch.tensor([t[-1] for t in tokens], device=logits.device)
    sum_logprobs += cur * (last != eot)
    nxt = torch.where(last == eot, torch.full_like(nxt, eot), nxt)
    out = [t + [int(n)] for t, n in zip(tokens, nxt)]
    return out, all(t[-1] == eot for t in out)


# ---- temperature sampling (decoding.py:283: Categorical(logits=logits / temperature).sample()) -------------
# The reference draws from torch's global generator, which cannot be replayed elsewhere.  What
# CAN be pinned is (1) the distribution - a Gumbel-max draw is an exact Categorical(softmax(logits / T)) sample -
# and (2) a counter-based generator so that (seed, row, step) determines the draw on any implementation.
# include/whisper_b200.h (wb200_decoder_set_sampling) states the same contract.
_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter: np.ndarray, key: Tuple[int, int]) -> np.ndarray:
    """Philox4x32-10 (Salmon et al., SC'11; Random123 reference) on an array of counters [..., 4] uint32."""
    c = np.asarray(counter, dtype=np.uint64).copy()
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _PHILOX_M0 * c[..., 0]
        p1 = _PHILOX_M1 * c[..., 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        n0 = hi1 ^ c[..., 1] ^ np.uint64(k0)
        n2 = hi0 ^ c[..., 3] ^ np.uint64(k1)
        c = np.stack([n0, lo1, n2, lo0], axis=-1)
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def gumbel_noise(seed: int, row: int, step: int, n_vocab: int) -> np.ndarray:
    """g_v for v < n_vocab: counter (v >> 2, row, step, 0), key (seed & 0xffffffff, seed >> 32), word v & 3,
    u = ((word >> 8) + 0.5) * 2^-24, g = -log(-log(u)) in fp32."""
    n4 = (n_vocab + 3) // 4
    ctr = np.zeros((n4, 4), dtype=np.uint64)
    ctr[:, 0] = np.arange(n4, dtype=np.uint64)
    ctr[:, 1] = row
    ctr[:, 2] = step
    words = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).reshape(-1)[:n_vocab]
    u = ((words >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
    return -np.log(-np.log(u, dtype=np.float32), dtype=np.float32)


def sample_update(tokens: List[List[int]], logits: torch.Tensor, sum_logprobs: torch.Tensor, eot: int,
                  temperature: float, seed: int) -> Tuple[List[List[int]], bool, List[float]]:
    """GreedyDecoder.update with temperature > 0 (decoding.py:277-293) under the Gumbel-max contract.
    Also returns, per row, the gap between the best and second-best perturbed score (for margin-gated parity)."""
    V = logits.shape[-1]
    inv_t = np.float32(1.0 / np.float32(temperature))
    nxt, gaps = [], []
    for r, row in enumerate(logits.float().numpy()):
        key = np.where(np.isneginf(row), -np.inf, row * inv_t + gumbel_noise(seed, r, len(tokens[r]), V)).astype(np.float32)
        order = np.argsort(-key, kind="stable")            # ties -> lower id
        nxt.append(int(order[0]))
        gaps.append(float(key[order[0]] - key[order[1]]) if V > 1 else float("inf"))
    nxt = torch.tensor(nxt)
    logprobs = torch.log_softmax(logits.float(), dim=-1)   # un-tempered (decoding.py:285)
    cur = logprobs[torch.arange(len(tokens)), nxt]
    last = torch.tensor([t[-1] for t in tokens])
    sum_logprobs += cur * (last != eot)
    nxt = torch.where(last == eot, torch.full_like(nxt, eot), nxt)
    out = [t + [int(n)] for t, n in zip(tokens, nxt)]
    return out, all(t[-1] == eot for t in out), gaps


class BeamState:
    """BeamSearchDecoder (decoding.py:301-404) with its per-audio `finished_sequences` dicts."""

    def __init__(self, beam_size: int, eot: int, patience: Optional[float]):
        self.beam = beam_size
        self.eot = eot
        self.max_candidates = round(beam_size * (patience or 1.0))          # decoding.py:312-313
        assert self.max_candidates > 0
        self.finished: Optional[List[Dict[tuple, float]]] = None
        self.min_gap = float("inf")     # diagnostics: smallest score gap that could reorder candidates
        self.step_gaps: List[float] = []   # the same, per update() call (margin-gated step-wise parity)

    def update(self, tokens: List[List[int]], logits: torch.Tensor, sum_logprobs: torch.Tensor):
        G = self.beam
        assert len(tokens) % G == 0
        n_audio = len(tokens) // G
        if self.finished is None:
            self.finished = [dict() for _ in range(n_audio)]
        logprobs = torch.log_softmax(logits.float(), dim=-1)
        new_tokens: List[List[int]] = []
        sources: List[int] = []
        newly: List[Dict[tuple, float]] = []
        gap_before = self.min_gap
        self.min_gap = float("inf")
        for a in range(n_audio):
            score: Dict[tuple, float] = {}
            origin: Dict[tuple, int] = {}
            for j in range(G):                                               # decoding.py:339-346
                r = a * G + j
                vals, idxs = logprobs[r].topk(G + 2)
                if torch.isfinite(vals[G + 1]):
                    self.min_gap = min(self.min_gap, float(vals[G] - vals[G + 1]))   # top-(G+1) boundary
                vals, idxs = vals[: G + 1], idxs[: G + 1]
                for lp, tok in zip(vals, idxs):
                    seq = tuple(tokens[r] + [int(tok)])
                    score[seq] = float((sum_logprobs[r] + lp).item())        # fp32 add, then widen
                    origin[seq] = r
            done: Dict[tuple, float] = {}
            kept = 0
            ranked = sorted(score.values(), reverse=True)[: 2 * G]
            for hi, lo in zip(ranked, ranked[1:]):
                if np.isfinite(lo):
                    self.min_gap = min(self.min_gap, hi - lo)
            for seq in sorted(score, key=score.get, reverse=True):           # stable; decoding.py:350-360
                if seq[-1] == self.eot:
                    done[seq] = score[seq]
                else:
                    sum_logprobs[len(new_tokens)] = score[seq]
                    new_tokens.append(list(seq))
                    sources.append(origin[seq])
                    kept += 1
                    if kept == G:
                        break
            newly.append(done)
        for prev, new in zip(self.finished, newly):                          # decoding.py:367-375
            for seq in sorted(new, key=new.get, reverse=True):
                if len(prev) >= self.max_candidates:
                    break
                prev[seq] = new[seq]
        completed = all(len(s) >= self.max_candidates for s in self.finished)
        self.step_gaps.append(self.min_gap)
        self.min_gap = min(self.min_gap, gap_before)
        return new_tokens, sources, completed

    def finalize(self, tokens: List[List[List[int]]], sum_logprobs: torch.Tensor):
        """decoding.py:384-404; tokens indexed [audio][beam]."""
        for a, seqs in enumerate(self.finished):
            if len(seqs) < self.beam:
                for j in list(np.argsort(sum_logprobs[a].cpu().numpy()))[::-1]:
                    seqs[tuple(tokens[a][j] + [self.eot])] = float(sum_logprobs[a][j])
                    if len(seqs) >= self.beam:
                        break
        return ([[list(s) for s in seqs.keys()] for seqs in self.finished],
                [list(seqs.values()) for seqs in self.finished])


def rank(candidates: List[List[List[int]]], sum_logprobs: List[List[float]],
         length_penalty: Optional[float]) -> List[int]:
    """MaximumLikelihoodRanker.rank (decoding.py:199-213)."""
    picks = []
    for seqs, lps in zip(candidates, sum_logprobs):
        scores = []
        for s, lp in zip(seqs, lps):
            n = len(s)
            pen = n if length_penalty is None else ((5 + n) / 6) ** length_penalty
            scores.append(lp / pen)
        picks.append(int(np.argmax(scores)))
    return picks


# ------------------------------------------------------------------------------------------------
# the task (decoding.py:508-789)
# ------------------------------------------------------------------------------------------------
def initial_tokens(ids: TokenIds, opt: Options, n_ctx: int, sample_len: int) -> Tuple[int, ...]:
    """decoding.py:587-613 (token-list prompts / prefixes only: string encoding is BPE, out of scope)."""
    seq = list(ids.sot_sequence(opt.language, opt.task))
    if opt.without_timestamps:
        seq.append(ids.no_timestamps)                                         # decoding.py:532-533
    if opt.prefix:
        pre = list(opt.prefix)
        keep = n_ctx // 2 - sample_len
        pre = pre[-keep:]                         # decoding.py:597-598 (note: -0 keeps everything)
        seq = seq + pre
    if opt.prompt:
        seq = [ids.sot_prev] + list(opt.prompt)[-(n_ctx // 2 - 1):] + seq
    return tuple(seq)


def suppress_list(ids: TokenIds, opt: Options) -> Tuple[int, ...]:
    """decoding.py:615-642."""
    sup = list(opt.suppress_tokens) if opt.suppress_tokens is not None else []
    if -1 in sup:
        sup = [t for t in sup if t >= 0]
        sup.extend(ids.non_speech)
    sup.extend([ids.transcribe, ids.translate, ids.sot, ids.sot_prev, ids.sot_lm, ids.no_speech])
    return tuple(sorted(set(sup)))


def apply_filters(logits: torch.Tensor, tokens: List[List[int]], ids: TokenIds, opt: Options, sample_begin: int,
                  sup: Sequence[int], mits: Optional[int], rule_gaps: Optional[List[float]] = None) -> None:
    """The LogitFilter chain in the order DecodingTask builds it (decoding.py:554-570), in place."""
    if opt.suppress_blank:
        suppress_blank(logits, tokens, ids, sample_begin)
    if sup:
        suppress_tokens(logits, sup)
    if not opt.without_timestamps:
        timestamp_rules(logits, tokens, ids, sample_begin, mits, rule_gaps)


def filter_context(dims: Dict[str, int], opt: Options):
    """(ids, initial tokens, sample_begin, suppress list, max_initial_timestamp_index) of a task - what
    apply_filters needs (decoding.py:521-570)."""
    ids = token_ids(dims["n_vocab"])
    n_ctx = dims["n_text_ctx"]
    init = initial_tokens(ids, opt, n_ctx, opt.sample_len or n_ctx // 2)
    sup = suppress_list(ids, opt) if opt.suppress_tokens else ()
    mits = None
    if not opt.without_timestamps and opt.max_initial_timestamp:
        mits = round(opt.max_initial_timestamp / (30.0 / dims["n_audio_ctx"]))
    return ids, init, len(init), sup, mits


def decode(W: M.Weights, dims: Dict[str, int], mel_or_features: torch.Tensor, opt: Options = Options(),
           record: Optional[dict] = None, max_steps: Optional[int] = None,
           timings: Optional[list] = None) -> List[Result]:
    """DecodingTask.run (decoding.py:713-789) for a batch; beam search is per audio.
    `timings`, if given, receives the wall-clock seconds of every loop iteration (bench.py's CPU baseline)."""
    import time as _time
    ids = token_ids(dims["n_vocab"])
    n_ctx = dims["n_text_ctx"]
    G = opt.beam_size or opt.best_of or 1                                    # decoding.py:524
    sample_len = opt.sample_len or n_ctx // 2
    init = initial_tokens(ids, opt, n_ctx, sample_len)
    sample_begin = len(init)
    sot_index = init.index(ids.sot)
    x = mel_or_features if mel_or_features.is_cuda else mel_or_features.float()     # CUDA: keep the caller's 16-bit type
    if x.dim() == 2:
        x = x[None]
    if tuple(x.shape[-2:]) == (dims["n_audio_ctx"], dims["n_audio_state"]):   # decoding.py:648-653
        feats = x
    else:
        feats = M.encoder_forward(W, dims, x)
    B = feats.shape[0]
    R = B * G
    tokens: List[List[int]] = [list(init) for _ in range(R)]
    languages: List[Optional[str]] = [opt.language] * B
    top_lang_prob = [float("nan")] * B
    if opt.language is None or opt.task == "lang_id":                         # decoding.py:666-678
        lang_tokens, lang_probs = detect_language(W, dims, feats)
        languages = [ids.language_codes[int(t) - ids.sot - 1] for t in lang_tokens]
        top_lang_prob = lang_probs.max(dim=-1).values.tolist()
        if opt.language is None:
            for r in range(R):
                tokens[r][sot_index + 1] = int(lang_tokens[r // G])            # write the language token
    if opt.task == "lang_id":                                                 # decoding.py:722-727
        return [Result(language=languages[a], top_language_prob=top_lang_prob[a], audio_features=feats[a]) for a in range(B)]
    xa = feats.repeat_interleave(G, dim=0) if G > 1 else feats
    cache = M.KVCache(dims["n_text_layer"])
    sum_lp = torch.zeros(R, device=feats.device)
    no_speech = [float("nan")] * R
    beam = BeamState(G, ids.eot, opt.patience) if opt.beam_size is not None else None
    sup = suppress_list(ids, opt) if opt.suppress_tokens else ()
    mits = None
    if not opt.without_timestamps and opt.max_initial_timestamp:
        mits = round(opt.max_initial_timestamp / (30.0 / dims["n_audio_ctx"]))  # decoding.py:560-565
    margins: List[List[float]] = [[] for _ in range(R)]
    steps = sample_len if max_steps is None else min(sample_len, max_steps)
    for i in range(steps):                                                   # decoding.py:686
        _t0 = _time.perf_counter()
        new = torch.tensor([t[cache.length:] for t in tokens], device=feats.device)   # decoding.py:159-161
        logits_all = M.decoder_forward(W, dims, new, xa, cache)
        if i == 0:                                                           # decoding.py:689-693
            probs = torch.softmax(logits_all[:, sot_index].float(), dim=-1)
            no_speech = probs[:, ids.no_speech].tolist()
        logits = logits_all[:, -1].clone()
        if record is not None:
            record.setdefault("raw_logits", []).append(logits.clone())
            record.setdefault("tokens_in", []).append([list(t) for t in tokens])
        rule_gaps: Optional[List[float]] = [] if record is not None else None
        apply_filters(logits, tokens, ids, opt, sample_begin, sup, mits, rule_gaps)
        if record is not None:
            record.setdefault("filtered_logits", []).append(logits.clone())
            record.setdefault("sum_logprobs_in", []).append(sum_lp.clone())
        if record is not None:                 # decision margins: diagnostics for the margin-gated parity tests only
            # the smaller of (a) the gap between the two best surviving logits and (b) the distance of the timestamp-vs-text
            # rule (decoding.py:498-505) from flipping: a near-tie there removes or keeps the whole text vocabulary
            top2 = logits.topk(2, dim=-1).values.cpu()
            for r in range(R):
                gap = float(top2[r, 0] - top2[r, 1])
                if rule_gaps:
                    gap = min(gap, rule_gaps[r])
                margins[r].append(gap)
            record.setdefault("rule_gaps", []).append(list(rule_gaps) if rule_gaps else None)
        if beam is None and opt.temperature > 0:
            tokens, completed, gaps = sample_update(tokens, logits, sum_lp, ids.eot, opt.temperature, opt.seed)
            if record is not None:
                record.setdefault("sample_gaps", []).append(gaps)
        elif beam is None:
            tokens, completed = greedy_update(tokens, logits, sum_lp, ids.eot)
        else:
            tokens, src, completed = beam.update(tokens, logits, sum_lp)
            cache.reorder(src)
            if record is not None:
                record.setdefault("source_indices", []).append(list(src))
                if rule_gaps:              # a flipped timestamp rule changes a beam's whole candidate set
                    g_rule = min(rule_gaps)
                    beam.step_gaps[-1] = min(beam.step_gaps[-1], g_rule)
                    beam.min_gap = min(beam.min_gap, g_rule)
        if record is not None:
            record.setdefault("tokens_out", []).append([list(t) for t in tokens])
            record.setdefault("sum_logprobs_out", []).append(sum_lp.clone())
        if timings is not None:
            timings.append(_time.perf_counter() - _t0)
        if completed or len(tokens[0]) > n_ctx:                              # decoding.py:705
            break
    grouped = [[tokens[a * G + j] for j in range(G)] for a in range(B)]
    lp_grouped = sum_lp.reshape(B, G).cpu()
    if beam is None:
        cands = [[t + [ids.eot] for t in grp] for grp in grouped]            # decoding.py:295-298
        cand_lp = lp_grouped.tolist()
    else:
        cands, cand_lp = beam.finalize(grouped, lp_grouped)
    cands = [[s[sample_begin: s.index(ids.eot)] for s in grp] for grp in cands]   # decoding.py:749-752
    pick = rank(cands, cand_lp, opt.length_penalty)
    out = []
    for a in range(B):
        toks = cands[a][pick[a]]
        lp = cand_lp[a][pick[a]]
        out.append(Result(tokens=toks, avg_logprob=lp / (len(toks) + 1), sum_logprob=lp,
                          no_speech_prob=no_speech[a * G], audio_features=feats[a],
                          step_margins=margins[a * G], language=languages[a], top_language_prob=top_lang_prob[a]))
    if record is not None:
        record["all_margins"] = margins
        record["beam_min_gap"] = beam.min_gap if beam is not None else None
        record["beam_step_gaps"] = list(beam.step_gaps) if beam is not None else None
    return out


def detect_language(W: M.Weights, dims: Dict[str, int], feats: torch.Tensor):
    """decoding.py:19-77 on pre-encoded features: one [sot] step, mask non-language ids."""
    ids = token_ids(dims["n_vocab"])
    B = feats.shape[0]
    logits = M.decoder_forward(W, dims, torch.full((B, 1), ids.sot), feats)[:, 0]
    mask = torch.ones(logits.shape[-1], dtype=torch.bool)
    mask[list(ids.all_language_tokens)] = False
    logits[:, mask] = NEG_INF
    lang_tokens = logits.argmax(dim=-1)
    probs = logits.softmax(dim=-1)
    return lang_tokens, probs[:, list(ids.all_language_tokens)]


def compression_ratio(text: str) -> float:
    """utils.py:45-47."""
    b = text.encode("utf-8")
    return len(b) / len(zlib.compress(b))
