"""Word-timing math with the reference's surface (whisper/timing.py:19-151): `median_filter` and
`dtw`, both on the GPU through the C ABI (csrc/timing.cu).  There is no Triton, no numba and no
CPU fallback: the DTW backtrace also runs on the device, only the final path (a few hundred int32)
comes back to the host."""
from __future__ import annotations

from ctypes import c_int, c_int64, c_size_t

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr


def _cuda(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("whisper_b200.timing needs a CUDA device (no CPU path)")
        x = x.cuda()
    return x


def median_filter(x: torch.Tensor, filter_width: int) -> torch.Tensor:
    """Median filter of width `filter_width` along the last dim with reflect padding (timing.py:19-54)."""
    pad_width = filter_width // 2
    if x.shape[-1] <= pad_width:
        return x                                                    # timing.py:22-24
    assert filter_width > 0 and filter_width % 2 == 1, "`filter_width` should be an odd number"
    xc = _cuda(x).to(torch.float32).contiguous()
    y = torch.empty_like(xc)
    T = xc.shape[-1]
    rows = xc.numel() // T
    with torch.cuda.device(xc.device):
        check(lib().wb200_median_filter(ptr(xc), ptr(y), c_int64(rows), c_int(T), c_int(filter_width), stream_ptr()),
              "wb200_median_filter")
    return y.to(x.dtype) if x.dtype != torch.float32 else y


def dtw(x: torch.Tensor, cpu_tie_break: bool = False) -> np.ndarray:
    """Dynamic time warping over cost matrix x (N text tokens, M frames) -> int array (2, path_len)
    (timing.py:82-151).  Ties follow the rule the reference applies to CUDA tensors
    (triton_ops.py:38-40) unless `cpu_tie_break` asks for dtw_cpu's (timing.py:95-100)."""
    if not torch.is_tensor(x):
        x = torch.from_numpy(np.ascontiguousarray(x))
    xc = _cuda(x).to(torch.float32).contiguous()
    N, M = xc.shape
    path = torch.empty((2, N + M + 1), device=xc.device, dtype=torch.int32)
    n = torch.zeros(1, device=xc.device, dtype=torch.int32)
    with torch.cuda.device(xc.device):
        nbytes = int(lib().wb200_dtw_workspace_bytes(c_int(N), c_int(M)))
        ws = torch.empty(nbytes, device=xc.device, dtype=torch.uint8)
        check(lib().wb200_dtw(ptr(xc), c_int(N), c_int(M), ptr(path), ptr(n), ptr(ws), c_size_t(nbytes),
                              c_int(int(cpu_tie_break)), stream_ptr()), "wb200_dtw")
    length = int(n.item())
    return path[:, :length].cpu().numpy().astype(np.int64)


# ------------------------------------------------------------------------------------------------
# token <-> audio-frame alignment (reference timing.py:154-242).  The string heuristics that turn an
# alignment into per-segment word lists (merge_punctuations / add_word_timestamps, timing.py:245-388)
# are host-side text post-processing that SURVEY.md section 2 leaves out of scope; they are not rebuilt.
# ------------------------------------------------------------------------------------------------
import ctypes
from dataclasses import dataclass as _dataclass
from typing import TYPE_CHECKING as _TC, List as _List

from .audio import HOP_LENGTH, SAMPLE_RATE, TOKENS_PER_SECOND

if _TC:
    from .model import Whisper
    from .tokenizer import Tokenizer


@_dataclass
class WordTiming:
    word: str
    tokens: _List[int]
    start: float
    end: float
    probability: float


def alignment_matrix(qk: torch.Tensor, n_frames: int, qk_scale: float = 1.0, medfilt_width: int = 7,
                     negate: bool = False) -> torch.Tensor:
    """timing.py:207-214 on exported cross-attention scores qk [heads, tokens, 1500] (fp32, CUDA):
    softmax over the first n_frames frames, z-score over tokens, median filter, mean over heads ->
    [tokens, n_frames]."""
    assert qk.is_cuda and qk.dtype == torch.float32 and qk.is_contiguous() and qk.dim() == 3
    H, N, T = qk.shape
    out = torch.empty((N, n_frames), device=qk.device, dtype=torch.float32)
    scratch = torch.empty(2 * H * N * n_frames, device=qk.device, dtype=torch.float32)
    from ctypes import c_float

    with torch.cuda.device(qk.device):
        check(lib().wb200_alignment_weights(ptr(qk), c_int(H), c_int(N), c_int(T), c_int(n_frames), c_float(qk_scale),
                                            c_int(medfilt_width), c_int(int(negate)), ptr(out), ptr(scratch),
                                            stream_ptr()), "wb200_alignment_weights")
    return out


def find_alignment(model: "Whisper", tokenizer: "Tokenizer", text_tokens: _List[int], mel: torch.Tensor,
                   num_frames: int, *, medfilt_width: int = 7, qk_scale: float = 1.0,
                   audio_features: torch.Tensor = None) -> _List[WordTiming]:
    """Reference timing.py:163-242.  The reference re-runs encoder AND decoder with SDPA disabled to
    capture every cross-attention matrix through hooks; here one teacher-forced decoder pass exports
    the scores of the alignment heads only, and `audio_features` (e.g. DecodingResult.audio_features)
    can be passed to skip the encoder."""
    if len(text_tokens) == 0:
        return []
    tokens = [*tokenizer.sot_sequence, tokenizer.no_timestamps, *text_tokens, tokenizer.eot]
    if audio_features is None:
        audio_features = model.embed_audio(mel)
    heads = model.alignment_heads.indices().T.tolist()                     # (layer, head) pairs, timing.py:207
    logits, qk = model.logits(torch.tensor([tokens]), audio_features, alignment_heads=heads)
    # probability of every text token under the softmax over the text vocabulary [0, eot) (timing.py:198-203)
    n_sot = len(tokenizer.sot_sequence)
    rows = logits[0, n_sot: n_sot + len(text_tokens)]
    assert rows.stride(1) == 1
    tok_dev = torch.tensor(text_tokens, device=rows.device, dtype=torch.int32)
    tok_prob = torch.empty((len(text_tokens),), device=rows.device, dtype=torch.float32)
    with torch.cuda.device(rows.device):
        check(lib().wb200_range_softmax(ptr(rows), ctypes.c_int64(rows.stride(0)), c_int(0), c_int(tokenizer.eot),
                                        c_int(len(text_tokens)), None, None, ptr(tok_dev), ptr(tok_prob), stream_ptr()),
              "wb200_range_softmax")
    text_token_probs = tok_prob.cpu().numpy()

    matrix = alignment_matrix(qk, num_frames // 2, qk_scale, medfilt_width, negate=True)   # already -matrix
    matrix = matrix[n_sot: -1].contiguous()                                # timing.py:215
    text_indices, time_indices = dtw(matrix)                               # timing.py:216

    words, word_tokens = tokenizer.split_to_word_tokens(text_tokens + [tokenizer.eot])
    if len(word_tokens) <= 1:
        return []                                                          # only <|endoftext|>: timing.py:219-225
    return word_spans(words, word_tokens, text_indices, time_indices, text_token_probs)


def word_spans(words, word_tokens, text_indices, time_indices, token_probs) -> _List[WordTiming]:
    """Start / end time and mean token probability of every word from a DTW path (timing.py:227-241).

    Word k owns the token positions [first_tok[k], first_tok[k + 1]); the monotone path reaches token position t for the
    first time at some path index, i.e. at frame time_indices[that index].  A word starts when the path reaches its first
    token and ends when it reaches the next word's first token; the trailing <|endoftext|> "word" only closes the last
    real one and is not returned."""
    first_tok = np.concatenate([[0], np.cumsum([len(t) for t in word_tokens[:-1]])]).astype(np.int64)
    text_indices = np.asarray(text_indices)
    advanced = np.concatenate([[True], text_indices[1:] != text_indices[:-1]])
    reach_time = np.asarray(time_indices)[advanced] / TOKENS_PER_SECOND
    timings = []
    for k in range(len(word_tokens) - 1):
        lo, hi = int(first_tok[k]), int(first_tok[k + 1])
        timings.append(WordTiming(words[k], word_tokens[k], reach_time[lo], reach_time[hi],
                                  float(np.mean(token_probs[lo:hi]))))
    return timings
