"""Oracle for the word-timing math (reference whisper/timing.py:19-151).  TEST INFRASTRUCTURE ONLY.

numpy only.  `dtw` follows the CPU implementation's tie-breaking (strict `<`, timing.py:95-100);
`dtw_gpu_tiebreak` follows the Triton kernel's (`<=` with diagonal written last,
triton_ops.py:38-40).  The two differ only on exact cost ties.
"""
from __future__ import annotations

import numpy as np


def median_filter(x: np.ndarray, filter_width: int) -> np.ndarray:
    """timing.py:19-54: reflect-pad by width//2 along the last axis, sliding median."""
    pad = filter_width // 2
    if x.shape[-1] <= pad:                                    # timing.py:22-24
        return x
    assert filter_width > 0 and filter_width % 2 == 1, "`filter_width` should be an odd number"
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(pad, pad)], mode="reflect")
    win = np.lib.stride_tricks.sliding_window_view(xp, filter_width, axis=-1)
    return np.sort(win, axis=-1)[..., pad]                    # timing.py:49


def backtrace(trace: np.ndarray) -> np.ndarray:
    """timing.py:57-79."""
    i, j = trace.shape[0] - 1, trace.shape[1] - 1
    trace = trace.copy()
    trace[0, :] = 2
    trace[:, 0] = 1
    path = []
    while i > 0 or j > 0:
        path.append((i - 1, j - 1))
        t = trace[i, j]
        if t == 0:
            i, j = i - 1, j - 1
        elif t == 1:
            i -= 1
        elif t == 2:
            j -= 1
        else:
            raise ValueError("Unexpected trace[i, j]")
    return np.array(path)[::-1, :].T


def _dtw(x: np.ndarray, gpu_ties: bool) -> np.ndarray:
    N, M = x.shape
    x = x.astype(np.float32)
    cost = np.full((N + 1, M + 1), np.inf, dtype=np.float32)
    trace = -np.ones((N + 1, M + 1), dtype=np.int32)
    cost[0, 0] = 0
    for j in range(1, M + 1):                                 # timing.py:90-103
        for i in range(1, N + 1):
            c0, c1, c2 = cost[i - 1, j - 1], cost[i - 1, j], cost[i, j - 1]
            if gpu_ties:
                # triton_ops.py:38-40: stores 2, then 1, then 0, each under `<=` -> diag wins ties
                t = 2
                if c1 <= c0 and c1 <= c2:
                    t = 1
                if c0 <= c1 and c0 <= c2:
                    t = 0
                c = min(c0, c1, c2)
            else:
                if c0 < c1 and c0 < c2:
                    c, t = c0, 0
                elif c1 < c0 and c1 < c2:
                    c, t = c1, 1
                else:
                    c, t = c2, 2
            cost[i, j] = np.float32(x[i - 1, j - 1] + c)
            trace[i, j] = t
    return backtrace(trace)


def dtw(x: np.ndarray) -> np.ndarray:
    """dtw_cpu, timing.py:82-105: returns int array (2, path_len) of (text_index, time_index)."""
    return _dtw(x, gpu_ties=False)


def dtw_gpu_tiebreak(x: np.ndarray) -> np.ndarray:
    """dtw_cuda / dtw_kernel semantics (timing.py:108-138, triton_ops.py:13-40)."""
    return _dtw(x, gpu_ties=True)


def alignment_matrix(qk: np.ndarray, n_frames: int, qk_scale: float = 1.0, medfilt_width: int = 7) -> np.ndarray:
    """timing.py:207-214: qk [heads, tokens, 1500] (pre-softmax cross-attention scores of the alignment
    heads) -> softmax over the first n_frames frames, z-score over the token axis (population std),
    median filter along frames, mean over heads.  Returns [tokens, n_frames] fp32."""
    w = qk[:, :, :n_frames].astype(np.float32) * np.float32(qk_scale)
    w = w - w.max(axis=-1, keepdims=True)
    w = np.exp(w)
    w = w / w.sum(axis=-1, keepdims=True)
    mean = w.mean(axis=-2, keepdims=True)
    std = np.sqrt(((w - mean) ** 2).mean(axis=-2, keepdims=True))
    w = (w - mean) / std
    w = median_filter(w.astype(np.float32), medfilt_width)
    return w.mean(axis=0).astype(np.float32)
